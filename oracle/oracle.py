"""ctypes wrapper over the C oracle (oracle/spx_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports it.
"""
import ctypes as C

import numpy as np

from . import build as _build
from . import spec

MAX_A, MAX_CELLS, MAX_MOVES = 9, 42, 42

NET_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_float), C.POINTER(C.c_float))


class Cfg(C.Structure):
    _fields_ = [("game", C.c_int), ("sims", C.c_int), ("evaluate", C.c_int), ("strong_play", C.c_int),
                ("tie_mode", C.c_int), ("noise_mode", C.c_int), ("alpha", C.c_double),
                ("seed", C.c_uint64), ("game_uid", C.c_uint64),
                ("noise_table", C.POINTER(C.c_double)), ("table_moves", C.c_int), ("threads", C.c_int)]


class Record(C.Structure):
    _fields_ = [("tree", C.c_int), ("ply", C.c_int), ("state", C.c_int8 * MAX_CELLS),
                ("probs", C.c_float * MAX_A), ("q", C.c_float), ("actual_val", C.c_float)]


class Move(C.Structure):
    _fields_ = [("tree", C.c_int), ("ply", C.c_int), ("action", C.c_int), ("root_n", C.c_int),
                ("root_w", C.c_double), ("n", C.c_int * MAX_A), ("w", C.c_double * MAX_A)]


class Episode(C.Structure):
    _fields_ = [("reward", C.c_int), ("n_moves", C.c_int), ("n_records", C.c_int),
                ("sims", C.c_long), ("net_calls", C.c_long), ("path_len_sum", C.c_long),
                ("final_state", C.c_int8 * MAX_CELLS),
                ("moves", Move * (MAX_MOVES + 1)), ("records", Record * (MAX_MOVES + 1))]


class HashNetState(C.Structure):
    _fields_ = [("game", C.c_int), ("net_seed", C.c_uint64), ("calls", C.c_long)]


class ReplayState(C.Structure):
    _fields_ = [("game", C.c_int), ("A", C.c_int), ("n", C.c_long * 2), ("cursor", C.c_long * 2),
                ("mismatches", C.c_long), ("overruns", C.c_long),
                ("own", C.POINTER(C.c_uint64) * 2), ("opp", C.POINTER(C.c_uint64) * 2),
                ("policy", C.POINTER(C.c_float) * 2), ("value", C.POINTER(C.c_float) * 2)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        L.ox_sizeof.restype = C.c_size_t
        for i, st in enumerate((Cfg, Episode, Record, Move, HashNetState, ReplayState)):
            assert L.ox_sizeof(i) == C.sizeof(st), (st.__name__, L.ox_sizeof(i), C.sizeof(st))
        L.ox_rng_uniform.restype = C.c_double
        L.ox_rng_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64]
        L.ox_rng_u64.restype = C.c_uint64
        L.ox_rng_u64.argtypes = L.ox_rng_uniform.argtypes
        L.ox_pow_int_exact.restype = C.c_double
        L.ox_pow_int_exact.argtypes = [C.c_uint32, C.c_int]
        L.ox_tree_new.restype = C.c_void_p
        L.ox_tree_new.argtypes = [C.POINTER(Cfg), C.c_int, C.c_void_p, C.c_void_p]
        for f in ("ox_tree_free", "ox_tree_search"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        L.ox_tree_reset.argtypes = [C.c_void_p, C.c_int]
        L.ox_tree_play.argtypes = [C.c_void_p, C.POINTER(Move)]
        L.ox_tree_play_action.argtypes = [C.c_void_p, C.c_int]
        L.ox_tree_root_stats.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.ox_tree_counter.argtypes = [C.c_void_p, C.c_int]
        L.ox_tree_counter.restype = C.c_long
        L.ox_play_episode.argtypes = [C.POINTER(Cfg), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Episode)]
        L.ox_play_episode_vs.argtypes = [C.POINTER(Cfg), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Episode)]
        L.ox_hashnet_bits.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ox_set_live_counters.argtypes = [C.c_void_p]
        L.ox_env_playout.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
        _lib = L
    return _lib


def fn_addr(name):
    return C.cast(getattr(lib(), name), C.c_void_p)


# ----------------------------------------------------------------------------- env
def env_playout(game, actions, first_player=None):
    """actions int32 [n, max_plies] (-1 = no move).  Returns per-ply boards/reward/done/valid/status."""
    W, H, A = spec.GAME_DIMS[game]
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    n, T = actions.shape
    boards = np.zeros((n, T, W, H), np.int8)
    reward = np.zeros((n, T), np.int8)
    done = np.zeros((n, T), np.uint8)
    valid = np.zeros((n, T, A), np.uint8)
    status = np.zeros((n, T), np.int8)
    fp = None if first_player is None else np.ascontiguousarray(first_player, dtype=np.int8)
    lib().ox_env_playout(game, n, T, actions.ctypes.data, None if fp is None else fp.ctypes.data,
                         boards.ctypes.data, reward.ctypes.data, done.ctypes.data, valid.ctypes.data, status.ctypes.data)
    return dict(boards=boards, reward=reward, done=done, valid=valid, status=status)


# ----------------------------------------------------------------------------- nets
def hashnet_bits(own, opp, A, net_seed=0):
    p = np.zeros(A, np.float32)
    v = np.zeros(1, np.float32)
    lib().ox_hashnet_bits(int(own), int(opp), A, net_seed, p.ctypes.data, v.ctypes.data)
    return p, v[0]


def make_cfg(game, sims, seed=0, game_uid=0, evaluate=False, strong_play=False, tie_mode=1, noise_table=None,
             noise_mode=None, alpha=1.0, threads=1):
    cfg = Cfg()
    cfg.threads = threads
    cfg.game, cfg.sims, cfg.evaluate, cfg.strong_play = game, sims, int(evaluate), int(strong_play)
    cfg.tie_mode, cfg.alpha, cfg.seed, cfg.game_uid = tie_mode, alpha, seed, game_uid
    keep = None
    if noise_table is not None:
        keep = np.ascontiguousarray(noise_table, dtype=np.float64)  # [2][moves][A] or [1][moves][A]
        cfg.noise_table = keep.ctypes.data_as(C.POINTER(C.c_double))
        cfg.table_moves = keep.shape[1]
        cfg.noise_mode = 1
    else:
        cfg.noise_mode = 0
    if noise_mode is not None:
        cfg.noise_mode = noise_mode
    cfg._keep = keep
    return cfg


class PyNet:
    """Wraps a python callable (state int8[W,H] in net frame) -> (policy[A], value) as an ox_net_fn."""

    def __init__(self, game, fn):
        W, H, A = spec.GAME_DIMS[game]

        def _cb(user, tree, state, policy, value):
            s = np.ctypeslib.as_array(state, shape=(W * H,)).reshape(W, H)
            p, v = fn(s, tree)
            for i in range(A):
                policy[i] = float(p[i])
            value[0] = float(v)
        self.cb = NET_FN(_cb)
        self.addr = C.cast(self.cb, C.c_void_p)


class Tree:
    """One MCTreeSearch restatement instance (for search-level tests)."""

    def __init__(self, cfg, tree_id=0, net_addr=None, net_user=None, hash_seed=0):
        self.cfg = cfg
        self.A = spec.GAME_DIMS[cfg.game][2]
        if net_addr is None:
            self._hs = HashNetState(cfg.game, hash_seed, 0)
            net_addr, net_user = fn_addr("ox_hashnet"), C.addressof(self._hs)
        self._h = lib().ox_tree_new(C.byref(cfg), tree_id, net_addr, net_user)

    def reset(self, player=1):
        lib().ox_tree_reset(self._h, player)

    def search(self):
        lib().ox_tree_search(self._h)

    def play(self):
        mv = Move()
        a = lib().ox_tree_play(self._h, C.byref(mv))
        return a, mv

    def play_action(self, a):
        lib().ox_tree_play_action(self._h, a)

    def root_stats(self):
        n = np.zeros(self.A, np.int32)
        w = np.zeros(self.A, np.float64)
        valid = np.zeros(self.A, np.uint8)
        rn = np.zeros(1, np.int32)
        rw = np.zeros(1, np.float64)
        q = np.zeros(1, np.float64)
        pl = np.zeros(1, np.int32)
        lib().ox_tree_root_stats(self._h, n.ctypes.data, w.ctypes.data, valid.ctypes.data, rn.ctypes.data,
                                 rw.ctypes.data, q.ctypes.data, pl.ctypes.data)
        return dict(n=n, w=w, valid=valid.astype(bool), root_n=int(rn[0]), root_w=float(rw[0]), q=float(q[0]),
                    player=int(pl[0]))

    def counter(self, which):
        return lib().ox_tree_counter(self._h, which)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ox_tree_free(self._h)
            self._h = None


def episode_to_dict(ep, game):
    W, H, A = spec.GAME_DIMS[game]
    cells = W * H
    moves = []
    for i in range(ep.n_moves):
        m = ep.moves[i]
        moves.append(dict(tree=m.tree, ply=m.ply, action=m.action, root_n=m.root_n, root_w=m.root_w,
                          n=list(m.n[:A]), w=list(m.w[:A])))
    records = []
    for i in range(ep.n_records):
        r = ep.records[i]
        records.append(dict(tree=r.tree, ply=r.ply,
                            state=np.array(r.state[:cells], dtype=np.int8).reshape(W, H),
                            tree_probs=np.array(r.probs[:A], dtype=np.float32),
                            q=np.float32(r.q), actual_val=float(r.actual_val)))
    return dict(reward=ep.reward, moves=moves, records=records, sims=ep.sims, net_calls=ep.net_calls,
                path_len_sum=ep.path_len_sum,
                final_state=np.array(ep.final_state[:cells], dtype=np.int8).reshape(W, H))


def play_episode(cfg, swap_sides=False, net_seed=0, net_seed_opp=None, nets=None):
    """nets: optional ((addr0, user0), (addr1, user1)); default = hash nets."""
    ep = Episode()
    if nets is None:
        hs0 = HashNetState(cfg.game, net_seed, 0)
        hs1 = HashNetState(cfg.game, net_seed if net_seed_opp is None else net_seed_opp, 0)
        nets = ((fn_addr("ox_hashnet"), C.addressof(hs0)), (fn_addr("ox_hashnet"), C.addressof(hs1)))
    lib().ox_play_episode(C.byref(cfg), int(swap_sides), nets[0][0], nets[0][1], nets[1][0], nets[1][1], C.byref(ep))
    return episode_to_dict(ep, cfg.game)


def make_replay(game, logs):
    """logs: per tree dict(own u64[n], opp u64[n], policy f32[n,A], value f32[n]) -> (ReplayState, keepalive)."""
    A = spec.GAME_DIMS[game][2]
    rs = ReplayState()
    rs.game, rs.A = game, A
    keep = []
    for t in (0, 1):
        own = np.ascontiguousarray(logs[t]["own"], dtype=np.uint64)
        opp = np.ascontiguousarray(logs[t]["opp"], dtype=np.uint64)
        pol = np.ascontiguousarray(logs[t]["policy"], dtype=np.float32).reshape(-1, A)
        val = np.ascontiguousarray(logs[t]["value"], dtype=np.float32)
        keep += [own, opp, pol, val]
        rs.n[t] = len(own)
        rs.own[t] = own.ctypes.data_as(C.POINTER(C.c_uint64))
        rs.opp[t] = opp.ctypes.data_as(C.POINTER(C.c_uint64))
        rs.policy[t] = pol.ctypes.data_as(C.POINTER(C.c_float))
        rs.value[t] = val.ctypes.data_as(C.POINTER(C.c_float))
    rs._keep = keep
    return rs


def play_episode_vs(cfg, swap_sides, kind, net_seed=0, net=None):
    """Policy (MCTS, hash net unless ``net`` = (addr, user)) against a hard-coded opponent (spec.OPP_LOOKAHEAD / OPP_RANDOM)."""
    ep = Episode()
    if net is None:
        hs = HashNetState(cfg.game, net_seed, 0)
        net = (fn_addr("ox_hashnet"), C.addressof(hs))
    lib().ox_play_episode_vs(C.byref(cfg), int(swap_sides), kind, net[0], net[1], C.byref(ep))
    return episode_to_dict(ep, cfg.game)
