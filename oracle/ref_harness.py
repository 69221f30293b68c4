"""Runs the UNMODIFIED reference (/root/reference) under injected, reproducible inputs.

TEST INFRASTRUCTURE ONLY -- used by ``oracle/make_golden.py`` (to write ``tests/golden``)
and by the ``-m "not gpu"`` tests that pin the C oracle when /root/reference is present.
It cannot travel to the GPU box (the reference does not exist there); the golden fixtures do.

Hooks: the three numpy RNG call sites on the hot path
(mcts.py:50 dirichlet, :355 rand, :280 choice) are replaced by readers of the
``oracle/spec.py`` counter stream / injected noise tables; the network is ``HashNet``
(spec.hashnet) or a real module.  Objects are constructed directly
(``MCTreeSearch(network, env, ...)``, ``SelfPlayer(...)``) because the reference's own
entry scripts are stale.
"""
import os
import sys
import queue

import numpy as np

from . import spec

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "games"))


def _import_reference():
    shims = os.path.join(_HERE, "_shims")
    for p in (REFERENCE_ROOT, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    import games.algos.mcts as mcts  # noqa
    from games.algos.selfplayworker import SelfPlayer  # noqa
    from games.connect4.connect4env import Connect4Env  # noqa
    from games.tictactoe.tictactoe_env import TicTacToeEnv  # noqa
    return mcts, SelfPlayer, Connect4Env, TicTacToeEnv


class _Ctx:
    """Where in the (game, tree, ply, sim, depth) key space the reference currently is."""
    seed = 0
    game_uid = 0
    tree = 0
    ply = 0
    sim = -1
    depth = 0
    tie_mode = 1          # 0: zeros, 1: spec stream
    noise_table = None    # [tree][move_idx][A] float64 or None -> uniform 1/A
    moves_played = 0


CTX = _Ctx()


def _hook_rand(*shape):
    n = shape[0]
    if CTX.tie_mode == 0:
        out = np.zeros(n)
    else:
        out = np.array([spec.rng_uniform(CTX.seed, CTX.game_uid, CTX.tree, spec.PURPOSE_TIE,
                                         CTX.ply, CTX.sim, CTX.depth, a) for a in range(n)])
    CTX.depth += 1
    return out


def _hook_dirichlet(alpha):
    n = len(alpha)
    if CTX.noise_table is None:
        return np.ones(n) / n
    return np.array(CTX.noise_table[CTX.tree][CTX.moves_played], dtype=np.float64)


def _hook_choice(a, p=None):
    # numpy legacy RandomState.choice with p, replace=True, size=None (mtrand.pyx):
    # cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(random_sample(), side='right')
    p = np.asarray(p, dtype=np.float64)
    if np.isnan(p).any():
        raise ValueError("probabilities contain NaN")
    cdf = p.cumsum()
    cdf /= cdf[-1]
    u = spec.rng_uniform(CTX.seed, CTX.game_uid, CTX.tree, spec.PURPOSE_ACTION, CTX.ply, 0, 0, 0)
    return int(cdf.searchsorted(u, side="right"))


class hooked:
    """Context manager installing the RNG hooks on the numpy module the reference uses."""

    def __enter__(self):
        self._saved = (np.random.rand, np.random.dirichlet, np.random.choice)
        np.random.rand, np.random.dirichlet, np.random.choice = _hook_rand, _hook_dirichlet, _hook_choice
        return self

    def __exit__(self, *exc):
        np.random.rand, np.random.dirichlet, np.random.choice = self._saved


class HashNet:
    """spec.hashnet behind the reference's network call convention
    (general/modules.py:109-112: state*player in, value*player out, python lists/floats)."""

    def __init__(self, game, net_seed=0):
        self.game = game
        self.net_seed = net_seed
        self.A = spec.GAME_DIMS[game][2]
        self.calls = 0

    def to(self, *a, **k):
        return self

    def train(self, *a, **k):
        return self

    def __call__(self, s, player=1):
        self.calls += 1
        own, opp = spec.board_to_bits(np.asarray(s) * player, self.game)
        p, v = spec.hashnet(own, opp, self.A, self.net_seed)
        return [float(x) for x in p], float(v) * player


def _coop_search(mcts, tree, K):
    """The reference's threaded search (mcts.py:328-331: ThreadPoolExecutor(thread_count) runs `iterations` search_node tasks)
    under ONE forced, legal interleaving: K real threads run the UNMODIFIED ``MCTreeSearch.search_node``; exactly one of them
    runs at a time; a worker keeps taking tasks until it blocks inside ``self.network(...)`` (the InferenceProxy round trip),
    then the next worker runs; when all K have blocked (or run out of tasks) the answers come back and worker 0 resumes, ...
    Locks (``child_node.lock``, held across the network call) and virtual losses are the reference's own."""
    import threading
    N = tree.iterations
    state = {"next": 0, "cur": 0}
    turn = [threading.Semaphore(0) for _ in range(K)]
    back = threading.Semaphore(0)
    status = ["run"] * K
    real_net = tree.network
    errors = []

    class YieldingNet:
        def __call__(self, *a, **kw):
            out = real_net(*a, **kw)          # evaluated now, delivered when the worker is resumed
            k = state["cur"]
            status[k] = "blocked"
            back.release()
            turn[k].acquire()
            status[k] = "run"
            return out

    def worker(k):
        turn[k].acquire()
        try:
            while state["next"] < N:
                CTX.sim = state["next"]
                CTX.depth = 0
                state["next"] += 1
                mcts.MCTreeSearch.search_node(tree)
        except BaseException as e:       # noqa: BLE001 -- reported by the scheduler thread
            errors.append(e)
        status[k] = "done"
        back.release()

    threads = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(K)]
    tree.network = YieldingNet()
    try:
        for th in threads:
            th.start()
        active = list(range(K))
        while active:
            for k in list(active):
                state["cur"] = k
                turn[k].release()
                back.acquire()
                if status[k] == "done":
                    active.remove(k)
    finally:
        tree.network = real_net
    if errors:
        raise errors[0]


def _make_tree_class(mcts):
    class TracedTree(mcts.MCTreeSearch):
        """MCTreeSearch that keeps CTX in sync and logs per-move root statistics."""
        tree_id = 0
        move_log = None
        coop_threads = 1     # > 1: the threaded search under the cooperative schedule of _coop_search

        def search(self):
            CTX.tree = self.tree_id
            CTX.ply = int(np.sum(np.abs(self.root_node.state)))
            CTX.sim = -1
            CTX.moves_played = self.moves_played
            if self.coop_threads > 1:     # search() of mcts.py:323-338 with the executor replaced by the forced schedule
                self.root_node.add_noise()
                _coop_search(mcts, self, self.coop_threads)
                self.root_node.remove_noise()
                return None
            return super().search()

        def search_node(self):
            CTX.sim += 1
            CTX.depth = 0
            return super().search_node()

        def _play(self, temp=0.05):
            CTX.tree = self.tree_id
            CTX.ply = int(np.sum(np.abs(self.root_node.state)))
            root = self.root_node
            entry = dict(tree=self.tree_id, ply=CTX.ply,
                         n=[int(c.n) for c in root.children], w=[float(c.w) for c in root.children],
                         root_n=int(root.n), root_w=float(root.w))
            a = super()._play(temp)
            entry["action"] = int(a)
            if self.move_log is not None:
                self.move_log.append(entry)
            return a
    return TracedTree


def _env_cls(game):
    _, _, C4, TTT = _import_reference()
    return C4 if game == spec.GAME_CONNECT4 else TTT


def run_search(game, sims, seed=0, game_uid=0, tie_mode=1, noise=None, net_seed=0, prefix=(),
               alpha=1, network=None, strong_play=False, threads=1):
    """Fresh MCTreeSearch (root player +1), optional ``prefix`` of (action, player) play_action
    calls, then one ``search()``.  Returns root child n, w, valid, root n, w, q."""
    mcts, _, _, _ = _import_reference()
    Tree = _make_tree_class(mcts)
    CTX.seed, CTX.game_uid, CTX.tie_mode = seed, game_uid, tie_mode
    CTX.noise_table = None if noise is None else [[noise]]
    net = network if network is not None else HashNet(game, net_seed)
    with hooked():
        t = Tree(net, _env_cls(game), iterations=sims, thread_count=1, alpha=alpha, strong_play=strong_play)
        t.tree_id = 0
        t.coop_threads = threads
        for a, pl in prefix:
            t.play_action(a, pl)
        if noise is not None:
            CTX.noise_table = [[noise] * 64]
        t.search()
    root = t.root_node
    return dict(n=np.array([c.n for c in root.children], dtype=np.int32),
                w=np.array([float(c.w) for c in root.children], dtype=np.float64),
                valid=np.array([bool(c._valid) for c in root.children]),
                root_n=int(root.n), root_w=float(root.w), q=float(root.q), player=int(root.player))


def run_episode(game, sims, seed=0, game_uid=0, swap_sides=False, evaluate=False, tie_mode=1,
                noise_table=None, net_seed=0, net_seed_opp=None, alpha=1, strong_play=False, threads=1):
    """One ``SelfPlayer.play_episode(swap_sides, update=True)`` with two traced trees
    (selfplayworker.py:67-90,172-194).  ``noise_table``: float64 [2][max_moves][A] or None."""
    mcts, SelfPlayer, _, _ = _import_reference()
    Tree = _make_tree_class(mcts)
    env_cls = _env_cls(game)
    CTX.seed, CTX.game_uid, CTX.tie_mode = seed, game_uid, tie_mode
    CTX.noise_table = noise_table
    memq, resq = queue.Queue(), queue.Queue()
    move_log = []
    with hooked():
        trees = []
        for tid in (0, 1):
            ns = net_seed if (tid == 0 or net_seed_opp is None) else net_seed_opp
            t = Tree(HashNet(game, ns), env_cls, memory_queue=memq, iterations=sims, thread_count=1,
                     alpha=alpha, strong_play=strong_play)
            t.tree_id = tid
            t.move_log = move_log
            t.coop_threads = threads
            t.train(False) if hasattr(t.network, "train") else None
            t.evaluate(evaluate)
            trees.append(t)
        sp = SelfPlayer(trees[0], trees[1], env_cls(), resq)
        out = sp.play_episode(swap_sides=swap_sides, update=True)
    assert out is not None, "reference play_episode swallowed an exception"
    state_list, r = out
    result = resq.get_nowait()
    records = []
    while not memq.empty():
        m = memq.get_nowait()
        records.append(dict(state=m.state.numpy().astype(np.int8), actual_val=float(m.actual_val),
                            tree_probs=m.tree_probs.numpy().astype(np.float32), q=np.float32(m.q.item())))
    return dict(reward=int(r), result=result, moves=move_log, records=records,
                final_state=np.asarray(state_list[-1]).astype(np.int8),
                net_calls=[t.network.calls for t in trees])


def ref_env_playout(game, actions, first_player=None):
    """Reference envs driven exactly like oracle.env_playout (same output layout)."""
    mcts, _, C4, TTT = _import_reference()
    from games.general.base_env import GameOver
    W, H, A = spec.GAME_DIMS[game]
    actions = np.asarray(actions, dtype=np.int32)
    n, T = actions.shape
    out = dict(boards=np.zeros((n, T, W, H), np.int8), reward=np.zeros((n, T), np.int8),
               done=np.zeros((n, T), np.uint8), valid=np.zeros((n, T, A), np.uint8), status=np.zeros((n, T), np.int8))
    for g in range(n):
        env = (C4 if game == spec.GAME_CONNECT4 else TTT)()
        env.reset()
        player = 1 if first_player is None else int(first_player[g])
        for t in range(T):
            a = int(actions[g, t])
            st, r = 0, 0
            if a < 0:
                st = -3
            else:
                try:
                    _, r, _, _ = env.step(a, player=player)
                except GameOver:
                    st = -1
                except ValueError:
                    st = -2
            out["status"][g, t] = st
            out["reward"][g, t] = r if st == 0 else 0
            out["done"][g, t] = bool(env.episode_over)
            out["boards"][g, t] = env.board
            out["valid"][g, t] = env.valid_moves()
            if st == 0:
                player = -player
    return out


def run_episode_vs_hardcoded(game, sims, kind, seed=0, game_uid=0, swap_sides=False, noise_table=None, net_seed=0):
    """SelfPlayer.play_episode(update=False) of a traced MCTreeSearch (evaluate mode, as SelfPlayWorker.set_up_policies
    does for evaluation games, selfplayworker.py:70-81) against the reference's OneStepLookahead / Random player whose
    ``random.choice`` is replaced by the spec stream (PURPOSE_OPPONENT)."""
    mcts, SelfPlayer, _, _ = _import_reference()
    import games.general.hardcoded_players as hp
    Tree = _make_tree_class(mcts)
    env_cls = _env_cls(game)
    CTX.seed, CTX.game_uid, CTX.tie_mode = seed, game_uid, 1
    CTX.noise_table = noise_table
    memq, resq = queue.Queue(), queue.Queue()
    move_log = []
    opp_cls = hp.OneStepLookahead if kind == spec.OPP_LOOKAHEAD else hp.Random

    class TracedOpp(opp_cls):
        def __call__(self, s):
            ply = int(np.sum(np.abs(self.env.board)))
            saved = hp.random.choice
            hp.random.choice = lambda seq: seq[min(int(spec.rng_uniform(seed, game_uid, 1, spec.PURPOSE_OPPONENT, ply, 0, 0, 0) * len(seq)), len(seq) - 1)]
            try:
                a = super().__call__(s)
            finally:
                hp.random.choice = saved
            move_log.append(dict(tree=1, ply=ply, action=int(a)))
            return a
    with hooked():
        t = Tree(HashNet(game, net_seed), env_cls, memory_queue=memq, iterations=sims, thread_count=1)
        t.tree_id, t.move_log = 0, move_log
        t.evaluate(True)
        opp = TracedOpp(env_cls)
        opp.env = env_cls()
        sp = SelfPlayer(t, opp, env_cls(), resq)
        out = sp.play_episode(swap_sides=swap_sides, update=False)
    assert out is not None, "reference play_episode swallowed an exception"
    return dict(reward=int(out[1]), moves=move_log, final_state=np.asarray(out[0][-1]).astype(np.int8))
