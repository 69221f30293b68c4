"""SelfPlayEngine: host-side driver of one spx_engine (one per GPU).

One ``tick`` = spx_advance (every game's state machine runs until it needs the network) followed by
one batched network evaluation of the dense leaf batch.  This replaces the reference's
SelfPlayWorker threads + InferenceProxy queues + InferenceWorker batching
(selfplayworker.py:95-142, inference_proxy.py:21-24, inference_worker.py:61-119).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import Config, Counters, MoveLog, Record, Result, check, lib

RECORD_DTYPE = np.dtype([("own", "<u8"), ("opp", "<u8"), ("game_index", "<u8"), ("tree_probs", "<f4", (9,)),
                         ("q", "<f4"), ("actual_val", "<f4"), ("tree", "u1"), ("ply", "u1"), ("pad0", "<u2"),
                         ("pad1", "<u8")])
RESULT_DTYPE = np.dtype([("game_index", "<u8"), ("reward", "i1"), ("swap_sides", "u1"), ("plies", "u1"),
                         ("pad", "u1", (5,))])
assert RECORD_DTYPE.itemsize == C.sizeof(Record) == 80
assert RESULT_DTYPE.itemsize == C.sizeof(Result) == 16


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view engine-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2}


def _view(ptr, shape, typestr, device):
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class HashNetEvaluator:
    """The synthetic oracle/spec.py hash network evaluated by a libspx kernel (tests, search-only bench)."""

    def __init__(self, game, net_seed=0, net_seed_opp=None):
        self.game = game
        self.seed0 = int(net_seed)
        self.seed1 = int(net_seed if net_seed_opp is None else net_seed_opp)

    def bind(self, engine):
        self.engine = engine

    def __call__(self, engine, events=None):
        check(lib().spx_hashnet_forward(self.game, engine.n_leaves, engine.leaf_own.data_ptr(), engine.leaf_opp.data_ptr(),
                                        engine.needs_eval.data_ptr(), engine.net_id.data_ptr(), self.seed0, self.seed1,
                                        engine.policy.data_ptr(), engine.value.data_ptr(), _stream_ptr()),
              "spx_hashnet_forward")


class SelfPlayEngine:
    """n_games concurrent self-play games (two MCTS trees each) on the current CUDA device."""

    def __init__(self, game, n_games, sims, evaluator, *, evaluate=False, strong_play=False, alpha=1.0, seed=0,
                 tie_mode=1, noise_mode=2, emit_records=True, max_sims_per_tick=8, nodes_per_tree=0, move_log=False,
                 two_nets=False, opponent_kind=0, slot_offset=0, slot_stride=None, games_target=None, record_capacity=None,
                 result_capacity=None, search_threads=1, eval_cache=0):
        """search_threads: MCTreeSearch(thread_count=K) behind an InferenceProxy (mcts.py:132,328-331; the reference's default is 4):
        K search_node tasks in flight per tree with virtual loss and per-child locks, under the cooperative round-robin
        schedule.  1 = the sequential search.  With K > 1 the leaf batch has n_games * K rows.
        eval_cache: 0 / False = off; True = 4096 entries per game slot; n >= 6 = 2**n entries (64 B each).  The fused tick kernel
        then answers requests for positions the slot has evaluated before (same weights) from the table instead of the
        network -- about half of all requests at 800 sims/move -- without changing the games."""
        if not torch.cuda.is_available():
            raise _lib.SpxError("SelfPlayEngine needs a CUDA device (B200); there is no CPU fallback")
        self.game, self.n_games, self.sims = game, int(n_games), int(sims)
        self.W, self.H, self.A = _lib.GAME_DIMS[game]
        self.device = torch.device("cuda", torch.cuda.current_device())
        cfg = Config()
        cfg.game, cfg.n_games, cfg.sims = game, n_games, sims
        cfg.evaluate, cfg.strong_play, cfg.tie_mode, cfg.noise_mode = int(evaluate), int(strong_play), tie_mode, noise_mode
        cfg.emit_records, cfg.max_sims_per_tick, cfg.nodes_per_tree = int(emit_records), max_sims_per_tick, nodes_per_tree
        cfg.move_log, cfg.two_nets, cfg.alpha, cfg.seed = int(move_log), int(two_nets), float(alpha), int(seed)
        cfg.opponent_kind = int(opponent_kind)
        cfg.search_threads = int(search_threads)
        cfg.eval_cache_log2 = 12 if eval_cache is True else int(eval_cache or 0)
        self.search_threads = max(1, int(search_threads))
        self.n_leaves = int(n_games) * self.search_threads
        cfg.reserved0 = int(os.environ.get("SPX_DBG_FLAGS", "0"), 0)   # timing experiments only (csrc/spx_tower.cu)
        cfg.slot_offset = slot_offset
        cfg.slot_stride = n_games if slot_stride is None else slot_stride
        cfg.games_target = (1 << 62) if games_target is None else games_target
        max_moves = self.W * self.H
        cfg.record_capacity = record_capacity or max(4 * n_games * max_moves, 1024)
        cfg.result_capacity = result_capacity or max(8 * n_games, 1024)
        self.cfg = cfg
        self._h = C.c_void_p()
        check(lib().spx_create(C.byref(cfg), C.byref(self._h)), "spx_create")
        own, opp, need, nid = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().spx_leaf_batch(self._h, C.byref(own), C.byref(opp), C.byref(need), C.byref(nid)), "spx_leaf_batch")
        nl = self.n_leaves
        self.leaf_own = _view(own.value, (nl,), "<i8", self.device)
        self.leaf_opp = _view(opp.value, (nl,), "<i8", self.device)
        self.needs_eval = _view(need.value, (nl,), "|u1", self.device)
        self.net_id = _view(nid.value, (nl,), "|u1", self.device)
        self.policy = torch.zeros(nl, self.A, dtype=torch.float32, device=self.device)
        self.value = torch.zeros(nl, dtype=torch.float32, device=self.device)
        self._noise_table = None
        self._first = True
        self.evaluator = evaluator
        if hasattr(evaluator, "bind"):
            evaluator.bind(self)
        self._rec_host = np.zeros(cfg.record_capacity, RECORD_DTYPE)
        self._res_host = np.zeros(cfg.result_capacity, RESULT_DTYPE)

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            torch.cuda.synchronize()
            lib().spx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(lib().spx_reset(self._h, _stream_ptr()), "spx_reset")
        self._first = True

    def set_sims(self, sims):
        """Simulations per move (MCTreeSearch.iterations) for the launches that follow; at most the creation-time value."""
        check(lib().spx_set_sims(self._h, int(sims)), "spx_set_sims")
        self.sims = int(sims)

    def stagger(self, generations=3, sims=2):
        """Decorrelate the slots' game phases: play `generations` rounds of quick games (`sims` simulations per move) so that
        the slots sit at different plies of different games, then restore the configured simulation count.  Benchmarks call this
        once before warm-up: all slots otherwise start their first game together and throughput drifts with the game phase
        (early plies have no terminal nodes, late plies many)."""
        full = self.sims
        self.set_sims(min(int(sims), full))
        max_moves = self.W * self.H
        # records/results of the quick games are discarded in bounded chunks (the rings must not overflow)
        ticks, chunk = int(generations * max_moves * (self.sims + 2)), int(max(8, min(512, 2 * self.sims)))
        done = 0
        while done < ticks:
            self.run_ticks(chunk, chunk=chunk)
            self.drain_records(); self.drain_results()
            done += chunk
        self.set_sims(full)

    def set_noise_table(self, table, first_game_index=0):
        """table: float64 [n_table_games, 2, table_moves, A] (host array or device tensor)."""
        t = torch.as_tensor(np.ascontiguousarray(table, dtype=np.float64) if not torch.is_tensor(table) else table)
        t = t.to(self.device, torch.float64).contiguous()
        if not (t.dim() == 4 and t.shape[1] == 2 and t.shape[3] == self.A):
            raise ValueError(f"noise table must be [n_table_games, 2, table_moves, {self.A}], got {tuple(t.shape)}")
        self._noise_table = t
        check(lib().spx_set_noise_table(self._h, t.data_ptr(), first_game_index, t.shape[0], t.shape[2]), "spx_set_noise_table")

    # ------------------------------------------------------------------ the hot loop
    def advance(self, events=None):
        """spx_advance only (consume last outputs, fill the leaf batch).  ``events``: optional pair of
        torch.cuda.Event recorded right before/after the launch on the launching stream."""
        if self.cfg.eval_cache_log2:   # the evaluation cache needs to know which weights the outputs it is about to consume came from
            vers = getattr(self.evaluator, "cache_versions", None)
            vers = tuple(vers()) if vers else (0, 0)
            if vers != getattr(self, "_cache_vers", None):
                check(lib().spx_set_eval_cache_versions(self._h, vers[0], vers[1]), "spx_set_eval_cache_versions")
                self._cache_vers = vers
        p = None if self._first else self.policy.data_ptr()
        v = None if self._first else self.value.data_ptr()
        if events is None:
            check(lib().spx_advance(self._h, p, v, _stream_ptr()), "spx_advance")
        else:
            check(lib().spx_advance_timed(self._h, p, v, _stream_ptr(), C.c_void_p(events[0].cuda_event),
                                          C.c_void_p(events[1].cuda_event)), "spx_advance_timed")
        self._first = False

    def tick(self, advance_events=None, net_events=None):
        self.advance(advance_events)
        if net_events is None:
            self.evaluator(self)
        else:
            self.evaluator(self, events=net_events)

    def run_ticks(self, n, fused=True, chunk=400, balanced=None):
        """n ticks.  With the native tower the whole loop runs as persistent launches of `chunk` ticks (spx_tick_fused: the
        network CTAs also advance their games); any other evaluator, or fused=False, launches advance + evaluation per tick.
        balanced: launches are work-conserving (spx_tick_fused_balanced: SM pairs draw their ticks from the launch's budget, a
        game gets `chunk` ticks per launch on average instead of exactly; games themselves do not depend on it).  Default: on
        for launches of at least 16 ticks (SPX_TICK_BALANCE=0 switches it off); callers that count ticks per game pass False."""
        fn = getattr(self.evaluator, "fused_ticks", None) if (fused and self.search_threads == 1) else None
        if balanced is None:
            balanced = os.environ.get("SPX_TICK_BALANCE", "1") != "0"
        done = 0
        while fn is not None and done < n:
            k = min(chunk, n - done)
            if not fn(self, k, balanced=bool(balanced) and k >= 16):
                break
            done += min(chunk, n - done)
            self._first = False
        for _ in range(n - done):
            self.tick()

    def run_until_idle(self, max_ticks=10_000_000, poll_every=64):
        t = 0
        while t < max_ticks:
            self.run_ticks(poll_every)
            t += poll_every
            if self.all_idle():
                return t
        raise _lib.SpxError("run_until_idle: max_ticks reached")

    # ------------------------------------------------------------------ readouts
    @property
    def safe_poll_interval(self):
        """Ticks between two drains that cannot overflow the default rings: a ply takes at least sims / max_sims_per_tick
        ticks and a game at least 5 plies, so 2 * sims ticks hold fewer than 4 finished games per slot."""
        return int(max(8, min(512, 2 * self.sims)))

    def check_overflow(self):
        """Raises if records were dropped since the engine was created (the device ring was drained too rarely): training
        data must not vanish silently.  (Lost game results make spx_drain_results itself fail.)"""
        c = self.counters()
        if c["records_dropped"]:
            raise _lib.SpxError(f"{c['records_dropped']} self-play records were dropped: the record ring overflowed between two "
                                "drains (drain more often or pass a larger record_capacity)")
        return c

    def all_idle(self):
        out = C.c_int32()
        check(lib().spx_all_idle(self._h, C.byref(out), _stream_ptr()), "spx_all_idle")
        return bool(out.value)

    def counters(self):
        c = Counters()
        check(lib().spx_counters_read(self._h, C.byref(c), _stream_ptr()), "spx_counters_read")
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    def drain_records(self):
        n = C.c_int64()
        check(lib().spx_drain_records(self._h, self._rec_host.ctypes.data, len(self._rec_host), C.byref(n), _stream_ptr()),
              "spx_drain_records")
        return self._rec_host[:n.value].copy()

    def drain_results(self):
        n = C.c_int64()
        check(lib().spx_drain_results(self._h, self._res_host.ctypes.data, len(self._res_host), C.byref(n), _stream_ptr()),
              "spx_drain_results")
        return self._res_host[:n.value].copy()

    def root_stats(self, tree):
        G, A = self.n_games, self.A
        n = torch.zeros(G, A, dtype=torch.int32, device=self.device)
        w = torch.zeros(G, A, dtype=torch.float64, device=self.device)
        rn = torch.zeros(G, dtype=torch.int32, device=self.device)
        rw = torch.zeros(G, dtype=torch.float64, device=self.device)
        valid = torch.zeros(G, dtype=torch.int16, device=self.device)
        check(lib().spx_root_stats(self._h, tree, n.data_ptr(), w.data_ptr(), rn.data_ptr(), rw.data_ptr(), valid.data_ptr(),
                                   _stream_ptr()), "spx_root_stats")
        torch.cuda.synchronize()
        vm = valid.cpu().numpy().astype(np.uint16)
        return dict(n=n.cpu().numpy(), w=w.cpu().numpy(), root_n=rn.cpu().numpy(), root_w=rw.cpu().numpy(),
                    valid=((vm[:, None] >> np.arange(A)) & 1).astype(bool))

    def move_log(self, slot):
        buf = (MoveLog * 44)()
        n = C.c_int32()
        check(lib().spx_read_move_log(self._h, slot, buf, 44, C.byref(n), _stream_ptr()), "spx_read_move_log")
        A = self.A
        return [dict(tree=m.tree, ply=m.ply, action=m.action, root_n=m.root_n, root_w=m.root_w, n=list(m.n[:A]),
                     w=list(m.w[:A]), noise=list(m.noise[:A])) for m in buf[:n.value]]

    def device_bytes(self):
        return int(lib().spx_device_bytes(self._h))
