"""BatchedSelfPlay: the per-GPU replacement of the reference's worker layer.

One object stands in for a whole set of ``SelfPlayWorker`` processes plus the ``InferenceWorker`` and the
``InferenceProxy`` queues between them (games/algos/selfplayworker.py:17-161, inference_worker.py:13-119,
inference_proxy.py:9-36): it keeps ``n_games`` games in flight on one B200, evaluates all their leaves in one
batch per tick, and speaks the scheduler's protocol on the way out:

  * results  -> ``{"reward": r, "swap_sides": b}`` dicts, one per finished game        (selfplayworker.py:185)
  * records  -> ``Move(state, actual_val, tree_probs, q)`` tuples of torch tensors with the reference's dtypes
                (int64 [W,H] state in the mover's own frame, float32 scalars/vectors)     (mcts.py:17,282-289,230)
  * tasks    -> ``{"play": {"swap_sides": b, "update": u}, "evaluate": e}`` dicts are accepted by ``run_tasks``
                (self_play_parallel.py:238,253,299,368); ``task_done()`` is called once per task.
Weights are refreshed with ``load_weights`` (one H2D/D2D copy of the packed blob) instead of the reference's
checkpoint-file polling (inference_worker.py:68-73).
"""
from collections import namedtuple

import numpy as np
import torch

from . import _lib, nets
from .engine import RECORD_DTYPE, RESULT_DTYPE, HashNetEvaluator, SelfPlayEngine
from .envs import game_id_of

Move = namedtuple("Move", ("state", "actual_val", "tree_probs", "q"))  # mcts.py:17


def _bits_to_board_np(own, opp, game):
    W, H, _ = _lib.GAME_DIMS[game]
    stride = 7 if game == _lib.GAME_CONNECT4 else 3
    idx = (np.arange(W)[:, None] * stride + np.arange(H)[None, :]).astype(np.uint64)
    o = (own[:, None, None] >> idx) & np.uint64(1)
    e = (opp[:, None, None] >> idx) & np.uint64(1)
    return o.astype(np.int64) - e.astype(np.int64)


def records_to_moves(records, game):
    """Structured record array (engine.RECORD_DTYPE) -> list of reference-format Move tuples."""
    A = _lib.GAME_DIMS[game][2]
    if len(records) == 0:
        return []
    boards = torch.from_numpy(_bits_to_board_np(records["own"], records["opp"], game))
    probs = torch.from_numpy(np.ascontiguousarray(records["tree_probs"][:, :A]))
    q = torch.from_numpy(np.ascontiguousarray(records["q"]))
    val = torch.from_numpy(np.ascontiguousarray(records["actual_val"]))
    # every field of every Move owns its storage, like the reference's (torch.tensor(...) per record, mcts.py:282-289): a view
    # into the batch tensors would drag the WHOLE batch along whenever one Move is pickled (memory_queue, save_memory)
    return [Move(b.clone(), v.clone(), p.clone(), x.clone()) for b, v, p, x in zip(boards.unbind(0), val.unbind(0), probs.unbind(0), q.unbind(0))]


def results_to_dicts(results):
    return [{"reward": int(r["reward"]), "swap_sides": bool(r["swap_sides"])} for r in results]


class BatchedSelfPlay:
    def __init__(self, network, game=None, env=None, n_games=1024, sims=800, net="tower", evaluation_network=None,
                 evaluate=False, update=True, alpha=1.0, strong_play=False, seed=0, rank=0, world=1, games_target=None,
                 max_sims_per_tick=None, noise_mode=2, tie_mode=1, move_log=False, net_dtype=torch.bfloat16, opponent=None,
                 search_threads=1, eval_cache=0):
        """network / evaluation_network: nn.Module (ResidualTower for the native tower; any board net for net='torch').
        env: a reference env class/instance (mapped by variant_string) or ``game`` id.  iterations == sims.
        opponent: None (MCTS self-play / evaluation network), "lookahead" or "random": the reference's hard-coded
        evaluation opponents (general/hardcoded_players.py; the default evaluation_policy_container of main.py:66).
        search_threads: MCTreeSearch(thread_count=K) (mcts.py:132; the reference's default is 4 behind its InferenceProxy): K
        simulations in flight per tree with virtual loss -- a move then takes iterations / K ticks, which is what fills the GPU
        when an epoch has fewer games than leaf slots (the reference's 750-1500 games); 1 = the sequential search.
        eval_cache: the engine's per-slot evaluation cache (SelfPlayEngine; native tower, one network, sequential search)."""
        self.opponent_kind = {None: 0, "mcts": 0, "lookahead": 1, "random": 2}[opponent]
        self.game = game_id_of(env if env is not None else game)
        self.network, self.evaluation_network = network, evaluation_network
        two = evaluation_network is not None
        if net == "tower":
            self.evaluator = nets.TwoTowerEvaluator(network, evaluation_network, self.game) if two else nets.TowerEvaluator(network, self.game)
        elif net == "tttnet":
            self.evaluator = nets.TTTNetEvaluator(network)
        elif net == "torch":
            self.evaluator = nets.TorchNetEvaluator(network, self.game, module_opp=evaluation_network, dtype=net_dtype)
        elif net == "hash":
            self.evaluator = HashNetEvaluator(self.game, seed, None if not two else seed + 1)
        else:
            raise ValueError(net)
        self.net_kind = net
        if max_sims_per_tick is None:   # simulations a slot may run between two evaluations (terminal re-visits, cache hits): measured optimum
            max_sims_per_tick = 16 if eval_cache else 8
        self.engine = SelfPlayEngine(self.game, n_games, sims, self.evaluator, evaluate=evaluate, strong_play=strong_play, alpha=alpha,
                                     seed=seed, tie_mode=tie_mode, noise_mode=noise_mode, emit_records=update, two_nets=two,
                                     max_sims_per_tick=max_sims_per_tick, move_log=move_log, opponent_kind=self.opponent_kind,
                                     slot_offset=rank * n_games,
                                     slot_stride=world * n_games, games_target=games_target, search_threads=search_threads,
                                     eval_cache=eval_cache if (net == "tower" and search_threads <= 1) else 0)
        self._pinned = None

    # ------------------------------------------------------------------ weights
    def packed_weights_pinned(self, module=None):
        """Packed weight blob of ``module`` (default: the construction network) in pinned host memory."""
        tw = self.evaluator.tower if self.net_kind == "tower" else None
        blob = nets.pack_tower_blob(module if module is not None else self.network, tw.ncta if tw else 2, tw.f16 if tw else None)
        self._pinned = blob.pin_memory()
        return self._pinned

    def load_weights(self, module_or_blob):
        """Replaces 'InferenceWorker reloads the newest checkpoint' (inference_worker.py:68-73)."""
        if self.net_kind == "tower":
            self.evaluator.load(module_or_blob)
            return int(module_or_blob.numel()) if torch.is_tensor(module_or_blob) else 0
        if not torch.is_tensor(module_or_blob) and hasattr(self.evaluator, "refresh"):
            self.evaluator.refresh(module_or_blob)     # the evaluator holds a private (device, bf16) copy of the master module
        return 0

    # ------------------------------------------------------------------ running
    def play_step(self, ticks, weights_host=None, fused=True):
        """One host-visible step: optional weight upload from (pinned) host memory, ``ticks`` engine ticks, then the
        step's records, results and counters copied back to the host."""
        h2d = 0
        if weights_host is not None:
            h2d += self.load_weights(weights_host)
        self.engine.run_ticks(ticks, fused=fused)
        recs = self.engine.drain_records()
        res = self.engine.drain_results()
        cnt = self.engine.counters()
        d2h = recs.nbytes + res.nbytes + 80 + 16
        return {"records": recs, "results": res, "counters": cnt, "h2d_bytes": h2d, "d2h_bytes": d2h}

    def play_games(self, max_ticks=50_000_000, poll_every=None):
        """Run until every slot is idle (needs a finite games_target); returns (Move list, result dict list)."""
        recs, res = [], []
        t = 0
        poll_every = poll_every or self.engine.safe_poll_interval
        while t < max_ticks:
            self.engine.run_ticks(poll_every)
            t += poll_every
            recs.append(self.engine.drain_records())
            res.append(self.engine.drain_results())
            if self.engine.all_idle():
                break
        self.engine.check_overflow()
        recs = np.concatenate(recs) if recs else np.zeros(0, RECORD_DTYPE)
        res = np.concatenate(res) if res else np.zeros(0, RESULT_DTYPE)
        order = np.argsort(res["game_index"], kind="stable")
        return records_to_moves(recs, self.game), results_to_dicts(res[order])

    def close(self):
        self.engine.close()
        if hasattr(self.evaluator, "close"):
            self.evaluator.close()
        else:
            tw = getattr(self.evaluator, "tower", None)
            if tw is not None:
                tw.close()


def run_tasks(network, env, tasks, result_queue=None, memory_queue=None, task_queue=None, iterations=800, n_games=None,
              evaluation_network=None, net="tower", seed=0, alpha=1.0, strong_play=False, **kw):
    """Consume a batch of scheduler task dicts the way SelfPlayWorker.run does (selfplayworker.py:105-142) -- but all at
    once on the GPU.  ``tasks``: list of {"play": {"swap_sides": b, "update": u}, "evaluate": e}.  All tasks of one call
    must share ``update``/``evaluate`` (the scheduler issues them that way: self_play_parallel.py:236-253,296-300,366-368).
    The engine derives swap_sides from the game index parity, so tasks are assigned indices of matching parity."""
    if not tasks:
        return [], []
    update = bool(tasks[0]["play"].get("update", False))
    evaluate = bool(tasks[0].get("evaluate", False))
    if not all(bool(t["play"].get("update", False)) == update and bool(t.get("evaluate", False)) == evaluate for t in tasks):
        raise ValueError("run_tasks: one call takes tasks of one kind (same 'update' and 'evaluate' flags), as one scheduler phase produces them")
    n_swap = sum(1 for t in tasks if t["play"].get("swap_sides", False))
    n_plain = len(tasks) - n_swap
    target = 2 * max(n_swap, n_plain)                 # even indices: swap_sides False, odd: True
    G = n_games or min(1024, max(2, target + (target & 1)))
    G += G & 1
    kw.setdefault("eval_cache", True)   # same games, fewer network passes
    sp = BatchedSelfPlay(network, env=env, n_games=G, sims=iterations, net=net, evaluation_network=evaluation_network if evaluate else None,
                         evaluate=evaluate, update=update, alpha=alpha, strong_play=strong_play, seed=seed, games_target=target, **kw)
    moves_all, results_all = [], []
    recs, res = [], []
    while True:
        sp.engine.run_ticks(sp.engine.safe_poll_interval)
        recs.append(sp.engine.drain_records())
        res.append(sp.engine.drain_results())
        if sp.engine.all_idle():
            break
    sp.engine.check_overflow()      # a dropped record or result would also leave task_queue.join() hanging
    recs, res = np.concatenate(recs), np.concatenate(res)
    keep_plain = set(sorted(int(i) for i in res["game_index"] if i % 2 == 0)[:n_plain])
    keep_swap = set(sorted(int(i) for i in res["game_index"] if i % 2 == 1)[:n_swap])
    keep = keep_plain | keep_swap
    for r in res[np.argsort(res["game_index"], kind="stable")]:
        if int(r["game_index"]) in keep:
            d = {"reward": int(r["reward"]), "swap_sides": bool(r["swap_sides"])}
            results_all.append(d)
            if result_queue is not None:
                result_queue.put(d)
            if task_queue is not None:
                task_queue.task_done()
    if update:
        mask = np.array([int(i) in keep for i in recs["game_index"]], dtype=bool)
        moves_all = records_to_moves(recs[mask], sp.game)
        if memory_queue is not None:
            for m in moves_all:
                memory_queue.put(m)
    sp.close()
    return moves_all, results_all
