"""Shared helpers for the GPU parity tests (engine vs oracle)."""
import ctypes as C
import json
import os

import numpy as np

from oracle import oracle as ox
from oracle import spec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhex(x):
    return float.fromhex(x)


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def record_board(rec, game):
    return spec.bits_to_board(int(rec["own"]), int(rec["opp"]), game).astype(np.int8)


def compare_game(game, engine_moves, engine_records, engine_result, oracle_ep, evaluate=False):
    """Engine output of ONE game vs an oracle episode dict.  Bit-exact on n, w (fp64), root stats,
    actions, record states, q and actual_val; tree_probs bit-exact."""
    A = spec.GAME_DIMS[game][2]
    assert engine_result["reward"] == oracle_ep["reward"], (engine_result, oracle_ep["reward"])
    assert len(engine_moves) == len(oracle_ep["moves"]), (len(engine_moves), len(oracle_ep["moves"]))
    for i, (a, b) in enumerate(zip(engine_moves, oracle_ep["moves"])):
        assert (a["tree"], a["ply"]) == (b["tree"], b["ply"]), (i, a, b)
        assert a["n"] == list(b["n"]), (i, a["n"], b["n"])
        assert a["w"] == list(b["w"]), (i, a["w"], b["w"])
        assert a["root_n"] == b["root_n"] and a["root_w"] == b["root_w"], (i, a, b)
        assert a["action"] == b["action"], (i, a, b)
    if engine_records is None:      # engine ran with emit_records off (evaluation games push nothing, selfplayworker.py:186-190)
        return
    # engine ring order == reference queue order: policy (tree 0) records first, then the opponent's
    assert len(engine_records) == len(oracle_ep["records"])
    for a, b in zip(engine_records, oracle_ep["records"]):
        assert np.array_equal(record_board(a, game), b["state"])
        assert np.array_equal(a["tree_probs"][:A], b["tree_probs"]), (a["tree_probs"], b["tree_probs"])
        assert a["q"] == b["q"] and a["actual_val"] == b["actual_val"]


def oracle_episode(game, sims, seed, game_index, noise_table, evaluate=False, strong_play=False, net_seed=0,
                   net_seed_opp=None, tie_mode=1, threads=1):
    cfg = ox.make_cfg(game, sims, seed=seed, game_uid=game_index, evaluate=evaluate, strong_play=strong_play,
                      noise_table=noise_table, tie_mode=tie_mode, threads=threads)
    return ox.play_episode(cfg, bool(game_index & 1), net_seed=net_seed, net_seed_opp=net_seed_opp)


def split_by_game(records, results):
    recs, res = {}, {}
    for r in records:
        recs.setdefault(int(r["game_index"]), []).append(r)
    for r in results:
        res[int(r["game_index"])] = dict(reward=int(r["reward"]), swap=int(r["swap_sides"]), plies=int(r["plies"]))
    return recs, res


def run_logged(engine):
    """Drive an engine tick by tick, logging every network evaluation per (game slot, tree) in call order.
    Returns logs[g][tree] = dict(own, opp, policy, value) usable by oracle.make_replay (one game per slot only)."""
    import ctypes as C
    import torch
    from self_play_reinforcement_learning_b200 import _lib
    G = engine.n_games
    logs = [[dict(own=[], opp=[], policy=[], value=[]) for _ in (0, 1)] for _ in range(G)]
    tree_t = torch.zeros(G, dtype=torch.int32, device=engine.device)
    for _ in range(10_000_000):
        engine.tick()
        torch.cuda.synchronize()
        need = engine.needs_eval.cpu().numpy().astype(bool)
        if not need.any() and engine.all_idle():
            break
        own = engine.leaf_own.cpu().numpy().view(np.uint64)
        opp = engine.leaf_opp.cpu().numpy().view(np.uint64)
        pol, val = engine.policy.cpu().numpy(), engine.value.cpu().numpy()
        _lib.check(_lib.lib().spx_pending_tree(engine._h, tree_t.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "spx_pending_tree")
        tree = tree_t.cpu().numpy()
        for g in np.flatnonzero(need):
            L = logs[g][tree[g]]
            L["own"].append(own[g]); L["opp"].append(opp[g]); L["policy"].append(pol[g].copy()); L["value"].append(val[g])
    return logs


def replay_in_oracle(game, sims, seed, game_index, noise_table, log, evaluate=False, strong_play=False, noise_mode=None):
    """Oracle episode fed with the logged network outputs; asserts the oracle asked for exactly the same leaves."""
    import ctypes as C
    rs = ox.make_replay(game, log)
    cfg = ox.make_cfg(game, sims, seed=seed, game_uid=game_index, noise_table=noise_table, evaluate=evaluate,
                      strong_play=strong_play, noise_mode=noise_mode)
    pair = (ox.fn_addr("ox_replaynet"), C.addressof(rs))
    o = ox.play_episode(cfg, bool(game_index & 1), nets=(pair, pair))
    assert rs.mismatches == 0 and rs.overruns == 0, (game_index, rs.mismatches, rs.overruns)
    assert rs.cursor[0] == rs.n[0] and rs.cursor[1] == rs.n[1], game_index
    return o


def run_logged_device(engine, slots, step, max_ticks=10_000_000, poll=1024):
    """Drive `engine` with step() (ONE tick per call: e.g. ``lambda: engine.run_ticks(1, fused=True, chunk=1)``) until every slot
    is idle, logging the evaluations of the sampled `slots` on the device (no host round trip per tick).
    Returns games[slot] = list (one entry per game played on the slot, in order) of per-tree logs
    ``[dict(own, opp, policy, value), dict(...)]`` usable by oracle.make_replay / replay_in_oracle."""
    import torch
    from self_play_reinforcement_learning_b200 import _lib
    dev, A, S = engine.device, engine.A, len(slots)
    idx = torch.as_tensor(np.asarray(slots, dtype=np.int64), device=dev)
    tree_all = torch.zeros(engine.n_games, dtype=torch.int32, device=dev)
    need_b = torch.zeros(poll, S, dtype=torch.uint8, device=dev)
    own_b = torch.zeros(poll, S, dtype=torch.int64, device=dev)
    opp_b = torch.zeros(poll, S, dtype=torch.int64, device=dev)
    pol_b = torch.zeros(poll, S, A, dtype=torch.float32, device=dev)
    val_b = torch.zeros(poll, S, dtype=torch.float32, device=dev)
    tree_b = torch.zeros(poll, S, dtype=torch.int32, device=dev)
    chunks, t, done = [], 0, False
    while not done and t < max_ticks:
        for i in range(poll):
            step()
            _lib.check(_lib.lib().spx_pending_tree(engine._h, tree_all.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "spx_pending_tree")
            torch.index_select(engine.needs_eval, 0, idx, out=need_b[i])
            torch.index_select(engine.leaf_own, 0, idx, out=own_b[i])
            torch.index_select(engine.leaf_opp, 0, idx, out=opp_b[i])
            torch.index_select(engine.policy, 0, idx, out=pol_b[i])
            torch.index_select(engine.value, 0, idx, out=val_b[i])
            torch.index_select(tree_all, 0, idx, out=tree_b[i])
        t += poll
        chunks.append(tuple(x.cpu().numpy().copy() for x in (need_b, own_b, opp_b, pol_b, val_b, tree_b)))
        done = engine.all_idle()
    assert done, "run_logged_device: max_ticks reached"
    need, own, opp, pol, val, tree = (np.concatenate([c[k] for c in chunks]) for k in range(6))
    own, opp = own.view(np.uint64), opp.view(np.uint64)
    games = {}
    for j, slot in enumerate(slots):
        out = []
        for i in np.flatnonzero(need[:, j]):
            tr = int(tree[i, j])
            if tr == 0 and own[i, j] == 0 and opp[i, j] == 0:       # MCTreeSearch.reset of the policy's tree: a new game begins
                out.append([dict(own=[], opp=[], policy=[], value=[]) for _ in (0, 1)])
            L = out[-1][tr]
            L["own"].append(own[i, j]); L["opp"].append(opp[i, j]); L["policy"].append(pol[i, j].copy()); L["value"].append(val[i, j])
        games[int(slot)] = out
    return games


def compare_records_and_result(game, engine_records, engine_result, oracle_ep):
    """compare_game without the per-move root statistics (the engine's move log holds only the slot's latest game)."""
    A = spec.GAME_DIMS[game][2]
    assert engine_result["reward"] == oracle_ep["reward"] and engine_result["plies"] == len(oracle_ep["moves"])
    assert len(engine_records) == len(oracle_ep["records"])
    for a, b in zip(engine_records, oracle_ep["records"]):
        assert np.array_equal(record_board(a, game), b["state"])
        assert np.array_equal(a["tree_probs"][:A], b["tree_probs"]), (a["tree_probs"], b["tree_probs"])
        assert a["q"] == b["q"] and a["actual_val"] == b["actual_val"]
