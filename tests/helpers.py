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
    # engine ring order == reference queue order: policy (tree 0) records first, then the opponent's
    assert len(engine_records) == len(oracle_ep["records"])
    for a, b in zip(engine_records, oracle_ep["records"]):
        assert np.array_equal(record_board(a, game), b["state"])
        assert np.array_equal(a["tree_probs"][:A], b["tree_probs"]), (a["tree_probs"], b["tree_probs"])
        assert a["q"] == b["q"] and a["actual_val"] == b["actual_val"]


def oracle_episode(game, sims, seed, game_index, noise_table, evaluate=False, strong_play=False, net_seed=0,
                   net_seed_opp=None, tie_mode=1):
    cfg = ox.make_cfg(game, sims, seed=seed, game_uid=game_index, evaluate=evaluate, strong_play=strong_play,
                      noise_table=noise_table, tie_mode=tie_mode)
    return ox.play_episode(cfg, bool(game_index & 1), net_seed=net_seed, net_seed_opp=net_seed_opp)


def split_by_game(records, results):
    recs, res = {}, {}
    for r in records:
        recs.setdefault(int(r["game_index"]), []).append(r)
    for r in results:
        res[int(r["game_index"])] = dict(reward=int(r["reward"]), swap=int(r["swap_sides"]), plies=int(r["plies"]))
    return recs, res
