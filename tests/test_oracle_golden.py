"""Pins the C oracle (oracle/spx_oracle.c) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only.  Bit-exact on every integer and fp64 quantity."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as ox
from oracle import spec


def unhex(x):
    return float.fromhex(x)


@pytest.mark.parametrize("name,game", [("env_connect4.npz", 0), ("env_tictactoe.npz", 1)])
def test_env_matches_reference(golden_dir, name, game):
    g = np.load(os.path.join(golden_dir, name))
    out = ox.env_playout(game, g["actions"], g["first_player"])
    for k in ("status", "reward", "done", "boards", "valid"):
        assert np.array_equal(out[k], g[k]), k
    # the fixture really exercises the edge cases (connect4env.py:30-31,36-37; tictactoe_env.py:28-29)
    if game == 0:
        assert (g["status"] == -1).any() and (g["status"] == -2).any()
        assert ((g["done"] == 1) & (g["reward"] == 0) & (g["status"] == 0)).any()  # a draw at ply 42
    assert (g["reward"] == 1).any()


def _search_cases(golden_dir):
    with open(os.path.join(golden_dir, "search.json")) as f:
        return json.load(f)


def test_search_matches_reference(golden_dir):
    cases = _search_cases(golden_dir)
    assert len(cases) >= 15
    for c in cases:
        p = c["params"]
        game, A = c["game"], spec.GAME_DIMS[c["game"]][2]
        noise = p.get("noise")
        table = None if noise is None else np.tile(np.asarray(noise, np.float64), (1, 64, 1))
        cfg = ox.make_cfg(game, c["sims"], seed=p.get("seed", 0), game_uid=p.get("game_uid", 0),
                          tie_mode=p.get("tie_mode", 1), noise_table=table, strong_play=p.get("strong_play", False))
        if c["net_kind"] == "uniform":
            net = ox.PyNet(game, lambda s, tree: ([1.0 / A] * A, 0.0))
            t = ox.Tree(cfg, net_addr=net.addr, net_user=None)
        else:
            t = ox.Tree(cfg, hash_seed=p.get("net_seed", 0))
        t.reset(1)
        for a, _pl in p.get("prefix", []):
            t.play_action(a)
        t.search()
        o = t.root_stats()
        assert o["n"].tolist() == c["n"], c["name"]
        assert o["w"].tolist() == [unhex(x) for x in c["w"]], c["name"]
        assert o["valid"].tolist() == c["valid"], c["name"]
        assert o["root_n"] == c["root_n"] and o["root_w"] == unhex(c["root_w"]) and o["q"] == unhex(c["q"]), c["name"]
        assert o["player"] == c["player"]


def test_known_answer_vectors(golden_dir):
    """SURVEY.md Appendix B #1-#5 as captured there (independent of the json round trip)."""
    by = {c["name"]: c for c in _search_cases(golden_dir)}
    assert by["B1"]["n"] == [8, 7, 7, 7, 7, 7, 7] and by["B1"]["root_n"] == 50
    assert by["B2"]["n"] == [115, 115, 114, 114, 114, 114, 114]
    assert by["B3"]["n"] == [12] + [11] * 8
    assert by["B4"]["n"] == [34, 34, 33, 0, 33, 33, 33] and by["B4"]["valid"] == [True] * 3 + [False] + [True] * 3
    assert by["B5"]["n"] == [7, 7, 7, 158, 7, 7, 7] and unhex(by["B5"]["q"]) == 158 / 201 and by["B5"]["root_n"] == 201


def test_episodes_match_reference(golden_dir):
    with open(os.path.join(golden_dir, "episodes.json")) as f:
        eps = json.load(f)
    assert len(eps) >= 12
    for e in eps:
        table = np.array([[[unhex(x) for x in row] for row in t] for t in e["noise_table"]], np.float64)
        cfg = ox.make_cfg(e["game"], e["sims"], seed=e["seed"], game_uid=e["game_uid"], evaluate=e["evaluate"],
                          strong_play=e["strong_play"], noise_table=table)
        o = ox.play_episode(cfg, e["swap"], net_seed=e["net_seed"], net_seed_opp=e["net_seed_opp"])
        assert o["reward"] == e["reward"]
        assert len(o["moves"]) == len(e["moves"])
        for a, b in zip(o["moves"], e["moves"]):
            assert (a["tree"], a["ply"], a["action"], a["root_n"]) == (b["tree"], b["ply"], b["action"], b["root_n"])
            assert a["n"] == b["n"]
            assert a["w"] == [unhex(x) for x in b["w"]] and a["root_w"] == unhex(b["root_w"])
        assert len(o["records"]) == len(e["records"])
        for a, b in zip(o["records"], e["records"]):
            assert a["state"].tolist() == b["state"]
            assert a["actual_val"] == b["actual_val"]
            assert a["q"] == np.float32(unhex(b["q"]))
            got, want = a["tree_probs"], np.array([unhex(x) for x in b["tree_probs"]], np.float32)
            if e["evaluate"]:
                # n**20: the oracle is correctly rounded (exact integer power); numpy's pow is within 1 ulp
                # of that in fp64, so the f32 records agree to <= 1 f32 ulp (documented in DESIGN.md)
                assert np.all(np.abs(got - want) <= np.spacing(np.maximum(np.abs(want), np.float32(1e-30))))
            else:
                assert np.array_equal(got, want)
        assert o["final_state"].tolist() == e["final_state"]


def test_spec_stream_and_hashnet_agree_between_python_and_c():
    rng = np.random.default_rng(0)
    for _ in range(200):
        args = [int(rng.integers(0, 2**62)), int(rng.integers(0, 2**40)), int(rng.integers(0, 2)), int(rng.integers(0, 3)),
                int(rng.integers(0, 43)), int(rng.integers(0, 1600)), int(rng.integers(0, 43)), int(rng.integers(0, 9))]
        assert spec.rng_uniform(*args) == ox.lib().ox_rng_uniform(*args)
        own, opp = int(rng.integers(0, 2**48)), int(rng.integers(0, 2**48))
        for A in (7, 9):
            p, v = spec.hashnet(own, opp, A, args[0])
            p2, v2 = ox.hashnet_bits(own, opp, A, args[0])
            assert np.array_equal(p, p2) and v == v2


def test_exact_integer_power_is_correctly_rounded():
    for n in list(range(0, 400)) + [799, 800, 801, 1599, 1600, 1601, 65535]:
        assert ox.lib().ox_pow_int_exact(n, 20) == float(n ** 20)
        assert ox.lib().ox_pow_int_exact(n, 3) == float(n ** 3)


def test_hardcoded_opponent_episodes_match_reference(golden_dir):
    """OneStepLookahead / Random evaluation games (general/hardcoded_players.py) incl. the reference's frame quirk."""
    with open(os.path.join(golden_dir, "episodes_vs_hardcoded.json")) as f:
        eps = json.load(f)
    assert len(eps) >= 16
    for e in eps:
        cfg = ox.make_cfg(e["game"], e["sims"], seed=e["seed"], game_uid=e["game_uid"], evaluate=True)
        o = ox.play_episode_vs(cfg, e["swap"], e["kind"], net_seed=e["net_seed"])
        assert o["reward"] == e["reward"]
        assert [[m["tree"], m["ply"], m["action"]] for m in o["moves"]] == e["moves"]
        assert o["final_state"].tolist() == e["final_state"]


def test_threaded_search_matches_reference_under_the_cooperative_schedule(golden_dir):
    """MCTreeSearch(thread_count=K) behind an InferenceProxy (mcts.py:328-331: virtual loss, per-child locks, the "all states
    in use" return): the unmodified reference run with its K threads forced into the cooperative round-robin interleaving
    (oracle/ref_harness._coop_search -> tests/golden/threaded.json) vs the C restatement of that schedule, bit for bit."""
    with open(os.path.join(golden_dir, "threaded.json")) as f:
        g = json.load(f)
    assert len(g["searches"]) >= 10 and len(g["episodes"]) >= 6
    for c in g["searches"]:
        noise = np.array([[[unhex(x) for x in c["noise"]]] * 64], np.float64)
        cfg = ox.make_cfg(c["game"], c["sims"], seed=c["seed"], game_uid=c["game_uid"], noise_table=noise, strong_play=c["strong_play"],
                          threads=c["threads"])
        t = ox.Tree(cfg, hash_seed=c["net_seed"])
        t.reset(1)
        for a, _pl in c["prefix"]:
            t.play_action(a)
        t.search()
        o = t.root_stats()
        assert o["n"].tolist() == c["n"], c["name"]
        assert o["w"].tolist() == [unhex(x) for x in c["w"]], c["name"]
        assert o["root_n"] == c["root_n"] and o["root_w"] == unhex(c["root_w"]) and o["q"] == unhex(c["q"]), c["name"]
    for e in g["episodes"]:
        table = np.array([[[unhex(x) for x in row] for row in t] for t in e["noise_table"]], np.float64)
        cfg = ox.make_cfg(e["game"], e["sims"], seed=e["seed"], game_uid=e["game_uid"], evaluate=e["evaluate"], noise_table=table,
                          threads=e["threads"])
        o = ox.play_episode(cfg, e["swap"], net_seed=e["net_seed"], net_seed_opp=e["net_seed_opp"])
        assert o["reward"] == e["reward"] and len(o["moves"]) == len(e["moves"])
        for a, b in zip(o["moves"], e["moves"]):
            assert (a["tree"], a["ply"], a["action"], a["root_n"]) == (b["tree"], b["ply"], b["action"], b["root_n"])
            assert a["n"] == b["n"] and a["w"] == [unhex(x) for x in b["w"]] and a["root_w"] == unhex(b["root_w"])
        assert len(o["records"]) == len(e["records"])
        for a, b in zip(o["records"], e["records"]):
            assert a["state"].tolist() == b["state"] and a["actual_val"] == b["actual_val"] and a["q"] == np.float32(unhex(b["q"]))
            got, want = a["tree_probs"], np.array([unhex(x) for x in b["tree_probs"]], np.float32)
            if e["evaluate"]:
                assert np.all(np.abs(got - want) <= np.spacing(np.maximum(np.abs(want), np.float32(1e-30))))
            else:
                assert np.array_equal(got, want)
        assert o["final_state"].tolist() == e["final_state"]
