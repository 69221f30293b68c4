"""CPU, only where /root/reference exists (this container; skipped on the GPU box): differential fuzz of the C oracle
against the LIVE unmodified reference under the injected-RNG harness -- more seeds and shapes than the committed goldens."""
import numpy as np
import pytest

from oracle import oracle as ox
from oracle import ref_harness as rh
from oracle import spec

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present (GPU box)")


@pytest.mark.parametrize("game", [0, 1])
def test_env_fuzz_against_live_reference(game):
    rng = np.random.default_rng(99 + game)
    W, H, A = spec.GAME_DIMS[game]
    actions = rng.integers(0, A, size=(150, 46 if game == 0 else 12)).astype(np.int32)
    fp = rng.choice([-1, 1], size=len(actions)).astype(np.int8)
    want = rh.ref_env_playout(game, actions, fp)
    got = ox.env_playout(game, actions, fp)
    for k in ("status", "reward", "done", "boards", "valid"):
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.parametrize("case", range(6))
def test_episode_fuzz_against_live_reference(case):
    rng = np.random.default_rng(500 + case)
    game = case % 2
    A = spec.GAME_DIMS[game][2]
    sims = int(rng.integers(20, 90))
    swap, evaluate, strong = bool(case & 1), case in (2, 5), case == 4
    uid = 2 * int(rng.integers(0, 1000)) + int(swap)
    table = rng.dirichlet([float(rng.choice([0.15, 1.0, 3.0]))] * A, size=(2, 22))
    seed = int(rng.integers(0, 2**31))
    r = rh.run_episode(game, sims, seed=seed, game_uid=uid, swap_sides=swap, evaluate=evaluate, noise_table=table,
                       net_seed=case, net_seed_opp=(case + 50) if evaluate else None, strong_play=strong)
    cfg = ox.make_cfg(game, sims, seed=seed, game_uid=uid, evaluate=evaluate, strong_play=strong, noise_table=table)
    o = ox.play_episode(cfg, swap, net_seed=case, net_seed_opp=(case + 50) if evaluate else None)
    assert r["reward"] == o["reward"] and len(r["moves"]) == len(o["moves"])
    for a, b in zip(r["moves"], o["moves"]):
        assert (a["tree"], a["ply"], a["action"], a["root_n"], a["root_w"]) == (b["tree"], b["ply"], b["action"], b["root_n"], b["root_w"])
        assert a["n"] == list(b["n"]) and a["w"] == list(b["w"])
    for a, b in zip(r["records"], o["records"]):
        assert np.array_equal(a["state"], b["state"]) and a["q"] == b["q"] and a["actual_val"] == b["actual_val"]
        if evaluate:
            assert np.all(np.abs(a["tree_probs"] - b["tree_probs"]) <= np.spacing(np.maximum(b["tree_probs"], np.float32(1e-30))))
        else:
            assert np.array_equal(a["tree_probs"], b["tree_probs"])


@pytest.mark.parametrize("case", range(6))
def test_threaded_episode_fuzz_against_live_reference(case):
    """thread_count = K: the unmodified reference's K search threads forced into the cooperative round-robin interleaving
    (ref_harness._coop_search) vs the oracle's restatement of that schedule, on random shapes."""
    rng = np.random.default_rng(900 + case)
    game = case % 2
    A = spec.GAME_DIMS[game][2]
    sims, K = int(rng.integers(16, 70)), int(rng.choice([2, 3, 4, 6]))
    swap, evaluate = bool(case & 1), case in (2, 5)
    uid = 2 * int(rng.integers(0, 1000)) + int(swap)
    table = rng.dirichlet([float(rng.choice([0.15, 1.0, 3.0]))] * A, size=(2, 22))
    seed = int(rng.integers(0, 2**31))
    r = rh.run_episode(game, sims, seed=seed, game_uid=uid, swap_sides=swap, evaluate=evaluate, noise_table=table,
                       net_seed=case, net_seed_opp=(case + 50) if evaluate else None, threads=K)
    cfg = ox.make_cfg(game, sims, seed=seed, game_uid=uid, evaluate=evaluate, noise_table=table, threads=K)
    o = ox.play_episode(cfg, swap, net_seed=case, net_seed_opp=(case + 50) if evaluate else None)
    assert r["reward"] == o["reward"] and len(r["moves"]) == len(o["moves"])
    for a, b in zip(r["moves"], o["moves"]):
        assert (a["tree"], a["ply"], a["action"], a["root_n"], a["root_w"]) == (b["tree"], b["ply"], b["action"], b["root_n"], b["root_w"])
        assert a["n"] == list(b["n"]) and a["w"] == list(b["w"])
    for a, b in zip(r["records"], o["records"]):
        assert np.array_equal(a["state"], b["state"]) and a["q"] == b["q"] and a["actual_val"] == b["actual_val"]
        if evaluate:
            assert np.all(np.abs(a["tree_probs"] - b["tree_probs"]) <= np.spacing(np.maximum(b["tree_probs"], np.float32(1e-30))))
        else:
            assert np.array_equal(a["tree_probs"], b["tree_probs"])
