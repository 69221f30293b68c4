"""Oracle parity at BASELINE.json's FULL sizes, through the default product path.

* configs[1]: 1024 concurrent games x 800 sims/move, ResidualTower-20, the fused tick kernel (spx_tick_fused): one launch per tick
  with every network output of 64 sampled slots logged on the device, their first TWO games replayed through the C oracle
  (records, results, per-move root statistics bit for bit: "visit counts bit-exact given identical network outputs"); then the
  same 2048 games again with 100 ticks per launch, and once more with the evaluation cache -> identical records and results for ALL games.
* configs[2] shard: 2048 games on one GPU (two board-group units per SM pair), same check on 32 sampled slots.
* configs[3]: 4096 head-to-head games x 400 sims, evaluate mode, two different native towers.
* the same three shapes with the hash network run to completion against oracle_episode (no logging needed): a 32-bit index bug
  in the 5-22 GB node pools would show here.
"""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _sample_slots(n_games, k, seed):
    rng = np.random.default_rng(seed)
    s = set(rng.choice(n_games, size=k - 4, replace=False).tolist()) | {0, 1, n_games - 2, n_games - 1}
    return sorted(s)


def _sorted(recs, res):
    return np.sort(recs, order=["game_index", "tree", "ply"]), np.sort(res, order=["game_index"])


def _play_all(engine, chunk):
    recs, res = [], []
    while True:
        engine.run_ticks(400, fused=True, chunk=chunk)
        recs.append(engine.drain_records()); res.append(engine.drain_results())
        if engine.all_idle():
            break
    engine.check_overflow()
    return _sorted(np.concatenate(recs), np.concatenate(res))


@pytest.mark.parametrize("n_games,games_per_slot,n_sample", [(1024, 2, 64), (2048, 1, 32)])
def test_fused_tower20_800_sims_full_size_replays_in_oracle(n_games, games_per_slot, n_sample):
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.engine import SelfPlayEngine
    sims, seed = 800, 17
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
    total = n_games * games_per_slot
    table = np.random.default_rng(5).dirichlet([1.0] * 7, size=(total, 2, 22))
    slots = _sample_slots(n_games, n_sample, 1)

    def make():
        e = SelfPlayEngine(game=0, n_games=n_games, sims=sims, evaluator=nets.TowerEvaluator(net), seed=seed, noise_mode=1,
                           games_target=total, move_log=True)
        e.set_noise_table(table)
        return e
    e = make()
    assert e.device_bytes() > (5.4e9 if n_games == 1024 else 10.8e9)        # the full-size node pool
    recs_parts, res_parts = [], []

    def step():
        e.run_ticks(1, fused=True, chunk=1)
    games = H.run_logged_device(e, slots, step, poll=2048)
    recs1, res1 = _sorted(e.drain_records(), e.drain_results())
    c1 = e.check_overflow()
    assert c1["errors"] == 0 and c1["games_finished"] == total and len(res1) == total
    by_rec, by_res = H.split_by_game(recs1, res1)
    for slot in slots:
        assert len(games[slot]) == games_per_slot
        for k, log in enumerate(games[slot]):
            gi = slot + k * n_games
            o = H.replay_in_oracle(0, sims, seed, gi, table[gi], log)
            if k == games_per_slot - 1:
                H.compare_game(0, e.move_log(slot), by_rec[gi], by_res[gi], o)
            else:
                H.compare_records_and_result(0, by_rec[gi], by_res[gi], o)
    mean_path = c1["path_len_sum"] / c1["sims"]
    assert 2.0 < mean_path < 12.0 and c1["sims"] >= c1["moves"] * sims
    e.close()
    # the same games with 100 ticks per launch (the bench's form): every record and result of every game is identical
    e = make()
    recs2, res2 = _play_all(e, 100)
    c2 = e.check_overflow()
    e.close()
    assert recs2.tobytes() == recs1.tobytes() and res2.tobytes() == res1.tobytes()
    for k in ("sims", "leaf_evals", "terminal_sims", "path_len_sum", "moves", "games_finished", "nodes_allocated", "errors"):
        assert c1[k] == c2[k], k
    # ... and with the evaluation cache (spx_config.eval_cache_log2 = 12): about half of the requests answered from the slots'
    # tables, every record and result of every game still identical
    e = SelfPlayEngine(game=0, n_games=n_games, sims=sims, evaluator=nets.TowerEvaluator(net), seed=seed, noise_mode=1,
                       games_target=total, eval_cache=True, max_sims_per_tick=16)
    e.set_noise_table(table)
    recs3, res3 = _play_all(e, 100)
    c3 = e.check_overflow()
    e.close()
    assert recs3.tobytes() == recs1.tobytes() and res3.tobytes() == res1.tobytes()
    for k in ("sims", "terminal_sims", "path_len_sum", "moves", "games_finished", "nodes_allocated", "errors"):
        assert c1[k] == c3[k], k
    assert c3["leaf_evals"] + c3["cache_hits"] == c1["leaf_evals"] and c3["cache_hits"] > 0.35 * c1["leaf_evals"]


def test_config4_4096_games_400_sims_two_towers_evaluate_mode():
    """BASELINE.json configs[3] at full size: two different random-init towers, evaluate mode (temp 1/20, no records), leaves
    partitioned by owning net on the device; 24 sampled games replayed through the oracle with the logged outputs."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    n_games, sims, seed = 4096, 400, 23
    torch.manual_seed(0)
    a = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
    torch.manual_seed(1)
    b = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
    sp = BatchedSelfPlay(a, game=0, n_games=n_games, sims=sims, net="tower", evaluation_network=b, evaluate=True, update=False, seed=seed,
                         games_target=n_games, noise_mode=0, move_log=True)
    e = sp.engine
    slots = _sample_slots(n_games, 24, 2)
    games = H.run_logged_device(e, slots, e.tick, poll=2048)
    _, res = H.split_by_game(e.drain_records(), e.drain_results())
    c = e.check_overflow()
    assert len(res) == n_games and c["errors"] == 0
    for slot in slots:
        assert len(games[slot]) == 1
        o = H.replay_in_oracle(0, sims, seed, slot, None, games[slot][0], evaluate=True)
        assert res[slot]["reward"] == o["reward"] and res[slot]["plies"] == len(o["moves"])
        ml = e.move_log(slot)
        assert [m["action"] for m in ml] == [m["action"] for m in o["moves"]]
        assert [m["n"] for m in ml] == [list(m["n"]) for m in o["moves"]] and [m["w"] for m in ml] == [list(m["w"]) for m in o["moves"]]
    firsts = [r["reward"] for g, r in res.items() if g % 2 == 0]
    assert len(firsts) == n_games // 2
    sp.close()


@pytest.mark.parametrize("n_games,sims,evaluate,two", [(1024, 800, False, False), (2048, 800, False, False), (4096, 400, True, True)])
def test_full_size_hash_net_games_equal_oracle_episodes(n_games, sims, evaluate, two):
    """Every slot plays one whole game at the full configuration size (node pools of 5.4 / 10.8 / 11 GB); sampled games -- the
    first, the last, and random ones -- equal the oracle's episodes move for move."""
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator, SelfPlayEngine
    seed = 31
    table = np.random.default_rng(9).dirichlet([1.0] * 7, size=(n_games, 2, 22))
    e = SelfPlayEngine(game=0, n_games=n_games, sims=sims, evaluator=HashNetEvaluator(0, 4, 5 if two else None), seed=seed, noise_mode=1,
                       games_target=n_games, move_log=True, evaluate=evaluate, two_nets=two, emit_records=not evaluate)
    e.set_noise_table(table)
    e.run_until_idle(max_ticks=200_000, poll_every=2048)
    recs, res = H.split_by_game(e.drain_records(), e.drain_results())
    c = e.check_overflow()
    assert c["errors"] == 0 and c["games_finished"] == n_games
    for g in _sample_slots(n_games, 20, 3):
        o = H.oracle_episode(0, sims, seed, g, table[g], evaluate=evaluate, net_seed=4, net_seed_opp=5 if two else None)
        H.compare_game(0, e.move_log(g), None if evaluate else recs[g], res[g], o, evaluate=evaluate)
    e.close()
