"""Edge cases of the C ABI on the GPU: empty calls, exhausted node pools, full record rings, a finite games_target, one game.
The engine must degrade loudly (counters / error codes), never corrupt memory or hang."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator, SelfPlayEngine
    game = kw.pop("game", 0)
    return SelfPlayEngine(game, kw.pop("n_games", 8), kw.pop("sims", 30), HashNetEvaluator(game, kw.get("seed", 0)), noise_mode=0, **kw)


def test_empty_calls_are_no_ops():
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200._lib import check, lib
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    z = torch.zeros(1, dtype=torch.int64, device="cuda")
    check(lib().spx_env_step(0, 0, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), st))
    check(lib().spx_env_valid_moves(1, 0, z.data_ptr(), z.data_ptr(), st))
    torch.manual_seed(0)
    tw = nets.NativeTower(nets.ResidualTower(7, 6, 7, num_blocks=1).eval())
    p, v = torch.full((1, 7), -1.0, device="cuda"), torch.full((1,), -1.0, device="cuda")
    check(lib().spx_tower_forward(tw._h, z.data_ptr(), z.data_ptr(), None, 0, p.data_ptr(), v.data_ptr(), st))
    # a batch in which nothing asks for an evaluation launches but writes nothing
    need = torch.zeros(5, dtype=torch.uint8, device="cuda")
    own = torch.zeros(5, dtype=torch.int64, device="cuda")
    p5, v5 = torch.full((5, 7), -1.0, device="cuda"), torch.full((5,), -1.0, device="cuda")
    tw.forward_bits(own, own.clone(), needs_eval=need, policy=p5, value=v5)
    torch.cuda.synchronize()
    assert bool((p5 == -1).all()) and bool((v5 == -1).all()) and bool((p == -1).all())
    tw.close()
    assert lib().spx_env_step(7, 4, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), st) < 0   # unknown game
    assert b"unknown game" in lib().spx_last_error()


def test_node_pool_exhaustion_is_counted_and_parks_the_game():
    e = _engine(n_games=4, sims=200, nodes_per_tree=64, games_target=4)   # 200 sims need ~200 nodes per move: cannot fit
    e.run_ticks(3000)
    c = e.counters()
    assert c["errors"] >= 4 and c["nodes_allocated"] <= 4 * 2 * 64 and e.all_idle()
    assert c["games_finished"] == 0 and len(e.drain_results()) == 0
    e.close()


def test_record_ring_overflow_is_counted_not_written():
    e = _engine(n_games=16, sims=20, games_target=64, record_capacity=100)
    e.run_ticks(6000)
    c = e.counters()
    recs = e.drain_records()
    assert e.all_idle() and c["games_finished"] == 64
    assert len(recs) == 100 and c["records_dropped"] > 0
    assert c["records_dropped"] + 100 == c["moves"]          # every move of a finished game is either kept or counted as dropped
    e.close()


def test_overflowing_rings_fail_loudly_on_the_host():
    """Lost training records / game results must not vanish silently (run_tasks would hang in task_queue.join())."""
    from self_play_reinforcement_learning_b200._lib import SpxError
    e = _engine(n_games=16, sims=20, games_target=64, record_capacity=100, result_capacity=10)
    e.run_ticks(6000)
    with pytest.raises(SpxError, match="dropped"):
        e.check_overflow()
    with pytest.raises(SpxError, match="results were lost"):
        e.drain_results()
    assert len(e.drain_results()) == 0                       # the ring was reset: later drains work again
    e.close()
    e = _engine(game=1, n_games=8, sims=2, games_target=200)      # tiny searches: many games per slot between two drains
    res = 0
    while not e.all_idle():
        e.run_ticks(e.safe_poll_interval)
        res += len(e.drain_results())
        e.drain_records()
    assert res == 200 and e.check_overflow()["records_dropped"] == 0
    e.close()


@pytest.mark.parametrize("game", [0, 1])
def test_single_game_with_finite_target_matches_oracle(game):
    e = _engine(game=game, n_games=1, sims=40, games_target=3, seed=9, move_log=True)
    logs_seen = []
    while not e.all_idle():
        e.run_ticks(64)
    recs, res = H.split_by_game(e.drain_records(), e.drain_results())
    assert sorted(res) == [0, 1, 2] and e.counters()["games_finished"] == 3
    for g in range(3):
        o = H.oracle_episode(game, 40, 9, g, None, net_seed=9)
        assert res[g]["reward"] == o["reward"] and res[g]["plies"] == len(o["moves"]) and len(recs[g]) == len(o["records"])
    e.run_ticks(10)                                           # idle engine: further ticks do nothing
    assert e.counters()["games_finished"] == 3 and len(e.drain_results()) == 0
    e.close()


def test_zero_target_engine_is_idle_from_the_start():
    e = _engine(n_games=4, sims=10, games_target=0)
    e.run_ticks(5)
    assert e.all_idle() and e.counters()["sims"] == 0
    e.close()
