"""The fused tick kernel (spx_tick_fused: the network CTAs also advance the games whose leaves they evaluate, n ticks per launch)
against the same ticks as separate spx_advance + spx_tower_forward launches: same counters, records, results, trees and
network outputs, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(net, game, n_games, sims, ticks, fused, games_target, chunk, **kw):
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    sp = BatchedSelfPlay(net, game=game, n_games=n_games, sims=sims, net="tower", seed=3, games_target=games_target, **kw)
    e = sp.engine
    e.run_ticks(ticks, fused=fused, chunk=chunk)
    torch.cuda.synchronize()
    recs, res = e.drain_records(), e.drain_results()
    out = dict(counters=e.counters(), recs=np.sort(recs, order=["game_index", "tree", "ply"]), res=np.sort(res, order=["game_index"]),
               stats=[e.root_stats(t) for t in (0, 1)], policy=e.policy.cpu().numpy().copy(), value=e.value.cpu().numpy().copy(),
               needs=e.needs_eval.cpu().numpy().copy(), leaf=e.leaf_own.cpu().numpy().copy())
    sp.close()
    return out


def _same(a, b):
    assert a["counters"] == b["counters"]
    assert a["recs"].tobytes() == b["recs"].tobytes() and a["res"].tobytes() == b["res"].tobytes()
    for sa, sb in zip(a["stats"], b["stats"]):
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), k
    assert np.array_equal(a["needs"], b["needs"]) and np.array_equal(a["leaf"], b["leaf"])
    live = a["needs"].astype(bool)
    assert np.array_equal(a["policy"][live], b["policy"][live]) and np.array_equal(a["value"][live], b["value"][live])


@pytest.mark.parametrize("game,n_games,sims,ticks,target", [(0, 50, 40, 2600, 120), (1, 23, 30, 900, 60), (0, 16, 25, 400, None)])
def test_fused_ticks_equal_separate_launches(game, n_games, sims, ticks, target):
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(1)
    net = (nets.ResidualTower(7, 6, 7, num_blocks=2) if game == 0 else nets.ResidualTower(3, 3, 9, num_blocks=2)).eval()
    a = _run(net, game, n_games, sims, ticks, False, target, 64)
    b = _run(net, game, n_games, sims, ticks, True, target, 97)         # chunks that do not divide the tick count
    assert a["counters"]["ticks"] == ticks and a["counters"]["errors"] == 0 and a["counters"]["games_finished"] > 0
    _same(a, b)
    if target is not None:
        assert a["counters"]["games_finished"] == target                # every slot ended idle: the all-idle skip path ran


def test_fused_ticks_more_groups_than_sm_pairs():
    """1100 games = 158 board groups = 79 units for 74 SM pairs: some clusters work on two units per tick."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(2)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    a = _run(net, 0, 1100, 12, 150, False, None, 64)
    b = _run(net, 0, 1100, 12, 150, True, None, 50)
    _same(a, b)


def test_fused_then_separate_then_fused_continues_the_same_games():
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(4)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    outs = []
    for plan in ([(300, False)], [(100, True), (100, False), (100, True)]):
        sp = BatchedSelfPlay(net, game=0, n_games=30, sims=20, net="tower", seed=9)
        for n, fused in plan:
            sp.engine.run_ticks(n, fused=fused, chunk=33)
        torch.cuda.synchronize()
        outs.append((sp.engine.counters(), np.sort(sp.engine.drain_records(), order=["game_index", "tree", "ply"]).tobytes()))
        sp.close()
    assert outs[0] == outs[1]


@pytest.mark.parametrize("kw", [dict(opponent="lookahead", evaluate=True, update=False), dict(opponent="random"), dict(strong_play=True, alpha=0.15)])
def test_fused_ticks_other_engine_modes(kw):
    """Hard-coded opponents (hardcoded_players.py), evaluate mode (temp 1/20), strong_play, another Dirichlet alpha: the same
    state machine runs inside the fused kernel."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(6)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    a = _run(net, 0, 20, 30, 700, False, 40, 64, **kw)
    b = _run(net, 0, 20, 30, 700, True, 40, 41, **kw)
    assert a["counters"]["games_finished"] > 0 and a["counters"]["errors"] == 0
    _same(a, b)
