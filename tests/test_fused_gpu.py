"""The fused tick kernel (spx_tick_fused: the persistent network CTAs also advance the games whose leaves they evaluate; terminal
re-visit chains run on a shadow warp under the network pass) against separate spx_advance + spx_tower_forward launches.

The two forms schedule simulations differently (the fused kernel attempts one simulation per game between two passes and lets the
shadow warp work ahead), so they are not compared tick by tick: every game is a deterministic function of its index -- its RNG
streams are keyed by (seed, game, tree, ply, sim, depth) and the network is bitwise independent of batch position -- and must
come out IDENTICAL: records, results, per-move root statistics, and the sums of all per-game counters."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GAME_COUNTERS = ("sims", "leaf_evals", "terminal_sims", "path_len_sum", "moves", "games_finished", "nodes_allocated", "errors",
                 "records_dropped")


def _finish(e, fused, chunk, plan=(), balanced=None):
    for n, f in plan:
        e.run_ticks(n, fused=f, chunk=chunk, balanced=balanced)
    for _ in range(100000):
        if e.all_idle():
            break
        e.run_ticks(150, fused=fused, chunk=chunk, balanced=balanced)
    assert e.all_idle()


def _run(net, game, n_games, sims, fused, games_target, chunk, plan=(), balanced=None, **kw):
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    sp = BatchedSelfPlay(net, game=game, n_games=n_games, sims=sims, net="tower", seed=3, games_target=games_target, move_log=True, **kw)
    e = sp.engine
    _finish(e, fused, chunk, plan, balanced)
    torch.cuda.synchronize()
    recs, res = e.drain_records(), e.drain_results()
    out = dict(counters=e.counters(), recs=np.sort(recs, order=["game_index", "tree", "ply"]), res=np.sort(res, order=["game_index"]),
               stats=[e.root_stats(t) for t in (0, 1)], logs=[e.move_log(g) for g in range(0, n_games, max(1, n_games // 16))])
    sp.close()
    return out


def _same(a, b):
    for k in GAME_COUNTERS:
        assert a["counters"][k] == b["counters"][k], (k, a["counters"], b["counters"])
    assert a["recs"].tobytes() == b["recs"].tobytes() and a["res"].tobytes() == b["res"].tobytes()
    for sa, sb in zip(a["stats"], b["stats"]):
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), k
    assert a["logs"] == b["logs"]


@pytest.mark.parametrize("game,n_games,sims,target", [(0, 50, 40, 120), (1, 23, 30, 60), (0, 16, 25, 16), (0, 7, 60, 30)])
def test_fused_games_equal_separate_launches(game, n_games, sims, target):
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(1)
    net = (nets.ResidualTower(7, 6, 7, num_blocks=2) if game == 0 else nets.ResidualTower(3, 3, 9, num_blocks=2)).eval()
    a = _run(net, game, n_games, sims, False, target, 64)
    b = _run(net, game, n_games, sims, True, target, 97)
    c = _run(net, game, n_games, sims, True, target, 1)          # one tick per launch: every hand-over crosses a launch boundary
    assert a["counters"]["errors"] == 0 and a["counters"]["games_finished"] == target and len(a["res"]) == target
    _same(a, b)
    _same(a, c)


@pytest.mark.parametrize("n_games,sims,target,blocks", [(50, 40, 120, 2), (1100, 12, 1100, 1), (2300, 12, 2300, 1)])
def test_exact_and_work_conserving_launches_play_the_same_games(n_games, sims, target, blocks):
    """spx_tick_fused (every game exactly n ticks per launch) and spx_tick_fused_balanced (SM pairs draw their ticks from the
    launch's budget of passes: the default of run_ticks for launches of >= 16 ticks) in fast mode, mixed and shadow-all mode."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(7)
    net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    a = _run(net, 0, n_games, sims, True, target, 40, balanced=False)
    b = _run(net, 0, n_games, sims, True, target, 40, balanced=True)
    c = _run(net, 0, n_games, sims, True, target, 17, balanced=True)
    assert a["counters"]["games_finished"] == target and a["counters"]["errors"] == 0
    _same(a, b)
    _same(a, c)


def test_fused_ticks_more_groups_than_sm_pairs():
    """2300 games = 329 board groups = 165 units for 74 SM pairs: clusters own two or three units each (shadow-all mode: the
    shadow warp advances a unit's games while the tensor pipe evaluates the cluster's other units)."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(2)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    a = _run(net, 0, 2300, 12, False, 2300, 64)
    b = _run(net, 0, 2300, 12, True, 2300, 50)
    assert a["counters"]["games_finished"] == 2300
    _same(a, b)


def test_fused_ticks_mixed_one_and_two_units_per_cluster():
    """1100 games = 158 groups = 79 units: five clusters own two units (shadow-all mode), the others one (fast mode)."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(2)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    a = _run(net, 0, 1100, 12, False, 1100, 64)
    b = _run(net, 0, 1100, 12, True, 1100, 50)
    _same(a, b)


def test_fused_then_separate_then_fused_continues_the_same_games():
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(4)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    a = _run(net, 0, 30, 20, False, 90, 33)
    b = _run(net, 0, 30, 20, True, 90, 33, plan=[(100, True), (77, False), (100, True), (3, False), (1, True), (50, False)])
    _same(a, b)


@pytest.mark.parametrize("kw", [dict(opponent="lookahead", evaluate=True, update=False), dict(opponent="random"), dict(strong_play=True, alpha=0.15)])
def test_fused_ticks_other_engine_modes(kw):
    """Hard-coded opponents (hardcoded_players.py), evaluate mode (temp 1/20), strong_play, another Dirichlet alpha: the same
    state machine runs inside the fused kernel."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(6)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    a = _run(net, 0, 20, 30, False, 40, 64, **kw)
    b = _run(net, 0, 20, 30, True, 40, 41, **kw)
    assert a["counters"]["games_finished"] == 40 and a["counters"]["errors"] == 0
    _same(a, b)


@pytest.mark.parametrize("game,n_games,sims,target,blocks,kw", [
    (0, 50, 60, 150, 2, {}),                                             # fast mode, three games per slot (cross-game reuse of the openings)
    (1, 23, 40, 69, 2, {}),                                              # TicTacToe: 9 priors per entry
    (0, 1100, 12, 2200, 1, {}),                                          # mixed fast / shadow-all mode
    (0, 30, 50, 60, 1, dict(opponent="lookahead", evaluate=True, update=False)),
    (0, 20, 40, 40, 1, dict(strong_play=True, alpha=0.15)),
])
def test_evaluation_cache_plays_the_same_games_with_fewer_network_evaluations(game, n_games, sims, target, blocks, kw):
    """spx_config.eval_cache_log2: requests for positions the slot has evaluated before are answered from its table.  The network
    is a pure function with batch-position-independent output, so everything but the number of network evaluations is identical."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(11)
    net = (nets.ResidualTower(7, 6, 7, num_blocks=blocks) if game == 0 else nets.ResidualTower(3, 3, 9, num_blocks=blocks)).eval()
    a = _run(net, game, n_games, sims, True, target, 40, **kw)
    b = _run(net, game, n_games, sims, True, target, 40, eval_cache=True, **kw)
    c = _run(net, game, n_games, sims, True, target, 23, eval_cache=6, balanced=False, **kw)   # 64 entries: constant eviction
    assert a["counters"]["games_finished"] == target and a["counters"]["errors"] == 0 and a["counters"]["cache_hits"] == 0
    for x in (b, c):
        assert x["counters"]["cache_hits"] > 0
        assert x["counters"]["leaf_evals"] + x["counters"]["cache_hits"] == a["counters"]["leaf_evals"]
        x["counters"]["leaf_evals"] = a["counters"]["leaf_evals"]
        _same(a, x)
    assert b["counters"]["cache_hits"] > c["counters"]["cache_hits"]
    if not kw:   # two searching trees per game: 18 % at 12 sims per move, about half at 800 (one tree against a hard-coded opponent at 50 sims: 3 %)
        assert b["counters"]["cache_hits"] > 0.1 * a["counters"]["leaf_evals"]


def test_evaluation_cache_is_keyed_by_the_weights_version():
    """New weights (spx_tower_load) must never be answered with evaluations of the old ones: play with net A, load net B into the
    same tower, and the games that follow equal an engine that only ever saw net B."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(12)
    net_a = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    net_b = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    outs = []
    for first in (net_a, net_b):
        sp = BatchedSelfPlay(first, game=0, n_games=12, sims=40, net="tower", seed=3, games_target=24, eval_cache=True)
        e = sp.engine
        if first is net_a:   # fill the tables with net A's evaluations, then start over with net B's weights
            e.run_ticks(400, chunk=50)
            sp.load_weights(net_b)
            e.reset()
            e.drain_records(); e.drain_results()
        _finish(e, True, 50)
        outs.append((np.sort(e.drain_records(), order=["game_index", "tree", "ply"]).tobytes(), np.sort(e.drain_results(), order=["game_index"]).tobytes()))
        sp.close()
    assert outs[0] == outs[1]


def test_evaluation_cache_with_separate_launches_and_with_two_networks():
    """spx_advance + a forward per tick (spx_set_eval_cache_versions tells the engine whose outputs it consumes), and head-to-head
    evaluation with two native towers (entries are keyed by network id + weights version: a position evaluated by network 0
    must not answer network 1)."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(13)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    other = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    a = _run(net, 0, 40, 50, False, 80, 64)
    b = _run(net, 0, 40, 50, False, 80, 64, eval_cache=True)
    c = _run(net, 0, 40, 50, True, 80, 64, plan=[(90, True), (60, False), (40, True)], eval_cache=10)   # fused and separate launches share the table
    for x in (b, c):
        assert x["counters"]["cache_hits"] > 0.1 * a["counters"]["leaf_evals"]
        assert x["counters"]["leaf_evals"] + x["counters"]["cache_hits"] == a["counters"]["leaf_evals"]
        x["counters"]["leaf_evals"] = a["counters"]["leaf_evals"]
        _same(a, x)
    kw = dict(evaluation_network=other, evaluate=True, update=False)
    a = _run(net, 0, 40, 50, False, 80, 64, **kw)
    b = _run(net, 0, 40, 50, False, 80, 64, eval_cache=True, **kw)
    assert b["counters"]["cache_hits"] > 0.01 * a["counters"]["leaf_evals"]   # 3 % at 50 sims in evaluate mode: each network only sees its own tree
    assert b["counters"]["leaf_evals"] + b["counters"]["cache_hits"] == a["counters"]["leaf_evals"]
    b["counters"]["leaf_evals"] = a["counters"]["leaf_evals"]
    _same(a, b)
