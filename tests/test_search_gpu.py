"""Engine (spx_advance + evaluator through the C ABI) vs golden reference vectors and vs the C oracle.
Visit counts, fp64 value sums, actions, records and results must be bit-exact given identical network
outputs, Dirichlet noise and tie-breaking (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import oracle as ox
from oracle import spec
from tests import helpers as H

pytestmark = pytest.mark.gpu


class UniformEvaluator:
    """policy == 1/A, value == 0 (SURVEY.md Appendix B hooks)."""

    def __call__(self, e):
        e.policy.fill_(1.0 / e.A)
        e.value.zero_()


def _engine(**kw):
    from self_play_reinforcement_learning_b200.engine import SelfPlayEngine
    return SelfPlayEngine(**kw)


def test_golden_first_searches():
    """search.json cases without a play_action prefix == the first search of game 0 (swap_sides False)."""
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    cases = [c for c in H.load_json("search.json") if not c["params"].get("prefix")]
    assert len(cases) >= 8
    for c in cases:
        p, game = c["params"], c["game"]
        A = spec.GAME_DIMS[game][2]
        ev = UniformEvaluator() if c["net_kind"] == "uniform" else HashNetEvaluator(game, p.get("net_seed", 0))
        uid = p.get("game_uid", 0)
        if uid & 1:
            continue  # an odd game index means swap_sides: not comparable with a root-player +1 search
        noise = p.get("noise")
        e = _engine(game=game, n_games=1, sims=c["sims"], evaluator=ev, seed=p.get("seed", 0), tie_mode=p.get("tie_mode", 1),
                    noise_mode=0 if noise is None else 1, move_log=True, slot_offset=uid, slot_stride=2,
                    games_target=uid + 1, strong_play=p.get("strong_play", False), max_sims_per_tick=4)
        if noise is not None:
            e.set_noise_table(np.tile(np.asarray(noise, np.float64), (1, 2, 22, 1)), first_game_index=uid)
        for _ in range(c["sims"] + 8):
            e.tick()
            ml = e.move_log(0)
            if ml:
                break
        m = ml[0]
        assert m["n"] == c["n"], c["name"]
        assert m["w"] == [H.unhex(x) for x in c["w"]], c["name"]
        assert m["root_n"] == c["root_n"] and m["root_w"] == H.unhex(c["root_w"]), c["name"]
        e.close()


def test_golden_episodes():
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    eps = H.load_json("episodes.json")
    for ep in eps:
        game, uid = ep["game"], ep["game_uid"]
        table = np.array([[[H.unhex(x) for x in row] for row in t] for t in ep["noise_table"]], np.float64)[None]
        ev = HashNetEvaluator(game, ep["net_seed"], ep["net_seed_opp"])
        e = _engine(game=game, n_games=1, sims=ep["sims"], evaluator=ev, seed=ep["seed"], noise_mode=1, move_log=True,
                    evaluate=ep["evaluate"], strong_play=ep["strong_play"], two_nets=ep["net_seed_opp"] is not None,
                    slot_offset=uid, slot_stride=2, games_target=uid + 1, max_sims_per_tick=3)
        e.set_noise_table(table, first_game_index=uid)
        e.run_until_idle(max_ticks=200000, poll_every=256)
        moves = e.move_log(0)
        recs, res = H.split_by_game(e.drain_records(), e.drain_results())
        want = dict(reward=ep["reward"],
                    moves=[dict(tree=m["tree"], ply=m["ply"], action=m["action"], n=m["n"], w=[H.unhex(x) for x in m["w"]],
                                root_n=m["root_n"], root_w=H.unhex(m["root_w"])) for m in ep["moves"]],
                    records=[dict(state=np.array(r["state"], np.int8), actual_val=r["actual_val"],
                                  tree_probs=np.array([H.unhex(x) for x in r["tree_probs"]], np.float32),
                                  q=np.float32(H.unhex(r["q"]))) for r in ep["records"]])
        if ep["evaluate"]:
            # reference n**20 goes through numpy's pow (<= 1 ulp from correctly rounded); compare probs to 1 f32 ulp
            got_r = recs[uid]
            for a, b in zip(got_r, want["records"]):
                assert np.all(np.abs(a["tree_probs"][:len(b["tree_probs"])] - b["tree_probs"]) <= np.spacing(np.maximum(b["tree_probs"], np.float32(1e-30))))
                b["tree_probs"] = a["tree_probs"][:len(b["tree_probs"])].copy()
        H.compare_game(game, moves, recs[uid], res[uid], want)
        assert res[uid]["swap"] == int(ep["swap"])
        assert e.counters()["errors"] == 0
        e.close()


@pytest.mark.parametrize("game,n_games,sims,evaluate,two,strong", [(0, 192, 120, False, False, False), (1, 256, 60, False, False, False),
                                                                   (0, 64, 64, True, True, False), (1, 64, 40, True, True, False),
                                                                   (0, 8, 800, False, False, False), (0, 96, 100, False, False, True),
                                                                   (1, 40, 50, True, True, True)])
def test_many_games_vs_oracle(game, n_games, sims, evaluate, two, strong):
    """Two generations of games per slot, injected Dirichlet tables, hash nets: every game must equal the oracle."""
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    A = spec.GAME_DIMS[game][2]
    rng = np.random.default_rng(1234 + n_games)
    total = 2 * n_games
    table = rng.dirichlet([0.6] * A, size=(total, 2, 22))
    ev = HashNetEvaluator(game, 5, 9 if two else None)
    e = _engine(game=game, n_games=n_games, sims=sims, evaluator=ev, seed=99, noise_mode=1, evaluate=evaluate, two_nets=two,
                games_target=total, max_sims_per_tick=8, move_log=False, strong_play=strong)
    e.set_noise_table(table, first_game_index=0)
    e.run_until_idle(max_ticks=400000, poll_every=512)
    c = e.counters()
    assert c["errors"] == 0 and c["games_finished"] == total and c["records_dropped"] == 0
    recs, res = H.split_by_game(e.drain_records(), e.drain_results())
    assert len(res) == total
    sims_total = 0
    for gi in range(total):
        o = H.oracle_episode(game, sims, 99, gi, table[gi], evaluate=evaluate, net_seed=5, net_seed_opp=9 if two else None, strong_play=strong)
        sims_total += o["sims"]
        assert res[gi]["reward"] == o["reward"] and res[gi]["plies"] == len(o["moves"]), gi
        got = recs[gi]
        assert len(got) == len(o["records"]), gi
        for a, b in zip(got, o["records"]):
            assert np.array_equal(H.record_board(a, game), b["state"]), gi
            assert np.array_equal(a["tree_probs"][:A], b["tree_probs"]), gi
            assert a["q"] == b["q"] and a["actual_val"] == b["actual_val"], gi
            assert (a["tree"], a["ply"]) == (b["tree"], b["ply"])
    assert c["sims"] == sims_total and c["moves"] == sum(r["plies"] for r in res.values())
    e.close()


def test_device_generated_noise_roundtrip():
    """noise_mode 2 (Gamma variates drawn on the GPU): read the noise back through the move log, inject it
    into the oracle and require identical games."""
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    for game, alpha in [(0, 1.0), (0, 0.15), (1, 2.5)]:
        A = spec.GAME_DIMS[game][2]
        n_games, sims = 16, 50
        e = _engine(game=game, n_games=n_games, sims=sims, evaluator=HashNetEvaluator(game, 1), seed=7, noise_mode=2, alpha=alpha,
                    games_target=n_games, move_log=True)
        e.run_until_idle(max_ticks=100000, poll_every=256)
        recs, res = H.split_by_game(e.drain_records(), e.drain_results())
        for g in range(n_games):
            ml = e.move_log(g)
            table = np.full((2, 22, A), 1.0 / A)
            cnt = [0, 0]
            for m in ml:
                table[m["tree"], cnt[m["tree"]]] = m["noise"]
                cnt[m["tree"]] += 1
                s = sum(m["noise"])
                assert abs(s - 1.0) < 1e-12 and min(m["noise"]) >= 0.0
            o = H.oracle_episode(game, sims, 7, g, table, net_seed=1)
            H.compare_game(game, ml, recs[g], res[g], o)
        e.close()


def test_full_size_invariants():
    """Config-2 shape (1024 games x 800 sims, hash net): size-independent properties."""
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    e = _engine(game=0, n_games=1024, sims=800, evaluator=HashNetEvaluator(0, 3), seed=1, noise_mode=2, max_sims_per_tick=8)
    e.run_ticks(2500)
    c = e.counters()
    assert c["errors"] == 0 and c["records_dropped"] == 0
    assert c["sims"] >= 1024 * 2400 and c["moves"] >= 1024 * 2
    assert c["leaf_evals"] + c["terminal_sims"] >= c["sims"]
    for tree in (0, 1):
        rs = e.root_stats(tree)
        diff = rs["root_n"] - rs["n"].sum(axis=1)
        assert set(np.unique(diff)).issubset({0, 1})       # node.n == own expansion visit + children visits
        assert (rs["n"][~rs["valid"]] == 0).all()           # illegal moves are never visited
        assert (rs["n"] >= 0).all() and (np.abs(rs["w"]) <= rs["n"] + 1e-9).all()
    recs = e.drain_records()
    A = 7
    assert np.allclose(recs["tree_probs"][:, :A].sum(axis=1), 1.0, atol=1e-5)
    assert set(np.unique(recs["actual_val"])).issubset({-1.0, 0.0, 1.0})
    e.close()


@pytest.mark.parametrize("game,kind", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_hardcoded_opponents_vs_oracle(game, kind):
    """Evaluation games against OneStepLookahead / Random (hardcoded_players.py): every game equals the oracle
    (which is pinned against the live reference in tests/test_oracle_vs_reference_live.py and the goldens)."""
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    n_games, sims = 48, 40
    e = _engine(game=game, n_games=n_games, sims=sims, evaluator=HashNetEvaluator(game, 4), seed=21, noise_mode=0, evaluate=True,
                emit_records=False, opponent_kind=kind, games_target=n_games, move_log=True)
    e.run_until_idle(max_ticks=200000, poll_every=256)
    _, res = H.split_by_game(e.drain_records(), e.drain_results())
    assert len(res) == n_games and e.counters()["errors"] == 0
    outcomes = set()
    for g in range(n_games):
        cfg = ox.make_cfg(game, sims, seed=21, game_uid=g, evaluate=True)
        o = ox.play_episode_vs(cfg, bool(g & 1), kind, net_seed=4)
        ml = e.move_log(g)
        assert res[g]["reward"] == o["reward"] and res[g]["plies"] == len(o["moves"]), g
        assert [(m["tree"], m["ply"], m["action"]) for m in ml] == [(m["tree"], m["ply"], m["action"]) for m in o["moves"]], g
        for a, b in zip(ml, o["moves"]):
            if a["tree"] == 0:
                assert a["n"] == list(b["n"]) and a["w"] == list(b["w"]) and a["root_n"] == b["root_n"]
        outcomes.add(o["reward"])
    assert len(outcomes) >= 2
    e.close()
