"""The threaded search on the device (spx_config.search_threads = K: MCTreeSearch(thread_count=K) behind an InferenceProxy,
mcts.py:132,328-331 -- virtual loss, per-child locks, the "all states in use" return) under the cooperative round-robin schedule:
vs golden vectors written by the UNMODIFIED reference forced into that schedule (tests/golden/threaded.json) and vs the C oracle
on many games.  Bit-exact: visit counts, fp64 value sums, actions, records, results."""
import numpy as np
import pytest
import torch

from oracle import spec
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from self_play_reinforcement_learning_b200.engine import SelfPlayEngine
    return SelfPlayEngine(**kw)


def test_threaded_golden_first_searches():
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    cases = [c for c in H.load_json("threaded.json")["searches"] if not c["prefix"] and not c["game_uid"] & 1]
    assert len(cases) >= 4
    for c in cases:
        game, uid = c["game"], c["game_uid"]
        noise = [H.unhex(x) for x in c["noise"]]
        e = _engine(game=game, n_games=1, sims=c["sims"], evaluator=HashNetEvaluator(game, c["net_seed"]), seed=c["seed"], noise_mode=1,
                    move_log=True, slot_offset=uid, slot_stride=2, games_target=uid + 1, strong_play=c["strong_play"],
                    search_threads=c["threads"])
        assert e.n_leaves == c["threads"]
        e.set_noise_table(np.tile(np.asarray(noise, np.float64), (1, 2, 22, 1)), first_game_index=uid)
        ml = []
        for _ in range(c["sims"] + 8):
            e.tick()
            ml = e.move_log(0)
            if ml:
                break
        m = ml[0]
        assert m["n"] == c["n"], (c["name"], m["n"], c["n"])
        assert m["w"] == [H.unhex(x) for x in c["w"]], c["name"]
        assert m["root_n"] == c["root_n"] and m["root_w"] == H.unhex(c["root_w"]), c["name"]
        assert e.counters()["errors"] == 0
        e.close()


def test_threaded_golden_episodes():
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    eps = H.load_json("threaded.json")["episodes"]
    for ep in eps:
        game, uid = ep["game"], ep["game_uid"]
        table = np.array([[[H.unhex(x) for x in row] for row in t] for t in ep["noise_table"]], np.float64)[None]
        ev = HashNetEvaluator(game, ep["net_seed"], ep["net_seed_opp"])
        e = _engine(game=game, n_games=1, sims=ep["sims"], evaluator=ev, seed=ep["seed"], noise_mode=1, move_log=True,
                    evaluate=ep["evaluate"], two_nets=ep["net_seed_opp"] is not None, slot_offset=uid, slot_stride=2,
                    games_target=uid + 1, search_threads=ep["threads"])
        e.set_noise_table(table, first_game_index=uid)
        e.run_until_idle(max_ticks=200000, poll_every=256)
        moves = e.move_log(0)
        recs, res = H.split_by_game(e.drain_records(), e.drain_results())
        want = dict(reward=ep["reward"],
                    moves=[dict(tree=m["tree"], ply=m["ply"], action=m["action"], n=m["n"], w=[H.unhex(x) for x in m["w"]],
                                root_n=m["root_n"], root_w=H.unhex(m["root_w"])) for m in ep["moves"]],
                    records=[dict(state=np.array(r["state"], np.int8), actual_val=r["actual_val"],
                                  tree_probs=np.array([H.unhex(x) for x in r["tree_probs"]], np.float32), q=np.float32(H.unhex(r["q"])))
                             for r in ep["records"]])
        if ep["evaluate"]:      # n**20 is exact-integer on the device: <= 1 f32 ulp from numpy's pow; compare through the oracle below
            want["records"] = None
        H.compare_game(game, moves, recs[uid] if want["records"] is not None else None, res[uid], want)
        assert e.counters()["errors"] == 0
        e.close()


@pytest.mark.parametrize("game,K,sims,n_games", [(0, 4, 48, 96), (0, 2, 33, 64), (1, 4, 40, 64), (0, 8, 100, 32)])
def test_threaded_many_games_vs_oracle(game, K, sims, n_games):
    """Two games per slot, hash network, Dirichlet tables injected: every game bit-exact against the oracle's threaded schedule."""
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    A = spec.GAME_DIMS[game][2]
    total = 2 * n_games
    table = np.random.default_rng(K + sims).dirichlet([1.0] * A, size=(total, 2, 22))
    e = _engine(game=game, n_games=n_games, sims=sims, evaluator=HashNetEvaluator(game, 3), seed=9, noise_mode=1, games_target=total,
                search_threads=K)
    e.set_noise_table(table)
    recs, res = [], []
    while True:
        e.run_ticks(128)
        recs.append(e.drain_records()); res.append(e.drain_results())
        if e.all_idle():
            break
    c = e.check_overflow()
    assert c["errors"] == 0 and c["games_finished"] == total
    by_rec, by_res = H.split_by_game(np.concatenate(recs), np.concatenate(res))
    for gi in range(total):
        o = H.oracle_episode(game, sims, 9, gi, table[gi], net_seed=3, threads=K)
        H.compare_records_and_result(game, by_rec[gi], by_res[gi], o)
    e.close()


def test_threaded_with_the_tower_network():
    """K = 4 workers per tree driven by the native tower (2 blocks): G x K leaves per tick through spx_advance + spx_tower_forward;
    every network output logged per leaf slot and the games replayed through the oracle's threaded schedule."""
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(1)
    net = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    game, n_games, sims, K = 0, 5, 24, 4
    table = np.random.default_rng(2).dirichlet([1.0] * 7, size=(n_games, 2, 22))
    e = _engine(game=game, n_games=n_games, sims=sims, evaluator=nets.TowerEvaluator(net), seed=4, noise_mode=1, games_target=n_games,
                search_threads=K)
    e.set_noise_table(table)
    # per game: evaluations in the order the oracle asks for them = tick by tick, leaf slots in worker order
    logs = [[dict(own=[], opp=[], policy=[], value=[]) for _ in (0, 1)] for _ in range(n_games)]
    tree_of = {}
    for _ in range(200000):
        e.advance()
        torch.cuda.synchronize()
        need = e.needs_eval.cpu().numpy().astype(bool)
        if not need.any() and e.all_idle():
            break
        own, opp, nid = e.leaf_own.cpu().numpy().view(np.uint64), e.leaf_opp.cpu().numpy().view(np.uint64), e.net_id.cpu().numpy()
        e.evaluator(e)
        torch.cuda.synchronize()
        pol, val = e.policy.cpu().numpy(), e.value.cpu().numpy()
        st = torch.zeros(n_games, dtype=torch.int32, device=e.device)
        import ctypes as C
        from self_play_reinforcement_learning_b200 import _lib
        _lib.check(_lib.lib().spx_pending_tree(e._h, st.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "spx_pending_tree")
        tree = st.cpu().numpy()
        for l in np.flatnonzero(need):
            g = l // K
            L = logs[g][int(tree[g])]
            L["own"].append(own[l]); L["opp"].append(opp[l]); L["policy"].append(pol[l].copy()); L["value"].append(val[l])
    recs, res = H.split_by_game(e.drain_records(), e.drain_results())
    for g in range(n_games):
        import ctypes as C
        from oracle import oracle as ox
        rs = ox.make_replay(game, logs[g])
        cfg = ox.make_cfg(game, sims, seed=4, game_uid=g, noise_table=table[g], threads=K)
        pair = (ox.fn_addr("ox_replaynet"), C.addressof(rs))
        o = ox.play_episode(cfg, bool(g & 1), nets=(pair, pair))
        assert rs.mismatches == 0 and rs.overruns == 0 and rs.cursor[0] == rs.n[0] and rs.cursor[1] == rs.n[1], g
        H.compare_records_and_result(game, recs[g], res[g], o)
    assert e.counters()["errors"] == 0
    e.close()


@pytest.mark.parametrize("game,kind", [(0, 1), (1, 2)])
def test_threaded_search_against_hardcoded_opponents(game, kind):
    """Evaluation games (evaluate mode, no records) of a thread_count = 4 policy against OneStepLookahead / Random: move for move
    and root statistics against the oracle."""
    from oracle import oracle as ox
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator
    n_games, sims, K = 32, 36, 4
    e = _engine(game=game, n_games=n_games, sims=sims, evaluator=HashNetEvaluator(game, 4), seed=21, noise_mode=0, evaluate=True,
                emit_records=False, opponent_kind=kind, games_target=n_games, move_log=True, search_threads=K)
    e.run_until_idle(max_ticks=200000, poll_every=256)
    _, res = H.split_by_game(e.drain_records(), e.drain_results())
    assert len(res) == n_games and e.counters()["errors"] == 0
    for g in range(n_games):
        cfg = ox.make_cfg(game, sims, seed=21, game_uid=g, evaluate=True, threads=K)
        o = ox.play_episode_vs(cfg, bool(g & 1), kind, net_seed=4)
        ml = e.move_log(g)
        assert res[g]["reward"] == o["reward"] and res[g]["plies"] == len(o["moves"]), g
        assert [(m["tree"], m["ply"], m["action"]) for m in ml] == [(m["tree"], m["ply"], m["action"]) for m in o["moves"]], g
        for a, b in zip(ml, o["moves"]):
            if a["tree"] == 0:
                assert a["n"] == list(b["n"]) and a["w"] == list(b["w"]) and a["root_n"] == b["root_n"]
    e.close()


def test_scheduler_epoch_with_four_search_threads():
    """The full loop (self-play -> native SGD step -> evaluation) with thread_count = 4 searches, toy sizes."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.scheduler import SelfPlayScheduler
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).cuda().eval()
    s = SelfPlayScheduler(net, 0, iterations=32, epoch_length=16, initial_games=8, evaluation_games=6, games_per_gpu=16, batch_size=32,
                          updates_per_epoch=3, lr=0.01, search_threads=4, evaluation_opponent="lookahead")
    hist = s.train_model(num_epochs=1)
    assert len(hist) == 1 and hist[0]["memory"] > 16 * 7 and np.isfinite(hist[0]["loss"]) and s.trainer_kind == "device"
