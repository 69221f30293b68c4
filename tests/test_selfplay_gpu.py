"""The public host API (selfplay.BatchedSelfPlay / run_tasks, nets.TorchNetEvaluator) on the GPU: reference-format
outputs, the scheduler task protocol, config-1 (TicTacToe, repo's tictactoe net, 1 game, 100 sims) and config-4
(head-to-head evaluation with two nets) shapes, each replayed bit-exactly through the oracle."""
import queue

import numpy as np
import pytest
import torch

from oracle import spec
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_run_tasks_protocol_and_move_format():
    from self_play_reinforcement_learning_b200.selfplay import Move, run_tasks

    class TaskQ:
        done = 0

        def task_done(self):
            self.done += 1
    tasks = [{"play": {"swap_sides": bool(i % 2), "update": True}} for i in range(7)]  # scheduler: swap_sides = i odd
    rq, mq, tq = queue.Queue(), queue.Queue(), TaskQ()
    moves, results = run_tasks(None, 0, tasks, result_queue=rq, memory_queue=mq, task_queue=tq, iterations=60, net="hash",
                               seed=5, noise_mode=0)
    assert tq.done == 7 and rq.qsize() == 7 and len(results) == 7
    assert sum(r["swap_sides"] for r in results) == 3 and all(r["reward"] in (-1, 0, 1) for r in results)
    assert mq.qsize() == len(moves) > 7 * 7
    m = moves[0]
    assert isinstance(m, Move) and m._fields == ("state", "actual_val", "tree_probs", "q")          # mcts.py:17
    assert m.state.dtype == torch.int64 and tuple(m.state.shape) == (7, 6)                          # mcts.py:284
    assert m.tree_probs.dtype == torch.float32 and tuple(m.tree_probs.shape) == (7,)                # mcts.py:286
    assert m.q.dtype == torch.float32 and m.q.dim() == 0 and m.actual_val.dtype == torch.float32    # mcts.py:287,230
    # game 0 (index 0, no swap) must be the oracle's game 0
    o = H.oracle_episode(0, 60, 5, 0, None, net_seed=5)
    first = [mv for mv in moves[:len(o["records"])]]
    for a, b in zip(first, o["records"]):
        assert np.array_equal(a.state.numpy(), b["state"]) and np.array_equal(a.tree_probs.numpy(), b["tree_probs"])
        assert float(a.q) == float(b["q"]) and float(a.actual_val) == b["actual_val"]
    # evaluation tasks: no records (update False), two nets
    tasks = [{"play": {"swap_sides": bool(i % 2), "update": False}, "evaluate": True} for i in range(4)]
    moves, results = run_tasks(None, 0, tasks, iterations=40, net="hash", seed=6, noise_mode=0, evaluation_network=object())
    assert moves == [] and len(results) == 4


def test_config1_tictactoe_repo_net_one_game_100_sims():
    """BASELINE.json configs[0] shape on the GPU path: TicTacToe, ConvNetTicTacToe, 1 game, 100 sims/move (generic
    nn.Module boundary through torch, fp32), replayed through the oracle."""
    from self_play_reinforcement_learning_b200 import envs, nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(1)
    net = nets.ConvNetTicTacToe(3, 3, 9).eval()
    table = np.random.default_rng(1).dirichlet([1.0] * 9, size=(1, 2, 22))
    sp = BatchedSelfPlay(net, env=envs.TicTacToeEnv, n_games=1, sims=100, net="torch", seed=2, games_target=1, noise_mode=1,
                         move_log=True, net_dtype=torch.float32)
    sp.engine.set_noise_table(table)
    logs = H.run_logged(sp.engine)
    recs, res = H.split_by_game(sp.engine.drain_records(), sp.engine.drain_results())
    o = H.replay_in_oracle(1, 100, 2, 0, table[0], logs[0])
    H.compare_game(1, sp.engine.move_log(0), recs[0], res[0], o)
    assert 5 <= res[0]["plies"] <= 9
    # the logged network outputs are the module's own fp32 outputs (value tolerance 1e-5, fp32 path)
    own, opp = logs[0][0]["own"][0], logs[0][0]["opp"][0]
    board = torch.from_numpy(spec.bits_to_board(int(own), int(opp), 1))[None]
    with torch.no_grad():
        p, v = net.cuda().float().forward(board.cuda())
    assert np.allclose(p.cpu().numpy()[0], logs[0][0]["policy"][0], atol=1e-5) and abs(float(v) - logs[0][0]["value"][0]) < 1e-5
    sp.close()


def test_config4_shape_two_nets_evaluate_mode():
    """elo.py head-to-head shape: two different random-init nets, evaluate mode (temp/20), no records; every game is
    replayed through the oracle with the logged outputs of both nets."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    a = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    torch.manual_seed(1)
    b = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    n_games, sims = 6, 40
    sp = BatchedSelfPlay(a, game=0, n_games=n_games, sims=sims, net="torch", evaluation_network=b, evaluate=True, update=False,
                         seed=3, games_target=n_games, noise_mode=0, move_log=True, net_dtype=torch.float32)
    logs = H.run_logged(sp.engine)
    recs, res = H.split_by_game(sp.engine.drain_records(), sp.engine.drain_results())
    assert recs == {} and len(res) == n_games
    for g in range(n_games):
        o = H.replay_in_oracle(0, sims, 3, g, None, logs[g], evaluate=True)
        assert res[g]["reward"] == o["reward"] and res[g]["plies"] == len(o["moves"])
        ml = sp.engine.move_log(g)
        assert [m["action"] for m in ml] == [m["action"] for m in o["moves"]] and [m["n"] for m in ml] == [list(m["n"]) for m in o["moves"]]
    sp.close()


def test_weight_refresh_changes_outputs():
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    a = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    torch.manual_seed(9)
    b = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    sp = BatchedSelfPlay(a, game=0, n_games=8, sims=20, net="tower", seed=1)
    sp.engine.run_ticks(3)
    torch.cuda.synchronize()
    p_a = sp.engine.policy.clone()
    blob_b = sp.packed_weights_pinned(b)
    out = sp.play_step(0, weights_host=blob_b)
    assert out["h2d_bytes"] == blob_b.numel()
    sp.evaluator(sp.engine)
    torch.cuda.synchronize()
    assert not torch.equal(p_a, sp.engine.policy)
    sp.close()


def test_scheduler_train_loop_small():
    """configs[4] shape at toy size: initial games -> epoch of self-play -> SGD updates -> evaluation, one GPU."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.scheduler import SelfPlayScheduler
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).cuda().eval()
    before = net.linear_output.weight.detach().clone()
    s = SelfPlayScheduler(net, 0, iterations=30, epoch_length=16, initial_games=8, evaluation_games=6, games_per_gpu=16, batch_size=32,
                          updates_per_epoch=5, lr=0.01)
    assert s.trainer_kind == "device"        # the native SGD step (csrc/spx_train.cu) is the default for the reference's tower
    hist = s.train_model(num_epochs=1)
    assert len(hist) == 1 and hist[0]["memory"] > 16 * 7 and np.isfinite(hist[0]["loss"])
    assert not torch.equal(before, net.linear_output.weight.detach())
    assert int(net.bn1.num_batches_tracked) == 5 and net.conv1.weight.dtype == torch.float32 and not net.training
    total, bd = s.compare_models()
    assert set(bd) == {"first", "second"} and sum(sum(v.values()) for v in bd.values()) == 16


def test_scheduler_with_the_torch_evaluator_keeps_fp32_master_weights(tmp_path):
    """A tower the native kernel is not built for (filter_factor != 32) runs through TorchNetEvaluator.  The evaluator casts a
    PRIVATE copy to bf16: the caller's module -- the one the scheduler trains (fp32, no amp) and checkpoints -- stays fp32."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.scheduler import SelfPlayScheduler
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1, filter_factor=8).cuda().eval()
    with pytest.raises(ValueError):
        nets.pack_tower_blob(net)
    before = net.linear_output.weight.detach().clone()
    s = SelfPlayScheduler(net, 0, iterations=20, epoch_length=12, initial_games=6, evaluation_games=4, games_per_gpu=8, batch_size=16,
                          updates_per_epoch=3, lr=0.01, net="torch", save_dir=str(tmp_path), save_memory=False)
    hist = s.train_model(num_epochs=1)
    assert np.isfinite(hist[0]["loss"]) and not torch.equal(before, net.linear_output.weight.detach())
    assert all(p.dtype == torch.float32 for p in net.parameters()) and all(b.dtype in (torch.float32, torch.int64) for b in net.buffers())
    with torch.no_grad():
        p, v = net.forward(torch.zeros(2, 7, 6, dtype=torch.int64))       # the fp32 forward still works
    assert p.dtype == torch.float32 and tuple(p.shape) == (2, 7)
    assert all(t.dtype != torch.bfloat16 for t in torch.load(hist[0]["saved_model"])["model"].values())


@pytest.mark.parametrize("replay", ["device", "host"])
def test_scheduler_deduplicate_option(replay):
    """UpdateWorker(deduplicate=True) (updateworker.py:88-89): after the epoch's records are in, the memory holds one averaged
    record per distinct position -- every game starts from the empty board, so duplicates are guaranteed -- and training still runs."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.scheduler import SelfPlayScheduler
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).cuda().eval()
    s = SelfPlayScheduler(net, 0, iterations=20, epoch_length=24, initial_games=8, evaluation_games=0, games_per_gpu=16, batch_size=16,
                          updates_per_epoch=3, lr=0.01, replay=replay, deduplicate=True, amp=torch.bfloat16 if replay == "device" else None)
    hist = s.train_model(num_epochs=1)
    assert np.isfinite(hist[0]["loss"])
    if replay == "device":
        recs = s.memory.read()
        keys = set(zip(recs["own"].tolist(), recs["opp"].tolist()))
        assert len(keys) == len(recs) == s.memory.unique_states and int(recs["pad1"].sum()) > len(recs)     # counts: merged records
        empty = recs[(recs["own"] == 0) & (recs["opp"] == 0)]
        assert len(empty) == 1 and int(empty["pad1"][0]) >= 16                                            # every game's first position
    else:
        states = [m.state.numpy().tobytes() for m in s.memory._buffer]
        assert len(set(states)) == len(states) > 0


@pytest.mark.parametrize("replay", ["device", "host"])
def test_scheduler_checkpoints_and_resume(tmp_path, replay):
    """save_dir / resume_model / resume_memory (self_play_parallel.py:213-267, updateworker.py:54-58,111-139, base_worker.py:26-62):
    an epoch leaves model-<time>:<games> ({"model": state_dict}) and memory-<time>:<size> (pickled Memory of Move tuples) in the
    run folder; a second run picks both up."""
    import glob
    import os
    import time
    from self_play_reinforcement_learning_b200 import checkpoint, nets
    from self_play_reinforcement_learning_b200.scheduler import SelfPlayScheduler
    kw = dict(iterations=20, epoch_length=12, initial_games=4, evaluation_games=4, games_per_gpu=8, batch_size=16, updates_per_epoch=2,
              lr=0.01, replay=replay, save_dir=str(tmp_path))
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).cuda().eval()
    s = SelfPlayScheduler(net, 0, **kw)
    hist = s.train_model(num_epochs=2)
    run = os.path.join(str(tmp_path), s.start_time)
    models, mems = sorted(glob.glob(os.path.join(run, "model-*"))), glob.glob(os.path.join(run, "memory-*"))
    assert [m.rsplit(":", 1)[1] for m in models] == ["12", "24"] and hist[1]["saved_model"] == models[1]
    assert len(mems) == 1 and mems[0].endswith(":" + str(hist[1]["memory"]))          # the previous memory file was removed
    assert set(torch.load(models[1]).keys()) == {"model"} and hist[1]["lr"] == 0.01
    saved_moves = list(checkpoint.load_memory(mems[0])._buffer)
    assert len(saved_moves) == hist[1]["memory"] and saved_moves[0].state.dtype == torch.int64 and tuple(saved_moves[0].state.shape) == (7, 6)
    time.sleep(0.01)
    torch.manual_seed(5)
    net2 = nets.ResidualTower(7, 6, 7, num_blocks=1).cuda().eval()
    s2 = SelfPlayScheduler(net2, 0, **kw)
    s2.resume(resume_model=True, resume_memory=True)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    assert len(s2.memory) == len(saved_moves)
    got = s2.memory.to_moves() if replay == "device" else list(s2.memory._buffer)
    for a, b in zip(got[:50], saved_moves[:50]):
        assert torch.equal(a.state.cpu(), b.state) and torch.equal(a.tree_probs.cpu(), b.tree_probs) and float(a.q) == float(b.q)
    s3 = SelfPlayScheduler(net2, 0, **{**kw, "save_dir": str(tmp_path / "nothing_here")})
    s3.resume(resume_model=True, resume_memory=True)                                   # no earlier run: a no-op


class _StreamRandomOpponent:
    """A host-side BasePlayer (general/base_model.py:10-29) that plays the spec-stream random move -- the same rule the
    oracle's OPP_RANDOM uses, so a facade-vs-host game can be compared with ox.play_episode_vs move for move."""

    def __init__(self, game, seed):
        self.game, self.seed = game, seed

    def reset(self, player=None, game_index=0):
        W, Hh, _ = spec.GAME_DIMS[self.game]
        self.board = np.zeros((W, Hh), np.int64)
        self.game_index = game_index

    def valid(self):
        return (np.abs(self.board).sum(1) < self.board.shape[1]) if self.game == 0 else (self.board.reshape(-1) == 0)

    def __call__(self, s=None):
        moves = np.flatnonzero(self.valid())
        ply = int(np.abs(self.board).sum())
        u = spec.rng_uniform(self.seed, self.game_index, 1, spec.PURPOSE_OPPONENT, ply, 0, 0, 0)
        return int(moves[min(int(u * len(moves)), len(moves) - 1)])

    def play_action(self, a, player):
        if self.game == 0:
            self.board[a, int(np.abs(self.board[a]).sum())] = player
        else:
            self.board[a // 3, a % 3] = player


@pytest.mark.parametrize("swap", [False, True])
def test_policy_facade_drives_reference_style_episode(swap):
    """The per-game Policy facade (mcts.MCTreeSearch) inside a SelfPlayer-style loop (selfplayworker.py:172-224) against a
    host-side opponent; the whole game must equal the oracle's play_episode_vs with the same random opponent."""
    import queue as pyqueue
    from oracle import oracle as ox
    from self_play_reinforcement_learning_b200 import envs
    from self_play_reinforcement_learning_b200.mcts import MCTreeSearch
    sims, seed = 60, 17
    memq = pyqueue.Queue()
    policy = MCTreeSearch(None, envs.Connect4Env, memory_queue=memq, iterations=sims, net="hash", seed=seed, noise_mode=0)
    policy.evaluate(True)
    opp = _StreamRandomOpponent(0, seed)
    policy.reset(player=-1 if swap else 1)
    opp.reset(game_index=policy.game_index)
    gi = policy.game_index
    assert (gi & 1) == int(swap)
    actions, r, done, player = [], 0, False, (-1 if swap else 1)
    env_pieces = np.zeros((7, 6), np.int64)
    while not done:
        a = policy(None) if player == 1 else opp(None)
        policy.play_action(a, player)            # selfplayworker.py:221-224
        opp.play_action(a, player * -1)
        env_pieces[a, int(np.abs(env_pieces[a]).sum())] = player
        actions.append(a)
        cfg = ox.make_cfg(0, sims, seed=seed, game_uid=gi, evaluate=True)
        done = len(actions) == len(ox.play_episode_vs(cfg, swap, spec.OPP_RANDOM, net_seed=seed)["moves"])
        player = -player
    o = ox.play_episode_vs(ox.make_cfg(0, sims, seed=seed, game_uid=gi, evaluate=True), swap, spec.OPP_RANDOM, net_seed=seed)
    assert actions == [m["action"] for m in o["moves"]]
    policy.push_to_queue(done=True, r=o["reward"])
    n_own = sum(1 for m in o["moves"] if m["tree"] == 0)
    assert memq.qsize() == n_own
    m = memq.get()
    assert float(m.actual_val) == float(o["reward"]) and m.state.dtype == torch.int64 and tuple(m.tree_probs.shape) == (7,)
    # a second episode on the same object works (new game index, engine restarted)
    policy.reset(player=1)
    assert policy.game_index == 4 and isinstance(policy(None), int)
    policy.close()


def test_elo_round_robin_with_native_towers_and_hardcoded_anchor():
    """elo.py flow: register models, compare all pairs (two native towers head to head; tower vs the device-side random
    player), accumulate under the reference's key convention, fit ratings with the anchor at 0."""
    from self_play_reinforcement_learning_b200 import elo, nets
    db = elo.ModelDatabase("connect4")
    db.add_model("random", "random")
    for i, name in enumerate(("neta", "netb")):
        torch.manual_seed(i)
        db.add_model(name, nets.ResidualTower(7, 6, 7, num_blocks=1).eval())
    e = elo.Elo(db, iterations=40, seed=3)
    e.compare_all()
    e.compare_models("neta", "netb", num_games=7)          # odd number: accumulates, exactly 7 more games
    assert set(db.result_shelf) == {"random__neta", "random__netb", "netb__neta"}
    assert sum(db.result_shelf["random__neta"].values()) == 100 and sum(db.result_shelf["netb__neta"].values()) == 107
    assert db.result_shelf["random__neta"]["losses"] > 60      # a 40-sim search beats the random player (losses: random's view)
    ratings = e.calculate_elo()
    assert ratings["random"] == 0 and ratings["neta"] > 100 and ratings["netb"] > 100 and db.elos()["elo"] == ratings
