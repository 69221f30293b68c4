"""The hand-written tcgen05 tower (spx_tower_forward through the C ABI) vs the fp32 torch forward of the
same ResidualTower.  Floating point: tolerance stated per assertion (bf16 weights/activations, fp32 accumulate)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# Absolute tolerance on the softmax policy / tanh value against the TRUE fp32 torch forward (TF32 off), 16-bit weights and
# activations, fp32 accumulation, 20 residual blocks.  Asserted = 1.5 x the largest deviation measured on 4096 positions
# (scripts/dbg_tower_err.py, profiles/r2_tower_accuracy.txt):
#   fp16 (default; the reference's own GPU arithmetic is fp16 autocast, inference_worker.py:117):
#       random-init net of the bench 7.1e-4 / 1.0e-3; nets with randomised BatchNorm statistics 3.0e-3 / 3.7e-3
#   bf16 (SPX_TOWER_DTYPE=bf16, 3.5 % faster): 6.0e-3 / 8.6e-3 and 2.6e-2 / 3.3e-2
# (torch's own .half() / .bfloat16() forward of the same module deviates MORE: 1.1e-3 / 1.4e-3 and 6.7e-3 / 1.1e-2.)
TOL = {"f16": dict(policy=4.5e-3, value=5.5e-3, policy_init=1.1e-3, value_init=1.5e-3),
       "bf16": dict(policy=3.9e-2, value=4.9e-2, policy_init=9.0e-3, value_init=1.3e-2)}


@pytest.fixture(params=["f16", "bf16"])
def dtype(request, monkeypatch):
    monkeypatch.setenv("SPX_TOWER_DTYPE", request.param)
    return request.param


def _random_positions(n, seed):
    """Reachable-looking Connect4 positions: random legal playouts of random length (net frame)."""
    rng = np.random.default_rng(seed)
    boards = np.zeros((n, 7, 6), np.int64)
    for i in range(n):
        h = np.zeros(7, int)
        player = 1
        for _ in range(rng.integers(0, 40)):
            legal = np.flatnonzero(h < 6)
            if not len(legal):
                break
            c = rng.choice(legal)
            boards[i, c, h[c]] = player
            h[c] += 1
            player = -player
    return boards


def _randomise_bn(net):
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.7, 1.3)
                m.bias.uniform_(-0.1, 0.1)


def test_bench_network_within_1e3_of_fp32(dtype):
    """The network bench.py runs (ResidualTower-20, torch.manual_seed(0), random init) on 4096 reachable positions."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.envs import boards_to_bits
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
    boards = torch.from_numpy(_random_positions(4096, 5))
    bits = boards_to_bits(boards.cuda(), 0)
    tw = nets.NativeTower(net)
    assert tw.f16 == (dtype == "f16")
    p, v = tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        pr, vr = net.cuda().float().forward(boards.cuda())
    dp, dv = (p - pr).abs().max().item(), (v - vr.reshape(-1)).abs().max().item()
    print(f"bench net {dtype}: max|dpolicy|={dp:.3e} max|dvalue|={dv:.3e}")
    assert dp < TOL[dtype]["policy_init"] and dv < TOL[dtype]["value_init"]
    tw.close()


@pytest.mark.parametrize("blocks,n", [(1, 7), (2, 50), (20, 300), (20, 1024), (20, 2500)])
def test_tower_matches_fp32_reference(blocks, n, dtype):
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.envs import boards_to_bits
    torch.manual_seed(blocks)
    net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    _randomise_bn(net)  # non-trivial BN statistics so that the folding is really exercised
    boards = torch.from_numpy(_random_positions(n, 5 + blocks))
    bits = boards_to_bits(boards.cuda(), 0)
    tw = nets.NativeTower(net)
    p, v = tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False          # the reference must be true fp32 (cuDNN convolutions default to TF32)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        pr, vr = net.cuda().float().forward(boards.cuda())
    dp = (p - pr).abs().max().item()
    dv = (v - vr.reshape(-1)).abs().max().item()
    print(f"{dtype} blocks={blocks} n={n} max|dpolicy|={dp:.3e} max|dvalue|={dv:.3e}")
    assert torch.allclose(p.sum(1), torch.ones(n, device="cuda"), atol=1e-5)
    assert dp < TOL[dtype]["policy"] and dv < TOL[dtype]["value"]
    # row independence: a board's outputs do not depend on its batch position or its neighbours
    perm = torch.randperm(n, device="cuda")
    p2, v2 = tw.forward_bits(bits[perm, 0].contiguous(), bits[perm, 1].contiguous())
    assert torch.equal(p2, p[perm]) and torch.equal(v2, v[perm])
    tw.close()


def test_selfplay_with_tower_replays_bit_exact_in_oracle():
    """Self-play driven by the native tower: log every evaluation, replay it through the C oracle and require
    identical games ("visit counts bit-exact given identical network outputs")."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.engine import SelfPlayEngine
    from tests import helpers as H
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    n_games, sims = 14, 40
    ev = nets.TowerEvaluator(net)
    table = np.random.default_rng(3).dirichlet([1.0] * 7, size=(n_games, 2, 22))
    e = SelfPlayEngine(game=0, n_games=n_games, sims=sims, evaluator=ev, seed=11, noise_mode=1, games_target=n_games, move_log=True,
                       max_sims_per_tick=4)
    e.set_noise_table(table)
    logs = H.run_logged(e)
    recs, res = H.split_by_game(e.drain_records(), e.drain_results())
    assert len(res) == n_games
    for g in range(n_games):
        o = H.replay_in_oracle(0, sims, 11, g, table[g], logs[g])
        H.compare_game(0, e.move_log(g), recs[g], res[g], o)
    e.close()


def test_two_native_towers_head_to_head_replays_in_oracle():
    """BASELINE.json configs[3] shape (elo.py head-to-head, two random-init nets, evaluate mode) on the native path:
    leaves are partitioned by network on the device; every game is replayed through the oracle with the logged outputs, and
    each tower's outputs equal what it returns for the same boards on its own."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    from tests import helpers as H
    torch.manual_seed(0)
    a = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    torch.manual_seed(1)
    b = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    n_games, sims = 30, 40
    sp = BatchedSelfPlay(a, game=0, n_games=n_games, sims=sims, net="tower", evaluation_network=b, evaluate=True, update=False, seed=5,
                         games_target=n_games, noise_mode=0, move_log=True)
    assert isinstance(sp.evaluator, nets.TwoTowerEvaluator)
    logs = H.run_logged(sp.engine)
    _, res = H.split_by_game(sp.engine.drain_records(), sp.engine.drain_results())
    assert len(res) == n_games
    for g in range(n_games):
        o = H.replay_in_oracle(0, sims, 5, g, None, logs[g], evaluate=True)
        assert res[g]["reward"] == o["reward"] and res[g]["plies"] == len(o["moves"])
        ml = sp.engine.move_log(g)
        assert [m["action"] for m in ml] == [m["action"] for m in o["moves"]] and [m["n"] for m in ml] == [list(m["n"]) for m in o["moves"]]
    # routing check: tree-k evaluations equal tower k evaluated directly on the logged boards
    for k, tw in enumerate(sp.evaluator.towers):
        own = torch.tensor(np.array(logs[3][k]["own"], dtype=np.uint64).view(np.int64), device="cuda")
        opp = torch.tensor(np.array(logs[3][k]["opp"], dtype=np.uint64).view(np.int64), device="cuda")
        p, v = tw.forward_bits(own, opp)
        assert np.array_equal(p.cpu().numpy(), np.array(logs[3][k]["policy"])) and np.array_equal(v.cpu().numpy(), np.array(logs[3][k]["value"], np.float32))
    sp.close()


def test_fused_heads_agree_with_the_separate_heads_kernel_and_respect_the_mask(monkeypatch, dtype):
    """The FC heads inside the tower kernel (default) vs heads_kernel after it (SPX_TOWER_FUSED_HEADS=0): same 16-bit inputs,
    fp32 accumulation in a different order -> 1e-5; the fused kernel never writes rows whose needs_eval is 0."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.envs import boards_to_bits
    torch.manual_seed(4)
    net = nets.ResidualTower(7, 6, 7, num_blocks=3).eval()
    _randomise_bn(net)
    n = 2100                                               # three units on some clusters: buffer reuse after the FC phase
    bits = boards_to_bits(torch.from_numpy(_random_positions(n, 11)).cuda(), 0)
    own, opp = bits[:, 0].contiguous(), bits[:, 1].contiguous()
    fused = nets.NativeTower(net)
    if not fused.fused_heads:
        fused.close()
        pytest.skip("single-CTA tower selected (SPX_TOWER_NCTA=1): the heads are never fused there")
    monkeypatch.setenv("SPX_TOWER_FUSED_HEADS", "0")
    plain = nets.NativeTower(net)
    monkeypatch.delenv("SPX_TOWER_FUSED_HEADS")
    assert fused.fused_heads and not plain.fused_heads
    pf, vf = fused.forward_bits(own, opp)
    pp, vp = plain.forward_bits(own, opp)
    torch.cuda.synchronize()
    assert (pf - pp).abs().max().item() < 1e-5 and (vf - vp).abs().max().item() < 1e-5
    need = (torch.arange(n, device="cuda") % 3 != 0).to(torch.uint8)
    need[14:42] = 0                                        # two whole SM-pair units without work
    for tw, (p0, v0) in ((fused, (pf, vf)), (plain, (pp, vp))):
        p = torch.full((n, 7), -1.0, device="cuda")
        v = torch.full((n,), -2.0, device="cuda")
        tw.forward_bits(own, opp, needs_eval=need, policy=p, value=v)
        torch.cuda.synchronize()
        m = need.bool()
        assert torch.equal(p[m], p0[m]) and torch.equal(v[m], v0[m])
        if tw is fused:   # the separate heads kernel skips whole 16-board groups only: rows without a request are unspecified there
            assert bool((p[~m] == -1.0).all()) and bool((v[~m] == -2.0).all())
    fused.close(); plain.close()


@pytest.mark.parametrize("blocks,n", [(1, 7), (15, 300), (15, 1100)])
def test_tictactoe_tower_matches_fp32_reference(blocks, n, dtype):
    """ResidualTower.from_env(TicTacToeEnv, ...) (main.py:74 with --g tictactoe): 3x3 boards embedded in the tower's board slots."""
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.envs import boards_to_bits
    torch.manual_seed(10 + blocks)
    net = nets.ResidualTower(3, 3, 9, num_blocks=blocks).eval()
    _randomise_bn(net)
    rng = np.random.default_rng(blocks)
    boards = torch.from_numpy(rng.integers(-1, 2, size=(n, 3, 3)).astype(np.int64))
    bits = boards_to_bits(boards.cuda(), 1)
    tw = nets.NativeTower(net)
    assert tw.game == 1 and tw.A == 9
    p, v = tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        pr, vr = net.cuda().float().forward(boards.cuda())
    dp, dv = (p - pr).abs().max().item(), (v - vr.reshape(-1)).abs().max().item()
    print(f"tictactoe {dtype} blocks={blocks} n={n} max|dpolicy|={dp:.3e} max|dvalue|={dv:.3e}")
    assert p.shape == (n, 9) and torch.allclose(p.sum(1), torch.ones(n, device="cuda"), atol=1e-5)
    assert dp < TOL[dtype]["policy"] and dv < TOL[dtype]["value"]
    perm = torch.randperm(n, device="cuda")
    p2, v2 = tw.forward_bits(bits[perm, 0].contiguous(), bits[perm, 1].contiguous())
    assert torch.equal(p2, p[perm]) and torch.equal(v2, v[perm])
    tw.close()


def test_tictactoe_selfplay_with_native_residual_tower_replays_in_oracle():
    from self_play_reinforcement_learning_b200 import envs, nets
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    from tests import helpers as H
    torch.manual_seed(2)
    net = nets.ResidualTower(3, 3, 9, num_blocks=2).eval()
    n_games, sims = 24, 60
    sp = BatchedSelfPlay(net, env=envs.TicTacToeEnv, n_games=n_games, sims=sims, net="tower", seed=8, games_target=n_games, noise_mode=0,
                         move_log=True)
    logs = H.run_logged(sp.engine)
    recs, res = H.split_by_game(sp.engine.drain_records(), sp.engine.drain_results())
    assert len(res) == n_games
    for g in range(n_games):
        o = H.replay_in_oracle(1, sims, 8, g, None, logs[g])
        H.compare_game(1, sp.engine.move_log(g), recs[g], res[g], o)
    sp.close()
