"""The hand-written tcgen05 tower (spx_tower_forward through the C ABI) vs the fp32 torch forward of the
same ResidualTower.  Floating point: tolerance stated per assertion (bf16 weights/activations, fp32 accumulate)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# bf16 trunk, fp32 accumulation: absolute tolerance on softmax policy and tanh value vs the fp32 reference
TOL_POLICY = 2e-2
TOL_VALUE = 6e-2


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


@pytest.mark.parametrize("blocks,n", [(1, 7), (2, 50), (20, 300), (20, 1024), (20, 2500)])
def test_tower_matches_fp32_reference(blocks, n):
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
    with torch.no_grad():
        pr, vr = net.cuda().float().forward(boards.cuda())
    dp = (p - pr).abs().max().item()
    dv = (v - vr.reshape(-1)).abs().max().item()
    print(f"blocks={blocks} n={n} max|dpolicy|={dp:.3e} max|dvalue|={dv:.3e}")
    assert torch.allclose(p.sum(1), torch.ones(n, device="cuda"), atol=1e-5)
    assert dp < TOL_POLICY and dv < TOL_VALUE
    # row independence: a board's outputs do not depend on its batch position or its neighbours
    perm = torch.randperm(n, device="cuda")
    p2, v2 = tw.forward_bits(bits[perm, 0].contiguous(), bits[perm, 1].contiguous())
    assert torch.equal(p2, p[perm]) and torch.equal(v2, v[perm])
    tw.close()


def _pending_tree(e):
    from self_play_reinforcement_learning_b200 import _lib
    t = torch.zeros(e.n_games, dtype=torch.int32, device=e.device)
    _lib.check(_lib.lib().spx_pending_tree(e._h, t.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "spx_pending_tree")
    return t.cpu().numpy()


def test_selfplay_with_tower_replays_bit_exact_in_oracle():
    """Self-play driven by the native tower: log every evaluation, replay it through the C oracle and require
    identical games ("visit counts bit-exact given identical network outputs")."""
    from oracle import oracle as ox
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.engine import SelfPlayEngine
    from tests import helpers as H
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    n_games, sims = 14, 40
    ev = nets.TowerEvaluator(net)
    rng = np.random.default_rng(3)
    table = rng.dirichlet([1.0] * 7, size=(n_games, 2, 22))
    e = SelfPlayEngine(game=0, n_games=n_games, sims=sims, evaluator=ev, seed=11, noise_mode=1, games_target=n_games, move_log=True,
                       max_sims_per_tick=4)
    e.set_noise_table(table)
    logs = [[dict(own=[], opp=[], policy=[], value=[]) for _ in (0, 1)] for _ in range(n_games)]
    for _ in range(200000):
        e.tick()
        torch.cuda.synchronize()
        need = e.needs_eval.cpu().numpy().astype(bool)
        if not need.any() and e.all_idle():
            break
        own, opp = e.leaf_own.cpu().numpy().view(np.uint64), e.leaf_opp.cpu().numpy().view(np.uint64)
        pol, val = e.policy.cpu().numpy(), e.value.cpu().numpy()
        tree = _pending_tree(e)
        for g in np.flatnonzero(need):
            L = logs[g][tree[g]]
            L["own"].append(own[g]); L["opp"].append(opp[g]); L["policy"].append(pol[g].copy()); L["value"].append(val[g])
    recs, res = H.split_by_game(e.drain_records(), e.drain_results())
    assert len(res) == n_games
    for g in range(n_games):
        rs = ox.make_replay(0, logs[g])
        cfg = ox.make_cfg(0, sims, seed=11, game_uid=g, noise_table=table[g])
        pair = (ox.fn_addr("ox_replaynet"), C.addressof(rs))
        o = ox.play_episode(cfg, bool(g & 1), nets=(pair, pair))
        assert rs.mismatches == 0 and rs.overruns == 0, (g, rs.mismatches, rs.overruns)
        assert rs.cursor[0] == rs.n[0] and rs.cursor[1] == rs.n[1]
        H.compare_game(0, e.move_log(g), recs[g], res[g], o)
    e.close()
