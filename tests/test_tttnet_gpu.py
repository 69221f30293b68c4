"""The hand-written fp32 TicTacToe network (spx_tttnet_forward) vs the fp32 torch forward of ConvNetTicTacToe, and
BASELINE.json configs[0] on the fully native path: TicTacToe, the repo's tictactoe net, 1 game, 100 sims/move."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-5  # fp32 path (north_star: 1e-5 fp32)


def _net(seed):
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(seed)
    net = nets.ConvNetTicTacToe(3, 3, 9).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.uniform_(-0.1, 0.1)
    return net


def test_tttnet_matches_fp32_reference_on_all_reachable_style_boards():
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.envs import boards_to_bits
    net = _net(1)
    rng = np.random.default_rng(0)
    boards = torch.from_numpy(rng.integers(-1, 2, size=(4000, 3, 3)).astype(np.int64))
    boards[0] = 0
    bits = boards_to_bits(boards.cuda(), 1)
    ev = nets.TTTNetEvaluator(net)
    p, v = ev.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
    torch.cuda.synchronize()
    with torch.no_grad():
        pr, vr = net.float().forward(boards)   # true fp32 on the CPU (cuDNN convolutions default to TF32 on the GPU)
    dp, dv = (p.cpu() - pr).abs().max().item(), (v.cpu() - vr.reshape(-1)).abs().max().item()
    print(f"max|dpolicy|={dp:.3e} max|dvalue|={dv:.3e}")
    assert dp < TOL and dv < TOL
    assert torch.allclose(p.sum(1), torch.ones(len(boards), device="cuda"), atol=1e-5)
    ev.close()


def test_config1_native_tictactoe_one_game_100_sims():
    from self_play_reinforcement_learning_b200 import envs
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    net = _net(2)
    table = np.random.default_rng(1).dirichlet([1.0] * 9, size=(1, 2, 22))
    sp = BatchedSelfPlay(net, env=envs.TicTacToeEnv, n_games=1, sims=100, net="tttnet", seed=2, games_target=1, noise_mode=1, move_log=True)
    sp.engine.set_noise_table(table)
    logs = H.run_logged(sp.engine)
    recs, res = H.split_by_game(sp.engine.drain_records(), sp.engine.drain_results())
    o = H.replay_in_oracle(1, 100, 2, 0, table[0], logs[0])
    H.compare_game(1, sp.engine.move_log(0), recs[0], res[0], o)
    assert 5 <= res[0]["plies"] <= 9 and sp.engine.counters()["errors"] == 0
    sp.close()
