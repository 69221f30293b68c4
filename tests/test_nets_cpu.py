"""CPU: the package's network definitions reproduce the reference classes (same random init under a
seed, same state_dict keys, same fp32 outputs) -- pinned by golden vectors made from the reference's
own games/general/modules.py and games/tictactoe/modules.py."""
import os

import numpy as np
import torch

from self_play_reinforcement_learning_b200 import nets

CASES = {"tower20": lambda: nets.ResidualTower(7, 6, 7, num_blocks=20), "tower2": lambda: nets.ResidualTower(7, 6, 7, num_blocks=2),
         "tower_ttt": lambda: nets.ResidualTower(3, 3, 9, num_blocks=3), "convttt": lambda: nets.ConvNetTicTacToe(3, 3, 9)}


def test_networks_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nets.npz"))
    torch.set_num_threads(1)
    for name, ctor in CASES.items():
        torch.manual_seed(int(g[name + "_seed"][0]))
        net = ctor().eval()
        wsum = sum(float(t.double().abs().sum()) for t in net.state_dict().values())
        assert wsum == float(g[name + "_wsum"][0]), name  # identical random init
        with torch.no_grad():
            p, v = net.forward(torch.from_numpy(g[name + "_x"].astype(np.int64)))
        # same weights, same math; allow fp32 reassociation differences between CPU kernels/hosts
        assert np.allclose(p.numpy(), g[name + "_policy"], atol=1e-5), name
        assert np.allclose(v.numpy(), g[name + "_value"], atol=1e-5), name
        assert np.allclose(p.numpy().sum(1), 1.0, atol=1e-5)


def test_call_convention_flips_frame():
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    s = np.zeros((7, 6), np.int64)
    s[3, 0] = 1
    with torch.no_grad():
        p1, v1 = net(s, 1)
        p2, v2 = net(-s, -1)
    assert p1 == p2 and v1 == -v2 and isinstance(p1, list) and isinstance(v1, float)


def test_planes_from_bits_equal_planes_from_boards():
    from oracle import spec
    rng = np.random.default_rng(0)
    for game in (0, 1):
        W, H, A = spec.GAME_DIMS[game]
        boards = rng.integers(-1, 2, size=(32, W, H))
        bits = np.array([spec.board_to_bits(b, game) for b in boards], dtype=np.int64)
        a = nets.board_planes(torch.from_numpy(boards), W, H)
        b = nets.bits_to_planes(torch.from_numpy(bits[:, 0]), torch.from_numpy(bits[:, 1]), game)
        assert torch.equal(a, b)


def test_pack_tower_blob_rejects_networks_the_kernel_is_not_built_for():
    import pytest
    from self_play_reinforcement_learning_b200 import nets
    with pytest.raises(ValueError, match="filter_factor"):
        nets.pack_tower_blob(nets.ResidualTower(7, 6, 7, num_blocks=1, filter_factor=16))
    with pytest.raises(ValueError, match="board"):
        nets.pack_tower_blob(nets.ResidualTower(5, 4, 5, num_blocks=1))
    with pytest.raises(ValueError, match="ResidualTower"):
        nets.pack_tower_blob(nets.ConvNetTicTacToe(3, 3, 9))
    assert nets.pack_tower_blob(nets.ResidualTower(7, 6, 7, num_blocks=1)).dtype.is_floating_point is False
