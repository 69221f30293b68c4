"""CPU: scheduler-level glue restated from the reference -- Memory, parse_results, the MCTS loss (pinned against
the reference's own MCTreeSearch.loss through a golden value)."""
import os

import numpy as np
import torch

from self_play_reinforcement_learning_b200 import nets
from self_play_reinforcement_learning_b200.scheduler import Memory, mcts_loss, parse_results
from self_play_reinforcement_learning_b200.selfplay import Move


def test_loss_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nets.npz"))
    torch.manual_seed(3)
    net = nets.ResidualTower(7, 6, 7, num_blocks=2).eval()
    batch = [Move(torch.from_numpy(g["loss_states"][i].astype(np.int64)), torch.tensor(float(g["loss_val"][i])),
                  torch.from_numpy(g["loss_probs"][i]), torch.tensor(float(g["loss_q"][i]))) for i in range(16)]
    with torch.no_grad():
        got = float(mcts_loss(net, batch))
    assert abs(got - float(g["loss_value"][0])) < 1e-5  # fp32, same weights, same formula (mcts.py:234-252)


def test_parse_results_breakdown():
    rl = [{"reward": 1, "swap_sides": False}, {"reward": -1, "swap_sides": True}, {"reward": 0, "swap_sides": True},
          {"reward": 1, "swap_sides": True}, {"reward": 1, "swap_sides": False}]
    total, bd = parse_results(rl)
    assert total == 2
    assert bd == {"first": dict(wins=2, draws=0, losses=0), "second": dict(wins=1, draws=1, losses=1)}


def test_memory_fifo_and_sampling():
    m = Memory(5)
    for i in range(8):
        m.add(i)
    assert len(m) == 5 and sorted(m.sample(5)) == [3, 4, 5, 6, 7]
    m.change_size(3)
    assert len(m) == 3 and sorted(m.sample(3)) == [5, 6, 7]


def test_training_reference_formula_reproduces_the_golden_loss(golden_dir):
    """tests/train_ref.torch_loss -- the autograd reference the device SGD step (csrc/spx_train.cu) is held against -- is the
    reference's MCTreeSearch.loss (mcts.py:234-252): with the network in eval mode and every Dropout unit kept-and-unscaled it
    reproduces the loss value the unmodified reference computed for the golden batch (q_average: target = actual_val + q)."""
    from tests import train_ref as R
    g = np.load(os.path.join(golden_dir, "nets.npz"))
    torch.manual_seed(3)
    net = R.patch_dropout(nets.ResidualTower(7, 6, 7, num_blocks=2)).eval()
    boards = torch.from_numpy(g["loss_states"].astype(np.int64))
    planes = torch.stack([(boards == 0), (boards == 1), (boards == -1)], 1).float()
    half = torch.full((16, 2, 1344), 0.5)            # FixedDropout computes x * keep * 2: keep = 0.5 is the identity (eval-mode Dropout)
    with torch.no_grad():
        loss, *_ = R.torch_loss(net, planes, torch.from_numpy(g["loss_probs"]), torch.from_numpy(g["loss_val"] + g["loss_q"]), half)
    assert abs(float(loss) - float(g["loss_value"][0])) < 1e-5
