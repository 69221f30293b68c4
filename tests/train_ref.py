"""fp32 PyTorch autograd reference of one training step (test infrastructure for csrc/spx_train.cu): the reference's
MCTreeSearch.loss (games/algos/mcts.py:234-252) on ResidualTower.train() with the Dropout masks injected."""
import torch
import torch.nn.functional as F


class FixedDropout(torch.nn.Module):
    """nn.Dropout(p=0.5) in training mode with a given keep-mask: x * keep * 2."""

    def __init__(self):
        super().__init__()
        self.keep = None

    def forward(self, x):
        return x * self.keep.to(x.dtype) * 2.0


def make_batch(B, seed=0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    boards = torch.randint(-1, 2, (B, 7, 6), generator=g)
    planes = torch.stack([(boards == 0), (boards == 1), (boards == -1)], 1).float()
    probs = torch.rand(B, 7, generator=g) + 0.05
    probs = probs / probs.sum(1, keepdim=True)
    target = torch.rand(B, generator=g) * 2 - 1
    mask = (torch.rand(B, 2, 1344, generator=g) < 0.5).to(torch.uint8)
    return planes.to(device), probs.to(device), target.to(device), mask.to(device)


def patch_dropout(net):
    net.policy_dropout, net.value_dropout = FixedDropout(), FixedDropout()
    return net


def torch_loss(net, planes, probs, target, mask):
    """(loss, value_loss, prob_loss, p, v) of net.train() in true fp32 (TF32 off)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net.policy_dropout.keep, net.value_dropout.keep = mask[:, 0], mask[:, 1]
    p, v = net.forward_planes(planes)
    value_loss = F.mse_loss(v.view(-1), target)
    prob_loss = -(p.log() * probs).sum() / p.size(0)
    return value_loss + prob_loss, value_loss, prob_loss, p, v
