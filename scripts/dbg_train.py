"""Layer-by-layer comparison of the device training step (csrc/spx_train.cu) with fp32 PyTorch autograd.
    python scripts/dbg_train.py [blocks] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.train import DeviceTrainer  # noqa: E402
from tests import train_ref as R  # noqa: E402

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(0)
net = R.patch_dropout(nets.ResidualTower(7, 6, 7, num_blocks=blocks)).cuda().float()
with torch.no_grad():   # non-trivial BatchNorm parameters
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
planes, probs, target, mask = R.make_batch(B)
tr = DeviceTrainer(net, batch_size=B, lr=0.01)
acts = {}
hooks = []


def keep(name):
    def fn(mod, inp, out):
        out.retain_grad()
        acts[name] = out
    return fn


hooks.append(net.conv1.register_forward_hook(keep("y0")))
for i, blk in enumerate(net.residual_blocks):
    hooks.append(blk.conv1.register_forward_hook(keep(f"y{2 * i + 1}")))
    hooks.append(blk.conv2.register_forward_hook(keep(f"y{2 * i + 2}")))
    hooks.append(blk.register_forward_hook(keep(f"a{2 * i + 2}")))
hooks.append(net.conv_policy.register_forward_hook(keep("yp")))
hooks.append(net.conv_value.register_forward_hook(keep("yv")))
net.train()
loss, lv, lp, p, v = R.torch_loss(net, planes, probs, target, mask)
loss.backward()
out = tr.step(planes, probs, target, dropout_mask=mask, apply_update=False)
torch.cuda.synchronize()
print("loss torch", loss.item(), lv.item(), lp.item(), " device", out.tolist())
dp, dv = tr.outputs()
print("max |dp|", (dp - p).abs().max().item(), "max |dv|", (dv - v.view(-1)).abs().max().item())


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


L = 2 * blocks + 1
for l in range(L):
    print(f"layer {l}: conv out rel err {rel(tr.debug_planes(1, l), acts[f'y{l}']):.2e}", end="")
    if f"a{l}" in acts:
        print(f"  block out rel err {rel(tr.debug_planes(2, l), acts[f'a{l}']):.2e}", end="")
    print()
yh = tr.debug_planes(3)
print("head conv out rel err", rel(yh[:, :32], acts["yp"]), rel(yh[:, 32:], acts["yv"]))
dyh = tr.debug_planes(8)
print("head conv-out grad rel err", rel(dyh[:, :32], acts["yp"].grad), rel(dyh[:, 32:], acts["yv"].grad))
# gradient w.r.t. conv outputs: only the LAST layer processed leaves its dy in the buffer (the stem); per-layer check through the weights
g = DeviceTrainer.unflatten(tr.gradients_flat(), net)
worst = 0.0
for name, prm in net.named_parameters():
    e = rel(g[name], prm.grad)
    a = (g[name] - prm.grad).abs().max().item()
    worst = max(worst, e if prm.grad.norm() > 1e-6 else 0.0)
    print(f"  grad {name:45s} rel {e:.2e} abs {a:.2e} |ref| {prm.grad.norm().item():.3e}")
print("worst relative gradient error", worst)
print("stem dy rel err", rel(tr.debug_planes(7), acts["y0"].grad))
