"""Times the device training step (20 blocks, batch 128) -- run under ncu for the per-kernel launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.train import DeviceTrainer  # noqa: E402
from tests import train_ref as R  # noqa: E402

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 50
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).cuda()
planes, probs, target, _ = R.make_batch(B)
tr = DeviceTrainer(net, batch_size=B)
for _ in range(3):
    tr.step(planes, probs, target)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    tr.step(planes, probs, target)
b.record()
torch.cuda.synchronize()
print("ms per step", a.elapsed_time(b) / n)

# where a conv_tf32_kernel launch goes: clock64 of CTA 0 at 8 points (the last launch of a step = the stem's backward... no: the last conv
# launch of a step is the backward-data conv of layer 1)
from self_play_reinforcement_learning_b200._lib import lib  # noqa: E402
tb = torch.zeros(8, dtype=torch.int64, device="cuda")
lib().spx_train_debug_trace(tb.data_ptr())
tr.step(planes, probs, target)
torch.cuda.synchronize()
lib().spx_train_debug_trace(None)
t = tb.cpu().numpy()
names = ["start", "setup done", "producer issued all", "A tile landed", "all MMAs issued", "accumulators complete", "TMEM->smem done", "epilogue done"]
print("conv kernel CTA 0 (cycles since start):", {n: int(t[i] - t[0]) for i, n in enumerate(names)})
