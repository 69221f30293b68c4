import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets
from self_play_reinforcement_learning_b200.envs import boards_to_bits
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 7
torch.manual_seed(blocks)
net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
rng = np.random.default_rng(0)
boards = torch.from_numpy(rng.integers(-1, 2, size=(n, 7, 6)).astype(np.int64))
bits = boards_to_bits(boards.cuda(), 0)
tw = nets.NativeTower(net)
p, v = tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
torch.cuda.synchronize()
with torch.no_grad():
    pr, vr = net.cuda().float().forward(boards.cuda())
print("policy native", p[:2].cpu().numpy())
print("policy ref   ", pr[:2].cpu().numpy())
print("value native", v[:7].cpu().numpy(), "ref", vr.reshape(-1)[:7].cpu().numpy())
print("max diff", (p - pr).abs().max().item(), (v - vr.reshape(-1)).abs().max().item())
import time
for _ in range(3):
    tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
flop = 2 * n * 42 * (9 * 3 * 128 + blocks * 2 * 9 * 128 * 128 + 128 * 64)
print(f"forward {ms:.3f} ms for n={n} blocks={blocks}: {flop / ms / 1e9:.1f} TFLOP/s useful")
