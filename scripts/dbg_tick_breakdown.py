"""Where does the tick go?  Sustained timings (>= 1 s each) of: the full tick, the tower alone on the loop's own leaf batch,
the tower alone on dense random boards, the tower alone on empty boards, and the advance kernel alone (hash net)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from self_play_reinforcement_learning_b200 import nets
from self_play_reinforcement_learning_b200.envs import boards_to_bits
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=1024, sims=800, net="tower", seed=0)
e, ev = sp.engine, sp.evaluator
e.run_ticks(2400)
torch.cuda.synchronize()


def timed(fn, n):
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / n


def smi():
    import subprocess
    try:
        return subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], text=True).strip()
    except Exception:
        return "?"


print("full tick            %.4f ms" % timed(e.tick, N), smi())
for chunk in (8, 64, 256, 800):
    print("fused ticks, chunk %-4d %.4f ms" % (chunk, timed(lambda: e.run_ticks(chunk, fused=True, chunk=chunk), N // chunk) / chunk), smi())
print("full tick            %.4f ms" % timed(e.tick, N), smi())
print("tower, loop leaves   %.4f ms" % timed(lambda: ev(e), N), smi())
tw = ev.tower
rng = np.random.default_rng(0)
dense = boards_to_bits(torch.from_numpy(rng.integers(-1, 2, size=(1024, 7, 6)).astype(np.int64)).cuda(), 0)
d0, d1 = dense[:, 0].contiguous(), dense[:, 1].contiguous()
print("tower, dense random  %.4f ms" % timed(lambda: tw.forward_bits(d0, d1), N), smi())
z = torch.zeros_like(d0)
print("tower, empty boards  %.4f ms" % timed(lambda: tw.forward_bits(z, z), N), smi())
print("tower, 10 launches   %.4f ms" % timed(lambda: tw.forward_bits(d0, d1), 10), smi())
time.sleep(2.0)
print("tower, 10 after idle %.4f ms" % timed(lambda: tw.forward_bits(d0, d1), 10), smi())
print("full tick again      %.4f ms" % timed(e.tick, N), smi())
# advance + tower with a host-side gap removed: two ticks' worth of launches queued before the GPU starts
print(e.counters())
