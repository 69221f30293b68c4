"""How the fused tick kernel's time depends on the ticks per launch (a fixed cost per launch vs a per-tick cost):
    python scripts/dbg_chunk_scaling.py [games]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=G, sims=800, net="tower", seed=0)
e = sp.engine
e.stagger()
e.run_ticks(2000, chunk=100)
torch.cuda.synchronize()
for chunk in (25, 100, 400, 800, 1600):
    n = 3200
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e.run_ticks(n, chunk=chunk)
    b.record()
    torch.cuda.synchronize()
    print(G, "chunk", chunk, "ms/tick", round(a.elapsed_time(b) / n, 4), flush=True)
