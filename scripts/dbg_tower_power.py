"""Sustained (power-capped) time of the plain tower kernel launched back to back for ~2 s, with the SM clock and board power
sampled during the run.  Used to price what each part of the kernel costs in ENERGY (the run is power limited): compare the
default library with -DSPX_DBG_NO_TMA (no weight streaming) / -DSPX_DBG_SKIP_EPI (no epilogue arithmetic) builds.
    SPX_LIB_PATH=variants/libspx_NO_TMA.so python scripts/dbg_tower_power.py [boards]"""
import os
import subprocess
import sys
import threading

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.envs import boards_to_bits  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
rng = np.random.default_rng(0)
boards = torch.from_numpy(rng.integers(-1, 2, size=(n, 7, 6)).astype(np.int64))
bits = boards_to_bits(boards.cuda(), 0)
own, opp = bits[:, 0].contiguous(), bits[:, 1].contiguous()
tw = nets.NativeTower(net)
lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l) for l in proc.stdout], daemon=True).start()
for _ in range(2000):
    tw.forward_bits(own, opp)
torch.cuda.synchronize()
lines.clear()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
N = 5000
for _ in range(N):
    tw.forward_bits(own, opp)
b.record()
torch.cuda.synchronize()
proc.terminate()
mhz = sorted(float(l.split(",")[0]) for l in lines if "," in l)
pw = sorted(float(l.split(",")[1]) for l in lines if "," in l)
print(os.environ.get("SPX_LIB_PATH", "default"), "boards", n, "ms/launch", round(a.elapsed_time(b) / N, 4), "median MHz", mhz[len(mhz) // 2] if mhz else None,
      "median W", pw[len(pw) // 2] if pw else None, flush=True)
