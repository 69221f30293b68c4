"""Where does a fused tick go?  (1) the plain tower kernel launched back to back for a second (sustained clocks under the power
cap: the honest 'network alone' number -- 64 instrumented launches between idle gaps run at a higher clock), (2) the fused tick
kernel, (3) the fused tick kernel with the search switched off (SPX_DBG_FLAGS=1 in a second process).
    python scripts/dbg_fused_gap.py [games]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=G, sims=800, net="tower", seed=0)
e, ev = sp.engine, sp.evaluator
e.stagger()   # steady-state mix of game phases
e.run_ticks(1600, chunk=100)
torch.cuda.synchronize()


import subprocess
import threading
lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l) for l in proc.stdout], daemon=True).start()
clk = {}


def timed(fn, n, tag=None):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    lines.clear()
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    mhz = sorted(float(l.split(",")[0]) for l in lines if "," in l)
    if tag and mhz:
        clk[tag] = mhz[len(mhz) // 2]
    return a.elapsed_time(b) / n


def tower_only(n):
    for _ in range(n):
        ev(e)


out = {}
for rep in range(2):
    c0 = e.counters()
    out["fused_ms_per_tick"] = timed(lambda: e.run_ticks(3000, chunk=100), 3000, "fused_mhz")
    c1 = e.counters()
    out["leaves_per_tick"] = (c1["leaf_evals"] - c0["leaf_evals"]) / 3000
    out["sims_per_tick"] = (c1["sims"] - c0["sims"]) / 3000
    if os.environ.get("SPX_GAP_TOWER", "1") != "0":
        out["tower_back_to_back_ms"] = timed(lambda: tower_only(3000), 3000, "tower_mhz")
print(os.environ.get("SPX_DBG_FLAGS", "0"), G, out, clk, flush=True)
proc.terminate()
