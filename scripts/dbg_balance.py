"""Exact launches (spx_tick_fused: the launch lasts as long as its slowest SM pair) against work-conserving launches
(spx_tick_fused_balanced), steady-state mix of game phases, seconds-long runs under the power cap:
    python scripts/dbg_balance.py [games]"""
import os
import subprocess
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=G, sims=800, net="tower", seed=0, eval_cache=int(os.environ.get("SPX_EVAL_CACHE", "0")), max_sims_per_tick=int(os.environ.get("SPX_MAX_SIMS", "8")))
e = sp.engine
e.stagger()
e.run_ticks(1600, chunk=100)
torch.cuda.synchronize()
lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l) for l in proc.stdout], daemon=True).start()
N = 3200
for rep in range(2):
    for balanced, chunk in (((True, 400),) if len(sys.argv) > 2 else ((False, 100), (True, 100), (False, 400), (True, 400), (True, 800))):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        lines.clear()
        c0 = e.counters()
        a.record()
        e.run_ticks(N, chunk=chunk, balanced=balanced)
        b.record()
        torch.cuda.synchronize()
        c1 = e.counters()
        e.drain_records(); e.drain_results()
        ms = a.elapsed_time(b)
        mhz = sorted(float(l.split(",")[0]) for l in lines if "," in l)
        print(f"games {G} balanced {int(balanced)} chunk {chunk}: {ms / N:.4f} ms/tick, {(c1['sims'] - c0['sims']) / ms / 1e3:.4f} M sims/s, "
              f"{(c1['leaf_evals'] - c0['leaf_evals']) / ms / 1e3:.4f} M leaf evals/s, leaves/tick {(c1['leaf_evals'] - c0['leaf_evals']) / N:.1f}, "
              f"cache hits/tick {(c1['cache_hits'] - c0['cache_hits']) / N:.1f}, sims/tick {(c1['sims'] - c0['sims']) / N:.1f}, errors {c1['errors']}, SM clock {mhz[len(mhz) // 2] if mhz else None}", flush=True)
proc.terminate()
