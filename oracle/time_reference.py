"""Calibration of the CPU baseline (runs only where /root/reference exists, i.e. in the build container, not on the GPU box):
the UNMODIFIED reference (games/algos/mcts.py MCTreeSearch + games/general/modules.py ResidualTower-20, direct mode, one
process, one torch thread) next to the oracle port bench.py times on the GPU box, same search budget, same machine.

    python -m oracle.time_reference [seconds]

TEST/BENCH INFRASTRUCTURE ONLY.  Writes nothing; prints one JSON line (recorded in DESIGN.md section 5)."""
import json
import os
import sys
import time

import numpy as np
import torch

from . import ref_harness as rh


def main(seconds=40.0, sims=800, blocks=20):
    assert rh.reference_available(), "needs /root/reference"
    torch.set_num_threads(1)
    mcts, SelfPlayer, Connect4Env, _ = rh._import_reference()
    from games.general.modules import ResidualTower
    torch.manual_seed(0)
    np.random.seed(0)
    net = ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    pol = mcts.MCTreeSearch(network=net, env=Connect4Env, iterations=sims, memory_queue=None)
    pol.reset()
    count = [0]
    orig = pol.search_node

    def counted(*a, **k):
        count[0] += 1
        return orig(*a, **k)
    pol.search_node = counted
    t0 = time.time()
    with torch.no_grad():
        while time.time() - t0 < seconds:      # one search() = `sims` search_node calls (mcts.py:323-338); ~12 s each on one core
            pol.search()
    dt_ref = time.time() - t0
    ref = count[0] / dt_ref

    # the port, same budget, same process limits (what bench.py's cpu_baseline runs per core)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    r = bench.cpu_baseline(seconds, blocks, sims, procs=1)
    print(json.dumps({"reference_sims_per_s_per_core": ref, "port_sims_per_s_per_core": r["sims_per_s"],
                      "port_over_reference": r["sims_per_s"] / ref, "seconds": seconds, "sims_per_move": sims, "blocks": blocks,
                      "host_cores_here": len(os.sched_getaffinity(0))}))


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 40.0)
