"""BASELINE.json configs[3]: elo.py head-to-head eval, two random-init ResidualTower-20 nets, 4096 games, 400 sims/move
(evaluate mode, no records) on one B200 through the native two-tower path.  Prints sims/s and games/s."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
torch.manual_seed(0)
a = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
torch.manual_seed(1)
b = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(a, game=0, n_games=games, sims=400, net="tower", evaluation_network=b, evaluate=True, update=False, seed=0)
sp.engine.run_ticks(400)
torch.cuda.synchronize()
c0 = sp.engine.counters()
t0 = time.time()
sp.engine.run_ticks(ticks)
torch.cuda.synchronize()
dt = time.time() - t0
c1 = sp.engine.counters()
print({"games": games, "sims_per_s": (c1["sims"] - c0["sims"]) / dt, "moves_per_s": (c1["moves"] - c0["moves"]) / dt,
       "ms_per_tick": 1e3 * dt / ticks, "leaf_evals_per_tick": (c1["leaf_evals"] - c0["leaf_evals"]) / ticks})
sp.close()
