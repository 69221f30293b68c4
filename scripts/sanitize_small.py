"""Small end-to-end run for compute-sanitizer: engine + native tower + heads + env kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from self_play_reinforcement_learning_b200 import envs, nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
sp = BatchedSelfPlay(net, game=0, n_games=20, sims=12, net="tower", seed=1, games_target=20)
moves, results = sp.play_games(max_ticks=4000, poll_every=64)
print("games", len(results), "moves", len(moves), sp.engine.counters())
sp.close()
sp = BatchedSelfPlay(None, game=1, n_games=9, sims=10, net="hash", seed=2, games_target=9, opponent="lookahead", evaluate=True, update=False)
print(len(sp.play_games(max_ticks=2000, poll_every=32)[1]))
sp.close()
e = envs.Connect4Env(1000, strict=False)
for t in range(10):
    e.step(torch.randint(0, 7, (1000,), device="cuda", dtype=torch.int32), 1 if t % 2 == 0 else -1)
torch.cuda.synchronize()
print("ok")
