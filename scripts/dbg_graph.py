"""Does replaying the tick (advance + tower + heads) from a CUDA graph beat stream launches?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_play_reinforcement_learning_b200 import nets
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=1024, sims=800, net="tower", seed=0)
e = sp.engine
e.run_ticks(2400)
torch.cuda.synchronize()
def timed(fn, n):
    c0 = e.counters()["sims"]; torch.cuda.synchronize(); t0 = time.time(); fn(n); torch.cuda.synchronize(); dt = time.time() - t0
    return (e.counters()["sims"] - c0) / dt, 1e3 * dt / n
print("stream launches: sims/s %.0f, ms/tick %.4f" % timed(lambda n: e.run_ticks(n), 2400))
K = 8
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    e.run_ticks(K)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    e.run_ticks(K)
def replay(n):
    for _ in range(n // K):
        g.replay()
print("graph of %d ticks: sims/s %.0f, ms/tick %.4f" % ((K,) + timed(replay, 2400)))
print("stream launches again: sims/s %.0f, ms/tick %.4f" % timed(lambda n: e.run_ticks(n), 2400))
print(e.counters())
