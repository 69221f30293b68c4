"""How many of a game slot's network evaluations repeat a position the slot has already evaluated (either tree, this game or
an earlier game on the slot)?  The two trees of a game search overlapping subtrees, Connect4 move orders transpose, and every
game starts from the same opening.  Simulates the per-slot evaluation cache (direct-mapped / 2-way, several sizes) on the
logged evaluation sequence of the cache-less engine (the searches are deterministic, so that IS the cached engine's lookup sequence).
    python scripts/dbg_transpositions.py [slots] [sims] [games per slot]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.engine import SelfPlayEngine  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 28
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
per_slot = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
e = SelfPlayEngine(game=0, n_games=G, sims=sims, evaluator=nets.TowerEvaluator(net), seed=0, games_target=G * per_slot)
poll = 4096
need_b = torch.zeros(poll, G, dtype=torch.uint8, device="cuda")
own_b = torch.zeros(poll, G, dtype=torch.int64, device="cuda")
opp_b = torch.zeros(poll, G, dtype=torch.int64, device="cuda")
chunks = []
while not e.all_idle():
    for i in range(poll):
        e.run_ticks(1, fused=True, chunk=1)
        need_b[i].copy_(e.needs_eval); own_b[i].copy_(e.leaf_own); opp_b[i].copy_(e.leaf_opp)
    chunks.append((need_b.cpu().numpy().copy(), own_b.cpu().numpy().copy(), opp_b.cpu().numpy().copy()))
need, own, opp = (np.concatenate([c[k] for c in chunks]) for k in range(3))
own, opp = own.view(np.uint64), opp.view(np.uint64)


def mix(o, p):
    h = (o * np.uint64(0x9E3779B97F4A7C15)) ^ (p * np.uint64(0xC2B2AE3D27D4EB4F))
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    return h


tot = 0
res = {}
for j in range(G):
    idx = np.flatnonzero(need[:, j])
    o, p = own[idx, j], opp[idx, j]
    with np.errstate(over="ignore"):
        h = mix(o, p)
    keys = list(zip(o.tolist(), p.tolist()))
    tot += len(keys)
    seen = set()
    hits = 0
    for k in keys:
        hits += k in seen
        seen.add(k)
    res["unbounded"] = res.get("unbounded", 0) + hits
    for slots in (2048, 4096, 8192, 16384, 65536):
        for ways in (1, 2, 4):
            sets = slots // ways
            table = [[None] * ways for _ in range(sets)]
            stamp = [[0] * ways for _ in range(sets)]
            hits = 0
            hs = (h % np.uint64(sets)).tolist()
            for t, (k, s) in enumerate(zip(keys, hs)):
                row = table[s]
                if k in row:
                    hits += 1
                    stamp[s][row.index(k)] = t
                else:
                    w = min(range(ways), key=lambda x: stamp[s][x])   # LRU within the set
                    row[w] = k; stamp[s][w] = t + 1
            res[(slots, ways)] = res.get((slots, ways), 0) + hits
print(f"{G} slots x {per_slot} games x {sims} sims: {tot} evaluations; hit rates:")
for k, v in res.items():
    print(f"  {k}: {v / tot:.4f}")
