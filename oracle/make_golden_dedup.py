"""Golden vectors for the Deduplicator (rl_utils/memory.py:47-94), produced by running the UNMODIFIED reference here.
TEST INFRASTRUCTURE ONLY.    python -m oracle.make_golden_dedup   ->   tests/golden/dedup.npz

The reference's own call site (mcts.py:385-386) raises TypeError (Move has a fourth field `q` that create_memory never fills), so
the reference Memory/Deduplicator is run with `q` listed among the averaged values; nothing else differs from the shipped code.
Three rounds over one stream of 600 Connect4 records drawn from 45 distinct positions: (records [0,380) into a Memory(300), dedup),
(+[380,470), dedup), (+[470,600), dedup with maxlen 30).
"""
import os

import numpy as np
import torch

from . import ref_harness as rh
from . import spec

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dedup.npz")
PHASES = [(0, 380, 0), (380, 470, 0), (470, 600, 30)]      # (first record, end, maxlen; 0 = None)
MAX_SIZE = 300


def main():
    ref_mcts = rh._import_reference()[0]
    import rl_utils.memory as ref_memory
    rng = np.random.default_rng(20261019)
    boards = [rng.integers(-1, 2, size=(7, 6)) for _ in range(44)] + [np.zeros((7, 6), np.int64)]
    n = PHASES[-1][1]
    pick = rng.integers(0, len(boards), size=n)
    pick[rng.random(n) < 0.15] = len(boards) - 1                       # a long run of the empty board
    bits = np.array([spec.board_to_bits(boards[k], 0) for k in pick], dtype=np.uint64)
    probs = rng.dirichlet([1.0] * 7, size=n).astype(np.float32)
    q = rng.uniform(-1, 1, n).astype(np.float32)
    val = rng.integers(-1, 2, n).astype(np.float32)
    mem = ref_memory.Memory(MAX_SIZE)
    out = dict(own=bits[:, 0], opp=bits[:, 1], tree_probs=probs, q=q, actual_val=val, phases=np.array(PHASES, np.int64),
               max_size=np.int64(MAX_SIZE), ply=np.array([int(np.abs(boards[k]).sum()) for k in pick], np.uint8))
    for p, (lo, hi, maxlen) in enumerate(PHASES):
        for i in range(lo, hi):
            mem.add(ref_mcts.Move(torch.from_numpy(boards[pick[i]].astype(np.int64)), torch.tensor(float(val[i])),
                                  torch.from_numpy(probs[i].copy()), torch.tensor(float(q[i]))))
        mem.deduplicate("state", ["actual_val", "tree_probs", "q"], ref_mcts.Move, maxlen=maxlen or None)
        buf = list(mem._buffer)
        b = np.array([spec.board_to_bits(m.state.numpy(), 0) for m in buf], dtype=np.uint64).reshape(-1, 2)
        out[f"own_{p}"], out[f"opp_{p}"] = b[:, 0], b[:, 1]
        out[f"tree_probs_{p}"] = np.stack([m.tree_probs.numpy() for m in buf]).astype(np.float32)
        out[f"actual_val_{p}"] = np.array([float(m.actual_val) for m in buf], np.float32)
        out[f"q_{p}"] = np.array([float(m.q) for m in buf], np.float32)
        out[f"unique_{p}"] = np.int64(len(mem.deduplicator.counter))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("_2") or k == "own"})


if __name__ == "__main__":
    main()
