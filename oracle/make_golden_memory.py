"""A replay-memory file written by the UNMODIFIED reference (rl_utils.memory.Memory of games.algos.mcts.Move tuples, pickled the
way UpdateWorker.save_memory does, updateworker.py:119-139), for the on-disk interop test.  TEST INFRASTRUCTURE ONLY.
    python -m oracle.make_golden_memory   ->   tests/golden/ref_memory.pkl  (+ ref_memory_dedup.pkl: after Memory.deduplicate)
"""
import os
import pickle

import numpy as np
import torch

from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref_mcts = rh._import_reference()[0]
    import rl_utils.memory as ref_memory
    rng = np.random.default_rng(7)
    boards = [rng.integers(-1, 2, size=(7, 6)).astype(np.int64) for _ in range(5)]
    mem = ref_memory.Memory(50)
    for i in range(12):
        mem.add(ref_mcts.Move(torch.from_numpy(boards[i % 5].copy()), torch.tensor(float(i % 3 - 1)),
                              torch.from_numpy(rng.dirichlet([1.0] * 7).astype(np.float32)), torch.tensor(float(i) / 16)))
    with open(os.path.join(OUT, "ref_memory.pkl"), "wb") as f:
        pickle.dump(mem, f)
    mem.deduplicate("state", ["actual_val", "tree_probs", "q"], ref_mcts.Move)
    mem.add(ref_mcts.Move(torch.from_numpy(boards[0].copy()), torch.tensor(1.0), torch.full((7,), 1 / 7), torch.tensor(0.5)))
    with open(os.path.join(OUT, "ref_memory_dedup.pkl"), "wb") as f:
        pickle.dump(mem, f)
    print("wrote", OUT, len(mem))


if __name__ == "__main__":
    main()
