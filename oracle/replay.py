"""CPU restatement of the reference's replay memory and loss-batch assembly.  TEST INFRASTRUCTURE ONLY (see oracle/README
note in DESIGN.md section 4): the product path is csrc/spx_replay.cu.

Follows rl_utils/memory.py:8-30 (Memory: deque(maxlen), add, change_size, reset, sample = uniform without replacement) and
games/algos/mcts.py:234-243 (stacking of a sampled batch) + games/general/modules.py:115-125 (preprocess planes).

The one thing that cannot be shared with the reference is numpy's global MT19937 (memory.py:28 np.random.choice): as for the
search (oracle/spec.py) the randomness is injected -- `sample_indices` defines the index stream, the live test
(tests/test_oracle_vs_reference_live.py) hooks np.random.choice in the reference's Memory with it and compares batches.
"""
from collections import deque

import numpy as np

from . import spec

PURPOSE_SAMPLE = 4


def sample_indices(seed, step, size, batch):
    """`batch` distinct indices in [0, size): partial Fisher-Yates over the virtual array a[k] = k, draw i taken from the
    counter stream as j = i + rng_u64(seed, step, 0, PURPOSE_SAMPLE, 0, i, 0, 0) % (size - i); result[i] = a[j], a[j] = a[i]."""
    assert 1 <= batch <= size
    moved = {}
    out = np.empty(batch, dtype=np.int64)
    for i in range(batch):
        j = i + spec.rng_u64(seed, step, 0, PURPOSE_SAMPLE, 0, i, 0, 0) % (size - i)
        vi, vj = moved.get(i, i), moved.get(j, j)
        out[i] = vj
        moved[j] = vi
    return out


class Memory:
    """memory.py:8-33 over plain record dicts / structured rows."""

    def __init__(self, max_size=None):
        self.max_size = max_size
        self._buffer = deque(maxlen=max_size)

    def __len__(self):
        return len(self._buffer)

    def add(self, experience):
        self._buffer.append(experience)

    def change_size(self, max_size):
        self.max_size = max_size
        self._buffer = deque(self._buffer, maxlen=max_size)

    def reset(self):
        self._buffer = deque(maxlen=self.max_size)

    def sample(self, batch_size, seed, step):
        return [self._buffer[int(i)] for i in sample_indices(seed, step, len(self._buffer), batch_size)]


class Deduplicator:
    """memory.py:56-94 over record dicts (key = the (own, opp) bitboards = the bytes of Move.state; values = tree_probs,
    actual_val and q as float32, summed one by one in insertion order like `count[value] += getattr(experience, value)` on
    float32 tensors, then divided by the count like `tensor / int`).  `counter` keeps first-insertion order (a dict)."""
    VALUES = ("tree_probs", "actual_val", "q")

    def __init__(self, buffer=()):
        self.counter = {}
        self.temp_queue = list(buffer)                                        # :62

    def add_temp(self, experience):                                           # :70-71
        self.temp_queue.append(experience)

    def add(self, r):                                                         # :73-84
        k = (int(r["own"]), int(r["opp"]))
        c = self.counter.get(k)
        if c is None:
            self.counter[k] = dict(count=1, first=r, tree_probs=np.asarray(r["tree_probs"], np.float32).copy(),
                                   actual_val=np.float32(r["actual_val"]), q=np.float32(r["q"]))
        else:
            c["count"] += 1
            c["tree_probs"] = (c["tree_probs"] + np.asarray(r["tree_probs"], np.float32)).astype(np.float32)
            c["actual_val"] = np.float32(c["actual_val"] + np.float32(r["actual_val"]))
            c["q"] = np.float32(c["q"] + np.float32(r["q"]))

    def deduplicate(self, max_size=None):                                     # :64-68, :86-94
        for r in self.temp_queue:
            self.add(r)
        self.temp_queue = []
        out = deque(maxlen=max_size)
        for (own, opp), c in self.counter.items():
            n = np.float32(c["count"])
            out.append(dict(own=own, opp=opp, count=c["count"], game_index=int(c["first"]["game_index"]), tree=int(c["first"]["tree"]),
                            ply=int(c["first"]["ply"]), tree_probs=(c["tree_probs"] / n).astype(np.float32),
                            actual_val=np.float32(c["actual_val"] / n), q=np.float32(c["q"] / n)))
        return out


def memory_deduplicate(memory, maxlen=None):
    """Memory.deduplicate (memory.py:47-54) + the add_temp hook of Memory.add (:19-20) for the oracle Memory above."""
    if getattr(memory, "deduplicator", None) is None:
        memory.deduplicator = Deduplicator(memory._buffer)
        plain_add = memory.add

        def add(experience):
            plain_add(experience)
            memory.deduplicator.add_temp(experience)
        memory.add = add
    memory._buffer = memory.deduplicator.deduplicate(max_size=maxlen)
    return memory._buffer


def assemble(records, game):
    """records: rows with own/opp/tree_probs/q/actual_val -> the tensors MCTreeSearch.loss stacks (mcts.py:236-243)."""
    W, H, A = spec.GAME_DIMS[game]
    boards = np.stack([spec.bits_to_board(int(r["own"]), int(r["opp"]), game) for r in records]).astype(np.int64)
    planes = np.stack([(boards == 0), (boards == 1), (boards == -1)], axis=1).astype(np.float32)   # modules.py:115-125
    probs = np.stack([np.asarray(r["tree_probs"], dtype=np.float32)[:A] for r in records])
    return dict(boards=boards, planes=planes, tree_probs=probs, actual_val=np.array([r["actual_val"] for r in records], np.float32),
                q=np.array([r["q"] for r in records], np.float32))
