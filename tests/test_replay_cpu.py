"""CPU: the replay-memory restatement (oracle/replay.py) -- index stream properties, deque semantics, batch assembly; and,
where /root/reference exists, the same against the reference's own Memory with np.random.choice hooked to the stream."""
import os
from collections import Counter

import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from oracle import replay as orp
from oracle import spec


def test_sample_indices_are_distinct_in_range_deterministic():
    for size, batch in [(1, 1), (5, 5), (64, 64), (1000, 128), (200000, 4096), (4097, 4096)]:
        a = orp.sample_indices(7, 3, size, batch)
        assert a.dtype == np.int64 and len(a) == batch and len(set(a.tolist())) == batch
        assert a.min() >= 0 and a.max() < size
        assert np.array_equal(a, orp.sample_indices(7, 3, size, batch))
        if size > 4:
            assert not np.array_equal(a, orp.sample_indices(7, 4, size, batch))
    assert sorted(orp.sample_indices(1, 0, 9, 9).tolist()) == list(range(9))   # batch == size: a permutation
    # prefix property of the partial Fisher-Yates: a smaller batch is a prefix of a larger one from the same stream
    assert np.array_equal(orp.sample_indices(5, 9, 500, 20), orp.sample_indices(5, 9, 500, 60)[:20])


def test_sample_indices_uniform():
    size, batch, steps = 20, 5, 4000
    cnt = Counter()
    for s in range(steps):
        cnt.update(orp.sample_indices(11, s, size, batch).tolist())
    exp = steps * batch / size
    chi2 = sum((cnt[k] - exp) ** 2 / exp for k in range(size))
    assert chi2 < 50.0      # 19 dof: P(chi2 > 50) ~ 1e-4


def test_memory_is_a_bounded_fifo():
    m = orp.Memory(5)
    for i in range(8):
        m.add(i)
    assert list(m._buffer) == [3, 4, 5, 6, 7]
    m.change_size(3)
    assert list(m._buffer) == [5, 6, 7] and m.max_size == 3
    m.change_size(6)
    for i in range(8, 12):
        m.add(i)
    assert list(m._buffer) == [6, 7, 8, 9, 10, 11]
    assert sorted(m.sample(6, 0, 0)) == [6, 7, 8, 9, 10, 11]
    m.reset()
    assert len(m) == 0 and m.max_size == 6


def _random_records(rng, game, n):
    W, H, A = spec.GAME_DIMS[game]
    recs = []
    for i in range(n):
        cells = rng.integers(-1, 2, size=(W, H))
        own, opp = spec.board_to_bits(cells, game)
        recs.append(dict(own=own, opp=opp, tree_probs=rng.dirichlet([1.0] * A).astype(np.float32), q=np.float32(rng.uniform(-1, 1)),
                         actual_val=np.float32(rng.integers(-1, 2)), board=cells))
    return recs


@pytest.mark.parametrize("game", [0, 1])
def test_assemble_matches_preprocess_semantics(game):
    rng = np.random.default_rng(game)
    recs = _random_records(rng, game, 12)
    b = orp.assemble(recs, game)
    W, H, A = spec.GAME_DIMS[game]
    assert b["boards"].shape == (12, W, H) and b["planes"].shape == (12, 3, W, H) and b["tree_probs"].shape == (12, A)
    for i, r in enumerate(recs):
        assert np.array_equal(b["boards"][i], r["board"])
        assert np.array_equal(b["planes"][i, 0], r["board"] == 0) and np.array_equal(b["planes"][i, 1], r["board"] == 1)
        assert np.array_equal(b["planes"][i, 2], r["board"] == -1)


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present (GPU box)")
def test_reference_memory_with_hooked_choice_samples_the_same_batch():
    """rl_utils/memory.py:8-33 run live: same FIFO content after add/change_size, and with np.random.choice hooked to the
    counter stream the same sampled batch; the reference's own loss on that batch == the loss from the assembled tensors."""
    ref_mcts = rh._import_reference()[0]
    import rl_utils.memory as ref_memory
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.replay import loss_from_batch
    rng = np.random.default_rng(5)
    recs = _random_records(rng, 0, 300)
    ref_mem, mem = ref_memory.Memory(200), orp.Memory(200)
    for r in recs:
        ref_mem.add(ref_mcts.Move(torch.from_numpy(r["board"].astype(np.int64)), torch.tensor(float(r["actual_val"])),
                                  torch.from_numpy(r["tree_probs"]), torch.tensor(float(r["q"]))))
        mem.add(r)
    ref_mem.change_size(150)
    mem.change_size(150)
    assert len(ref_mem) == len(mem) == 150
    seed, step, batch = 9, 4, 32
    real_choice = ref_memory.np.random.choice

    def choice(a, size=None, replace=True, p=None):
        assert replace is False and p is None
        return np.asarray(a)[orp.sample_indices(seed, step, len(a), size)]
    ref_memory.np.random.choice = choice
    try:
        ref_batch = ref_mem.sample(batch)
    finally:
        ref_memory.np.random.choice = real_choice
    ours = orp.assemble(mem.sample(batch, seed, step), 0)
    assert np.array_equal(np.stack([m.state.numpy() for m in ref_batch]), ours["boards"])
    assert np.array_equal(np.stack([m.tree_probs.numpy() for m in ref_batch]), ours["tree_probs"])
    assert np.array_equal(np.array([float(m.q) for m in ref_batch], np.float32), ours["q"])
    torch.manual_seed(3)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()

    class Holder:
        network, q_average = net, True
    with torch.no_grad():
        want = float(ref_mcts.MCTreeSearch.loss(Holder, ref_batch))
        got = float(loss_from_batch(net, {k: torch.from_numpy(v) for k, v in ours.items()}))
    assert abs(want - got) < 1e-5
