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


def _records_with_duplicates(rng, game, n, n_distinct):
    """n records over n_distinct positions (so that most keys repeat), fresh values each time."""
    W, H, A = spec.GAME_DIMS[game]
    boards = [rng.integers(-1, 2, size=(W, H)) for _ in range(n_distinct)]
    recs = []
    for i in range(n):
        cells = boards[int(rng.integers(0, n_distinct))]
        own, opp = spec.board_to_bits(cells, game)
        recs.append(dict(own=own, opp=opp, game_index=i, tree=int(i & 1), ply=int(np.abs(cells).sum()),
                         tree_probs=rng.dirichlet([1.0] * A).astype(np.float32), q=np.float32(rng.uniform(-1, 1)),
                         actual_val=np.float32(rng.integers(-1, 2)), board=cells))
    return recs


def test_deduplicator_restatement_properties():
    """oracle/replay.Deduplicator: one entry per distinct state in first-seen order, counts add up, values are the float32
    running sums divided by the count, the table persists across calls and sees records the FIFO has already evicted."""
    rng = np.random.default_rng(2)
    recs = _records_with_duplicates(rng, 0, 400, 37)
    mem = orp.Memory(250)
    for r in recs[:300]:
        mem.add(r)
    out = orp.memory_deduplicate(mem)
    last250 = recs[50:300]
    keys = []
    for r in last250:
        if (r["own"], r["opp"]) not in keys:
            keys.append((r["own"], r["opp"]))
    assert [(o["own"], o["opp"]) for o in out] == keys and sum(o["count"] for o in out) == 250
    k0 = keys[0]
    members = [r for r in last250 if (r["own"], r["opp"]) == k0]
    acc = np.float32(0)
    for i, r in enumerate(members):
        acc = np.float32(r["actual_val"]) if i == 0 else np.float32(acc + np.float32(r["actual_val"]))
    assert out[0]["actual_val"] == np.float32(acc / np.float32(len(members))) and out[0]["count"] == len(members)
    for r in recs[300:]:
        mem.add(r)                                   # raw appends after the first call land behind the averaged entries
    assert len(mem) == len(keys) + 100
    out2 = orp.memory_deduplicate(mem, maxlen=20)
    assert sum(c["count"] for c in mem.deduplicator.counter.values()) == 350 and len(out2) == min(20, len(mem.deduplicator.counter))
    assert [(o["own"], o["opp"]) for o in out2] == list(mem.deduplicator.counter.keys())[-20:]


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present (GPU box)")
def test_deduplicator_matches_reference_live():
    """rl_utils/memory.py:47-94 run live.  The reference's own call (mcts.py:385-386) raises TypeError because Move has a
    fourth field `q` that create_memory never fills, so the live run uses the same Memory/Deduplicator with `q` listed as a
    value (the generic algorithm is unchanged); its output must equal the restatement and the host Memory bit for bit."""
    ref_mcts = rh._import_reference()[0]
    import rl_utils.memory as ref_memory
    from self_play_reinforcement_learning_b200.scheduler import Memory as HostMemory
    rng = np.random.default_rng(8)
    recs = _records_with_duplicates(rng, 0, 500, 41)
    ref_mem, mem, host = ref_memory.Memory(300), orp.Memory(300), HostMemory(300)

    def as_move(r):   # fresh tensors: the reference's `count[value] += ...` adds IN PLACE into the first occurrence's tensors
        return ref_mcts.Move(torch.from_numpy(r["board"].astype(np.int64)), torch.tensor(float(r["actual_val"])),
                             torch.from_numpy(r["tree_probs"].copy()), torch.tensor(float(r["q"])))
    with pytest.raises(TypeError):
        probe = ref_memory.Memory(10)
        probe.add(as_move(recs[0]))
        probe.deduplicate("state", ["actual_val", "tree_probs"], ref_mcts.Move)
    for phase, (lo, hi, maxlen) in enumerate([(0, 350, None), (350, 430, None), (430, 500, 25)]):
        for r in recs[lo:hi]:
            ref_mem.add(as_move(r)); mem.add(r); host.add(as_move(r))
        ref_mem.deduplicate("state", ["actual_val", "tree_probs", "q"], ref_mcts.Move, maxlen=maxlen)
        ours = orp.memory_deduplicate(mem, maxlen=maxlen)
        host.deduplicate("state", ["actual_val", "tree_probs"], ref_mcts.Move, maxlen=maxlen)
        assert len(ref_mem) == len(ours) == len(host)
        for m, o, hm in zip(ref_mem._buffer, ours, host._buffer):
            assert spec.board_to_bits(m.state.numpy(), 0) == (o["own"], o["opp"])
            assert np.array_equal(m.tree_probs.numpy(), o["tree_probs"][:7]) and np.float32(m.actual_val) == o["actual_val"]
            assert np.float32(m.q) == o["q"]
            assert torch.equal(hm.state, m.state) and torch.equal(hm.tree_probs, m.tree_probs)
            assert torch.equal(hm.actual_val, m.actual_val) and torch.equal(hm.q, m.q)


def _golden_dedup_records(g):
    return [dict(own=int(g["own"][i]), opp=int(g["opp"][i]), game_index=i, tree=0, ply=int(g["ply"][i]), tree_probs=g["tree_probs"][i],
                 q=g["q"][i], actual_val=g["actual_val"][i]) for i in range(len(g["own"]))]


def test_deduplicator_restatement_matches_reference_golden(golden_dir):
    """tests/golden/dedup.npz was written by the unmodified reference (oracle/make_golden_dedup.py); the restatement must
    reproduce every round bit for bit (this pin travels to the GPU box, where the reference does not exist)."""
    g = np.load(os.path.join(golden_dir, "dedup.npz"))
    recs = _golden_dedup_records(g)
    mem = orp.Memory(int(g["max_size"]))
    for p, (lo, hi, maxlen) in enumerate(g["phases"].tolist()):
        for r in recs[lo:hi]:
            mem.add(r)
        out = orp.memory_deduplicate(mem, maxlen=maxlen or None)
        assert [o["own"] for o in out] == g[f"own_{p}"].tolist() and [o["opp"] for o in out] == g[f"opp_{p}"].tolist()
        assert np.stack([o["tree_probs"][:7] for o in out]).tobytes() == g[f"tree_probs_{p}"].tobytes()
        assert np.array([o["actual_val"] for o in out], np.float32).tobytes() == g[f"actual_val_{p}"].tobytes()
        assert np.array([o["q"] for o in out], np.float32).tobytes() == g[f"q_{p}"].tobytes()
        assert len(mem.deduplicator.counter) == int(g[f"unique_{p}"])
