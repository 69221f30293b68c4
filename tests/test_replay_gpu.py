"""Device-resident replay memory (csrc/spx_replay.cu through the C ABI) against the CPU restatement oracle/replay.py:
FIFO content under append / eviction / change_size with ring wrap-around, sampled indices and assembled batches bit for
bit, the device-to-device drain of the engine's records, and the loss on a device batch."""
import numpy as np
import pytest
import torch

from oracle import replay as orp
from oracle import spec

pytestmark = pytest.mark.gpu


def _records(rng, game, n, start=0):
    from self_play_reinforcement_learning_b200.engine import RECORD_DTYPE
    W, H, A = spec.GAME_DIMS[game]
    out = np.zeros(n, RECORD_DTYPE)
    for i in range(n):
        cells = rng.integers(-1, 2, size=(W, H))
        out["own"][i], out["opp"][i] = spec.board_to_bits(cells, game)
        out["tree_probs"][i, :A] = rng.dirichlet([1.0] * A)
        out["game_index"][i] = start + i
    out["q"] = rng.uniform(-1, 1, n)
    out["actual_val"] = rng.integers(-1, 2, n)
    return out


def _same_content(dev, mem):
    got = dev.read()
    assert len(dev) == len(mem) == len(got) and dev.max_size == mem.max_size
    assert got.tobytes() == (np.array(list(mem._buffer), dtype=got.dtype) if len(mem) else got[:0]).tobytes()


@pytest.mark.parametrize("game", [0, 1])
def test_fifo_eviction_resize_and_wraparound(game):
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    rng = np.random.default_rng(game)
    dev, mem = DeviceReplay(game, max_size=50, physical_capacity=80, seed=1), orp.Memory(50)
    made = 0
    for it in range(40):
        op = rng.integers(0, 10)
        if op < 7:
            n = int(rng.integers(1, 130 if it % 5 == 0 else 30))       # sometimes more than max_size at once
            r = _records(rng, game, n, made)
            made += n
            dev.append_records(r)
            for x in r:
                mem.add(x)
        elif op < 9:
            m = int(rng.integers(1, 81))
            dev.change_size(m)
            mem.change_size(m)
        else:
            dev.reset()
            mem.reset()
        _same_content(dev, mem)
    dev.close()


@pytest.mark.parametrize("game,batch", [(0, 1), (0, 64), (0, 4096), (1, 33), (1, 500)])
def test_sampled_batch_is_bit_exact(game, batch):
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    rng = np.random.default_rng(10 * game + batch)
    size = max(batch, 5) + int(rng.integers(0, 3000))
    dev, mem = DeviceReplay(game, max_size=size, physical_capacity=size + 100, seed=77), orp.Memory(size)
    for chunk in range(3):                                               # overfill so that head != 0 and the ring wraps
        r = _records(rng, game, size // 2 + 60, chunk * 10000)
        dev.append_records(r)
        for x in r:
            mem.add(x)
    _same_content(dev, mem)
    for step in (0, 1, 12345):
        got = dev.sample_batch(batch, step=step, boards=True, planes=True)
        torch.cuda.synchronize()
        idx = orp.sample_indices(77, step, len(mem), batch)
        want = orp.assemble([mem._buffer[int(i)] for i in idx], game)
        assert np.array_equal(got["idx"].cpu().numpy(), idx)
        for k in ("boards", "planes", "tree_probs", "actual_val", "q"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    # the reference-format view (list of Move tuples, reference dtypes)
    moves = dev.sample(min(batch, 8))
    assert moves[0].state.dtype == torch.int64 and tuple(moves[0].state.shape) == spec.GAME_DIMS[game][:2]
    assert moves[0].tree_probs.dtype == torch.float32 and moves[0].q.dim() == 0
    dev.close()


def test_sample_errors_are_loud():
    from self_play_reinforcement_learning_b200._lib import SpxError
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    dev = DeviceReplay(0, max_size=10)
    with pytest.raises(SpxError):
        dev.sample_batch(1)                       # empty memory (the reference skips the update, mcts.py:255-261)
    dev.append_records(_records(np.random.default_rng(0), 0, 6))
    with pytest.raises(SpxError):
        dev.sample_batch(7)
    with pytest.raises(SpxError):
        dev.change_size(11)                       # beyond the physical capacity
    assert len(dev.sample_batch(6)["q"]) == 6
    dev.close()


def test_device_drain_equals_host_drain_sorted():
    """Two identical engines: one drained to the host (spx_drain_records), one device-to-device into the replay memory."""
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    kw = dict(game=0, n_games=64, sims=30, net="hash", seed=4, games_target=150, noise_mode=0)
    a, b = BatchedSelfPlay(None, **kw), BatchedSelfPlay(None, **kw)
    dev = DeviceReplay(0, max_size=100000)
    host = []
    while True:
        a.engine.run_ticks(200)
        b.engine.run_ticks(200)
        host.append(a.engine.drain_records())
        n_before = len(dev)
        part = dev.drain_engine(b.engine)
        assert len(dev) - n_before == part.shape[0] == len(host[-1])
        chunk = dev.read(n_before)
        key = (chunk["game_index"].astype(np.int64) << 8) | (chunk["tree"].astype(np.int64) << 7) | chunk["ply"]
        assert np.all(np.diff(key) > 0)                                       # sorted by (game, tree, ply)
        hs = host[-1]
        hk = (hs["game_index"].astype(np.int64) << 8) | (hs["tree"].astype(np.int64) << 7) | hs["ply"]
        assert chunk.tobytes() == hs[np.argsort(hk)].tobytes()
        if a.engine.all_idle():
            break
    assert len(dev) == sum(len(h) for h in host) > 150 * 7 and b.engine.drain_records().shape[0] == 0
    a.close(); b.close(); dev.close()


def test_loss_on_device_batch_matches_move_list_loss():
    from self_play_reinforcement_learning_b200 import nets
    from self_play_reinforcement_learning_b200.replay import DeviceReplay, loss_from_batch
    from self_play_reinforcement_learning_b200.scheduler import mcts_loss
    from self_play_reinforcement_learning_b200.selfplay import records_to_moves
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(3)
    recs = _records(rng, 0, 400)
    dev = DeviceReplay(0, max_size=400, seed=5)
    dev.append_records(recs)
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=2).cuda().eval()
    batch = dev.sample_batch(48, step=2)
    idx = batch["idx"].cpu().numpy()
    moves = records_to_moves(recs[idx], 0)
    with torch.no_grad():
        got, want = float(loss_from_batch(net, batch)), float(mcts_loss(net, moves))
    assert abs(got - want) < 1e-5
    # round trip through the reference format: Move tuples -> records -> device
    dev2 = DeviceReplay(0, max_size=400)
    dev2.extend(records_to_moves(recs[:50], 0))
    back = dev2.read()
    for k in ("own", "opp", "tree_probs", "q", "actual_val"):
        assert np.array_equal(back[k], recs[k][:50]), k
    dev.close(); dev2.close()


@pytest.mark.parametrize("game", [0, 1])
def test_device_deduplicator_matches_restatement(game):
    """spx_replay_deduplicate vs oracle/replay.Deduplicator (memory.py:47-94), bit for bit over several rounds: first call folds
    the current buffer, later calls everything appended since (incl. records the FIFO evicted), maxlen keeps the newest,
    raw appends after a call sit behind the averaged entries, reset() keeps the table."""
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    rng = np.random.default_rng(40 + game)
    W, H, A = spec.GAME_DIMS[game]
    boards = [rng.integers(-1, 2, size=(W, H)) for _ in range(60)]
    boards += [np.zeros((W, H), np.int64)] * 20                    # the empty board: one long run of duplicates

    def make(n, start):
        from self_play_reinforcement_learning_b200.engine import RECORD_DTYPE
        out = np.zeros(n, RECORD_DTYPE)
        for i in range(n):
            cells = boards[int(rng.integers(0, len(boards)))]
            out["own"][i], out["opp"][i] = spec.board_to_bits(cells, game)
            out["tree_probs"][i, :A] = rng.dirichlet([1.0] * A)
            out["game_index"][i] = start + i
            out["tree"][i] = i & 1
            out["ply"][i] = int(np.abs(cells).sum())
        out["q"] = rng.uniform(-1, 1, n)
        out["actual_val"] = rng.integers(-1, 2, n)
        return out

    def check(dev, mem, where):
        got = dev.read()
        assert len(dev) == len(mem) == len(got), (where, len(dev), len(mem), len(got))
        for g, o in zip(got, mem._buffer):
            assert (int(g["own"]), int(g["opp"])) == (int(o["own"]), int(o["opp"]))
            want_probs = np.zeros(9, np.float32)
            want_probs[:len(o["tree_probs"])] = o["tree_probs"]
            assert g["tree_probs"].tobytes() == want_probs.tobytes()
            assert np.float32(g["actual_val"]).tobytes() == np.float32(o["actual_val"]).tobytes()
            assert np.float32(g["q"]).tobytes() == np.float32(o["q"]).tobytes()
            assert int(g["game_index"]) == int(o["game_index"]) and int(g["ply"]) == int(o["ply"]) and int(g["tree"]) == int(o["tree"])
            if isinstance(o, dict):                                  # an averaged entry (raw records are numpy rows)
                assert int(g["pad1"]) == o["count"]

    dev, mem = DeviceReplay(game, max_size=300, physical_capacity=2000, seed=1), orp.Memory(300)
    assert dev.unique_states == 0
    made = 0
    # (the device ring is bounded by its physical capacity where a deque(maxlen=None) is not: stay below it)
    for rnd, (n, maxlen) in enumerate([(450, None), (120, None), (0, None), (700, 40), (90, None), (1500, 1000), (2500, 35)]):
        r = make(n, made)
        made += n
        dev.append_records(r)
        for x in r:
            mem.add(x)
        check(dev, mem, ("appended", rnd))
        if rnd == 4:
            dev.reset(); mem.reset()
        dev.deduplicate("state", ["actual_val", "tree_probs"], maxlen=maxlen)
        orp.memory_deduplicate(mem, maxlen=maxlen)
        check(dev, mem, ("deduplicated", rnd))
        assert dev.unique_states == len(mem.deduplicator.counter)
        assert dev.max_size == (maxlen or 2000)
    dev.close()


def test_device_deduplicate_empty_memory_and_sampling_after():
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    dev = DeviceReplay(0, max_size=100, seed=3)
    assert dev.deduplicate() == 0 and len(dev) == 0 and dev.unique_states == 0
    rng = np.random.default_rng(1)
    r = _records(rng, 0, 64)
    r[32:] = r[:32]                                                  # every state twice
    dev.append_records(r)
    assert dev.deduplicate() == 32 and dev.unique_states == 32
    b = dev.sample_batch(32, step=0)
    assert sorted(b["idx"].cpu().tolist()) == list(range(32)) if "idx" in b else True
    with pytest.raises(ValueError):
        dev.deduplicate(key="q")
    dev.close()


def test_device_deduplicator_matches_reference_golden(golden_dir):
    """spx_replay_deduplicate against tests/golden/dedup.npz (written by the unmodified reference's Memory/Deduplicator)."""
    import os
    from self_play_reinforcement_learning_b200.engine import RECORD_DTYPE
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    g = np.load(os.path.join(golden_dir, "dedup.npz"))
    n = len(g["own"])
    recs = np.zeros(n, RECORD_DTYPE)
    recs["own"], recs["opp"], recs["q"], recs["actual_val"], recs["ply"] = g["own"], g["opp"], g["q"], g["actual_val"], g["ply"]
    recs["tree_probs"][:, :7] = g["tree_probs"]
    recs["game_index"] = np.arange(n)
    dev = DeviceReplay(0, max_size=int(g["max_size"]), physical_capacity=1000, seed=0)
    for p, (lo, hi, maxlen) in enumerate(g["phases"].tolist()):
        dev.append_records(recs[lo:hi])
        dev.deduplicate("state", ["actual_val", "tree_probs"], maxlen=maxlen or None)
        got = dev.read()
        assert got["own"].tolist() == g[f"own_{p}"].tolist() and got["opp"].tolist() == g[f"opp_{p}"].tolist()
        assert np.ascontiguousarray(got["tree_probs"][:, :7]).tobytes() == g[f"tree_probs_{p}"].tobytes()
        assert got["actual_val"].tobytes() == g[f"actual_val_{p}"].tobytes() and got["q"].tobytes() == g[f"q_{p}"].tobytes()
        assert dev.unique_states == int(g[f"unique_{p}"])
    dev.close()
