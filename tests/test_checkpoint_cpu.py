"""CPU: the reference's on-disk formats ({"model": state_dict} checkpoints, pickled Memory, newest-file discovery)."""
import datetime
import os

import torch

from self_play_reinforcement_learning_b200 import checkpoint, nets
from self_play_reinforcement_learning_b200.scheduler import Memory
from self_play_reinforcement_learning_b200.selfplay import Move


def test_checkpoint_roundtrip_and_discovery(tmp_path):
    save_dir = str(tmp_path)
    torch.manual_seed(0)
    a = nets.ResidualTower(7, 6, 7, num_blocks=1)
    t0 = datetime.datetime(2024, 1, 1, 10, 0, 0)
    for run, games in (("2024-01-01T09:00:00", 750), ("2024-01-02T09:00:00", 1500)):
        for k in range(2):
            name = checkpoint.model_file_name(save_dir, run, games * (k + 1), now=t0 + datetime.timedelta(hours=k))
            checkpoint.save_model(a, name)
    os.makedirs(os.path.join(save_dir, "2024-01-03T00:00:00"))  # empty folder is ignored
    newest = checkpoint.recent_save_file(save_dir, None, False, "model")
    assert "2024-01-02T09:00:00" in newest and newest.endswith(":3000")
    prev = checkpoint.recent_save_file(save_dir, "2024-01-02T09:00:00", True, "model")
    assert "2024-01-01T09:00:00" in prev
    assert set(torch.load(newest).keys()) == {"model"}                      # the reference's checkpoint schema
    torch.manual_seed(1)
    b = nets.ResidualTower(7, 6, 7, num_blocks=1)
    checkpoint.load_model(b, newest)
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    # the packed engine weights of a reloaded checkpoint are identical
    assert torch.equal(nets.pack_tower_blob(a), nets.pack_tower_blob(b))


def test_memory_pickle_roundtrip(tmp_path):
    m = Memory(10)
    for i in range(4):
        m.add(Move(torch.zeros(7, 6, dtype=torch.int64), torch.tensor(1.0), torch.full((7,), 1 / 7), torch.tensor(0.25 * i)))
    f1 = checkpoint.save_memory(m, str(tmp_path), "run", now=datetime.datetime(2024, 1, 1))
    m.add(Move(torch.ones(7, 6, dtype=torch.int64), torch.tensor(-1.0), torch.full((7,), 1 / 7), torch.tensor(0.0)))
    f2 = checkpoint.save_memory(m, str(tmp_path), "run", previous=f1, now=datetime.datetime(2024, 1, 2))
    assert not os.path.exists(f1) and f2.endswith(":5")
    m2 = checkpoint.load_memory(checkpoint.recent_save_file(str(tmp_path), None, False, "memory"))
    assert len(m2) == 5 and float(m2.sample(5)[0].tree_probs.sum()) > 0.99


def test_moves_from_records_pickle_independently():
    """records_to_moves must hand out tensors that own their storage: a Move that is a view into the batch would serialise the
    whole batch with every record (mp queues, save_memory) -- 200 k records once meant 200 k copies of a 67 MB tensor."""
    import pickle
    import numpy as np
    from self_play_reinforcement_learning_b200.engine import RECORD_DTYPE
    from self_play_reinforcement_learning_b200.selfplay import records_to_moves
    r = np.zeros(5000, RECORD_DTYPE)
    r["own"] = np.arange(5000, dtype=np.uint64) & np.uint64(0x3F)
    r["tree_probs"][:, :7] = 1 / 7
    moves = records_to_moves(r, 0)
    assert len(pickle.dumps(moves[17])) < 4000
    m = Memory(10000)
    for mv in moves:
        m.add(mv)
    assert len(pickle.dumps(m)) < 5000 * 4000
    assert moves[63].state.dtype == torch.int64 and int(moves[63].state.sum()) == 6 and moves[63].state.shape == (7, 6)


def test_memory_file_written_by_the_unmodified_reference_loads_here(golden_dir):
    """tests/golden/ref_memory*.pkl were pickled by the reference's own rl_utils.memory.Memory (oracle/make_golden_memory.py).
    They must open WITHOUT the reference tree (or anytree) importable, as local Memory / Move objects."""
    import subprocess
    import sys
    code = (
        "import sys, os, torch\n"
        "assert not any('reference' in p for p in sys.path)\n"
        "from self_play_reinforcement_learning_b200 import checkpoint\n"
        "from self_play_reinforcement_learning_b200.scheduler import Memory\n"
        "from self_play_reinforcement_learning_b200.selfplay import Move\n"
        f"m = checkpoint.load_memory(os.path.join({golden_dir!r}, 'ref_memory.pkl'))\n"
        "assert type(m) is Memory and len(m) == 12 and m.max_size == 50 and m._buffer.maxlen == 50\n"
        "assert all(type(x) is Move for x in m._buffer) and m._buffer[3].state.dtype == torch.int64\n"
        "assert abs(float(m._buffer[5].q) - 5 / 16) < 1e-7 and len(m.sample(4)) == 4\n"
        f"d = checkpoint.load_memory(os.path.join({golden_dir!r}, 'ref_memory_dedup.pkl'))\n"
        "assert len(d) == 6 and len(d._dedup) == 5 and len(d._dedup_pending) == 1\n"
        "d.deduplicate('state', ['actual_val', 'tree_probs'], Move)\n"      # the reference's running sums carry on
        "assert len(d) == 5 and sorted(c['count'] for c in d._dedup.values()) == [2, 2, 2, 3, 4]\n"
        "assert 'rl_utils' not in sys.modules and 'games' not in sys.modules and 'anytree' not in sys.modules\n"
        "print('ok')\n")
    import os as _os
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env={**_os.environ, "PYTHONPATH": root})
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_memory_file_written_here_carries_the_reference_class_paths(tmp_path):
    """What save_memory writes names rl_utils.memory.Memory / games.algos.mcts.Move (and nothing of this package), so the
    unmodified reference opens it with plain pickle.load (base_worker.py:36-42)."""
    import pickletools
    import sys
    m = Memory(10)
    for i in range(3):
        m.add(Move(torch.full((7, 6), i, dtype=torch.int64), torch.tensor(1.0), torch.full((7,), 1 / 7), torch.tensor(0.5)))
    f = checkpoint.save_memory(m, str(tmp_path), "run", now=datetime.datetime(2024, 1, 1))
    assert "rl_utils" not in sys.modules and "games" not in sys.modules       # the stand-in modules are gone again
    globs = set()
    for op, arg, _ in pickletools.genops(open(f, "rb").read()):
        if op.name in ("GLOBAL", "STACK_GLOBAL") and arg:
            globs.add(arg)
    blob = open(f, "rb").read()
    assert b"rl_utils.memory" in blob and b"games.algos.mcts" in blob and b"self_play_reinforcement_learning_b200" not in blob
    back = checkpoint.load_memory(f)
    assert type(back) is Memory and len(back) == 3 and int(back._buffer[2].state[0, 0]) == 2 and back._buffer.maxlen == 10
    # live check where the reference exists (this container): its own classes, plain pickle.load
    from oracle import ref_harness as rh
    if rh.reference_available():
        import subprocess
        code = (f"import sys; sys.path[:0] = ['/root/reference', {os.path.join(os.path.dirname(rh.__file__), '_shims')!r}]\n"
                "import pickle, rl_utils.memory, games.algos.mcts\n"
                f"m = pickle.load(open({f!r}, 'rb'))\n"
                "assert type(m) is rl_utils.memory.Memory and len(m) == 3 and m.deduplicator is None and m.max_size == 10\n"
                "assert type(m._buffer[0]) is games.algos.mcts.Move and len(m.sample(2)) == 2\n"
                "m.add(m._buffer[0]); print('ok')\n")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
        assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
