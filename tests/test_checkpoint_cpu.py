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
