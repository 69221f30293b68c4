"""World-size-2 gloo tests of the multi-GPU host logic (sharding, weight broadcast, record/result gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from self_play_reinforcement_learning_b200 import nets, parallel
from self_play_reinforcement_learning_b200.engine import RECORD_DTYPE, RESULT_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. weights: rank 0 packs, everyone receives the identical blob
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=1).eval()
    ref = nets.pack_tower_blob(net)
    blob = ref.clone() if rank == 0 else torch.zeros_like(ref)
    parallel.broadcast_blob(blob, src=0)
    assert torch.equal(blob, ref)
    # 2. sharding: disjoint slots, every global game owned by exactly one rank
    G = 8
    off, stride = parallel.shard(rank, world, G)
    assert (off, stride) == (rank * G, world * G)
    mine = [g for g in range(64) if parallel.owner_of_game(g, world, G) == rank]
    assert mine == [g for g in range(64) if off <= g % stride < off + G]
    # 3. records / results: variable-length gather, ordered by rank, only dst receives
    recs = np.zeros(3 + 2 * rank, RECORD_DTYPE)
    recs["game_index"] = np.arange(len(recs)) * world + rank
    recs["q"] = rank + 0.5
    res = np.zeros(2, RESULT_DTYPE)
    res["game_index"] = [rank + 2, rank]
    res["reward"] = [1, -1]
    all_recs = parallel.gather_structured(recs, dst=0)
    all_res = parallel.gather_structured(res, dst=0)
    # 3b. the same gather for device-resident record rows ([n, 80] uint8 tensors; CPU tensors under gloo)
    rows = torch.from_numpy(recs.view(np.uint8).reshape(len(recs), RECORD_DTYPE.itemsize).copy())
    parts = parallel.gather_device_rows(rows, dst=0)
    if rank == 0:
        back = np.concatenate([p.numpy().reshape(-1).view(RECORD_DTYPE) for p in parts])
        assert [len(p) for p in parts] == [3, 5] and back.tobytes() == all_recs.tobytes()
    else:
        assert parts is None
    tot = parallel.reduce_counters({"sims": 10 * (rank + 1), "moves": rank})
    assert tot == {"sims": 30, "moves": 1}
    if rank == 0:
        assert len(all_recs) == 3 + 5 and (all_recs["q"][:3] == 0.5).all() and (all_recs["q"][3:] == 1.5).all()
        ordered = parallel.merge_results_in_game_order(all_res)
        assert ordered["game_index"].tolist() == [0, 1, 2, 3]
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert all_recs is None and all_res is None
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_single_process_paths_are_noops():
    recs = np.zeros(2, RECORD_DTYPE)
    assert parallel.gather_structured(recs) is recs
    assert parallel.reduce_counters({"a": 1}) == {"a": 1}
    b = torch.zeros(4, dtype=torch.uint8)
    assert parallel.broadcast_blob(b) is b
