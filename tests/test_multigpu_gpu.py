"""Needs >= 2 GPUs (skipped otherwise): the sharded engine under NCCL -- disjoint game indices per rank, weight blob
broadcast, records/results gathered to rank 0, and every gathered game bit-equal to the oracle."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from self_play_reinforcement_learning_b200 import nets, parallel
    from self_play_reinforcement_learning_b200.engine import RECORD_DTYPE  # noqa: F401
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    G, sims = 8, 30
    sp = BatchedSelfPlay(None, game=0, n_games=G, sims=sims, net="hash", seed=13, rank=rank, world=world, games_target=2 * world * G,
                         noise_mode=0)
    recs, res = [], []
    while True:
        sp.engine.run_ticks(256)
        recs.append(sp.engine.drain_records()); res.append(sp.engine.drain_results())
        if sp.engine.all_idle():
            break
    recs, res = np.concatenate(recs), np.concatenate(res)
    mine = {int(g) for g in res["game_index"]}
    assert all(parallel.owner_of_game(g, world, G) == rank for g in mine) and len(mine) == 2 * G
    dev = torch.device("cuda", rank)
    all_recs = parallel.gather_structured(recs, dst=0, device=dev)
    all_res = parallel.gather_structured(res, dst=0, device=dev)
    # weight blob broadcast over NCCL
    torch.manual_seed(0)
    blob_ref = nets.pack_tower_blob(nets.ResidualTower(7, 6, 7, num_blocks=1).eval())
    blob = (blob_ref.clone() if rank == 0 else torch.zeros_like(blob_ref)).cuda()
    parallel.broadcast_blob(blob, src=0)
    assert torch.equal(blob.cpu(), blob_ref)
    if rank == 0:
        np.save(os.path.join(out_dir, "recs.npy"), all_recs)
        np.save(os.path.join(out_dir, "res.npy"), all_res)
    sp.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_nccl_shard_gather_broadcast(tmp_path):
    import torch.multiprocessing as mp
    from tests import helpers as H
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    recs, res = np.load(tmp_path / "recs.npy"), np.load(tmp_path / "res.npy")
    by_rec, by_res = H.split_by_game(recs, res)
    assert sorted(by_res) == list(range(2 * world * 8))
    for g in sorted(by_res):
        o = H.oracle_episode(0, 30, 13, g, None, net_seed=13)
        assert by_res[g]["reward"] == o["reward"] and len(by_rec[g]) == len(o["records"])
        for a, b in zip(by_rec[g], o["records"]):
            assert np.array_equal(H.record_board(a, 0), b["state"]) and np.array_equal(a["tree_probs"][:7], b["tree_probs"])
