"""Multi-GPU plumbing: one process per GPU, games sharded by slot, no collective on the data path.

The reference is single-GPU and moves weights through checkpoint files and records/results through
multiprocessing queues (updateworker.py:111-117 -> inference_worker.py:68-73; memory_queue/result_queue,
selfplayworker.py:185-190).  Here rank 0 broadcasts the packed weight blob (NCCL over NVLink; gloo in the CPU
tests) and the self-play records/results are gathered to rank 0.  Everything takes plain tensors / numpy arrays so
the same code runs under gloo on CPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard(rank, world, games_per_rank):
    """Global game-index layout: rank r owns slots [r*G, (r+1)*G); slot s plays games s, s+world*G, s+2*world*G, ...
    Returns the (slot_offset, slot_stride) pair of spx_config."""
    return rank * games_per_rank, world * games_per_rank


def owner_of_game(game_index, world, games_per_rank):
    return (game_index % (world * games_per_rank)) // games_per_rank


def broadcast_blob(blob, src=0, group=None):
    """In-place broadcast of the packed uint8 weight blob (every rank passes a tensor of the same size)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(blob, src=src, group=group)
    return blob


def _gather_bytes(arr, dst, group, device):
    """Gather variable-length byte arrays to dst: all_gather the lengths, pad, all_gather the payloads."""
    world = dist.get_world_size(group)
    raw = torch.from_numpy(np.frombuffer(arr.tobytes(), dtype=np.uint8).copy()).to(device)
    n = torch.tensor([raw.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros(cap, dtype=torch.uint8, device=device)
    padded[:raw.numel()] = raw
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    if dist.get_rank(group) != dst:
        return None
    return [o[:s].cpu().numpy().tobytes() for o, s in zip(out, sizes)]


def gather_structured(arr, dst=0, group=None, device="cpu"):
    """Gather a numpy structured array (engine.RECORD_DTYPE / RESULT_DTYPE) from every rank; rank ``dst`` gets the
    concatenation ordered by rank, the others get None."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arr
    parts = _gather_bytes(arr, dst, group, device)
    if parts is None:
        return None
    return np.concatenate([np.frombuffer(p, dtype=arr.dtype) for p in parts]) if parts else arr[:0]


def gather_device_rows(rows, dst=0, group=None):
    """Gather a [n_r, k] tensor with a different n_r on every rank (device-resident Move records, n_r x 80 bytes) to rank
    ``dst`` without a host copy: all_gather of the row counts, then one padded all_gather.  Returns the list of per-rank
    tensors on ``dst`` and None elsewhere; the memory_queue of the reference (mcts.py:225-232) for ranks > 0."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [rows]
    world = dist.get_world_size(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    padded = torch.zeros((max(max(sizes), 1),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    padded[:rows.shape[0]] = rows
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    if dist.get_rank(group) != dst:
        return None
    return [o[:k] for o, k in zip(out, sizes)]


def reduce_counters(counters, group=None, device="cpu"):
    """Sum the per-rank counter dicts (every rank gets the total)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(counters)
    keys = sorted(counters)
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def merge_results_in_game_order(results):
    """Results gathered from all ranks, ordered by global game index (== the order tasks were issued in,
    self_play_parallel.py:250-253)."""
    return results[np.argsort(results["game_index"], kind="stable")]
