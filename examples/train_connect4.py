#!/usr/bin/env python
"""BASELINE.json configs[4] shape: the full self_play_parallel loop -- batched self-play on every GPU, records gathered
to rank 0, training steps (mcts.py:234-270 loss, SGD momentum 0.9 wd 1e-4), weights broadcast over NCCL, evaluation.

    python examples/train_connect4.py --epochs 2                                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_connect4.py --epochs 2
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.scheduler import SelfPlayScheduler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--iterations", type=int, default=800)
    ap.add_argument("--epoch-length", type=int, default=1500)      # run_self_play_connect4.py:58
    ap.add_argument("--evaluation-games", type=int, default=150)   # :59
    ap.add_argument("--initial-games", type=int, default=256)
    ap.add_argument("--games-per-gpu", type=int, default=1024)
    ap.add_argument("--save-dir", default=None, help="run folder root: model-*/memory-* files as the reference writes them")
    ap.add_argument("--resume", action="store_true", help="resume model and memory from the newest earlier run in --save-dir")
    ap.add_argument("--no-save-memory", action="store_true", help="with --save-dir: write only the model checkpoints, not the pickled replay memory")
    ap.add_argument("--amp", action="store_true", help="with --trainer torch: bf16 autocast for the SGD steps")
    ap.add_argument("--trainer", default="auto", choices=["auto", "device", "torch"], help="device: the native SGD step (csrc/spx_train.cu)")
    ap.add_argument("--threads", type=int, default=1, help="thread_count of the searches (the reference's default behind its InferenceProxy is 4)")
    ap.add_argument("--eval-cache", action="store_true", help="per-slot evaluation cache inside the fused tick kernel: the same games in about half the network passes (DESIGN.md 3.9)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=args.blocks).cuda().eval()
    sched = SelfPlayScheduler(net, 0, iterations=args.iterations, epoch_length=args.epoch_length, initial_games=args.initial_games,
                              evaluation_games=args.evaluation_games, games_per_gpu=args.games_per_gpu, save_dir=args.save_dir,
                              save_memory=not args.no_save_memory, amp=torch.bfloat16 if args.amp else None, trainer=args.trainer,
                              search_threads=args.threads, eval_cache=args.eval_cache)
    t0 = time.time()
    hist = sched.train_model(num_epochs=args.epochs, resume_model=args.resume, resume_memory=args.resume)
    if sched.rank == 0:
        print(json.dumps({"world": world, "seconds": time.time() - t0, "history": hist}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
