"""max |policy - fp32|, |value - fp32| of the native tower for both element types (fp16 default, bf16) against the true-fp32
torch forward (TF32 off), on the bench's random-init net and on nets with randomised BatchNorm statistics; and speed of each.
    python scripts/dbg_tower_err.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets  # noqa: E402
from self_play_reinforcement_learning_b200.envs import boards_to_bits  # noqa: E402
from tests.test_tower_gpu import _random_positions, _randomise_bn  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rows = []
for blocks, seed, rand_bn in ((20, 0, False), (20, 20, True), (20, 21, True), (15, 3, True), (2, 2, True)):
    torch.manual_seed(seed)
    net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    if rand_bn:
        _randomise_bn(net)
    boards = torch.from_numpy(_random_positions(4096, 5 + seed))
    bits = boards_to_bits(boards.cuda(), 0)
    with torch.no_grad():
        pr, vr = net.cuda().float().forward(boards.cuda())
        net16 = __import__("copy").deepcopy(net).half()
        ph, vh = net16.forward_planes(nets.board_planes(boards.cuda(), 7, 6, torch.float16))
        netb = __import__("copy").deepcopy(net).bfloat16()
        pb, vb = netb.forward_planes(nets.board_planes(boards.cuda(), 7, 6, torch.bfloat16))
    net.cpu()
    for dt in ("f16", "bf16"):
        os.environ["SPX_TOWER_DTYPE"] = dt
        tw = nets.NativeTower(net)
        p, v = tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        own, opp = bits[:1024, 0].contiguous(), bits[:1024, 1].contiguous()
        for _ in range(20):
            tw.forward_bits(own, opp)
        a.record()
        for _ in range(200):
            tw.forward_bits(own, opp)
        b.record()
        torch.cuda.synchronize()
        rows.append((blocks, rand_bn, dt, (p - pr).abs().max().item(), (v - vr.reshape(-1)).abs().max().item(), a.elapsed_time(b) / 200,
                     bool(torch.isfinite(p).all() and torch.isfinite(v).all())))
        tw.close()
    rows.append((blocks, rand_bn, "torch.half", (ph.float() - pr).abs().max().item(), (vh.float() - vr).abs().max().item(), 0, True))
    rows.append((blocks, rand_bn, "torch.bf16", (pb.float() - pr).abs().max().item(), (vb.float() - vr).abs().max().item(), 0, True))
for r in rows:
    print("blocks=%d randBN=%d %-10s max|dp|=%.3e max|dv|=%.3e ms/1024=%.4f finite=%s" % r)
