"""Decode what the MN-major tcgen05 backward-weights kernel computes: one-hot / ramp inputs, several descriptor stride choices."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200._lib import check, lib  # noqa: E402

dev = "cuda"
rows = 16 + 128
N = 128


def planes(t):   # [rows, C] -> bf16 [C/8][rows][8]
    r, c = t.shape
    return t.view(r, c // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)


def run(x, dy, strides, taps=1):
    out = torch.zeros(1, taps, 128, dy.shape[1], device=dev)
    px, pdy = planes(x), planes(dy)     # keep the temporaries alive until the kernel has run
    check(lib().spx_train_debug_wgrad(px.data_ptr(), pdy.data_ptr(), dy.shape[1], taps, rows, 1, out.data_ptr(), *strides,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "wgrad")
    torch.cuda.synchronize()
    return out[0]


g = torch.Generator().manual_seed(0)
x = torch.zeros(rows, 128); dy = torch.zeros(rows, N)
x[8:136] = torch.randint(-4, 5, (128, 128), generator=g).float()
dy[8:136] = torch.randint(-4, 5, (128, N), generator=g).float()
ref = x[8:136].t() @ dy[8:136]          # [ci][co]
x, dy = x.to(dev), dy.to(dev)
PIECE, DT = 144 * 16, 128 * 16
for name, st in [("lbo128/sbo=plane", (128, PIECE, 128, DT)), ("lbo=plane/sbo128", (PIECE, 128, DT, 128)),
                 ("A swapped only", (PIECE, 128, 128, DT)), ("B swapped only", (128, PIECE, DT, 128))]:
    out = run(x, dy, st)[0].cpu()
    print(name, "max err", (out - ref).abs().max().item(), "ref max", ref.abs().max().item(), "out norm", out.norm().item(), "ref norm", ref.norm().item())
# one-hot decode with the default strides: x one-hot at (row r0, ci c0); dy[row][co] = row  -> out[ci'][co'] = r' paired
for (r0, c0) in [(0, 0), (1, 0), (9, 0), (0, 1), (0, 4), (0, 5), (17, 37)]:
    x1 = torch.zeros(rows, 128); x1[8 + r0, c0] = 1.0
    d1 = torch.zeros(rows, N); d1[8:136] = torch.arange(128).float()[:, None].expand(128, N) + 1
    d2 = torch.zeros(rows, N); d2[8:136] = torch.arange(N).float()[None, :].expand(128, N) + 1
    for name, st in [("default", (128, PIECE, 128, DT)), ("swapped", (PIECE, 128, DT, 128))]:
        o1 = run(x1.to(dev), d1.to(dev), st)[0].cpu(); o2 = run(x1.to(dev), d2.to(dev), st)[0].cpu()
        nz = o1.nonzero()
        print(name, "one-hot row", r0, "ci", c0, "-> nonzero rows(ci') of out:", sorted(set(nz[:, 0].tolist()))[:8], "n nonzero", len(nz),
              "paired row+1 values", sorted(set(o1[o1 != 0].tolist()))[:6], "co map sample", o2[nz[0, 0]][:6].tolist() if len(nz) else None)
