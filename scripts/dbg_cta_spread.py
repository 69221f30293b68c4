"""Per-CTA duration of ONE fused-tick launch (entry / exit globaltimer of every CTA; needs the -DSPX_DBG_TRACE build):
which clusters finish last, and by how much -- the launch lasts as long as its slowest cluster.
    SPX_LIB_PATH=variants/libspx_trace.so python scripts/dbg_cta_spread.py [games] [ticks]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import _lib, nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=G, sims=800, net="tower", seed=0)
e = sp.engine
e.stagger()
e.run_ticks(3000, chunk=100)
torch.cuda.synchronize()
for rep in range(3):
    e.run_ticks(T, chunk=T)
    torch.cuda.synchronize()
    buf = np.zeros(64 * 16 + 3 * 160, np.int64)
    _lib.lib().spx_debug_trace(C.c_void_p(buf.ctypes.data))
    blk = buf[1024:].reshape(160, 3)[:148]
    dur = (blk[:, 1] - blk[:, 0]) / 1e3
    t0 = blk[:, 0].min()
    order = np.argsort(dur)
    print(f"games {G} ticks {T}: launch {(blk[:, 1].max() - t0) / 1e3:.0f} us; CTA duration min {dur.min():.0f} p25 {np.percentile(dur, 25):.0f} median {np.median(dur):.0f} "
          f"p75 {np.percentile(dur, 75):.0f} max {dur.max():.0f} us; slowest (block, sm): {[(int(b), int(blk[b, 2])) for b in order[-6:]]} fastest: {[(int(b), int(blk[b, 2])) for b in order[:4]]}", flush=True)
