"""Per-layer clock64 trace of CTA 0 inside the FUSED tick kernel (last pass of each launch; -DSPX_DBG_TRACE build), averaged
over launches: does the search running next to the network (shadow warp) slow the MMA issue stream, the epilogue, or neither?
    SPX_LIB_PATH=variants/libspx_trace.so [SPX_DBG_FLAGS=1] python scripts/dbg_layer_trace_fused.py [games]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import _lib, nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=G, sims=800, net="tower", seed=0)
e = sp.engine
e.stagger()
e.run_ticks(2000, chunk=100)
torch.cuda.synchronize()
acc = []
for rep in range(24):
    e.run_ticks(37 + rep, chunk=64)
    torch.cuda.synchronize()
    buf = np.zeros(64 * 16 + 3 * 160, np.int64)
    _lib.lib().spx_debug_trace(C.c_void_p(buf.ctypes.data))
    t = buf[:1024].reshape(64, 16)
    lay = slice(2, 40)
    acc.append(dict(period=np.diff(t[2:41, 1]).mean(), issue=(t[lay, 2] - t[lay, 1]).mean(), issued_to_acc=(t[lay, 4] - t[lay, 2]).mean(),
                    epi_tiles=(t[lay, 8] - t[lay, 4]).mean(), epi_total=(t[lay, 11] - t[lay, 4]).mean(),
                    signalled_to_go=(t[3:41, 1] - t[2:40, 11]).mean(), wait_epi=(t[lay, 1] - t[lay, 0]).mean()))
keys = acc[0].keys()
print("games", G, "flags", os.environ.get("SPX_DBG_FLAGS", "0"), {k: round(float(np.median([a[k] for a in acc])), 0) for k in keys})
