"""Timeline of one fused tick (CTA 0, first epilogue thread, globaltimer ns) -- needs a library built with -DSPX_DBG_TRACE:
    nvcc ... -DSPX_DBG_TRACE -o /tmp/libspx_trace.so ...;  SPX_LIB_PATH=/path/libspx_trace.so python scripts/dbg_tick_trace.py [games]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import _lib, nets  # noqa: E402
from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
sp = BatchedSelfPlay(net, game=0, n_games=G, sims=800, net="tower", seed=0)
e = sp.engine
e.stagger()
e.run_ticks(2000, chunk=100)
torch.cuda.synchronize()
e.run_ticks(100, chunk=100)
torch.cuda.synchronize()
buf = np.zeros(64 * 12, np.int64)
_lib.lib().spx_debug_tick_trace(C.c_void_p(buf.ctypes.data))
t = buf.reshape(64, 12)[:, :11]
names = ["top", "leaf_ready", "bar1", "preproc_signalled", "layer0_acc", "layers_done", "policy_dots", "fc_done", "partials", "value", "end_bar"]
d = np.diff(t, axis=1)
print("games", G, "mean ns between points:", {f"{names[i]}->{names[i+1]}": float(d[:, i].mean().round(0)) for i in range(10)})
print("pass period (top->top):", np.diff(t[:, 0]).mean(), "min", np.diff(t[:, 0]).min(), "max", np.diff(t[:, 0]).max())
print("end_bar -> next top:", (t[1:, 0] - t[:-1, 10]).mean())
