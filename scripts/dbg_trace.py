"""Per-layer clock64 trace of CTA 0 of the tower kernel (needs a library built with -DSPX_DBG_TRACE; SPX_LIB_PATH=...)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from self_play_reinforcement_learning_b200 import nets, _lib
from self_play_reinforcement_learning_b200.envs import boards_to_bits
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
rng = np.random.default_rng(0)
boards = torch.from_numpy(rng.integers(-1, 2, size=(1024, 7, 6)).astype(np.int64))
bits = boards_to_bits(boards.cuda(), 0)
tw = nets.NativeTower(net)
for _ in range(5):
    tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
torch.cuda.synchronize()
buf = np.zeros(64 * 16 + 3 * 160, np.int64)
_lib.lib().spx_debug_trace(C.c_void_p(buf.ctypes.data))
t = buf[:1024].reshape(64, 16)
blk = buf[1024:].reshape(160, 3)[:148]
names = ["mma:wait_epi", "mma:go", "mma:issued", "epi:wait_acc", "epi:acc_full", "epi:t0", "epi:t1", "epi:t2", "epi:tiles_done", "epi:bias_st", "epi:fenced", "epi:signalled"]
t0 = t[2, 1]
for l in range(2, 8):
    print("layer", l, {n: int(t[l, i] - t0) for i, n in enumerate(names)})
    t0 = t[l + 1, 1] if False else t0
d = np.diff(t[2:40, 1])
print("layer period (cycles): mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
print("issue time (go -> issued): %.0f" % (t[2:40, 2] - t[2:40, 1]).mean())
print("issued -> acc_full seen by epi: %.0f" % (t[2:40, 4] - t[2:40, 2]).mean())
print("epi acc_full -> t0/t1/t2/tiles_done/bias/fenced/signalled:", [(t[2:40, k] - t[2:40, 4]).mean().round() for k in range(5, 12)])
print("epi signalled -> next mma go: %.0f" % (t[3:41, 1] - t[2:40, 11]).mean())
L = 42
print("head layer:", {n: int(t[L - 1, i] - t[L - 1, 1]) for i, n in enumerate(names)})
print("FC phase (offsets from the head layer's mma:go):", dict(mma_wait=int(t[L, 0] - t[L - 1, 1]), mma_go=int(t[L, 1] - t[L - 1, 1]), mma_issued=int(t[L, 2] - t[L - 1, 1]),
      policy_start=int(t[L, 3] - t[L - 1, 1]), policy_done=int(t[L, 4] - t[L - 1, 1]), fc_done_seen=int(t[L, 5] - t[L - 1, 1]), value_partial_done=int(t[L, 6] - t[L - 1, 1])))
print("whole kernel (layer 0 go -> value partial done): %d cycles" % int(t[L, 6] - t[0, 1]))
print("kernel entry -> setup done: %d cycles; setup done -> layer 0 go: %d; value partial done -> exit path: %d; entry -> exit path: %d" % (
    t[62, 1] - t[62, 0], t[0, 1] - t[62, 1], t[62, 2] - t[L, 6], t[62, 2] - t[62, 0]))
t0 = blk[:, 0].min()
dur = blk[:, 1] - blk[:, 0]
print("per-CTA (globaltimer ns): first entry -> last exit %d ns; CTA duration min %d median %d max %d; entry skew max %d ns" % (
    blk[:, 1].max() - t0, dur.min(), np.median(dur), dur.max(), (blk[:, 0] - t0).max()))
slow = np.argsort(dur)[-6:]
print("slowest CTAs (block, sm, ns):", [(int(b), int(blk[b, 2]), int(dur[b])) for b in slow], "fastest:", [(int(b), int(blk[b, 2]), int(dur[b])) for b in np.argsort(dur)[:4]])
