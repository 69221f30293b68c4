"""env_step bandwidth vs problem size, next to a torch copy of the same number of bytes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_play_reinforcement_learning_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for logn in (22, 24, 26):
    n = 1 << logn
    state = torch.zeros(n, 2, dtype=torch.int64, device=dev); done = torch.zeros(n, dtype=torch.uint8, device=dev)
    reward = torch.zeros(n, dtype=torch.int8, device=dev); valid = torch.zeros(n, dtype=torch.int16, device=dev)
    status = torch.zeros(n, dtype=torch.int8, device=dev); player = torch.ones(n, dtype=torch.int8, device=dev)
    a = torch.randint(0, 7, (n,), device=dev, dtype=torch.int32)
    ts = []
    for i in range(7):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.spx_env_step(0, n, state.data_ptr(), done.data_ptr(), a.data_ptr(), player.data_ptr(), reward.data_ptr(), valid.data_ptr(), status.data_ptr(), st))
        e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    src = torch.empty(n * 43 // 2 // 8, dtype=torch.int64, device=dev); dst = torch.empty_like(src)
    tc = []
    for i in range(7):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src); e1.record(); torch.cuda.synchronize()
        if i >= 2: tc.append(e0.elapsed_time(e1))
    print(f"n=2^{logn}: env_step {ms:.4f} ms = {n * 43 / ms / 1e6:.0f} GB/s ; torch copy of the same bytes {min(tc):.4f} ms = {n * 43 / min(tc) / 1e6:.0f} GB/s")
    del state, done, reward, valid, status, player, a, src, dst
