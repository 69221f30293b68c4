"""Five spx_env_step launches over 16 Mi Connect4 boards with a fresh action array each (operands from DRAM): the ncu target
for the env kernel (ncu -k regex:env_step -s 2 -c 3 --set full)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_play_reinforcement_learning_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n = 1 << 24
state = torch.zeros(n, 2, dtype=torch.int64, device=dev); done = torch.zeros(n, dtype=torch.uint8, device=dev)
reward = torch.zeros(n, dtype=torch.int8, device=dev); valid = torch.zeros(n, dtype=torch.int16, device=dev)
status = torch.zeros(n, dtype=torch.int8, device=dev); player = torch.ones(n, dtype=torch.int8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
acts = [torch.randint(0, 7, (n,), generator=g, device=dev, dtype=torch.int32) for _ in range(5)]
ts = []
for a in acts:
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    _lib.check(L.spx_env_step(0, n, state.data_ptr(), done.data_ptr(), a.data_ptr(), player.data_ptr(), reward.data_ptr(), valid.data_ptr(), status.data_ptr(), st))
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); player = -player
print("env_step ms per launch:", [round(t, 4) for t in ts], "GB/s:", [round(n * 43 / t / 1e6) for t in ts])
