import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets, _lib
from self_play_reinforcement_learning_b200.envs import boards_to_bits
torch.manual_seed(0)
net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
n = 1024
rng = np.random.default_rng(0)
boards = torch.from_numpy(rng.integers(-1, 2, size=(n, 7, 6)).astype(np.int64))
bits = boards_to_bits(boards.cuda(), 0)
own, opp = bits[:, 0].contiguous(), bits[:, 1].contiguous()
tw = nets.NativeTower(net)
for _ in range(5):
    tw.forward_bits(own, opp)
torch.cuda.synchronize()
tt, hh = [], []
for _ in range(30):
    ev = tuple(_lib.Event() for _ in range(3))
    tw.forward_bits(own, opp, events=ev)
    torch.cuda.synchronize()
    tt.append(ev[0].elapsed_time(ev[1])); hh.append(ev[1].elapsed_time(ev[2]))
print("tower ms", np.median(tt), "heads ms", np.median(hh), "min heads", min(hh))
