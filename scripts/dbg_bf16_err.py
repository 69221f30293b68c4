import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import nets
from self_play_reinforcement_learning_b200.envs import boards_to_bits
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
for randomize in (False, True):
    torch.manual_seed(20)
    net = nets.ResidualTower(7, 6, 7, num_blocks=20).eval()
    if randomize:
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.uniform_(-0.1, 0.1)
    rng = np.random.default_rng(0)
    boards = torch.from_numpy(rng.integers(-1, 2, size=(1024, 7, 6)).astype(np.int64)).cuda()
    bits = boards_to_bits(boards, 0)
    tw = nets.NativeTower(net)
    p, v = tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
    with torch.no_grad():
        pr, vr = net.cuda().float().forward(boards)
        import copy
        nb = copy.deepcopy(net).cuda().to(torch.bfloat16)
        pb, vb = nb.forward_planes(nets.board_planes(boards, 7, 6, torch.bfloat16))
    print("randomized_bn" if randomize else "default_init", "native:", (p-pr).abs().max().item(), (v-vr.reshape(-1)).abs().max().item(),
          "| torch bf16:", (pb.float()-pr).abs().max().item(), (vb.float().reshape(-1)-vr.reshape(-1)).abs().max().item())
