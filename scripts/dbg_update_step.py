"""Time of one SGD step of the config-5 update (batch 128, ResidualTower-20, loss of mcts.py:234-252) in a few PyTorch settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_play_reinforcement_learning_b200 import nets
from self_play_reinforcement_learning_b200.replay import loss_from_batch

def run(tag, channels_last=False, benchmark=False, tf32=True, graph=False, B=128, steps=30, amp=None):
    torch.backends.cudnn.benchmark = benchmark
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=20).cuda().train()
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator(device="cuda").manual_seed(0)
    cells = torch.randint(-1, 2, (B, 7, 6), device="cuda", generator=g)
    planes = torch.stack([(cells == 0), (cells == 1), (cells == -1)], 1).float()
    if channels_last:
        planes = planes.contiguous(memory_format=torch.channels_last)
    batch = dict(planes=planes, tree_probs=torch.softmax(torch.randn(B, 7, device="cuda", generator=g), 1),
                 actual_val=torch.randint(-1, 2, (B,), device="cuda", generator=g).float(), q=torch.rand(B, device="cuda", generator=g))
    def step():
        if amp is not None:
            with torch.autocast("cuda", dtype=amp):
                loss = loss_from_batch(net, batch)
        else:
            loss = loss_from_batch(net, batch)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        return loss
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if graph:
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(gr):
            step()
        fn = gr.replay
    else:
        fn = step
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    print(f"{tag:40s} {1e3 * (time.time() - t0) / steps:7.2f} ms/step")

run("default (TF32 convs, NCHW)")
run("bf16 autocast NCHW", amp=torch.bfloat16)
run("bf16 autocast NCHW + benchmark", amp=torch.bfloat16, benchmark=True)
run("bf16 autocast channels_last + benchmark", amp=torch.bfloat16, channels_last=True, benchmark=True)
run("fp16 autocast NCHW + benchmark", amp=torch.float16, benchmark=True)
run("batch 512 default", B=512, steps=10)
run("batch 512 bf16 + benchmark", B=512, steps=10, amp=torch.bfloat16, benchmark=True)
