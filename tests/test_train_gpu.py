"""The SGD step on the device (csrc/spx_train.cu) against PyTorch autograd: MCTreeSearch.loss
(games/algos/mcts.py:234-252) on ResidualTower.train() with injected Dropout masks, torch.optim.SGD(momentum 0.9, weight decay 1e-4).

Tolerances.  The convolutions run in TF32 (forward / backward-data) and bf16 (backward-weights) with fp32 accumulation, the
reference below in true fp32 (TF32 off) or fp64.  Forward values agree to ~1e-3 relative (TF32's 10-bit mantissa); gradients
additionally see ReLU masks flip where a forward value sits within that error of zero (a fraction f of flipped units costs
~sqrt(f) in relative L2), measured 3-4e-2 per tensor on 2 blocks -- the same size as the error of PyTorch's own TF32 step, which
the test measures next to ours and uses as the yardstick."""
import ctypes as C

import pytest
import torch

from self_play_reinforcement_learning_b200 import nets
from self_play_reinforcement_learning_b200._lib import check, lib
from tests import train_ref as R

pytestmark = pytest.mark.gpu


def _net(blocks, seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    net = R.patch_dropout(nets.ResidualTower(7, 6, 7, num_blocks=blocks))
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
    return net.cuda().to(dtype).train()


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _planes16(t):   # [rows, C] -> bf16 [C/8][rows][8]
    r, c = t.shape
    return t.view(r, c // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)


@pytest.mark.parametrize("N,taps,chunks,S", [(128, 9, 2, 2), (64, 1, 3, 2), (128, 9, 3, 1)])
def test_backward_weights_kernel_is_exact_on_integers(N, taps, chunks, S):
    """dW[tap][ci][co] = sum_rows x[row + shift(tap)][ci] * dy[row][co]: small integers are exact in bf16 and in the fp32 accumulator."""
    g = torch.Generator().manual_seed(N + taps)
    rows = 16 + 128 * chunks
    x = torch.zeros(rows, 128); dy = torch.zeros(rows, N)
    x[8:-8] = torch.randint(-3, 4, (rows - 16, 128), generator=g).float()
    dy[8:-8] = torch.randint(-3, 4, (rows - 16, N), generator=g).float()
    px, pdy = _planes16(x.cuda()), _planes16(dy.cuda())
    out = torch.zeros(S, taps, 128, N, device="cuda")
    check(lib().spx_train_debug_wgrad(px.data_ptr(), pdy.data_ptr(), N, taps, rows, S, out.data_ptr(), -1, -1, -1, -1,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "spx_train_debug_wgrad")
    torch.cuda.synchronize()
    got = out.sum(0).cpu()
    for tap in range(taps):
        shift = (tap // 3 - 1) * 7 + (tap % 3 - 1) if taps == 9 else 0
        ref = x[8 + shift:rows - 8 + shift].t() @ dy[8:-8]
        assert torch.equal(got[tap], ref), tap


def _reference_step(blocks, planes, probs, target, mask, dtype, tf32=False):
    net = _net(blocks, dtype=dtype)
    loss, lv, lp, p, v = R.torch_loss(net, planes.to(dtype), probs.to(dtype), target.to(dtype), mask)
    if tf32:   # PyTorch's own TF32 step (stock behaviour of fp32 convolutions on this GPU): the yardstick
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        net.policy_dropout.keep, net.value_dropout.keep = mask[:, 0], mask[:, 1]
        p, v = net.forward_planes(planes)
        lv = torch.nn.functional.mse_loss(v.view(-1), target); lp = -(p.log() * probs).sum() / p.size(0)
        loss = lv + lp
    loss.backward()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return net, loss, lv, lp, p, v


@pytest.mark.parametrize("blocks,B", [(2, 32), (3, 128), (1, 20)])
def test_forward_loss_and_gradients_match_autograd(blocks, B):
    from self_play_reinforcement_learning_b200.train import DeviceTrainer
    planes, probs, target, mask = R.make_batch(B, seed=blocks)
    net64, loss64, lv64, lp64, p64, v64 = _reference_step(blocks, planes, probs, target, mask, torch.float64)
    net_tf, *_ = _reference_step(blocks, planes, probs, target, mask, torch.float32, tf32=True)
    net = _net(blocks)
    tr = DeviceTrainer(net, batch_size=B)
    out = tr.step(planes, probs, target, dropout_mask=mask, apply_update=False).cpu()
    p, v = tr.outputs()
    # forward: loss terms, network outputs (TF32 convolutions: stated tolerance 2e-3 absolute on probabilities / value)
    assert abs(out[0].item() - loss64.item()) < 2e-3 * abs(loss64.item()) and abs(out[1].item() - lv64.item()) < 2e-3 and abs(out[2].item() - lp64.item()) < 2e-3
    assert (p.double() - p64).abs().max().item() < 2e-3 and (v.double() - v64.view(-1)).abs().max().item() < 2e-3
    # gradients per tensor against fp64 autograd, next to PyTorch's own TF32 step
    g = DeviceTrainer.unflatten(tr.gradients_flat(), net)
    ours, yard, ref_all, our_all = {}, {}, [], []
    for (name, p64_), (_, ptf) in zip(net64.named_parameters(), net_tf.named_parameters()):
        if p64_.grad.norm() < 1e-6:          # conv biases in front of a training-mode BatchNorm: exactly 0 here, ~1e-9 noise in autograd
            assert g[name].abs().max().item() < 1e-6, name
            continue
        ours[name], yard[name] = _rel(g[name], p64_.grad), _rel(ptf.grad, p64_.grad)
        ref_all.append(p64_.grad.reshape(-1)); our_all.append(g[name].reshape(-1).double())
    worst = max(ours, key=ours.get)
    total = _rel(torch.cat(our_all), torch.cat(ref_all))
    print(f"blocks {blocks} B {B}: worst tensor {worst} ours {ours[worst]:.3e} (torch TF32 {yard[worst]:.3e}); all gradients {total:.3e}; "
          f"torch TF32 worst {max(yard.values()):.3e}")
    assert total < 6e-2 and all(e < 1e-1 for e in ours.values()), ours
    assert ours[worst] < max(3.0 * max(yard.values()), 2e-2)
    # BatchNorm running statistics after one training-mode forward (momentum 0.1, unbiased variance)
    run = tr.running_flat()
    want = torch.cat([torch.cat([m.running_mean, m.running_var]) for m in net64.modules() if isinstance(m, torch.nn.BatchNorm2d)])
    assert _rel(run, want) < 2e-3
    tr.close()


def test_ten_sgd_steps_follow_torch_sgd():
    """10 update_from_memory steps (lr 0.01, momentum 0.9, weight decay 1e-4) against torch.optim.SGD on true-fp32 autograd.
    SGD trajectories of a ReLU network separate under ANY rounding difference, so the yardstick is measured in the same test:
    PyTorch's own TF32 run of the same 10 steps against its fp32 run.  Stated tolerance: the weight change (final - initial)
    agrees to 15 % relative L2 or twice the yardstick, the loss curves to 1e-2."""
    from self_play_reinforcement_learning_b200.train import DeviceTrainer
    blocks, B = 2, 64
    net = _net(blocks)
    ref = _net(blocks)
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    tf = _net(blocks)
    opt_tf = torch.optim.SGD(tf.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    tr = DeviceTrainer(net, batch_size=B, lr=0.01, momentum=0.9, weight_decay=1e-4)
    init = torch.cat([p.detach().reshape(-1) for p in ref.parameters()]).clone()
    l_ref, l_dev = [], []
    for step in range(10):
        planes, probs, target, mask = R.make_batch(B, seed=100 + step)
        loss, *_ = R.torch_loss(ref, planes, probs, target, mask)
        opt.zero_grad(); loss.backward(); opt.step()
        l_ref.append(loss.item())
        torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
        tf.policy_dropout.keep, tf.value_dropout.keep = mask[:, 0], mask[:, 1]
        p_, v_ = tf.forward_planes(planes)
        loss_tf = torch.nn.functional.mse_loss(v_.view(-1), target) - (p_.log() * probs).sum() / B
        opt_tf.zero_grad(); loss_tf.backward(); opt_tf.step()
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        l_dev.append(tr.step(planes, probs, target, dropout_mask=mask)[0].item())
    got = tr.parameters_flat() - init
    want = torch.cat([p.detach().reshape(-1) for p in ref.parameters()]) - init
    err = _rel(got, want)
    yard = _rel(torch.cat([p.detach().reshape(-1) for p in tf.parameters()]) - init, want)
    print("10-step trajectory: relative error of the weight change", err, "(torch TF32 vs torch fp32:", yard, ") losses", l_ref[-1], l_dev[-1])
    assert err < max(0.15, 2.0 * yard)
    assert max(abs(a - b) for a, b in zip(l_ref, l_dev)) < 1e-2 * max(l_ref)
    # store(): the module carries the trained state (parameters, running statistics, step count)
    tr.store(net)
    assert _rel(torch.cat([p.detach().reshape(-1) for p in net.parameters()]), tr.parameters_flat()) == 0.0
    assert int(net.bn1.num_batches_tracked) == 10
    assert _rel(net.bn1.running_var, ref.bn1.running_var) < 2e-3
    tr.close()


def test_generated_dropout_masks_keep_half_and_differ_per_step():
    from self_play_reinforcement_learning_b200.train import DeviceTrainer
    B = 32
    net = _net(1)
    tr = DeviceTrainer(net, batch_size=B, seed=3)
    planes, probs, target, _ = R.make_batch(B)
    l0 = tr.step(planes, probs, target)[0].item()
    l1 = tr.step(planes, probs, target)[0].item()
    assert l0 == l0 and l1 == l1 and l0 != l1
    tr.close()


def test_hundred_steps_of_batch_128_take_under_a_second():
    """100 SGD steps of batch 128 on ResidualTower-20 in < 1 s (PyTorch autograd: 2.9-4.2 s)."""
    from self_play_reinforcement_learning_b200.train import DeviceTrainer
    B = 128
    net = _net(20)
    tr = DeviceTrainer(net, batch_size=B)
    planes, probs, target, _ = R.make_batch(B)
    for _ in range(5):
        tr.step(planes, probs, target)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        tr.step(planes, probs, target)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("100 steps of batch 128, 20 blocks:", ms, "ms")
    assert ms < 1000.0
    tr.close()


def test_step_on_the_golden_loss_batch(golden_dir):
    """The batch of the reference's golden loss fixture (tests/golden/nets.npz: states, tree_probs, actual_val, q written by the
    unmodified reference; target = actual_val + q, mcts.py:243-244) through one device step in train mode vs autograd on the same
    module (tests/test_scheduler_cpu.py ties that autograd formula to the reference's own loss value)."""
    import os
    import numpy as np
    from self_play_reinforcement_learning_b200.train import DeviceTrainer
    g = np.load(os.path.join(golden_dir, "nets.npz"))
    boards = torch.from_numpy(g["loss_states"].astype(np.int64)).cuda()
    planes = torch.stack([(boards == 0), (boards == 1), (boards == -1)], 1).float()
    probs = torch.from_numpy(g["loss_probs"]).cuda()
    target = torch.from_numpy(g["loss_val"] + g["loss_q"]).cuda()
    mask = (torch.rand(16, 2, 1344, generator=torch.Generator().manual_seed(1)) < 0.5).to(torch.uint8).cuda()
    torch.manual_seed(3)
    net = R.patch_dropout(nets.ResidualTower(7, 6, 7, num_blocks=2)).cuda().train()
    ref64 = R.patch_dropout(nets.ResidualTower(7, 6, 7, num_blocks=2))
    ref64.load_state_dict({k: v for k, v in net.state_dict().items()}, strict=False)
    ref64 = ref64.cuda().double().train()
    tr = DeviceTrainer(net, batch_size=16)
    out = tr.step(planes, probs, target, dropout_mask=mask, apply_update=False).cpu()
    loss64, lv64, lp64, _, _ = R.torch_loss(ref64, planes.double(), probs.double(), target.double(), mask)
    loss64.backward()
    assert abs(out[0].item() - loss64.item()) < 2e-3 * abs(loss64.item()) and abs(out[1].item() - lv64.item()) < 2e-3 and abs(out[2].item() - lp64.item()) < 2e-3
    gflat = DeviceTrainer.unflatten(tr.gradients_flat(), net)
    errs = {n: _rel(gflat[n], p.grad) for n, p in ref64.named_parameters() if p.grad.norm() > 1e-6}
    assert max(errs.values()) < 1e-1, errs
    tr.close()
