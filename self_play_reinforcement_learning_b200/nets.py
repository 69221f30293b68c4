"""Policy/value networks evaluated at the MCTS leaves, and their evaluators.

Model definitions keep the reference's plugin contract (games/general/modules.py:43-125,
games/tictactoe/modules.py:14-81): same constructor arguments, same ``state_dict`` keys (so reference
checkpoints ``{"model": state_dict}`` load unchanged), same ``forward(int boards[B,W,H]) ->
(policy[B,A] softmaxed, value[B,1] tanh)`` and the same ``net(state, player) -> (list, float)`` call
convention with the frame flip of modules.py:109-112.  Parameter creation order matches the
reference so that ``torch.manual_seed(s)`` yields the same random-init weights
(tests/test_nets_cpu.py pins this against golden vectors from the reference classes).

Evaluators turn the engine's dense leaf batch (bitboards in the net frame) into policy/value rows:
  * TowerEvaluator    -- the hand-written sm_100a tcgen05 tower in libspx (the product path)
  * TorchNetEvaluator -- any nn.Module through PyTorch (generic-module boundary and the in-repo
                         cuDNN/cuBLAS comparison point)
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from ._lib import GAME_CONNECT4, check, lib


def init_weights(m):
    """rl_utils/weights.py:5-8: xavier-uniform conv weights, bias 0.01 (Linear layers keep torch defaults)."""
    if isinstance(m, nn.Conv2d):
        nn.init.xavier_uniform_(m.weight)
        m.bias.data.fill_(0.01)


def board_planes(s, width, height, dtype=torch.float32):
    """int boards [...,W,H] -> one-hot planes [B,3,W,H] ordered (empty, own == +1, enemy == -1)
    (general/modules.py:115-125)."""
    s = torch.as_tensor(s)
    s = s.reshape(-1, width, height)
    return torch.stack([(s == 0), (s == 1), (s == -1)], dim=1).to(dtype)


def bits_to_planes(own, opp, game, dtype=torch.float32):
    """bitboards int64[B] (net frame) -> planes [B,3,W,H] without materialising int boards."""
    W, H, _ = _lib.GAME_DIMS[game]
    stride = 7 if game == GAME_CONNECT4 else 3
    idx = (torch.arange(W, device=own.device)[:, None] * stride + torch.arange(H, device=own.device)[None, :])
    o = ((own[:, None, None] >> idx) & 1)
    e = ((opp[:, None, None] >> idx) & 1)
    return torch.stack([1 - o - e, o, e], dim=1).to(dtype)


class BasicBlock(nn.Module):
    """conv-BN-ReLU-conv-BN (+identity) ReLU, 3x3 'same' convolutions (general/modules.py:13-40)."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=True)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=True)
        self.bn2 = nn.BatchNorm2d(planes)

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + x)


class _BoardNet(nn.Module):
    """Shared call convention: ``net(state, player)`` flips the frame in and the value out."""
    width = 0
    height = 0

    def __call__(self, state, player=1):
        policy, value = nn.Module.__call__(self, torch.as_tensor(np.asarray(state) * player))
        return policy.tolist()[0], value.item() * player

    def forward(self, x):
        dev = next(self.parameters()).device
        return self.forward_planes(board_planes(x, self.width, self.height).to(dev))


class ResidualTower(_BoardNet):
    """general/modules.py:43-112.  channels = 4*filter_factor, heads reduce to filter_factor channels."""

    def __init__(self, width=7, height=6, action_size=7, num_blocks=15, default_kernel_size=3, filter_factor=32):
        super().__init__()
        ch = filter_factor * 4
        self.width, self.height, self.action_size, self.num_blocks, self.filter_factor = width, height, action_size, num_blocks, filter_factor
        self.conv1 = nn.Conv2d(3, ch, default_kernel_size, 1, 1, bias=True)
        self.bn1 = nn.BatchNorm2d(ch)
        self.relu = nn.ReLU(inplace=True)
        self.residual_blocks = nn.Sequential(*[BasicBlock(ch, ch) for _ in range(num_blocks)])
        flat = width * height * filter_factor
        self.conv_policy = nn.Conv2d(ch, filter_factor, 1)
        self.policy_bn = nn.BatchNorm2d(filter_factor)
        self.policy_dropout = nn.Dropout(p=0.5)
        self.linear_policy = nn.Linear(flat, action_size)
        self.conv_value = nn.Conv2d(ch, filter_factor, 1)
        self.value_bn = nn.BatchNorm2d(filter_factor)
        self.value_dropout = nn.Dropout(p=0.5)
        self.fc_value = nn.Linear(flat, filter_factor * 8)
        self.linear_output = nn.Linear(filter_factor * 8, 1)
        self.apply(init_weights)

    @staticmethod
    def from_env(env, num_blocks=15, filter_factor=32):
        return ResidualTower(env.width, env.height, env.num_actions(), num_blocks, filter_factor=filter_factor)

    def forward_planes(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x = self.residual_blocks(x)
        p = F.relu(self.policy_bn(self.conv_policy(x))).flatten(1)
        p = F.softmax(self.linear_policy(self.policy_dropout(p)), dim=1)
        v = F.relu(self.value_bn(self.conv_value(x))).flatten(1)
        v = F.relu(self.fc_value(self.value_dropout(v)))
        return p, torch.tanh(self.linear_output(v))


class ConvNetTicTacToe(_BoardNet):
    """games/tictactoe/modules.py:14-81 (3 conv layers 3->128->128->64, leaky-ReLU, no dropout in forward).
    The reference's default action_size=3 is wrong for a 3x3 board; pass 9."""

    def __init__(self, width=3, height=3, action_size=3):
        super().__init__()
        self.width, self.height, self.action_size = width, height, action_size
        self.conv1 = nn.Conv2d(3, 128, 3, 1, 1, bias=True)
        self.bn1 = nn.BatchNorm2d(128)
        self.conv2 = nn.Conv2d(128, 128, 3, 1, 1, bias=True)
        self.bn2 = nn.BatchNorm2d(128)
        self.conv3 = nn.Conv2d(128, 64, 3, 1, 1, bias=True)
        self.bn3 = nn.BatchNorm2d(64)
        cells = width * height
        self.conv_policy = nn.Conv2d(64, 2, 1)
        self.policy_bn = nn.BatchNorm2d(2)
        self.policy_dropout = nn.Dropout(p=0.5)
        self.linear_policy = nn.Linear(cells * 2, action_size)
        self.conv_value = nn.Conv2d(64, 1, 1)
        self.value_bn = nn.BatchNorm2d(1)
        self.value_dropout = nn.Dropout(p=0.5)
        self.fc_value = nn.Linear(cells, 256)
        self.linear_output = nn.Linear(256, 1)
        self.apply(init_weights)

    def forward_planes(self, x):
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = F.leaky_relu(bn(conv(x)))
        p = F.leaky_relu(self.policy_bn(self.conv_policy(x))).flatten(1)
        p = F.softmax(self.linear_policy(p), dim=1)
        v = F.leaky_relu(self.value_bn(self.conv_value(x))).flatten(1)
        v = F.leaky_relu(self.fc_value(v))
        return p, torch.tanh(self.linear_output(v))


# --------------------------------------------------------------------------------------------- evaluators
class TorchNetEvaluator:
    """Evaluates the leaf batch with an arbitrary nn.Module through PyTorch (library kernels).

    ``module.forward_planes(planes[B,3,W,H])`` is used when present, otherwise ``module.forward`` on int
    boards (the generic reference contract).  ``module_opp`` evaluates tree 1 in two-net mode."""

    def __init__(self, module, game, module_opp=None, dtype=torch.bfloat16, channels_last=True):
        self.game, self.dtype, self.channels_last = game, dtype, channels_last
        self.modules = [module] + ([module_opp] if module_opp is not None else [])

    def bind(self, engine):
        """Evaluation runs on PRIVATE copies (device, eval mode, evaluation dtype, channels_last): nn.Module.to works in place,
        so casting the caller's module would turn the fp32 master weights it trains and checkpoints into bf16."""
        import copy
        dev = engine.device
        self.masters = list(self.modules)
        mods = []
        for m in self.masters:
            c = copy.deepcopy(m).to(dev).eval()
            for p in c.parameters():
                p.requires_grad_(False)
            if self.dtype != torch.float32:
                c = c.to(self.dtype)
            if self.channels_last:
                c = c.to(memory_format=torch.channels_last)
            mods.append(c)
        self.modules = mods

    def refresh(self, module=None, which=0):
        """Copies the weights of `module` (default: the master this evaluator was built from) into the evaluation copy."""
        src = module if module is not None else self.masters[which]
        with torch.no_grad():
            self.modules[which].load_state_dict(src.state_dict())
        self.modules[which].eval()

    @torch.no_grad()
    def forward_bits(self, own, opp, which=0):
        m = self.modules[which]
        W, H, _ = _lib.GAME_DIMS[self.game]
        planes = bits_to_planes(own, opp, self.game, self.dtype)
        if self.channels_last:
            planes = planes.contiguous(memory_format=torch.channels_last)
        if hasattr(m, "forward_planes"):
            p, v = m.forward_planes(planes)
        else:
            boards = (planes[:, 1] - planes[:, 2]).to(torch.int64)
            p, v = m.forward(boards)
        return p.float(), v.float().reshape(-1)

    @torch.no_grad()
    def __call__(self, engine):
        p, v = self.forward_bits(engine.leaf_own, engine.leaf_opp, 0)
        if len(self.modules) > 1:
            p1, v1 = self.forward_bits(engine.leaf_own, engine.leaf_opp, 1)
            sel = engine.net_id.bool()
            p = torch.where(sel[:, None], p1, p)
            v = torch.where(sel, v1, v)
        engine.policy.copy_(p)
        engine.value.copy_(v)


# --------------------------------------------------------------------------------------------- native tower
def _align(x, a=256):
    return (x + a - 1) // a * a


def _fold(conv, bn):
    """eval-mode BatchNorm folded into the preceding convolution (fp32)."""
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv.weight.detach().float() * s[:, None, None, None]
    b = (conv.bias.detach().float() - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w.cpu(), b.cpu()


def _stage_blocks(w, ncta=1, dtype=torch.bfloat16):
    """[N, Cin(multiple of 16), kh, kw] -> bf16 elements in MMA consumption order: tap-major, then 16-channel K
    slices, each slice stored as [2 k-chunks][N][8] (the no-swizzle K-major core-matrix layout).
    ncta=2 (SM-pair kernel): ring stages carry two K slices (one for the 3-channel stem) and every stage is stored as
    [half of the output channels][K slice][2 k-chunks][N/2][8], so each CTA of the pair copies one contiguous half."""
    N, Cin, KH, KW = w.shape
    ks = Cin // 16
    if ncta == 1:
        x = w.reshape(N, ks, 2, 8, KH, KW).permute(4, 5, 1, 2, 0, 3).contiguous()
        return x.to(dtype).reshape(-1)
    per = 2 if ks >= 2 else 1                       # K slices per ring stage
    x = w.reshape(2, N // 2, ks // per, per, 2, 8, KH, KW)   # [h, n, kpair, j, kc, e, kh, kw]
    x = x.permute(6, 7, 2, 0, 3, 4, 1, 5).contiguous()       # [kh, kw, kpair, h, j, kc, n, e]
    return x.to(dtype).reshape(-1)


def _bias_slice(b, dtype=torch.bfloat16):
    """Folded-BN bias [N] -> the first ring stage of its layer in the SM-pair kernel: the B operand of the bias MMA
    (ones * bias^T), [half][2 k-chunks][N/2][8] with k = 0 -> round(b), k = 1 -> round(b - round(b)) in the stream's
    element type (16 significant bits for bf16, 22 for fp16), zeros elsewhere."""
    N = b.numel()
    hi = b.float().to(dtype)
    lo = (b.float() - hi.float()).to(dtype)
    x = torch.zeros(2, 2, N // 2, 8, dtype=dtype)
    x[:, 0, :, 0] = hi.reshape(2, N // 2)
    x[:, 0, :, 1] = lo.reshape(2, N // 2)
    return x.reshape(-1)


def _fc_stream(w1, dtype=torch.bfloat16):
    """fc_value.weight [256 hidden, 1344] -> the weight-ring stream of the fused value layer (SM-pair kernel): the hidden
    units are the M rows of the MMA (128 per CTA), so every K step of 16 inputs is stored per CTA half as
    [2 k-chunks][128 hidden][8] bf16 = 4 KB; a ring stage carries two K steps: [stage][half][2 K steps][4 KB]."""
    HID, K = w1.shape
    x = w1.detach().float().cpu().reshape(2, HID // 2, K // 32, 2, 2, 8)    # [half, n, stage, j, kc, e]
    return x.permute(2, 0, 3, 4, 1, 5).contiguous().to(dtype).reshape(-1)   # [stage, half, j, kc, n, e]


def pack_tower_blob(module, ncta=2, f16=None):
    """ResidualTower (128 trunk channels; the 7x6 Connect4 or the 3x3 TicTacToe board of ResidualTower.from_env) -> one flat
    uint8 tensor in the layout spx_tower_load expects (spx_tower.cu: tower_layout).  f16: element type of the weight stream the
    kernel consumes (conv trunk + fused value layer) -- fp16 (True; what spx_tower_create selects unless SPX_TOWER_DTYPE=bf16)
    or bf16; None = what a tower created now would take."""
    m = module
    if f16 is None:
        f16 = ncta == 2 and not os.environ.get("SPX_TOWER_DTYPE", "f16").lower().startswith("b")
    sdt = torch.float16 if f16 else torch.bfloat16
    _bs, _sb, _fs = _bias_slice, _stage_blocks, _fc_stream
    _bias_slice_d = lambda b: _bs(b, sdt)                  # noqa: E731
    _stage_blocks_d = lambda w, n: _sb(w, n, sdt)          # noqa: E731

    def require(ok, what):
        if not ok:
            raise ValueError(f"the native tower cannot run this network ({what}); use net='torch' (nets.TorchNetEvaluator) for it")
    require(hasattr(m, "residual_blocks"), "not a ResidualTower")
    require((m.width, m.height) in ((7, 6), (3, 3)), f"board {m.width}x{m.height}: built for the 7x6 Connect4 and 3x3 TicTacToe boards")
    require(m.conv1.out_channels == 128 and m.conv_policy.out_channels == 32, "filter_factor must be 32 (128 trunk channels)")
    require(m.conv1.kernel_size == (3, 3), "default_kernel_size must be 3")
    blocks = list(m.residual_blocks)
    n_layers = 2 * len(blocks) + 2
    A, FLAT, HID = m.linear_policy.out_features, 32 * m.width * m.height, 256
    require(A == (7 if m.width == 7 else 9), "action_size must be the board's own (7 / 9)")
    require(m.fc_value.out_features == HID and m.linear_policy.in_features == FLAT, "head sizes differ from general/modules.py:66-78")
    conv_parts, biases = [], torch.zeros(n_layers, 128)
    w, b = _fold(m.conv1, m.bn1)
    wp = torch.zeros(128, 16, 3, 3)
    wp[:, :3] = w
    conv_parts.append(_bias_slice_d(b))
    conv_parts.append(_stage_blocks_d(wp, ncta))
    biases[0] = b
    li = 1
    for blk in blocks:
        for conv, bn in ((blk.conv1, blk.bn1), (blk.conv2, blk.bn2)):
            w, b = _fold(conv, bn)
            conv_parts.append(_bias_slice_d(b))
            conv_parts.append(_stage_blocks_d(w, ncta))
            biases[li] = b
            li += 1
    wpol, bpol = _fold(m.conv_policy, m.policy_bn)
    wval, bval = _fold(m.conv_value, m.value_bn)
    conv_parts.append(_bias_slice_d(torch.cat([bpol, bval])))
    conv_parts.append(_stage_blocks_d(torch.cat([wpol, wval], 0), ncta))
    biases[li, :64] = torch.cat([bpol, bval])
    conv_parts.append(_fs(m.fc_value.weight, sdt))
    f32 = lambda t: t.detach().float().cpu().contiguous()  # noqa: E731
    pieces = [torch.cat(conv_parts).view(torch.int16).view(torch.uint8),
              biases.reshape(-1).contiguous().view(torch.uint8),
              f32(m.linear_policy.weight).reshape(-1).view(torch.uint8),
              torch.cat([f32(m.linear_policy.bias), torch.zeros(16 - A)]).view(torch.uint8),
              f32(m.fc_value.weight).to(sdt).reshape(-1).view(torch.int16).view(torch.uint8),
              f32(m.fc_value.bias).view(torch.uint8),
              f32(m.linear_output.weight).reshape(-1).view(torch.uint8),
              torch.cat([f32(m.linear_output.bias), torch.zeros(3)]).view(torch.uint8)]
    total = sum(_align(p.numel()) for p in pieces)
    blob = torch.zeros(total, dtype=torch.uint8)
    off = 0
    for p in pieces:
        blob[off:off + p.numel()] = p
        off += _align(p.numel())
    return blob


class NativeTower:
    """Handle on one spx_tower (the tcgen05 network of libspx) with weights loaded from a ResidualTower."""

    def __init__(self, module, game=None):
        if not torch.cuda.is_available():
            raise _lib.SpxError("the native tower runs on sm_100a only (no CPU fallback)")
        board_game = GAME_CONNECT4 if (module.width, module.height) == (7, 6) else _lib.GAME_TICTACTOE
        if game is not None and game != board_game:
            raise ValueError("the network's board size does not match the game")
        self.game, self.num_blocks = board_game, len(module.residual_blocks)
        self.A = module.linear_policy.out_features
        self._h = C.c_void_p()
        check(lib().spx_tower_create(self.game, self.num_blocks, C.byref(self._h)), "spx_tower_create")
        self.ncta = lib().spx_tower_ncta(self._h)
        self.fused_heads = bool(lib().spx_tower_fused_heads(self._h))
        self.f16 = bool(lib().spx_tower_f16(self._h))
        self.load(module)

    def load(self, module_or_blob):
        """module (packed on the host, then H2D) or an already packed uint8 blob (pinned host or device tensor)."""
        blob = module_or_blob if torch.is_tensor(module_or_blob) else pack_tower_blob(module_or_blob, self.ncta, self.f16)
        want = lib().spx_tower_blob_bytes(self.game, self.num_blocks)
        if blob.numel() != want:
            raise ValueError(f"packed weight blob has {blob.numel()} bytes, this tower ({self.num_blocks} blocks) takes {want}")
        self.blob_dev = blob.to("cuda", non_blocking=True)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().spx_tower_load(self._h, self.blob_dev.data_ptr(), self.blob_dev.numel(), stream), "spx_tower_load")

    def forward_bits(self, own, opp, needs_eval=None, policy=None, value=None, events=None):
        """events: optional (start, tower_done, end) torch.cuda.Event triple recorded around the two kernels."""
        n = own.numel()
        policy = torch.empty(n, self.A, dtype=torch.float32, device=own.device) if policy is None else policy
        value = torch.empty(n, dtype=torch.float32, device=own.device) if value is None else value
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        need = None if needs_eval is None else needs_eval.data_ptr()
        if events is None:
            check(lib().spx_tower_forward(self._h, own.data_ptr(), opp.data_ptr(), need, n, policy.data_ptr(), value.data_ptr(),
                                          stream), "spx_tower_forward")
        else:
            ev = [C.c_void_p(e.cuda_event) for e in events]
            check(lib().spx_tower_forward_timed(self._h, own.data_ptr(), opp.data_ptr(), need, n, policy.data_ptr(),
                                                value.data_ptr(), stream, ev[0], ev[1], ev[2]), "spx_tower_forward_timed")
        return policy, value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            torch.cuda.synchronize()
            lib().spx_tower_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TowerEvaluator:
    """Evaluates the engine's leaf batch with the hand-written tcgen05 tower (the product path)."""

    def __init__(self, module, game=GAME_CONNECT4):
        self.tower = NativeTower(module, game)

    def bind(self, engine):
        pass

    def load(self, module_or_blob):
        self.tower.load(module_or_blob)

    def __call__(self, engine, events=None):
        self.tower.forward_bits(engine.leaf_own, engine.leaf_opp, engine.needs_eval, engine.policy, engine.value, events)

    def cache_versions(self):
        """Weights versions behind the outputs this evaluator writes (network 0, network 1): the engine's evaluation-cache tags."""
        return (int(lib().spx_tower_version(self.tower._h)), 0)

    def fused_ticks(self, engine, n, balanced=False):
        """n whole ticks (advance + evaluation) in one persistent launch (spx_tick_fused; balanced: the work-conserving form
        spx_tick_fused_balanced, n ticks per game ON AVERAGE); False if this tower cannot."""
        if not (self.tower.ncta == 2 and self.tower.fused_heads) or os.environ.get("SPX_FUSED_TICK", "1") == "0":
            return False
        fn = lib().spx_tick_fused_balanced if balanced else lib().spx_tick_fused
        check(fn(engine._h, self.tower._h, int(n), engine.policy.data_ptr(), engine.value.data_ptr(),
                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "spx_tick_fused")
        return True


class TwoTowerEvaluator:
    """Head-to-head evaluation with two native towers (tree 0 -> network 0, tree 1 -> network 1): the leaf batch is
    partitioned by network on the device, each tower evaluates only its own rows, the outputs are scattered back."""

    def __init__(self, module, module_opp, game=GAME_CONNECT4):
        self.towers = [NativeTower(module, game), NativeTower(module_opp, game)]
        self.tower = self.towers[0]

    def bind(self, engine):
        n, dev, A = engine.n_leaves, engine.device, engine.A
        self.own2 = torch.zeros(2, n, dtype=torch.int64, device=dev)
        self.opp2 = torch.zeros(2, n, dtype=torch.int64, device=dev)
        self.needs2 = torch.zeros(2, n, dtype=torch.uint8, device=dev)
        self.map2 = torch.zeros(2, n, dtype=torch.int32, device=dev)
        self.pol2 = torch.zeros(2, n, A, dtype=torch.float32, device=dev)
        self.val2 = torch.zeros(2, n, dtype=torch.float32, device=dev)

    def load(self, module_or_blob, which=0):
        self.towers[which].load(module_or_blob)

    def cache_versions(self):
        return tuple(int(lib().spx_tower_version(t._h)) for t in self.towers)

    def __call__(self, engine, events=None):
        n = engine.n_leaves
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().spx_partition_leaves(n, engine.leaf_own.data_ptr(), engine.leaf_opp.data_ptr(), engine.needs_eval.data_ptr(),
                                         engine.net_id.data_ptr(), self.own2.data_ptr(), self.opp2.data_ptr(), self.needs2.data_ptr(),
                                         self.map2.data_ptr(), st), "spx_partition_leaves")
        for k in (0, 1):
            self.towers[k].forward_bits(self.own2[k], self.opp2[k], self.needs2[k], self.pol2[k], self.val2[k])
        check(lib().spx_scatter_outputs(n, engine.A, self.needs2.data_ptr(), self.map2.data_ptr(), self.pol2.data_ptr(),
                                        self.val2.data_ptr(), engine.policy.data_ptr(), engine.value.data_ptr(), st), "spx_scatter_outputs")

    def close(self):
        for t in self.towers:
            t.close()


# --------------------------------------------------------------------------------------------- native TicTacToe net
def pack_tttnet_blob(module):
    """ConvNetTicTacToe (3x3, action_size 9) -> fp32 blob in the layout of csrc/spx_tttnet.cu: conv weights as
    [ic][tap][oc] with eval-mode BN folded, then the head convs, linear_policy, fc_value, linear_output."""
    m = module
    if not (m.width == 3 and m.height == 3 and m.linear_policy.out_features == 9 and m.conv3.out_channels == 64):
        raise ValueError("the native TicTacToe kernel runs ConvNetTicTacToe(3, 3, 9) (tictactoe/modules.py:14-53) only; use net='torch' otherwise")
    parts = []
    for conv, bn in ((m.conv1, m.bn1), (m.conv2, m.bn2), (m.conv3, m.bn3)):
        w, b = _fold(conv, bn)                                   # [oc, ic, 3, 3]
        parts += [w.permute(1, 2, 3, 0).reshape(-1), b]          # [ic][tap = kh*3+kw][oc]
    wp, bp = _fold(m.conv_policy, m.policy_bn)                   # [2, 64, 1, 1]
    wv, bv = _fold(m.conv_value, m.value_bn)                     # [1, 64, 1, 1]
    f32 = lambda t: t.detach().float().cpu().contiguous()        # noqa: E731
    parts += [wp.reshape(2, 64).t().reshape(-1), bp, wv.reshape(-1), bv,
              f32(m.linear_policy.weight).reshape(-1), f32(m.linear_policy.bias),
              f32(m.fc_value.weight).reshape(-1), f32(m.fc_value.bias),
              f32(m.linear_output.weight).reshape(-1), f32(m.linear_output.bias)]
    blob = torch.cat([p.contiguous().reshape(-1) for p in parts]).contiguous()
    if blob.numel() != lib().spx_tttnet_blob_floats():
        raise ValueError(f"packed TicTacToe net has {blob.numel()} floats, the kernel takes {lib().spx_tttnet_blob_floats()}")
    return blob


class TTTNetEvaluator:
    """Evaluates TicTacToe leaves with the hand-written fp32 kernel of ConvNetTicTacToe (the repo's tictactoe net)."""

    def __init__(self, module):
        if not torch.cuda.is_available():
            raise _lib.SpxError("the native TicTacToe net runs on the GPU only (no CPU fallback)")
        self._h = C.c_void_p()
        check(lib().spx_tttnet_create(C.byref(self._h)), "spx_tttnet_create")
        self.load(module)

    def bind(self, engine):
        pass

    def load(self, module_or_blob):
        blob = module_or_blob if torch.is_tensor(module_or_blob) else pack_tttnet_blob(module_or_blob)
        self.blob_dev = blob.to("cuda", torch.float32)
        check(lib().spx_tttnet_load(self._h, self.blob_dev.data_ptr(), self.blob_dev.numel(),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "spx_tttnet_load")

    def forward_bits(self, own, opp, needs_eval=None, policy=None, value=None):
        n = own.numel()
        policy = torch.empty(n, 9, dtype=torch.float32, device=own.device) if policy is None else policy
        value = torch.empty(n, dtype=torch.float32, device=own.device) if value is None else value
        check(lib().spx_tttnet_forward(self._h, own.data_ptr(), opp.data_ptr(), None if needs_eval is None else needs_eval.data_ptr(), n,
                                       policy.data_ptr(), value.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "spx_tttnet_forward")
        return policy, value

    def __call__(self, engine, events=None):
        self.forward_bits(engine.leaf_own, engine.leaf_opp, engine.needs_eval, engine.policy, engine.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            torch.cuda.synchronize()
            lib().spx_tttnet_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def smoke_check():
    """Tiny tower forward on cuda:0 against the fp32 torch forward of the same module."""
    torch.manual_seed(0)
    net = ResidualTower(7, 6, 7, num_blocks=2).eval()
    boards = torch.randint(-1, 2, (16, 7, 6))
    from .envs import boards_to_bits
    bits = boards_to_bits(boards.cuda(), GAME_CONNECT4)
    tw = NativeTower(net)
    p, v = tw.forward_bits(bits[:, 0].contiguous(), bits[:, 1].contiguous())
    with torch.no_grad():
        pr, vr = net.forward(boards)
    torch.cuda.synchronize()
    tol = 1e-3 if tw.f16 else 8e-3      # two blocks: measured 4e-5 / 1e-4 (fp16), 3e-4 / 7e-4 (bf16)
    assert (p.cpu() - pr).abs().max() < tol and (v.cpu() - vr.reshape(-1)).abs().max() < tol, "native tower diverges from fp32 torch"
    tw.close()
