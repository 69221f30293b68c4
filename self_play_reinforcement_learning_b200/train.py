"""DeviceTrainer: the reference's SGD step (MCTreeSearch.loss + update_from_memory, games/algos/mcts.py:234-270, driven by
UpdateWorker.update, games/algos/updateworker.py:141-149) as hand-written sm_100a kernels (csrc/spx_train.cu) behind the C ABI.

The trainer owns flat fp32 copies of the network's parameters, BatchNorm running statistics and SGD momentum buffers on the
device; ``load`` takes them from a ``nets.ResidualTower`` (named_parameters order), ``step`` runs one training step on a batch
from ``replay.DeviceReplay.sample_batch``, ``store`` writes them back into the module (whose state_dict then feeds
``nets.pack_tower_blob`` / ``checkpoint.save_model`` as before).  There is no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _bn_modules(module):
    return [m for m in module.modules() if isinstance(m, torch.nn.BatchNorm2d)]


def supports(module):
    """The native step is built for the reference's default tower on Connect4: 128 trunk channels (filter_factor 32), 7x6 boards."""
    return (getattr(module, "filter_factor", None) == 32 and getattr(module, "width", None) == 7 and getattr(module, "height", None) == 6
            and getattr(module, "action_size", None) == 7 and hasattr(module, "residual_blocks"))


class DeviceTrainer:
    def __init__(self, module, batch_size=128, lr=0.01, momentum=0.9, weight_decay=1e-4, seed=0):
        if not torch.cuda.is_available():
            raise _lib.SpxError("DeviceTrainer needs a CUDA device (B200); there is no CPU fallback")
        if not supports(module):
            raise ValueError("DeviceTrainer is built for ResidualTower(7, 6, 7, filter_factor=32); train other networks with PyTorch autograd "
                             "(SelfPlayScheduler(trainer='torch'))")
        self.num_blocks, self.batch_size = int(module.num_blocks), int(batch_size)
        self.lr, self.momentum, self.weight_decay, self.seed = float(lr), float(momentum), float(weight_decay), int(seed)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._h = C.c_void_p()
        check(lib().spx_train_create(self.num_blocks, self.batch_size, C.byref(self._h)), "spx_train_create")
        self.n_params = int(lib().spx_train_param_count(self._h))
        self.n_running = int(lib().spx_train_running_count(self._h))
        self.steps = 0
        self._loss = torch.zeros(3, device=self.device)
        self.load(module)

    # ------------------------------------------------------------------ state in / out
    def load(self, module, reset_momentum=True):
        """Take parameters and BatchNorm running statistics from ``module`` (any device / dtype; converted to fp32)."""
        flat = torch.cat([p.detach().reshape(-1).to(self.device, torch.float32) for p in module.parameters()])
        run = torch.cat([torch.cat([m.running_mean.detach().reshape(-1), m.running_var.detach().reshape(-1)]).to(self.device, torch.float32)
                         for m in _bn_modules(module)])
        if flat.numel() != self.n_params or run.numel() != self.n_running:
            raise ValueError(f"network has {flat.numel()} parameters / {run.numel()} running statistics, the trainer ({self.num_blocks} blocks) "
                             f"takes {self.n_params} / {self.n_running}")
        check(lib().spx_train_set_state(self._h, flat.data_ptr(), run.data_ptr(), int(bool(reset_momentum)), _stream()), "spx_train_set_state")
        torch.cuda.current_stream().synchronize()     # flat / run are temporaries

    def _get(self, what, n):
        out = torch.empty(n, device=self.device)
        check(lib().spx_train_get_state(self._h, what, out.data_ptr(), _stream()), "spx_train_get_state")
        return out

    def parameters_flat(self):
        return self._get(0, self.n_params)

    def running_flat(self):
        return self._get(1, self.n_running)

    def gradients_flat(self):
        return self._get(2, self.n_params)

    def momentum_flat(self):
        return self._get(3, self.n_params)

    @staticmethod
    def unflatten(flat, module):
        """flat parameter-shaped vector -> {name: tensor} in named_parameters order."""
        out, off = {}, 0
        for name, p in module.named_parameters():
            out[name] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        return out

    def store(self, module):
        """Write parameters, running statistics and the step count back into ``module`` (in place, keeping its dtype/device)."""
        flat, run = self.parameters_flat(), self.running_flat()
        off = 0
        with torch.no_grad():
            for p in module.parameters():
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            off = 0
            for m in _bn_modules(module):
                c = m.running_mean.numel()
                m.running_mean.copy_(run[off:off + c]); m.running_var.copy_(run[off + c:off + 2 * c])
                m.num_batches_tracked += self.steps - getattr(self, "_stored_steps", 0)
                off += 2 * c
        self._stored_steps = self.steps

    # ------------------------------------------------------------------ the step
    def step(self, planes, tree_probs, target, dropout_mask=None, apply_update=True, lr=None):
        """One update_from_memory step.  planes f32 [B, 3, 7, 6]; tree_probs f32 [B, 7]; target f32 [B] (actual_val + q when
        q_average); dropout_mask: optional uint8 keep-masks [B, 2, 1344] (tests).  Returns a device tensor (total, value, policy loss)."""
        B = self.batch_size
        planes = planes.to(self.device, torch.float32).contiguous()
        tree_probs = tree_probs.to(self.device, torch.float32).contiguous()
        target = target.to(self.device, torch.float32).contiguous()
        if planes.shape != (B, 3, 7, 6) or tree_probs.shape != (B, 7) or target.shape != (B,):
            raise ValueError(f"batch shapes {tuple(planes.shape)}, {tuple(tree_probs.shape)}, {tuple(target.shape)} do not fit batch size {B}")
        if dropout_mask is not None:
            dropout_mask = dropout_mask.to(self.device, torch.uint8).contiguous()
            if dropout_mask.shape != (B, 2, 1344):
                raise ValueError("dropout_mask must be uint8 [B, 2, 1344]")
        check(lib().spx_train_step(self._h, planes.data_ptr(), tree_probs.data_ptr(), target.data_ptr(),
                                   None if dropout_mask is None else dropout_mask.data_ptr(), self.seed, self.steps,
                                   self.lr if lr is None else float(lr), self.momentum, self.weight_decay, int(bool(apply_update)),
                                   self._loss.data_ptr(), _stream()), "spx_train_step")
        if apply_update:
            self.steps += 1
        return self._loss

    def step_from_batch(self, batch, q_average=True, lr=None):
        """``batch``: a DeviceReplay.sample_batch dict (planes, tree_probs, actual_val, q)."""
        target = batch["actual_val"] + batch["q"] if q_average else batch["actual_val"]
        return self.step(batch["planes"], batch["tree_probs"], target, lr=lr)

    def outputs(self):
        """Train-mode network outputs of the last step's forward pass: (probs [B, 7], value [B])."""
        p = torch.empty(self.batch_size, 7, device=self.device)
        v = torch.empty(self.batch_size, device=self.device)
        check(lib().spx_train_outputs(self._h, p.data_ptr(), v.data_ptr(), _stream()), "spx_train_outputs")
        return p, v

    def debug_planes(self, which, layer=0):
        """Internal plane tensor as [B, C, 7, 6] (tests)."""
        ptr, n, rows = C.c_void_p(), C.c_int64(), C.c_int32()
        check(lib().spx_train_debug_planes(self._h, which, layer, C.byref(ptr), C.byref(n), C.byref(rows)), "spx_train_debug_planes")
        from .engine import _view
        flat = _view(ptr.value, (n.value,), "<f4", self.device)
        chunks = n.value // (rows.value * 4)
        t = flat.view(chunks, rows.value, 4)[:, 8:8 + 56 * self.batch_size].reshape(chunks, self.batch_size, 56, 4)[:, :, :49]
        t = t.reshape(chunks, self.batch_size, 7, 7, 4)[:, :, :, :6]          # [chunk, b, col, row, 4]
        return t.permute(1, 0, 4, 2, 3).reshape(self.batch_size, chunks * 4, 7, 6).clone()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            torch.cuda.synchronize()
            lib().spx_train_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
