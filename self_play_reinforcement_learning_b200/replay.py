"""DeviceReplay: the reference's ``Memory`` (rl_utils/memory.py:8-30) kept in HBM (csrc/spx_replay.cu).

The reference moves every ``Move`` through a multiprocessing queue into a host deque, then stacks a sampled batch and copies
it to the GPU for every SGD step (mcts.py:217-222,234-243).  Here the engine's record ring is drained device-to-device
(``spx_drain_records_device``, sorted by game/tree/ply so the content is reproducible), appended to a device ring
(``spx_replay_append``) and a training batch is produced by two small kernels (``spx_replay_sample``): indices without
replacement, then boards / preprocess planes / tree_probs / actual_val / q as dense device tensors.

Kept Memory API: ``len()``, ``add(Move)``, ``change_size(max_size)`` (UpdateWorker.stagger_memory, updateworker.py:107-109),
``reset()``, ``sample(batch_size)`` (list of Move tuples, device tensors), ``max_size``, ``deduplicate(...)`` (memory.py:47-94).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .engine import RECORD_DTYPE


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def moves_to_records(moves, game):
    """Move tuples (reference format, mcts.py:17) -> structured records (resume from a pickled Memory, Memory.add)."""
    W, H, A = _lib.GAME_DIMS[game]
    stride = 7 if game == _lib.GAME_CONNECT4 else 3
    out = np.zeros(len(moves), RECORD_DTYPE)
    shift = (np.arange(W)[:, None] * stride + np.arange(H)[None, :]).astype(np.uint64)
    for i, m in enumerate(moves):
        b = np.asarray(m.state.cpu() if hasattr(m.state, "cpu") else m.state).astype(np.int64)
        out["own"][i] = np.bitwise_or.reduce(((b == 1).astype(np.uint64) << shift).reshape(-1))
        out["opp"][i] = np.bitwise_or.reduce(((b == -1).astype(np.uint64) << shift).reshape(-1))
        out["tree_probs"][i, :A] = np.asarray(m.tree_probs.cpu() if hasattr(m.tree_probs, "cpu") else m.tree_probs, dtype=np.float32)
        out["q"][i] = float(m.q)
        out["actual_val"][i] = float(m.actual_val) if m.actual_val is not None else 0.0
        out["ply"][i] = int(np.abs(b).sum())
    return out


class DeviceReplay:
    def __init__(self, game, max_size=200000, physical_capacity=None, seed=0, device=None):
        self.game = game
        self.W, self.H, self.A = _lib.GAME_DIMS[game]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.seed, self.step = seed, 0
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().spx_replay_create(int(max_size), int(physical_capacity or max_size), C.byref(self._h)), "spx_replay_create")
        self._staging = None

    # ------------------------------------------------------------------ Memory API
    def __len__(self):
        return int(lib().spx_replay_size(self._h))

    @property
    def max_size(self):
        return int(lib().spx_replay_max_size(self._h))

    def change_size(self, max_size):
        check(lib().spx_replay_change_size(self._h, int(max_size)), "spx_replay_change_size")

    def reset(self):
        check(lib().spx_replay_reset(self._h), "spx_replay_reset")

    def add(self, experience):
        self.extend([experience])

    def deduplicate(self, key="state", values=("actual_val", "tree_probs"), named_tuple=None, maxlen=None):
        """Memory.deduplicate (memory.py:47-54) with the Deduplicator (memory.py:56-94) kept on the device
        (``spx_replay_deduplicate``): the buffer becomes one record per distinct state ever seen, tree_probs / actual_val / q
        averaged over its occurrences (f32 sums in insertion order, then sum / count), first-seen order, the last ``maxlen``.
        The signature is the reference's; only its one call site's arguments (mcts.py:385-386) are supported."""
        if key != "state" or not set(values) <= {"actual_val", "tree_probs", "q"}:
            raise ValueError("DeviceReplay.deduplicate: key must be 'state' and values a subset of actual_val / tree_probs / q")
        check(lib().spx_replay_deduplicate(self._h, int(maxlen or 0), _stream()), "spx_replay_deduplicate")
        return len(self)

    @property
    def unique_states(self):
        """len(deduplicator.counter): distinct states folded so far (0 before the first deduplicate)."""
        return int(lib().spx_replay_unique(self._h))

    def extend(self, moves):
        """Memory.add for host-side Move tuples (e.g. a pickled Memory being resumed)."""
        self.append_records(moves_to_records(list(moves), self.game))

    def append_records(self, records):
        """records: structured numpy array (engine.RECORD_DTYPE) on the host, or a uint8 device tensor [n, 80]."""
        if isinstance(records, np.ndarray):
            if len(records) == 0:
                return 0
            records = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(len(records), RECORD_DTYPE.itemsize)).to(self.device)
        n = int(records.shape[0])
        if n:
            assert records.dtype == torch.uint8 and records.is_contiguous() and records.shape[1] == RECORD_DTYPE.itemsize
            check(lib().spx_replay_append(self._h, records.data_ptr(), n, _stream()), "spx_replay_append")
        return n

    def drain_engine(self, engine, append=True):
        """pull_from_queue (mcts.py:217-222): the engine's finished-game records, device to device.  Returns the drained
        records as a uint8 device tensor view [n, 80] (valid until the next drain) after appending them (append=False: only
        drain, e.g. on a rank that ships its records to the trainer rank)."""
        cap = int(engine.cfg.record_capacity)
        if self._staging is None or self._staging.shape[0] < cap:
            self._staging = torch.empty(cap, RECORD_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        n = C.c_int64()
        check(lib().spx_drain_records_device(engine._h, self._staging.data_ptr(), cap, C.byref(n), _stream()), "spx_drain_records_device")
        recs = self._staging[:n.value]
        if append:
            self.append_records(recs)
        return recs

    def sample_batch(self, batch_size, step=None, boards=False, planes=True):
        """Memory.sample + the stacking of MCTreeSearch.loss: dict of device tensors (idx, [boards], [planes], tree_probs,
        actual_val, q).  `step` defaults to an internal counter (one draw stream per SGD step)."""
        if step is None:
            step, self.step = self.step, self.step + 1
        B, dev = int(batch_size), self.device
        out = dict(idx=torch.empty(B, dtype=torch.int64, device=dev), tree_probs=torch.empty(B, self.A, device=dev),
                   actual_val=torch.empty(B, device=dev), q=torch.empty(B, device=dev))
        if boards:
            out["boards"] = torch.empty(B, self.W, self.H, dtype=torch.int64, device=dev)
        if planes:
            out["planes"] = torch.empty(B, 3, self.W, self.H, device=dev)
        ptr = lambda k: out[k].data_ptr() if k in out else None
        check(lib().spx_replay_sample(self._h, self.game, B, self.seed, int(step), ptr("idx"), ptr("boards"), ptr("planes"), ptr("tree_probs"),
                                      ptr("actual_val"), ptr("q"), _stream()), "spx_replay_sample")
        return out

    def sample(self, batch_size):
        """Reference-format batch: a list of Move(state, actual_val, tree_probs, q) (device tensors, reference dtypes)."""
        from .selfplay import Move
        b = self.sample_batch(batch_size, boards=True, planes=False)
        return [Move(b["boards"][i], b["actual_val"][i], b["tree_probs"][i], b["q"][i]) for i in range(int(batch_size))]

    def read(self, first=0, n=None):
        """Logical records [first, first+n) as a structured host array (save_memory, tests)."""
        n = len(self) - first if n is None else n
        out = np.zeros(n, RECORD_DTYPE)
        if n:
            check(lib().spx_replay_read(self._h, int(first), int(n), out.ctypes.data, _stream()), "spx_replay_read")
        return out

    def to_moves(self):
        from .selfplay import records_to_moves
        return records_to_moves(self.read(), self.game)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            torch.cuda.synchronize()
            lib().spx_replay_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def loss_from_batch(network, batch, q_average=True):
    """MCTreeSearch.loss (mcts.py:234-252) on a ``sample_batch`` dict: value MSE (mean) against actual_val (+ q when
    q_average) plus the policy cross-entropy -sum(log p * tree_probs) / B."""
    if "planes" in batch and hasattr(network, "forward_planes"):
        net_probs, predict_val = network.forward_planes(batch["planes"])
    else:
        net_probs, predict_val = network.forward(batch["boards"])
    target = batch["actual_val"] + batch["q"] if q_average else batch["actual_val"]
    value_loss = torch.nn.functional.mse_loss(predict_val.view(-1).float(), target)
    prob_loss = -(net_probs.float().log() * batch["tree_probs"]).sum() / net_probs.size(0)
    return value_loss + prob_loss
