"""On-disk formats of the reference, so existing runs can be resumed and produced.

* model checkpoints: ``torch.save({"model": state_dict}, <save_dir>/<start_time>/model-<iso timestamp>:<games>)``
  (updateworker.py:111-117, name built at self_play_parallel.py:263-267)
* replay memory: ``pickle.dump(Memory)`` to ``memory-<iso timestamp>:<size>`` with the previous file removed
  (updateworker.py:119-139); written with the reference's class paths (``rl_utils.memory.Memory``, ``games.algos.mcts.Move``)
  so either side opens the other's files
* discovery: lexicographically newest non-empty run folder, then the lexicographically newest file with the prefix
  (base_worker.py:44-62)
Checkpoints feed the engine through ``nets.pack_tower_blob(module)`` / ``BatchedSelfPlay.load_weights``.
"""
import datetime
import os
import pickle

import torch


def model_file_name(save_dir, start_time, games_played, now=None):
    now = now or datetime.datetime.now()
    return os.path.join(save_dir, start_time, "model-" + now.isoformat() + ":" + str(games_played))


def save_model(network, saved_name):
    os.makedirs(os.path.dirname(saved_name), exist_ok=True)
    torch.save({"model": network.state_dict()}, saved_name)
    return saved_name


def load_model(network, model_file, map_location="cpu"):
    checkpoint = torch.load(model_file, map_location=map_location)
    network.load_state_dict(checkpoint["model"])
    return network


# ---- pickle interop with the reference ---------------------------------------------------------------------------------
# The reference pickles ``rl_utils.memory.Memory`` objects holding ``games.algos.mcts.Move`` tuples (updateworker.py:119-139)
# and opens them with plain ``pickle.load`` (base_worker.py:36-42).  Files written here carry exactly those class paths, so
# the unmodified reference opens them; files written by the reference load here without its tree (or anytree) on sys.path.
_REF_MEMORY, _REF_MCTS = "rl_utils.memory", "games.algos.mcts"


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        from . import scheduler, selfplay
        local = {(_REF_MEMORY, "Memory"): scheduler.Memory, (_REF_MEMORY, "Deduplicator"): scheduler.DeduplicatorState,
                 (_REF_MCTS, "Move"): selfplay.Move}.get((module, name))
        return local if local is not None else super().find_class(module, name)


class _reference_class_paths:
    """Context manager: makes ``rl_utils.memory.Memory`` / ``games.algos.mcts.Move`` resolvable while pickling.  When the
    reference's tree is importable its own classes are used; otherwise bare stand-ins with the same module and name are
    installed in sys.modules for the duration of the dump (pickle stores only the path, never the class body)."""

    def __enter__(self):
        import collections
        import sys
        import types
        self.installed = []
        try:
            import importlib
            self.Memory = importlib.import_module(_REF_MEMORY).Memory
            self.Move = importlib.import_module(_REF_MCTS).Move
            return self
        except Exception:
            pass
        Memory = type("Memory", (), {"__module__": _REF_MEMORY})
        Move = collections.namedtuple("Move", ("state", "actual_val", "tree_probs", "q"), module=_REF_MCTS)   # mcts.py:17
        for name, attrs in (("rl_utils", {}), (_REF_MEMORY, {"Memory": Memory}), ("games", {}), ("games.algos", {}),
                            (_REF_MCTS, {"Move": Move})):
            if name not in sys.modules:
                mod = types.ModuleType(name)
                sys.modules[name] = mod
                self.installed.append(name)
            for k, v in attrs.items():
                setattr(sys.modules[name], k, v)
        self.Memory, self.Move = Memory, Move
        return self

    def __exit__(self, *exc):
        import sys
        for name in self.installed:
            sys.modules.pop(name, None)
        return False


def dump_memory(memory, f):
    """Pickles a Memory of Move tuples with the reference's class paths and attribute names (memory.py:9-12)."""
    from collections import deque
    with _reference_class_paths() as ref:
        out = ref.Memory.__new__(ref.Memory)
        out.__dict__.update(max_size=memory.max_size, deduplicator=None,
                            _buffer=deque((ref.Move(*m) for m in memory._buffer), maxlen=memory._buffer.maxlen))
        pickle.dump(out, f)


def save_memory(memory, save_dir, start_time, previous=None, now=None):
    now = now or datetime.datetime.now()
    name = os.path.join(save_dir, start_time, "memory-" + now.isoformat() + ":" + str(len(memory)))
    os.makedirs(os.path.dirname(name), exist_ok=True)
    with open(name, "wb") as f:
        dump_memory(memory, f)
    if previous and os.path.exists(previous):
        os.remove(previous)
    return name


def load_memory(memory_file):
    """A memory file of this package or of the reference -> scheduler.Memory of selfplay.Move tuples."""
    with open(memory_file, "rb") as f:
        return _RefUnpickler(f).load()


def recent_save_file(save_dir, start_time=None, prev_run=False, starting_str="model"):
    """base_worker.py:44-62: newest (by name) non-empty run folder -- excluding ``start_time`` when ``prev_run`` -- and
    in it the newest (by name) file whose name starts with ``starting_str``."""
    folders = [os.path.join(save_dir, f) for f in os.listdir(save_dir)
               if not os.path.isfile(os.path.join(save_dir, f)) and not (prev_run and f == start_time)]
    non_empty = [f for f in folders if os.listdir(f)]
    recent_folder = max(non_empty)
    saves = [os.path.join(recent_folder, f) for f in os.listdir(recent_folder)
             if os.path.isfile(os.path.join(recent_folder, f)) and f.startswith(starting_str)]
    return max(saves)
