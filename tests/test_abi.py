"""CPU-only: libspx.so builds for sm_100a, loads, and exports every symbol include/spx.h declares.
No compute calls are made (no GPU here)."""
import ctypes as C
import os
import re

from self_play_reinforcement_learning_b200 import _lib
from self_play_reinforcement_learning_b200 import build as b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        src = open(os.path.join(inc, fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(spx_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_builds_and_exports_header_symbols():
    path = b.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    declared = _declared()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert set(_lib.EXPORTS) <= declared


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Record) == 80 and C.sizeof(_lib.Result) == 16
    assert C.sizeof(_lib.MoveLog) == 4 * 4 + 8 + 9 * 4 + 4 + 9 * 8 * 2
    assert C.sizeof(_lib.Counters) == 80
    assert C.sizeof(_lib.Config) == 14 * 4 + 8 + 8 + 5 * 8


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        return
    L = _lib.lib()
    assert L.spx_version() >= 100
    cfg = _lib.Config()
    cfg.game, cfg.n_games, cfg.sims, cfg.slot_stride = 0, 4, 10, 4
    h = C.c_void_p()
    rc = L.spx_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and b"no CUDA device" in L.spx_last_error()
    import pytest
    from self_play_reinforcement_learning_b200.engine import SelfPlayEngine
    with pytest.raises(_lib.SpxError):
        SelfPlayEngine(game=0, n_games=1, sims=1, evaluator=None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "self_play_reinforcement_learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "spx_oracle" not in src.replace("ox_pow_int_exact", ""), fn
