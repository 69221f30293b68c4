"""The integer-pipe binary64 arithmetic of csrc/spx_softf64.cuh (what the fused tick kernel's shadow warp computes the PUCT scores
with, because FP64 instructions next to running tcgen05 MMAs slow the tensor pipe down) against the FP64 instructions themselves:
bit-identical results for every operation on 2 x 10^8 operand sets (close exponents so that sums cancel and round, zeros, powers of
two, exact ties, divisors up to 2^31, float and 53-bit conversions, comparisons)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OPS = ("mul", "add", "div by integer", "sqrt of integer", "f32 -> f64", "uniform from 53 bits", "comparisons", "scaling / negation")


@pytest.mark.parametrize("seed", [1, 2])
def test_integer_pipe_fp64_equals_the_fp64_instructions(seed):
    from self_play_reinforcement_learning_b200 import _lib
    torch.cuda.set_device(0)
    out = np.zeros(8, np.uint64)
    _lib.check(_lib.lib().spx_softf64_selftest(100_000_000, seed, C.c_void_p(out.ctypes.data),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)), "spx_softf64_selftest")
    assert not out.any(), dict(zip(OPS, out.tolist()))
