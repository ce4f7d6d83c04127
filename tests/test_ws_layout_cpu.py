"""Host-side checks of two design constants of the warp-specialised gradient kernel (csrc/grad_kernel_ws.cuh); no GPU."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "resilient-consensus-based-marl_b200", "csrc", "grad_kernel_ws.cuh")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_tile_buffer_row_layout_is_bank_conflict_free():
    """The nine operand regions of a buffer row tile the 23 sixteen-byte units without overlap, and the five LDS.128 of a
    consumer step cost the ideal 20 wavefronts under the quarter-warp bank model of tools/ws_bank_layout.py."""
    b = _load(os.path.join(ROOT, "tools", "ws_bank_layout.py"), "ws_bank_layout")
    S, aoff, doff = b.header_layout(HEADER)
    assert S == 23 and S % 2 == 1                       # odd stride: the producers' row-per-thread STS.128 spread too
    used = []
    for a in aoff:
        used += [a, a + 1]
    for d in doff.values():
        used += [d, d + 1, d + 2]
    assert len(used) == len(set(used)) == 22 and max(used) < S
    assert b.cost(S, aoff, doff, b.lm_cur) == 20
    assert b.cost(23, [0, 2, 4, 6, 8], {(0, 0): 10, (0, 1): 13, (1, 0): 16, (1, 1): 19}, b.lm_cur) == 26   # the naive order


def test_tf32x3_split_error_stays_at_fp32_level():
    """3xTF32 with a truncating split of the activations (the shipped form) against fp64: within 4e-7 of max|z| for the two
    layer shapes of the nets, without bias (tools/experiments/tf32x3_error.py restates the kernel's arithmetic in NumPy)."""
    t = _load(os.path.join(ROOT, "tools", "experiments", "tf32x3_error.py"), "tf32x3_error")
    rs = np.random.RandomState(1)
    for K in (16, 24):
        A = rs.uniform(-2, 2, (512, K)).astype(np.float32)
        lim = np.sqrt(6.0 / (K + 20))
        B = rs.uniform(-lim, lim, (K, 20)).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        b_hi = t.tf32_rna(B)
        b_lo = t.tf32_trunc(B - b_hi)
        a_hi = t.tf32_trunc(A)
        a_lo = t.tf32_trunc(A - a_hi)
        x = t.mm32(a_hi, b_hi) + t.mm32(a_lo, b_hi) + t.mm32(a_hi, b_lo)
        scale = np.abs(ref).max()
        assert np.abs(x - ref).max() / scale < 4e-7
        assert abs((x - ref).mean()) / scale < 2e-8
        one_pass = t.mm32(t.tf32_trunc(A), t.tf32_trunc(B))
        assert np.abs(one_pass - ref).max() / scale > 1e-4          # what the split buys
