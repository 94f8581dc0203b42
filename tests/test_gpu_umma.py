"""Bring-up test of the tcgen05 building block: SWIZZLE_NONE K-major descriptors over channel-group planes,
where a convolution tap is a shifted start address.  bf16 products are exact in fp32, so the only
difference to numpy is the summation order."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16(x):
    """float32 -> bf16 bits (round to nearest even) and the rounded float32 values."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r, (r.astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize("n_cg,N,n_pos,shift", [(2, 64, 128, 0), (2, 64, 200, 37), (8, 128, 180, 5), (4, 16, 300, 171), (6, 112, 160, 32)])
def test_shifted_gemm(n_cg, N, n_pos, shift):
    from trex_b200 import _capi
    rng = np.random.default_rng(n_cg * 1000 + N + shift)
    a_bits, a = _bf16(rng.standard_normal((n_cg, n_pos, 8)))
    b_bits, b = _bf16(rng.standard_normal((n_cg, N, 8)))
    d = np.zeros((128, N), np.float32)
    _capi.check(_capi.lib().tbdbg_umma_shifted_gemm(a_bits.ctypes.data_as(C.c_void_p), n_pos, n_cg, shift,
                                                       b_bits.ctypes.data_as(C.c_void_p), N, d.ctypes.data_as(C.c_void_p)))
    A = a[:, shift:shift + 128, :].transpose(1, 0, 2).reshape(128, n_cg * 8).astype(np.float64)
    B = b.transpose(1, 0, 2).reshape(N, n_cg * 8).astype(np.float64)
    ref = A @ B.T
    assert np.abs(d - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
