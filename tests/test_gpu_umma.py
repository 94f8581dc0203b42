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


def _e5m2(x):
    """float32 -> e5m2 bytes (round to nearest even through float16's top byte) and the rounded values."""
    h = np.ascontiguousarray(x, np.float32).astype(np.float16).view(np.uint16).astype(np.uint32)
    r = ((h + 0x7F + ((h >> 8) & 1)) >> 8).astype(np.uint8)              # RNE on the low mantissa byte (values stay far from inf here)
    return r, (r.astype(np.uint16) << 8).view(np.float16).astype(np.float32)


@pytest.mark.parametrize("n_cg,n_pl8,N,n_pos,shift", [(2, 2, 64, 160, 0), (2, 2, 64, 200, 37), (8, 8, 128, 180, 5), (8, 8, 224, 200, 47)])
def test_mixed_f16_plus_e5m2_accumulate(n_cg, n_pl8, N, n_pos, shift):
    """kind::f16 and kind::f8f6f4 (e5m2, K = 32 over two planes) MMAs into the same fp32 accumulator: the building block of the
    "fp16c" precision (fp16 main term + 8-bit correction terms).  Products of representable operands are exact in fp32."""
    from trex_b200 import _capi
    rng = np.random.default_rng(n_cg * 100 + N + shift)
    a16 = rng.standard_normal((n_cg, n_pos, 8)).astype(np.float16); b16 = rng.standard_normal((n_cg, N, 8)).astype(np.float16)
    a8b, a8 = _e5m2(rng.standard_normal((n_pl8, n_pos, 16)) * 4); b8b, b8 = _e5m2(rng.standard_normal((n_pl8, N, 16)) * 0.25)
    d = np.zeros((128, N), np.float32)
    L = _capi.lib()
    L.tbdbg_umma_mixed_gemm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    _capi.check(L.tbdbg_umma_mixed_gemm(a16.ctypes.data_as(C.c_void_p), n_pos, n_cg, shift, b16.ctypes.data_as(C.c_void_p),
                                        a8b.ctypes.data_as(C.c_void_p), b8b.ctypes.data_as(C.c_void_p), n_pl8, N, d.ctypes.data_as(C.c_void_p)))
    A = a16[:, shift:shift + 128, :].transpose(1, 0, 2).reshape(128, -1).astype(np.float64)
    B = b16.transpose(1, 0, 2).reshape(N, -1).astype(np.float64)
    A8 = a8[:, shift:shift + 128, :].transpose(1, 0, 2).reshape(128, -1).astype(np.float64)
    B8 = b8.transpose(1, 0, 2).reshape(N, -1).astype(np.float64)
    ref = A @ B.T + A8 @ B8.T
    assert np.abs(d - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


def test_tmem_ld_16x256b_layout():
    """The register layout of tcgen05.ld shape 16x256b.x4 (the mma accumulator fragment: a thread holds TMEM lanes t/4 and t/4 + 8 of two
    adjacent columns per 8-column group) -- what conv2's tap-pair epilogue relies on to add lane r + 8 without shuffles."""
    from trex_b200 import _capi
    rng = np.random.default_rng(5)
    n_cg, n_pos, N, shift = 2, 160, 128, 3
    a = rng.standard_normal((n_cg, n_pos, 8)).astype(np.float32); b = rng.standard_normal((n_cg, N, 8)).astype(np.float32)
    import torch
    a16 = torch.from_numpy(a).to(torch.bfloat16); b16 = torch.from_numpy(b).to(torch.bfloat16)
    d = np.zeros((128, N), np.float32); raw = np.zeros((4, N // 32, 2, 32, 16), np.float32)
    L = _capi.lib()
    L.tbdbg_tmem_ld_16x256b.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    _capi.check(L.tbdbg_tmem_ld_16x256b(C.c_void_p(a16.data_ptr()), n_pos, n_cg, shift, C.c_void_p(b16.data_ptr()), N,
                                        d.ctypes.data_as(C.c_void_p), raw.ctypes.data_as(C.c_void_p)))
    exp = np.zeros_like(raw)
    for w in range(4):
        for cb in range(N // 32):
            for lh in range(2):
                for t in range(32):
                    for j in range(4):
                        for h in range(2):
                            for e in range(2):
                                exp[w, cb, lh, t, 4 * j + 2 * h + e] = d[32 * w + 16 * lh + t // 4 + 8 * h, 32 * cb + 8 * j + 2 * (t % 4) + e]
    assert np.array_equal(raw, exp)
