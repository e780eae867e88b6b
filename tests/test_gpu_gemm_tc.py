"""The complex64 tensor-core GEMM (csrc/gemm_tc.cu: tcgen05.mma kind::tf32 split products, TMA-staged tiles, A planes in
tensor memory) through the C ABI (``qg_gemm``), against a complex128 product of the same operands.

Covers what the expm-chain tests do not pin directly: ragged tile edges (128 x 96 tiles over shapes that are not multiples),
pitched outputs (nothing outside the result may be written), the switch back to the FMA kernel, the coherent (bias) part of
the error that a chain of squarings would amplify, and that operands with entries no TF32 pair can represent exactly still
meet the bar.
"""
import numpy as np
import pytest
import torch

from tests._tol import check, relerr

pytestmark = pytest.mark.gpu
C64 = torch.complex64


@pytest.fixture(scope="module")
def lib():
    from qgrad_b200 import _lib
    assert torch.cuda.is_available()
    return _lib.load()


def _gemm(A, B, C=None, ldc=None):
    from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
    b, m, k = A.shape
    n = B.shape[-1]
    if C is None:
        C = torch.empty((b, m, n), dtype=A.dtype, device=A.device)
        ldc = n
    _call("qg_gemm", _code(A.dtype), _ptr(A), _ptr(B), _ptr(C), b, m, n, k, k, n, ldc, m * k, k * n, C.shape[1] * C.shape[2], _stream())
    return C


def _rand(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, dtype=C64, device="cuda", generator=g)


@pytest.mark.parametrize("b,m,n,k", [(1, 128, 128, 64), (2, 128, 96, 1024), (1, 256, 384, 128), (3, 192, 320, 192), (2, 320, 224, 80),
                                     (1, 1024, 1024, 1024), (4, 512, 1024, 512)])
def test_tensor_gemm_matches_complex128_product(lib, b, m, n, k):
    A, B = _rand((b, m, k), 7 * m + n), _rand((b, k, n), 3 * n + k)
    ref = (A.to(torch.complex128) @ B.to(torch.complex128)).cpu().numpy()
    prev = lib.qg_set_c64_tensor(1)
    try:
        C = _gemm(A, B)
    finally:
        lib.qg_set_c64_tensor(prev)
    check(relerr(C.cpu().numpy(), ref), C64, f"tensor-core gemm {b}x{m}x{n}x{k}")
    if m % 64 or n % 64:
        return                                      # the FMA kernel takes multiples of its 64 x 64 tile only
    # the FMA kernel on the same operands agrees to float32 round-off
    prev = lib.qg_set_c64_tensor(0)
    try:
        C0 = _gemm(A, B)
    finally:
        lib.qg_set_c64_tensor(prev)
    check(relerr(C.cpu().numpy(), C0.cpu().numpy()), C64, f"tensor-core vs FMA gemm {b}x{m}x{n}x{k}")


def test_tensor_gemm_writes_nothing_outside_the_result(lib):
    """ragged tiles (m, n not multiples of 128 / 96) into a pitched output: the padding columns keep their sentinel"""
    b, m, n, k, ldc = 2, 192, 200, 64, 264
    A, B = _rand((b, m, k), 1), _rand((b, k, n), 2)
    C = torch.full((b, m + 8, ldc), 777.0, dtype=C64, device="cuda")
    prev = lib.qg_set_c64_tensor(1)
    try:
        _gemm(A, B, C, ldc)
    finally:
        lib.qg_set_c64_tensor(prev)
    ref = (A.to(torch.complex128) @ B.to(torch.complex128)).cpu().numpy()
    out = C.cpu().numpy()
    check(relerr(out[:, :m, :n], ref), C64, "pitched ragged tensor-core gemm")
    assert np.all(out[:, :m, n:] == 777.0) and np.all(out[:, m:, :] == 777.0)


def test_tensor_gemm_coherent_error_is_removed(lib):
    """<C - C_ref, C_ref> / <C_ref, C_ref>: the truncating accumulator's mean loss (-8.6e-8 per product uncorrected) is what a
    chain of squarings would double at every step; after the correction it is below 1e-8 on dense operands"""
    A, B = _rand((4, 512, 512), 11), _rand((4, 512, 512), 12)
    ref = A.to(torch.complex128) @ B.to(torch.complex128)
    prev = lib.qg_set_c64_tensor(1)
    try:
        C = _gemm(A, B).to(torch.complex128)
    finally:
        lib.qg_set_c64_tensor(prev)
    coh = float((torch.vdot(ref.reshape(-1), (C - ref).reshape(-1)) / torch.vdot(ref.reshape(-1), ref.reshape(-1))).real)
    assert abs(coh) < 1e-8, coh


def test_tensor_gemm_exact_on_small_integers(lib):
    """entries that TF32 holds exactly (Pauli-like 0 / +-1 / +-i, small integers): hi carries everything, lo = 0, and the
    products and sums are exact in float32 -- apart from the 8.6e-8 bias correction, which must stay below one ulp"""
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.complex(torch.randint(-3, 4, (2, 256, 128), device="cuda", generator=g).float(),
                      torch.randint(-3, 4, (2, 256, 128), device="cuda", generator=g).float())
    B = torch.complex(torch.randint(-3, 4, (2, 128, 192), device="cuda", generator=g).float(),
                      torch.randint(-3, 4, (2, 128, 192), device="cuda", generator=g).float())
    ref = (A.to(torch.complex128) @ B.to(torch.complex128)).cpu().numpy()
    prev = lib.qg_set_c64_tensor(1)
    try:
        C = _gemm(A, B).cpu().numpy()
    finally:
        lib.qg_set_c64_tensor(prev)
    assert np.max(np.abs(C - ref) / np.maximum(np.abs(ref), 1.0)) < 2.4e-7      # <= 2 ulp of float32


def test_chain_of_squarings_stays_with_the_fma_kernel(lib):
    """eight dependent squarings of a unitary (the expm chain's worst regime; squaring doubles whatever part of the error
    commutes with the matrix): the tensor path ends within 3x of the FMA kernel's error and inside 1e-4 -- measured 1.4e-5
    against 6.8e-6 at n = 512, 1.3e-5 against 1.2e-5 at n = 1024"""
    n = 512
    g = torch.Generator(device="cuda").manual_seed(5)
    H = torch.randn((n, n), dtype=torch.complex128, device="cuda", generator=g)
    H = (H + H.mH) / 2
    w, V = torch.linalg.eigh(H)
    U0 = ((V * torch.exp(1j * w / 40)) @ V.mH).to(C64).reshape(1, n, n).contiguous()
    errs = {}
    for on in (1, 0):
        prev = lib.qg_set_c64_tensor(on)
        try:
            R, R64 = U0.clone(), U0.to(torch.complex128)
            for _ in range(8):
                R = _gemm(R, R)
                R64 = R64 @ R64
            errs[on] = relerr(R.cpu().numpy(), R64.cpu().numpy())
        finally:
            lib.qg_set_c64_tensor(prev)
    assert errs[1] < 1e-4 and errs[1] < 3.0 * errs[0] + 1e-6, errs


def test_tensor_gemm_in_a_cuda_graph_and_on_side_streams(lib):
    """the tensor maps travel as kernel parameters: a captured launch replays on new operand VALUES in the same buffers, and
    launches on two side streams do not interfere"""
    A, B = _rand((2, 256, 128), 21), _rand((2, 128, 192), 22)
    C = torch.empty((2, 256, 192), dtype=C64, device="cuda")
    prev = lib.qg_set_c64_tensor(1)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _gemm(A, B, C, 192)                               # warm-up outside the capture (first-use attribute call)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            _gemm(A, B, C, 192)
        A.copy_(_rand((2, 256, 128), 23)); B.copy_(_rand((2, 128, 192), 24))
        g.replay()
        torch.cuda.synchronize()
        ref = (A.to(torch.complex128) @ B.to(torch.complex128)).cpu().numpy()
        check(relerr(C.cpu().numpy(), ref), C64, "tensor-core gemm replayed from a CUDA graph")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        outs = []
        for st, seed in ((s1, 31), (s2, 32)):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                a, b = _rand((3, 384, 256), seed), _rand((3, 256, 384), seed + 100)
                outs.append((a, b, _gemm(a, b)))
        torch.cuda.synchronize()
        for a, b, c in outs:
            check(relerr(c.cpu().numpy(), (a.to(torch.complex128) @ b.to(torch.complex128)).cpu().numpy()), C64, "tensor-core gemm on a side stream")
    finally:
        lib.qg_set_c64_tensor(prev)
