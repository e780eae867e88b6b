"""Per-kernel roofline table for the HBM-bound kernels of the hot path.

    python tools/kernel_rooflines.py [out.json]

Each kernel is called through the C ABI on device-resident inputs whose footprint exceeds L2
(126 MB), timed with CUDA events on the launching stream (3 warm-ups, 10 timed calls), and reported as
ALGORITHMIC bytes / time against the measured copy bandwidth in MEASURED_PEAKS.json (fallback 6650
GB/s, B200_PROFILING.md).  Algorithmic bytes = what must cross HBM once: inputs read + outputs written
(SURVEY.md 8d), nothing for re-reads."""
import ctypes as C
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qgrad_b200 import _lib                                                   # noqa: E402
from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream, _displace_constants  # noqa: E402

try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        PEAK, PEAK_SRC = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
except Exception:
    PEAK, PEAK_SRC = 6650.0, "fallback 6650 GB/s (of fallback)"

rows = []


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def report(kernel, cfg, nbytes, sec, flops=None):
    row = {"kernel": kernel, "config": cfg, "algorithmic_bytes": int(nbytes), "us": sec * 1e6,
           "achieved_gbs": nbytes / sec / 1e9, "frac_hbm": nbytes / sec / 1e9 / PEAK}
    if flops:
        row["tflops"] = flops / sec / 1e12
    rows.append(row)
    print(f"{kernel:28s} {cfg:34s} {row['us']:9.1f} us {row['achieved_gbs']:8.1f} GB/s  {100 * row['frac_hbm']:5.1f}% of HBM copy peak"
          + (f"  {row['tflops']:.2f} TFLOP/s" if flops else ""), flush=True)


def rnd(shape, dt):
    return torch.randn(shape, dtype=dt, device="cuda")


def main():
    dev = torch.device("cuda")
    for cdt, rdt, w, name in ((torch.complex64, torch.float32, 8, "c64"), (torch.complex128, torch.float64, 16, "c128")):
        code = _code(cdt)
        r = w // 2
        # ---- a6: fused rot + fidelity value and gradient (config #2)
        B = 1 << 25
        ang = torch.rand((B, 3), dtype=rdt, device=dev) * 6.28
        ket = torch.tensor([[0.6, 0.8j]], dtype=cdt, device=dev)
        fid = torch.empty(B, dtype=rdt, device=dev); grad = torch.empty((B, 3), dtype=rdt, device=dev)
        t = timeit(lambda: _call("qg_rot_fidelity_fwd_grad", code, _ptr(ang), _ptr(ket), 0, _ptr(fid), _ptr(grad), B, _stream()))
        report("rot_fidelity_fwd_grad", f"{name} B=2^25", B * 7 * r, t)
        R = torch.empty((B, 2, 2), dtype=cdt, device=dev)
        t = timeit(lambda: _call("qg_rot_zyz_fwd", code, _ptr(ang), _ptr(R), B, _stream()))
        report("rot_zyz_fwd", f"{name} B=2^25", B * (3 * r + 4 * w), t)
        Rb = rnd((B, 2, 2), cdt); ab = torch.empty_like(ang)
        t = timeit(lambda: _call("qg_rot_zyz_bwd", code, _ptr(ang), _ptr(Rb), _ptr(ab), B, _stream()))
        report("rot_zyz_bwd", f"{name} B=2^25", B * (3 * r + 4 * w + 3 * r), t)
        del R, Rb, ab, ang, fid, grad
        # ---- a3: expect, ket branch (n = 2 shared operator; n = 8 per-sample operator) and dm branch
        for n, B, ob in ((2, 1 << 25, 0), (8, 1 << 21, 1), (32, 1 << 17, 1)):
            O = rnd((B if ob else 1, n, n), cdt); k = rnd((B, n), cdt); out = torch.empty(B, dtype=cdt, device=dev)
            t = timeit(lambda: _call("qg_expect_ket_fwd", code, _ptr(O), _ptr(k), _ptr(out), B, n, ob, _stream()))
            nb = B * (w * n + w) + (B if ob else 1) * w * n * n
            report("expect_ket_fwd", f"{name} n={n} B=2^{int(math.log2(B))} oper_batched={ob}", nb, t)
            g = rnd((B,), cdt); kb = torch.empty_like(k)
            t = timeit(lambda: _call("qg_expect_ket_bwd", code, _ptr(O), _ptr(k), _ptr(g), _ptr(kb), None, B, n, ob, _stream()))
            report("expect_ket_bwd(ketbar)", f"{name} n={n} B=2^{int(math.log2(B))} oper_batched={ob}", nb + B * w * n, t)
            del O, k, out, g, kb
        for n, B in ((4, 1 << 22), (32, 1 << 16)):
            O = rnd((1, n, n), cdt); rho = rnd((B, n, n), cdt); out = torch.empty(B, dtype=cdt, device=dev)
            t = timeit(lambda: _call("qg_expect_dm_fwd", code, _ptr(O), _ptr(rho), _ptr(out), B, n, 0, _stream()))
            report("expect_dm_fwd", f"{name} n={n} B=2^{int(math.log2(B))}", B * (w * n * n + w), t)
            g = rnd((B,), cdt); rb = torch.empty_like(rho)
            t = timeit(lambda: _call("qg_expect_dm_bwd", code, _ptr(O), _ptr(rho), _ptr(g), _ptr(rb), None, B, n, 0, _stream()))
            report("expect_dm_bwd(rhobar)", f"{name} n={n} B=2^{int(math.log2(B))}", B * (w + w * n * n), t)
            del O, rho, out, g, rb
        # ---- a4: fidelity, ket-ket (config #1 n=8, config #3 n=32, config #5 n=1024) and dm surrogate
        for n, B in ((8, 1 << 22), (32, 1 << 20), (1024, 1 << 15)):
            a = rnd((B, n), cdt); b = rnd((B, n), cdt); out = torch.empty(B, dtype=rdt, device=dev)
            t = timeit(lambda: _call("qg_fidelity_ket_fwd", code, _ptr(a), _ptr(b), _ptr(out), B, n, _stream()))
            report("fidelity_ket_fwd", f"{name} n={n} B=2^{int(math.log2(B))}", B * (2 * w * n + r), t)
            g = torch.rand(B, dtype=rdt, device=dev); abar = torch.empty_like(a); bbar = torch.empty_like(b)
            t = timeit(lambda: _call("qg_fidelity_ket_bwd", code, _ptr(a), _ptr(b), _ptr(g), _ptr(abar), _ptr(bbar), B, n, _stream()))
            report("fidelity_ket_bwd", f"{name} n={n} B=2^{int(math.log2(B))}", B * (4 * w * n + r), t)
            del a, b, out, g, abar, bbar
        n, B = 32, 1 << 16
        r1 = rnd((B, n, n), cdt); r2 = rnd((B, n, n), cdt); out = torch.empty(B, dtype=rdt, device=dev)
        t = timeit(lambda: _call("qg_fidelity_dm_fwd", code, _ptr(r1), _ptr(r2), _ptr(out), B, n, _stream()))
        report("fidelity_dm_fwd (diagonals)", f"{name} n={n} B=2^16", B * (2 * w * n + r), t)
        del r1, r2, out
        # ---- a5: Displace build / apply (config #3)
        for n, B in ((10, 1 << 19), (40, 1 << 15)):
            V, lam = _displace_constants(n, cdt)
            alpha = rnd((B,), cdt); D = torch.empty((B, n, n), dtype=cdt, device=dev)
            t = timeit(lambda: _call("qg_displace_fwd", code, _ptr(V), _ptr(lam), _ptr(alpha), _ptr(D), B, n, _stream()))
            report("displace_fwd", f"{name} n={n} B=2^{int(math.log2(B))}", B * (w + w * n * n), t, flops=B * 4.0 * n ** 3)
            Db = rnd((B, n, n), cdt); ab = torch.empty_like(alpha)
            t = timeit(lambda: _call("qg_displace_bwd", code, _ptr(V), _ptr(lam), _ptr(alpha), _ptr(Db), _ptr(ab), B, n, _stream()))
            report("displace_bwd", f"{name} n={n} B=2^{int(math.log2(B))}", B * (2 * w + w * n * n), t, flops=B * 8.0 * n ** 3)
            del D, Db
            B2 = B * 8
            alpha = rnd((B2,), cdt); x = rnd((B2, n), cdt); y = torch.empty_like(x)
            t = timeit(lambda: _call("qg_displace_apply_fwd", code, _ptr(V), _ptr(lam), _ptr(alpha), _ptr(x), _ptr(y), B2, n, _stream()))
            report("displace_apply_fwd", f"{name} n={n} B=2^{int(math.log2(B2))}", B2 * (w + 2 * w * n), t, flops=B2 * 12.0 * n * n)
            yb = rnd((B2, n), cdt); ab = torch.empty_like(alpha); xb = torch.empty_like(x)
            t = timeit(lambda: _call("qg_displace_apply_bwd", code, _ptr(V), _ptr(lam), _ptr(alpha), _ptr(x), _ptr(yb), _ptr(ab), _ptr(xb), B2, n, _stream()))
            report("displace_apply_bwd", f"{name} n={n} B=2^{int(math.log2(B2))}", B2 * (2 * w + 4 * w * n), t)
            del alpha, x, y, yb, ab, xb
        # ---- a6: Givens-product Unitary (config #1: N=8) and _make_rot
        for N, B in ((8, 1 << 20), (16, 1 << 17)):
            K = N * (N - 1) // 2
            th = torch.rand((B, K), dtype=rdt, device=dev) * 6.28; ph = torch.rand((B, K), dtype=rdt, device=dev) * 6.28
            om = torch.rand((B, N), dtype=rdt, device=dev) * 6.28
            U = torch.empty((B, N, N), dtype=cdt, device=dev)
            t = timeit(lambda: _call("qg_givens_unitary_fwd", code, _ptr(th), _ptr(ph), _ptr(om), _ptr(U), B, N, _stream()))
            report("givens_unitary_fwd", f"{name} N={N} B=2^{int(math.log2(B))}", B * ((2 * K + N) * r + w * N * N), t, flops=B * 6.0 * N ** 3)
            Ub = rnd((B, N, N), cdt); tb, pb, ob_ = torch.empty_like(th), torch.empty_like(ph), torch.empty_like(om)
            t = timeit(lambda: _call("qg_givens_unitary_bwd", code, _ptr(th), _ptr(ph), _ptr(om), _ptr(Ub), _ptr(tb), _ptr(pb), _ptr(ob_), B, N, _stream()))
            report("givens_unitary_bwd", f"{name} N={N} B=2^{int(math.log2(B))}", B * (2 * (2 * K + N) * r + w * N * N), t)
            del th, ph, om, U, Ub, tb, pb, ob_
        # ---- fused Unitary-learning cost (config #1: N = 8, 80 training pairs, batches of independent learners)
        N, M, B = 8, 80, 1 << 17
        K = N * (N - 1) // 2
        th = torch.rand((B, K), dtype=rdt, device=dev) * 6.28; ph = torch.rand((B, K), dtype=rdt, device=dev) * 6.28
        om = torch.rand((B, N), dtype=rdt, device=dev) * 6.28
        kin = rnd((M, N), cdt); kout = rnd((M, N), cdt)
        loss = torch.empty(B, dtype=rdt, device=dev); tb, pb, ob_ = torch.empty_like(th), torch.empty_like(ph), torch.empty_like(om)
        t = timeit(lambda: _call("qg_unitary_cost_fwd_grad", code, _ptr(th), _ptr(ph), _ptr(om), _ptr(kin), _ptr(kout), M, _ptr(loss),
                                 _ptr(tb), _ptr(pb), _ptr(ob_), B, N, _stream()))
        report("unitary_cost_fwd_grad", f"{name} N=8 M=80 B=2^17", B * (2 * (2 * K + N) + 1) * r, t, flops=B * (M * N * N * 16.0 + 12.0 * N ** 3))
        rows[-1]["parameter_sets_per_s"] = B / t
        del th, ph, om, loss, tb, pb, ob_
        N, B = 8, 1 << 20
        prm = torch.rand((B, 2), dtype=rdt, device=dev); Rm = torch.empty((B, N, N), dtype=cdt, device=dev)
        t = timeit(lambda: _call("qg_make_rot_fwd", code, _ptr(prm), _ptr(Rm), B, N, 1, 5, _stream()))
        report("make_rot_fwd", f"{name} N=8 B=2^20", B * (2 * r + w * N * N), t)
        del prm, Rm
        # ---- a7: learning-step glue at config #5 shapes (8 settings, 512 kets each)
        nq, S, m = 10, 8, 512
        n = 1 << nq
        P = 38
        theta = torch.rand(P, dtype=rdt, device=dev); ctrl = torch.rand((S, P), dtype=rdt, device=dev)
        xm = torch.randint(0, n, (P,), dtype=torch.int32, device=dev); zm = torch.randint(0, n, (P,), dtype=torch.int32, device=dev)
        A = torch.empty((S, n, n), dtype=cdt, device=dev)
        t = timeit(lambda: _call("qg_pauli_hamiltonian", code, _ptr(theta), _ptr(ctrl), _ptr(xm), _ptr(zm), P, S, nq, 1.0, _ptr(A), _stream()))
        report("pauli_hamiltonian (+memset)", f"{name} 10 qubits, 8 settings, 38 terms", S * w * n * n, t)
        tb = torch.empty(P, dtype=rdt, device=dev)
        t = timeit(lambda: _call("qg_pauli_project", code, _ptr(A), _ptr(ctrl), _ptr(xm), _ptr(zm), P, S, nq, 1.0, _ptr(tb), 0, _stream()))
        report("pauli_project", f"{name} 10 qubits, 8 settings, 38 terms", S * n * P * w, t)
        phi = rnd((S, n, m), cdt); out = rnd((S, n, m), cdt); G = torch.empty_like(phi); loss = torch.zeros((), dtype=rdt, device=dev)
        t = timeit(lambda: _call("qg_overlap_loss", code, _ptr(phi), _ptr(out), _ptr(G), _ptr(loss), S, n, m, 1.0 / (S * m), _stream()))
        report("overlap_loss", f"{name} n=1024, 8 x 512 kets", S * n * m * 3 * w, t)
        del A, phi, out, G
        torch.cuda.empty_cache()
    path = sys.argv[1] if len(sys.argv) > 1 else None
    if path:
        with open(path, "w") as f:
            json.dump({"peak_hbm_gbs": PEAK, "peak_source": PEAK_SRC, "gpu": torch.cuda.get_device_name(0), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
