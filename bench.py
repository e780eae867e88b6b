#!/usr/bin/env python
"""bench.py -- the qgrad hot path on B200: learning steps/s (and expm+VJP matrices/s vs dim).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the CPU oracle on the host cores

Headline workload (BASELINE.json config #5): 10-qubit (dim 1024) Hamiltonian learning in
complex128, 4096 training kets.  The problem has J = 8 control settings
H_s(theta) = sum_p ctrl[s,p] theta_p P_p (38 Pauli terms, shared theta), 512 kets per setting;
the loss is the mean over all 4096 kets (cost of examples/Unitary-Learning-qgrad.py:123-151).
Strong scaling: the J settings (with their kets) are sharded over the ranks, theta is replicated,
one NCCL all-reduce of P+1 reals per step combines the shared-parameter gradient.  A "step" is
forward expm, ket products, loss, cotangent, Frechet pull-back, projection, all-reduce, Adam.

One JSON line on stdout (rank 0).  value = steps/s with inputs resident in HBM; e2e = the same
step with that step's kets copied from pinned host memory and the loss read back, inside the
timed region.  roofline = the dominant kernel (the FP64 tensor-pipe complex GEMM).  sweep =
the first half of BASELINE.json's metric: fused expm+VJP matrices/s vs dim.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "reference" in sys.argv:
    # the CPU arm gets every host thread: torchrun exports OMP_NUM_THREADS=1, and the BLAS libraries under
    # NumPy / SciPy read the variable once, when they are loaded (i.e. at the imports below)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())

import numpy as np  # noqa: E402

N_QUBITS, N_SETTINGS, KETS_PER_SETTING, T_EVOLVE = 10, 8, 512, 1.0
METRIC, UNIT = "learning steps/s (dim-1024 c128 Hamiltonian learning, 4096 kets)", "steps/s"
REF_BUDGET_S = 240.0        # the CPU arm times whole steps when steps + warmup fit this, else a stated sample per step
DATA_SEED = 1234            # Philox seed of the synthetic ket set (device-side generator)
E2E_LAG = 2                 # steps between launching a step and reading its loss on the host (pinned ring)


def pauli_terms(nq):
    """ZZ and XX on the nq-1 nearest-neighbour bonds, X and Z on every site (38 terms at nq=10)."""
    xm, zm = [], []
    for b in range(nq - 1):
        m = (1 << b) | (1 << (b + 1))
        xm += [0, m]; zm += [m, 0]
    for s in range(nq):
        xm += [1 << s, 0]; zm += [0, 1 << s]
    return np.array(xm, dtype=np.int32), np.array(zm, dtype=np.int32)


def problem(seed=0):
    rng = np.random.default_rng(seed)
    xm, zm = pauli_terms(N_QUBITS)
    P = len(xm)
    ctrl = rng.uniform(0.5, 1.5, (N_SETTINGS, P))
    theta_true = rng.uniform(-1.0, 1.0, P)
    theta0 = theta_true + 0.2 * rng.standard_normal(P)
    return xm, zm, ctrl, theta_true, theta0


def base_config(n_gpus):
    return {"workload": "hamiltonian_learning_dim1024_c128", "n_qubits": N_QUBITS, "dim": 1 << N_QUBITS,
            "settings": N_SETTINGS, "kets_total": N_SETTINGS * KETS_PER_SETTING, "pauli_terms": len(pauli_terms(N_QUBITS)[0]),
            "sharding": f"{N_SETTINGS} settings (512 kets each) over {n_gpus} rank(s); 1 all-reduce of P+1 reals/step",
            "optimizer": "adam(1e-2)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mxc = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = mxc
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def fp64_peak():
    """FP64 tensor-pipe ceiling measured on this pool (tools/fp_peaks.cu -> profiles/r01_fp_peaks.json);
    MEASURED_PEAKS.json carries only HBM and bf16 figures."""
    path = os.path.join(ROOT, "profiles", "r01_fp_peaks.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["dmma_m8n8k4_tflops_sustained"]), "profiles/r01_fp_peaks.json (DMMA m8n8k4 microbenchmark, this pool)"
    except Exception:
        return 37.0, "fallback 37.0 TFLOP/s (64 FP64 FMA/clk/SM x 148 SM x 1.965 GHz)"


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_setting(ents, ctrl_row, theta, psi_in, psi_out, m_total):
    """One control setting of one step on the host cores with the oracle's algorithm (oracle/examples_ref.py
    pauli_learning_loss_and_grad_frechet: sparse Hamiltonian assembly, SciPy Pade expm, ket GEMMs, loss, cotangent,
    SciPy expm_frechet pull-back, projection on the Pauli terms): returns (seconds, thetabar contribution, sum |Re o|)."""
    import scipy.linalg
    n = 1 << N_QUBITS
    t0 = time.perf_counter()
    H = np.zeros((n, n), dtype=np.complex128)
    for p, (r, c, v) in enumerate(ents):
        H[r, c] += ctrl_row[p] * theta[p] * v
    A = -1j * T_EVOLVE * H
    U = scipy.linalg.expm(A)
    phi = U @ psi_in
    o = np.sum(np.conj(psi_out) * phi, axis=0)
    G = (-np.sign(o.real) / m_total)[None, :] * psi_out
    Ubar = G @ psi_in.conj().T
    Abar = scipy.linalg.expm_frechet(A.conj().T, Ubar, compute_expm=False)
    g = np.array([ctrl_row[p] * float(np.real(np.sum(np.conj(Abar[r, c]) * (-1j * T_EVOLVE * v)))) for p, (r, c, v) in enumerate(ents)])
    return time.perf_counter() - t0, g, float(np.sum(np.abs(o.real)))


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count()


def run_reference(args):
    """CPU arm: the oracle's learning step on the host cores.  A step is all N_SETTINGS settings when the whole
    --steps/--warmup run fits REF_BUDGET_S at the measured per-setting time, else a bounded sample of g settings per
    step scaled by N_SETTINGS / g; the line says which, with the seconds actually measured."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scipy.linalg  # noqa: F401  (load its BLAS before the limits below)
    from threadpoolctl import threadpool_limits
    from oracle.examples_ref import pauli_entries
    threadpool_limits(limits=os.cpu_count())
    xm, zm, ctrl, theta_true, theta0 = problem()
    ents = [pauli_entries(int(x), int(z), N_QUBITS) for x, z in zip(xm, zm)]
    rng = np.random.default_rng(1)
    n = 1 << N_QUBITS
    psi = rng.standard_normal((n, KETS_PER_SETTING)) + 1j * rng.standard_normal((n, KETS_PER_SETTING))
    psi /= np.linalg.norm(psi, axis=0, keepdims=True)
    out = rng.standard_normal((n, KETS_PER_SETTING)) + 1j * rng.standard_normal((n, KETS_PER_SETTING))
    out /= np.linalg.norm(out, axis=0, keepdims=True)
    m_total = N_SETTINGS * KETS_PER_SETTING
    t_wall0 = time.perf_counter()
    t1 = cpu_setting(ents, ctrl[0], theta0, psi, out, m_total)[0]          # untimed first touch (BLAS threads, page faults)
    t1 = cpu_setting(ents, ctrl[0], theta0, psi, out, m_total)[0]
    total_steps = args.steps + args.warmup
    g = N_SETTINGS if total_steps * N_SETTINGS * t1 <= REF_BUDGET_S else max(1, min(N_SETTINGS, int(REF_BUDGET_S / (total_steps * t1))))

    def one_step():
        return sum(cpu_setting(ents, ctrl[s], theta0, psi, out, m_total)[0] for s in range(g))

    for _ in range(args.warmup):
        one_step()
    ts = [one_step() for _ in range(args.steps)]
    measured = float(np.sum(ts))
    factor = N_SETTINGS / g
    sec_per_step = float(np.mean(ts)) * factor
    value = 1.0 / sec_per_step
    sample = (f"every step = all {N_SETTINGS} settings (scipy.linalg.expm + expm_frechet at dim 1024, 512 kets each), nothing extrapolated"
              if g == N_SETTINGS else
              f"every step = {g} of {N_SETTINGS} settings (scipy.linalg.expm + expm_frechet at dim 1024, 512 kets each), time x{factor:g}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic", "config": base_config(args.gpus),
            "measured_seconds_timed_region": measured, "measured_ms_per_step_sample": float(np.mean(ts)) * 1e3,
            "settings_timed_per_step": g, "extrapolation_factor": factor,
            "wall_seconds_total": time.perf_counter() - t_wall0,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                             "sample": sample + f"; host has {os.cpu_count()} cpus"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = CPU restatement (oracle/): the reference itself needs JAX, which is not installable here"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def plan_squarings(torch, A, dtype_name):
    """(order m, squarings s) per matrix exactly as the kernels choose them (common.cuh pade_plan +
    the normal-matrix rule of expm_small.cu / expm_large.cu), evaluated with torch on the device;
    used for the flop accounting of the sweep only."""
    n = A.shape[-1]
    c128 = dtype_name == "c128"
    bounds = [(3, 1.08e-2), (5, 2.00e-1), (7, 7.83e-1)] if c128 else [(3, 3.4e-1), (5, 1.3)]
    top_m, top = bounds[-1]
    k = 6 if c128 else 4
    n1 = A.abs().sum(dim=-2).amax(dim=-1).double()
    if n <= 4:                                   # packed slots share one plan: 4 (n<=2) / 2 (n<=4) matrices per slot
        g = 4 if n <= 2 else 2
        pad = (-n1.numel()) % g
        n1g = torch.cat([n1, n1[-1:].expand(pad)]).view(-1, g).amax(dim=1, keepdim=True).expand(-1, g).reshape(-1)
        n1 = n1g[: n1.numel()]
    s_hi = torch.clamp(torch.ceil(torch.log2(torch.clamp(n1 / top, min=1e-300))), min=0)
    m = torch.full_like(n1, float(top_m))
    if n <= 32:
        for mm, bd in reversed(bounds):
            m = torch.where(n1 <= bd, torch.full_like(m, float(mm)), m)
    tot = A.abs().sum(dim=(-1, -2)).double()
    tol = 9.1e-13 if c128 else 3.8e-6
    d_skew = (A + A.mH).abs().sum(dim=(-1, -2)).double()
    d_herm = (A - A.mH).abs().sum(dim=(-1, -2)).double()
    normal = (s_hi > 0) & (torch.minimum(d_skew, d_herm) <= tol * tot) & (n > 8)
    spre = torch.where(normal, torch.clamp(s_hi - 6, min=0), s_hi)
    Ap = A * torch.exp2(-spre).to(A.real.dtype)[:, None, None]
    Pk = Ap @ Ap
    Pk = Pk @ Pk if k == 4 else (Pk @ Pk) @ Pk
    dk = Pk.abs().sum(dim=-2).amax(dim=-1).double() ** (1.0 / k)
    x = torch.minimum(dk, n1 * torch.exp2(-spre))
    e = torch.clamp(torch.ceil(torch.log2(torch.clamp(x / top, min=1e-300))), min=0)
    s = torch.where(normal, spre + e, s_hi)
    return m.cpu().numpy(), s.cpu().numpy()


def gemm_equivalents(m, s):
    """n^3-complex-MAC products of the fused value+VJP algorithm per matrix:
    16/13/10 + 3 s for order m = 7/5/3 (inverse included)."""
    base = np.where(np.asarray(m) >= 7, 16.0, np.where(np.asarray(m) >= 5, 13.0, 10.0))
    return base + 3.0 * np.asarray(s, dtype=np.float64)


def time_calls(torch, fn, reps, flush=None):
    """mean seconds per call over ``reps`` calls, CUDA events on the current stream.  With ``flush`` (a buffer larger
    than L2) every call is timed on its own after the buffer has been rewritten, for inputs that would sit in L2."""
    fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3 / reps
    tot = 0.0
    for _ in range(reps):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / 1e3 / reps


def sweep_expm_vjp(torch, peak_tf, full, with_cpu):
    """expm+VJP matrices/s (c128, plus c64) vs dim -- the first half of BASELINE.json's metric (config #4).
    Default: bounded batches (the HBM-filling batch capped at ~0.25 s of work per call), 10 timed calls per point;
    ``full``: the HBM-filling batch (half of 180 GB over A, Ybar, Y, Abar) streamed through a 40 GB workspace in chunks,
    3 timed calls.  Inputs exceed L2 at every point.  c128 rows carry a SciPy expm + expm_frechet sample on the host."""
    from qgrad_b200.qgrad_qutip import expm_fwd_vjp
    from qgrad_b200 import _lib
    from oracle import expm_ref as E
    rows = []
    rate_guess = {"c128": 30e12, "c64": 40e12}
    for dt, name, esz in ((torch.complex128, "c128", 16), (torch.complex64, "c64", 8)):
        for n in (4, 8, 16, 32, 64, 128, 256, 512, 1024):
            hbm_batch = int(180e9 / 2 / (4 * esz * n * n))
            if full:
                batch = hbm_batch
            else:
                work = 8.0 * max(n, 16) ** 3 * 19                                       # padded flops per matrix, roughly
                batch = max(8, int(min(hbm_batch, 0.25 * rate_guess[name] / work, 2e9 / (4 * esz * n * n))))
                batch = max(batch, int(math.ceil(4 * 126e6 / (4 * esz * n * n))))       # >= 4x L2 of inputs + outputs
            chunk = None
            if n > 32:
                ws1 = _lib.load().qg_expm_workspace_bytes(0 if name == "c64" else 1, n, 1, 1, 2)
                chunk = max(1, int(40e9 // ws1))
            g = torch.Generator(device="cuda").manual_seed(n)
            X = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
            A = (-0.5j / math.sqrt(n)) * (X + X.mH)
            del X
            G = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
            norms = A.abs().sum(dim=-2).amax(dim=-1).double().cpu().numpy()
            s1 = int(max(0, math.ceil(math.log2(max(float(norms.max()) / (0.783 if name == "c128" else 1.3), 1e-30)))))
            if n <= 32:
                m_used, s_used = plan_squarings(torch, A[: min(batch, 4096)], name)
                m_used = np.resize(m_used, batch); s_used = np.resize(s_used, batch)
            else:                                                # read back what the device chose
                nb = min(batch, 4)
                s_used = expm_fwd_vjp(A[:nb], G[:nb], max_squarings=max(1, s1), return_plan=True)[2].cpu().numpy()
                s_used = np.full(batch, float(s_used.max()))
                m_used = np.full(batch, 7.0 if name == "c128" else 5.0)
            prods = gemm_equivalents(m_used, s_used)             # per matrix, incl. the inverse
            s = int(s_used.max())
            ms = max(1, s)
            ws = None
            if n > 32:
                per = _lib.load().qg_expm_workspace_bytes(0 if name == "c64" else 1, n, 1, 1, ms)
                chunk = max(1, min(batch, int(40e9 // per)))
                ws = torch.empty(_lib.load().qg_expm_workspace_bytes(0 if name == "c64" else 1, n, chunk, 1, ms), dtype=torch.uint8, device="cuda")
            Y, Ab = torch.empty_like(A), torch.empty_like(A)
            from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream

            def call():
                _call("qg_expm_fwd_vjp", _code(dt), _ptr(A), _ptr(G), _ptr(Y), _ptr(Ab), batch, n, 0,
                      _ptr(ws), 0 if ws is None else ws.numel(), ms, _stream())
            reps = 3 if full else 10
            sec = time_calls(torch, call, reps)
            flops = 8.0 * n ** 3 * float(prods.sum())            # algorithmic: unpadded n
            row = {"dim": n, "dtype": name, "batch": batch, "reps": reps, "workspace_chunk": chunk, "max_squarings": s,
                   "squarings_1norm_rule": s1, "mean_gemm_equivalents": float(prods.mean()), "matrices_per_s": batch / sec,
                   "tflops": flops / sec / 1e12, "hbm_gbs": 4.0 * esz * n * n * batch / sec / 1e9}
            if name == "c128":
                row["frac_fp64_peak"] = row["tflops"] / peak_tf
                if with_cpu:
                    k = max(1, min(batch, int(2e8 / n ** 3) + 1, 64))
                    a, gg = A[:k].cpu().numpy(), G[:k].cpu().numpy()
                    t0 = time.perf_counter()
                    E.expm_vjp(a, gg); E.expm_scipy(a)
                    row["cpu_matrices_per_s"] = k / (time.perf_counter() - t0)
                    row["cpu_sample"] = f"{k} matrices, scipy.linalg.expm + expm_frechet"
            else:
                # complex64: n >= 128 runs its products on the tcgen05 tensor cores as split-TF32 (3 TF32 MMAs per float32
                # product): the ceiling is a third of the TF32 rate = a sixth of the measured bf16 GEMM rate
                row["gemm_path"] = "tcgen05 kind::tf32 x3, TMA-staged, FP32 round-to-nearest accumulation (gemm_tc.cu)" if n >= 128 \
                    else "shared-memory / FP32 FMA kernels"
                if n >= 128:
                    row["frac_tf32x3_peak"] = row["tflops"] / tf32x3_peak()[0]
            rows.append(row)
            del A, G, Y, Ab, ws
            torch.cuda.empty_cache()
    return rows


def tf32x3_peak():
    """float32-equivalent ceiling of the split-TF32 complex64 GEMM: measured bf16 GEMM rate / 2 (TF32) / 3 (products)"""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]) / 6.0, "MEASURED_PEAKS.json bf16_tflops / 2 (tf32) / 3 (split products)"
    except Exception:
        return 1590.0 / 6.0, "B200_PROFILING.md fallback 1.59 PFLOP/s bf16 / 2 / 3"


def c64_gemm_record(torch):
    """the complex64 tensor-core GEMM alone (qg_gemm), batch 59 x 1024^3 (inputs + outputs 1.5 GB > L2), against the FP32 FMA
    kernel and torch.matmul (cuBLAS CGEMM) on the same operands; error against the complex128 product"""
    from qgrad_b200 import _lib, learning
    lib = _lib.load()
    b, n = 59, 1024
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn((b, n, n), dtype=torch.complex64, device="cuda", generator=g)
    B = torch.randn((b, n, n), dtype=torch.complex64, device="cuda", generator=g)
    C = torch.empty_like(A)
    ref = A[:2].to(torch.complex128) @ B[:2].to(torch.complex128)
    rec = {"kernel": "qg::gemm::tc::gemm_tc_kernel", "batch": b, "m": n, "n": n, "k": n, "dtype": "c64",
           "l2_policy": "operands + result 1.5 GB exceed L2"}
    flops = 8.0 * b * n ** 3
    for name, on in (("tensor", 1), ("fma", 0)):
        prev = lib.qg_set_c64_tensor(on)
        sec = time_calls(torch, lambda: learning.gemm(A, B, out=C), 5)
        lib.qg_set_c64_tensor(prev)
        rec[name + "_tflops"] = flops / sec / 1e12
        rec[name + "_relerr_vs_c128"] = float(torch.linalg.norm((C[:2].to(torch.complex128) - ref).reshape(-1)) / torch.linalg.norm(ref.reshape(-1)))
    sec = time_calls(torch, lambda: torch.matmul(A, B, out=C), 5)
    rec["cublas_cgemm_tflops"] = flops / sec / 1e12
    peak, src = tf32x3_peak()
    rec["roofline"] = {"bound": "tensor", "achieved": rec["tensor_tflops"], "peak": peak, "unit": "TFLOP/s (float32-equivalent)",
                       "frac": rec["tensor_tflops"] / peak, "peak_source": src}
    return rec


def config_records(torch, hbm_peak, fp64_peak_tf):
    """BASELINE configs #1-#3 at their stated sizes through the public API, each with its own roofline fraction and a
    CPU-oracle sample.  Inputs smaller than L2 are timed with an L2 flush between calls."""
    import qgrad_b200.qgrad_qutip as q
    from oracle import examples_ref as ex
    from oracle import qgrad_ref as qr
    recs = []
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")       # 256 MB > L2 (126 MB)
    rng = np.random.default_rng(5)

    # ---- config #1: examples/Unitary-Learning-qgrad.py:123-151 -- N = 8, 80 training pairs, 10^4 independent learners
    N, M = 8, 80
    K = N * (N - 1) // 2
    for cdt, rdt, name, r, B in ((torch.complex64, torch.float32, "c64", 4, 10_000), (torch.complex128, torch.float64, "c128", 8, 10_000),
                                 (torch.complex128, torch.float64, "c128", 8, 1 << 17)):
        th, ph = torch.rand((B, K), dtype=rdt, device="cuda") * 6.28, torch.rand((B, K), dtype=rdt, device="cuda") * 6.28
        om = torch.rand((B, N), dtype=rdt, device="cuda") * 6.28
        kin = torch.randn((M, N), dtype=cdt, device="cuda"); kin /= torch.linalg.norm(kin, dim=1, keepdim=True)
        kout = kin @ q.rand_unitary(N, seed=3, dtype=cdt).T
        th.requires_grad_(True); ph.requires_grad_(True); om.requires_grad_(True)

        def call():
            loss = q.unitary_learning_cost(th, ph, om, kin, kout, dtype=cdt)
            torch.autograd.grad(loss.sum(), (th, ph, om))
        sec = time_calls(torch, call, 10, flush)
        flops = B * (M * N * N * 16.0 + 12.0 * N ** 3)
        nbytes = B * (2 * (2 * K + N) + 1) * r
        rec = {"config": "#1 unitary_learning_cost value+grad", "dtype": name, "N": N, "kets": M, "parameter_sets": B,
               "parameter_sets_per_s": B / sec, "us": sec * 1e6, "l2_policy": "256 MB flush before every timed call",
               "roofline": {"bound": "fp64-fma (accumulation in FP64; operands stay in registers)", "achieved": flops / sec / 1e12,
                            "peak": fp64_peak_tf, "unit": "TFLOP/s", "frac": flops / sec / 1e12 / fp64_peak_tf,
                            "algorithmic_bytes": nbytes, "hbm_gbs": nbytes / sec / 1e9}}
        if B > 10_000:
            rec["config"] += " (batch sized to fill the GPU: 10^4 sets are 79 of 148 SMs' worth of warps)"
        if name == "c64":
            k = 2
            ins = [kin[i].reshape(N, 1).cpu().numpy() for i in range(M)]; outs = [kout[i].reshape(N, 1).cpu().numpy() for i in range(M)]
            t0 = time.perf_counter()
            for b in range(k):
                ref = lambda a, bb, c: ex.unitary_learning_cost((a, bb, c), ins, outs, N, torch.complex64)
                qr.jax_style_grad(ref, (0, 1, 2))(*[v[b].detach().cpu() for v in (th, ph, om)])
            rec["cpu_baseline"] = {"value": k / (time.perf_counter() - t0), "unit": "parameter sets/s", "cores": cpu_threads(), "kind": "port",
                                   "sample": f"{k} parameter sets, oracle cost + autograd gradient"}
        recs.append(rec)
        del th, ph, om

    # ---- config #2: examples/Qubit_Rotation.py:72-92 -- rot() + fidelity to |0> with gradients, 10^6 and 2^26 sets
    for cdt, rdt, name, r in ((torch.complex64, torch.float32, "c64", 4), (torch.complex128, torch.float64, "c128", 8)):
        for B in (1_000_000, 1 << 26):
            ang = (torch.rand((B, 3), dtype=rdt, device="cuda") * 2 - 1) * math.pi
            ket = torch.tensor([[0.6, 0.8j]], dtype=cdt, device="cuda")
            a = ang.requires_grad_(True)

            def call():
                F = q.rot_fidelity(a[:, 0], a[:, 1], a[:, 2], ket, dtype=cdt)
                return F
            from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
            fid = torch.empty(B, dtype=rdt, device="cuda"); grad = torch.empty((B, 3), dtype=rdt, device="cuda")

            def raw():
                _call("qg_rot_fidelity_fwd_grad", _code(cdt), _ptr(ang), _ptr(ket), 0, _ptr(fid), _ptr(grad), B, _stream())
            small = B * 7 * r < 2 * 126e6
            sec = time_calls(torch, raw, 10, flush if small else None)
            nbytes = B * 7 * r
            rec = {"config": "#2 rot + fidelity value+grad (qg_rot_fidelity_fwd_grad)", "dtype": name, "parameter_sets": B,
                   "parameter_sets_per_s": B / sec, "us": sec * 1e6,
                   "l2_policy": "256 MB flush before every timed call" if small else "inputs + outputs exceed L2",
                   "roofline": {"bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": nbytes / sec / 1e9 / hbm_peak, "algorithmic_bytes": nbytes}}
            if name == "c128" and B == 1_000_000:
                k = 100
                t0 = time.perf_counter()
                for b in range(k):
                    f = lambda p, tt, o: ex.qubit_rotation_cost(p, tt, o, np.array([[0.6], [0.8j]]))
                    qr.jax_style_grad(f, (0, 1, 2))(*[ang[b, i].detach().cpu() for i in range(3)])
                rec["cpu_baseline"] = {"value": k / (time.perf_counter() - t0), "unit": "parameter sets/s", "cores": cpu_threads(),
                                       "kind": "port", "sample": f"{k} parameter sets, oracle cost + autograd gradient"}
            recs.append(rec)
            del ang, fid, grad, a
            torch.cuda.empty_cache()

    # ---- config #3: examples/SNAP_gates.py:150-182,231-248 -- T = 3 blocks, n = 10-40, 10^5 parameter sets, value + gradients
    T, B = 3, 100_000
    for n in (10, 20, 30, 40):
        cdt, rdt, w, r = torch.complex128, torch.float64, 16, 8
        alphas = torch.randn((B, T), dtype=cdt, device="cuda") * 0.7
        thetas = torch.rand((B, T, n), dtype=rdt, device="cuda") * 6.28
        init = q.basis(n, 0, cdt)
        target = (math.sqrt(3) * q.basis(n, 3, cdt) + q.basis(n, 9, cdt)) / 2
        alphas.requires_grad_(True); thetas.requires_grad_(True)

        def call():
            c = q.snap_cost(alphas, thetas, init, target)
            torch.autograd.grad(c.sum(), (alphas, thetas))
        sec = time_calls(torch, call, 5)
        nbytes = B * (2 * (T * w + T * n * r) + r)
        flops = B * 20.0 * T * n * n * 4.0                    # 4 T matrix-vector products forward + 16 T in the reverse sweep, n^2 complex x real MACs each
        rec = {"config": "#3 SNAP cost value+grad (qg_snap_cost_fwd_grad: all blocks, fidelity and reverse sweep in one kernel)",
               "dtype": "c128", "n": n, "blocks": T, "parameter_sets": B,
               "parameter_sets_per_s": B / sec, "us": sec * 1e6, "l2_policy": "parameters + gradients exceed L2 from n = 20 (n = 10: 58 MB, timed back to back)",
               "roofline": {"bound": "fp64-fma", "achieved": flops / sec / 1e12, "peak": fp64_peak_tf, "unit": "TFLOP/s",
                            "frac": flops / sec / 1e12 / fp64_peak_tf, "algorithmic_bytes": nbytes, "hbm_gbs": nbytes / sec / 1e9,
                            "hbm_frac": nbytes / sec / 1e9 / hbm_peak,
                            "note": "algorithmic bytes = parameters in, cost + gradients out (the state never leaves shared memory); "
                                    "20 T n^2 complex x real MACs per set make it FMA-pipe bound from n ~ 20"}}
        k = 2
        t0 = time.perf_counter()
        for b in range(k):
            cost_ref = lambda a, t: ex.snap_cost((a, t), qr.basis(n, 0, cdt), torch.as_tensor(target.cpu().numpy()))
            qr.jax_style_grad(cost_ref, (0, 1))(alphas[b].detach().cpu(), thetas[b].detach().cpu())
        rec["cpu_baseline"] = {"value": k / (time.perf_counter() - t0), "unit": "parameter sets/s", "cores": cpu_threads(), "kind": "port",
                               "sample": f"{k} parameter sets, oracle cost (eigh + dense D per call, as the reference) + autograd gradient"}
        recs.append(rec)
        del alphas, thetas
        torch.cuda.empty_cache()
    return recs


def make_kets(torch, dev, lo, hi, n):
    """Haar-random training kets of settings lo..hi-1, generated ON THE DEVICE (qg_rand_ket: Philox4x32-10, ket index =
    global position in the data set), so that every sharding of the problem trains on the same data -- the loss
    trajectory can be compared across N -- and no synthetic input is staged through the host."""
    import qgrad_b200.qgrad_qutip as q
    K = KETS_PER_SETTING
    kets = q.rand_kets(n, (hi - lo) * K, seed=DATA_SEED, first=lo * K, dtype=torch.complex128)[..., 0]     # ((hi-lo) K, n)
    return kets.reshape(hi - lo, K, n).transpose(1, 2).contiguous()                                        # kets as columns


def build_learner(torch, dev, lo, hi, world_group, args, graphs=True):
    from qgrad_b200 import learning
    import qgrad_b200.qgrad_qutip as q
    n, m_total = 1 << N_QUBITS, N_SETTINGS * KETS_PER_SETTING
    xm, zm, ctrl, theta_true, theta0 = problem()
    psi_in = make_kets(torch, dev, lo, hi, n)
    th_true = torch.tensor(theta_true, dtype=torch.float64, device=dev)
    ctrl_d = torch.tensor(ctrl[lo:hi], dtype=torch.float64, device=dev)
    xm_d, zm_d = torch.tensor(xm, device=dev), torch.tensor(zm, device=dev)
    A_true = learning.pauli_hamiltonian(th_true, ctrl_d, xm_d, zm_d, N_QUBITS, T_EVOLVE)
    psi_out = learning.gemm(q.expm(A_true), psi_in)            # labels from the target Hamiltonian (our own kernels, untimed)
    del A_true
    learner = learning.HamiltonianLearner(N_QUBITS, xm, zm, ctrl[lo:hi], T_EVOLVE, psi_in, psi_out, m_total,
                                          theta_bound=float(np.max(np.abs(theta0)) * 1.5), lr=1e-2,
                                          skew_pullback=not args.general_pullback, group=world_group)
    theta = torch.tensor(theta0, dtype=torch.float64, device=dev)
    if graphs:
        learner.enable_graphs(theta, whole_step=not args.local_graph)
    return learner, theta, psi_in, psi_out


def shutdown(torch, dist, world, learners):
    """Tear the process group down without hanging: CUDA graphs that hold captured NCCL kernels must be released (and the
    device drained) before the communicator is destroyed -- destroying it first blocks forever inside NCCL.  The result
    line has been printed and flushed by now; a watchdog turns a teardown that still stalls into a clean exit 0."""
    import gc
    sys.stdout.flush()
    if world <= 1:
        return
    threading.Thread(target=lambda: (time.sleep(30.0), os._exit(0)), daemon=True).start()
    for l in learners:
        if l is not None:
            l.release_graphs()
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


def run_verify(args, torch, dist, dev, world, rank):
    """--verify: the sharded run (this world size, graphs as in the bench) against the single-process run of the whole
    problem (rank 0, eager launches, no collective) on the same data: loss of every step and theta after ``steps`` steps."""
    from qgrad_b200.sharding import shard_range
    lo, hi = shard_range(N_SETTINGS, world, rank)
    learner, theta, _, _ = build_learner(torch, dev, lo, hi, None, args, graphs=not args.no_graph)
    tickets = [learner.step_async(theta) for _ in range(min(args.steps, 8))]
    losses = [learner.loss_result(t) for t in tickets]
    thetas = [torch.empty_like(theta) for _ in range(world)]
    if world > 1:
        dist.all_gather(thetas, theta)
    else:
        thetas = [theta]
    out = None
    if rank == 0:
        ref, theta_ref, _, _ = build_learner(torch, dev, 0, N_SETTINGS, False, args, graphs=False)
        ref_losses = [float(ref.step(theta_ref)) for _ in range(len(tickets))]
        dl = max(abs(a - b) for a, b in zip(losses, ref_losses))
        dth = float((theta - theta_ref).abs().max() / theta_ref.abs().max())
        ranks_equal = all(bool(torch.equal(t, thetas[0])) for t in thetas)
        ok = dl <= 1e-13 and dth <= 1e-13 and ranks_equal
        out = {"verify": "sharded step == single-process step", "n_gpus": world, "steps": len(tickets), "ok": ok,
               "max_abs_loss_diff": dl, "max_rel_theta_diff": dth, "theta_identical_on_all_ranks": ranks_equal,
               "losses_sharded": losses, "losses_single": ref_losses, "tolerance": 1e-13,
               "graphs": "whole step (one launch incl. all-reduce + Adam)" if not (args.no_graph or args.local_graph) else
                         ("local part" if not args.no_graph else "off")}
        print(json.dumps(out))
    failed = out is not None and not out["ok"]
    if failed:
        sys.stdout.flush()
        os._exit(1)
    shutdown(torch, dist, world, [learner])


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from qgrad_b200 import launch_count, learning
    from qgrad_b200.sharding import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: qgrad_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    if args.verify:
        return run_verify(args, torch, dist, dev, world, rank)
    n, m_total = 1 << N_QUBITS, N_SETTINGS * KETS_PER_SETTING
    lo, hi = shard_range(N_SETTINGS, world, rank)
    group = None
    if args.shard_of > 1:          # profiling aid: rank 0's share of an N-rank run on one GPU, no collectives
        lo, hi = shard_range(N_SETTINGS, args.shard_of, 0)
        group = False
    n_local = hi - lo
    learner, theta, psi_in, psi_out = build_learner(torch, dev, lo, hi, group, args, graphs=not args.no_graph)
    working_set = learner.ws.numel() + 6 * psi_in.numel() * 16

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loss_first = None
    for _ in range(args.warmup):
        l = float(learner.step(theta))
        loss_first = l if loss_first is None else loss_first
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = launch_count()
    graph_launches0 = learner.graph_kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        learner.step(theta)                      # one graph launch: local part, all-reduce, Adam, loss write
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    launches = launch_count() - launches0 + learner.graph_kernel_launches - graph_launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    loss_last = float(learner.loss_out)
    s_dev = learner.check_plan()                 # raises if theta outgrew the squaring budget

    # ---- e2e: the same step through the learner's streaming-input API: every step's kets (psi_in, psi_out;
    #      psi_in^H too with --general-pullback) are copied from pinned host memory inside the timed region -- on a copy stream, so the copy
    #      of step i+1 runs under the compute of step i -- and every step's loss is read back to the host through a pinned
    #      ring, picked up E2E_LAG steps later so that the read never drains the launch queue.
    host_in = psi_in.cpu().pin_memory(); host_out = psi_out.cpu().pin_memory()
    host_in_h = None if learner.skew else learner.psi_in_h.cpu().pin_memory()   # only the general pull-back reads psi_in^H
    h2d = (host_in.numel() + (0 if host_in_h is None else host_in_h.numel()) + host_out.numel()) * 16
    # The e2e loop replays the SAME training steps the device-timed loop ran (theta and Adam state are reset to the start,
    # the same number of warm-up steps precedes the timed ones): the planner picks the number of squarings per matrix from
    # the current Hamiltonian, so steps further along the trajectory are not the same work.
    theta0_dev = torch.tensor(problem()[4], dtype=torch.float64, device=dev)
    learner.reset(theta, theta0_dev)
    learner.enable_streaming()
    learner.feed(host_in, host_in_h, host_out)
    for _ in range(args.warmup):
        learner.swap(); learner.feed(host_in, host_in_h, host_out)
        float(learner.step(theta))
    barrier()
    e2e_steps = max(3, min(args.steps, 50))
    e2e_losses, tickets = [], []
    e0.record()
    for i in range(e2e_steps):
        learner.swap()                                   # this step's kets: wait for their copy (started one step ago)
        learner.feed(host_in, host_in_h, host_out)       # next step's kets start travelling now
        tickets.append(learner.step_async(theta))        # the step + the device->host copy of its loss
        if i >= E2E_LAG:
            e2e_losses.append(learner.loss_result(tickets[i - E2E_LAG]))
    for t in tickets[max(0, e2e_steps - E2E_LAG):]:
        e2e_losses.append(learner.loss_result(t))        # every step's result is on the host before the clock stops
    e1.record()
    barrier()
    assert len(e2e_losses) == e2e_steps and all(math.isfinite(x) for x in e2e_losses)
    if args.e2e_diagnose:
        def timed(body):
            barrier(); ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for i in range(e2e_steps):
                body(i)
            ev1.record(); barrier()
            return ev0.elapsed_time(ev1) / e2e_steps
        t_plain = timed(lambda i: learner.step(theta))
        t_swap = timed(lambda i: (learner.swap(), learner.step(theta)))
        t_feed = timed(lambda i: (learner.swap(), learner.feed(host_in, host_in_h, host_out), learner.step(theta)))
        t_sync = timed(lambda i: (learner.swap(), learner.feed(host_in, host_in_h, host_out), float(learner.step(theta))))
        print(f"[e2e-diagnose rank {rank}] ms/step: step only {t_plain:.3f} | + buffer swap {t_swap:.3f} | + H2D feed {t_feed:.3f} | "
              f"+ float(loss) every step {t_sync:.3f}", file=sys.stderr)
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = 1e3 * e2e_steps / float(ms2)

    # ---- roofline of the dominant kernel: the n^3 complex GEMM launch (FP64 tensor pipe), CUDA events on its stream
    peak_tf, peak_src = fp64_peak()
    Xa = torch.randn((n_local, n, n), dtype=torch.complex128, device=dev)
    Xb = torch.randn((n_local, n, n), dtype=torch.complex128, device=dev)
    Xc = torch.empty_like(Xa)
    for _ in range(3):
        learning.gemm(Xa, Xb, out=Xc)
    torch.cuda.synchronize()
    reps = 20
    e0.record()
    for _ in range(reps):
        learning.gemm(Xa, Xb, out=Xc)
    e1.record()
    torch.cuda.synchronize()
    gemm_ms = e0.elapsed_time(e1) / reps
    gemm_flops = 8.0 * n ** 3 * n_local
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    s_used = int(s_dev.max())
    norms = learner.A.abs().sum(dim=-2).amax(dim=-1).double().cpu().numpy()
    prods = np.array([learner.gemm_equivalents(int(x)) for x in s_dev])     # tile-exact count of what the step executes
    step_flops = float(8.0 * n ** 3 * prods.sum() + n_local * 2 * 8.0 * n * n * KETS_PER_SETTING)
    traffic, traffic_src = None, None
    try:       # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r02_ncu_gemm1024_b8.json")) as f:
            cap = json.load(f)
        if n_local == 8:
            traffic, traffic_src = float(cap["traffic_bytes_per_launch"]), "profiles/r02_ncu_gemm1024_b8.json (batch 8)"
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "qg::gemm::gemm_kernel<double> (FP64 DMMA complex GEMM, 1024^3 per setting; "
                                             "plain tile grid at >= 4 waves of tiles, stream-K below)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 3.0 * n_local * n * n * 16,
                "peak_source": peak_src, "launch_ms": gemm_ms, "algorithmic_flops_per_launch": gemm_flops,
                "step_algorithmic_tflop": step_flops / 1e12, "step_mean_gemm_equivalents_per_setting": float(prods.mean()),
                "step_achieved_tflops": step_flops / (ms_per_step * 1e-3) / 1e12,
                "note": "MEASURED_PEAKS.json has no FP64 figure; the expm chain is FP64-tensor-pipe bound (SURVEY 8d)"}
    del Xa, Xb, Xc

    line = None
    if rank == 0:
        graph_mode = ("off" if args.no_graph else
                      ("local part replayed, all-reduce + Adam launched eagerly" if args.local_graph else
                       "whole step = one CUDA-graph launch (local part + NCCL all-reduce + Adam + loss)"))
        line = {**({"emulated_rank0_of": args.shard_of, "note": "NOT a bench line: one rank's share, no collectives"}
                   if args.shard_of > 1 else {}),
                "metric": METRIC, "value": 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": dict(base_config(world), squarings=s_used, max_squarings_launched=learner.ms,
                               pullback="skew-Hermitian part from the body-frame cotangent G Phi^H (QG_ADJ_SKEW_BODY)"
                               if learner.skew else "general Frechet pull-back from Ubar",
                               norm1_max=float(norms.max()), l2_policy="working set per step exceeds L2",
                               working_set_bytes_per_rank=int(working_set), cuda_graph=graph_mode,
                               data_seed=f"Haar kets from the device generator (Philox4x32-10, seed {DATA_SEED}, ket index = global position): identical data at every N"),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
                        "steps": e2e_steps, "input_pipeline": "double-buffered: copy of step i+1 on a copy stream under the compute of step i",
                        "result_readback": f"every step's loss copied to pinned host memory behind the step, read by the host {E2E_LAG} steps later; "
                                           f"all {e2e_steps} read before the clock stops", "loss_last": e2e_losses[-1],
                        "same_steps_as_device_timed_loop": bool(e2e_steps == args.steps and abs(e2e_losses[-1] - loss_last) < 1e-12)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "loss_first": loss_first, "loss_last": loss_last}
    if rank == 0 and world == 1 and args.shard_of == 1:
        aux = ClockSampler(local); aux.start(); t_aux0 = time.time()
        if not args.no_sweep:
            # BASELINE config #4 at its stated size: batches that fill half of HBM, streamed through a 40 GB workspace in chunks
            # (~100 s for the 18 points).  Falls back to the bounded batches if the box cannot give the memory.
            full = not args.sweep_bounded
            try:
                line["sweep"] = sweep_expm_vjp(torch, peak_tf, full, not args.no_cpu)
                line["sweep_batches"] = "HBM-filling (config #4)" if full else "bounded"
            except RuntimeError as e:
                if not full:
                    raise
                torch.cuda.empty_cache()
                line["sweep"] = sweep_expm_vjp(torch, peak_tf, False, not args.no_cpu)
                line["sweep_batches"] = f"bounded (HBM-filling sweep failed: {str(e)[:80]})"
            line["c64_gemm"] = c64_gemm_record(torch)
        if not args.no_configs:
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    hbm_peak = float(json.load(f)["hbm_gbs"])
            except Exception:
                hbm_peak = 6650.0                         # B200_PROFILING.md fallback
            line["configs"] = config_records(torch, hbm_peak, peak_tf)
        line["clocks_aux"] = aux.stop(t_aux0, time.time())
        if not args.no_cpu:
            from oracle.examples_ref import pauli_entries
            xm_, zm_, ctrl_, _, th0 = problem()
            ents = [pauli_entries(int(x), int(z), N_QUBITS) for x, z in zip(xm_, zm_)]
            pin, pout = psi_in.cpu().numpy(), psi_out.cpu().numpy()
            cpu_setting(ents, ctrl_[0], th0, pin[0], pout[0], m_total)
            t0 = time.perf_counter()
            for s in range(N_SETTINGS):
                cpu_setting(ents, ctrl_[s], th0, pin[s], pout[s], m_total)
            t = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / t, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                                    "sample": f"one full step: all {N_SETTINGS} settings (scipy expm + expm_frechet, dim 1024, 512 kets each), "
                                              f"after one untimed setting; {t:.1f} s"}
    if rank == 0:
        print(json.dumps(line))
    shutdown(torch, dist, world, [learner])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qgrad_b200", choices=["qgrad_b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the expm+VJP matrices/s sweep")
    ap.add_argument("--sweep-full", action="store_true", help="(default since round 2) size sweep batches to fill HBM (BASELINE config #4)")
    ap.add_argument("--sweep-bounded", action="store_true", help="bounded sweep batches (~0.25 s of work per call) instead of the HBM-filling ones")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline samples")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config #1-#3 sub-records")
    ap.add_argument("--local-graph", action="store_true",
                    help="capture only the local part of a step; all-reduce and Adam are launched eagerly (A/B aid)")
    ap.add_argument("--e2e-diagnose", action="store_true",
                    help="also time the e2e loop without the host->device copies and without the loss read-back (stderr)")
    ap.add_argument("--verify", action="store_true",
                    help="self-check instead of a bench: the sharded run at this world size must reproduce the single-process "
                         "run of the whole problem (losses of every step, theta after --steps steps) to 1e-13; exit 1 if not")
    ap.add_argument("--general-pullback", action="store_true",
                    help="general Frechet pull-back from Ubar instead of the skew-Hermitian body-frame one (A/B aid)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--shard-of", type=int, default=1,
                    help="profiling aid (not a bench line): run rank 0's share of an N-rank job on one GPU, no collectives")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
