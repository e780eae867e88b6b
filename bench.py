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
def cpu_step_sample(xm, zm, ctrl_row, theta, psi_in, psi_out, m_total):
    """One setting of one step on the host cores with the oracle (SciPy Pade expm + expm_frechet,
    NumPy GEMMs): returns seconds.  The full step is N_SETTINGS such samples."""
    import scipy.linalg
    n = 1 << N_QUBITS
    t0 = time.perf_counter()
    H = np.zeros((n, n), dtype=np.complex128)
    cols = np.arange(n)
    for p in range(len(xm)):
        ph = (1j) ** bin(int(xm[p]) & int(zm[p])).count("1") * (-1.0) ** np.array([bin(c & int(zm[p])).count("1") for c in cols])
        H[cols ^ int(xm[p]), cols] += ctrl_row[p] * theta[p] * ph
    A = -1j * T_EVOLVE * H
    U = scipy.linalg.expm(A)
    phi = U @ psi_in
    o = np.sum(np.conj(psi_out) * phi, axis=0)
    G = (-np.sign(o.real) / m_total)[None, :] * psi_out
    Ubar = G @ psi_in.conj().T
    Abar = scipy.linalg.expm_frechet(A.conj().T, Ubar, compute_expm=False)
    _ = np.real(np.sum(np.conj(Abar) * A))       # stands in for the P projections (O(n) each)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scipy.linalg  # noqa: F401  (load its BLAS before the limits below)
    from threadpoolctl import threadpool_info, threadpool_limits
    threadpool_limits(limits=os.cpu_count())
    xm, zm, ctrl, theta_true, theta0 = problem()
    rng = np.random.default_rng(1)
    n = 1 << N_QUBITS
    psi = rng.standard_normal((n, KETS_PER_SETTING)) + 1j * rng.standard_normal((n, KETS_PER_SETTING))
    psi /= np.linalg.norm(psi, axis=0, keepdims=True)
    out = rng.standard_normal((n, KETS_PER_SETTING)) + 1j * rng.standard_normal((n, KETS_PER_SETTING))
    out /= np.linalg.norm(out, axis=0, keepdims=True)
    m_total = N_SETTINGS * KETS_PER_SETTING
    for _ in range(args.warmup):
        cpu_step_sample(xm, zm, ctrl[0], theta0, psi, out, m_total)
    ts = [cpu_step_sample(xm, zm, ctrl[0], theta0, psi, out, m_total) for _ in range(args.steps)]
    sec_per_step = float(np.mean(ts)) * N_SETTINGS
    threads = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    value = 1.0 / sec_per_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic", "config": base_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"1 of {N_SETTINGS} settings per step (scipy.linalg.expm + expm_frechet at dim 1024, "
                                       f"512 kets), time x{N_SETTINGS}; host has {os.cpu_count()} cpus"},
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


def sweep_expm_vjp(torch, peak_tf, full):
    """expm+VJP matrices/s (c128, plus c64) vs dim -- the first half of BASELINE.json's metric."""
    from qgrad_b200.qgrad_qutip import expm_fwd_vjp
    from qgrad_b200 import _lib
    rows = []
    for dt, name, esz in ((torch.complex128, "c128", 16), (torch.complex64, "c64", 8)):
        for n in (4, 8, 16, 32, 64, 128, 256, 512, 1024):
            if full:
                batch = int(180e9 / 2 / (4 * esz * n * n))
            else:   # bounded: inputs+outputs >= 4x L2 for the small dims, ~1 s of work for the large ones
                batch = max(2, int(min(600e6 / (4 * esz * n * n), 4e12 / (8.0 * n ** 3 * 34))))
            if n > 32:
                ws1 = _lib.load().qg_expm_workspace_bytes(0 if name == "c64" else 1, n, 1, 1, 8)
                batch = max(1, min(batch, int(60e9 // ws1)))
            g = torch.Generator(device="cuda").manual_seed(n)
            X = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
            A = (-0.5j / math.sqrt(n)) * (X + X.mH)
            G = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
            del X
            norms = A.abs().sum(dim=-2).amax(dim=-1).double().cpu().numpy()
            s1 = int(max(0, math.ceil(math.log2(max(float(norms.max()) / (0.783 if name == "c128" else 1.3), 1e-30)))))
            if n <= 32:
                m_used, s_used = plan_squarings(torch, A, name)
            else:                                                # read back what the device chose
                nb = min(batch, 4)
                s_used = expm_fwd_vjp(A[:nb], G[:nb], max_squarings=max(1, s1), return_plan=True)[2].cpu().numpy()
                s_used = np.full(batch, float(s_used.max()))
                m_used = np.full(batch, 7.0 if name == "c128" else 5.0)
            prods = gemm_equivalents(m_used, s_used)             # per matrix, incl. the inverse
            s = int(s_used.max())
            ms = max(1, s)
            expm_fwd_vjp(A, G, max_squarings=ms)                 # full-size warm-up: the workspace allocation happens here
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                expm_fwd_vjp(A, G, max_squarings=ms)
            e1.record()
            torch.cuda.synchronize()
            sec = e0.elapsed_time(e1) / 1e3 / reps
            flops = 8.0 * n ** 3 * float(prods.sum())            # algorithmic: unpadded n
            row = {"dim": n, "dtype": name, "batch": batch, "max_squarings": s, "squarings_1norm_rule": s1,
                   "mean_gemm_equivalents": float(prods.mean()), "matrices_per_s": batch / sec,
                   "tflops": flops / sec / 1e12, "hbm_gbs": 4.0 * esz * n * n * batch / sec / 1e9}
            if name == "c128":
                row["frac_fp64_peak"] = row["tflops"] / peak_tf
            rows.append(row)
            del A, G
            torch.cuda.empty_cache()
    return rows


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
    n, m_total = 1 << N_QUBITS, N_SETTINGS * KETS_PER_SETTING
    xm, zm, ctrl, theta_true, theta0 = problem()
    lo, hi = shard_range(N_SETTINGS, world, rank)
    if args.shard_of > 1:          # profiling aid: rank 0's share of an N-rank run on one GPU, no collectives
        lo, hi = shard_range(N_SETTINGS, args.shard_of, 0)
    n_local = hi - lo
    theta_bound = float(np.max(np.abs(theta0)) * 1.5)

    # synthetic data: Haar-random kets, labels from the target Hamiltonian (generated with our own kernels, untimed)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    psi_in = torch.randn((n_local, n, KETS_PER_SETTING), dtype=torch.complex128, device=dev, generator=g)
    psi_in /= torch.linalg.norm(psi_in, dim=1, keepdim=True)
    ctrl_loc = ctrl[lo:hi]
    th_true = torch.tensor(theta_true, dtype=torch.float64, device=dev)
    ctrl_d = torch.tensor(ctrl_loc, dtype=torch.float64, device=dev)
    xm_d, zm_d = torch.tensor(xm, device=dev), torch.tensor(zm, device=dev)
    import qgrad_b200.qgrad_qutip as q
    A_true = learning.pauli_hamiltonian(th_true, ctrl_d, xm_d, zm_d, N_QUBITS, T_EVOLVE)
    psi_out = learning.gemm(q.expm(A_true), psi_in)
    del A_true
    learner = learning.HamiltonianLearner(N_QUBITS, xm, zm, ctrl_loc, T_EVOLVE, psi_in, psi_out, m_total,
                                          theta_bound=theta_bound, lr=1e-2, skew_pullback=not args.general_pullback)
    theta = torch.tensor(theta0, dtype=torch.float64, device=dev)
    if not args.no_graph:
        learner.enable_graphs(theta)
    working_set = learner.ws.numel() + 6 * psi_in.numel() * 16

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        losses.append(learner.step(theta))
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
        losses.append(learner.step(theta))
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    launches = launch_count() - launches0 + learner.graph_kernel_launches - graph_launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    loss_first, loss_last = float(losses[0]), float(losses[-1])

    # ---- e2e: the same step through the learner's streaming-input API: every step's kets (psi_in, psi_out;
    #      psi_in^H too with --general-pullback) are copied from pinned host memory inside the timed region -- on a copy stream, so the copy
    #      of step i+1 runs under the compute of step i -- and every step's loss is read back to the host.
    host_in = psi_in.cpu().pin_memory(); host_out = psi_out.cpu().pin_memory()
    host_in_h = None if learner.skew else learner.psi_in_h.cpu().pin_memory()   # only the general pull-back reads psi_in^H
    h2d = (host_in.numel() + (0 if host_in_h is None else host_in_h.numel()) + host_out.numel()) * 16
    learner.enable_streaming()
    learner.feed(host_in, host_in_h, host_out)
    for _ in range(2):
        learner.swap(); learner.feed(host_in, host_in_h, host_out)
        float(learner.step(theta))
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for i in range(e2e_steps):
        learner.swap()                                   # this step's kets: wait for their copy (started one step ago)
        learner.feed(host_in, host_in_h, host_out)       # next step's kets start travelling now
        loss_host = float(learner.step(theta))           # device->host read of the step's result
    e1.record()
    barrier()
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
    from qgrad_b200.qgrad_qutip import expm_plan
    s_dev = expm_plan(learner.ws, n_local).cpu().numpy()
    s_used = int(s_dev.max())
    norms = learner.A.abs().sum(dim=-2).amax(dim=-1).double().cpu().numpy()
    prods = np.array([learner.gemm_equivalents(int(x)) for x in s_dev])     # tile-exact count of what the step executes
    step_flops = float(8.0 * n ** 3 * prods.sum() + n_local * 2 * 8.0 * n * n * KETS_PER_SETTING)
    traffic, traffic_src = None, None
    try:       # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r01_ncu_gemm1024_b8.json")) as f:
            cap = json.load(f)
        if n_local == 8:
            traffic, traffic_src = float(cap["traffic_bytes_per_launch"]), "profiles/r01_ncu_gemm1024_b8.json (batch 8)"
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
        line = {**({"emulated_rank0_of": args.shard_of, "note": "NOT a bench line: one rank's share, no collectives"}
                   if args.shard_of > 1 else {}),
                "metric": METRIC, "value": 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": dict(base_config(world), squarings=s_used, max_squarings_launched=learner.ms,
                               pullback="skew-Hermitian part from the body-frame cotangent G Phi^H (QG_ADJ_SKEW_BODY)"
                               if learner.skew else "general Frechet pull-back from Ubar",
                               norm1_max=float(norms.max()), l2_policy="working set per step exceeds L2",
                               working_set_bytes_per_rank=int(working_set)),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
                        "steps": e2e_steps, "input_pipeline": "double-buffered: copy of step i+1 on a copy stream under the compute of step i"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "loss_first": loss_first, "loss_last": loss_last}
    if rank == 0 and world == 1 and not args.no_sweep:
        line["sweep"] = sweep_expm_vjp(torch, peak_tf, args.sweep_full)
    if rank == 0 and world == 1 and not args.no_cpu:
        xm_, zm_, ctrl_, _, th0 = problem()
        t = cpu_step_sample(xm_, zm_, ctrl_[0], th0, psi_in[0].cpu().numpy(), psi_out[0].cpu().numpy(), m_total)
        t = min(t, cpu_step_sample(xm_, zm_, ctrl_[0], th0, psi_in[0].cpu().numpy(), psi_out[0].cpu().numpy(), m_total))
        try:
            from threadpoolctl import threadpool_info
            threads = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
        except Exception:
            threads = os.cpu_count()
        line["cpu_baseline"] = {"value": 1.0 / (t * N_SETTINGS), "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"1 of {N_SETTINGS} settings of one step (scipy expm + expm_frechet, dim 1024, 512 kets), "
                                          f"best of 2, time x{N_SETTINGS}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qgrad_b200", choices=["qgrad_b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the expm+VJP matrices/s sweep")
    ap.add_argument("--sweep-full", action="store_true", help="size sweep batches to fill HBM (BASELINE config #4)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
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
