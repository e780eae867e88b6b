"""complex64 GEMM: the tcgen05 split-TF32 kernel (gemm_tc.cu) against the FP32 FMA kernel and a float64 product.

    python tools/tc_gemm_check.py [quick]

Prints, per shape, the relative Frobenius error of both paths against the complex128 product and their time per launch;
then the error after a chain of 16 dependent squarings of a unitary matrix (the expm chain's regime)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from qgrad_b200 import _lib, learning  # noqa: E402

lib = _lib.load()
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def rel(x, ref):
    return float(torch.linalg.norm((x.to(torch.complex128) - ref).reshape(-1)) / torch.linalg.norm(ref.reshape(-1)))


def coherent(x, ref):
    """signed relative component of the error ALONG the result: x ~ (1 + delta) ref -- what a chain of squarings doubles per step"""
    d = (x.to(torch.complex128) - ref).reshape(-1)
    r = ref.reshape(-1)
    return float((torch.vdot(r, d) / torch.vdot(r, r)).real)


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


out = []
shapes = [(1, 128, 128, 64), (2, 128, 128, 1024), (1, 256, 384, 128), (3, 192, 320, 192), (8, 1024, 1024, 1024)]
if not quick:
    shapes += [(1, 1024, 1024, 1024), (59, 1024, 1024, 1024), (1, 4096, 4096, 4096), (8, 1024, 512, 1024), (16, 512, 512, 512)]
for (b, m, n, k) in shapes:
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k + b)
    A = torch.randn((b, m, k), dtype=torch.complex64, device="cuda", generator=g)
    B = torch.randn((b, k, n), dtype=torch.complex64, device="cuda", generator=g)
    ref = A.to(torch.complex128) @ B.to(torch.complex128)
    rec = {"batch": b, "m": m, "n": n, "k": k}
    for name, on in (("tc", 1), ("fma", 0)):
        lib.qg_set_c64_tensor(on)
        C = torch.full((b, m, n), float("nan"), dtype=torch.complex64, device="cuda")
        learning.gemm(A, B, out=C)
        torch.cuda.synchronize()
        rec[name + "_err"] = rel(C, ref)
        rec[name + "_coh"] = coherent(C, ref)
        us = timed(lambda: learning.gemm(A, B, out=C), 5 if b * m * n * k > 2 ** 33 else 20)
        rec[name + "_us"] = us
        rec[name + "_tflops"] = 8.0 * b * m * n * k / us / 1e6
    rec["torch_err"] = rel(A @ B, ref)
    us = timed(lambda: torch.matmul(A, B), 5 if b * m * n * k > 2 ** 33 else 20)
    rec["torch_us"] = us
    rec["torch_tflops"] = 8.0 * b * m * n * k / us / 1e6
    print(json.dumps(rec), flush=True)
    out.append(rec)

# chain of dependent squarings of a unitary (norm-preserving: errors add, they do not shrink)
n = 1024
g = torch.Generator(device="cuda").manual_seed(5)
H = torch.randn((n, n), dtype=torch.complex128, device="cuda", generator=g)
H = (H + H.mH) / 2
w, V = torch.linalg.eigh(H)
U0 = (V * torch.exp(1j * w / 50)) @ V.mH
for name, on in (("tc", 1), ("fma", 0)):
    lib.qg_set_c64_tensor(on)
    R = U0.to(torch.complex64).reshape(1, n, n).contiguous()
    R64 = R.to(torch.complex128)
    errs = []
    for it in range(16):
        Rn = torch.empty_like(R)
        learning.gemm(R, R, out=Rn)
        R = Rn
        R64 = R64 @ R64
        errs.append(rel(R, R64))
    rec = {"chain": name, "n": n, "err_after": {4: errs[3], 8: errs[7], 16: errs[15]}}
    print(json.dumps(rec), flush=True)
    out.append(rec)
lib.qg_set_c64_tensor(1)
json.dump(out, open("gpurun_out/tc_gemm_check.json", "w"), indent=1)
