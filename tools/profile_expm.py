"""Runs the fused expm+VJP of one (dim, dtype) sweep point a few times: the target of ncu captures.
usage: python tools/profile_expm.py DIM [BATCH] [c128|c64] [REPS]"""
import math, sys, torch
sys.path.insert(0, ".")
from qgrad_b200.qgrad_qutip import expm_fwd_vjp
n = int(sys.argv[1])
esz = 16
dt = torch.complex128 if (len(sys.argv) <= 3 or sys.argv[3] == "c128") else torch.complex64
if dt == torch.complex64: esz = 8
batch = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else max(2, int(min(600e6 / (4 * esz * n * n), 4e12 / (8.0 * n ** 3 * 34))))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
g = torch.Generator(device="cuda").manual_seed(n)
X = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
A = (-0.5j / math.sqrt(n)) * (X + X.mH)
G = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
ms = 5
expm_fwd_vjp(A, G, max_squarings=ms)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    expm_fwd_vjp(A, G, max_squarings=ms)
e1.record()
torch.cuda.synchronize()
print(f"n={n} batch={batch} {dt} {batch * reps / (e0.elapsed_time(e1) / 1e3):.4g} mat/s")
