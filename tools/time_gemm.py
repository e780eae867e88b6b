"""Times the complex GEMM kernel alone: python tools/time_gemm.py BATCH M N K [c128|c64]  (QG_STREAMK=G forces a stream-K grid, 0 disables)"""
import sys, torch
sys.path.insert(0, ".")
from qgrad_b200 import learning
b, m, n, k = (int(x) for x in sys.argv[1:5])
dt = torch.complex64 if (len(sys.argv) > 5 and sys.argv[5] == "c64") else torch.complex128
A = torch.randn((b, m, k), dtype=dt, device="cuda"); B = torch.randn((b, k, n), dtype=dt, device="cuda")
C = torch.empty((b, m, n), dtype=dt, device="cuda")
for _ in range(3): learning.gemm(A, B, out=C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps): learning.gemm(A, B, out=C)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
err = float((C - A @ B).abs().max() / (A @ B).abs().max())
print(f"b={b} {m}x{n}x{k} {dt}: {us:.1f} us  {8.0*b*m*n*k/us/1e6:.2f} TFLOP/s  relerr {err:.2e}")
