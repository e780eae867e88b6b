"""Times the complex64 tensor-core GEMM alone: python tools/tc_gemm_time.py BATCH M N K [reps]"""
import sys, torch
sys.path.insert(0, ".")
from qgrad_b200 import learning
b, m, n, k = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
A = torch.randn((b, m, k), dtype=torch.complex64, device="cuda"); B = torch.randn((b, k, n), dtype=torch.complex64, device="cuda")
C = torch.empty((b, m, n), dtype=torch.complex64, device="cuda")
for _ in range(3): learning.gemm(A, B, out=C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): learning.gemm(A, B, out=C)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"b={b} {m}x{n}x{k} c64: {us:.1f} us  {8.0*b*m*n*k/us/1e6:.2f} TFLOP/s")
