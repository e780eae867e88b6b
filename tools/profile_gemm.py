"""Runs the dominant kernel alone (the 1024^3 complex128 GEMM, batch 8, as in the learning step)
a few times: the target of the `ncu --set full` capture."""
import sys, torch
sys.path.insert(0, ".")
from qgrad_b200 import learning
b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 1024
A = torch.randn((b, n, n), dtype=torch.complex128, device="cuda")
B = torch.randn((b, n, n), dtype=torch.complex128, device="cuda")
C = torch.empty_like(A)
for _ in range(6):
    learning.gemm(A, B, out=C)
torch.cuda.synchronize()
print("ok", float((C - A @ B).abs().max()))
