"""small complex64 cases through the tensor-core GEMM for compute-sanitizer: python tools/tc_sanitize_case.py"""
import math, sys, torch
sys.path.insert(0, ".")
from qgrad_b200 import learning
import qgrad_b200.qgrad_qutip as q
for (b, m, n, k) in ((1, 256, 384, 128), (3, 192, 320, 192)):
    A = torch.randn((b, m, k), dtype=torch.complex64, device="cuda"); B = torch.randn((b, k, n), dtype=torch.complex64, device="cuda")
    C = learning.gemm(A, B)
    ref = A.to(torch.complex128) @ B.to(torch.complex128)
    print(b, m, n, k, float((C.to(torch.complex128) - ref).abs().max() / ref.abs().max()))
X = torch.randn((2, 200, 200), dtype=torch.complex64, device="cuda")
A = (-0.5j / math.sqrt(200)) * (X + X.mH)            # skew-Hermitian: structure-hint tiles (mirror writes)
Y = q.expm(A)
print("unitarity", float((Y @ Y.mH - torch.eye(200, device="cuda")).abs().max()))
