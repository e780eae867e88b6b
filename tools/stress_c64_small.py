"""Stress test for the complex64 shared-memory expm kernel (9 <= n <= 32): large batches, several runs; every run must
be bit-identical to the first, and every matrix within the complex64 tolerance of the complex128 kernel's result."""
import math, sys, torch
sys.path.insert(0, ".")
from qgrad_b200.qgrad_qutip import expm_fwd_vjp
import qgrad_b200.qgrad_qutip as q
ok = True
for n in (9, 12, 16, 17, 24, 32):
    for kind in ("scaled", "stiff", "general"):
        batch = 20000
        g = torch.Generator(device="cuda").manual_seed(n)
        X = torch.randn((batch, n, n), dtype=torch.complex128, device="cuda", generator=g)
        if kind == "general":
            A = X * (1.5 / math.sqrt(n))
        else:
            A = (-0.5j / (math.sqrt(n) if kind == "scaled" else 2.0)) * (X + X.mH)
        G = torch.randn((batch, n, n), dtype=torch.complex128, device="cuda", generator=g)
        Yd, Ad = expm_fwd_vjp(A, G)
        A32, G32 = A.to(torch.complex64), G.to(torch.complex64)
        first = None
        for rep in range(4):
            Y, Ab = expm_fwd_vjp(A32, G32)
            Yf = q.expm(A32)
            if first is None:
                first = (Y.clone(), Ab.clone(), Yf.clone())
            else:
                same = torch.equal(Y, first[0]) and torch.equal(Ab, first[1]) and torch.equal(Yf, first[2])
                if not same:
                    ok = False; print("NOT REPRODUCIBLE", n, kind, rep)
        ey = (torch.linalg.norm(Y.to(torch.complex128) - Yd, dim=(1, 2)) / torch.linalg.norm(Yd, dim=(1, 2))).max()
        ea = (torch.linalg.norm(Ab.to(torch.complex128) - Ad, dim=(1, 2)) / torch.linalg.norm(Ad, dim=(1, 2))).max()
        ef = (torch.linalg.norm(Yf.to(torch.complex128) - Yd, dim=(1, 2)) / torch.linalg.norm(Yd, dim=(1, 2))).max()
        tol = 1e-5 if kind != "stiff" else 2e-4
        flag = "" if max(float(ey), float(ea), float(ef)) < tol else "  <-- FAIL"
        if flag: ok = False
        print(f"n={n:2d} {kind:8s} max rel err: Y {float(ey):.2e}  Abar {float(ea):.2e}  Y(fwd only) {float(ef):.2e}{flag}")
print("STRESS", "OK" if ok else "FAILED")
