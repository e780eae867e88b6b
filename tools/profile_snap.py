"""one fused SNAP cost + gradient call for ncu: python tools/profile_snap.py N [BATCH]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import qgrad_b200.qgrad_qutip as q
n = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
T = 3
rng = np.random.default_rng(0)
alphas = torch.as_tensor(rng.standard_normal((B, T)) + 1j * rng.standard_normal((B, T))).cuda().requires_grad_(True)
thetas = torch.as_tensor(rng.uniform(0, 6.28, (B, T, n))).cuda().requires_grad_(True)
init = q.basis(n, 0).to(torch.complex128); target = q.basis(n, 1).to(torch.complex128)
for _ in range(3):
    out = q.snap_cost(alphas, thetas, init, target)
torch.cuda.synchronize()
print("ok", float(out.sum()))
