import sys, numpy as np, torch
sys.path.insert(0, ".")
import qgrad_b200.qgrad_qutip as q
from qgrad_b200.qgrad_qutip import expm_fwd_vjp
from oracle import expm_ref as E
rng = np.random.default_rng(0)
quick = len(sys.argv) > 1
for dt, npdt in (((torch.complex64, np.complex64),) if quick else ((torch.complex64, np.complex64), (torch.complex128, np.complex128))):
    for n in ((10, 16, 32) if quick else (8, 9, 10, 12, 16, 17, 24, 32)):
        A = (-1j * E.rand_hermitian(rng, 5, n, True)).astype(npdt)
        G = E.rand_cotangent(rng, 5, n).astype(npdt)
        Yr = E.expm_scipy(A.astype(np.complex128)); Lr = E.expm_vjp(A.astype(np.complex128), G.astype(np.complex128))
        Ad, Gd = torch.as_tensor(A).cuda(), torch.as_tensor(G).cuda()
        Y1 = q.expm(Ad).cpu().numpy(); Y2 = q.expm(Ad).cpu().numpy()
        Yv, Lv = expm_fwd_vjp(Ad, Gd); Yv = Yv.cpu().numpy(); Lv = Lv.cpu().numpy()
        per = lambda X, R: ["%.1e" % E.rel_fro(X[b], R[b]) for b in range(len(X))]
        print(dt, n, "fwd", per(Y1, Yr), "det", bool((Y1 == Y2).all()), "| vjpY", per(Yv, Yr), "L", per(Lv, Lr), flush=True)
