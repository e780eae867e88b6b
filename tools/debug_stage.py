import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, ".")
import qgrad_b200.qgrad_qutip as q
from qgrad_b200 import _lib
from oracle import expm_ref as E
lib = _lib.load()
lib.qg_debug_set_stage.argtypes = [ctypes.c_int]
rng = np.random.default_rng(0)
n = 16
A = (-1j * E.rand_hermitian(rng, 3, n, True)).astype(np.complex64)
Ad = torch.as_tensor(A).cuda()
# reference intermediates in complex128
norm = np.abs(A).sum(axis=1).max(axis=-1)
res = {}
for b in range(3):
    a = A[b].astype(np.complex128)
    m, s = E.pade_plan(float(np.abs(a).sum(0).max()), True, False)
    m = 5 if m >= 5 else 3
    bound = 1.3
    s = max(0, int(np.ceil(np.log2(np.abs(a).sum(0).max() / bound)))) if np.abs(a).sum(0).max() > bound else 0
    a = a * 2.0 ** -s
    bco = [1.0, 0.5, 1/9, 1/72, 1/1008, 1/30240]
    I = np.eye(n)
    A2 = a @ a; A4 = A2 @ A2
    W = bco[5]*A4 + bco[3]*A2 + bco[1]*I
    V = bco[4]*A4 + bco[2]*A2 + bco[0]*I
    U = a @ W
    Q = V - U; P = V + U; Qi = np.linalg.inv(Q); R = Qi @ P
    res[b] = [a, A2, A4, W, Q, P, Qi, R]
names = ["As", "A2", "A4", "W", "Q", "P", "Qinv", "R0"]
for stage in range(8):
    lib.qg_debug_set_stage(stage)
    Y1 = q.expm(Ad).cpu().numpy(); Y2 = q.expm(Ad).cpu().numpy()
    errs = ["%.1e" % E.rel_fro(Y1[b], res[b][stage]) for b in range(3)]
    print(stage, names[stage], errs, "det", bool((Y1 == Y2).all()), flush=True)
lib.qg_debug_set_stage(-1)
Y = q.expm(Ad).cpu().numpy()
print("final", ["%.1e" % E.rel_fro(Y[b], E.expm_scipy(A[b].astype(np.complex128))) for b in range(3)])
