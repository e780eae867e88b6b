"""fused expm + pull-back at one small dim: python tools/time_small.py DIM [c128|c64]   (QGRAD_B200_LIB selects a library variant)"""
import math, sys, torch
sys.path.insert(0, ".")
from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
n = int(sys.argv[1]); dt = torch.complex64 if (len(sys.argv) > 2 and sys.argv[2] == "c64") else torch.complex128
esz = 8 if dt == torch.complex64 else 16
batch = int(2e9 / (4 * esz * n * n))
g = torch.Generator(device="cuda").manual_seed(n)
X = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
A = (-0.5j / math.sqrt(n)) * (X + X.mH); del X
G = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
Y, Ab = torch.empty_like(A), torch.empty_like(A)
call = lambda: _call("qg_expm_fwd_vjp", _code(dt), _ptr(A), _ptr(G), _ptr(Y), _ptr(Ab), batch, n, 0, None, 0, 3, _stream())
for _ in range(2): call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): call()
e1.record(); torch.cuda.synchronize()
print(f"n={n} {dt} batch={batch}: {batch / (e0.elapsed_time(e1) / 5e3):.4e} matrices/s")
