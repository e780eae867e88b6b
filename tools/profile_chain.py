"""one fused expm + pull-back call for a launch list: python tools/profile_chain.py DIM BATCH [c128|c64]"""
import math, sys, torch
sys.path.insert(0, ".")
from qgrad_b200 import _lib
from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
n, batch = int(sys.argv[1]), int(sys.argv[2])
dt = torch.complex64 if (len(sys.argv) > 3 and sys.argv[3] == "c64") else torch.complex128
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(n)
X = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
A = (-0.5j / math.sqrt(n)) * (X + X.mH); del X
G = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
Y, Ab = torch.empty_like(A), torch.empty_like(A)
ws = torch.empty(lib.qg_expm_workspace_bytes(_code(dt), n, batch, 1, 1), dtype=torch.uint8, device="cuda")
for _ in range(3):
    _call("qg_expm_fwd_vjp", _code(dt), _ptr(A), _ptr(G), _ptr(Y), _ptr(Ab), batch, n, 0, _ptr(ws), ws.numel(), 1, _stream())
torch.cuda.synchronize()
print("ok")
