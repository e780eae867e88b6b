"""A few launches of qg_overlap_loss at the learning step's size (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
cdt = torch.complex64 if (len(sys.argv) > 1 and sys.argv[1] == "c64") else torch.complex128
rdt = torch.float32 if cdt == torch.complex64 else torch.float64
S, n, m = 8, 1024, 512
phi = torch.randn((S, n, m), dtype=cdt, device="cuda"); out = torch.randn((S, n, m), dtype=cdt, device="cuda")
G = torch.empty_like(phi); loss = torch.zeros((), dtype=rdt, device="cuda")
for _ in range(5):
    _call("qg_overlap_loss", _code(cdt), _ptr(phi), _ptr(out), _ptr(G), _ptr(loss), S, n, m, 1.0 / (S * m), _stream())
torch.cuda.synchronize()
print("ok", float(loss))
