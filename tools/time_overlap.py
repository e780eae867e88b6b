"""qg_overlap_loss timing at the learning step's sizes (settings 8 and 1), CUDA events, L2 flushed by size."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
for cdt in (torch.complex128, torch.complex64):
    rdt = torch.float32 if cdt == torch.complex64 else torch.float64
    for S in (8, 1):
        n, m = 1024, 512
        reps = 4 if S == 8 else 32                                  # rotate over buffers so that each call misses L2
        bufs = [(torch.randn((S, n, m), dtype=cdt, device="cuda"), torch.randn((S, n, m), dtype=cdt, device="cuda"),
                 torch.empty((S, n, m), dtype=cdt, device="cuda")) for _ in range(reps)]
        loss = torch.zeros((), dtype=rdt, device="cuda")
        def run():
            for phi, out, G in bufs:
                _call("qg_overlap_loss", _code(cdt), _ptr(phi), _ptr(out), _ptr(G), _ptr(loss), S, n, m, 1.0 / (S * m), _stream())
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (5 * reps) * 1e3
        nbytes = S * n * m * 3 * (16 if cdt == torch.complex128 else 8)
        print(f"variant {os.environ.get('QG_OV_VARIANT', 'default')} {str(cdt):16s} S={S} {us:8.1f} us {nbytes / us / 1e3:8.1f} GB/s", flush=True)
        del bufs
