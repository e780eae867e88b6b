"""complex64 fused expm + pull-back (qg_expm_fwd_vjp) per dim with the tensor-core GEMM on and off:
    python tools/tc_expm_sweep.py [dims...]"""
import json, math, sys, torch
sys.path.insert(0, ".")
from qgrad_b200 import _lib
from qgrad_b200.qgrad_qutip import _call, _code, _ptr, _stream
lib = _lib.load()
dims = [int(x) for x in sys.argv[1:]] or [128, 256, 512, 1024]
dt = torch.complex64
out = []
for n in dims:
    batch = max(8, int(min(2e9 / (32 * n * n), 0.25 * 40e12 / (8.0 * n ** 3 * 19))))
    g = torch.Generator(device="cuda").manual_seed(n)
    X = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
    A = (-0.5j / math.sqrt(n)) * (X + X.mH)
    G = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
    ms = 1
    ws = torch.empty(lib.qg_expm_workspace_bytes(0, n, batch, 1, ms), dtype=torch.uint8, device="cuda")
    Y, Ab = torch.empty_like(A), torch.empty_like(A)
    rec = {"dim": n, "batch": batch}
    for name, on in (("tc", 1), ("fma", 0)):
        lib.qg_set_c64_tensor(on)
        call = lambda: _call("qg_expm_fwd_vjp", _code(dt), _ptr(A), _ptr(G), _ptr(Y), _ptr(Ab), batch, n, 0, _ptr(ws), ws.numel(), ms, _stream())
        for _ in range(2): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): call()
        e1.record(); torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 5e3
        rec[name + "_matrices_per_s"] = batch / sec
        rec[name + "_tflops_16_products"] = 8.0 * n ** 3 * 16 * batch / sec / 1e12
        rec[name + "_Y"] = Y.clone(); rec[name + "_Ab"] = Ab.clone()
    rel = lambda a, b: float(torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1)))
    rec["tc_vs_fma_value"] = rel(rec.pop("tc_Y"), rec.pop("fma_Y")); rec["tc_vs_fma_pullback"] = rel(rec.pop("tc_Ab"), rec.pop("fma_Ab"))
    print(json.dumps(rec), flush=True); out.append(rec)
lib.qg_set_c64_tensor(1)
json.dump(out, open("gpurun_out/tc_expm_sweep.json", "w"), indent=1)
