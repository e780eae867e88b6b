"""Library ceilings on this box: cuBLAS ZGEMM / CGEMM / DGEMM through torch.matmul, and
torch.linalg.matrix_exp (the 'library baseline' BASELINE.md section 4 names).  JSON to stdout."""
import json, time, torch

def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {}
for name, dt, flop_per_mac in (("zgemm", torch.complex128, 8), ("cgemm", torch.complex64, 8),
                               ("dgemm", torch.float64, 2), ("sgemm", torch.float32, 2)):
    for n in (1024, 4096):
        a = torch.randn(n, n, dtype=dt, device="cuda"); b = torch.randn(n, n, dtype=dt, device="cuda")
        ms = t(lambda: a @ b)
        out[f"{name}_{n}_tflops"] = round(flop_per_mac * n**3 / ms / 1e9, 3)
torch.backends.cuda.matmul.allow_tf32 = False
for n, B in ((4, 1 << 18), (16, 1 << 15), (64, 4096), (256, 64), (1024, 4)):
    for dt, nm in ((torch.complex128, "c128"), (torch.complex64, "c64")):
        x = torch.randn(B, n, n, dtype=dt, device="cuda")
        h = (x + x.mH) / (2 * n ** 0.5)
        a = (-1j * h).requires_grad_(True)
        g = torch.randn(B, n, n, dtype=dt, device="cuda")
        ms = t(lambda: torch.linalg.matrix_exp(a.detach()), 3)
        out[f"torch_matrix_exp_{nm}_n{n}_mat_per_s"] = round(B / ms * 1e3, 2)
        def fb():
            y = torch.linalg.matrix_exp(a)
            a.grad = None
            y.backward(g)
        ms = t(fb, 3)
        out[f"torch_matrix_exp_vjp_{nm}_n{n}_mat_per_s"] = round(B / ms * 1e3, 2)
print(json.dumps(out, indent=1))
