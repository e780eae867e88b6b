"""Times forward-only expm (qg_expm) for one dim: python tools/time_expm.py DIM BATCH [c128|c64]"""
import math, sys, torch
sys.path.insert(0, ".")
import qgrad_b200.qgrad_qutip as q
n, batch = int(sys.argv[1]), int(sys.argv[2])
dt = torch.complex64 if (len(sys.argv) > 3 and sys.argv[3] == "c64") else torch.complex128
g = torch.Generator(device="cuda").manual_seed(n)
X = torch.randn((batch, n, n), dtype=dt, device="cuda", generator=g)
A = (-0.5j / math.sqrt(n)) * (X + X.mH)
for _ in range(2): q.expm(A, max_squarings=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): q.expm(A, max_squarings=3)
e1.record(); torch.cuda.synchronize()
print(f"expm fwd n={n} batch={batch} {dt}: {e0.elapsed_time(e1)/5*1e3:.1f} us per call")
