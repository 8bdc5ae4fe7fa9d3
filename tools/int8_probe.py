"""Probe: rate of int8 x int8 -> int32 GEMMs (cuBLASLt through torch._int_mm) at the Omega-build shapes.
Run on the GPU box:  python tools/int8_probe.py [q]"""
import sys, time, torch
q = int(sys.argv[1]) if len(sys.argv) > 1 else 10016
dev = "cuda:0"
torch.manual_seed(0)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for d in (1, 2, 4, 6):
    K = q * d
    a = torch.randint(-64, 65, (q, K), dtype=torch.int8, device=dev)          # row-major: K contiguous
    bt = torch.randint(-64, 65, (q, K), dtype=torch.int8, device=dev)         # B' row-major -> B column-major
    for name, fn in (("TN", lambda: torch._int_mm(a, bt.t())), ("NN", lambda: torch._int_mm(a, bt.t().contiguous()))):
        try:
            ms = t(fn)
            print(f"q={q} K={K} {name}: {ms:.3f} ms  {2.0*q*q*K/ms/1e9:.0f} TOPS", flush=True)
        except Exception as ex:
            print(f"q={q} K={K} {name}: failed {ex}", flush=True)
    del a, bt
# exactness spot check
a = torch.randint(-64, 65, (256, 4096), dtype=torch.int8, device=dev)
b = torch.randint(-64, 65, (256, 4096), dtype=torch.int8, device=dev)
c = torch._int_mm(a, b.t())
ref = a.double() @ b.double().t()
print("exact:", bool((c.double() == ref).all()))
# FP64 DGEMM for comparison
x = torch.randn(q, q, dtype=torch.float64, device=dev); y = torch.randn(q, q, dtype=torch.float64, device=dev)
ms = t(lambda: x @ y, 3)
print(f"dgemm q={q}: {ms:.2f} ms {2.0*q**3/ms/1e9:.1f} TF")
