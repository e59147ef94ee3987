"""One launch each of the kernels that have an ncu summary under profiles/ besides the headline ones:
the TMA-staged strided pass (N=65536), the fused CRT polymul (C4 shape) and the fused external product (C3 shape).
  ncu --set full -k regex:'tma|polymul_fused|ext_product' python tools/ncu_targets.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfhe_ntt_b200 as T

st = torch.cuda.current_stream()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
# large-N transform: 2048 polynomials of 65536 u64
n, p = 65536, T.prime64.SOLINAS_PRIME
plan = T.prime64.Plan.try_new(n, p)
d = torch.randint(0, 1 << 62, (2048, n), dtype=torch.int64, device="cuda")
for _ in range(reps):
    plan.fwd_device(d, 2048, stream=st)
    plan.inv_device(d, 2048, stream=st)
del d
# C4: native64::Plan32 N=4096, 4096 products
pl = T.native64.Plan32.try_new(4096)
a, b = (torch.randint(-(1 << 62), 1 << 62, (4096, 4096), dtype=torch.int64, device="cuda") for _ in range(2))
o = torch.empty_like(a)
for _ in range(reps):
    pl.negacyclic_polymul_device(o, a, b, stream=st)
del a, b, o
# C3: external product k=1, l=2, 4096 LWEs
n = 2048
plan = T.prime64.Plan.try_new(n, p)
x = torch.randint(0, 1 << 62, (4096, 4, n), dtype=torch.int64, device="cuda")
g = torch.randint(0, 1 << 62, (4, 2, n), dtype=torch.int64, device="cuda")
out = torch.empty((4096, 2, n), dtype=torch.int64, device="cuda")
for _ in range(reps):
    plan.ext_product_device(out, x, g, 4, 2, stream=st)
torch.cuda.synchronize()
print("ok")
