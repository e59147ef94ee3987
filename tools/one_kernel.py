"""Run one batched transform a few times (for ncu): python tools/one_kernel.py <bits> <n> <p> [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfhe_ntt_b200 as T

bits, n, p = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mod = T.prime64 if bits == 64 else T.prime32
plan = mod.Plan.try_new(n, p)
batch = (1 << 30) // (n * bits // 8)
d = torch.randint(0, 1 << 29, (batch, n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
st = torch.cuda.current_stream()
for _ in range(iters):
    plan.fwd_device(d, batch, stream=st)
    plan.inv_device(d, batch, stream=st)
torch.cuda.synchronize()
print("ok")
