import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import tfhe_ntt_b200 as T
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
st = torch.cuda.current_stream()
n, batch = 4096, 8192
pl = T.native64.Plan32.try_new(n)
val = torch.randint(-(1 << 62), 1 << 62, (batch, n), dtype=torch.int64, device="cuda")
res = [torch.empty((batch, n), dtype=torch.int32, device="cuda") for _ in range(5)]
tf = timeit(lambda: pl.fwd_device(val, res, batch, stream=st))
ti = timeit(lambda: pl.inv_device(val, res, batch, stream=st))
p32 = T.prime32.Plan.try_new(n, 0x3F5A0001)
t1 = timeit(lambda: p32.fwd_device(res[0], batch, stream=st))
t2 = timeit(lambda: p32.inv_device(res[0], batch, stream=st))
print("native64::Plan32 n=4096 batch=8192: fwd %.3f ms, inv %.3f ms; one residue NTT fwd %.3f ms inv %.3f ms (x5 = %.3f / %.3f)" % (tf, ti, t1, t2, 5 * t1, 5 * t2))
