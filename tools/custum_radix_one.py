"""Runs the custum_radix forward transform a few times on a 512 MiB device batch (for ncu):
python tools/custum_radix_one.py <n> <p> [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import tfhe_ntt_b200.custum_radix as cr

n, p = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
batch = (1 << 29) // (4 * n)
tw = torch.from_numpy(cr.make_twiddles(n, p).view(np.int32)).cuda()
d = torch.randint(0, p, (batch, n), dtype=torch.int64, device="cuda").to(torch.int32)
st = torch.cuda.current_stream()
for _ in range(iters):
    cr.fft_device(cr.RADIX2, d, n, batch, tw, p, stream=st)
torch.cuda.synchronize()
print("ok")
