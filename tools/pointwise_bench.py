"""Pointwise entry points (normalize, mul_assign_normalize, mul_accumulate) on device-resident 1 GiB arrays:
achieved fraction of the measured HBM copy peak.  Developer tool; CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfhe_ntt_b200 as T

HBM = 6543.4


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    st = torch.cuda.current_stream()
    for bits, n, p in [(64, 2048, T.prime64.SOLINAS_PRIME), (64, 2048, 4611686018427322369), (32, 2048, 1073479681), (32, 2048, 4293918721)]:
        mod = T.prime64 if bits == 64 else T.prime32
        plan = mod.Plan.try_new(n, p)
        eb = bits // 8
        batch = (1 << 30) // (n * eb)
        dt = torch.int64 if bits == 64 else torch.int32
        a, b, c = (torch.randint(0, 1 << 29, (batch, n), dtype=dt, device="cuda") for _ in range(3))
        gib = batch * n * eb
        t = timeit(lambda: plan.normalize_device(a, stream=st))
        print("u%d p=%-20d normalize            %.3f ms  %.0f GB/s (%.0f%% of HBM peak; 2 arrays)" % (bits, p, t, 2 * gib / t / 1e6, 2 * gib / t / 1e6 / HBM * 100))
        t = timeit(lambda: plan.mul_assign_normalize_device(a, b, stream=st))
        print("u%d p=%-20d mul_assign_normalize %.3f ms  %.0f GB/s (%.0f%%; 3 arrays)" % (bits, p, t, 3 * gib / t / 1e6, 3 * gib / t / 1e6 / HBM * 100))
        t = timeit(lambda: plan.mul_accumulate_device(a, b, c, stream=st))
        print("u%d p=%-20d mul_accumulate       %.3f ms  %.0f GB/s (%.0f%%; 4 arrays)" % (bits, p, t, 4 * gib / t / 1e6, 4 * gib / t / 1e6 / HBM * 100), flush=True)
        del a, b, c


if __name__ == "__main__":
    main()
