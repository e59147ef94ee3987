"""Developer timing sweep (not the contract bench): device-resident kernels, CUDA events."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
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
    cases = [(64, 1024, T.prime64.SOLINAS_PRIME), (64, 2048, T.prime64.SOLINAS_PRIME), (64, 4096, T.prime64.SOLINAS_PRIME),
             (64, 2048, 4611686018427322369), (64, 2048, 9223372036853661697), (64, 2048, 18446744073707716609),
             (32, 2048, 1073479681), (32, 4096, 1068236801), (32, 2048, 2147352577), (32, 2048, 4293918721),
             (64, 16384, T.prime64.SOLINAS_PRIME), (64, 65536, T.prime64.SOLINAS_PRIME), (64, 256, T.prime64.SOLINAS_PRIME),
             (64, 512, T.prime64.SOLINAS_PRIME), (64, 512, 4611686018427322369), (32, 256, 1073479681), (32, 512, 1073479681),
             (32, 1024, 1073479681)]
    for bits, n, p in cases:
        mod = T.prime64 if bits == 64 else T.prime32
        plan = mod.Plan.try_new(n, p)
        eb = bits // 8
        total_bytes = 1 << 30
        batch = total_bytes // (n * eb)
        dt = torch.int64 if bits == 64 else torch.int32
        d = torch.randint(0, 1 << 30, (batch, n), dtype=dt, device="cuda")
        tf = timeit(lambda: plan.fwd_device(d, batch, stream=st))
        ti = timeit(lambda: plan.inv_device(d, batch, stream=st))
        alg = 2 * batch * n * eb
        print("u%d n=%-6d p=%-20d batch=%-7d fwd %.3f ms (%.1f M NTT/s, %.1f%% HBM)  inv %.3f ms (%.1f M NTT/s, %.1f%% HBM)" % (
            bits, n, p, batch, tf, batch / tf / 1e3, alg / tf / 1e6 / HBM * 100, ti, batch / ti / 1e3, alg / ti / 1e6 / HBM * 100), flush=True)
        if n <= 4096:
            rhs = torch.randint(0, 1 << 30, (batch, n), dtype=dt, device="cuda")
            out = torch.empty_like(d)
            tm = timeit(lambda: plan.fwd_mac_inv_device(out, d, rhs, None, stream=st))
            alg3 = 3 * batch * n * eb
            print("     fused fwd*rhs->inv %.3f ms (%.1f M units/s, %.1f%% HBM of 3 arrays)" % (tm, batch / tm / 1e3, alg3 / tm / 1e6 / HBM * 100), flush=True)
            del rhs, out
        del d
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
