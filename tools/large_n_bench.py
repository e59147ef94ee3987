"""Polynomials longer than one CTA (2^13 .. 2^16): the default dispatch, a forced number of strided top
stages (NTT_B200_DEPTH=k) or the cluster-of-eight single-pass kernels (NTT_B200_CLUSTER=1).
Developer tool; CUDA events, 1 GiB working sets."""
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
    tag = "cluster8" if os.environ.get("NTT_B200_CLUSTER") == "1" else "default" if "NTT_B200_DEPTH" not in os.environ else "depth=" + os.environ["NTT_B200_DEPTH"]
    cases = [(64, 1 << k, T.prime64.SOLINAS_PRIME) for k in (13, 14, 15, 16)]
    cases += [(64, 1 << 14, 4611686018427322369), (32, 1 << 13, 1073479681), (32, 1 << 14, 1073479681), (32, 1 << 15, 1073479681), (32, 1 << 16, 1073479681)]
    for bits, n, p in cases:
        mod = T.prime64 if bits == 64 else T.prime32
        plan = mod.Plan.try_new(n, p)
        eb = bits // 8
        batch = (1 << 30) // (n * eb)
        d = torch.randint(0, 1 << 29, (batch, n), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
        tf = timeit(lambda: plan.fwd_device(d, batch, stream=st))
        ti = timeit(lambda: plan.inv_device(d, batch, stream=st))
        alg = 2 * batch * n * eb
        print("%-10s u%d n=%-6d p=%-20d batch=%-6d fwd %.3f ms (%.2f M NTT/s, %.1f%% HBM)  inv %.3f ms (%.2f M NTT/s, %.1f%% HBM)" % (
            tag, bits, n, p, batch, tf, batch / tf / 1e3, alg / tf / 1e6 / HBM * 100, ti, batch / ti / 1e3, alg / ti / 1e6 / HBM * 100), flush=True)
        del d


if __name__ == "__main__":
    main()
