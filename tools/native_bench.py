"""negacyclic_polymul throughput of all ten CRT plans with the reference's bench ids
(tfhe-ntt/benches/ntt.rs:157-235: native32-32-<n>, nativebinary32-32-<n>, native32-52-<n>, ...,
native128-32-<n>), device resident, CUDA events, working set beyond L2.  Developer tool."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfhe_ntt_b200 as T

WORKING_SET = 256 << 20


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    st = torch.cuda.current_stream()
    ns = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1024,2048,4096,8192".split(","))]
    kinds = [("native32-32", T.native32.Plan32, 4, 3), ("nativebinary32-32", T.native_binary32.Plan32, 4, 2),
             ("native32-52", T.native32.Plan52, 4, 2), ("nativebinary32-52", T.native_binary32.Plan52, 4, 1),
             ("native64-32", T.native64.Plan32, 8, 5), ("nativebinary64-32", T.native_binary64.Plan32, 8, 3),
             ("native64-52", T.native64.Plan52, 8, 3), ("nativebinary64-52", T.native_binary64.Plan52, 8, 2),
             ("native128-32", T.native128.Plan32, 16, 10), ("nativebinary128-32", T.native_binary128.Plan32, 16, 5)]
    for name, cls, vb, primes in kinds:
        for n in ns:
            plan = cls.try_new(n)
            if plan is None:
                continue
            batch = max(16, (WORKING_SET // 2) // (n * vb))
            words = n * vb // 8
            lhs = torch.randint(-(1 << 62), 1 << 62, (batch, words), dtype=torch.int64, device="cuda")
            if "binary" in name:
                rhs = torch.zeros((batch, words), dtype=torch.int64, device="cuda")
                if vb == 4:
                    rhs.view(torch.int32)[:] = torch.randint(0, 2, (batch, n), dtype=torch.int32, device="cuda")
                elif vb == 8:
                    rhs[:] = torch.randint(0, 2, (batch, n), dtype=torch.int64, device="cuda")
                else:
                    rhs.view(batch, n, 2)[:, :, 0] = torch.randint(0, 2, (batch, n), dtype=torch.int64, device="cuda")
            else:
                rhs = torch.randint(-(1 << 62), 1 << 62, (batch, words), dtype=torch.int64, device="cuda")
            prod = torch.empty_like(lhs)
            t = timeit(lambda: plan.negacyclic_polymul_device(prod, lhs, rhs, batch, stream=st), iters=3, warm=1)
            ntts = (2 if "binary" not in name else 2) * primes + primes  # 2 forward + 1 inverse per prime
            print("%-22s batch %6d  %9.3f ms  %8.3f M products/s  %7.1f M residue NTT/s  %5.1f %% of the HBM bound (3 value arrays)" % (
                "%s-%d" % (name, n), batch, t * 1e3, batch / t / 1e6, batch * ntts / t / 1e6,
                3 * batch * n * vb / t / 6543.4e9 * 100), flush=True)
            del lhs, rhs, prod


if __name__ == "__main__":
    main()
