"""The reference's own benchmark matrix (tfhe-ntt/benches/ntt.rs:83-235) on the GPU engine, with the
same bench ids: fwd-32-<p>-<n>, inv-32-..., fwd-64-..., inv-64-..., native32-32-<n>, nativebinary32-32-<n>,
native32-52-<n>, ... native128-32-<n>.  Reports batched device-resident throughput (units/s) and the
equivalent time per unit in ns, so the numbers line up with criterion's per-call latencies.
Developer tool; writes JSON to stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfhe_ntt_b200 as T

LP = T.prime.largest_prime_in_arithmetic_progression64
WORKING_SET = 256 << 20  # bytes per operand: larger than the 126 MB L2


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
    ns = [256, 512, 1024, 2048, 4096, 8192, 16384, 32768]
    out = {}
    p32 = [LP(1 << 16, 1, 1 << 29, 1 << 30), LP(1 << 16, 1, 1 << 30, 1 << 31), LP(1 << 16, 1, 1 << 31, 1 << 32)]
    p64 = [LP(1 << 16, 1, 1 << 49, 1 << 50), LP(1 << 16, 1, 1 << 50, 1 << 51), LP(1 << 16, 1, 1 << 61, 1 << 62),
           LP(1 << 16, 1, 1 << 62, 1 << 63), T.prime64.SOLINAS_PRIME, LP(1 << 16, 1, 1 << 63, (1 << 64) - 1)]
    for bits, primes in ((32, p32), (64, p64)):
        mod = T.prime32 if bits == 32 else T.prime64
        dt = torch.int32 if bits == 32 else torch.int64
        for n in ns:
            batch = WORKING_SET // (n * bits // 8)
            d = torch.randint(0, 1 << 29, (batch, n), dtype=dt, device="cuda")
            for p in primes:
                plan = mod.Plan.try_new(n, p)
                for name, fn in (("fwd", plan.fwd_device), ("inv", plan.inv_device)):
                    t = timeit(lambda: fn(d, batch, stream=st))
                    out["%s-%d-%d-%d" % (name, bits, p, n)] = {"batch": batch, "units_per_s": batch / t, "ns_per_unit": t / batch * 1e9}
            del d
    kinds = [("native32-32", T.native32.Plan32, 4), ("nativebinary32-32", T.native_binary32.Plan32, 4),
             ("native32-52", T.native32.Plan52, 4), ("nativebinary32-52", T.native_binary32.Plan52, 4),
             ("native64-32", T.native64.Plan32, 8), ("nativebinary64-32", T.native_binary64.Plan32, 8),
             ("native64-52", T.native64.Plan52, 8), ("nativebinary64-52", T.native_binary64.Plan52, 8),
             ("native128-32", T.native128.Plan32, 16), ("nativebinary128-32", T.native_binary128.Plan32, 16)]
    for name, cls, vb in kinds:
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
            out["%s-%d" % (name, n)] = {"batch": batch, "units_per_s": batch / t, "ns_per_unit": t / batch * 1e9}
            del lhs, rhs, prod
    print(json.dumps(out, indent=0))


if __name__ == "__main__":
    main()
