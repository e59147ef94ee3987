"""Latency of the per-polynomial drop-in calls (Plan::fwd / inv on one host polynomial, BASELINE config C1) and of
small host batches; wall clock, after warm-up, next to the CPU port of the reference doing the same call on one core.
Developer tool: NTT_B200_ZERO_COPY=0 python tools/latency_bench.py gives the staged (two DMA copies) path."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import tfhe_ntt_b200 as T


def bench(fn, reps=500, warm=50):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def cpu_per_poly_us(n, p, bits=64):
    import oracle_lib as O
    ref = O.OraclePlan(bits, n, p, _lib=O._load(native=True))  # -march=native build: AVX-512 where the host has it
    buf = (np.arange(n, dtype=np.uint64) * np.uint64(12345) % np.uint64(p)).astype(np.uint64 if bits == 64 else np.uint32)
    isa = ref.fwd_batch_inplace(buf, 1, simd=True)
    return bench(lambda: ref.fwd_batch_inplace(buf, 1, simd=True), reps=2000), isa


def main():
    mode = "zero-copy (mapped pinned buffer)" if os.environ.get("NTT_B200_ZERO_COPY", "1") != "0" else "staged (H2D + D2H copies)"
    print("# small-call path: %s" % mode)
    p = T.prime64.SOLINAS_PRIME
    for n in (1024, 2048, 4096):
        plan = T.prime64.Plan.try_new(n, p)
        buf = (np.arange(n, dtype=np.uint64) * np.uint64(12345)) % np.uint64(p)
        cpu, isa = cpu_per_poly_us(n, p)
        print("prime64 Solinas n=%d: fwd(one polynomial) %.1f us, inv %.1f us   | CPU port (%s, one core) %.1f us" % (
            n, bench(lambda: plan.fwd(buf)), bench(lambda: plan.inv(buf)), isa, cpu))
        for batch in (4, 16, 256):
            b = np.tile(buf, (batch, 1))
            print("   fwd_batch(%d polynomials) %.1f us  (%.2f us per polynomial)" % (batch, bench(lambda: plan.fwd_batch(b), reps=100), bench(lambda: plan.fwd_batch(b), reps=100) / batch))
    pl = T.native64.Plan32.try_new(1024)
    a = np.arange(1024, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    b = a[::-1].copy()
    prod = np.zeros_like(a)
    print("native64::Plan32 n=1024: negacyclic_polymul(one product) %.1f us" % bench(lambda: pl.negacyclic_polymul(prod, a, b)))
    plan = T.prime32.Plan.try_new(2048, 1073479681)
    buf = (np.arange(2048, dtype=np.uint32) * np.uint32(12345)) % np.uint32(1073479681)
    print("prime32 n=2048: fwd(one polynomial) %.1f us" % bench(lambda: plan.fwd(buf)))


if __name__ == "__main__":
    main()
