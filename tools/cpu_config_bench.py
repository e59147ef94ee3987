"""The BASELINE.json configs on the HOST cores: the oracle's C port of the reference's CPU path (scalar u32 /
AVX-512 Solinas where the host has it), independent units over all host threads in static contiguous chunks --
the decomposition the reference's callers use with rayon.  A bounded sample per config (a few seconds each).
Reported baseline beside tools/config_bench.py's GPU numbers; developer tool (it executes oracle/, like
bench.py's cpu_baseline leg)."""
import json
import os
import sys
import time
import multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O

THREADS = os.cpu_count() or 1


def chunks(total):
    per = (total + THREADS - 1) // THREADS
    return [(b, min(total, b + per)) for b in range(0, total, per)]


class ForkPool:
    """One forked process per host thread (the units are microsecond-sized C calls: Python threads would
    serialise on the interpreter lock between them).  map() runs chunk_fn over the static contiguous chunks."""

    def map(self, chunk_fn, parts):
        procs = []
        for part in parts:
            pr = mp.get_context("fork").Process(target=chunk_fn, args=(part,))
            pr.start()
            procs.append(pr)
        for pr in procs:
            pr.join()
        return [pr.exitcode for pr in procs]


def timed(fn, target_s=2.0):
    fn()
    t0 = time.perf_counter()
    fn()
    one = time.perf_counter() - t0
    reps = int(max(1, min(50, target_s / max(one, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    lib = O._load(native=True)
    rng = np.random.default_rng(0)
    out = {"threads": THREADS}
    pool = ForkPool()
    # C1: prime64 N=1024 Solinas, one polynomial, fwd + inv + normalize roundtrip on one core
    plan = O.OraclePlan(64, 1024, O.SOLINAS_P, _lib=lib)
    x = (rng.integers(0, 1 << 63, 1024, dtype=np.uint64) * np.uint64(2)) % np.uint64(O.SOLINAS_P)
    isa = plan.fwd_batch_inplace(x, 1, simd=True)
    t = timed(lambda: (plan.fwd_batch_inplace(x, 1, simd=True), plan.inv_batch_inplace(x, 1, simd=True)), 1.0)
    out["C1"] = {"us_per_fwd_inv_roundtrip": t * 1e6, "isa": isa, "cores": 1}
    # C2: prime32 N=2048 30-bit, fwd + mul_accumulate + inv per polynomial (scalar port)
    n, p, batch = 2048, 1073479681, 2048 * THREADS
    plan = O.OraclePlan(32, n, p, _lib=lib)
    lhs, rhs, acc = (rng.integers(0, p, (batch, n), dtype=np.uint64).astype(np.uint32) for _ in range(3))

    def c2_chunk(be):
        b, e = be
        for i in range(b, e):
            lib.tfo_plan32_fwd(plan.h, O._ptr(lhs[i]))
            lib.tfo_plan32_mul_accumulate(plan.h, O._ptr(acc[i]), O._ptr(lhs[i]), O._ptr(rhs[i]), n)
            lib.tfo_plan32_inv(plan.h, O._ptr(acc[i]))
    t = timed(lambda: list(pool.map(c2_chunk, chunks(batch))))
    out["C2"] = {"units_per_s": batch / t, "ntt_per_s": 2 * batch / t, "isa": "scalar", "cores": THREADS}
    # C3: Solinas N=2048 external product k=1 l=2: 4 fwd, 8 mul_accumulate, 2 inv per LWE (AVX-512 transforms)
    n, lwes = 2048, 512 * THREADS
    plan = O.OraclePlan(64, n, O.SOLINAS_P, _lib=lib)
    dig = (rng.integers(0, 1 << 63, (lwes, 4, n), dtype=np.uint64) * np.uint64(2)) % np.uint64(O.SOLINAS_P)
    ggsw = (rng.integers(0, 1 << 63, (4, 2, n), dtype=np.uint64) * np.uint64(2)) % np.uint64(O.SOLINAS_P)
    accs = np.zeros((lwes, 2, n), dtype=np.uint64)

    def c3_chunk(be):
        b, e = be
        for i in range(b, e):
            for r in range(4):
                if not lib.tfo_plan64_fwd_simd1(plan.h, O._ptr(dig[i, r])):
                    lib.tfo_plan64_fwd(plan.h, O._ptr(dig[i, r]))
                for c in range(2):
                    lib.tfo_plan64_mul_accumulate(plan.h, O._ptr(accs[i, c]), O._ptr(dig[i, r]), O._ptr(ggsw[r, c]), n)
            for c in range(2):
                if not lib.tfo_plan64_inv_simd1(plan.h, O._ptr(accs[i, c])):
                    lib.tfo_plan64_inv(plan.h, O._ptr(accs[i, c]))
    t = timed(lambda: list(pool.map(c3_chunk, chunks(lwes))))
    out["C3"] = {"external_products_per_s": lwes / t, "ntt_per_s": 6 * lwes / t, "cores": THREADS}
    # C4: native64::Plan32 N=4096 negacyclic_polymul (scalar port)
    n, batch = 4096, 128 * THREADS
    nplan = O.OracleNativePlan(O.NATIVE64_PLAN32, n)
    a = rng.integers(0, 1 << 63, (batch, n), dtype=np.uint64)
    b_ = rng.integers(0, 1 << 63, (batch, n), dtype=np.uint64)

    def c4_chunk(be):
        b, e = be
        for i in range(b, e):
            nplan.negacyclic_polymul(a[i], b_[i])
    t = timed(lambda: list(pool.map(c4_chunk, chunks(batch))))
    out["C4"] = {"products_per_s": batch / t, "residue_ntt_per_s": 15 * batch / t, "isa": "scalar", "cores": THREADS}
    # C5: Solinas N=65536 forward / inverse (AVX-512)
    n, batch = 65536, THREADS
    plan = O.OraclePlan(64, n, O.SOLINAS_P, _lib=lib)
    buf = (rng.integers(0, 1 << 63, (batch, n), dtype=np.uint64) * np.uint64(2)) % np.uint64(O.SOLINAS_P)
    tf = timed(lambda: plan.fwd_batch_inplace(buf, THREADS, simd=True))
    ti = timed(lambda: plan.inv_batch_inplace(buf, THREADS, simd=True))
    out["C5"] = {"fwd_ntt_per_s": batch / tf, "inv_ntt_per_s": batch / ti, "cores": THREADS}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
