"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the headers compile as C and C++, constructor rejections that need no GPU, the
sharding rule, and a world_size-2 gloo run of the multi-rank host logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tfhe_ntt_b200.h")


@pytest.fixture(scope="module")
def T():
    import __graft_entry__
    __graft_entry__.build()
    import tfhe_ntt_b200
    return tfhe_ntt_b200


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ntt_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(T):
    syms = declared_symbols()
    assert len(syms) > 60
    lib = ctypes.CDLL(T.library_path())
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # and the python binding knows every one of them
    from tfhe_ntt_b200 import _binding
    assert sorted(_binding.SIGNATURES) == syms


def test_headers_compile_as_c_and_cpp():
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)
    for src in ("example_prime.cpp", "example_pbs.cpp", "example_native.cpp"):
        subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", inc,
                        os.path.join(ROOT, "tests", "cpp", src)], check=True)


def test_try_new_none_needs_no_gpu(T):
    # the acceptance rules run on the host before any CUDA call (prime64.rs:769-774, prime32.rs:667-672)
    assert T.prime64.Plan.try_new(2048, 1024) is None
    assert T.prime64.Plan.try_new(8, T.prime64.SOLINAS_PRIME) is None
    assert T.prime64.Plan.try_new(24, T.prime64.SOLINAS_PRIME) is None
    assert T.prime32.Plan.try_new(16, 1062862849) is None
    assert T.prime32.Plan.try_new(1 << 17, 1062862849) is None
    assert T.native64.Plan32.try_new(65536) is None
    assert T.prime.is_prime64(T.prime64.SOLINAS_PRIME)
    assert T.prime.largest_prime_in_arithmetic_progression64(6, 5, 0, (1 << 64) - 1) == 18446744073709551557
    assert T.prime.largest_prime_in_arithmetic_progression64(1, 0, 14, 16) is None


def test_no_cpu_fallback(T):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.NttB200Error):
        T.prime64.Plan.try_new(2048, T.prime64.SOLINAS_PRIME)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "tfhe-rs-main_modified_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "tfhe_ntt_oracle" not in text, f


def test_shard_ranges(T):
    sh = T.sharding
    for batch in [0, 1, 7, 8, 4096, 65537]:
        for g in [1, 2, 3, 4, 8]:
            ranges = [sh.shard_range(batch, g, r) for r in range(g)]
            assert ranges[0][0] == 0 and ranges[-1][1] == batch
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import tfhe_ntt_b200 as T
import oracle_lib as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n, p, batch = 256, O.SOLINAS_P, 37
rng = np.random.default_rng(5)
x = (rng.integers(0, 1 << 63, size=(batch, n), dtype=np.uint64) * 2) %% np.uint64(p)
b, e = T.sharding.shard_range(batch, world, rank)
plan = O.OraclePlan(64, n, p)           # CPU stand-in for the per-GPU kernel (no GPU in this test)
mine = plan.fwd(x[b:e])
parts = [None] * world
dist.all_gather_object(parts, (b, e, mine))
t = T.sharding.max_over_ranks(1.0 + rank)
if rank == 0:
    out = np.zeros_like(x)
    for (bb, ee, part) in parts:
        out[bb:ee] = part
    assert (out == plan.fwd(x)).all()
    assert t == float(world)
    print("gloo ok")
dist.destroy_process_group()
'''


def test_gloo_world_size_2_sharding(T, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gloo ok" in r.stdout


def test_rust_crate_build_lists_every_cuda_source():
    """rust/tfhe-ntt-b200/build.rs compiles the same translation units as build.py (no rustc here to try it)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "tfhe-rs-main_modified_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text = open(os.path.join(ROOT, "rust", "tfhe-ntt-b200", "build.rs")).read()
    for src in mod.SOURCES:
        assert '"%s"' % src in text, src
    # and every extern "C" symbol the Rust ffi block names exists in the header
    ffi = open(os.path.join(ROOT, "rust", "tfhe-ntt-b200", "src", "ffi.rs")).read()
    header = open(HEADER).read()
    import re
    for name in re.findall(r"pub fn (ntt_b200_\w+)\(", ffi):
        assert re.search(r"\b%s\s*\(" % name, header), name


def test_prime_helpers_agree_with_the_oracle_and_a_definition(T):
    """is_prime64 / largest_prime_in_arithmetic_progression64 are host functions of the library (prime.rs:76-186): against
    the oracle's restatement and, for small arguments, trial division; strong pseudoprimes and Carmichael numbers included."""
    import ctypes as C
    import random
    import oracle_lib as O
    L = O.lib()
    rnd = random.Random(7)
    hard = [2047, 1373653, 9080191, 25326001, 3215031751, 4759123141, 1122004669633, 2152302898747, 3474749660383,
            341550071728321, 3825123056546413051, 318665857834031151167461 % (1 << 64), 561, 1105, 1729, 2465, 2821,
            6601, 8911, 41041, 825265, 321197185, 5394826801, 232250619601, 9746347772161, (1 << 61) - 1, (1 << 64) - 59,
            (1 << 64) - 1, (1 << 32) + 15, (1 << 32) - 5, 0, 1, 2, 3, 4]
    values = hard + [rnd.getrandbits(rnd.randint(2, 64)) for _ in range(3000)] + [rnd.getrandbits(64) | 1 for _ in range(2000)]
    for v in values:
        assert bool(T.prime.is_prime64(v)) == bool(L.tfo_is_prime64(v)), v
    for v in range(0, 3000):
        want = v >= 2 and all(v % d for d in range(2, int(v ** 0.5) + 1))
        assert bool(T.prime.is_prime64(v)) == want, v
    import math
    for _ in range(300):
        factor = rnd.choice([1, 2, 6, 64, 1 << 12, 1 << 16, 2 * rnd.getrandbits(16) + 2])
        offset = rnd.choice([1, 5, 7, 11])
        if math.gcd(factor, offset) != 1:
            continue  # no prime in the progression: the search walks the whole range
        lo = rnd.getrandbits(rnd.randint(1, 62))
        hi = lo + factor * rnd.randint(0, 3000)
        got = T.prime.largest_prime_in_arithmetic_progression64(factor, offset, lo, hi)
        first = offset if lo <= offset else lo + (offset - lo) % factor  # smallest term >= lo
        if first > hi:
            want = None  # (the reference walks below zero on a range without a term, prime.rs:165-178)
        else:
            out = C.c_uint64(0)
            ok = L.tfo_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, C.byref(out))
            want = int(out.value) if ok else None
        assert got == want, (factor, offset, lo, hi, got, want)
    assert T.prime.largest_prime_in_arithmetic_progression64(0, 7, 5, 9) == 7
    assert T.prime.largest_prime_in_arithmetic_progression64(0, 8, 5, 9) is None
