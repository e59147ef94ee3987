"""NTT-PBS throughput on one GPU at the reference's parameters (test/mod.rs:106-130:
n_lwe 742, k 1, N 2048, base_log 23, level 1, Solinas prime), device-resident, fused vs composed,
with the CPU oracle's single-thread time beside it.  Prints one JSON object per line.

Usage: python tools/pbs_bench.py [--batches 148,592,2368] [--bnf] [--no-cpu] [--composed]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import tfhe_ntt_b200 as T  # noqa: E402
from tfhe_ntt_b200 import ntt64_pbs as G  # noqa: E402

P = (1 << 64) - (1 << 32) + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="148,296,592,1184,2368")
    ap.add_argument("--n-lwe", type=int, default=742)
    ap.add_argument("--poly", type=int, default=2048)
    ap.add_argument("--glwe-dim", type=int, default=1)
    ap.add_argument("--base-log", type=int, default=23)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--bnf", action="store_true")
    ap.add_argument("--composed", action="store_true")
    ap.add_argument("--cluster", action="store_true", help="two-CTA cluster (latency) kernel")
    ap.add_argument("--fused", action="store_true", help="one-CTA fused kernel only")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    n_lwe, N, gs = a.n_lwe, a.poly, a.glwe_dim + 1
    plan = T.prime64.Plan.try_new(N, P)
    bsk = (rng.integers(0, 1 << 63, n_lwe * a.level * gs * gs * N, dtype=np.uint64) * np.uint64(2)) % np.uint64(P)
    key = G.NttLweBootstrapKey.from_container(plan, bsk, n_lwe, gs, a.base_log, a.level)
    dev = torch.device("cuda:0")
    lut = torch.from_numpy((rng.integers(0, 1 << 62, gs * N, dtype=np.uint64)).view(np.int64)).to(dev)
    path = G.PATH_COMPOSED if a.composed else G.PATH_CLUSTER if a.cluster else G.PATH_FUSED if a.fused else G.PATH_AUTO
    for batch in [int(x) for x in a.batches.split(",")]:
        lwe_h = rng.integers(1, 1 << 62, (batch, n_lwe + 1), dtype=np.uint64)
        lwe = torch.from_numpy(lwe_h.view(np.int64)).to(dev)
        acc = torch.empty((batch, gs * N), dtype=torch.int64, device=dev)
        out = torch.empty((batch, (gs - 1) * N + 1), dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream()

        def step():
            G.blind_rotate_ntt64_device(key, lwe, lut, 1, acc, batch, bnf=a.bnf, path=path, stream=st)
            G.extract_lwe_sample_device(key, acc, out, batch, bnf=a.bnf, stream=st)

        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        ntts = batch * n_lwe * (gs * a.level + gs)
        print(json.dumps({"what": "pbs", "variant": "bnf" if a.bnf else "classic",
                          "path": "composed" if a.composed else "cluster" if a.cluster else "fused" if a.fused else "auto", "batch": batch, "ms": ms,
                          "pbs_per_s": batch / ms * 1e3, "ms_per_pbs_latency": ms,
                          "ntt_per_s": ntts / ms * 1e3, "n_lwe": n_lwe, "N": N, "k": a.glwe_dim,
                          "level": a.level}), flush=True)
    if not a.no_cpu:
        import oracle_lib as O
        oplan = O.OraclePlan(64, N, P)
        opbs = O.OraclePbs(oplan, bsk, n_lwe, gs, a.base_log, a.level)
        lut_h = lut.cpu().numpy().view(np.uint64) % np.uint64(P)
        ct = lwe_h[0] % np.uint64(P)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            (opbs.pbs_bnf if a.bnf else opbs.pbs)(ct, lut_h)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"what": "cpu_oracle_pbs", "threads": 1, "ms_per_pbs": dt * 1e3, "pbs_per_s": 1 / dt}))


if __name__ == "__main__":
    main()
