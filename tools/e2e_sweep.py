"""Sweeps the staging chunk size / stream count of the host-buffer pipeline (env NTT_B200_CHUNK_MIB,
NTT_B200_STREAMS are read once per process, so each point runs in its own process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, time, numpy as np, torch
sys.path.insert(0, %r)
import tfhe_ntt_b200 as T
P = T.prime64.SOLINAS_PRIME
n, eb = 2048, int(sys.argv[1])
plan = T.prime64.Plan.try_new(n, P)
rng = np.random.default_rng(0)
lhs = torch.from_numpy((rng.integers(0, 1 << 62, (eb, n), dtype=np.uint64)).view(np.int64)).pin_memory()
out = torch.empty_like(lhs).pin_memory()
rhs = (rng.integers(0, 1 << 62, n, dtype=np.uint64))
lv, ov = lhs.numpy().view(np.uint64), out.numpy().view(np.uint64)
for _ in range(2):
    plan.fwd_mac_inv_batch(ov, lv, rhs)
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    plan.fwd_mac_inv_batch(ov, lv, rhs)
dt = (time.perf_counter() - t0) / reps
print(json.dumps({"eb": eb, "ms": dt * 1e3, "ntt_per_s": 2 * eb / dt, "GBps_each_way": eb * n * 8 / dt / 1e9}))
''' % ROOT

for eb in (32768, 65536):
    for mib in [int(x) for x in os.environ.get("SWEEP_MIB", "4,8,16,32").split(",")]:
        for ns in [int(x) for x in os.environ.get("SWEEP_STREAMS", "3,4,6").split(",")]:
            env = dict(os.environ, NTT_B200_CHUNK_MIB=str(mib), NTT_B200_STREAMS=str(ns))
            r = subprocess.run([sys.executable, "-c", CODE, str(eb)], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(mib, ns, line, flush=True)
