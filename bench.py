#!/usr/bin/env python
"""Headline benchmark: forward + inverse negacyclic NTTs per second, N=2048, u64 Solinas prime
(2^64 - 2^32 + 1), batched and device-resident, on N B200s (one process per GPU, no collective:
the polynomial batch shards trivially).

  python bench.py --gpus 1 --steps K --warmup W          # our arm
  python bench.py --impl reference ...                   # tfhe-ntt's CPU path (C port) on host cores

One "step" = fwd over the whole per-GPU batch, then inv over it (2 * batch transforms).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N = 2048
SOLINAS_P = (1 << 64) - (1 << 32) + 1
BATCH_PER_GPU = 65536          # 65536 * 16 KiB = 1 GiB per GPU: far beyond the 126 MB L2
ALG_BYTES_PER_NTT = 2 * N * 8  # one in-place NTT reads and writes each coefficient once
E2E_BATCH = 65536              # host-buffer leg: 1 GiB in + 1 GiB out per step
NCU_TRAFFIC_PER_LAUNCH = {"fwd": 2.1018e9, "inv": 2.0891e9}  # dram read+write bytes of one launch, ncu --set full (profiles/r02_ncu_full_solinas2048_summary.md)
# Issue cost of one thread of the shipped kernels (two polynomials, 88 butterflies), from the SASS of
# ntt_fast_{fwd,inv}_kernel<Solinas64,11,1,2,false> weighted by the measured issue costs of
# profiles/r01_int_pipe_microbench.txt (IMAD.WIDE 4 clk, other IMAD 2 clk on the FMA-heavy pipe; IADD3 / LOP3 /
# SEL / ISETP 2 clk on the ALU pipe); a polynomial pair is 8 warps.
PIPE_CLK_PER_THREAD = {"fwd": {"fmaheavy": 2768, "alu": 2670}, "inv": {"fmaheavy": 2768, "alu": 3062}}  # profiles/r02_sass_hist_shipped_solinas2048.txt
# Busy fraction of the busier integer pipe that the shipped butterfly sustains when nothing else runs (register-resident
# loop, profiles/r02_solinas_bf_variants.md: 33.75 pipe clocks needed per warp-butterfly, 39.9 measured): the ceiling of
# roofline_int.
INT_PIPE_PEAK_FRAC = 33.75 / 39.9
METRIC = "fwd+inv NTTs/sec, N=2048 u64 prime, batched"
UNIT = "NTT/s"
WORKLOAD = "prime64 Solinas p=2^64-2^32+1 N=2048, batch %d polynomials per GPU, fwd then inv (in place, HBM-resident)" % BATCH_PER_GPU


def workload_config(world):
    """The `config` object, identical in the GPU arm and the reference arm (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH_PER_GPU, "n": N, "modulus": SOLINAS_P,
            "l2_policy": "inputs (1 GiB per GPU) far larger than the 126 MB L2",
            "parallelism": "batch sharded over %d GPU(s), no collective" % world}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs: NVML every ~2 ms when
    pynvml is importable (the timed region is tens of milliseconds), nvidia-smi otherwise."""

    def __init__(self, index):
        self.index = index
        self.samples, self.max_mhz, self.reasons = [], None, set()
        self._stop = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
        self._t = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)

    def _run_nvml(self):
        nv = self._nvml
        flags = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = get_reasons(self._h)
                for name, bit in flags.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s),
                "source": "nvml" if self._nvml else "nvidia-smi"}


def synth(batch, seed):
    """i.i.d. uniform coefficients in [0, p) from a fixed seed."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << 64, size=(batch, N), dtype=np.uint64, endpoint=False)
    x[x >= np.uint64(SOLINAS_P)] -= np.uint64(SOLINAS_P)
    return x


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle port of tfhe-ntt's scalar path; the one place bench.py may execute oracle/)
# ---------------------------------------------------------------------------------------------
CPU_ISA = ["scalar"]
CPU_NOTE = ("C port of tfhe-ntt's CPU path (oracle/): AVX-512 Solinas butterflies as in generic_solinas.rs:415-446 when the "
            "host has AVX-512F+DQ, else the scalar path; batch split over all host threads in static contiguous chunks; "
            "the Rust crate itself cannot be built in this image")


def cpu_leg(sample_polys, repeats, threads):
    import oracle_lib
    lib = oracle_lib._load(native=True)   # rebuilt with -march=native on this host
    plan = oracle_lib.OraclePlan(64, N, SOLINAS_P, _lib=lib)
    buf = synth(sample_polys, 0xC0FFEE03)
    isa = plan.fwd_batch_inplace(buf, threads, simd=True)  # warm
    plan.inv_batch_inplace(buf, threads, simd=True)
    t0 = time.perf_counter()
    for _ in range(repeats):
        plan.fwd_batch_inplace(buf, threads, simd=True)
        plan.inv_batch_inplace(buf, threads, simd=True)
    dt = time.perf_counter() - t0
    CPU_ISA[0] = isa
    return 2.0 * sample_polys * repeats / dt, dt


PBS_PARAMS = {"n_lwe": 742, "n": N, "k": 1, "base_log": 23, "level": 1}  # tfhe test/mod.rs:106-130 (TEST_PARAMS_3_BITS_SOLINAS_U64)


def pbs_inputs(batch, seed=1):
    """Random NTT-domain key, ciphertexts and LUT of the reference's parameter set (the timing does not depend on the
    values; parity of the same call is tests/test_pbs_gpu.py)."""
    rng = np.random.default_rng(seed)
    n_lwe, gs, level = PBS_PARAMS["n_lwe"], PBS_PARAMS["k"] + 1, PBS_PARAMS["level"]
    p = np.uint64(SOLINAS_P)
    bsk = (rng.integers(0, 1 << 63, n_lwe * level * gs * gs * N, dtype=np.uint64) * np.uint64(2)) % p
    lut = rng.integers(0, 1 << 62, gs * N, dtype=np.uint64)
    lwe = rng.integers(1, 1 << 62, (batch, n_lwe + 1), dtype=np.uint64)
    return bsk, lut, lwe


def pbs_cpu_leg(threads, seconds=8.0):
    """The PBS on the host cores: the oracle's restatement of programmable_bootstrap_ntt64_lwe_ciphertext with the
    vectorised transforms, independent ciphertexts over all host threads (static contiguous chunks)."""
    import oracle_lib
    lib = oracle_lib._load(native=True)
    plan = oracle_lib.OraclePlan(64, N, SOLINAS_P, _lib=lib)
    gs = PBS_PARAMS["k"] + 1
    bsk, lut, lwe = pbs_inputs(threads)
    pbs = oracle_lib.OraclePbs(plan, bsk, PBS_PARAMS["n_lwe"], gs, PBS_PARAMS["base_log"], PBS_PARAMS["level"])
    lib.tfo_use_simd_transforms(1)
    try:
        t0 = time.perf_counter()
        pbs.pbs_batch(lwe, lut, threads)  # one ciphertext per thread: sizes the sample
        one = time.perf_counter() - t0
        rounds = int(max(1, min(64, seconds / max(one, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(rounds):
            pbs.pbs_batch(lwe, lut, threads)
        dt = time.perf_counter() - t0
    finally:
        lib.tfo_use_simd_transforms(0)
    return {"value": threads * rounds / dt, "unit": "PBS/s", "cores": threads, "kind": "port", "isa": CPU_ISA[0],
            "sample": "%d rounds x %d ciphertexts (one per thread), %.1f s" % (rounds, threads, dt),
            "config": dict(PBS_PARAMS, workload="programmable_bootstrap_ntt64_lwe_ciphertext (classic)")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 4096
    # size the sample so the whole run stays within a couple of minutes
    rate, dt = cpu_leg(sample, 1, threads)
    steps, warm = args.steps, args.warmup
    per_step_polys = int(max(256, min(BATCH_PER_GPU, rate * 1.0 / 2)))  # about 1 s of CPU work per step
    import oracle_lib
    lib = oracle_lib._load(native=True)
    plan = oracle_lib.OraclePlan(64, N, SOLINAS_P, _lib=lib)
    buf = synth(per_step_polys, 0xC0FFEE03)
    for _ in range(warm):
        plan.fwd_batch_inplace(buf, threads, simd=True)
        plan.inv_batch_inplace(buf, threads, simd=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.fwd_batch_inplace(buf, threads, simd=True)
        plan.inv_batch_inplace(buf, threads, simd=True)
    dt = time.perf_counter() - t0
    value = 2.0 * per_step_polys * steps / dt
    sample_txt = "%d polynomials per step (fwd then inv), %d threads, static contiguous chunks" % (per_step_polys, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt,
                         "isa": CPU_ISA[0], "note": CPU_NOTE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_pbs:
        try:
            line["pbs"] = pbs_cpu_leg(threads)
        except Exception as e:
            line["pbs"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import tfhe_ntt_b200 as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T.set_device(local)
    plan = T.prime64.Plan.try_new(N, SOLINAS_P)
    assert plan is not None
    stream = torch.cuda.current_stream()

    batch = BATCH_PER_GPU
    host = synth(4096, 0xC0FFEE03 + rank)
    d = torch.from_numpy(host.view(np.int64)).cuda().repeat(batch // 4096, 1).contiguous()
    # make the replicas distinct so no two polynomials are equal
    d[:, 0] = torch.arange(batch, device="cuda", dtype=torch.int64)
    torch.cuda.synchronize()

    def step():
        plan.fwd_device(d, batch, stream=stream)
        plan.inv_device(d, batch, stream=stream)
        plan.normalize_device(d, batch * N, stream=stream) if args.normalize else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # per-kernel timing for the roofline: CUDA events on the launching stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    with ClockSampler(local) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        ev[0].record(stream)
        for s in range(args.steps):
            plan.fwd_device(d, batch, stream=stream)
            ev[2 * s + 1].record(stream)
            plan.inv_device(d, batch, stream=stream)
            ev[2 * s + 2].record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    total_ms = ev[0].elapsed_time(ev[-1])
    fwd_ms = sum(ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps)) / args.steps
    inv_ms = sum(ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps)) / args.steps
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = 2.0 * batch * world * args.steps / (total_ms_max * 1e-3)

    # ---- e2e: host buffers through the public batch API, copies inside the timed region ----
    # The user-facing call for "forward, pointwise, inverse" on host data is the fused
    # prime64.Plan.fwd_mac_inv_batch (out = inv(fwd(lhs) * rhs), rhs = one resident GGSW row set):
    # per step the batch crosses PCIe once in each direction and undergoes 2 transforms/polynomial.
    eb = E2E_BATCH
    h_in = torch.from_numpy(synth(eb, 0xE2E + rank).view(np.int64)).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    ggsw = synth(8, 0x66 + rank)
    in_np, out_np = h_in.numpy().view(np.uint64), h_out.numpy().view(np.uint64)
    e2e_steps = max(2, min(args.steps, 5))
    plan.fwd_mac_inv_batch(out_np, in_np, ggsw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.fwd_mac_inv_batch(out_np, in_np, ggsw)   # H2D + fwd + mul + inv + D2H, pipelined
    barrier()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * eb * world * e2e_steps / float(te.item())
    e2e_bytes = eb * N * 8  # per step and direction
    chunk = 32 << 20  # staging chunk of the host pipeline (csrc/capi_prime.cu chunk_bytes())
    e2e_launches = e2e_steps * ((eb * N * 8 + chunk - 1) // chunk)

    # ---- sustained leg: the same step back to back for >= 2 s, clocks sampled throughout ----
    sustained = None
    if not args.no_sustained:
        per = max(args.steps, 20)
        with ClockSampler(local) as sclk:
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            done, t_s = 0, time.perf_counter()
            while time.perf_counter() - t_s < args.sustained_seconds:
                for _ in range(per):
                    plan.fwd_device(d, batch, stream=stream)
                    plan.inv_device(d, batch, stream=stream)
                done += per
                torch.cuda.synchronize()
            s1.record(stream)
            barrier()
        sus_ms = s0.elapsed_time(s1)
        ts = torch.tensor([sus_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sustained = {"seconds": float(ts.item()) * 1e-3, "steps": done, "value": 2.0 * batch * world * done / (float(ts.item()) * 1e-3),
                     "unit": UNIT, "ms_per_step": float(ts.item()) / done, "clocks": sclk.summary()}

    # ---- C1: the per-polynomial drop-in call (Plan::fwd on one host polynomial, N = 1024) next to one CPU core ----
    c1 = None
    if rank == 0 and not args.no_cpu:
        try:
            c1 = c1_latency_leg(T)
        except Exception as e:
            c1 = {"error": repr(e)}

    if rank == 0:
        peaks, which = measured_peaks()
        hbm = float(peaks["hbm_gbs"])
        dom_ms = max(fwd_ms, inv_ms)
        dom = "fwd" if fwd_ms >= inv_ms else "inv"
        achieved = batch * ALG_BYTES_PER_NTT / (dom_ms * 1e-3) / 1e9
        # the binding roofline: busy fraction of the two integer pipes, = pipe clocks the launch needs
        # (SASS mix) / pipe clocks available (SMs x 4 sub-partitions x duration x sampled SM clock)
        clk = clocks.summary()
        sm_hz = (clk.get("sm_mhz") or 1965.0) * 1e6
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        int_pipes = {}
        for name, ms in (("fwd", fwd_ms), ("inv", inv_ms)):
            avail = sms * 4 * ms * 1e-3 * sm_hz
            warps = batch / 2 * 8
            int_pipes[name] = {k: warps * v / avail for k, v in PIPE_CLK_PER_THREAD[name].items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(world),
            "roofline": {"bound": "hbm", "kernel": "ntt %s (N=2048 Solinas)" % dom, "achieved": achieved, "peak": hbm,
                         "unit": "GB/s", "frac": achieved / hbm, "traffic": NCU_TRAFFIC_PER_LAUNCH[dom], "peak_source": which,
                         "traffic_source": "profiles/r02_ncu_full_solinas2048_summary.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                         "limiter": "integer instruction throughput (ncu r02: ALU pipe 70-72 %, FMA-heavy pipe 71-74 %, issue 66 %, DRAM 28-30 %): "
                                    "the Solinas butterfly is carry-chain adds, see DESIGN.md section 5",
                         "int_pipe_busy_frac": int_pipes,
                         "int_pipe_note": "pipe clocks needed by the SASS instruction mix / pipe clocks available; ncu measures "
                                          "71-74 % (fmaheavy) and 70-72 % (alu), profiles/r02_ncu_full_solinas2048_summary.md; "
                                          "the register-resident butterfly loop reaches 83-85 % (profiles/r02_solinas_bf_variants.md)",
                         "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                         "algorithmic_bytes_per_launch": batch * ALG_BYTES_PER_NTT},
            # the binding roofline: the busier integer pipe of the dominant kernel against what the same butterfly
            # sustains register-resident (no memory, no barriers)
            "roofline_int": {"bound": "integer issue (ALU + FMA-heavy pipes)", "kernel": "ntt %s (N=2048 Solinas)" % dom,
                             "achieved": max(int_pipes[dom].values()), "peak": INT_PIPE_PEAK_FRAC,
                             "unit": "busy fraction of the busier pipe", "frac": max(int_pipes[dom].values()) / INT_PIPE_PEAK_FRAC,
                             "peak_source": "profiles/microbench/solinas_bf_variants.cu (V0, register resident): "
                                            "profiles/r02_solinas_bf_variants.md; pipe issue costs from profiles/microbench/int_pipe.cu"},
            "sustained": sustained,
            "c1_latency": c1,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": e2e_bytes,
                    "api": "prime64.Plan.fwd_mac_inv_batch (C ABI ntt_b200_plan64_fwd_mac_inv_batch) on pinned host "
                           "buffers: %d polynomials in, fwd + pointwise + inv, %d polynomials out per step" % (eb, eb),
                    "steps": e2e_steps, "kernel_launches": e2e_launches},
            "gpu_launches": 2 * args.steps,
            "clocks": clk,
            "wall_s": t_wall,
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            try:
                rate, dt = cpu_leg(2048, 1, threads)
                reps = int(max(1, min(4000, 12.0 / max(dt, 1e-3))))
                rate, dt = cpu_leg(2048, reps, threads)
                line["cpu_baseline"] = {
                    "value": rate, "unit": UNIT, "cores": threads, "kind": "port", "isa": CPU_ISA[0], "note": CPU_NOTE,
                    "sample": "%d x (fwd+inv over 2048 polynomials), %.1f s of CPU work, %d threads" % (reps, dt, threads)}
            except Exception as e:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"error": repr(e)}
    # ---- the consumer of the hot path: NTT-PBS, ciphertexts sharded over the ranks like the polynomials ----
    pbs = None
    if not args.no_pbs:
        try:
            pbs = pbs_leg(plan, torch, dist if world > 1 else None, world, barrier)
        except Exception as e:
            pbs = {"error": repr(e)}
    if rank == 0:
        if pbs is not None:
            if world == 1 and not args.no_cpu and "error" not in pbs:
                try:
                    pbs["cpu_baseline"] = pbs_cpu_leg(os.cpu_count() or 1)
                except Exception as e:
                    pbs["cpu_baseline"] = {"error": repr(e)}
            line["pbs"] = pbs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c1_latency_leg(T, n=1024, reps=400):
    """BASELINE config C1: Plan::fwd on ONE host polynomial (N = 1024, Solinas) through the C ABI mirror, wall clock per
    call, and the CPU port of the reference doing the same transform on one core."""
    import oracle_lib
    plan = T.prime64.Plan.try_new(n, SOLINAS_P)
    buf = (np.arange(n, dtype=np.uint64) * np.uint64(12345)) % np.uint64(SOLINAS_P)
    for _ in range(50):
        plan.fwd(buf)
    t0 = time.perf_counter()
    for _ in range(reps):
        plan.fwd(buf)
    gpu_us = (time.perf_counter() - t0) / reps * 1e6
    lib = oracle_lib._load(native=True)
    ref = oracle_lib.OraclePlan(64, n, SOLINAS_P, _lib=lib)
    isa = ref.fwd_batch_inplace(buf, 1, simd=True)
    t0 = time.perf_counter()
    for _ in range(4 * reps):
        ref.fwd_batch_inplace(buf, 1, simd=True)
    cpu_us = (time.perf_counter() - t0) / (4 * reps) * 1e6
    return {"gpu_us": gpu_us, "cpu_us": cpu_us, "cpu_isa": isa, "n": n,
            "call": "prime64::Plan::fwd on one host polynomial (ntt_b200_plan64_fwd; mapped pinned staging, one launch)",
            "note": "a single small call is launch- and PCIe-latency-bound; profiles/r02_latency.txt has the batch sizes from which the GPU path is ahead"}


def pbs_leg(plan, torch, dist, world, barrier, batch=888, reps=3, host_batch=1776):
    """Programmable bootstraps per second at the reference's parameter set, per GPU `batch` ciphertexts device
    resident (value) and `host_batch` ciphertexts through the host-buffer call
    programmable_bootstrap_ntt64_lwe_ciphertext (e2e: LWE in, LWE out, copies inside the timed region)."""
    from tfhe_ntt_b200 import ntt64_pbs as G
    n_lwe, gs, base_log, level = PBS_PARAMS["n_lwe"], PBS_PARAMS["k"] + 1, PBS_PARAMS["base_log"], PBS_PARAMS["level"]
    n = plan.ntt_size()
    bsk, lut_h, lwe_h = pbs_inputs(max(batch, host_batch))
    key = G.NttLweBootstrapKey.from_container(plan, bsk, n_lwe, gs, base_log, level)
    dev = torch.device("cuda", torch.cuda.current_device())
    lut = torch.from_numpy(lut_h.view(np.int64)).to(dev)
    lwe = torch.from_numpy(lwe_h[:batch].view(np.int64)).to(dev)
    acc = torch.empty((batch, gs * n), dtype=torch.int64, device=dev)
    out = torch.empty((batch, (gs - 1) * n + 1), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()

    def step():
        G.blind_rotate_ntt64_device(key, lwe, lut, 1, acc, batch, stream=st)
        G.extract_lwe_sample_device(key, acc, out, batch, stream=st)

    for _ in range(2):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    # host-buffer call
    lwe_in = np.ascontiguousarray(lwe_h[:host_batch]).reshape(-1)
    lwe_out = np.zeros(host_batch * ((gs - 1) * n + 1), dtype=np.uint64)
    G.programmable_bootstrap_ntt64_lwe_ciphertext(lwe_in, lwe_out, lut_h, key)
    barrier()
    t0 = time.perf_counter()
    G.programmable_bootstrap_ntt64_lwe_ciphertext(lwe_in, lwe_out, lut_h, key)
    barrier()
    host_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([ms, host_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, host_s = float(t[0].item()), float(t[1].item())
    return {"value": batch * world / ms * 1e3, "unit": "PBS/s", "ms_per_batch": ms, "n_gpus": world,
            "config": dict(PBS_PARAMS, workload="programmable_bootstrap_ntt64 (classic), persistent two-CTA-cluster blind "
                                                "rotation + sample extraction; ciphertexts sharded over the GPUs, no collective",
                           batch_per_gpu=batch),
            "ntt_per_s_inside": batch * world * n_lwe * 4 / ms * 1e3,
            "e2e": {"value": host_batch * world / host_s, "unit": "PBS/s", "batch_per_gpu": host_batch,
                    "h2d_bytes_per_step": int(lwe_in.nbytes + lut_h.nbytes), "d2h_bytes_per_step": int(lwe_out.nbytes),
                    "api": "ntt64_pbs.programmable_bootstrap_ntt64_lwe_ciphertext (C ABI ntt_b200_programmable_bootstrap_ntt64) on host buffers"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--normalize", action="store_true", help="also run normalize in warm-up steps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pbs", action="store_true", help="skip the informational NTT-PBS leg")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s back-to-back leg")
    ap.add_argument("--sustained-seconds", type=float, default=2.2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
