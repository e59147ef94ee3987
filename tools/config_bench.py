"""Device-resident throughput of the BASELINE.json configs C2..C5 (developer tool; the contract
bench is bench.py).  CUDA events, inputs larger than L2."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfhe_ntt_b200 as T

HBM = 6543.4e9


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
    out = {}
    # C2: prime32 N=2048 30-bit prime, batch 65536, fwd + mul_accumulate + inv (fused)
    n, p, batch = 2048, 1073479681, 65536
    plan = T.prime32.Plan.try_new(n, p)
    lhs, rhs, acc = (torch.randint(0, p, (batch, n), dtype=torch.int64, device="cuda").to(torch.int32) for _ in range(3))
    o = torch.empty_like(lhs)
    t = timeit(lambda: plan.fwd_mac_inv_device(o, lhs, rhs, acc, stream=st))
    out["C2"] = {"units_per_s": batch / t, "ms": t * 1e3, "algorithmic_bytes_per_unit": 4 * n * 4,
                 "hbm_frac": batch * 4 * n * 4 / t / HBM, "ntt_per_s": 2 * batch / t}
    def unfused():
        x = lhs.clone()
        plan.fwd_device(x, stream=st)
        a = acc.clone()
        plan.mul_accumulate_device(a, x, rhs, stream=st)
        plan.inv_device(a, stream=st)
    t2 = timeit(unfused)
    out["C2"]["unfused_ms_incl_2_clones"] = t2 * 1e3
    del lhs, rhs, acc, o
    # C3: prime64 Solinas N=2048, k=1, l=2, 4096 LWEs: 16384 fwd, 8 shared GGSW polys, 8192 inv
    n, p, lwes = 2048, T.prime64.SOLINAS_PRIME, 4096
    plan = T.prime64.Plan.try_new(n, p)
    dig = torch.randint(0, 1 << 62, (lwes, 4, n), dtype=torch.int64, device="cuda")
    ggsw = torch.randint(0, 1 << 62, (4, 2, n), dtype=torch.int64, device="cuda")
    accs = torch.zeros((2, lwes, n), dtype=torch.int64, device="cuda")
    work = torch.empty_like(dig)
    rows = [torch.empty((lwes, n), dtype=torch.int64, device="cuda") for _ in range(4)]
    def c3_step():
        work.copy_(dig)
        plan.fwd_device(work, lwes * 4, stream=st)
        accs.zero_()
        for r in range(4):
            rows[r].copy_(work[:, r, :])
            for c in range(2):
                plan.mul_accumulate_device(accs[c], rows[r], ggsw[r, c].contiguous(), stream=st)
        plan.inv_device(accs, lwes * 2, stream=st)
    t = timeit(c3_step)
    out["C3"] = {"external_products_per_s": lwes / t, "ms": t * 1e3, "ntt_per_s": (lwes * 6) / t,
                 "note": "unfused: copy + 16384 fwd + 8 shared-GGSW mul_accumulate passes + 8192 inv"}
    outp = torch.empty((lwes, 2, n), dtype=torch.int64, device="cuda")
    t_e = timeit(lambda: plan.ext_product_device(outp, dig, ggsw, 4, 2, stream=st))
    out["C3"]["fused_ext_product_ms"] = t_e * 1e3
    out["C3"]["fused_external_products_per_s"] = lwes / t_e
    out["C3"]["fused_ntt_per_s"] = lwes * 6 / t_e
    out["C3"]["fused_hbm_frac"] = lwes * 6 * n * 8 / t_e / HBM
    t_f = timeit(lambda: plan.fwd_device(work, lwes * 4, stream=st))
    t_i = timeit(lambda: plan.inv_device(accs, lwes * 2, stream=st))
    out["C3"]["fwd_only_ms"] = t_f * 1e3
    out["C3"]["inv_only_ms"] = t_i * 1e3
    out["C3"]["ntt_only_hbm_frac"] = (lwes * 6 * 2 * n * 8) / (t_f + t_i) / HBM
    del dig, ggsw, accs, work, rows
    # C4: native64::Plan32 negacyclic_polymul N=4096 batch 16384
    n, batch = 4096, 16384
    plan = T.native64.Plan32.try_new(n)
    lhs = torch.randint(-(1 << 63), (1 << 63) - 1, (batch, n), dtype=torch.int64, device="cuda")
    rhs = torch.randint(-(1 << 63), (1 << 63) - 1, (batch, n), dtype=torch.int64, device="cuda")
    prod = torch.empty_like(lhs)
    t = timeit(lambda: plan.negacyclic_polymul_device(prod, lhs, rhs, stream=st), iters=3, warm=1)
    out["C4"] = {"products_per_s": batch / t, "ms": t * 1e3, "algorithmic_bytes_per_unit": 3 * n * 8,
                 "hbm_frac": batch * 3 * n * 8 / t / HBM, "residue_ntt_per_s": 15 * batch / t}
    del lhs, rhs, prod
    # C5: prime64 Solinas N=65536 batch 1024
    n, batch = 65536, 1024
    plan = T.prime64.Plan.try_new(n, T.prime64.SOLINAS_PRIME)
    x = torch.randint(0, 1 << 62, (batch, n), dtype=torch.int64, device="cuda")
    tf = timeit(lambda: plan.fwd_device(x, batch, stream=st))
    ti = timeit(lambda: plan.inv_device(x, batch, stream=st))
    out["C5"] = {"fwd_ntt_per_s": batch / tf, "inv_ntt_per_s": batch / ti, "fwd_ms": tf * 1e3, "inv_ms": ti * 1e3,
                 "algorithmic_bytes_per_unit": 2 * n * 8, "hbm_frac_fwd": batch * 2 * n * 8 / tf / HBM,
                 "hbm_frac_inv": batch * 2 * n * 8 / ti / HBM, "passes_over_hbm": 2}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
