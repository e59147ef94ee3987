"""Aggregate host<->device copy ceiling of the box when several GPUs copy at the same time (one process
per GPU under torchrun, pinned memory, plain cudaMemcpyAsync, no kernels of ours).  The end-to-end number
of `bench.py --gpus N` is read against this: the host side of the box (root complex / host DRAM), not the
per-GPU PCIe link, bounds it.  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/pcie_peak_multi.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 29  # 512 MiB each way per GPU
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(mode, reps=6):
        def once():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        per_dir = n * reps * world / float(dt.item()) / 1e9
        return per_dir * (2 if mode == "both" else 1)

    out = {"n_gpus": world, "bytes_each_way_per_gpu": n}
    for mode in ("h2d", "d2h", "both"):
        out[mode + "_aggregate_GBps"] = run(mode)
    out["ntt_per_s_ceiling_at_16KiB_each_way"] = out["both_aggregate_GBps"] / 2 * 1e9 / 16384 * 2
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
