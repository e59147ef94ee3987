"""Measures the host<->device copy ceiling of the box (pinned memory, large transfers) so that the
end-to-end number of bench.py can be read against it.  Prints one JSON line."""
import json

import torch


def bw(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    n = 1 << 29  # 512 MiB each way
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2d = bw(lambda: d_in.copy_(h_in, non_blocking=True), n)
    d2h = bw(lambda: h_out.copy_(d_out, non_blocking=True), n)

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    bi = bw(both, 2 * n)
    print(json.dumps({"h2d_GBps": h2d, "d2h_GBps": d2h, "bidirectional_total_GBps": bi,
                      "ntt_per_s_ceiling_at_16KiB_each_way": bi / 2 * 1e9 / 16384 * 2}))


if __name__ == "__main__":
    main()
