"""Does running the strided TMA pass of one chunk of long polynomials beside the single-CTA kernels of another chunk
(two streams) beat the one-stream sequence?  The pass is HBM-bound, the sub-block kernels are integer-issue-bound."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import tfhe_ntt_b200 as T

P = T.prime64.SOLINAS_PRIME


def bench(n, batch, chunks, streams, reps=20, inverse=False):
    plan = T.prime64.Plan.try_new(n, P)
    x = torch.randint(0, (1 << 62), (batch, n), dtype=torch.int64, device="cuda")
    sts = [torch.cuda.Stream() for _ in range(streams)]
    per = batch // chunks
    call = plan.inv_device if inverse else plan.fwd_device

    def step():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        for s in sts:
            s.wait_event(ev)
        for c in range(chunks):
            s = sts[c % streams]
            call(x[c * per:(c + 1) * per], per, stream=s)
        for s in sts:
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n, batch in ((65536, 1024), (16384, 4096), (8192, 8192)):
    for inverse in (False, True):
        base = bench(n, batch, 1, 1, inverse=inverse)
        line = "n=%d batch=%d %s: one call %.3f ms" % (n, batch, "inv" if inverse else "fwd", base)
        for chunks, streams in ((2, 2), (4, 2), (8, 2), (4, 4), (8, 4), (16, 4), (8, 1)):
            t = bench(n, batch, chunks, streams, inverse=inverse)
            line += " | %dch/%dst %.3f" % (chunks, streams, t)
        print(line, flush=True)
