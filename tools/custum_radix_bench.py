"""custum_radix transforms on device-resident batches: vectors per second and the fraction of the measured HBM
copy peak (algorithmic bytes = 2 * n * 4 per transform), with the oracle's recursive radix-2 routine -- the
reference's formulation -- timed on one host core beside it.  Developer tool; CUDA events."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import tfhe_ntt_b200.custum_radix as cr

HBM = 6543.4


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cpu_rate(n, p, tw):
    import oracle_lib as O
    L = O.lib()
    L.tfo_cr_fft_radix2_recursive.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]
    L.tfo_cr_fft_radix2_recursive.restype = None
    a = np.random.default_rng(1).integers(0, p, size=n, dtype=np.uint64).astype(np.uint32)
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 1.0:
        L.tfo_cr_fft_radix2_recursive(a.ctypes.data, n, tw.ctypes.data, p)
        reps += 1
    return reps / (time.perf_counter() - t0)


def main():
    st = torch.cuda.current_stream()
    for n, p in [(64, 65537), (1024, 2013265921), (2048, 2013265921), (4096, 2013265921), (32768, 2013265921),
                 (1 << 17, 2013265921)]:
        batch = (1 << 29) // (4 * n)  # 512 MiB, far larger than L2
        tw = cr.make_twiddles(n, p)
        inv = cr.make_inv_twiddles(tw, p)
        d = torch.randint(0, p, (batch, n), dtype=torch.int64, device="cuda").to(torch.int32)
        d_tw = torch.from_numpy(tw.view(np.int32)).cuda()
        d_inv = torch.from_numpy(inv.view(np.int32)).cuda()
        n_inv = pow(n, p - 2, p)
        tf = timeit(lambda: cr.fft_device(cr.RADIX2, d, n, batch, d_tw, p, stream=st))
        ti = timeit(lambda: cr.ifft_device(cr.RADIX2, d, n, batch, d_inv, p, n_inv, True, stream=st))
        cpu = cpu_rate(n, p, tw)
        for name, t in (("fft ", tf), ("ifft", ti)):
            rate = batch / t * 1e3
            print("n=%-7d p=%-11d %s  %.3f ms  %8.2f M transforms/s  %5.1f %% of HBM peak   (oracle recursion, 1 core: "
                  "%.1f K/s)" % (n, p, name, t, rate / 1e6, rate * 8 * n / 1e9 / HBM * 100, cpu / 1e3), flush=True)
        del d


def stats_bench():
    """the `_mut` routines with MultStats: host arrays in, values and counters out (one launch per 32 MiB chunk)"""
    import oracle_lib as O
    L = O.lib()

    class MS(C.Structure):
        _fields_ = [("nz", C.c_size_t), ("sk", C.c_size_t)]
    L.tfo_cr_fft_split_radix_recursive_mut.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.POINTER(MS)]
    L.tfo_cr_fft_split_radix_recursive_mut.restype = None
    p = 2013265921
    for n, batch in [(64, 65536), (1024, 8192), (4096, 2048)]:
        tw = cr.make_twiddles(n, p)
        a = np.random.default_rng(2).integers(0, p, size=(batch, n), dtype=np.uint64).astype(np.uint32)
        cr.fft_mut_batch(cr.SPLIT_RADIX, a.copy(), tw, p)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            cr.fft_mut_batch(cr.SPLIT_RADIX, a.copy(), tw, p)
        gpu = reps * batch / (time.perf_counter() - t0)
        v, st, k, t0 = a[0].copy(), MS(), 0, time.perf_counter()
        while time.perf_counter() - t0 < 1.0:
            L.tfo_cr_fft_split_radix_recursive_mut(v.ctypes.data, n, tw.ctypes.data, p, C.byref(st))
            k += 1
        cpu = k / (time.perf_counter() - t0)
        print("split-radix _mut with MultStats, n=%-5d batch=%-6d host call: %8.1f K vectors/s   (oracle recursion, 1 core: "
              "%.1f K/s)" % (n, batch, gpu / 1e3, cpu / 1e3), flush=True)


if __name__ == "__main__":
    main()
    stats_bench()
