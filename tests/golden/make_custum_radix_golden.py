"""Generates tests/golden/custum_radix_golden.json: frozen input/output vectors of the fork's custum_radix transforms.

The module has no known-answer test of its own (its only test, custum_radix/fwd_1.rs:433-463, prints), and the crate
cannot be built here (no rustc), so the outputs are computed from the DEFINITION with Python integers -- no oracle, no
CUDA: with g the smallest generator of (Z/p)^* (fwd.rs:42-68), root = g^((p-1)/n) and tw[k] = root^k (fwd.rs:72-93),

    fft_*_recursive(a, tw, p)[k]                 = sum_j a[j] root^(jk)                      (fwd.rs:105-272)
    ifft_radix2 / ifft_split_radix(.., top)[k]   = n^-1 * sum_j a[j] root^(-jk)              (inv.rs:178-303)
    ifft_radix4(.., top)[k]                      = the same, times 1/2 when log2 n is odd    (inv.rs:106-176)

The first case is the input of that test (n = 64, p = 65537).  tests/test_custum_radix.py requires the oracle (CPU) and the
CUDA path (GPU) to reproduce every vector bit for bit.

    python tests/golden/make_custum_radix_golden.py        # rewrites the file (deterministic)
"""
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))

REFERENCE_TEST_INPUT = [5, 11, 3, 12, 8, 13, 2, 14, 4, 15, 7, 16, 6, 17, 1, 18, 3, 19, 9, 20, 2, 21, 5, 22,
                        7, 23, 4, 24, 1, 25, 8, 26, 9, 27, 6, 28, 3, 29, 5, 30, 2, 31, 8, 32, 4, 33, 7, 34,
                        1, 35, 6, 36, 3, 37, 9, 38, 2, 39, 5, 40, 7, 41, 4, 42]


def smallest_generator(p):
    m, factors, i = p - 1, [], 2
    while i * i <= m:
        if m % i == 0:
            factors.append(i)
            while m % i == 0:
                m //= i
        i += 1
    if m > 1:
        factors.append(m)
    return next(g for g in range(2, p) if all(pow(g, (p - 1) // f, p) != 1 for f in factors))


def dft(a, root, p):
    n = len(a)
    pw = [pow(root, e, p) for e in range(n)]
    return [sum(a[j] * pw[(j * k) % n] for j in range(n)) % p for k in range(n)]


def case(name, a, p):
    n = len(a)
    root = pow(smallest_generator(p), (p - 1) // n, p)
    n_inv, half = pow(n, p - 2, p), pow(2, p - 2, p)
    inv = dft(a, pow(root, p - 2, p), p)
    odd = (n.bit_length() - 1) % 2 == 1
    return {"name": name, "n": n, "p": p, "root": root, "input": a, "fft": dft(a, root, p),
            "ifft_radix2_top": [x * n_inv % p for x in inv],
            "ifft_radix4_top": [x * n_inv * (half if odd else 1) % p for x in inv]}


def main():
    rng = random.Random(0xC0FFEE)
    cases = [case("reference_test_input_fwd_1.rs:447", REFERENCE_TEST_INPUT, 65537)]
    for n, p in ((8, 17), (32, 2013265921), (128, 4293918721), (256, 65537)):
        a = [rng.randrange(p) for _ in range(n)]
        a[3] = 0
        a[5] = p - 1
        cases.append(case("random_n%d_p%d" % (n, p), a, p))
    with open(os.path.join(HERE, "custum_radix_golden.json"), "w") as f:
        json.dump({"cases": cases}, f, separators=(",", ":"))
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
