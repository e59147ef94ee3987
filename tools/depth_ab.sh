# A/B of the number of strided top stages for long polynomials (tools/large_n_bench.py)
for d in 1 2 3 4 5; do NTT_B200_DEPTH=$d timeout 300 python tools/large_n_bench.py 2>&1; done
NTT_B200_CLUSTER=1 timeout 300 python tools/large_n_bench.py 2>&1 | grep "u64 n=8192\|u64 n=16384"
timeout 300 python tools/large_n_bench.py 2>&1
