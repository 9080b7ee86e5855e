"""Throughput of the chirp-z K2 kernels over transform lengths and cell classes (synthetic weights, device RNG):
one warm-up and one timed launch per (N, lo, P).  python tools/chirpz_sizes.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_b200 import _lib as lib

CASES = [(100, 35, 30), (164, 41, 82), (200, 78, 44), (236, 108, 20), (104, 40, 24), (58, 26, 6),
         (300, 60, 180), (330, 110, 110), (460, 210, 50), (700, 250, 200), (900, 400, 120), (1500, 500, 549)]
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = torch.device('cuda')
for N, lo, P in CASES:
    gen = torch.Generator(device='cuda').manual_seed(N)
    weight = lib.make_weight(torch.rand(N, N, dtype=torch.float64, device=dev, generator=gen) * 1e-5, 1.5)
    U = torch.rand(P, P, dtype=torch.float32, device=dev, generator=gen)
    n = max(2000, int(pairs * (164 * 164) / (N * N)))
    rp = lib.RunParams()
    rp.n, rp.n_pup, rp.lo, rp.n_pairs, rp.pairs_per_chunk, rp.seed, rp.algo = N, P, lo, n, n, 17, lib.ALGO_AUTO
    rp.u_sum, rp.sigma_chi = float(U.sum()), 0.02
    ws = torch.empty(lib.screen_detect_workspace_bytes(rp), dtype=torch.uint8, device=dev)
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)
    lib.screen_detect(rp, weight, U, a, b, ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.screen_detect(rp, weight, U, a, b, ws); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    M = 64
    while M < N + P - 1:
        M *= 2
    print(f'N={N:5d} P={P:4d} M={M:5d} stride={lib.noise_stride(N, P):4d}  {n:7d} pairs {ms:9.3f} ms  '
          f'{2 * n / ms / 1e3:8.3f} M realisations/s  {2 * n * (8 * N * N + 4) / ms / 1e6:8.1f} GB/s model A', flush=True)
