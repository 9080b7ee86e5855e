"""Reference distribution for the statistical parity check (TEST INFRASTRUCTURE).

Runs the UNMODIFIED reference on the C2 configuration (N=256, SMF detection) in 8 processes of
12 500 realisations each with distinct seeds, and stores the 1e5 per-realisation values `_r`
(float32) in tests/golden/c2_dist_1e5.npz.  The CUDA path with device RNG must reproduce this
distribution (mean / variance of dB_rel, KS test) -- tests/test_gpu_statistics.py.
"""
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SEEDS = [101, 102, 103, 104, 105, 106, 107, 108]
PER = 12500


def one(seed):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, '/root/reference')
    sys.path.insert(0, os.path.join(HERE, 'shim'))
    import fast
    from oracle import configs
    p = configs.c2(niter=PER, nchunks=50, seed=seed)
    sim = fast.Fast(p)
    return sim.run()._r


if __name__ == '__main__':
    import numpy as np
    with Pool(len(SEEDS)) as pool:
        parts = pool.map(one, SEEDS)
    r = np.concatenate(parts).astype(np.float32)
    out = os.path.join(ROOT, 'tests', 'golden', 'c2_dist_1e5.npz')
    np.savez_compressed(out, r=r, seeds=np.array(SEEDS), per_seed=np.int64(PER))
    db = 10 * np.log10(r.astype(float))
    print(out, r.size, db.mean(), db.var(), os.path.getsize(out))
