"""Reference distributions for the statistical parity checks (TEST INFRASTRUCTURE).

    python oracle/make_golden_dist.py [c2|c3_el10|c4|c5|c1prime]

Runs the UNMODIFIED reference in 8 processes with distinct seeds and stores the per-realisation
values `_r` in tests/golden/<case>_dist_*.npz (c2: 1e5 samples, float32; c3_el10: 5e4; c4: 5e4
complex64; c5: 1e4; c1prime -- the auto-sized 164 x 164 grid of the reference's example, TEMPORAL off: 1e5).  The CUDA path with device RNG must reproduce these distributions (mean /
variance of dB_rel, KS test) -- tests/test_gpu_statistics.py.
"""
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SEEDS = [101, 102, 103, 104, 105, 106, 107, 108]
# case -> (config factory, kwargs, realisations per process, chunks, output file)
CASES = {
    'c2': ('c2', {}, 12500, 50, 'c2_dist_1e5.npz'),
    'c3_el10': ('c3_elevation', {'el_deg': 10.0}, 6250, 25, 'c3_el10_dist_5e4.npz'),
    'c4': ('c4', {}, 6250, 125, 'c4_dist_5e4.npz'),
    'c5': ('c5', {}, 1250, 125, 'c5_dist_1e4.npz'),
    'c1prime': ('c1prime', {}, 12500, 50, 'c1prime_dist_1e5.npz'),
}


def one(args):
    case, seed = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, '/root/reference')
    sys.path.insert(0, os.path.join(HERE, 'shim'))
    import fast
    from oracle import configs
    factory, kw, per, chunks, _ = CASES[case]
    p = getattr(configs, factory)(niter=per, nchunks=chunks, seed=seed, **kw)
    sim = fast.Fast(p)
    return sim.run()._r


if __name__ == '__main__':
    import numpy as np
    for case in (sys.argv[1:] or ['c2']):
        _, _, per, _, fname = CASES[case]
        with Pool(len(SEEDS)) as pool:
            parts = pool.map(one, [(case, s) for s in SEEDS])
        r = np.concatenate(parts)
        r = r.astype(np.complex64) if np.iscomplexobj(r) else r.astype(np.float32)
        out = os.path.join(ROOT, 'tests', 'golden', fname)
        np.savez_compressed(out, r=r, seeds=np.array(SEEDS), per_seed=np.int64(per))
        db = 10 * np.log10(np.abs(r.astype(complex)) ** 2 if np.iscomplexobj(r) else r.astype(float))
        print(out, r.size, db.mean(), db.var(), os.path.getsize(out), flush=True)
