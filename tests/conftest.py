import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden_arrays(name):
    """A golden npz without a config (tests/golden/comms.npz)."""
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def load_golden(name):
    """Golden npz written by oracle/make_golden.py + the params dict that produced it."""
    from oracle import configs
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))
    factory = str(g['case_factory'])
    kwargs = ast.literal_eval(str(g['case_kwargs']))
    params = getattr(configs, factory)(**kwargs)
    return g, params


def reference_noise_stream(seed, niter, nchunks, N):
    """Replays the reference's RNG order (fast/funcs.py:352-365): log-amp normals (niter real,
    niter discarded imaginary), then per chunk a real (J/2,N,N) block and an imaginary one.
    Yields ('chi', array) once, then ('noise', chunk, complex array) per chunk."""
    rng = np.random.default_rng(seed)
    a = rng.normal(0, 1, size=(niter,))
    rng.normal(0, 1, size=(niter,))
    yield ('chi', a)
    J = niter // nchunks
    for c in range(nchunks):
        re = rng.normal(0, 1, size=(J // 2, N, N))
        im = rng.normal(0, 1, size=(J // 2, N, N))
        yield ('noise', c, re + 1j * im)


@pytest.fixture(scope='session')
def golden():
    return load_golden
