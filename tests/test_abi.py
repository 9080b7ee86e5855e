"""CPU checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/fastb.h declares, argument validation works without a GPU, and the product path fails
loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import build_fastb
    build_fastb.build()
    from fast_b200 import _lib
    return _lib


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'fastb.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(fastb_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib.lib, s), s
    assert sorted(lib.EXPORTED) == syms


def test_struct_layout_matches_header(lib):
    # sizes follow from the C declarations: 8 int32 + 11 double + 4*32 double
    assert ctypes.sizeof(lib.PsdParams) == 8 * 4 + 11 * 8 + 4 * 32 * 8
    assert ctypes.sizeof(lib.RunParams) == 6 * 4 + 3 * 8 + 8 + 8 + 4 + 4
    assert ctypes.sizeof(lib.PsdOutputs) == 10 * 8 and ctypes.sizeof(lib.PsdInputs) == 3 * 8
    assert ctypes.sizeof(lib.RunBatch) == 2 * 4 + 8 + 2 * 8
    assert ctypes.sizeof(lib.RunStats) == 2 * 8 + 2 * 4 + 3 * 8
    assert lib.RunParams.flags.offset == 20 and lib.RunParams.n_pairs.offset == 24


def test_run_argument_validation_without_gpu(lib):
    """Validation happens before any CUDA call: bad flags, algorithms and batch geometry are
    refused with a message, on a box without a GPU too."""
    rp = lib.RunParams()
    rp.n, rp.n_pup, rp.lo, rp.n_pairs, rp.pairs_per_chunk, rp.u_sum = 64, 8, 28, 1, 1, 1.0
    rp.flags = 64
    assert lib.lib.fastb_screen_detect_batch_workspace_bytes(ctypes.byref(rp), 1) == -1
    assert b'flags' in lib.lib.fastb_last_error()
    rp.flags, rp.algo = 0, lib.ALGO_BLUESTEIN + 1
    assert lib.lib.fastb_screen_detect_batch_workspace_bytes(ctypes.byref(rp), 1) == -1
    assert b'algo' in lib.lib.fastb_last_error()
    rp.algo, rp.n, rp.n_pup, rp.lo = lib.ALGO_BLUESTEIN, 2000, 200, 900
    assert lib.lib.fastb_screen_detect_batch_workspace_bytes(ctypes.byref(rp), 1) == -1
    assert b'chirp-z' in lib.lib.fastb_last_error()


def test_version_and_error_text(lib):
    assert lib.version() == 200
    rp = lib.RunParams()
    rp.n, rp.n_pup, rp.lo, rp.n_pairs, rp.pairs_per_chunk, rp.u_sum = 63, 8, 0, 1, 1, 1.0
    assert lib.lib.fastb_screen_detect_workspace_bytes(ctypes.byref(rp)) == -1
    assert b'even' in lib.lib.fastb_last_error()
    assert lib.lib.fastb_psd_build(None, None, None, None) == 1       # FASTB_ERR_ARG


@pytest.mark.skipif(__import__('torch').cuda.is_available(), reason='CPU-only check')
def test_product_path_fails_loudly_without_gpu(lib):
    import fast_b200
    assert lib.lib.fastb_device_count() == 0
    with pytest.raises(lib.FastbError, match='no CPU fallback'):
        fast_b200.Fast({'NITER': 4, 'NCHUNKS': 1, 'LOGLEVEL': 'ERROR'})
    import torch
    with pytest.raises(lib.FastbError, match='CUDA tensor'):
        lib.make_weight(torch.zeros(4, 4, dtype=torch.float64), 1.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'fast_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'fast_oracle' not in text, f
