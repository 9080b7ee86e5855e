"""Runs the CPU emulation of the device line FFT (tests/host/host_fft_emul.cu): the exact
per-thread phases of fast_b200/csrc/fft_core.cuh against a float64 DFT for N = 64..2048, and the
constexpr register masks of the window-specialised kernel instances."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc not available')
def test_line_fft_index_algebra(tmp_path):
    exe = str(tmp_path / 'host_fft_emul')
    subprocess.run(['nvcc', '-O1', '-std=c++17', '--expt-relaxed-constexpr', '-o', exe, os.path.join(ROOT, 'tests', 'host', 'host_fft_emul.cu')],
                   check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count('max_err') == 17
    # compile-time output pruning: keep_mask<F>() equals a brute-force scan for 6 sizes x 3 window classes
    assert out.stdout.count('keeps') == 18 and 'MISMATCH' not in out.stdout


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc not available')
def test_chirp_z_line_algebra(tmp_path):
    """tests/host/host_chirpz_emul.cu: one line of the chirp-z K2 kernel on the CPU -- shifted kernel, Bhat in register
    order, chaining of the two transforms (in registers at M = 256, through the line buffer otherwise), the
    class invariants behind the compile-time pruning -- against a float64 DFT, M = 64 .. 2048, classes 5 .. 8."""
    exe = str(tmp_path / 'host_chirpz_emul')
    subprocess.run(['nvcc', '-O1', '-std=c++17', '--expt-relaxed-constexpr', '-o', exe,
                    os.path.join(ROOT, 'tests', 'host', 'host_chirpz_emul.cu')], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count('max_err') == 13
    assert 'N=164 lo=41 P=82 M=256 C=6 keep=0x003f identity=1' in out.stdout


def test_noise_stride_rule_matches_the_oracle_restatement():
    """fastb_noise_stride (host function of the library, no GPU needed) against oracle.fast_oracle.noise_stride over
    every even grid up to 2300 and a spread of crops: N / 16 for the radix sizes, M / 16 where the chirp-z kernel
    applies, ceil(N / 16) beyond."""
    import sys
    sys.path.insert(0, ROOT)
    from fast_b200 import _lib
    from oracle import fast_oracle as fo
    for n in list(range(4, 700, 2)) + list(range(700, 2300, 38)):
        for n_pup in sorted({1, 2, n // 6 + 1, n // 3, n // 2, n - 1, n}):
            if n_pup < 1:
                continue
            assert _lib.noise_stride(n, n_pup) == fo.noise_stride(n, n_pup), (n, n_pup)
            assert 16 * _lib.noise_stride(n, n_pup) >= n
    assert _lib.noise_stride(256, 82) == 16 and _lib.noise_stride(164, 82) == 16 and _lib.noise_stride(2100, 100) == 132
    with pytest.raises(_lib.FastbError):
        _lib.noise_stride(163, 10)
