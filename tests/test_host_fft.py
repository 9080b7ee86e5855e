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
