"""TEST INFRASTRUCTURE ONLY -- aotools.fouriertransform restated (see ../__init__.py).
Centred DFT pairs scaled by the grid spacing.  Used by fast/funcs.py:218,309,373 and
fast/fast.py:745-754."""
from numpy import fft as _fft


def ft(data, delta):
    return _fft.fftshift(_fft.fft(_fft.fftshift(data, axes=-1)), axes=-1) * delta


def ift(DATA, delta_f):
    n = DATA.shape[-1]
    return _fft.ifftshift(_fft.ifft(_fft.ifftshift(DATA, axes=-1)), axes=-1) * n * delta_f


def ft2(data, delta):
    ax = (-1, -2)
    return _fft.fftshift(_fft.fft2(_fft.fftshift(data, axes=ax)), axes=ax) * delta ** 2


def ift2(DATA, delta_f):
    ax = (-1, -2)
    n = DATA.shape[-1]
    return _fft.ifftshift(_fft.ifft2(_fft.ifftshift(DATA, axes=ax)), axes=ax) * (n * delta_f) ** 2
