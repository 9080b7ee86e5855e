"""TEST INFRASTRUCTURE ONLY -- aotools.functions.zernike.zernIndex (Noll index -> [n, m]).
Used by fast/ao_power_spectra.py:4,11."""
import numpy as _np


def zernIndex(j):
    n = int((-1.0 + _np.sqrt(8 * (j - 1) + 1)) / 2.0)
    p = j - (n * (n + 1)) / 2.0
    k = n % 2
    m = int((p + k) / 2.0) * 2 - k
    if m != 0:
        m *= 1 if j % 2 == 0 else -1
    return [n, m]
