"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party `aotools` package.

The reference (ojdf/fast) depends on aotools>=1.0.7 (requirements.txt:3), which is
not installable in this image (no network).  This shim restates the published
behaviour of the handful of aotools functions the reference calls, so that the
UNMODIFIED reference under /root/reference can be imported in the build container
to generate golden vectors (oracle/make_golden.py).  It is never imported by the
product package `fast_b200`.

Call sites in the reference: fast/fast.py:5, fast/funcs.py:8, fast/comms.py:8,
fast/ao_power_spectra.py:4-5.
"""
import numpy as _np
from . import fouriertransform          # noqa: F401
from .functions import zernike          # noqa: F401


def circle(radius, size, circle_centre=(0, 0), origin="middle"):
    # aotools.functions.pupil.circle: pixel centres sit at i + 0.5 - size/2, so the
    # disc is centred between pixels size/2-1 and size/2.
    c = _np.arange(size, dtype=float) + 0.5
    xx, yy = _np.meshgrid(c, c)
    if origin == "middle":
        xx = xx - size / 2.0
        yy = yy - size / 2.0
    xx = xx - circle_centre[0]
    yy = yy - circle_centre[1]
    out = _np.zeros((size, size))
    out[xx ** 2 + yy ** 2 <= radius ** 2] = 1
    return out


def gaussian2d(size, width, amplitude=1.0, cent=None):
    # aotools.functions._functions.gaussian2d: (y, x) ordering for size/width/cent,
    # centred ON pixel size/2 (half a pixel away from `circle`'s centre).
    if _np.ndim(size) == 0:
        ny = nx = int(size)
    else:
        ny, nx = int(size[0]), int(size[1])
    if _np.ndim(width) == 0:
        wy = wx = float(width)
    else:
        wy, wx = float(width[0]), float(width[1])
    if not cent:
        cy, cx = ny / 2.0, nx / 2.0
    else:
        cy, cx = cent[0], cent[1]
    X, Y = _np.meshgrid(_np.arange(nx), _np.arange(ny))
    return amplitude * _np.exp(-(((cx - X) / wx) ** 2 + ((cy - Y) / wy) ** 2) / 2)


def cn2_to_r0(cn2, lamda=500.e-9):
    return (0.423 * (2 * _np.pi / lamda) ** 2 * cn2) ** (-3.0 / 5.0)


def isoplanaticAngle(cn2, h, lamda=500.e-9):
    # aotools.turbulence.atmos_conversions: returns ARCSECONDS (restated from memory like the rest of this
    # shim; unverifiable offline -- the value only feeds the `theta0` attribute / FITS header)
    Jsum = (cn2 * (h ** (5.0 / 3.0))).sum()
    return 0.057 * lamda ** (6.0 / 5.0) * Jsum ** (-3.0 / 5.0) * 180.0 * 3600.0 / _np.pi


def coherenceTime(cn2, v, lamda=500.e-9):
    return float(_np.sum(cn2 * v ** (5.0 / 3.0)) ** (-3.0 / 5.0) * 0.057 * lamda ** (6.0 / 5.0))


def rytov_variance(cn2, h, lamda=500.e-9):
    # diagnostic only in the reference (fast/fast.py:267,273)
    k = 2 * _np.pi / lamda
    return 2.25 * k ** (7.0 / 6.0) * _np.sum(cn2 * h ** (5.0 / 6.0))
