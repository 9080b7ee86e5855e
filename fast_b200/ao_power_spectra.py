"""Host-side counterparts of the reference's `fast.ao_power_spectra` mask helpers
(fast/ao_power_spectra.py:10-141), kept for API compatibility (`mask_lf`, `zernike_ft`,
`zernike_squared_filter` are public names of the reference) and for the TEMPORAL frequency
grids.  `Fast` itself builds the modal / tip-tilt / LGS masks on the device
(`fastb_zernike_filter`); the per-pixel PSD terms (G_AO_PAOLA, Jol_alias_openloop,
Jol_noise_openloop, logamp_powerspec) are computed by the CUDA library."""
import warnings

import numpy
from scipy.special import jv


def zernIndex(j):
    """Noll index -> [n, m] (aotools.functions.zernike.zernIndex semantics)."""
    n = int((-1. + numpy.sqrt(8 * (j - 1) + 1)) / 2.)
    p = j - (n * (n + 1)) / 2.
    k = n % 2
    m = int((p + k) / 2.) * 2 - k
    if m != 0:
        m *= 1 if j % 2 == 0 else -1
    return [n, m]


def zernike_ft(fabs, phi, D, n_noll):
    """Fourier transform of Zernike polynomial n_noll over a disc of diameter D (Noll 1976)."""
    n, m = zernIndex(n_noll)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        x = fabs * D / 2
        radial = 2 * jv(n + 1, x) / x
        if m == 0:
            return numpy.sqrt(n + 1) * (-1) ** (n / 2.) * radial
        azim = numpy.cos(m * phi) if n_noll % 2 == 0 else numpy.sin(m * phi)
        return numpy.sqrt(2 * (n + 1)) * (-1) ** ((n - m) / 2.) * (1j) ** m * radial * azim


def zernike_squared_filter(fabs, fx, fy, D, n_noll, n_noll_start=1):
    """sum_{j=start}^{n_noll} |Z_j(f)|^2 with the DC pixel set to 1 (start == 1) or 0."""
    phi = numpy.arctan2(fy, fx)
    out = numpy.zeros(fabs.shape, dtype=complex)
    for j in range(n_noll_start, n_noll + 1):
        out += numpy.abs(zernike_ft(fabs, phi, D, j)) ** 2
    out[..., int(fabs.shape[-2] / 2), int(fabs.shape[-1] / 2)] = 1 if n_noll_start == 1 else 0
    return out


def mask_lf(freq, d_WFS, modal=False, modal_mult=1, Zmax=None, D=None, Gtilt=False):
    """AO-corrected spatial-frequency region (fast/ao_power_spectra.py:119-141): the WFS box
    |fx|,|fy| <= pi/d, times a DM term (box, disc, or <=1-clipped Zernike filter)."""
    fx, fy = freq.fx, freq.fy
    fmax = numpy.pi / d_WFS
    box = numpy.logical_and(abs(fx) <= fmax, abs(fy) <= fmax)
    if not modal:
        dm = box
    else:
        fabs = numpy.sqrt(fx ** 2 + fy ** 2)
        if Zmax is None:
            dm = fabs <= fmax * modal_mult
        elif Gtilt:
            gt = (zernike_squared_filter(fabs, fx, fy, D, 1) + jv(1, fabs * D / 2.) ** 2).real
            dm = numpy.minimum(gt, 1.)
        else:
            dm = zernike_squared_filter(fabs, fx, fy, D, Zmax).real
    dm = numpy.where(dm < 1, dm, 1)
    return box * dm
