"""Host-side helpers with the names of the reference's `fast.funcs` that the drop-in path
keeps on the CPU: geometry scalars, pupil / fibre-mode construction (run once per config) and
the numpy RNG wrappers used when RNG='numpy'.  The hot functions of fast/funcs.py
(turb_powerspectrum_vonKarman, make_phase_fft, integrate_powerspectrum, pupil_filter) live in
the CUDA library instead (include/fastb.h)."""
import logging

import numpy
from scipy.integrate import simpson
from scipy.optimize import minimize_scalar

logger = logging.getLogger(__name__)

# module-global generator, replaced by Fast.set_seed (fast/funcs.py:21, fast/fast.py:768-769)
_R = numpy.random.default_rng()


def circle(radius, size):
    """Filled disc on a size x size grid whose pixel centres sit at i + 0.5 - size/2
    (the aotools.circle convention the reference relies on, fast/funcs.py:263)."""
    c = numpy.arange(size) + 0.5 - size / 2.
    return ((c[None, :] ** 2 + c[:, None] ** 2) <= radius ** 2).astype(float)


def gaussian2d(size, width):
    """Unit-amplitude Gaussian exp(-r^2 / 2 width^2) centred ON pixel (size/2, size/2)
    (aotools.gaussian2d convention: half a pixel off the disc centre)."""
    ny, nx = (size, size) if numpy.ndim(size) == 0 else size
    y = (ny / 2. - numpy.arange(ny))[:, None] / width
    x = (nx / 2. - numpy.arange(nx))[None, :] / width
    return numpy.exp(-(x ** 2 + y ** 2) / 2)


def compute_pupil(N, dx, D, obsc=0, Ny=None):
    """Annular aperture of unit power: sum(P^2) dx^2 = 1 (fast/funcs.py:261-277)."""
    ap = circle(D / dx / 2, N) - circle(obsc / dx / 2, N)
    if Ny is not None:
        assert ((Ny - N) % 2) == 0, "(Nx-Ny)/2 must be even"
        if Ny > N:
            pad = (Ny - N) // 2
            ap = numpy.pad(ap, [(0, 0), (pad, pad)])
        elif Ny < N:
            cut = (N - Ny) // 2
            ap = ap[:, cut:-cut]
    return ap / numpy.sqrt(ap.sum() * dx ** 2)


def _unit_gaussian(shape, W, dx):
    return gaussian2d(shape, W / dx / numpy.sqrt(2)) * numpy.sqrt(2. / (numpy.pi * W ** 2))


def coupling_loss(W, N, pupil, dx):
    """1 - |<g_W, P>|^2 for a unit-power Gaussian of 1/e^2 radius W (fast/funcs.py:347-350)."""
    return 1 - numpy.abs((_unit_gaussian(N, W, dx) * pupil).sum() * dx ** 2) ** 2


def optimize_fibre(pupil, dx, size_min=None, size_max=None, return_size=False):
    """Brent search for the Gaussian mode best coupled to `pupil` (fast/funcs.py:317-345)."""
    shape = pupil.shape
    size_max = max(shape) * dx if size_max is None else size_max
    size_min = dx if size_min is None else size_min

    def loss(W):
        return coupling_loss(W, shape, pupil, dx)

    opt = minimize_scalar(loss, bracket=[size_min, size_max]).x
    if abs(opt) < dx:
        # the bracket search occasionally collapses to ~0; retry once with a wider bracket
        logger.info("Gaussian mode optimisation failed, trying with different parameters")
        opt = minimize_scalar(loss, bracket=[size_min, 2 * size_max]).x
        if abs(opt) < dx:
            raise Exception("Cannot optimise gaussian mode, try changing DX?")
    g = _unit_gaussian(shape, opt, dx)
    return (g, numpy.abs(opt)) if return_size else g


def compute_gaussian_mode(pupil, dx, W0=None, D=None, obsc=None, ptype='gauss'):
    """Fibre (or launch) mode over the aperture, divided by pupil.max(); returns (mode, W0)
    (fast/funcs.py:280-305)."""
    if ptype == 'gauss':
        if isinstance(W0, str) and W0 == "opt":
            g, opt = optimize_fibre(pupil, dx, return_size=True)
            logger.debug(f"Optimised gaussian size: {opt}")
            return g / pupil.max(), opt
        return _unit_gaussian(pupil.shape, W0, dx) / pupil.max(), W0
    if ptype == 'axicon':
        if isinstance(W0, str) and W0 == "opt":
            raise TypeError("Using 'axicon' and W0='opt' not supported, please set a value for W0")
        nx, ny = pupil.shape
        yy = (numpy.arange(nx) - nx / 2)[:, None] * dx
        xx = (numpy.arange(ny) - ny / 2)[None, :] * dx
        ring = numpy.exp(-(numpy.sqrt(xx ** 2 + yy ** 2) - (obsc / 2 + (D / 2 - obsc / 2) / 2)) ** 2 / W0 ** 2)
        return ring / numpy.sqrt((ring ** 2).sum() * dx ** 2) / pupil.max(), W0
    raise Exception('ptype must be one of "gauss" or "axicon"')


def simpson_weights(f):
    """w such that scipy.integrate.simpson(y, x=f) == w @ y (end correction of the installed
    scipy included); the device evaluates fast/funcs.py:100-115 as w^T P w."""
    return simpson(numpy.eye(len(f)), x=f)


def generate_random_coefficients(shape):
    """Complex unit normals, real block drawn first (fast/funcs.py:352-356)."""
    re = _R.normal(0, 1, size=shape)
    im = _R.normal(0, 1, size=shape)
    return re + 1j * im


def generate_random_coefficients_logamp(Nscrns, powerspec, temporal=False, temporal_powerspecs=None):
    """Log-amplitude draws scaled by sqrt(variance) (fast/funcs.py:358-375).  temporal=True:
    complex noise coloured by the normalised temporal PSD, centred FFT along time."""
    if not temporal:
        shape = (Nscrns, *numpy.shape(powerspec))
        rand = _R.normal(0, 1, size=shape) + 1j * _R.normal(0, 1, size=shape)
        return rand * numpy.sqrt(powerspec)
    shape = (*numpy.shape(powerspec), Nscrns)
    spec = _R.normal(0, 1, size=shape) + 1j * _R.normal(0, 1, size=shape)
    spec = spec * numpy.sqrt(temporal_powerspecs / temporal_powerspecs.sum())
    series = numpy.fft.fftshift(numpy.fft.fft(numpy.fft.fftshift(spec, axes=-1)), axes=-1)
    return series.T * numpy.sqrt(powerspec)


def l_path(h_sat, zeta):
    """Slant range [m] to a satellite at altitude h_sat seen at zenith angle zeta [deg]
    (law of cosines on the Earth-centre triangle; fast/funcs.py:388-399)."""
    r_earth = 6.371009e6
    z = numpy.radians(zeta)
    b = -2 * r_earth * numpy.cos(numpy.pi - z)
    c = r_earth ** 2 - (r_earth + h_sat) ** 2
    root = numpy.sqrt(b ** 2 - 4 * c)
    first = (-b + root) / 2
    return first if first >= 0 else (-b - root) / 2


def calculate_wind_correction(h, theta_loop, Tloop):
    """Pseudo-wind [m/s] per layer from the apparent slew of a LEO satellite during one loop
    period (fast/funcs.py:403-406)."""
    sx = numpy.sin(numpy.radians(theta_loop[0] / 3600))
    sy = numpy.sin(numpy.radians(theta_loop[1] / 3600))
    return -numpy.array([sx * h / Tloop, sy * h / Tloop]).T
