"""Host-side bookkeeping of the TEMPORAL (frozen-flow) mode, mirroring what the reference does
on the CPU around its hot loop (fast/fast.py:394-405, 538-587, 617-635, 846-875): per-layer
temporal frequency grids, the elongated pupil-filter spline, the temporal log-amplitude PSD and
the per-step sample coordinates.  These are small (L x J x Npup) or once-per-config arrays; the
screens, the bilinear gather and the detector run on the device (fastb_layer_screens,
fastb_temporal_detect)."""
import numpy
from scipy.interpolate import RectBivariateSpline

from . import funcs


def temporal_axes(nlayer, Ny, Nx, wind_speed, dt, dfy):
    """Per-layer axes of SpatialFrequencies.make_temporal_freqs (fast/fast.py:846-864): the x
    axis is the LINEAR frequency 1/(Nx v dt) conjugate to time (the reference's own NOTE), the
    y axis repeats the main grid."""
    steps = numpy.arange(-Nx / 2, Nx / 2)
    rows = numpy.arange(-Ny / 2, Ny / 2)
    fx_axes = numpy.array([steps * (1 / (Nx * (wind_speed[i] * dt))) for i in range(nlayer)])
    fy_axes = numpy.array([rows * dfy for _ in range(nlayer)])
    return fx_axes, fy_axes


def elongated_pupil_filter(sim, fx_axes, fy_axes):
    """Bilinear spline of |FT(P M)|^2 on a grid fine enough for the temporal axes
    (fast/fast.py:394-405; funcs.pupil_filter(spline=True), fast/funcs.py:308-313)."""
    f_max = max(fx_axes.max(), fy_axes.max())
    dx_req = numpy.pi / f_max
    N_req = int(2 * numpy.ceil(2 * numpy.pi / (sim.freq.main.df * dx_req) / 2))
    Ny = 2 * sim.Npxls_pup
    pupil = funcs.compute_pupil(N_req, dx_req, sim.D_ground, sim.obsc_ground, Ny=Ny)
    mode, _ = funcs.compute_gaussian_mode(pupil, dx_req, W0=sim.W0, ptype="gauss")
    pm = pupil * mode
    ax = (-1, -2)
    spec = numpy.fft.fftshift(numpy.fft.fft2(numpy.fft.fftshift(pm, axes=ax)), axes=ax)
    P = numpy.abs(spec) ** 2 / pm.sum() ** 2
    fx = numpy.arange(-N_req / 2., N_req / 2.) * (2 * numpy.pi / (N_req * dx_req))
    fy = numpy.arange(-Ny / 2., Ny / 2.) * (2 * numpy.pi / (Ny * sim.dx))
    return RectBivariateSpline(fx, fy, P, kx=1, ky=1, s=0)


def temporal_logamp_powerspec(sim, fx_axes, fy_axes, fabs, spline):
    """Log-amplitude PSD on the per-layer temporal grids, integrated over the axis orthogonal
    to the wind (fast/fast.py:581-587 with ao_power_spectra.logamp_powerspec :272-301).  The
    spline is sampled at (fy_axis, fx_axis), un-rotated, as in the reference."""
    km2 = (5.92 / sim.l0) ** 2
    k02 = (2 * numpy.pi / sim.L0) ** 2
    total = numpy.zeros(fabs.shape[1:])
    with numpy.errstate(all='ignore'):
        for i in range(fabs.shape[0]):
            f2 = fabs[i] ** 2
            vk = 0.033 * numpy.exp(-f2 / km2) / (f2 + k02) ** (11 / 6.) * sim.cn2[i]
            vk[numpy.isinf(vk)] = 0.
            ps = vk * 2 * numpy.pi * sim.k ** 2 * numpy.sin(sim.wvl * sim.h[i] * f2 / (4 * numpy.pi)) ** 2
            total += ps * spline(fy_axes[i], fx_axes[i])
    return total.sum(-2) * sim.freq.main.dfy


def sample_coordinates(interp_coords, N):
    """(..., L, 2, J, Npup) wind-shifted pupil coordinates -> integer/fraction sample positions
    per output pixel, reproducing fast/fast.py:621-633 literally: wrap mod N, sort, roll back
    by the argmax of the gaps (0 without wrap; one short of the true inverse with a wrap, which
    is what the reference computes), then FITPACK's clamp of arguments beyond the last knot.
    Same values as the plain numpy statement (tests/test_host_mirror.py keeps it), with the two
    slow steps -- the float modulo and the per-row roll -- applied only where they do something."""
    x = numpy.asarray(interp_coords, dtype=float)
    outside = (x < 0) | (x >= N)
    wrapped = x.copy()
    if outside.any():
        wrapped[outside] = x[outside] % N          # x % N == x for 0 <= x < N
    coord = numpy.sort(wrapped, axis=-1)
    gaps = numpy.abs(numpy.diff(coord, axis=-1))
    roll = gaps.argmax(-1)
    roll[(numpy.abs(gaps - 1) <= 1e-8 + 1e-5).all(-1)] = 0        # numpy.isclose(gaps, 1), spelled out
    npup = coord.shape[-1]
    at = coord
    rows = numpy.nonzero(roll)
    if rows[0].size:
        at = coord.copy()
        sel = coord[rows]                                          # (n_rolled, Npup)
        at[rows] = numpy.take_along_axis(sel, (numpy.arange(npup) + roll[rows][:, None]) % npup, axis=-1)
    at = numpy.minimum(at, N - 1.0)
    i0 = numpy.minimum(numpy.floor(at).astype(numpy.int32), N - 2)
    frac = (at - i0).astype(numpy.float32)
    # -> xi, xf, yi, yf each (L, J, Npup); leading axes (e.g. one per chunk) are kept
    return (numpy.ascontiguousarray(i0[..., 0, :, :]), numpy.ascontiguousarray(frac[..., 0, :, :]),
            numpy.ascontiguousarray(i0[..., 1, :, :]), numpy.ascontiguousarray(frac[..., 1, :, :]))
