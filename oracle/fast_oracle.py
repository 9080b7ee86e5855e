"""CPU ORACLE (TEST INFRASTRUCTURE -- NOT PRODUCT CODE).

A float64 numpy restatement of the Monte-Carlo hot path of ojdf/fast (FAST): residual phase
PSD build -> coloured-noise phase screens -> pupil crop -> fibre-overlap detector.  It is
written independently of the reference's classes (functions over plain arrays) so that a
disagreement between the CUDA path and the reference can be localised term by term.

Who may import this: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  The product package `fast_b200` never imports it; the product path has no
CPU fallback.

Parity status: PINNED.  The reference's own tests hold no golden vectors (test/tests_pytest.py
asserts finiteness only), so the pin is "outputs of the reference itself run here":
oracle/make_golden.py imports the unmodified /root/reference through oracle/shim and commits
its outputs under tests/golden/; tests/test_oracle_vs_golden.py checks every function below
against them (<=1e-12 relative on PSD terms, <=1e-10 on per-realisation results).

Every function cites the reference lines it restates (paths relative to /root/reference).
The third-party arithmetic not vendored in the reference (aotools 1.0.x circle / gaussian2d /
ft2 / ift2 / cn2_to_r0 ...) is restated from its published behaviour; see SURVEY.md App. A.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.integrate import simpson as _simpson
from scipy.optimize import minimize_scalar as _minimize_scalar

ARCSEC = 206265.0
R_EARTH = 6.371009e6


# --------------------------------------------------------------------------------------
# config defaults (fast/conf.py:67-115) -- key names are the API contract
# --------------------------------------------------------------------------------------
def default_params():
    return {
        'NPXLS': 'auto', 'DX': 'auto', 'NITER': 1000, 'SUBHARM': False, 'FFTW': False,
        'FFTW_THREADS': 1, 'NCHUNKS': 10, 'TEMPORAL': False, 'DT': 0.001, 'LOGFILE': None,
        'LOGLEVEL': 'INFO', 'SEED': None,
        'W0': 'opt', 'D_GROUND': 1.0, 'OBSC_GROUND': 0, 'D_SAT': 0.1, 'OBSC_SAT': 0,
        'WVL': 1550e-9, 'AXICON': False, 'POWER': 1, 'SMF': True,
        'H_SAT': 36e6, 'L_SAT': None, 'H_TURB': np.array([0, 10e3]),
        'CN2_TURB': np.array([100e-15, 100e-15]), 'WIND_SPD': np.array([10, 10]),
        'WIND_DIR': np.array([90., 0.]), 'L0': np.inf, 'l0': 1e-6, 'ZENITH_ANGLE': 0,
        'PROP_DIR': 'up', 'DTHETA': [4, 0], 'TRANSMISSION': 1,
        'AO_MODE': 'AO', 'DSUBAP': 0.02, 'TLOOP': 0.001, 'TEXP': 0.001, 'ALIAS': True,
        'NOISE': 0.0, 'MODAL': False, 'MODAL_MULT': 1, 'ZMAX': None,
        'COHERENT': False, 'MODULATION': None, 'EsN0': None,
    }


def with_defaults(params):
    out = default_params()
    out.update(params)
    return out


# --------------------------------------------------------------------------------------
# turbulence / wind profile generators (fast/turbulence_models.py:4-105) -- config inputs
# --------------------------------------------------------------------------------------
def hv57(h, w=21.0, A=1.7e-14):
    """Hufnagel-Valley 5/7 Cn2(h) (fast/turbulence_models.py:4-19)."""
    h = np.asarray(h, dtype=float)
    return (0.00594 * (w / 27) ** 2 * (1e-5 * h) ** 10 * np.exp(-h / 1000)
            + 2.7e-16 * np.exp(-h / 1500) + A * np.exp(-h / 100.))


def bufton_wind(h, vg=8.0, vt=30.0, ht=9400.0, Lt=4800.0):
    """Bufton wind-speed profile (fast/turbulence_models.py:22-38)."""
    h = np.asarray(h, dtype=float)
    return vg + vt * np.exp(-((h - ht) / Lt) ** 2)


def compress_layers(h, p, nlayers, w=None):
    """Equivalent-layers compression (fast/turbulence_models.py:65-105): slab sums of Cn2,
    5/3-moment effective heights (and wind speeds)."""
    step = (h.max() - h.min()) / nlayers
    slab = np.digitize(h, np.arange(h.min(), h.max(), step))
    h_out, c_out, w_out = np.zeros(nlayers), np.zeros(nlayers), np.zeros(nlayers)
    for i in range(nlayers):
        sel = slab == i + 1
        tot = p[sel].sum()
        c_out[i] = tot
        h_out[i] = ((p[sel] * h[sel] ** (5 / 3)).sum() / tot) ** (3 / 5)
        if w is not None:
            w_out[i] = ((p[sel] * w[sel] ** (5 / 3)).sum() / tot) ** (3 / 5)
    return (h_out, c_out, w_out) if w is not None else (h_out, c_out)


def hv57_bufton_profile(nlayers, **kw):
    """fast/turbulence_models.py:41-62: 1 m bins to 30 km, then compress."""
    h0 = np.arange(0, 30000)
    keys_c = {k: kw[k] for k in ('w', 'A') if k in kw}
    keys_w = {k: kw[k] for k in ('vg', 'vt', 'ht', 'Lt') if k in kw}
    return compress_layers(h0, hv57(h0, **keys_c), nlayers, w=bufton_wind(h0, **keys_w))


# --------------------------------------------------------------------------------------
# geometry scalars (fast/fast.py:229-276, 286; fast/funcs.py:388-406)
# --------------------------------------------------------------------------------------
def path_length(h_sat, zenith_deg):
    """Slant range to a satellite at altitude h_sat (fast/funcs.py:388-399)."""
    z = np.radians(zenith_deg)
    b = -2 * R_EARTH * np.cos(np.pi - z)
    c = R_EARTH ** 2 - (R_EARTH + h_sat) ** 2
    disc = np.sqrt(b ** 2 - 4 * c)
    r1 = (-b + disc) / 2
    return r1 if r1 >= 0 else (-b - disc) / 2


def atmosphere(p):
    """Per-layer geometry feeding the PSD filters (fast/fast.py:229-276)."""
    gamma = 1 / np.cos(np.radians(p['ZENITH_ANGLE']))
    h = np.asarray(p['H_TURB'], dtype=float) * gamma
    cn2 = np.asarray(p['CN2_TURB'], dtype=float) * gamma
    L = p['L_SAT'] if p['L_SAT'] is not None else path_length(p['H_SAT'], p['ZENITH_ANGLE'])
    dtheta = p['DTHETA']
    paa = np.sqrt(dtheta[0] ** 2 + dtheta[1] ** 2)
    wdir = p['WIND_DIR']
    if 'AZIMUT_SAT' in p:
        wdir = [(x - p['AZIMUT_SAT']) % 380 for x in wdir]           # sic: % 380 (fast/fast.py:250)
    wrad = np.radians(wdir)
    wind = (np.asarray(p['WIND_SPD']) * np.array([np.cos(wrad), np.sin(wrad) / gamma])).T
    if 'ANISO_DL' in p:                                              # fast/funcs.py:403-406
        a = p['ANISO_DL']
        wind = wind + (-np.array([np.sin(np.radians(a[0] / 3600)) * h / p['TLOOP'],
                                  np.sin(np.radians(a[1] / 3600)) * h / p['TLOOP']]).T)
    speed = np.sqrt(wind[:, 0] ** 2 + wind[:, 1] ** 2)
    k500 = 2 * np.pi / 500e-9
    kw = 2 * np.pi / p['WVL']
    cn2_0 = np.asarray(p['CN2_TURB'], dtype=float)
    h_0 = np.asarray(p['H_TURB'], dtype=float)
    w_0 = np.asarray(p['WIND_SPD'], dtype=float)
    out = dict(gamma=gamma, h=h, cn2=cn2, L=L, dtheta=dtheta, paa=paa, wind_dir=wdir,
               wind_vector=wind, wind_speed=speed,
               r0=(0.423 * k500 ** 2 * cn2_0.sum()) ** (-3 / 5),
               theta0=0.057 * 500e-9 ** (6 / 5) * np.sum(cn2_0 * h_0 ** (5 / 3)) ** (-3 / 5) * 180.0 * 3600.0 / np.pi,
               tau0=float(np.sum(cn2_0 * w_0 ** (5 / 3)) ** (-3 / 5) * 0.057 * 500e-9 ** (6 / 5)),
               r0_los=(0.423 * kw ** 2 * cn2.sum()) ** (-3 / 5),
               theta0_los=0.057 * p['WVL'] ** (6 / 5) * np.sum(cn2 * h ** (5 / 3)) ** (-3 / 5) * 180.0 * 3600.0 / np.pi,
               tau0_los=float(np.sum(cn2 * speed ** (5 / 3)) ** (-3 / 5) * 0.057 * p['WVL'] ** (6 / 5)))
    return out


def grid_size(p, atm):
    """DX / NPXLS 'auto' rules and the pupil crop window (fast/fast.py:151-213,390)."""
    D = p['D_GROUND']
    if p['DX'] == 'auto':
        dx = np.min([p['DSUBAP'] / 2, atm['r0_los'] / 2, D / 10])
        if p['AO_MODE'] == 'NOAO':
            dx = atm['r0_los'] / 2
    else:
        dx = p['DX']
    if p['NPXLS'] == 'auto':
        nyq = np.min([np.pi / (atm['h'][-1] * atm['paa'] / ARCSEC),
                      np.pi / (max(atm['wind_speed']) * p['TLOOP']),
                      np.pi / p['DSUBAP'] / 5])
        n_nyq = int(2 * np.ceil(2 * np.pi / (nyq * dx) / 2))
        n_ap = int(2 * np.ceil(D / dx / 2)) + 2
        n_t = 0
        if p['TEMPORAL']:
            n_t = int(np.asarray(p['WIND_SPD']).max() * p['DT'] * p['NITER'] / p['DX'] / 2)
        N = int(np.max([n_nyq, n_ap, n_t]))
    else:
        N = int(p['NPXLS'])
    npup = int(np.ceil(D / dx)) + 2
    lo, hi = (N - npup) // 2, (N + npup) // 2
    return dict(dx=dx, N=N, Npup=npup, lo=lo, hi=hi, df=2 * np.pi / (N * dx))


def freq_axis(N, dx):
    """Centred angular-frequency axis (fast/fast.py:830-833): (i - N/2) * 2pi/(N dx)."""
    return np.arange(-N / 2., N / 2.) * (2 * np.pi / (N * dx))


# --------------------------------------------------------------------------------------
# PSD terms.  Grids: fx[r, c] = ax[c], fy[r, c] = ay[r]  (numpy.meshgrid, fast/fast.py:911)
# --------------------------------------------------------------------------------------
def von_karman_base(fabs, L0, l0):
    """0.033 exp(-f^2/km^2) / (f^2 + k0^2)^(11/6), +-inf -> 0 (fast/funcs.py:151-170);
    per unit Cn2.  km = 5.92/l0, k0 = 2 pi/L0."""
    km = 5.92 / l0
    k0 = 2 * np.pi / L0
    with np.errstate(all='ignore'):
        base = 0.033 * np.exp(-fabs ** 2 / km ** 2) / (fabs ** 2 + k0 ** 2) ** (11 / 6.)
    base = np.array(base, dtype=float)
    base[np.isinf(base)] = 0.
    return base


def von_karman(fabs, cn2, L0, l0):
    """(L, ...) stack: base * cn2_l (fast/funcs.py:162-164)."""
    base = von_karman_base(fabs, L0, l0)
    return np.asarray(cn2, dtype=float).reshape((-1,) + (1,) * base.ndim) * base[None]


def zonal_mask(fx, fy, dsubap):
    """|fx| <= pi/d and |fy| <= pi/d (fast/ao_power_spectra.py:123-124,135-140)."""
    fmax = np.pi / dsubap
    return np.logical_and(np.abs(fx) <= fmax, np.abs(fy) <= fmax)


def _zernike_ft(fabs, phi, D, j):
    """Fourier transform of Noll Zernike j on a disc of diameter D
    (fast/ao_power_spectra.py:10-21)."""
    from scipy.special import jv
    n = int((-1. + np.sqrt(8 * (j - 1) + 1)) / 2.)
    pp = j - (n * (n + 1)) / 2.
    kk = n % 2
    m = int((pp + kk) / 2.) * 2 - kk
    if m != 0:
        m *= 1 if j % 2 == 0 else -1
    with np.errstate(all='ignore'):
        rad = 2 * jv(n + 1, fabs * D / 2) / (fabs * D / 2)
        if m == 0:
            return np.sqrt(n + 1) * (-1) ** (n / 2.) * rad
        ang = np.cos(m * phi) if j % 2 == 0 else np.sin(m * phi)
        return np.sqrt(2 * (n + 1)) * (-1) ** ((n - m) / 2.) * (1j) ** m * rad * ang


def zernike_sq_filter(fabs, fx, fy, D, jmax, jstart=1):
    """sum_j |Z_j(f)|^2 with DC fixed to 1 (jstart==1) or 0 (fast/ao_power_spectra.py:54-76)."""
    phi = np.arctan2(fy, fx)
    out = np.zeros(fabs.shape, dtype=complex)
    for j in range(jstart, jmax + 1):
        out += np.abs(_zernike_ft(fabs, phi, D, j)) ** 2
    out[..., int(fabs.shape[-2] / 2), int(fabs.shape[-1] / 2)] = 1 if jstart == 1 else 0
    return out


def lf_mask(fx, fy, dsubap, modal=False, modal_mult=1, Zmax=None, D=None):
    """AO-corrected region mask (fast/ao_power_spectra.py:119-141).  Zonal: boolean box;
    modal: disc, or the (<=1-clipped) Zernike squared filter when Zmax is given."""
    box = zonal_mask(fx, fy, dsubap)
    if not modal:
        dm = box
    else:
        fabs = np.sqrt(fx ** 2 + fy ** 2)
        if Zmax is None:
            dm = fabs <= (np.pi / dsubap) * modal_mult
        else:
            dm = zernike_sq_filter(fabs, fx, fy, D, Zmax).real
    dm = np.where(dm < 1, dm, 1)
    return box * dm


def _np_sinc(x):
    return np.sinc(x)


def g_ao(fx, fy, mask, mode, h, wind, dtheta, tloop, texp, D=None):
    """Aniso-servo transfer function per layer (fast/ao_power_spectra.py:225-270):
    G = M (1 - 2 cos(dr.k - tl v.k) s + s^2) + (1 - M),  s = sinc(texp v.k / 2pi)."""
    if mode not in ('NOAO', 'AO', 'TT', 'LGSAO'):
        raise Exception('Mode not recognised')
    if mode == 'NOAO':
        return 1
    h = np.asarray(h, dtype=float)
    drx = (dtheta[0] / ARCSEC * h)[:, None, None]
    dry = (dtheta[1] / ARCSEC * h)[:, None, None]
    a = fx[None] * drx + fy[None] * dry
    b = fx[None] * wind[:, 0][:, None, None] + fy[None] * wind[:, 1][:, None, None]
    s = _np_sinc(texp * b / (2 * np.pi))
    aniso = 1 - 2 * np.cos(a - tloop * b) * s + s ** 2
    if mode in ('AO', 'TT'):
        return aniso * mask + (1 - mask)
    # LGSAO: low orders (Noll 1..4) from the NGS direction, the rest servo-lag only (:262-267)
    aniso_lgs = 1 - 2 * np.cos(-tloop * b) * s + s ** 2
    Z = zernike_sq_filter(np.sqrt(fx ** 2 + fy ** 2), fx, fy, D, 4).real
    return mask * (Z * aniso + (1 - Z) * aniso_lgs) + (1 - mask)


def alias_psd(ax, ay, dsubap, cn2, mask, wind, texp, L0, l0, lmax=5, kmax=5):
    """Open-loop WFS aliasing PSD per layer (fast/ao_power_spectra.py:163-223).  Sum over the
    (2lmax+1)(2kmax+1)-1 replicas of the von Karman spectrum shifted by 2pi(k, l)/d, with the
    reference's row/column/DC overrides applied in its order, then x sinc^2 x mask, NaN -> 0.
    Summation order (l outer, k inner) follows the reference."""
    fx, fy = np.meshgrid(ax, ay)
    cn2 = np.asarray(cn2, dtype=float)
    nl = len(cn2)
    N_r, N_c = fx.shape
    mr, mc = int(N_r / 2.), int(N_c / 2.)
    acc = np.zeros((nl, N_r, N_c))
    b = fx[None] * wind[:, 0][:, None, None] + fy[None] * wind[:, 1][:, None, None]
    with np.errstate(all='ignore'):
        sinc2 = _np_sinc(texp * b / (2 * np.pi)) ** 2
        fabs = np.sqrt(fx ** 2 + fy ** 2)
        t0 = fx ** 2 * fy ** 2 / fabs ** 4
        for l in range(-lmax, lmax + 1):
            ys = ay - 2 * np.pi * l / dsubap
            for k in range(-kmax, kmax + 1):
                if l == 0 and k == 0:
                    continue
                xs = ax - 2 * np.pi * k / dsubap
                sx, sy = np.meshgrid(xs, ys)
                t1 = (fx / sy + fy / sx) ** 2
                t2 = von_karman(np.sqrt(sx ** 2 + sy ** 2), cn2, L0, l0)
                m = t1 * t2 * t0
                m[:, mr, mc] = 0.
                if l == 0:
                    m[:, mr, :] = t2[:, mr, :]
                if k == 0:
                    m[:, :, mc] = t2[:, :, mc]
                    m[:, mr, mc] = t2[:, mr, mc]
                acc += m
        acc *= sinc2 * mask
    acc[np.isnan(acc)] = 0.
    return acc


def noise_psd(fx, fy, dsubap, noise_var, mask):
    """Open-loop WFS noise PSD (fast/ao_power_spectra.py:148-161); DC forced to 0."""
    with np.errstate(all='ignore'):
        fabs = np.sqrt(fx ** 2 + fy ** 2)
        ps = noise_var / (fabs ** 2 * _np_sinc(dsubap * fx / (2 * np.pi)) ** 2
                          * _np_sinc(dsubap * fy / (2 * np.pi)) ** 2)
    ps[..., int(ps.shape[-2] / 2.), int(ps.shape[-1] / 2.)] = 0.
    return mask * ps


def simpson2d(P, f):
    """simpson(simpson(P, x=f), x=f) over the last two axes (fast/funcs.py:100-115)."""
    return _simpson(_simpson(P, x=f), x=f)


def simpson_weights(f):
    """The weight vector w with simpson(y, x=f) == w @ y for this scipy (even-N end correction
    included).  2-D integral = w^T P w.  """
    return _simpson(np.eye(len(f)), x=f)


# --------------------------------------------------------------------------------------
# pupil, fibre mode, pupil filter (fast/funcs.py:261-350; fast/fast.py:375-392)
# --------------------------------------------------------------------------------------
def disc(radius, n):
    """aotools.circle: 1 where (i+0.5-n/2)^2 + (j+0.5-n/2)^2 <= radius^2."""
    c = np.arange(n) + 0.5 - n / 2.
    xx, yy = np.meshgrid(c, c)
    return (xx * xx + yy * yy <= radius * radius).astype(float)


def gauss2d(shape, width):
    """aotools.gaussian2d, unit amplitude, centred on pixel (n/2, n/2)."""
    ny, nx = shape
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny))
    return np.exp(-(((nx / 2. - X) / width) ** 2 + ((ny / 2. - Y) / width) ** 2) / 2)


def pupil_aperture(N, dx, D, obsc=0):
    """Annular aperture normalised to unit power (fast/funcs.py:261-277)."""
    ap = disc(D / dx / 2, N) - disc(obsc / dx / 2, N)
    return ap / np.sqrt(ap.sum() * dx ** 2)


def fibre_mode(pupil, dx, W0, D=None, obsc=None, ptype='gauss'):
    """Backpropagated fibre mode over the aperture, divided by pupil.max()
    (fast/funcs.py:280-350).  W0 == 'opt' -> Brent minimisation of 1 - |sum g P dx^2|^2."""
    shape = pupil.shape

    def unit_gauss(W):
        return gauss2d(shape, W / dx / np.sqrt(2)) * np.sqrt(2. / (np.pi * W ** 2))

    if ptype == 'gauss':
        if isinstance(W0, str) and W0 == 'opt':
            def loss(W):
                return 1 - np.abs((unit_gauss(W) * pupil).sum() * dx ** 2) ** 2
            smax = max(shape) * dx
            opt = _minimize_scalar(loss, bracket=[dx, smax]).x
            if abs(opt) < dx:
                opt = _minimize_scalar(loss, bracket=[dx, 2 * smax]).x
                if abs(opt) < dx:
                    raise Exception('Cannot optimise gaussian mode, try changing DX?')
            return unit_gauss(opt) / pupil.max(), np.abs(opt)
        return unit_gauss(W0) / pupil.max(), W0
    if ptype == 'axicon':
        if isinstance(W0, str) and W0 == 'opt':
            raise TypeError("Using 'axicon' and W0='opt' not supported, please set a value for W0")
        nx, ny = shape
        xx, yy = np.meshgrid(np.arange(-ny / 2, ny / 2, 1) * dx, np.arange(-nx / 2, nx / 2, 1) * dx)
        r = np.sqrt(xx ** 2 + yy ** 2)
        mid = obsc / 2 + (D / 2 - obsc / 2) / 2
        ring = np.exp(-(r - mid) ** 2 / W0 ** 2)
        return ring / np.sqrt((ring ** 2).sum() * dx ** 2) / pupil.max(), W0
    raise Exception('ptype must be one of "gauss" or "axicon"')


def pupil_filter(pm):
    """|FT2(P M)|^2 / (sum P M)^2 on the full grid, centred transform, unit spacing
    (fast/funcs.py:308-311 with aotools.ft2)."""
    ax = (-1, -2)
    F = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(pm, axes=ax)), axes=ax)
    return np.abs(F) ** 2 / pm.sum() ** 2


def logamp_psd(fabs, h, cn2, wvl, pf, L0, l0):
    """Aperture-filtered Rytov log-amplitude PSD, summed over layers
    (fast/ao_power_spectra.py:272-301)."""
    h = np.asarray(h, dtype=float)
    ps = von_karman(fabs, cn2, L0, l0) * 2 * np.pi * (2 * np.pi / wvl) ** 2
    ps = ps * np.sin((wvl * h)[:, None, None] * (fabs ** 2)[None] / (4 * np.pi)) ** 2
    if pf is not None:
        ps = ps * pf
    return ps.sum(0)


# --------------------------------------------------------------------------------------
# link budget (fast/fast.py:670-734) -- analytic scalars; diffraction_limit scales power
# --------------------------------------------------------------------------------------
def link_budget(p, L, W0, W0_sat, pupil, mode, pupil_sat, mode_sat, dx, dx_sat):
    up = p['PROP_DIR'] == 'up'
    D_t, D_r = (p['D_GROUND'], p['D_SAT']) if up else (p['D_SAT'], p['D_GROUND'])
    ob_t, ob_r = (p['OBSC_GROUND'], p['OBSC_SAT']) if up else (p['OBSC_SAT'], p['OBSC_GROUND'])
    m, dx_r, pup_r, w0 = (mode_sat, dx_sat, pupil_sat, W0) if up else (mode, dx, pupil, W0_sat)
    wvl = p['WVL']
    lb = {}
    lb['power'] = 10 * np.log10(p['POWER'] / 1e-3)
    lb['free_space'] = 10 * np.log10((wvl / (4 * np.pi * L)) ** 2)
    alpha = D_t / (2 * w0)
    gam = ob_t / D_t
    g_t = 2 / alpha ** 2 * (np.exp(-alpha ** 2) - np.exp(-gam ** 2 * alpha ** 2)) ** 2
    lb['transmitter_gain'] = 10 * np.log10((np.pi * D_t ** 2) * 4 * np.pi / wvl ** 2 * g_t)
    A = np.pi * ((D_r / 2) ** 2 - (ob_r / 2) ** 2)
    lb['receiver_gain'] = 10 * np.log10(4 * np.pi * A / wvl ** 2)
    lb['transmission_loss'] = 10 * np.log10(p['TRANSMISSION'])
    lb['smf_coupling'] = 10 * np.log10(((pup_r * m).sum() * dx_r) ** 2 / (m ** 2).sum())
    return lb, 10 ** (sum(lb.values()) / 10) / 1e3


# --------------------------------------------------------------------------------------
# whole init (fast/fast.py:71-113, 445-492)
# --------------------------------------------------------------------------------------
def build(params, alias_terms=5):
    """Everything Fast.__init__ computes for the non-temporal Monte-Carlo path, as a dict."""
    p = with_defaults(params)
    atm = atmosphere(p)
    g = grid_size(p, atm)
    N, dx, npup, lo, hi = g['N'], g['dx'], g['Npup'], g['lo'], g['hi']
    ax = freq_axis(N, dx)
    fx, fy = np.meshgrid(ax, ax)
    fabs = np.sqrt(fx ** 2 + fy ** 2)
    df = ax[1] - ax[0]
    mode_name = p['AO_MODE']
    zmax, modal, mmult = p['ZMAX'], p['MODAL'], p['MODAL_MULT']
    if mode_name == 'TT':
        zmax, modal, mmult = 3, True, 1
    D = p['D_GROUND']
    mask = lf_mask(fx, fy, p['DSUBAP'], modal=modal, modal_mult=mmult, Zmax=zmax, D=D)

    # pupil / mode on the full grid, pupil filter, then crop (fast/fast.py:375-392)
    pupil_full = pupil_aperture(N, dx, D, p['OBSC_GROUND'])
    dx_sat = p['D_SAT'] / 32
    pupil_sat = pupil_aperture(32, dx_sat, p['D_SAT'], p['OBSC_SAT'])
    ptype = 'axicon' if p['AXICON'] else 'gauss'
    mode_full, W0 = fibre_mode(pupil_full, dx, p['W0'], D=D, obsc=p['OBSC_GROUND'], ptype=ptype)
    mode_sat, W0_sat = fibre_mode(pupil_sat, dx_sat, 'opt')
    pf = pupil_filter(pupil_full * mode_full)
    pupil = pupil_full[lo:hi, lo:hi]
    mode = mode_full[lo:hi, lo:hi]

    lb, difflim = link_budget(p, atm['L'], W0, W0_sat, pupil, mode, pupil_sat, mode_sat, dx, dx_sat)

    k = 2 * np.pi / p['WVL']
    h, cn2, wind = atm['h'], atm['cn2'], atm['wind_vector']
    L0, l0 = p['L0'], p['l0']
    turb = von_karman(fabs, cn2, L0, l0)
    G = g_ao(fx, fy, mask, mode_name, h, wind, atm['dtheta'], p['TLOOP'], p['TEXP'], D=D)
    aniso_servo_error = simpson2d((G * turb).sum(0) * mask * 2 * np.pi * k ** 2, ax)
    if p['ALIAS'] and mode_name != 'NOAO':
        alias = alias_psd(ax, ax, p['DSUBAP'], cn2, mask, wind, p['TEXP'], L0, l0,
                          lmax=alias_terms, kmax=alias_terms)
        alias_error = simpson2d((alias * 2 * np.pi * k ** 2).sum(0), ax)
    else:
        alias, alias_error = 0., 0.
    if p['NOISE'] > 0 and mode_name != 'NOAO':
        noise = noise_psd(fx, fy, p['DSUBAP'], p['NOISE'], mask)
        noise_error = simpson2d(noise, ax)
    else:
        noise, noise_error = 0., 0.
    per_layer = 2 * np.pi * k ** 2 * (turb * G + alias) + noise / len(h)
    W = per_layer.sum(0)
    fitting_error = simpson2d(W * (1 - mask), ax)
    phs_var = simpson2d(W, ax)
    phs_var_weights = simpson2d(per_layer, ax) / phs_var
    Wchi = logamp_psd(fabs, h, cn2, p['WVL'], pf, L0, l0)
    logamp_var = simpson2d(Wchi, ax)
    out = dict(params=p, atm=atm, N=N, dx=dx, Npup=npup, lo=lo, hi=hi, df=df, f=ax, k=k,
               lf_mask=mask, pupil=pupil, pupil_mode=mode, W0=W0, W0_sat=W0_sat,
               pupil_filter=pf, link_budget=lb, diffraction_limit=difflim,
               turb_powerspec=turb, G_ao=G, alias_powerspec=alias, noise_powerspec=noise,
               powerspec_per_layer=per_layer, powerspec=W, logamp_powerspec=Wchi,
               aniso_servo_error=aniso_servo_error, alias_error=alias_error,
               noise_error=noise_error, fitting_error=fitting_error, phs_var=phs_var,
               phs_var_weights=phs_var_weights, logamp_var=logamp_var)
    return out


# --------------------------------------------------------------------------------------
# Monte-Carlo loop (fast/fast.py:115-140, 589-605, 639-668; fast/funcs.py:210-223, 352-365)
# --------------------------------------------------------------------------------------
def draw_complex(rng, shape):
    """Real block first, then the imaginary block (fast/funcs.py:352-356)."""
    return rng.normal(0, 1, size=shape) + 1j * rng.normal(0, 1, size=shape)


def draw_logamp(rng, niter, logamp_var):
    """chi_i = sigma_chi * Re(a_i + i b_i): niter real draws THEN niter imaginary draws that
    are discarded (fast/funcs.py:358-365; fast/fast.py:639-645)."""
    a = rng.normal(0, 1, size=(niter,))
    rng.normal(0, 1, size=(niter,))
    return a * np.sqrt(logamp_var)


def screens_from_noise(noise, W, df, lo, hi):
    """Colour by sqrt(W), centred inverse DFT scaled by (N df)^2 / N^2, stack Re then Im, crop
    (fast/fast.py:593-596; fast/funcs.py:218-221 with aotools.ift2).
    noise: (J/2, N, N) complex -> (J, Npup, Npup) float."""
    ax = (-1, -2)
    n = noise.shape[-1]
    spec = noise * np.sqrt(W) * df
    scr = np.fft.ifftshift(np.fft.ifft2(np.fft.ifftshift(spec, axes=ax)), axes=ax) * (n * 1) ** 2
    both = np.vstack([scr.real, scr.imag])
    return both[:, lo:hi, :][:, :, lo:hi]


def detector(phs, U, chi, coherent=False):
    """z_i = exp(chi_i) sum(U exp(i phi_i)) / sum(U); |z|^2 unless coherent
    (fast/fast.py:647-668; dx^2 cancels)."""
    z = np.exp(chi) * (U * np.exp(1j * phs)).sum((1, 2)) / U.sum()
    return z if coherent else np.abs(z) ** 2


def run_mc(init, rng, niter=None, nchunks=None, noise_hook=None):
    """The chunk loop of Fast.run() (fast/fast.py:115-140) on a build() dict.
    Returns the flattened per-realisation array (`FastResult._r`)."""
    p = init['params']
    niter = p['NITER'] if niter is None else niter
    nchunks = p['NCHUNKS'] if nchunks is None else nchunks
    if niter % nchunks != 0:
        raise Exception('NCHUNKS must divide NITER without remainder')
    J = niter // nchunks
    if J % 2 != 0:
        raise Exception('NITER/NCHUNKS must be even number')
    coherent = bool(p['COHERENT'])
    N = init['N']
    U = init['pupil'] * init['pupil_mode']
    chi = draw_logamp(rng, niter, init['logamp_var'])
    out = np.zeros((nchunks, J), dtype=complex if coherent else float)
    for c in range(nchunks):
        noise = draw_complex(rng, (J // 2, N, N))
        if noise_hook is not None:
            noise_hook(c, noise)
        phs = screens_from_noise(noise, init['powerspec'], init['df'], init['lo'], init['hi'])
        out[c] = detector(phs, U, chi[c * J:(c + 1) * J], coherent)
    return out.flatten(), chi


# --------------------------------------------------------------------------------------
# Device RNG contract (NEW -- no reference equivalent; restated here so the CUDA generator
# can be checked bit-for-bit).  Philox4x32-10 (Salmon et al., SC'11; same round function and
# constants as cuRAND's curand_philox4x32_x.h) + Box-Muller.
# --------------------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
STREAM_NOISE = 0x5CE7E000      # counter word 3 tag: phase-noise cells
STREAM_CHI = 0x10CA3900        # counter word 3 tag: log-amplitude draws


STREAM_NOISE_FAST = 0x5CE7F000  # counter word 3 tag: phase-noise cells of the 'device-fast' stream


def philox4x32_10(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R (R = 10 by default; the 'device-fast' stream uses R = 7).  Inputs
    broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    sh = np.uint64(32)
    for _ in range(rounds):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> sh, p0 & mask
        hi1, lo1 = p1 >> sh, p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


def u32_to_unit_open(u):
    """(0, 1]: 1 - (u >> 9) * 2^-23 -- exactly representable in fp32 (radius argument)."""
    return 1.0 - (np.asarray(u, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) * 2.0 ** -23


def u32_to_unit(u):
    """[0, 1): (u >> 9) * 2^-23 (angle fraction)."""
    return (np.asarray(u, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) * 2.0 ** -23


def box_muller(ua, ub):
    """(r cos t, r sin t), r = sqrt(-2 ln ua'), t = 2 pi ub'."""
    r = np.sqrt(-2.0 * np.log(u32_to_unit_open(ua)))
    t = 2.0 * np.pi * u32_to_unit(ub)
    return r * np.cos(t), r * np.sin(t)


def _bm_fields(mr, ma):
    """Box-Muller on 23-bit fields: radius argument 1 - mr 2^-23, angle 2 pi ma 2^-23."""
    r = np.sqrt(-2.0 * np.log(1.0 - mr.astype(np.float64) * 2.0 ** -23))
    t = 2.0 * np.pi * ma.astype(np.float64) * 2.0 ** -23
    return r * np.cos(t) + 1j * r * np.sin(t)


def noise_stride(N, n_pup=None):
    """Noise blocks per row S of the K2 device RNG (include/fastb.h, "Device RNG"): N / 16 for the
    radix sizes (powers of two 64..2048); M / 16 with M = 2^ceil(log2(N + n_pup - 1)) >= 64 for the
    other even N with N + n_pup - 1 <= 2048 (the chirp-z kernel); ceil(N / 16) otherwise, and when
    n_pup is not given (the K4 layer screens)."""
    if n_pup is None or (64 <= N <= 2048 and N & (N - 1) == 0):
        return (N + 15) // 16
    if N % 2 == 0 and N >= 4 and N + n_pup - 1 <= 2048:
        M = 64
        while M < N + n_pup - 1:
            M *= 2
        return M // 16
    return (N + 15) // 16


def device_noise_pair_fast(seed, pair, N, S=None):
    """The 'device-fast' stream (include/fastb.h, FASTB_RUN_RNG_FAST): same block / cell mapping,
    five Philox4x32-7 calls q with counter (b, pair lo, pair hi, STREAM_NOISE_FAST + q) give 20
    words; cell m owns word W[m] and byte m % 4 of the extra word W[16 + m // 4]:
    radius field = W[m] & 0x7FFFFF, angle field = ((W[m] >> 9) & 0x7FC000) ^ (byte << 8)."""
    S = (N + 15) // 16 if S is None else S
    b = (np.arange(N, dtype=np.uint64)[:, None] * np.uint64(S) + np.arange(S, dtype=np.uint64)[None, :])
    W = []
    for q in range(5):
        W.extend(philox4x32_10(b, np.uint64(pair & 0xFFFFFFFF), np.uint64((pair >> 32) & 0xFFFFFFFF),
                               np.uint64(STREAM_NOISE_FAST + q), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF,
                               rounds=7))
    out = np.zeros((N, S * 16), dtype=complex)
    t = np.arange(S)
    for m in range(16):
        w = W[m]
        byte = (W[16 + m // 4] >> np.uint32(8 * (m % 4))) & np.uint32(0xFF)
        mr = w & np.uint32(0x7FFFFF)
        ma = ((w >> np.uint32(9)) & np.uint32(0x7FC000)) ^ (byte << np.uint32(8))
        out[:, t + S * m] = _bm_fields(mr, ma)
    return out[:, :N]


def device_noise_pair(seed, pair, N, fast=False, S=None):
    """The complex white-noise tile the CUDA generator produces for global pair index `pair`
    (contract: include/fastb.h).  Noise block b = r*S + t holds cells (r, t + S m), m < 16 (those
    beyond the grid are dropped), S = noise_stride(N, n_pup) (default ceil(N/16)); six Philox calls q
    with counter (b, pair lo, pair hi, STREAM_NOISE + q) give 24 words; each word triple feeds two
    Box-Muller pairs (top 23 bits of each word, plus one field mixed from the three low 9-bit
    remainders)."""
    if fast:
        return device_noise_pair_fast(seed, pair, N, S)
    S = (N + 15) // 16 if S is None else S
    b = (np.arange(N, dtype=np.uint64)[:, None] * np.uint64(S) + np.arange(S, dtype=np.uint64)[None, :])
    W = []
    for q in range(6):
        w = philox4x32_10(b, np.uint64(pair & 0xFFFFFFFF), np.uint64((pair >> 32) & 0xFFFFFFFF),
                          np.uint64(STREAM_NOISE + q), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        W.extend(w)
    out = np.zeros((N, S * 16), dtype=complex)
    t = np.arange(S)
    nine, five = np.uint32(0x1FF), np.uint32(0x1F)
    for G in range(8):
        a, bb, c = W[3 * G], W[3 * G + 1], W[3 * G + 2]
        mix = ((a & nine) << np.uint32(14)) | ((bb & nine) << np.uint32(5)) | ((c >> np.uint32(4)) & five)
        out[:, t + S * (2 * G)] = _bm_fields(a >> np.uint32(9), bb >> np.uint32(9))
        out[:, t + S * (2 * G + 1)] = _bm_fields(c >> np.uint32(9), mix)
    return out[:, :N]


def device_chi_normals(seed, first, count):
    """Standard normals for the log-amplitude of realisations [first, first+count): Philox
    call i = index // 4 with counter (i_lo, i_hi, 0, STREAM_CHI); the 4 words give
    (n0, n1) = BM(w0, w1), (n2, n3) = BM(w2, w3); realisation index -> n[index % 4]."""
    idx = np.arange(first, first + count, dtype=np.uint64)
    call = idx >> np.uint64(2)
    w = philox4x32_10(call & np.uint64(0xFFFFFFFF), call >> np.uint64(32), np.uint64(0),
                      np.uint64(STREAM_CHI), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    n0, n1 = box_muller(w[0], w[1])
    n2, n3 = box_muller(w[2], w[3])
    return np.choose((idx & np.uint64(3)).astype(int), [n0, n1, n2, n3])


def run_mc_device_rng(init, seed, n_pairs, pairs_per_chunk, coherent=None, fast=False, noise_of=None):
    """Oracle of the CUDA path in device-RNG mode for global pairs [0, n_pairs): same
    realisation layout as Fast.run() (chunk-major, Re-half then Im-half inside a chunk) with
    noise from device_noise_pair and chi from device_chi_normals."""
    p = init['params']
    coherent = bool(p['COHERENT']) if coherent is None else coherent
    N = init['N']
    U = init['pupil'] * init['pupil_mode']
    chi_all = math.sqrt(init['logamp_var']) * device_chi_normals(seed, 0, 2 * n_pairs)
    out = np.zeros(2 * n_pairs, dtype=complex if coherent else float)
    for g in range(n_pairs):
        chunk, pp = divmod(g, pairs_per_chunk)
        i_re = chunk * 2 * pairs_per_chunk + pp
        i_im = i_re + pairs_per_chunk
        # noise_of(g): the tile dumped from the device (fastb_rng_dump) instead of the restatement,
        # which removes the ~1e-6 MUFU difference of the noise itself from the comparison
        noise = (noise_of(g) if noise_of is not None
                 else device_noise_pair(seed, g, N, fast, noise_stride(N, init['Npup'])))[None]
        phs = screens_from_noise(noise, init['powerspec'], init['df'], init['lo'], init['hi'])
        r = detector(phs, U, chi_all[[i_re, i_im]], coherent)
        out[i_re], out[i_im] = r[0], r[1]
    return out


# --------------------------------------------------------------------------------------
# TEMPORAL (frozen-flow) path: fast/fast.py:181-206,217-219,394-405,538-587,607-637,846-875;
# fast/funcs.py:367-375.  L per-layer screens are drawn once; each time step samples them at the
# pupil coordinates shifted by the layer wind, bilinearly, and sums the layers.
# --------------------------------------------------------------------------------------
def temporal_setup(init):
    """Everything TEMPORAL adds to Fast.__init__: pixel shifts per step, the per-layer temporal
    frequency grids, the elongated pupil filter spline and the temporal log-amplitude PSD."""
    from scipy.interpolate import RectBivariateSpline
    p, atm = init['params'], init['atm']
    N, dx, npup, df = init['N'], init['dx'], init['Npup'], init['df']
    niter, J = p['NITER'], p['NITER'] // p['NCHUNKS']
    h, cn2, wind = atm['h'], atm['cn2'], atm['wind_vector']
    L = len(h)
    shifts = (np.arange(1, J + 1) * p['DT']) * wind[..., None] / dx            # (L, 2, J)  fast.py:543-544

    # per-layer frequency grids: x axis LINEAR frequency 1/(Niter v dt) (sic), y axis = main axis,
    # rotated by the wind direction (fast.py:846-864, 903-907)
    fxa = np.array([np.arange(-niter / 2, niter / 2) * (1 / (niter * atm['wind_speed'][i] * p['DT']))
                    for i in range(L)])
    fya = np.array([np.arange(-N / 2, N / 2) * df for _ in range(L)])
    rot = np.radians(atm['wind_dir'])
    fabs = np.zeros((L, N, niter))
    for i in range(L):
        gx, gy = np.meshgrid(fxa[i], fya[i])
        xr = gx * np.cos(rot[i]) - gy * np.sin(rot[i])
        yr = gx * np.sin(rot[i]) + gy * np.cos(rot[i])
        fabs[i] = np.sqrt(xr ** 2 + yr ** 2)

    # elongated high-resolution pupil filter, as a bilinear spline (fast.py:394-405)
    f_max = max(fxa.max(), fya.max())
    dx_req = np.pi / f_max
    n_req = int(2 * np.ceil(2 * np.pi / (df * dx_req) / 2))
    ny = 2 * npup
    D, obsc = p['D_GROUND'], p['OBSC_GROUND']
    ap = disc(D / dx_req / 2, n_req) - disc(obsc / dx_req / 2, n_req)
    assert (ny - n_req) % 2 == 0
    if ny > n_req:
        pad = (ny - n_req) // 2
        ap = np.pad(ap, [(0, 0), (pad, pad)])
    elif ny < n_req:
        cut = (n_req - ny) // 2
        ap = ap[:, cut:-cut]
    pup_t = ap / np.sqrt(ap.sum() * dx_req ** 2)
    W0 = init['W0']
    yy, xx = np.meshgrid(np.arange(ny), np.arange(n_req))          # gaussian2d((n_req, ny), w)
    w = W0 / dx_req / np.sqrt(2)
    mode_t = np.exp(-(((ny / 2. - yy) / w) ** 2 + ((n_req / 2. - xx) / w) ** 2) / 2) \
        * np.sqrt(2 / (np.pi * W0 ** 2)) / pup_t.max()
    pf_t = pupil_filter(pup_t * mode_t)
    fxl = np.arange(-n_req / 2., n_req / 2.) * (2 * np.pi / (n_req * dx_req))
    fyl = np.arange(-ny / 2., ny / 2.) * (2 * np.pi / (ny * dx))
    spline = RectBivariateSpline(fxl, fyl, pf_t, kx=1, ky=1, s=0)

    # temporal log-amplitude PSD, integrated over the axis orthogonal to the wind (fast.py:582-587);
    # the spline is sampled at (fy_axis, fx_axis) un-rotated, as the reference does
    total = np.zeros((N, niter))
    k = init['k']
    for i in range(L):
        ps = von_karman_base(fabs[i], p['L0'], p['l0']) * cn2[i] * 2 * np.pi * k ** 2
        ps = ps * np.sin(p['WVL'] * h[i] * fabs[i] ** 2 / (4 * np.pi)) ** 2
        total += ps * spline(fya[i], fxa[i])
    tps = total.sum(-2) * df
    return dict(pixel_shifts=shifts, temporal_logamp_powerspec=tps, J=J)


def temporal_logamp(rng, niter, logamp_var, tps):
    """Temporally coloured log-amplitude: centred FFT of coloured complex noise, real part
    (fast/funcs.py:367-375 with aotools.ft)."""
    rf = rng.normal(0, 1, size=(niter,)) + 1j * rng.normal(0, 1, size=(niter,))
    rf = rf * np.sqrt(tps / tps.sum())
    r = np.fft.fftshift(np.fft.fft(np.fft.fftshift(rf)))
    return (r * np.sqrt(logamp_var)).real


def layer_screens(noise, per_layer, df):
    """L real screens from (L, N, N) complex noise coloured per layer (fast/fast.py:611-614,
    make_phase_fft double=False)."""
    ax = (-1, -2)
    n = noise.shape[-1]
    spec = noise * np.sqrt(per_layer) * df
    return (np.fft.ifftshift(np.fft.ifft2(np.fft.ifftshift(spec, axes=ax)), axes=ax) * n ** 2).real


def temporal_sample_coords(interp, N):
    """Per layer/axis/step: the coordinates at which the reference evaluates the screen for
    each output pixel (fast/fast.py:621-633): wrap mod N, sort, and 'un-sort' by rolling with
    the argmax of the gaps (0 when there is no wrap).  NOTE the roll is one short of restoring
    the original order when a wrap occurs -- reproduced here because it is what the reference
    computes."""
    coord = np.sort(interp % N, axis=-1)
    gaps = np.abs(np.diff(coord, axis=-1))
    sh = gaps.argmax(-1)
    sh[np.isclose(gaps, 1).all(-1)] = 0
    npup = coord.shape[-1]
    idx = (np.arange(npup) + sh[..., None]) % npup
    return np.take_along_axis(coord, idx, axis=-1)


def bilinear_clamped(scr, x, y):
    """Grid evaluation of a degree-1 spline on knots 0..N-1 at rows x, columns y; arguments
    beyond N-1 are clamped (FITPACK bispev behaviour of RectBivariateSpline(kx=ky=1, s=0))."""
    n = scr.shape[0]
    def split(c):
        c = np.minimum(c, n - 1.0)
        i0 = np.minimum(np.floor(c).astype(int), n - 2)
        return i0, c - i0
    ix, fx = split(x)
    iy, fy = split(y)
    a = scr[np.ix_(ix, iy)] * (1 - fy) + scr[np.ix_(ix, iy + 1)] * fy
    b = scr[np.ix_(ix + 1, iy)] * (1 - fy) + scr[np.ix_(ix + 1, iy + 1)] * fy
    return a * (1 - fx)[:, None] + b * fx[:, None]


def run_mc_temporal(init, rng, screens_hook=None):
    """Fast.run() in TEMPORAL mode (fast/fast.py:115-140, 607-637).  Returns (_r, chi, last phs)."""
    p = init['params']
    niter, nch = p['NITER'], p['NCHUNKS']
    ts = temporal_setup(init)
    J, shifts = ts['J'], ts['pixel_shifts']
    N, npup, lo, hi = init['N'], init['Npup'], init['lo'], init['hi']
    coherent = bool(p['COHERENT'])
    U = init['pupil'] * init['pupil_mode']
    chi = temporal_logamp(rng, niter, init['logamp_var'], ts['temporal_logamp_powerspec'])
    L = len(init['atm']['h'])
    noise = draw_complex(rng, (L, N, N))
    scr = layer_screens(noise, init['powerspec_per_layer'], init['df'])
    if screens_hook is not None:
        screens_hook(scr)
    base = np.arange(lo, hi).astype(float)
    interp = base[None, None, None, :] + shifts[:, :, :, None]                 # (L, 2, J, Pp)
    out = np.zeros((nch, J), dtype=complex if coherent else float)
    phs = None
    for c in range(nch):
        at = temporal_sample_coords(interp, N)
        phs = np.zeros((J, npup, npup))
        for i in range(L):
            for j in range(J):
                phs[j] += bilinear_clamped(scr[i], at[i, 0, j], at[i, 1, j])
        out[c] = detector(phs, U, chi[c * J:(c + 1) * J], coherent)
        interp = interp + shifts[:, :, -1, None, None]
    return out.flatten(), chi, phs


# --------------------------------------------------------------------------------------
# Sub-harmonics (Lane et al. style low-frequency correction): fast/fast.py:494-531, 598-603,
# 835-844; fast/funcs.py:225-258.  Three levels p = 1..3 of 3 x 3 frequencies spaced
# 2 pi / (3^p N dx); the residual PSD is evaluated on them with the same terms as the main grid.
# --------------------------------------------------------------------------------------
def subharm_axes(N, dx, pmax=3):
    """(pmax, 3) frequency axes {-1, 0, 1} * 2 pi / (3^p N dx) (fast/fast.py:835-844)."""
    D = dx * N
    return np.array([np.arange(-1, 2) * (2 * np.pi / (3 ** p * D)) for p in range(1, pmax + 1)])


def subharm_psd(init):
    """powerspec_subharm_per_layer (L, 3, 3, 3) and its layer sum, by applying the main-grid
    terms to each 3 x 3 level (fast/fast.py:494-523)."""
    p, atm = init['params'], init['atm']
    axes = subharm_axes(init['N'], init['dx'])
    h, cn2, wind = atm['h'], atm['cn2'], atm['wind_vector']
    L = len(h)
    mode_name = p['AO_MODE']
    zmax, modal, mmult = p['ZMAX'], p['MODAL'], p['MODAL_MULT']
    if mode_name == 'TT':
        zmax, modal, mmult = 3, True, 1
    k = init['k']
    out = np.zeros((L, 3, 3, 3))
    for i, ax in enumerate(axes):
        fx, fy = np.meshgrid(ax, ax)
        fabs = np.sqrt(fx ** 2 + fy ** 2)
        mask = lf_mask(fx, fy, p['DSUBAP'], modal=modal, modal_mult=mmult, Zmax=zmax, D=p['D_GROUND'])
        turb = von_karman(fabs, cn2, p['L0'], p['l0'])
        G = g_ao(fx, fy, mask, mode_name, h, wind, atm['dtheta'], p['TLOOP'], p['TEXP'], D=p['D_GROUND'])
        if p['ALIAS'] and mode_name != 'NOAO':
            alias = alias_psd(ax, ax, p['DSUBAP'], cn2, mask, wind, p['TEXP'], p['L0'], p['l0'])
        else:
            alias = 0.
        if p['NOISE'] > 0 and mode_name != 'NOAO':
            noise = noise_psd(fx, fy, p['DSUBAP'], p['NOISE'], mask)
        else:
            noise = 0.
        out[:, i] = 2 * np.pi * k ** 2 * (turb * G + alias) + noise / L
    return out, out.sum(0), axes


def subharm_screens(noise_lo, W_sh, axes, N, dx):
    """make_phase_subharm(double=True) (fast/funcs.py:225-258): noise_lo (J/2, 3, 3, 3) complex
    -> (J, N, N): sum of the 27 plane waves, full-grid mean removed per complex screen,
    Re stacked over Im."""
    D = dx * N
    coords = np.arange(-D / 2, D / 2, dx)[:N]
    x, y = np.meshgrid(coords, coords)
    acc = np.zeros((noise_lo.shape[0], N, N), dtype=complex)
    for i in range(axes.shape[0]):
        df_lo = axes[i, 1] - axes[i, 0]
        fx_lo, fy_lo = np.meshgrid(axes[i], axes[i])
        amp = noise_lo[:, i] * np.sqrt(W_sh[i]) * df_lo                      # (J/2, 3, 3)
        modes = np.exp(1j * (x[None, None] * fx_lo[..., None, None] + y[None, None] * fy_lo[..., None, None]))
        acc = acc + np.einsum('pqs,qsrc->prc', amp, modes)
    acc = acc - acc.mean((1, 2))[:, None, None]
    return np.vstack([acc.real, acc.imag])


def run_mc_subharm(init, rng):
    """Fast.run() with SUBHARM=True (fast/fast.py:589-605): per chunk the main noise block is
    drawn first, then the (J/2, 3, 3, 3) sub-harmonic block.  Returns (_r, last chunk phs)."""
    p = init['params']
    niter, nch = p['NITER'], p['NCHUNKS']
    J = niter // nch
    N, dx, lo, hi = init['N'], init['dx'], init['lo'], init['hi']
    _, W_sh, axes = subharm_psd(init)
    U = init['pupil'] * init['pupil_mode']
    chi = draw_logamp(rng, niter, init['logamp_var'])
    out = np.zeros((nch, J), dtype=complex if p['COHERENT'] else float)
    phs = None
    for c in range(nch):
        noise = draw_complex(rng, (J // 2, N, N))
        phs = screens_from_noise(noise, init['powerspec'], init['df'], lo, hi)
        noise_lo = draw_complex(rng, (J // 2, 3, 3, 3))
        phs = phs + subharm_screens(noise_lo, W_sh, axes, N, dx)[:, lo:hi, :][:, :, lo:hi]
        out[c] = detector(phs, U, chi[c * J:(c + 1) * J], bool(p['COHERENT']))
    return out.flatten(), phs


STREAM_SUBHARM = 0x5AB4A200


def device_subharm_noise(seed, pair):
    """(3, 3, 3) complex unit normals the CUDA generator uses for the sub-harmonics of global
    pair `pair`: call j < 14, counter (j, pair lo, pair hi, STREAM_SUBHARM) -> flat amplitudes
    2j (words 0,1) and 2j+1 (words 2,3), flat index m = (level*3 + q)*3 + s.  (include/fastb.h)"""
    j = np.arange(14, dtype=np.uint64)
    w = philox4x32_10(j, np.uint64(pair & 0xFFFFFFFF), np.uint64((pair >> 32) & 0xFFFFFFFF),
                      np.uint64(STREAM_SUBHARM), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    re0, im0 = box_muller(w[0], w[1])
    re1, im1 = box_muller(w[2], w[3])
    flat = np.empty(28, dtype=complex)
    flat[0::2] = re0 + 1j * im0
    flat[1::2] = re1 + 1j * im1
    return flat[:27].reshape(3, 3, 3)


def run_mc_device_rng_subharm(init, seed, n_pairs, pairs_per_chunk):
    """run_mc_device_rng with the sub-harmonic term (device noise restated)."""
    p = init['params']
    coherent = bool(p['COHERENT'])
    N, dx, lo, hi = init['N'], init['dx'], init['lo'], init['hi']
    _, W_sh, axes = subharm_psd(init)
    U = init['pupil'] * init['pupil_mode']
    chi_all = math.sqrt(init['logamp_var']) * device_chi_normals(seed, 0, 2 * n_pairs)
    out = np.zeros(2 * n_pairs, dtype=complex if coherent else float)
    for g in range(n_pairs):
        chunk, pp = divmod(g, pairs_per_chunk)
        i_re = chunk * 2 * pairs_per_chunk + pp
        i_im = i_re + pairs_per_chunk
        phs = screens_from_noise(device_noise_pair(seed, g, N, S=noise_stride(N, init['Npup']))[None], init['powerspec'],
                                 init['df'], lo, hi)
        phs = phs + subharm_screens(device_subharm_noise(seed, g)[None], W_sh, axes, N, dx)[:, lo:hi, :][:, :, lo:hi]
        r = detector(phs, U, chi_all[[i_re, i_im]], coherent)
        out[i_re], out[i_im] = r[0], r[1]
    return out
