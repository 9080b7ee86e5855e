"""`Fast` / `FastResult`: the reference's public object protocol (fast/fast.py:20-140,931-1002)
over the B200 CUDA library.

    import fast_b200 as fast
    sim = fast.Fast(p)          # p: dict or path to a .py config, same keys as ojdf/fast
    res = sim.run()             # FastResult: .power .dB_rel .dB_abs .dBm .scintillation_index

What runs where
  host (numpy, once per config): config, geometry scalars, pupil and fibre mode, link budget
  device (libfastb.so):          residual-PSD build + Simpson integrals + pupil filter
                                 (fastb_psd_build / fastb_simpson2d / fastb_pupil_filter),
                                 noise -> screens -> detector (fastb_screen_detect),
                                 result statistics (fastb_stats)
There is no CPU fallback for the device part.
"""
import functools
import logging
import math

import numpy
import torch

from . import _lib
from . import ao_power_spectra
from . import conf
from . import dist
from . import funcs
from . import temporal

logger = logging.getLogger(__name__)

# the reference's optional-dependency flag (fast/fast.py:9-15, read by test/tests_pytest.py:50): pyFFTW is never
# used on this path
_pyfftw = False

_AO_MODES = {'NOAO': _lib.AO_NOAO, 'AO': _lib.AO_AO, 'TT': _lib.AO_AO, 'LGSAO': _lib.AO_LGSAO}
_RNG_MODES = ('device', 'device-fast', 'numpy')
_GOLDEN64 = 0x9E3779B97F4A7C15
# Aperture and fibre mode depend only on (N, dx, D, obscuration, W0 request, mode type): the samples of a pass
# sweep (one Fast per elevation, fast/complete_orbit_simulation.py:217-228) share them, and the W0 optimisation
# over N x N arrays is the bulk of the host-side construction time (10 ms at N = 256, 200 ms at N = 1024).
_PUPIL_CACHE = {}
_PUPIL_CACHE_MAX = 8


def _on_device(method):
    """Run a method with the simulation's CUDA device current: the library launches on the current
    device and stream, so a Fast built for 'cuda:1' must not depend on what the caller left current."""
    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapped


class SpatialFrequencyStruct():
    """Centred angular-frequency grid (fast/fast.py:877-921): fx[r, c] = fx_axis[c],
    fy[r, c] = fy_axis[r]; 2-D axes give one (optionally rotated) grid per layer.  The 2-D
    arrays are built on first use; the device never reads them."""

    def __init__(self, fx_axis, fy_axis=None, rot=None, freq_per_layer=False):
        self.fx_axis = fx_axis
        self.fy_axis = fx_axis if fy_axis is None else fy_axis
        self.freq_per_layer = freq_per_layer
        if fy_axis is None:
            self.f = fx_axis
            self.df = fx_axis[..., 1] - fx_axis[..., 0]
        self.dfx = fx_axis[..., 1] - fx_axis[..., 0]
        self.dfy = self.fy_axis[..., 1] - self.fy_axis[..., 0]
        if fx_axis.ndim not in (1, 2):
            raise Exception('fx_axis ndim sould be either 1 or 2')
        self._rot = rot
        self._grid = None

    def _mesh(self):
        if self._grid is None:
            def one(ax, ay, angle):
                gx, gy = numpy.meshgrid(ax, ay)
                if angle is None:
                    return gx, gy
                return (gx * numpy.cos(angle) - gy * numpy.sin(angle),
                        gx * numpy.sin(angle) + gy * numpy.cos(angle))
            if self.fx_axis.ndim == 1:
                self._grid = one(self.fx_axis, self.fy_axis, self._rot)
            else:
                pairs = [one(self.fx_axis[i], self.fy_axis[i], None if self._rot is None else self._rot[i])
                         for i in range(self.fx_axis.shape[0])]
                self._grid = (numpy.array([a for a, _ in pairs]), numpy.array([b for _, b in pairs]))
        return self._grid

    @property
    def fx(self):
        return self._mesh()[0]

    @property
    def fy(self):
        return self._mesh()[1]

    @property
    def fabs(self):
        return numpy.sqrt(self.fx ** 2 + self.fy ** 2)

    def realspace_sampling(self):
        """Pixel scales of the real-space grid conjugate to this one (fast/fast.py:923-928)."""
        Nx, Ny = self.fx.shape[-1], self.fx.shape[-2]
        return 2 * numpy.pi / (Nx * self.dfx), 2 * numpy.pi / (Ny * self.dfy)


class SpatialFrequencies():
    """fast/fast.py:814-875: `main` grid with df = 2 pi / (N dx); `temporal` per-layer grids."""

    def __init__(self, N, dx):
        self.N, self.dx = N, dx
        self.main = SpatialFrequencyStruct(numpy.arange(-N / 2., N / 2.) * (2 * numpy.pi / (N * dx)))
        self.f = self.main.f
        self.df = self.main.df

    fx = property(lambda self: self.main.fx)
    fy = property(lambda self: self.main.fy)
    fabs = property(lambda self: self.main.fabs)

    def make_main_freqs(self, N, dx):
        """fast/fast.py:830-833 (the constructor already calls it)."""
        self.main = SpatialFrequencyStruct(numpy.arange(-N / 2., N / 2.) * (2 * numpy.pi / (N * dx)))

    def make_logamp_freqs(self, Nx=None, dx=None, Ny=None, dy=None):
        """Grid of the log-amplitude PSD: the main grid unless a sampling is given (fast/fast.py:866-875)."""
        if Nx is None and dx is None:
            self.logamp = self.main
        else:
            self.logamp = SpatialFrequencyStruct(numpy.arange(-Nx / 2., Nx / 2.) * (2 * numpy.pi / (Nx * dx)),
                                                 numpy.arange(-Ny / 2., Ny / 2.) * (2 * numpy.pi / (Ny * dy)))

    def make_subharm_freqs(self, pmax=3):
        """Three 3 x 3 levels spaced 2 pi / (3^p N dx) (fast/fast.py:835-844)."""
        D = self.dx * self.N
        self.subharm = SpatialFrequencyStruct(
            numpy.array([numpy.arange(-1, 2) * (2 * numpy.pi / (3 ** p * D)) for p in range(1, pmax + 1)]))

    def make_temporal_freqs(self, nlayer, Ny, Nx, wind_speed, wind_dir, dt):
        fx_axes, fy_axes = temporal.temporal_axes(nlayer, Ny, Nx, wind_speed, dt, self.main.dfy)
        self.temporal = SpatialFrequencyStruct(fx_axes, fy_axes, rot=numpy.radians(wind_dir),
                                               freq_per_layer=True)


class Fast():
    """Drop-in for `fast.Fast` on the Monte-Carlo path (fast/fast.py:20-140).

    Readable attributes follow the reference: I, result, link_budget, diffraction_limit,
    powerspec, powerspec_per_layer, logamp_powerspec, logamp_var, phs_var, phs_var_weights,
    aniso_servo_error, alias_error, noise_error, fitting_error, r0, theta0, tau0, *_los,
    pupil, pupil_mode, W0, dx, Npxls, Npxls_pup, freq, L, h, cn2, wind_vector, params.
    Array-valued PSD attributes are fetched from the device on first access."""

    def __init__(self, params):
        self.conf = conf.ConfigParser(params)
        self.params = self.conf.config

        self.Niter = self.params['NITER']
        self.Nchunks = self.params['NCHUNKS']
        self.fftw = False                      # no FFTW on this path; the key is accepted
        self.nthreads = self.params['FFTW_THREADS']
        self.seed = self.params['SEED']
        if self.seed != None:  # noqa: E711  (reference semantics: 0 is a valid seed)
            self.set_seed(self.seed)
        self.temporal = self.params['TEMPORAL']
        self.dt = self.params['DT']
        self.rng_mode = self.params.get('RNG', conf.EXTRA_DEFAULTS['RNG'])
        if self.rng_mode not in _RNG_MODES:
            raise Exception("RNG must be 'device', 'device-fast' or 'numpy'")

        if self.Niter % self.Nchunks != 0:
            raise Exception('NCHUNKS must divide NITER without remainder')
        self.Niter_per_chunk = self.Niter // self.Nchunks
        if not (self.Niter_per_chunk % 2 == 0) and not self.temporal:
            raise Exception('NITER/NCHUNKS must be even number')
        _lib.require_cuda()
        dev = self.params.get('DEVICE', None)
        self.device = torch.device(dev) if dev is not None else torch.device('cuda', torch.cuda.current_device())
        self._d = {}                           # device tensors by name
        self._host_cache = {}
        self._runs = 0                         # run() calls so far; _run_index = the current / last one
        self._run_index = 0
        self._prep_key = None

        with torch.cuda.device(self.device):
            self.init_logging()
            self.init_atmos()
            self.init_beam_params()
            self.init_frequency_grid()
            self.init_ao_params()
            self.init_pupil_mask()
            self.init_phs_logamp()
            self.compute_link_budget()
            self.compute_powerspec()
        self.fftw_objs = None

    # ------------------------------------------------------------------ init (host scalars)
    def init_logging(self):
        logging.basicConfig(filename=self.params['LOGFILE'],
                            level=logging.getLevelName(self.params['LOGLEVEL']),
                            format="[%(levelname)s] %(name)s.%(funcName)s | %(message)s")

    def calc_zenith_correction(self, zenith_angle):
        return 1 / numpy.cos(numpy.radians(zenith_angle))

    def set_seed(self, seed):
        funcs._R = numpy.random.default_rng(seed)

    def init_atmos(self):
        """Line-of-sight geometry per layer (fast/fast.py:229-276)."""
        p = self.params
        g = self.zenith_correction = self.calc_zenith_correction(p['ZENITH_ANGLE'])
        self.h = p['H_TURB'] * g
        self.cn2 = p['CN2_TURB'] * g
        self.L = p['L_SAT'] if p['L_SAT'] != None else funcs.l_path(p['H_SAT'], p['ZENITH_ANGLE'])  # noqa: E711
        self.dtheta = p['DTHETA']
        self.paa = numpy.sqrt(self.dtheta[0] ** 2 + self.dtheta[1] ** 2)

        self.wind_dir = p['WIND_DIR']
        if 'AZIMUT_SAT' in p:
            # modulus 380 is the reference's (fast/fast.py:250); kept for parity
            self.wind_dir = [(x - p['AZIMUT_SAT']) % 380 for x in self.wind_dir]
        ang = numpy.radians(self.wind_dir)
        self.wind_vector = (p['WIND_SPD'] * numpy.array([numpy.cos(ang), numpy.sin(ang) / g])).T
        if 'ANISO_DL' in p:
            self.wind_correction = funcs.calculate_wind_correction(self.h, p['ANISO_DL'], p['TLOOP'])
            self.wind_vector = self.wind_vector + self.wind_correction
        self.wind_speed = numpy.hypot(self.wind_vector[:, 0], self.wind_vector[:, 1])

        def r0_of(cn2_sum, lam):
            return (0.423 * (2 * numpy.pi / lam) ** 2 * cn2_sum) ** (-3. / 5.)

        def theta0_of(cn2, h, lam):
            # arcseconds, like aotools.isoplanaticAngle (attribute / FITS header only, not used by the MC)
            return 0.057 * lam ** (6. / 5.) * numpy.sum(cn2 * h ** (5. / 3.)) ** (-3. / 5.) * 180. * 3600. / numpy.pi

        def tau0_of(cn2, v, lam):
            return float(numpy.sum(cn2 * v ** (5. / 3.)) ** (-3. / 5.) * 0.057 * lam ** (6. / 5.))

        def rytov_of(cn2, h, lam):
            return 2.25 * (2 * numpy.pi / lam) ** (7. / 6.) * numpy.sum(cn2 * h ** (5. / 6.))

        cn2_0, h_0, w_0 = (numpy.asarray(p[k], dtype=float) for k in ('CN2_TURB', 'H_TURB', 'WIND_SPD'))
        self.r0 = r0_of(cn2_0.sum(), 500e-9)
        self.theta0 = theta0_of(cn2_0, h_0, 500e-9)
        self.tau0 = tau0_of(cn2_0, w_0, 500e-9)
        self.rytov_variance = rytov_of(cn2_0, h_0, 500e-9)
        wvl = p['WVL']
        self.r0_los = r0_of(self.cn2.sum(), wvl)
        self.theta0_los = theta0_of(self.cn2, self.h, wvl)
        self.tau0_los = tau0_of(self.cn2, self.wind_speed, wvl)
        self.rytov_variance_los = rytov_of(self.cn2, self.h, wvl)
        self.L0 = p['L0']
        self.l0 = p['l0']

    def init_beam_params(self):
        p = self.params
        self.power = p['POWER']
        self.W0 = p['W0']
        self.F0 = numpy.inf
        self.wvl = p['WVL']
        self.k = 2 * numpy.pi / self.wvl
        self.D_ground, self.obsc_ground = p['D_GROUND'], p['OBSC_GROUND']
        self.D_sat, self.obsc_sat = p['D_SAT'], p['OBSC_SAT']

    def init_frequency_grid(self):
        """DX / NPXLS 'auto' rules and the pupil window (fast/fast.py:147-227)."""
        p = self.params
        if p['DX'] == 'auto':
            self.dx = numpy.min([p['DSUBAP'] / 2, self.r0_los / 2, self.D_ground / 10])
            if p['AO_MODE'] == 'NOAO':
                self.dx = self.r0_los / 2
            logger.info(f"Auto set DX to {self.dx}")
        else:
            self.dx = p['DX']

        if p['NPXLS'] == 'auto':
            nyq = numpy.min([numpy.pi / (self.h[-1] * self.paa / 206265.),       # anisoplanatism
                             numpy.pi / (max(self.wind_speed) * p['TLOOP']),      # servo lag
                             numpy.pi / p['DSUBAP'] / 5])                         # corrected region
            n_nyq = int(2 * numpy.ceil(2 * numpy.pi / (nyq * self.dx) / 2))
            n_ap = int(2 * numpy.ceil(p['D_GROUND'] / self.dx / 2)) + 2
            n_t = 0
            if p['TEMPORAL']:       # enough pixels for the wind not to wrap during the run
                n_t = int(p['WIND_SPD'].max() * p['DT'] * p['NITER'] / p['DX'] / 2)
            self.Npxls = int(numpy.max([n_nyq, n_ap, n_t]))
            logger.info(f"Auto set NPXLS to {self.Npxls}")
            if p['AO_MODE'] == 'NOAO' and not numpy.isinf(p['L0']):
                n_L0 = int(2 * numpy.ceil((p['L0'] * 2) / self.dx) / 2)
                if n_L0 > self.Npxls:
                    logger.warning(f"L0 set with NOAO mode, low orders may be undersampled. "
                                   f"Recommended NPXLS: {n_L0}")
        else:
            self.Npxls = int(p['NPXLS'])
            if p['TEMPORAL']:
                n_t = int(p['WIND_SPD'].max() * p['DT'] * p['NITER'] / p['DX'] / 2)
                if self.Npxls < n_t:
                    logger.warning("NPXLS is likely too small -- some periodicity may occur in your "
                                   "resulting time series")
                    logger.warning(f"Current value: {self.Npxls}")
                    logger.warning(f"Recommended value: {n_t}")
        if self.Npxls % 2:
            raise Exception('NPXLS must be even')
        if self.Npxls > 2048:
            logger.warning(f"NPXLS is large ({self.Npxls}) and may cause very high memory usage")
        self.Npxls_pup = int(numpy.ceil(self.D_ground / self.dx)) + 2
        if self.Npxls_pup > self.Npxls:
            raise Exception('aperture does not fit in the grid: increase NPXLS or DX')
        self.freq = SpatialFrequencies(self.Npxls, self.dx)
        self.subharmonics = False
        if self.temporal:
            self.freq.make_temporal_freqs(len(self.h), self.Npxls, self.Niter, self.wind_speed,
                                          self.wind_dir, self.dt)
            if p['SUBHARM']:
                logger.info("SUBHARM not used in TEMPORAL mode")
        elif p['SUBHARM']:
            self.subharmonics = True
            self.freq.make_subharm_freqs()

    def init_ao_params(self):
        p = self.params
        self.ao_mode = p['AO_MODE']
        if self.ao_mode not in _AO_MODES:
            raise Exception('Mode not recognised, note that "AO_PA", "TT_PA" and "LGS_PA" are now '
                            '"AO" and "TT" and "LGSAO')
        self.Dsubap, self.tloop, self.texp = p['DSUBAP'], p['TLOOP'], p['TEXP']
        self.Zmax, self.alias, self.noise = p['ZMAX'], p['ALIAS'], p['NOISE']
        self.modal, self.modal_mult = p['MODAL'], p['MODAL_MULT']
        if self.ao_mode == 'TT':
            self.Zmax, self.modal, self.modal_mult = 3, True, 1   # tip/tilt = modal, Noll 1..3

    def _modal_mask_device(self, n, df):
        """mask_lf(modal=True) on the n x n grid of spacing df (fast/ao_power_spectra.py:119-141),
        evaluated on the device (Bessel-based Zernike filter when ZMAX is given)."""
        if self.Zmax is None:
            return _lib.zernike_filter(n, df, self.device, noll_first=1, noll_last=0, d_wfs=self.Dsubap,
                                       modal_mult=self.modal_mult, clip_box=True)
        return _lib.zernike_filter(n, df, self.device, noll_first=1, noll_last=self.Zmax,
                                   diameter=self.D_ground, d_wfs=self.Dsubap, clip_box=True)

    def _lgs_filter_device(self, n, df):
        """Zernike(1..4) squared filter of the LGSAO branch (fast/ao_power_spectra.py:262-267)."""
        return _lib.zernike_filter(n, df, self.device, noll_first=1, noll_last=4, diameter=self.D_ground)

    @property
    def lf_mask(self):
        """Corrected-region mask as a host array (fast/fast.py:317-319).  Zonal masks are
        recomputed inside K1 from fx, fy; modal ones come from fastb_zernike_filter."""
        if 'lf_mask' not in self._host_cache:
            if self.modal:
                m = self._modal_mask_device(self.Npxls, self.freq.main.df).cpu().numpy()
            else:
                m = ao_power_spectra.mask_lf(self.freq.main, self.Dsubap)
            self._host_cache['lf_mask'] = m
        return self._host_cache['lf_mask']

    @property
    def hf_mask(self):
        return 1 - self.lf_mask

    def init_pupil_mask(self):
        """Aperture, fibre mode, crop window (fast/fast.py:332-392); host numpy, once."""
        N, dx = self.Npxls, self.dx
        self.dx_sat = self.D_sat / 32
        ptype = 'axicon' if self.params['AXICON'] else 'gauss'
        key = (int(N), float(dx), float(self.D_ground), float(self.obsc_ground), repr(self.W0), ptype)
        hit = _PUPIL_CACHE.get(key)
        if hit is None:
            pupil_full = funcs.compute_pupil(N, dx, self.D_ground, self.obsc_ground)
            mode_full, w0 = funcs.compute_gaussian_mode(pupil_full, dx, self.W0, D=self.D_ground,
                                                        obsc=self.obsc_ground, ptype=ptype)
            pupil_full.flags.writeable = mode_full.flags.writeable = False
            if len(_PUPIL_CACHE) >= _PUPIL_CACHE_MAX:
                _PUPIL_CACHE.pop(next(iter(_PUPIL_CACHE)))
            hit = _PUPIL_CACHE[key] = (pupil_full, mode_full, w0)
        pupil_full, mode_full, self.W0 = hit
        self.pupil_sat = funcs.compute_pupil(32, self.dx_sat, self.D_sat, self.obsc_sat)
        self.pupil_mode_sat, self.W0_sat = funcs.compute_gaussian_mode(self.pupil_sat, self.dx_sat,
                                                                       "opt", ptype="gauss")
        self._pm_full = pupil_full * mode_full            # input of the device pupil filter
        lo, hi = (N - self.Npxls_pup) // 2, (N + self.Npxls_pup) // 2
        self._lo = lo
        self.pup_coords = numpy.array((numpy.arange(lo, hi), numpy.arange(lo, hi))).astype(int)
        self.pupil = pupil_full[lo:hi, lo:hi].copy()          # own, writable crops (the full arrays are shared)
        self.pupil_mode = mode_full[lo:hi, lo:hi].copy()
        if self.temporal:
            ft = self.freq.temporal
            self.pupil_filter_temporal = temporal.elongated_pupil_filter(self, ft.fx_axis, ft.fy_axis)
        return self.pupil

    def init_fftw(self):
        """fast/fast.py:419-438 plans pyFFTW transforms; the transforms of this path are the K2 / K4 kernels, so
        there is nothing to plan (FFTW / FFTW_THREADS are accepted and ignored)."""
        self.fftw_objs = {}

    def init_phs_logamp(self):
        # screens are never materialised on this path; `logamp` holds the host draws in
        # RNG='numpy' mode (fast/fast.py:440-443)
        self.phs = None
        self.logamp = numpy.zeros((self.Niter))

    def compute_link_budget(self):
        """Analytic link budget [dB] and the diffraction-limited power (fast/fast.py:670-734)."""
        up = self.params['PROP_DIR'] == "up"
        if up:
            D_t, D_r, obsc_t, obsc_r = self.D_ground, self.D_sat, self.obsc_ground, self.obsc_sat
            mode, dx_r, pupil_r, w0 = self.pupil_mode_sat, self.dx_sat, self.pupil_sat, self.W0
        else:
            D_t, D_r, obsc_t, obsc_r = self.D_sat, self.D_ground, self.obsc_sat, self.obsc_ground
            mode, dx_r, pupil_r, w0 = self.pupil_mode, self.dx, self.pupil, self.W0_sat
        lb = {}
        lb['power'] = 10 * numpy.log10(self.power / 1e-3)
        lb['free_space'] = 10 * numpy.log10((self.wvl / (4 * numpy.pi * self.L)) ** 2)
        # truncated-Gaussian transmitter gain, Klein & Degnan, Appl. Opt. 13 (1974) eq. 9
        alpha, gamma = D_t / (2 * w0), obsc_t / D_t
        g_t = 2 / alpha ** 2 * (numpy.exp(-alpha ** 2) - numpy.exp(-gamma ** 2 * alpha ** 2)) ** 2
        lb['transmitter_gain'] = 10 * numpy.log10((numpy.pi * D_t ** 2) * 4 * numpy.pi / self.wvl ** 2 * g_t)
        area = numpy.pi * ((D_r / 2) ** 2 - (obsc_r / 2) ** 2)
        lb['receiver_gain'] = 10 * numpy.log10(4 * numpy.pi * area / self.wvl ** 2)
        lb['transmission_loss'] = 10 * numpy.log10(self.params['TRANSMISSION'])
        lb['smf_coupling'] = 10 * numpy.log10(((pupil_r * mode).sum() * dx_r) ** 2 / (mode ** 2).sum())
        self.link_budget = lb
        self.diffraction_limit = 10 ** (sum(lb.values()) / 10) / 1e3      # W
        return self.link_budget

    # ------------------------------------------------------------------ K1 on the device
    def _psd_params(self, n=None, df=None):
        pp = _lib.PsdParams()
        L = len(self.h)
        if L > _lib.MAX_LAYERS:
            raise Exception(f'at most {_lib.MAX_LAYERS} turbulence layers are supported')
        pp.n, pp.n_layers = (self.Npxls if n is None else n), L
        pp.ao_mode = _AO_MODES[self.ao_mode]
        pp.alias = 1 if self.alias else 0
        pp.lmax = pp.kmax = 5
        pp.df = float(self.freq.main.df if df is None else df)
        pp.k, pp.wvl = float(self.k), float(self.wvl)
        pp.L0, pp.l0 = float(self.L0), float(self.l0)
        pp.dsubap, pp.tloop, pp.texp = float(self.Dsubap), float(self.tloop), float(self.texp)
        pp.noise_var = float(self.noise)
        pp.dtheta[0], pp.dtheta[1] = float(self.dtheta[0]), float(self.dtheta[1])
        for i in range(L):
            pp.h[i], pp.cn2[i] = float(self.h[i]), float(self.cn2[i])
            pp.vx[i], pp.vy[i] = float(self.wind_vector[i, 0]), float(self.wind_vector[i, 1])
        return pp

    @_on_device
    def compute_powerspec(self):
        """Residual phase PSD, log-amplitude PSD and the error-budget integrals
        (fast/fast.py:445-492) -- one fused kernel + one batched Simpson reduction."""
        N, L, dev = self.Npxls, len(self.h), self.device
        f64 = torch.float64
        d = self._d
        d['pupil_filter'] = _lib.pupil_filter(torch.from_numpy(numpy.ascontiguousarray(self._pm_full)).to(dev))
        lf = zf = None
        if self.modal:
            lf = self._modal_mask_device(N, self.freq.main.df)
        if self.ao_mode == 'LGSAO':
            zf = self._lgs_filter_device(N, self.freq.main.df)
        # one slab: [aniso_servo, alias, fitting | noise | W | logamp | per-layer (L)]
        slab = torch.zeros((6 + L, N, N), dtype=f64, device=dev)
        d['turb'] = torch.empty((L, N, N), dtype=f64, device=dev)
        d['g_ao'] = torch.empty((L, N, N), dtype=f64, device=dev)
        d['alias'] = torch.empty((L, N, N), dtype=f64, device=dev)
        d['weight'] = torch.empty((N, N), dtype=torch.float32, device=dev)
        outs = {'integrands': slab[0:3], 'noise': slab[3], 'powerspec': slab[4], 'logamp': slab[5],
                'powerspec_per_layer': slab[6:], 'turb': d['turb'], 'g_ao': d['g_ao'],
                'alias': d['alias'], 'weight': d['weight']}
        if self.temporal:
            d['weight_per_layer'] = torch.empty((L, N, N), dtype=torch.float32, device=dev)
            outs['weight_per_layer'] = d['weight_per_layer']
        _lib.psd_build(self._psd_params(), outs, lf_mask=lf, zfilter=zf, pupil_filter=d['pupil_filter'])
        self._prep_key = None
        d['noise'], d['powerspec'], d['logamp'], d['powerspec_per_layer'] = slab[3], slab[4], slab[5], slab[6:]
        w = torch.from_numpy(funcs.simpson_weights(self.freq.main.f)).to(dev)
        ints = _lib.simpson2d(slab, w).cpu().numpy()
        noao = self.ao_mode == 'NOAO'
        self.aniso_servo_error = float(ints[0])
        self.alias_error = float(ints[1]) if (self.alias and not noao) else 0.
        self.fitting_error = float(ints[2])
        self.noise_error = float(ints[3]) if (self.noise > 0 and not noao) else 0.
        self.phs_var = float(ints[4])
        self.logamp_var = float(ints[5])
        self.phs_var_weights = ints[6:] / self.phs_var
        self.powerspec_subharm = self.phs_var_subharm = self.phs_var_weights_sh = None
        if self.subharmonics:
            self._compute_powerspec_subharm()
        self.temporal_powerspec = self.temporal_logamp_powerspec = None
        self.shifts = self.shifts_sh = None
        U = numpy.ascontiguousarray(self.pupil * self.pupil_mode)
        self._u_sum = float(U.sum())
        u32 = U.astype(numpy.float32)
        self._u_digest = hash(u32.tobytes())       # sweep.run_sweep groups the samples that share U
        d['U'] = torch.from_numpy(u32).to(dev)
        self._u_version = d['U']._version
        if self.temporal:
            # per-step wind shifts in pixels (fast/fast.py:543-544) and the temporal log-amp PSD
            dts = numpy.arange(1, self.Niter_per_chunk + 1) * self.dt
            self.pixel_shifts = dts * self.wind_vector[..., numpy.newaxis] / self.dx
            d['pixel_shifts'] = torch.from_numpy(numpy.ascontiguousarray(self.pixel_shifts, dtype=numpy.float64)).to(dev)
            ft = self.freq.temporal
            self.temporal_logamp_powerspec = temporal.temporal_logamp_powerspec(
                self, ft.fx_axis, ft.fy_axis, ft.fabs, self.pupil_filter_temporal)

    def _compute_powerspec_subharm(self):
        """Residual PSD on the three 3 x 3 sub-harmonic levels (fast/fast.py:494-531): the same
        K1 kernel with N = 3 and the level's spacing, plus the plane-wave tables K2 needs."""
        dev, L, f64 = self.device, len(self.h), torch.float64
        sub = self.freq.subharm
        W = torch.zeros((3, 3, 3), dtype=f64, device=dev)
        per_layer = torch.zeros((3, L, 3, 3), dtype=f64, device=dev)
        for i in range(3):
            lf = self._modal_mask_device(3, sub.df[i]) if self.modal else None
            zf = self._lgs_filter_device(3, sub.df[i]) if self.ao_mode == 'LGSAO' else None
            _lib.psd_build(self._psd_params(n=3, df=sub.df[i]),
                           {'powerspec': W[i], 'powerspec_per_layer': per_layer[i]}, lf_mask=lf, zfilter=zf)
        self.powerspec_subharm = W.cpu().numpy()
        self.powerspec_subharm_per_layer = per_layer.cpu().numpy().transpose(1, 0, 2, 3).copy()
        self.phs_var_subharm = self.powerspec_subharm_per_layer.sum((-1, -2)) * sub.df ** 2
        self.phs_var_weights_sh = self.phs_var_subharm / self.phs_var_subharm.sum()
        # plane-wave tables for K2 (fast/funcs.py:227-253): pixel coordinates, per-level phasors
        # on the pupil window, and the full-grid mean of each of the 27 waves
        N, lo, P = self.Npxls, self._lo, self.Npxls_pup
        D = self.dx * N
        coords = numpy.arange(-D / 2, D / 2, self.dx)[:N]
        phasor = numpy.exp(1j * coords[None, :] * sub.df[:, None])                  # (3, N): f = +df_i
        grid_mean = numpy.stack([phasor.conj().mean(1), numpy.ones(3), phasor.mean(1)], axis=1)   # [i][s]
        mean27 = grid_mean[:, None, :] * grid_mean[:, :, None]                       # [i][q][s]
        weight27 = numpy.sqrt(self.powerspec_subharm) * sub.df[:, None, None]

        def c64(a):
            return torch.view_as_real(torch.from_numpy(numpy.ascontiguousarray(a, dtype=numpy.complex64))).to(dev)
        win = phasor[:, lo:lo + P]
        self._d['subharm'] = {'weight': torch.from_numpy(weight27.astype(numpy.float32).reshape(27)).to(dev),
                              'ex': c64(win), 'ey': c64(win), 'mean': c64(mean27.reshape(27))}

    def _host(self, name):
        if name not in self._host_cache:
            self._host_cache[name] = self._d[name].cpu().numpy()
        return self._host_cache[name]

    powerspec = property(lambda self: self._host('powerspec'))
    powerspec_per_layer = property(lambda self: self._host('powerspec_per_layer'))
    logamp_powerspec = property(lambda self: self._host('logamp'))
    turb_powerspec = property(lambda self: self._host('turb'))
    pupil_filter = property(lambda self: self._host('pupil_filter'))

    @property
    def G_ao(self):
        return 1 if self.ao_mode == 'NOAO' else self._host('g_ao')

    @property
    def alias_powerspec(self):
        return self._host('alias') if (self.alias and self.ao_mode != 'NOAO') else 0.

    @property
    def noise_powerspec(self):
        return self._host('noise') if (self.noise > 0 and self.ao_mode != 'NOAO') else 0.

    # ------------------------------------------------------------------ K2 on the device
    def _run_params(self, n_pairs, first_pair, algo=_lib.ALGO_AUTO):
        rp = _lib.RunParams()
        rp.n, rp.n_pup, rp.lo = self.Npxls, self.Npxls_pup, self._lo
        rp.coherent = 1 if self.params['COHERENT'] else 0
        rp.algo = algo
        rp.flags = _lib.RUN_RNG_FAST if self.rng_mode == 'device-fast' else 0
        rp.n_pairs, rp.first_pair = int(n_pairs), int(first_pair)
        rp.pairs_per_chunk = self.Niter_per_chunk // 2
        rp.seed = self._run_seed()
        rp.u_sum = self._u_sum
        rp.sigma_chi = math.sqrt(self.logamp_var)
        return rp

    def _run_seed(self):
        """Philox key of the current (or last) run.  Run k > 0 of one object uses the key
        seed + k * 0x9E3779B97F4A7C15 (mod 2^64), so that successive run() calls draw fresh,
        independent realisations like the reference's persistent generator does (fast/funcs.py:21),
        while a new object with the same SEED reproduces the first run."""
        base = int(self.seed) if self.seed != None else self._auto_seed()  # noqa: E711
        return (base + self._run_index * _GOLDEN64) & 0xFFFFFFFFFFFFFFFF

    def _auto_seed(self):
        if not hasattr(self, '_seed_drawn'):
            drawn = int(numpy.random.SeedSequence().generate_state(2, numpy.uint32).view(numpy.uint64)[0])
            # unseeded runs under torch.distributed: every rank uses rank 0's draw, so that the
            # sharded result does not depend on the number of ranks
            self._seed_drawn = dist.broadcast_seed(drawn, self.device)
        return self._seed_drawn

    def _workspace(self, rp, n_items=1):
        nbytes = _lib.screen_detect_workspace_bytes(rp, n_items)
        ws = self._d.get('workspace')
        if ws is None or ws.numel() < nbytes:
            ws = self._d['workspace'] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._prep_key = None
        return ws

    def _prepared_workspace(self, rp):
        """Workspace whose derived tables (transposed U, pre-scaled weight copy / chirp tables) match
        the current weight and U: prepared once, re-prepared when either tensor was written to."""
        ws = self._workspace(rp)
        w, U = self._d['weight'], self._d['U']
        key = (ws.data_ptr(), w.data_ptr(), w._version, U.data_ptr(), U._version, int(rp.algo))
        if self._prep_key != key:
            _lib.screen_detect_prepare(rp, w, U, ws)
            self._prep_key = key
        return ws

    @_on_device
    def screen_detect(self, first_pair, n_pairs, noise=None, chi=None, algo=_lib.ALGO_AUTO, noise_lo=None,
                      stats=None):
        """Run K2 for global pairs [first_pair, first_pair + n_pairs).  Returns two device
        tensors (results of the Re and Im realisations); complex64 when COHERENT.
        noise: optional (n_pairs, N, N) complex64 device tensor; chi: optional float32 device
        tensor indexed by global realisation index; stats: optional dist.StatsBuffers that the
        kernel accumulates the moments / extrema / dB histogram of these results into."""
        rp = self._run_params(n_pairs, first_pair, algo)
        width = 2 if rp.coherent else 1
        out_a = torch.empty(n_pairs * width, dtype=torch.float32, device=self.device)
        out_b = torch.empty(n_pairs * width, dtype=torch.float32, device=self.device)
        if n_pairs:
            fused = noise is None and not self.subharmonics and algo != _lib.ALGO_RADIX_PAIR
            if fused:
                # device RNG: one launch -- tables prepared once per (weight, U), statistics fused
                ws = self._prepared_workspace(rp)
                rp.flags |= _lib.RUN_PREPARED
                st = None if stats is None else _lib.run_stats(stats.db_lo, stats.db_hi, stats.nbins,
                                                               stats.sums, stats.minmax, stats.hist)
                _lib.screen_detect_batch(rp, self._d['weight'], self._d['U'], out_a, out_b, ws, stats=st, chi=chi)
            else:
                nz = None if noise is None else torch.view_as_real(noise.contiguous())
                sh = None
                if self.subharmonics:
                    sh = dict(self._d['subharm'])
                    if noise_lo is not None:
                        sh['noise'] = torch.view_as_real(noise_lo.contiguous())
                self._prep_key = None
                _lib.screen_detect(rp, self._d['weight'], self._d['U'], out_a, out_b, self._workspace(rp),
                                   chi=chi, noise=nz, subharm=sh)
        if rp.coherent:
            out_a = torch.view_as_complex(out_a.view(-1, 2))
            out_b = torch.view_as_complex(out_b.view(-1, 2))
        if stats is not None and n_pairs and not fused:
            r = torch.cat([out_a, out_b])
            r = (r.real ** 2 + r.imag ** 2) if r.is_complex() else r
            _lib.stats(r.contiguous(), stats.db_lo, stats.db_hi, stats.nbins, stats.sums[0], stats.minmax[0],
                       stats.hist[0])
        return out_a, out_b

    @_on_device
    def screens(self, first_pair, n_pairs, noise=None, noise_lo=None):
        """The cropped phase screens of global pairs [first_pair, first_pair + n_pairs) as a
        (2 n_pairs, Npup, Npup) float32 device tensor ordered [Re_0, Im_0, Re_1, Im_1, ...]
        (inspection seam, fastb_screens_crop; the run itself never materialises screens)."""
        rp = self._run_params(n_pairs, first_pair, _lib.ALGO_DIRECT)
        P = self.Npxls_pup
        phs = torch.empty((2 * n_pairs, P, P), dtype=torch.float32, device=self.device)
        if n_pairs:
            nz = None if noise is None else torch.view_as_real(noise.contiguous())
            sh = None
            if self.subharmonics:
                sh = dict(self._d['subharm'])
                if noise_lo is not None:
                    sh['noise'] = torch.view_as_real(noise_lo.contiguous())
            self._prep_key = None                  # the direct kernel's scratch overlaps the tables
            _lib.screens_crop(rp, self._d['weight'], phs, self._workspace(rp), noise=nz, subharm=sh)
        return phs

    def compute_logamp(self):
        """RNG='numpy': host draws in the reference's order (fast/fast.py:639-645);
        RNG='device': chi is generated inside the kernel; `logamp` keeps the zeros it was created with (no pass over
        NITER host values per run: 1 ms at 1.2e6 realisations, 7 ms for the 9.6e6 of an 8-GPU step)."""
        if self.temporal:
            # temporally coloured chi: Niter complex draws + one 1-D FFT on the host
            # (fast/funcs.py:367-375).  RNG='device' uses a generator derived from the seed.
            keep = funcs._R
            if self.rng_mode != 'numpy':
                funcs._R = numpy.random.default_rng([self._run_seed(), 0xC41])
            try:
                self.logamp[:] = funcs.generate_random_coefficients_logamp(
                    self.Niter, self.logamp_var, True, self.temporal_logamp_powerspec).real
            finally:
                funcs._R = keep
            self._d['chi'] = torch.from_numpy(self.logamp.astype(numpy.float32)).to(self.device)
        elif self.rng_mode == 'numpy':
            self.logamp[:] = funcs.generate_random_coefficients_logamp(self.Niter, self.logamp_var).real
            self._d['chi'] = torch.from_numpy(self.logamp.astype(numpy.float32)).to(self.device)
        else:
            self._d.pop('chi', None)
        return self.logamp

    def compute_phs(self, chunk=0):
        """RNG='numpy': draw this chunk's complex noise on the host exactly like the reference
        (fast/fast.py:593) and stage it on the device.  Screens are not materialised."""
        if self.rng_mode == 'numpy':
            J2, N = self.Niter_per_chunk // 2, self.Npxls
            rand = funcs.generate_random_coefficients((J2, N, N)).astype(numpy.complex64)
            self._d['noise'] = torch.from_numpy(rand).to(self.device)
            if self.subharmonics:       # drawn after the main block, like fast/fast.py:600
                lo = funcs.generate_random_coefficients((J2, 3, 3, 3)).astype(numpy.complex64)
                self._d['noise_lo'] = torch.from_numpy(lo.reshape(J2, 27)).to(self.device)
            if self.params.get('KEEP_PHS', False):
                # opt-in: also materialise this chunk's screens in the reference's layout
                # (J, Npup, Npup) = [Re screens | Im screens] (fast/funcs.py:220-221, fast/fast.py:596)
                s = self.screens(chunk * J2, J2, noise=self._d['noise'], noise_lo=self._d.get('noise_lo'))
                self.phs = torch.cat([s[0::2], s[1::2]]).cpu().numpy().astype(float)
        return self.phs

    def _temporal_layer_screens(self):
        """The chunk-0 branch of fast/fast.py:609-616: one real screen per layer on the device
        (fastb_layer_screens) and the sample coordinates of the first step."""
        noise = None
        if self.rng_mode == 'numpy':
            shape = (len(self.h), self.Npxls, self.Npxls)
            noise = torch.from_numpy(funcs.generate_random_coefficients(shape).astype(numpy.complex64)).to(self.device)
        self._d['layer_screens'] = _lib.layer_screens(self._d['weight_per_layer'], self._run_seed(), noise=noise)
        self.interp_coords = self.pup_coords[numpy.newaxis, :, numpy.newaxis, :].astype(float) \
            + self.pixel_shifts[:, :, :, numpy.newaxis]

    def compute_phs_temporal(self, chunk=0):
        """TEMPORAL mode (fast/fast.py:607-637), one chunk at a time as the reference's loop does.  Chunk 0: the
        layer screens.  Every chunk: the wind-shifted sample coordinates of its J steps (host numpy, the literal
        restatement in fast_b200/temporal.py) staged on the device; the gather itself is fused with the detector.
        Fast.run() does all chunks at once (_run_temporal)."""
        if chunk == 0:
            self._temporal_layer_screens()
        coords = self._advance_temporal_coords()
        self._d['tcoords'] = tuple(torch.from_numpy(c).to(self.device) for c in coords)
        return None

    def _run_temporal(self):
        """Every chunk of a TEMPORAL run in TWO launches after the layer screens: fastb_temporal_coords does the
        coordinate bookkeeping of all Nchunks x J steps on the device (the reference's arithmetic operation for
        operation in float64: chunk-after-chunk accumulation, wrap, sort, roll, clamp -- fast/fast.py:617-635) and
        fastb_temporal_detect gathers and detects them, instead of Nchunks rounds of host bookkeeping + copies +
        launches (the run is latency-bound at the reference's sizes).  Leaves the object in the state the chunk
        loop would."""
        self._temporal_layer_screens()
        coords = _lib.temporal_coords(self.Npxls, self.Npxls_pup, self._lo, self._d['pixel_shifts'], self.Nchunks)
        step = self.pixel_shifts[:, :, -1, numpy.newaxis, numpy.newaxis]
        for _ in range(self.Nchunks):                       # fast/fast.py:635, once per chunk
            self.interp_coords = self.interp_coords + step
        J = self.Niter_per_chunk
        self._d['tcoords'] = tuple(c[:, -J:].contiguous() for c in coords)
        return self._temporal_detector(0, steps=self.Niter, coords=coords)

    def _advance_temporal_coords(self):
        """Sample coordinates of the next chunk from self.interp_coords, which is then moved on by the
        chunk's total wind shift (fast/fast.py:621-635)."""
        coords = temporal.sample_coordinates(self.interp_coords, self.Npxls)
        self.interp_coords = self.interp_coords + self.pixel_shifts[:, :, -1, numpy.newaxis, numpy.newaxis]
        return coords

    def _temporal_detector(self, chunk, steps=None, coords=None):
        J = self.Niter_per_chunk if steps is None else steps
        tp = _lib.TemporalParams()
        tp.n, tp.n_pup, tp.n_layers = self.Npxls, self.Npxls_pup, len(self.h)
        tp.coherent = 1 if self.params['COHERENT'] else 0
        tp.n_steps, tp.u_sum = J, self._u_sum
        out = torch.empty(J * (2 if tp.coherent else 1), dtype=torch.float32, device=self.device)
        xi, xf, yi, yf = self._d['tcoords'] if coords is None else coords
        chi = self._d['chi'][chunk * J:(chunk + 1) * J].contiguous()
        _lib.temporal_detect(tp, self._d['layer_screens'], xi, xf, yi, yf, self._d['U'], chi, out)
        return torch.view_as_complex(out.view(-1, 2)) if tp.coherent else out

    def compute_detector(self, chunk=0):
        """Screens + detector for one chunk on the device; returns the chunk's J results in
        the reference's order [Re half | Im half] (fast/fast.py:647-668) as a device tensor."""
        if self.temporal:
            self.random_iters = self._temporal_detector(chunk)
            return self.random_iters
        ppc = self.Niter_per_chunk // 2
        numpy_rng = self.rng_mode == 'numpy'
        a, b = self.screen_detect(chunk * ppc, ppc, noise=self._d.get('noise') if numpy_rng else None,
                                  chi=self._d.get('chi'),
                                  noise_lo=self._d.get('noise_lo') if numpy_rng else None)
        self.random_iters = torch.cat([a, b])
        return self.random_iters

    @_on_device
    def run(self):
        """Monte-Carlo run (fast/fast.py:115-140).  With torch.distributed initialised and the
        device RNG, pair ranges are sharded over the ranks and gathered (fast_b200/dist.py).
        Successive calls give fresh realisations (see _run_seed)."""
        self._run_index = self._runs
        self._runs += 1
        self._d.pop('run_stats', None)
        logger.debug("Compute log amplitude values")
        self.compute_logamp()
        ppc = self.Niter_per_chunk // 2
        total = self.Nchunks * ppc
        if self.temporal:
            flat = self._run_temporal()
        elif self.rng_mode != 'numpy':
            rank, world = dist.rank_world()
            lo, hi = dist.shard_range(total, rank, world)
            # moments / extrema / dB histogram of the run accumulated in the kernel epilogue and combined over the
            # ranks with one small collective: result_stats() reads them without another pass over the results
            sb = self._d.get('stats_buffers')
            if sb is None:
                sb = self._d['stats_buffers'] = dist.StatsBuffers(4096, self.device)
            sb.reset()
            a, b = self.screen_detect(lo, hi - lo, stats=sb)
            sb.allreduce()
            self._d['run_stats'] = sb
            a, b = dist.gather_pairs(a, b, total, world)
            flat = dist.assemble(a, b, self.Nchunks, ppc)
        else:
            parts = []
            for i in range(self.Nchunks):
                logger.debug(f"Compute phase for chunk {i+1}")
                self.compute_phs(chunk=i)
                logger.debug(f"Compute detector for chunk {i+1}")
                parts.append(self.compute_detector(chunk=i))
            flat = torch.cat(parts)
            self._d.pop('noise', None)
        self._d['result'] = flat
        self._publish(flat)
        logger.info(self.result)
        return self.result

    def _to_host(self, flat):
        """Device results -> the reference's host array (float64, complex128 when COHERENT).  The widening
        happens on the device and the copy lands in pinned memory from torch's caching host allocator; the
        returned numpy array is a view of that block (no pass over the data on the host)."""
        wide = torch.complex128 if flat.is_complex() else torch.float64
        if flat.dtype != wide:
            flat = flat.to(wide)
        if flat.numel() < 65536:                   # short runs (e.g. TEMPORAL chunks): a pinned block is not worth it
            return flat.cpu().numpy()
        host = torch.empty(flat.shape, dtype=wide, pin_memory=True)
        host.copy_(flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()

    def _publish(self, flat):
        """self.result / self.I from the device results (fast/fast.py:136-137).  `I = result.power` is the
        IEEE product diffraction_limit * r: formed on the device in float64 (bit-identical to numpy's) so that
        the host makes no pass over NITER values."""
        wide = flat.to(torch.complex128 if flat.is_complex() else torch.float64)
        self.result = FastResult(self._to_host(wide), self.diffraction_limit)
        self.I = self._to_host(wide * self.diffraction_limit)
        return self.result

    @_on_device
    def result_stats(self, db_lo=-60.0, db_hi=3.0, nbins=4096):
        """Moments / extrema / dB histogram of the last run computed on the device
        (fastb_stats) and, under torch.distributed, all-reduced over the ranks' shards."""
        sb = self._d.get('run_stats')
        if sb is not None and (sb.db_lo, sb.db_hi, sb.nbins) == (float(db_lo), float(db_hi), int(nbins)):
            return sb.summary()                    # fused into the run's kernel epilogue, already global
        r = self._d['result']
        if r.is_complex():
            r = (r.real ** 2 + r.imag ** 2).contiguous()
        return dist.reduced_stats(r, db_lo, db_hi, nbins, already_global=True)

    def compute_mean_irradiance(self, onaxis=True):
        """Analytic (no Monte Carlo) mean coupled power from the residual PSD: long-exposure OTF
        = exp(-D_phi) x pupil OTF (fast/fast.py:736-761).  Host numpy on the device-built PSD;
        one-off N x N transforms, not on the hot path."""
        def ift2(a, df):
            ax = (-1, -2)
            return numpy.fft.ifftshift(numpy.fft.ifft2(numpy.fft.ifftshift(a, axes=ax)), axes=ax) * (a.shape[-1] * df) ** 2

        def ft2(a, dx):
            ax = (-1, -2)
            return numpy.fft.fftshift(numpy.fft.fft2(numpy.fft.fftshift(a, axes=ax)), axes=ax) * dx ** 2

        W = self.powerspec
        pupil = numpy.zeros(W.shape)
        pupil[:self.pupil.shape[0], :self.pupil.shape[1]] = self.pupil * self.pupil_mode
        cov = ift2(W, self.freq.df)
        mid = (cov.shape[0] // 2, cov.shape[1] // 2)
        structure = cov[mid] - cov
        pupil_otf = ift2(numpy.abs(ft2(pupil, self.dx)) ** 2, self.freq.df) / (2 * numpy.pi) ** 2
        otf = numpy.exp(-structure) * pupil_otf
        psf = ft2(otf, self.dx).real if not onaxis else otf.sum().real * self.dx ** 2
        return psf * self.diffraction_limit / (pupil.sum() * self.dx ** 2) ** 2

    # ------------------------------------------------------------------ FITS (optional dep)
    def make_header(self, params):
        from astropy.io import fits
        hdr = fits.Header()
        for key, val in (('ZENITH', params['ZENITH_ANGLE']), ('WVL', int(params['WVL'] * 1e9)),
                         ('OTRSCALE', str(params['L0']) if numpy.isinf(params['L0']) else params['L0']),
                         ('INRSCALE', params['l0']), ('POWER', params['POWER']), ('PAA', self.paa),
                         ('AO_MODE', self.ao_mode), ('TLOOP', params['TLOOP']), ('TEXP', params['TEXP']),
                         ('DSUBAP', params['DSUBAP']), ('ALIAS', str(params['ALIAS'])),
                         ('NOISE', params['NOISE']), ('D_GND', params['D_GROUND']),
                         ('OBSC_GND', params['OBSC_GROUND']), ('D_SAT', params['D_SAT']),
                         ('OBSC_SAT', params['OBSC_SAT']), ('AXICON', str(params['AXICON'])),
                         ('W0', self.W0), ('L_SAT', self.L), ('H_SAT', params['H_SAT']), ('DX', self.dx),
                         ('NPXLS', self.Npxls), ('NITER', self.Niter), ('R0', self.r0),
                         ('THETA0', self.theta0), ('TAU0', self.tau0), ('DIFFLIM', self.diffraction_limit)):
            hdr[key] = val
        if self.seed != None:  # noqa: E711
            hdr['SEED'] = self.seed
        return hdr

    def save(self, fname, **kwargs):
        from astropy.io import fits
        fits.writeto(fname, self.result.power, header=self.make_header(self.params), **kwargs)


class FastResult():
    """Unit conversions over the per-realisation array (fast/fast.py:931-994): `_r` is the
    received power relative to the diffraction limit (complex field when COHERENT)."""

    def __init__(self, random_iters, diffraction_limit, header=None):
        self._r = random_iters
        self._dl = diffraction_limit
        if header != None:  # noqa: E711
            self.hdr = header

    dB_rel = property(lambda self: 10 * numpy.log10(self._r))
    dB_abs = property(lambda self: 10 * numpy.log10(self._r * self._dl))
    dBm = property(lambda self: 10 * numpy.log10(self._r * self._dl / 1e-3))
    power = property(lambda self: self._dl * self._r)
    scintillation_index = property(lambda self: (self._r / self._r.mean()).var())
    avg_power_W = property(lambda self: self.power.mean())
    avg_power_dBm = property(lambda self: 10 * numpy.log10(self.avg_power_W / 1e-3))
    avg_power_dB_rel = property(lambda self: 10 * numpy.log10((self.power / self._dl).mean()))
    avg_power_dB_abs = property(lambda self: 10 * numpy.log10(self.avg_power_W))

    def __str__(self):
        return ("FAST result statistics:\n"
                f"            Avg. power (W): {self.avg_power_W}\n"
                f"            Avg. power (dBm): {self.avg_power_dBm}\n"
                f"            Avg. power (dB_rel): {self.avg_power_dB_rel}\n"
                f"            Avg. power (dB_abs): {self.avg_power_dB_abs}\n"
                f"            Scintillation index: {self.scintillation_index}\n        ")


def load(fname):
    """Read a result saved by Fast.save (fast/fast.py:998-1002)."""
    from astropy.io import fits
    hdr = fits.getheader(fname)
    data = fits.getdata(fname)
    data /= hdr['DIFFLIM']
    return FastResult(data, hdr['DIFFLIM'], header=hdr)
