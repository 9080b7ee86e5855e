"""Named synthetic configurations of the hot path for bench.py and
examples; product-side twin of oracle/configs.py (tests check that the two agree).

All use the 4-layer HV57/Bufton profile of the reference's example config
(test/test_params.py:8-9,13-66)."""
import numpy as np

from . import funcs, turbulence_models

_H, _CN2, _W = turbulence_models.HV57_Bufton_profile(4)


def base():
    """test/test_params.py:13-66 verbatim (TEMPORAL uplink example), as a fresh dict."""
    return {
        'NPXLS': 'auto', 'DX': 0.01, 'NITER': 100, 'SUBHARM': False, 'FFTW': False,
        'FFTW_THREADS': 1, 'NCHUNKS': 10, 'TEMPORAL': True, 'DT': 0.001, 'LOGFILE': None,
        'LOGLEVEL': 'ERROR', 'SEED': None,
        'WVL': 1550e-9, 'POWER': 1, 'W0': 'opt', 'D_GROUND': 0.8, 'OBSC_GROUND': 0,
        'D_SAT': 0.1, 'OBSC_SAT': 0, 'AXICON': False, 'SMF': True,
        'H_SAT': 36e6, 'L_SAT': None, 'H_TURB': _H.copy(), 'CN2_TURB': _CN2.copy(),
        'WIND_SPD': _W.copy(), 'WIND_DIR': [0, 90, 180, 270], 'L0': np.inf, 'l0': 1e-6,
        'ZENITH_ANGLE': 55, 'PROP_DIR': 'up', 'DTHETA': [4, 0], 'TRANSMISSION': 1,
        'AO_MODE': 'AO', 'DSUBAP': 0.1, 'TLOOP': 0.001, 'TEXP': 0.001, 'ALIAS': True,
        'NOISE': 0, 'MODAL': False, 'MODAL_MULT': 1, 'ZMAX': None,
        'COHERENT': False, 'MODULATION': None, 'EsN0': None,
    }


def _mk(**kw):
    p = base()
    p.update(kw)
    return p


def c1(seed=1):
    """Config 1 verbatim: test/test_params.py (TEMPORAL uplink, N auto -> 164, NITER=100)."""
    return _mk(SEED=seed)


def c1prime(niter=100, nchunks=10, seed=1, **kw):
    """Config 1 with TEMPORAL off (test/tests_pytest.py:56-59): N auto -> 164."""
    d = dict(TEMPORAL=False, NITER=niter, NCHUNKS=nchunks, SEED=seed)
    d.update(kw)
    return _mk(**d)


def c2(niter=100000, nchunks=1, seed=1):
    """GEO downlink, 256x256, SMF detection (BASELINE.json configs[1])."""
    return _mk(NPXLS=256, DX=0.01, PROP_DIR='down', TEMPORAL=False, NITER=niter,
               NCHUNKS=nchunks, SEED=seed)


def c3_elevation(el_deg, niter=10000, nchunks=1, seed=1):
    """One sample of the synthetic LEO pass: the keys FAST_sat_orbit sets
    (fast/complete_orbit_simulation.py:218-225) for a 550 km orbit at elevation el_deg."""
    zen = 90.0 - el_deg
    L = funcs.l_path(550e3, zen)
    return _mk(NPXLS=256, DX=0.01, PROP_DIR='down', TEMPORAL=False, NITER=niter, NCHUNKS=nchunks,
               SEED=seed, ZENITH_ANGLE=zen, L_SAT=L,
               DTHETA=[10.5 * np.sin(np.radians(el_deg)), 0.0],
               ANISO_DL=[2.85 * 550e3 / L, 0.0], AZIMUT_SAT=0.0)


C3_ELEVATIONS = [10.0 + 5.0 * i for i in range(16)]


def c4(niter=100000, nchunks=1, seed=1):
    """Coherent detection, 512x512 (BASELINE.json configs[3])."""
    return _mk(NPXLS=512, DX=0.005, PROP_DIR='down', TEMPORAL=False, COHERENT=True,
               NITER=niter, NCHUNKS=nchunks, SEED=seed)


def c5(niter=1000000, nchunks=1, seed=1):
    """Large sweep, 1024x1024 (BASELINE.json configs[4])."""
    return _mk(NPXLS=1024, DX=0.005, PROP_DIR='down', TEMPORAL=False, NITER=niter,
               NCHUNKS=nchunks, SEED=seed)


def mini(niter=40, nchunks=2, seed=3, **kw):
    """64x64 miniature of C2 (DX=0.04 -> Npup=22) for exhaustive term-by-term checks."""
    d = dict(NPXLS=64, DX=0.04, PROP_DIR='down', TEMPORAL=False, NITER=niter, NCHUNKS=nchunks,
             SEED=seed)
    d.update(kw)
    return _mk(**d)
