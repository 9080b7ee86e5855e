"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) through
oracle/shim in the build container.  TEST INFRASTRUCTURE.

    python oracle/make_golden.py [case ...]

/root/reference does not exist on the GPU box, so the outputs are committed; nothing at test
time reads the reference.  Each file holds the reference's init scalars / PSD arrays and the
per-realisation result `_r` of `Fast(p).run()` for one named case of oracle/configs.py.  The
noise itself is not stored: it is `numpy.random.default_rng(SEED)` drawn in the reference's
order (fast/funcs.py:352-365), and a few leading values are stored as a stream guard.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'shim'))

import numpy as np  # noqa: E402

import fast  # noqa: E402  (the reference)
from oracle import configs  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

# name -> (config factory name, kwargs, storage level)
CASES = {
    'mini_ao':       ('mini', {}, 'full'),
    'mini_noise_L0': ('mini', {'NOISE': 1.0, 'L0': 25.0}, 'full'),
    'mini_noao':     ('mini', {'AO_MODE': 'NOAO'}, 'full'),
    'mini_tt':       ('mini', {'AO_MODE': 'TT'}, 'full'),
    'mini_modal':    ('mini', {'MODAL': True}, 'full'),
    'mini_lgsao':    ('mini', {'AO_MODE': 'LGSAO'}, 'full'),
    'mini_axicon':   ('mini', {'W0': 0.2, 'AXICON': True, 'OBSC_GROUND': 0.2}, 'full'),
    'mini_coherent': ('mini', {'COHERENT': True}, 'full'),
    'mini_up_w0':    ('mini', {'PROP_DIR': 'up', 'W0': 0.3, 'ZENITH_ANGLE': 20, 'DTHETA': [3, -2]}, 'full'),
    'c1prime':       ('c1prime', {'niter': 20, 'nchunks': 2}, 'psd'),
    'c2':            ('c2', {'niter': 4000, 'nchunks': 20}, 'psd'),
    'c3_el10':       ('c3_elevation', {'el_deg': 10.0, 'niter': 4}, 'scalars'),
    'c3_el45':       ('c3_elevation', {'el_deg': 45.0, 'niter': 4}, 'psd'),
    'c3_el85':       ('c3_elevation', {'el_deg': 85.0, 'niter': 4}, 'scalars'),
    'c4':            ('c4', {'niter': 4}, 'sub'),
    'c5':            ('c5', {'niter': 2}, 'sub'),
    # TEMPORAL frozen-flow path (fast/fast.py:607-637): config 1 verbatim, and a 64x64 variant
    'c1_temporal':   ('c1', {}, 'temporal'),
    # sub-harmonics (fast/funcs.py:225-258, fast/fast.py:494-531,598-603)
    'mini_subharm':  ('mini', {'SUBHARM': True}, 'subharm'),
    'mini_subharm_noao': ('mini', {'SUBHARM': True, 'AO_MODE': 'NOAO', 'L0': 25.0}, 'subharm'),
    'c1prime_subharm': ('c1prime', {'niter': 8, 'nchunks': 2, 'SUBHARM': True}, 'subharm'),
    'mini_temporal': ('mini', {'TEMPORAL': True, 'NITER': 60, 'NCHUNKS': 3, 'DT': 0.002, 'COHERENT': True},
                      'temporal'),
}

SCALARS = ['W0', 'W0_sat', 'dx', 'Npxls', 'Npxls_pup', 'L', 'paa', 'r0', 'theta0', 'tau0',
           'r0_los', 'theta0_los', 'tau0_los', 'k', 'diffraction_limit', 'aniso_servo_error',
           'alias_error', 'noise_error', 'fitting_error', 'phs_var', 'logamp_var']


def run_case(name):
    factory, kw, level = CASES[name]
    p = getattr(configs, factory)(**kw)
    guard = {}
    orig = fast.funcs.generate_random_coefficients

    def spy(shape):
        r = orig(shape)
        if 'noise_head' not in guard:
            guard['noise_head'] = r.reshape(-1)[:4].copy()
            guard['noise_shape'] = np.array(shape)
        return r

    orig_fft = fast.funcs.make_phase_fft

    def spy_fft(*a, **k):
        out = orig_fft(*a, **k)
        guard.setdefault('first_screens', out.copy())
        return out

    fast.funcs.generate_random_coefficients = spy
    fast.funcs.make_phase_fft = spy_fft
    try:
        sim = fast.Fast(dict(p))
        res = sim.run()
    finally:
        fast.funcs.generate_random_coefficients = orig
        fast.funcs.make_phase_fft = orig_fft

    d = {'case_factory': np.array(factory), 'case_kwargs': np.array(repr(kw)),
         'r': res._r, 'logamp': sim.logamp.copy(), 'noise_head': guard['noise_head'],
         'noise_shape': guard['noise_shape'],
         'H_TURB': np.asarray(p['H_TURB']), 'CN2_TURB': np.asarray(p['CN2_TURB']),
         'WIND_SPD': np.asarray(p['WIND_SPD']),
         'h': sim.h, 'cn2': sim.cn2, 'wind_vector': sim.wind_vector, 'df': np.float64(sim.freq.main.df),
         'phs_var_weights': np.asarray(sim.phs_var_weights),
         'link_budget_keys': np.array(list(sim.link_budget.keys())),
         'link_budget_vals': np.array(list(sim.link_budget.values()), dtype=float),
         'pupil': sim.pupil, 'pupil_mode': sim.pupil_mode,
         'phs_last_re0': sim.phs[0].copy(), 'phs_last_im0': sim.phs[sim.Niter_per_chunk // 2].copy()}
    if level == 'temporal':
        d['pixel_shifts'] = sim.pixel_shifts
        d['temporal_logamp_powerspec'] = sim.temporal_logamp_powerspec
        d['layer_screens_sub'] = guard['first_screens'][:, ::2, ::2].copy()
        d['powerspec_per_layer_sub'] = sim.powerspec_per_layer[:, ::2, ::2].copy()
        d['phs_last_all'] = sim.phs.copy()
        d['powerspec'] = sim.powerspec
        d['logamp_powerspec'] = sim.logamp_powerspec
    if level == 'subharm':
        d['powerspec_subharm'] = sim.powerspec_subharm
        d['powerspec_subharm_per_layer'] = np.asarray(sim.powerspec_subharm_per_layer)
        d['phs_var_subharm'] = sim.phs_var_subharm
        d['powerspec'] = sim.powerspec
        d['logamp_powerspec'] = sim.logamp_powerspec
        d['phs_last_all'] = sim.phs.copy()
    for s in SCALARS:
        d[s] = np.float64(getattr(sim, s))
    N = sim.Npxls
    if level in ('full', 'psd'):
        d['powerspec'] = sim.powerspec
        d['logamp_powerspec'] = sim.logamp_powerspec
        d['lf_mask'] = np.asarray(sim.lf_mask, dtype=float)
    if level == 'full':
        for a in ('turb_powerspec', 'G_ao', 'alias_powerspec', 'noise_powerspec',
                  'powerspec_per_layer', 'pupil_filter'):
            d[a] = np.asarray(getattr(sim, a), dtype=float)
    if level == 'sub':
        # large grids: every 8th row/column plus the two central rows, and global sums
        d['powerspec_sub'] = sim.powerspec[::8, ::8].copy()
        d['powerspec_mid'] = sim.powerspec[N // 2 - 1:N // 2 + 1].copy()
        d['logamp_powerspec_sub'] = sim.logamp_powerspec[::8, ::8].copy()
        d['powerspec_sum'] = np.float64(sim.powerspec.sum())
        d['logamp_powerspec_sum'] = np.float64(sim.logamp_powerspec.sum())
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **d)
    print(f'{name}: N={N} Npup={sim.Npxls_pup} niter={sim.Niter} -> {path} '
          f'({os.path.getsize(path) / 1024:.0f} KiB)  r[:3]={res._r[:3]}')


if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n)
