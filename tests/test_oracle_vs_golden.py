"""Pins the CPU oracle (oracle/fast_oracle.py) to outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import fast_oracle as fo
from conftest import load_golden

MINI = ['mini_ao', 'mini_noise_L0', 'mini_noao', 'mini_tt', 'mini_modal', 'mini_lgsao',
        'mini_axicon', 'mini_coherent', 'mini_up_w0']
SCALARS = ['W0', 'W0_sat', 'dx', 'L', 'paa', 'r0', 'theta0', 'tau0', 'r0_los', 'theta0_los',
           'tau0_los', 'k', 'diffraction_limit', 'aniso_servo_error', 'alias_error',
           'noise_error', 'fitting_error', 'phs_var', 'logamp_var']


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.max(np.abs(b))
    return 0.0 if scale == 0 else float(np.max(np.abs(a - b)) / scale)


def check_init(g, init, arrays):
    assert init['N'] == int(g['Npxls']) and init['Npup'] == int(g['Npxls_pup'])
    for s in SCALARS:
        want = float(g[s])
        got = float(init[s]) if s in init else float(init['atm'][s])
        assert got == pytest.approx(want, rel=1e-10, abs=1e-300), s
    np.testing.assert_allclose(init['phs_var_weights'], g['phs_var_weights'], rtol=1e-10)
    np.testing.assert_allclose(init['atm']['wind_vector'], g['wind_vector'], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(list(init['link_budget'].values()), g['link_budget_vals'], rtol=1e-11)
    np.testing.assert_allclose(init['pupil'], g['pupil'], rtol=1e-13)
    np.testing.assert_allclose(init['pupil_mode'], g['pupil_mode'], rtol=1e-9)
    for a in arrays:
        want = g[a]
        got = np.broadcast_to(np.asarray(init[a], dtype=float), want.shape)
        assert rel(got, want) < 1e-12, a


@pytest.mark.parametrize('name', MINI)
def test_mini_cases_every_term(name):
    g, p = load_golden(name)
    init = fo.build(p)
    check_init(g, init, ['turb_powerspec', 'G_ao', 'alias_powerspec', 'noise_powerspec',
                         'powerspec_per_layer', 'powerspec', 'logamp_powerspec', 'lf_mask',
                         'pupil_filter'])
    r, chi = fo.run_mc(init, np.random.default_rng(p['SEED']))
    np.testing.assert_allclose(chi, g['logamp'], rtol=1e-12)
    assert rel(r, g['r']) < 1e-10


def test_profile_generator_matches_reference():
    g, p = load_golden('c2')
    h, cn2, w = fo.hv57_bufton_profile(4)
    np.testing.assert_allclose(h, g['H_TURB'], rtol=1e-14)
    np.testing.assert_allclose(cn2, g['CN2_TURB'], rtol=1e-14)
    np.testing.assert_allclose(w, g['WIND_SPD'], rtol=1e-14)


def test_c1prime_auto_grid_164():
    g, p = load_golden('c1prime')
    init = fo.build(p)
    assert init['N'] == 164
    check_init(g, init, ['powerspec', 'logamp_powerspec', 'lf_mask'])
    r, _ = fo.run_mc(init, np.random.default_rng(p['SEED']))
    assert rel(r, g['r']) < 1e-10


def test_c2_first_chunks_and_screens():
    g, p = load_golden('c2')
    init = fo.build(p)
    check_init(g, init, ['powerspec', 'logamp_powerspec', 'lf_mask'])
    # replay the first two of the 20 chunks (200 realisations each) with the reference's stream
    niter, nch, N = p['NITER'], p['NCHUNKS'], init['N']
    J = niter // nch
    rng = np.random.default_rng(p['SEED'])
    chi = fo.draw_logamp(rng, niter, init['logamp_var'])
    U = init['pupil'] * init['pupil_mode']
    for c in range(2):
        noise = fo.draw_complex(rng, (J // 2, N, N))
        if c == 0:
            np.testing.assert_array_equal(noise.reshape(-1)[:4], g['noise_head'])
        phs = fo.screens_from_noise(noise, init['powerspec'], init['df'], init['lo'], init['hi'])
        r = fo.detector(phs, U, chi[c * J:(c + 1) * J])
        assert rel(r, g['r'][c * J:(c + 1) * J]) < 1e-10
    # survey known answers of the full 4000-realisation reference run
    db = 10 * np.log10(g['r'])
    assert db.mean() == pytest.approx(-3.06228, abs=1e-5)
    assert db.var() == pytest.approx(2.64810, abs=1e-5)


@pytest.mark.parametrize('name', ['c3_el10', 'c3_el45', 'c3_el85'])
def test_c3_orbit_sample_keys(name):
    g, p = load_golden(name)
    init = fo.build(p)
    check_init(g, init, ['powerspec', 'logamp_powerspec'] if 'powerspec' in g else [])
    r, _ = fo.run_mc(init, np.random.default_rng(p['SEED']))
    assert rel(r, g['r']) < 1e-10


@pytest.mark.parametrize('name', ['c4', 'c5'])
def test_large_grids_subsampled(name):
    g, p = load_golden(name)
    if name == 'c5':
        pytest.importorskip('scipy')
    init = fo.build(p)
    check_init(g, init, [])
    N = init['N']
    assert rel(init['powerspec'][::8, ::8], g['powerspec_sub']) < 1e-12
    assert rel(init['powerspec'][N // 2 - 1:N // 2 + 1], g['powerspec_mid']) < 1e-12
    assert rel(init['logamp_powerspec'][::8, ::8], g['logamp_powerspec_sub']) < 1e-12
    assert init['powerspec'].sum() == pytest.approx(float(g['powerspec_sum']), rel=1e-12)
    r, _ = fo.run_mc(init, np.random.default_rng(p['SEED']))
    assert rel(r, g['r']) < 1e-10


def test_simpson_weights_are_the_integral():
    f = fo.freq_axis(64, 0.04)
    w = fo.simpson_weights(f)
    P = np.random.default_rng(0).random((64, 64))
    assert w @ P @ w == pytest.approx(fo.simpson2d(P, f), rel=1e-13)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors)."""
    def h(*a):
        return [int(x) for x in fo.philox4x32_10(*a)]
    assert h(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert h(f, f, f, f, f, f) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_7_round_known_answers():
    """Random123 known-answer vectors for philox4x32-7 (kat_vectors): the round function of the
    opt-in 'device-fast' noise stream."""
    def h(*a):
        return [int(x) for x in fo.philox4x32_10(*a, rounds=7)]
    assert h(0, 0, 0, 0, 0, 0) == [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]
    f = 0xffffffff
    assert h(f, f, f, f, f, f) == [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]
    assert h(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]


def test_fast_stream_noise_is_standard_normal_and_independent_of_the_default_stream():
    z = fo.device_noise_pair(seed=7, pair=3, N=128, fast=True)
    x = np.concatenate([z.real.ravel(), z.imag.ravel()])
    assert abs(x.mean()) < 4 / np.sqrt(x.size)
    assert abs(x.var() - 1) < 4 * np.sqrt(2 / x.size)
    assert abs(np.mean(z.real * z.imag)) < 4 / np.sqrt(z.size)
    # fourth moment of a standard normal is 3; lag-1 correlations along both axes vanish
    assert abs(np.mean(x ** 4) - 3) < 4 * np.sqrt(96 / x.size)
    assert abs(np.mean(z.real[:, 1:] * z.real[:, :-1])) < 4 / np.sqrt(z.size)
    assert abs(np.mean(z.real[1:] * z.real[:-1])) < 4 / np.sqrt(z.size)
    # angles lie on the 2^-15 turn lattice the contract states
    turn = (np.angle(z) / (2 * np.pi)) % 1.0
    assert np.abs(turn * 2 ** 15 - np.round(turn * 2 ** 15)).max() < 1e-6
    z0 = fo.device_noise_pair(seed=7, pair=3, N=128)
    assert abs(np.corrcoef(z.real.ravel(), z0.real.ravel())[0, 1]) < 4 / np.sqrt(z.size)


def test_device_noise_is_standard_normal():
    z = fo.device_noise_pair(seed=7, pair=3, N=128)
    x = np.concatenate([z.real.ravel(), z.imag.ravel()])
    assert abs(x.mean()) < 4 / np.sqrt(x.size)
    assert abs(x.var() - 1) < 4 * np.sqrt(2 / x.size)
    assert abs(np.mean(z.real * z.imag)) < 4 / np.sqrt(z.size)
    chi = fo.device_chi_normals(7, 0, 40000)
    assert abs(chi.mean()) < 0.02 and abs(chi.var() - 1) < 0.03
    # sub-ranges are consistent with the whole (counter-based => any range recomputable)
    np.testing.assert_array_equal(fo.device_chi_normals(7, 1001, 50), chi[1001:1051])


@pytest.mark.parametrize('name', ['c1_temporal', 'mini_temporal'])
def test_temporal_path(name):
    """TEMPORAL frozen-flow path (config 1 verbatim = c1_temporal)."""
    g, p = load_golden(name)
    init = fo.build(p)
    check_init(g, init, ['powerspec', 'logamp_powerspec'])
    ts = fo.temporal_setup(init)
    np.testing.assert_allclose(ts['pixel_shifts'], g['pixel_shifts'], rtol=1e-13, atol=1e-13)
    assert rel(ts['temporal_logamp_powerspec'], g['temporal_logamp_powerspec']) < 1e-11
    assert rel(init['powerspec_per_layer'][:, ::2, ::2], g['powerspec_per_layer_sub']) < 1e-12
    seen = {}
    r, chi, phs = fo.run_mc_temporal(init, np.random.default_rng(p['SEED']), screens_hook=lambda s: seen.update(s=s))
    assert rel(seen['s'][:, ::2, ::2], g['layer_screens_sub']) < 1e-11
    np.testing.assert_allclose(chi, g['logamp'], rtol=1e-9, atol=1e-14)
    assert rel(phs, g['phs_last_all']) < 1e-10
    assert rel(r, g['r']) < 1e-10


def test_bilinear_matches_scipy_spline_including_clamp():
    from scipy.interpolate import RectBivariateSpline
    rng = np.random.default_rng(2)
    scr = rng.normal(size=(12, 12))
    s = RectBivariateSpline(np.arange(12), np.arange(12), scr, kx=1, ky=1, s=0)
    x = np.sort(rng.uniform(0, 12, 9))
    y = np.sort(rng.uniform(0, 12, 7))
    np.testing.assert_allclose(fo.bilinear_clamped(scr, x, y), s(x, y), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('name', ['mini_subharm', 'mini_subharm_noao', 'c1prime_subharm'])
def test_subharmonics(name):
    g, p = load_golden(name)
    init = fo.build(p)
    check_init(g, init, ['powerspec', 'logamp_powerspec'])
    per_layer, W_sh, axes = fo.subharm_psd(init)
    assert rel(per_layer, g['powerspec_subharm_per_layer']) < 1e-12
    assert rel(W_sh, g['powerspec_subharm']) < 1e-12
    df_lo = axes[:, 1] - axes[:, 0]
    np.testing.assert_allclose(per_layer.sum((-1, -2)) * df_lo ** 2, g['phs_var_subharm'], rtol=1e-12)
    r, phs = fo.run_mc_subharm(init, np.random.default_rng(p['SEED']))
    assert rel(phs, g['phs_last_all']) < 1e-10
    assert rel(r, g['r']) < 1e-10
