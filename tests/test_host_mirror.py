"""CPU tests of the host-side mirror of the reference interface (no GPU): config handling,
profile generators, pupil / mode construction and geometry helpers against the golden outputs
of the unmodified reference, and agreement of the two config tables."""
import numpy as np
import pytest

from conftest import load_golden


def test_config_parser_contract(tmp_path):
    from fast_b200 import conf
    d = {'NITER': 10}
    c = conf.ConfigParser(d)
    assert c.config is d                      # kept by reference and completed in place
    assert d['NCHUNKS'] == 10 and d['AO_MODE'] == 'AO' and d['DTHETA'] == [4, 0]
    assert set(conf.DEFAULTS) <= set(d)
    f = tmp_path / 'cfg.py'
    f.write_text("p = {'NITER': 4, 'NCHUNKS': 2}\n")
    c2 = conf.ConfigParser(str(f))
    assert c2.config['NITER'] == 4 and c2.config['WVL'] == 1550e-9
    with pytest.raises(Exception, match='Require .py config file'):
        conf.ConfigParser('cfg.yaml')
    with pytest.raises(Exception, match='Either config file name or params dict required'):
        conf.ConfigParser(12)


def test_turbulence_models_match_reference():
    from fast_b200 import turbulence_models as tm
    g, _ = load_golden('c2')
    h, cn2, w = tm.HV57_Bufton_profile(4)
    np.testing.assert_allclose(h, g['H_TURB'], rtol=1e-14)
    np.testing.assert_allclose(cn2, g['CN2_TURB'], rtol=1e-14)
    np.testing.assert_allclose(w, g['WIND_SPD'], rtol=1e-14)
    hh = np.linspace(0, 20000, 10)
    assert tm.HV57(hh).dtype == float and len(tm.Bufton_wind(hh)) == 10       # test/tests_pytest.py:12-27


@pytest.mark.parametrize('name', ['mini_ao', 'mini_axicon', 'mini_up_w0', 'c1prime', 'c2', 'c4'])
def test_pupil_and_mode_match_reference(name):
    from fast_b200 import funcs
    g, p = load_golden(name)
    N, dx, npup = int(g['Npxls']), float(g['dx']), int(g['Npxls_pup'])
    pupil = funcs.compute_pupil(N, dx, p['D_GROUND'], p['OBSC_GROUND'])
    ptype = 'axicon' if p['AXICON'] else 'gauss'
    mode, W0 = funcs.compute_gaussian_mode(pupil, dx, p['W0'], D=p['D_GROUND'], obsc=p['OBSC_GROUND'], ptype=ptype)
    lo, hi = (N - npup) // 2, (N + npup) // 2
    np.testing.assert_allclose(pupil[lo:hi, lo:hi], g['pupil'], rtol=1e-13)
    np.testing.assert_allclose(mode[lo:hi, lo:hi], g['pupil_mode'], rtol=1e-9)
    assert W0 == pytest.approx(float(g['W0']), rel=1e-10)


def test_geometry_helpers():
    from fast_b200 import funcs
    g, p = load_golden('c3_el45')
    assert funcs.l_path(550e3, p['ZENITH_ANGLE']) == pytest.approx(float(g['L']), rel=1e-13)
    wc = funcs.calculate_wind_correction(g['h'], p['ANISO_DL'], p['TLOOP'])
    assert wc.shape == (4, 2) and np.all(wc[:, 1] == 0)
    with pytest.raises(TypeError, match="axicon"):
        funcs.compute_gaussian_mode(np.ones((8, 8)), 0.1, 'opt', ptype='axicon')
    with pytest.raises(Exception, match='ptype must be one of'):
        funcs.compute_gaussian_mode(np.ones((8, 8)), 0.1, 0.2, ptype='bessel')


@pytest.mark.parametrize('name', ['mini_tt', 'mini_modal', 'mini_ao'])
def test_host_masks_match_reference(name):
    from fast_b200 import ao_power_spectra as aps
    from fast_b200.fast import SpatialFrequencies
    g, p = load_golden(name)
    freq = SpatialFrequencies(int(g['Npxls']), float(g['dx']))
    zmax, modal = p['ZMAX'], p['MODAL']
    if p['AO_MODE'] == 'TT':
        zmax, modal = 3, True
    m = aps.mask_lf(freq.main, p['DSUBAP'], modal=modal, modal_mult=p['MODAL_MULT'], Zmax=zmax, D=p['D_GROUND'])
    np.testing.assert_allclose(np.asarray(m, dtype=float), g['lf_mask'], rtol=1e-12, atol=1e-15)


def test_product_and_oracle_config_tables_agree():
    from fast_b200 import configs as a
    from oracle import configs as b
    for f, kw in [('c2', {}), ('c4', {}), ('c5', {}), ('c1prime', {}), ('mini', {}),
                  ('c3_elevation', {'el_deg': 30.})]:
        pa, pb = getattr(a, f)(**kw), getattr(b, f)(**kw)
        assert pa.keys() == pb.keys()
        for k in pa:
            assert np.all(np.asarray(pa[k] == pb[k])), (f, k)


def test_fast_result_properties():
    from fast_b200 import FastResult
    r = np.array([0.5, 0.25, 1.0])
    res = FastResult(r, 2e-6)
    np.testing.assert_allclose(res.power, 2e-6 * r)
    np.testing.assert_allclose(res.dB_rel, 10 * np.log10(r))
    np.testing.assert_allclose(res.dBm, 10 * np.log10(r * 2e-6 / 1e-3))
    assert res.scintillation_index == pytest.approx((r / r.mean()).var())
    assert res.avg_power_dB_rel == pytest.approx(10 * np.log10(r.mean()))
    assert 'Scintillation index' in str(res)


@pytest.mark.parametrize('name', ['c1_temporal', 'mini_temporal'])
def test_temporal_host_bookkeeping_matches_oracle(name):
    """The product's TEMPORAL coordinate bookkeeping (wrap / sort / roll / clamp) against the
    oracle restatement (itself pinned to the reference in test_oracle_vs_golden)."""
    from fast_b200 import temporal
    from oracle import fast_oracle as fo
    g, p = load_golden(name)
    N, npup = int(g['Npxls']), int(g['Npxls_pup'])
    lo = (N - npup) // 2
    shifts = g['pixel_shifts']
    interp = np.arange(lo, lo + npup).astype(float)[None, None, None, :] + shifts[:, :, :, None]
    for chunk in range(4):
        xi, xf, yi, yf = temporal.sample_coordinates(interp, N)
        at = fo.temporal_sample_coords(interp, N)
        atc = np.minimum(at, N - 1.0)
        np.testing.assert_allclose(xi + xf.astype(float), atc[:, 0], atol=2e-5)
        np.testing.assert_allclose(yi + yf.astype(float), atc[:, 1], atol=2e-5)
        assert xi.min() >= 0 and xi.max() <= N - 2 and 0 <= xf.min() and xf.max() <= 1
        interp = interp + shifts[:, :, -1, None, None]


def test_spatial_frequency_helpers_follow_the_reference():
    """fast/fast.py:830-833, 866-875, 923-928: main / log-amplitude grids and the conjugate real-space sampling."""
    from fast_b200 import fast as fb
    fr = fb.SpatialFrequencies(64, 0.04)
    assert fr.main.df == pytest.approx(2 * np.pi / (64 * 0.04))
    dx, dy = fr.main.realspace_sampling()
    assert dx == pytest.approx(0.04) and dy == pytest.approx(0.04)
    fr.make_logamp_freqs()
    assert fr.logamp is fr.main
    fr.make_logamp_freqs(Nx=32, dx=0.02, Ny=16, dy=0.05)
    assert fr.logamp.fx.shape == (16, 32)
    assert fr.logamp.dfx == pytest.approx(2 * np.pi / (32 * 0.02)) and fr.logamp.dfy == pytest.approx(2 * np.pi / (16 * 0.05))
    assert fr.logamp.realspace_sampling() == pytest.approx((0.02, 0.05))
    fr.make_main_freqs(128, 0.01)
    assert fr.main.fx_axis.shape == (128,) and fr.main.fx_axis[64] == 0.0


def _sample_coordinates_plain(interp_coords, N):
    """The literal numpy statement of fast/fast.py:621-633 that temporal.sample_coordinates must reproduce."""
    coord = np.sort(interp_coords % N, axis=-1)
    gaps = np.abs(np.diff(coord, axis=-1))
    roll = gaps.argmax(-1)
    roll[np.isclose(gaps, 1).all(-1)] = 0
    npup = coord.shape[-1]
    at = np.take_along_axis(coord, (np.arange(npup) + roll[..., None]) % npup, axis=-1)
    at = np.minimum(at, N - 1.0)
    i0 = np.minimum(np.floor(at).astype(np.int32), N - 2)
    frac = (at - i0).astype(np.float32)
    return i0[..., 0, :, :], frac[..., 0, :, :], i0[..., 1, :, :], frac[..., 1, :, :]


@pytest.mark.parametrize('scale', [0.0, 3.0, 40.0, 400.0])
def test_temporal_sample_coordinates_equal_the_plain_statement(scale):
    """No wrap, partial wraps (negative and beyond N), many wraps; with and without a leading chunk axis."""
    from fast_b200 import temporal
    rng = np.random.default_rng(int(scale) + 1)
    L, J, P, N = 3, 7, 22, 64
    pup = np.stack([np.arange(P) + 21.0, np.arange(P) + 21.0])
    shifts = (rng.random((L, 2, J)) - 0.5) * scale
    ic = pup[None, :, None, :] + shifts[:, :, :, None]
    ic[0, 0, 0] = np.arange(P) + 42.0            # reaches exactly N - 1 + ... and the last knot
    stacked = np.stack([ic, ic + shifts[:, :, -1, None, None], ic - 17.25])
    for arr in (ic, stacked):
        got = temporal.sample_coordinates(arr, N)
        want = _sample_coordinates_plain(arr, N)
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
            assert g.flags['C_CONTIGUOUS']
