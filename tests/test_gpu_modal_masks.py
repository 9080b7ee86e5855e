"""fastb_zernike_filter (device Bessel / Zernike masks) against the CPU oracle's
restatement of fast/ao_power_spectra.py:10-141 (itself pinned to the reference by the golden cases
mini_tt / mini_modal / mini_lgsao)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _grid(n, df):
    ax = (np.arange(n) - n // 2) * df
    fx, fy = np.meshgrid(ax, ax)
    return fx, fy, np.sqrt(fx ** 2 + fy ** 2)


@pytest.mark.parametrize('n,df', [(64, 2.45), (164, 0.9), (3, 0.011), (257, 1.7)])
@pytest.mark.parametrize('first,last', [(1, 3), (1, 4), (1, 21), (1, 66), (2, 3), (4, 10)])
def test_zernike_squared_filter_matches_oracle(n, df, first, last):
    import torch
    from fast_b200 import _lib
    from oracle import fast_oracle as fo
    fx, fy, fabs = _grid(n, df)
    want = fo.zernike_sq_filter(fabs, fx, fy, 0.8, last, first).real
    got = _lib.zernike_filter(n, df, torch.device('cuda'), noll_first=first, noll_last=last, diameter=0.8).cpu().numpy()
    # the terms are O(1) near DC and decay as |f|^-3: compare on an absolute scale
    # (CUDA jn vs scipy jv agree to ~3e-13 absolute on these O(1) values; K1 parity needs 1e-9)
    assert np.max(np.abs(got - want)) < 1e-11 * max(1.0, np.max(np.abs(want)))
    assert got[n // 2, n // 2] == (1.0 if first == 1 else 0.0)


@pytest.mark.parametrize('zmax', [None, 3, 10, 36])
def test_modal_lf_mask_matches_oracle(zmax):
    import torch
    from fast_b200 import _lib
    from oracle import fast_oracle as fo
    n, df, d = 128, 2.45, 0.1
    fx, fy, _ = _grid(n, df)
    want = np.asarray(fo.lf_mask(fx, fy, d, modal=True, modal_mult=0.7, Zmax=zmax, D=0.8), dtype=float)
    if zmax is None:
        got = _lib.zernike_filter(n, df, torch.device('cuda'), noll_first=1, noll_last=0, d_wfs=d, modal_mult=0.7,
                                  clip_box=True)
    else:
        got = _lib.zernike_filter(n, df, torch.device('cuda'), noll_first=1, noll_last=zmax, diameter=0.8, d_wfs=d,
                                  clip_box=True)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-11)


def test_gtilt_term_and_argument_errors():
    import torch
    from scipy.special import jv
    from fast_b200 import _lib
    from oracle import fast_oracle as fo
    n, df = 96, 1.3
    fx, fy, fabs = _grid(n, df)
    want = fo.zernike_sq_filter(fabs, fx, fy, 0.5, 1).real + jv(1, fabs * 0.5 / 2.) ** 2
    got = _lib.zernike_filter(n, df, torch.device('cuda'), 1, 1, diameter=0.5, gtilt=True).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-11)
    with pytest.raises(_lib.FastbError, match='diameter'):
        _lib.zernike_filter(n, df, torch.device('cuda'), 1, 3, diameter=0.0)
    with pytest.raises(_lib.FastbError, match='d_wfs'):
        _lib.zernike_filter(n, df, torch.device('cuda'), 1, 3, diameter=0.5, clip_box=True)


@pytest.mark.parametrize('kw', [{'AO_MODE': 'TT'}, {'AO_MODE': 'LGSAO'}, {'MODAL': True, 'ZMAX': 15},
                                {'MODAL': True, 'MODAL_MULT': 0.8}])
def test_modal_psd_with_subharmonics_matches_oracle(kw):
    """K1 fed by the device masks, on the main grid and the three 3 x 3 sub-harmonic levels."""
    import fast_b200
    from oracle import configs, fast_oracle as fo
    p = configs.mini()
    p.update(kw)
    p.update({'SUBHARM': True, 'NITER': 8, 'NCHUNKS': 2})
    sim = fast_b200.Fast(dict(p))
    init = fo.build(dict(p))
    scale = np.max(np.abs(init['powerspec']))
    assert np.max(np.abs(sim.powerspec - init['powerspec'])) < 1e-9 * scale
    np.testing.assert_allclose(sim.lf_mask, np.asarray(init['lf_mask'], dtype=float), rtol=0, atol=1e-12)
    _, want_sub, _ = fo.subharm_psd(init)
    assert np.max(np.abs(sim.powerspec_subharm - want_sub)) < 1e-9 * np.max(np.abs(want_sub))
