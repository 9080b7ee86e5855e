"""Drop-in acceptance: the reference's own test matrix (test/tests_pytest.py:36-127 -- smoke
assertions: finite results, dtypes, attribute values) replayed against `fast_b200`, plus the
edge cases of the C ABI.  The numerical parity of the same modes is in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def fast():
    import fast_b200
    return fast_b200


@pytest.fixture(scope='module')
def p(fast):
    # test/test_params.py verbatim (TEMPORAL uplink example), as written by the reference
    return dict(fast.configs.base(), LOGLEVEL='ERROR')


def run_sim(fast, params):
    sim = fast.Fast(dict(params))
    sim.run()
    assert np.isfinite(sim.I).all()
    return sim


def test_sim_default(fast, p):                                   # tests_pytest.py:36-43
    sim = fast.Fast(dict(p))
    sim.run()
    assert sim.Npxls == 164 and sim.temporal
    assert np.isfinite(sim.result.power).all()
    assert np.isfinite(sim.result.dB_rel).all()
    assert np.isfinite(sim.result.dB_abs).all()
    assert len(sim.I) == p['NITER']


def test_sim_mean_irradiance(fast, p):                           # :45-48
    sim = fast.Fast(dict(p, TEMPORAL=False))
    psf = sim.compute_mean_irradiance()
    assert np.isfinite(psf)
    # the analytic mean coupled power agrees with the Monte-Carlo mean to a few per cent
    mc = fast.Fast(dict(p, TEMPORAL=False, NITER=20000, NCHUNKS=2, SEED=3)).run()
    assert psf == pytest.approx(mc.power.mean(), rel=0.05)


def test_sim_fftw_flag_is_accepted(fast, p):                     # :50-54 (pyfftw is optional there)
    run_sim(fast, dict(p, FFTW=True, TEMPORAL=False))


def test_sim_randomScrns(fast, p):                               # :56-59
    run_sim(fast, dict(p, TEMPORAL=False))


def test_sim_subharm(fast, p):                                   # :61-65
    run_sim(fast, dict(p, SUBHARM=True, TEMPORAL=False))


def test_sim_subharm_ignored_in_temporal(fast, p):               # fast/fast.py:221-224
    sim = run_sim(fast, dict(p, SUBHARM=True))
    assert sim.subharmonics is False


def test_sim_obsc(fast, p):                                      # :67-70
    run_sim(fast, dict(p, OBSC_GROUND=0.1))


def test_sim_obsc_sat(fast, p):                                  # :72-75
    run_sim(fast, dict(p, OBSC_SAT=0.05))


def test_sim_axicon(fast, p):                                    # :77-82
    run_sim(fast, dict(p, W0=0.1, AXICON=True, OBSC_GROUND=0.1))
    with pytest.raises(TypeError, match="axicon"):
        fast.Fast(dict(p, AXICON=True))                          # W0 == 'opt' is not supported with axicon


def test_sim_L_SAT(fast, p):                                     # :84-88
    sim = fast.Fast(dict(p, L_SAT=500e3))
    assert sim.L == 500e3


def test_sim_L0(fast, p):                                        # :90-93
    run_sim(fast, dict(p, L0=25))


def test_sim_down(fast, p):                                      # :95-98
    run_sim(fast, dict(p, PROP_DIR='down'))


def test_sim_NOAO(fast, p):                                      # :100-103
    run_sim(fast, dict(p, PROP_DIR='down', AO_MODE='NOAO'))


def test_sim_TT(fast, p):                                        # :105-108
    run_sim(fast, dict(p, PROP_DIR='down', AO_MODE='TT'))


def test_sim_noise(fast, p):                                     # :110-113
    run_sim(fast, dict(p, PROP_DIR='down', NOISE=1))


def test_sim_modal(fast, p):                                     # :115-118
    run_sim(fast, dict(p, PROP_DIR='down', NOISE=1, MODAL=True))


def test_sim_coherent(fast, p):                                  # :122-127
    sim = fast.Fast(dict(p, PROP_DIR='down', COHERENT=True))
    sim.run()
    assert sim.I.dtype == complex
    sim = fast.Fast(dict(p, PROP_DIR='down', COHERENT=True, TEMPORAL=False))
    sim.run()
    assert sim.I.dtype == complex


@pytest.mark.parametrize('scheme', ['OOK', 'BPSK', 'QAM'])
def test_fast_fsoc(fast, p, scheme):                             # :129-157 (EsN0 left at its default None)
    sim = fast.comms.FastFSOC(dict(p, MODULATION=scheme))
    sim.run()
    assert np.isfinite(sim.I).all()
    assert hasattr(sim.modulator, "sep")
    assert np.isfinite(sim.modulator.sep)
    assert np.isfinite(sim.modulator.evm)


def test_fast_fsoc_with_noise(fast, p):
    sim = fast.comms.FastFSOC(dict(p, MODULATION='16-QAM', EsN0=12.0, TEMPORAL=False, NITER=1000, NCHUNKS=2))
    sim.run()
    assert 0.0 < sim.modulator.sep < 1.0 and sim.modulator.evm > 0.0


def test_ber_ook(fast, p):                                       # :168-176
    sim = run_sim(fast, p)
    assert np.isfinite(fast.comms.ber_ook(10, sim.result.power))
    assert np.isfinite(fast.comms.ber_ook(10))


def test_ber_qam(fast, p):                                       # :178-187
    sim = run_sim(fast, p)
    assert np.isfinite(fast.comms.ber_qam(4, 10, samples=sim.result.power))
    assert np.isfinite(fast.comms.ber_qam(4, 10))


def test_consumer_style_postprocessing(fast, p):
    """What fast/comms.py consumes (FastFSOC.run -> Modulator(self.result.power, ...),
    fast/comms.py:159-162): a finite 1-D float array of length NITER and the fade statistics."""
    sim = run_sim(fast, dict(p, TEMPORAL=False, NITER=2000, NCHUNKS=2, SEED=1))
    pw = sim.result.power
    assert pw.ndim == 1 and pw.dtype == float and len(pw) == 2000
    fade_prob = (sim.result.dB_rel < -6).mean()
    assert 0 <= fade_prob < 0.5
    st = sim.result_stats()
    assert st['n'] == 2000 and st['hist'].sum() == 2000


def test_config_file_path(fast, tmp_path):                       # :31-33, Fast("…/test_params.py")
    f = tmp_path / 'my_params.py'
    f.write_text("import numpy\np = {'NITER': 40, 'NCHUNKS': 2, 'NPXLS': 64, 'DX': 0.04, 'D_GROUND': 0.8,\n"
                 "     'DSUBAP': 0.1, 'LOGLEVEL': 'ERROR', 'SEED': 2}\n")
    sim = fast.Fast(str(f))
    assert sim.params['H_TURB'].tolist() == [0, 10e3]            # defaults filled (fast/conf.py:67-115)
    assert np.isfinite(sim.run().power).all()


# ---------------------------------------------------------------------------------------------
# edge cases at the C ABI
# ---------------------------------------------------------------------------------------------
def test_zero_pairs_is_a_no_op(fast):
    sim = fast.Fast(dict(fast.configs.mini(), LOGLEVEL='ERROR'))
    a, b = sim.screen_detect(0, 0)
    assert a.numel() == 0 and b.numel() == 0


def test_single_pair_and_odd_splits(fast):
    sim = fast.Fast(dict(fast.configs.mini(niter=2, nchunks=1), LOGLEVEL='ERROR'))
    r = sim.run()._r
    assert r.shape == (2,) and np.isfinite(r).all()
    sim = fast.Fast(dict(fast.configs.mini(niter=14, nchunks=7), LOGLEVEL='ERROR'))      # 1 pair per chunk
    r7 = sim.run()._r
    a, b = sim.screen_detect(0, 7)
    np.testing.assert_array_equal(r7, np.stack([a.cpu().numpy(), b.cpu().numpy()], 1).reshape(-1))


def test_pupil_as_large_as_the_grid_and_one_pixel_pupil(fast):
    """Ragged crops: n_pup == N (no pruning) and n_pup == 1, through the raw wrapper."""
    lib = fast._lib
    N = 64
    dev = torch.device('cuda')
    w = torch.rand(N, N, dtype=torch.float64, device=dev) * 1e-4
    weight = lib.make_weight(w, 2.0)
    for P, lo in ((N, 0), (1, 31), (3, 0), (5, 59)):
        U = torch.ones(P, P, dtype=torch.float32, device=dev)
        rp = lib.RunParams()
        rp.n, rp.n_pup, rp.lo, rp.n_pairs, rp.pairs_per_chunk, rp.seed = N, P, lo, 6, 3, 11
        rp.u_sum, rp.sigma_chi = float(P * P), 0.0
        outs = []
        for algo in (lib.ALGO_RADIX_PAIR, lib.ALGO_RADIX, lib.ALGO_DIRECT):
            rp.algo = algo
            ws = torch.empty(lib.screen_detect_workspace_bytes(rp), dtype=torch.uint8, device=dev)
            a = torch.empty(6, dtype=torch.float32, device=dev)
            b = torch.empty(6, dtype=torch.float32, device=dev)
            lib.screen_detect(rp, weight, U, a, b, ws)
            outs.append(torch.cat([a, b]).cpu().numpy())
        assert np.isfinite(outs[0]).all() and (outs[0] <= 1 + 1e-5).all()
        np.testing.assert_allclose(outs[0], outs[2], rtol=2e-4)
        np.testing.assert_allclose(outs[1], outs[2], rtol=2e-4)
        if P == 1:
            np.testing.assert_allclose(outs[0], 1.0, rtol=1e-6)      # |exp(i phi)|^2 = 1


def test_unsupported_sizes_fail_loudly(fast):
    lib = fast._lib
    rp = lib.RunParams()
    rp.n, rp.n_pup, rp.lo, rp.n_pairs, rp.pairs_per_chunk, rp.u_sum = 4098, 8, 0, 1, 1, 1.0
    with pytest.raises(lib.FastbError, match='4096'):
        lib.screen_detect_workspace_bytes(rp)
    rp.n, rp.algo = 96, lib.ALGO_RADIX
    with pytest.raises(lib.FastbError, match='power of two'):
        lib.screen_detect_workspace_bytes(rp)
    with pytest.raises(Exception, match='NPXLS must be even'):
        fast.Fast(dict(fast.configs.mini(), NPXLS=65, LOGLEVEL='ERROR'))


def test_largest_radix_grid_2048(fast):
    """Maximum radix size: N = 2048 (two realisations), radix vs direct."""
    p = dict(fast.configs.c5(niter=2, nchunks=1, seed=3), NPXLS=2048, DX=0.0025, LOGLEVEL='ERROR')
    sim = fast.Fast(p)
    assert sim.Npxls == 2048 and sim.Npxls_pup == 322
    a2, b2 = sim.screen_detect(0, 1, algo=fast._lib.ALGO_DIRECT)
    for algo in (fast._lib.ALGO_RADIX, fast._lib.ALGO_RADIX_PAIR):
        a1, b1 = sim.screen_detect(0, 1, algo=algo)
        np.testing.assert_allclose(a1.cpu().numpy(), a2.cpu().numpy(), rtol=2e-4)
        np.testing.assert_allclose(b1.cpu().numpy(), b2.cpu().numpy(), rtol=2e-4)


def test_orbit_sweep_driver_with_injected_geometry(fast):
    """FAST_sat_orbit (fast/complete_orbit_simulation.py:190-232) with precomputed pass geometry:
    one Fast per sample with the reference's keys, zero-Cn2 layers dropped, then one batched sweep."""
    cos = fast.complete_orbit_simulation
    base = fast.configs.mini()
    base.update({'NITER': 200, 'NCHUNKS': 2, 'SEED': 2, 'PROP_DIR': 'down'})
    base['CN2_TURB'] = list(base['CN2_TURB']) + [0.0]
    base['H_TURB'] = list(base['H_TURB']) + [30000.0]
    base['WIND_SPD'] = list(base['WIND_SPD']) + [5.0]
    base['WIND_DIR'] = list(base['WIND_DIR']) + [10.0]
    alt = np.array([20.0, 45.0, 80.0])
    geometry = (np.array([[8.0, 1.0], [9.5, 0.5], [10.4, 0.1]]),      # point-ahead angle ["]
                np.array([[0.3, 0.0], [0.8, 0.1], [2.5, 0.2]]),       # downlink anisoplanatism ["]
                alt, np.array([10.0, 40.0, 170.0]), 550e3 / np.sin(np.radians(alt)))
    sims = cos.FAST_sat_orbit(base, {}, None, geometry=geometry)
    assert sorted(k for k in sims if k != 'altitudes') == ['simulation_0', 'simulation_1', 'simulation_2']
    np.testing.assert_array_equal(sims['altitudes'], alt)
    lst = [sims[f'simulation_{i}'] for i in range(3)]
    assert all(len(s.h) == len(base['H_TURB']) - 1 for s in lst)
    # lower elevation: longer path through the turbulence -> larger residual phase variance
    assert lst[0].phs_var > lst[1].phs_var > lst[2].phs_var
    res = fast.sweep.run_sweep(lst)
    assert all(np.isfinite(r.power).all() and len(r.power) == 200 for r in res)
    assert res[0].dB_rel.mean() < res[2].dB_rel.mean()
    one = cos.FAST_sat(np.array([2.0, 0.0]), dict(fast.configs.mini(), NITER=20, NCHUNKS=2))
    np.testing.assert_allclose(one.params['ANISO_DL'], np.array([2.0, 0.0]) * one.params['TLOOP'])
