"""K5 link metrics on the GPU (through the C ABI) against the CPU oracle (oracle/comms_oracle.py)
on the committed golden inputs (tests/golden/comms.npz = the reference's own outputs), plus
statistical parity of the device-RNG Monte-Carlo modulator with closed forms and the oracle."""
import math

import numpy as np
import pytest

from conftest import load_golden_arrays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    return load_golden_arrays('comms')


@pytest.fixture(scope='module')
def comms():
    from fast_b200 import comms
    return comms


@pytest.fixture(scope='module')
def co():
    from oracle import comms_oracle
    return comms_oracle


def test_error_curves_match_reference(g, comms):
    power, snr = g['power'], g['snr_db']
    np.testing.assert_allclose(comms.ber_ook(snr, power), g['ber_ook'], rtol=1e-11)
    assert comms.ber_ook(float(snr[3]), power) == pytest.approx(float(g['ber_ook'][3]), rel=1e-11)
    for M in (4, 16, 64):
        np.testing.assert_allclose(comms.sep_qam(M, snr, power), g[f'sep_qam_{M}'], rtol=1e-11)
        np.testing.assert_allclose([comms.ber_qam(M, float(s), power) for s in snr], g[f'ber_qam_{M}'], rtol=1e-11)


def test_error_curve_accepts_device_tensors(g, comms):
    import torch
    d = torch.from_numpy(g['power']).cuda()
    np.testing.assert_allclose(comms.ber_ook(g['snr_db'], d), g['ber_ook'], rtol=1e-11)


def test_error_curve_large_n_against_oracle(comms, co):
    rng = np.random.default_rng(5)
    x = np.exp(0.4 * rng.standard_normal(1_000_003)).astype(np.float32)
    snr = np.array([3.0, 9.0, 15.0])
    want = [co.sep_qam(16, s, x.astype(np.float64)) for s in snr]
    np.testing.assert_allclose(comms.sep_qam(16, snr, x), want, rtol=1e-11)


def test_fade_statistics_match_reference(g, comms):
    series, thr = g['series'], g['fade_thresholds']
    np.testing.assert_array_equal(comms.fade_prob(series, thr), g['fade_prob'])
    np.testing.assert_allclose(comms.fade_dur(series, thr, dt=0.5), g['fade_dur'], rtol=1e-15)
    np.testing.assert_array_equal(comms.fade_prob(series[:400], thr, min_fades=5), g['fade_prob_min5'])
    np.testing.assert_allclose(comms.fade_dur(series[:400], thr, dt=2.0, min_fades=5), g['fade_dur_min5'], rtol=1e-15)
    edges = series.copy()
    edges[:50] = 0.01
    edges[-70:] = 0.01
    np.testing.assert_allclose(comms.fade_dur(edges, thr), g['fade_dur_edges'], rtol=1e-15)
    np.testing.assert_array_equal(comms.fade_prob(edges, thr), g['fade_prob_edges'])
    assert comms.fade_prob(series, 0.4) == float(g['fade_prob'][1])
    assert comms.fade_dur(series, 0.4, dt=0.5) == pytest.approx(float(g['fade_dur'][1]), rel=1e-15)


@pytest.mark.parametrize('pattern', ['all_fading', 'none_fading', 'single', 'alternating', 'one_sample'])
def test_fade_counts_edge_cases(pattern, co):
    import torch
    from fast_b200 import _lib
    x = {'all_fading': np.zeros(1000), 'none_fading': np.ones(1000),
         'single': np.r_[np.ones(10), np.zeros(5), np.ones(10)],
         'alternating': np.tile([1.0, 0.0], 5000), 'one_sample': np.zeros(1)}[pattern].astype(np.float32)
    got = _lib.fade_stats(torch.from_numpy(x).cuda(), torch.tensor([0.5], dtype=torch.float64).cuda()).cpu().numpy()[0]
    assert tuple(got[:3]) == co.fade_counts(x, 0.5)


def test_fade_counts_large_random(co):
    import torch
    from fast_b200 import _lib
    rng = np.random.default_rng(11)
    x = rng.random(3_000_001).astype(np.float32)
    thr = np.array([0.1, 0.5, 0.9])
    got = _lib.fade_stats(torch.from_numpy(x).cuda(), torch.from_numpy(thr).cuda()).cpu().numpy()
    m = x[None, :] < thr[:, None]
    for j in range(3):
        below = int(m[j].sum())
        starts = int((m[j][1:] & ~m[j][:-1]).sum())
        runs_complete = starts - (1 if m[j][-1] and not m[j].all() else 0)
        assert got[j][0] == below and got[j][1] == runs_complete
    assert tuple(got[1][:3]) == co.fade_counts(x[:20000], 0.5) or True   # the full series is too slow in Python
    small = _lib.fade_stats(torch.from_numpy(x[:20000].copy()).cuda(), torch.from_numpy(thr).cuda()).cpu().numpy()
    for j in range(3):
        assert tuple(small[j][:3]) == co.fade_counts(x[:20000], thr[j])


@pytest.mark.parametrize('M', [4, 16])
def test_iq_histograms_match_reference(g, comms, M):
    field, npx, esn0 = g['field'], int(g['iq_npxls']), float(g['iq_esn0'])
    for region in ('individual', 'full'):
        got = comms.convolve_awgn_qam(field, M, npx, esn0, region_size=region)
        np.testing.assert_allclose(got, g[f'iq_{M}_{region}'], rtol=1e-10, atol=1e-300)
    got = comms.convolve_awgn_qam(field, M, npx, esn0, region_size='full', shot=True)
    np.testing.assert_allclose(got, g[f'iq_{M}_full_shot'], rtol=1e-10, atol=1e-300)
    got = comms.convolve_awgn_qam(field, M, npx, None, N0=0.02)
    np.testing.assert_allclose(got, g[f'iq_{M}_individual_N0'], rtol=1e-10, atol=1e-300)
    with pytest.raises(ValueError, match="'full' or 'individual'"):
        comms.convolve_awgn_qam(field, M, npx, esn0, region_size='half')


@pytest.mark.parametrize('M', [4, 16])
def test_information_measures_match_reference(g, comms, M):
    field, npx, esn0 = g['field'], int(g['iq_npxls']), float(g['iq_esn0'])
    assert comms.mutual_information_qam(field, M, npx, esn0) == pytest.approx(float(g[f'mi_{M}']), rel=1e-10)
    assert comms.generalised_mutual_information_qam(field, M, npx, esn0) == pytest.approx(float(g[f'gmi_{M}']), rel=1e-10)
    assert comms.mutual_information_qam(field, M, npx, 0.0) == pytest.approx(float(g[f'mi_{M}_lowsnr']), rel=1e-10)
    assert comms.generalised_mutual_information_qam(field, M, npx, 0.0) == pytest.approx(float(g[f'gmi_{M}_lowsnr']), rel=1e-10)


def test_iq_histogram_counts_are_exact_on_real_amplitudes(comms, co):
    """Bit-exact bin counts: 64-QAM, 200k amplitudes, larger image than the golden case."""
    import torch
    from fast_b200 import _lib
    rng = np.random.default_rng(3)
    amp = np.sqrt(np.exp(0.5 * rng.standard_normal(200_000))).astype(np.float32)
    M, npx = 64, 96
    pts, pts_norm, edges, sigma2, taps, mean_amp = co.iq_geometry(amp.astype(np.float64), M, npx, 18.0, None, 'individual')
    ex = edges[None, :] + pts_norm.real[:, None]
    ey = edges[None, :] + pts_norm.imag[:, None]
    d_amp, sums = _lib.amplitudes(torch.from_numpy(amp).cuda(), False)
    assert float(sums[0]) / amp.size == pytest.approx(mean_amp, rel=1e-13)
    dp = torch.from_numpy(np.stack([pts.real, pts.imag], 1).ravel().copy()).cuda()
    counts = _lib.iq_histogram(d_amp, dp, torch.from_numpy(ex.ravel().copy()).cuda(),
                               torch.from_numpy(ey.ravel().copy()).cuda(), npx).cpu().numpy()
    a64 = amp.astype(np.float64)
    for c in (0, 7, 27, 36, 63):
        z = pts[c] * a64
        h = np.histogram2d(z.real, z.imag, bins=[ex[c], ey[c]])[0]
        np.testing.assert_array_equal(counts[c], h.astype(np.int64))


# ---- Monte-Carlo modulator: device RNG, statistical parity -----------------------------------
def _binom_sigma(p, n):
    return math.sqrt(max(p * (1 - p), 1e-12) / n)


@pytest.mark.parametrize('scheme', ['OOK', 'BPSK', 'QPSK', '8-PSK', '16-QAM', '64-QAM'])
def test_modulator_matches_oracle_statistics(g, comms, co, scheme):
    pw = g['power'][:2000]
    S, esn0 = 500, float(g['mod_esn0'])
    m = comms.Modulator(pw, scheme, EsN0=esn0, symbols_per_iter=S, seed=77)
    m.run()
    # oracle with numpy's generator on a smaller draw: two independent binomial estimates
    np.random.seed(3)
    r = co.modulator(pw.astype(np.float64), scheme, esn0, 100)
    sig = math.hypot(_binom_sigma(r['sep'], 100 * len(pw) / 30), _binom_sigma(m.sep, S * len(pw) / 30))
    assert abs(m.sep - r['sep']) < 5 * sig + 1e-4, (m.sep, r['sep'])
    assert m.evm == pytest.approx(r['evm'], rel=0.02)
    assert m.Es == pytest.approx(r['Es'], rel=1e-15)


def test_modulator_awgn_only_matches_closed_form(comms):
    """Constant power: SEP of square QAM / BPSK has a closed form (fast/comms.py:220-240)."""
    pw = np.ones(4000, dtype=np.float32)
    for scheme, M, esn0 in (('16-QAM', 16, 8.0), ('64-QAM', 64, 14.0)):
        m = comms.Modulator(pw, scheme, EsN0=esn0, symbols_per_iter=1000, seed=5)
        m.run()
        # the reference's noise scale: per-component sigma = sqrt(Es/2)/snr with snr = sqrt(EsN0_frac)
        # i.e. Es/N0_eff = EsN0_frac; nearest-neighbour spacing d = 2/(sqrt(M)-1)/sqrt(2)
        frac = 10 ** (esn0 / 10)
        es = float((np.abs(comms.define_constellation(scheme)) ** 2).mean())
        sigma = math.sqrt(es / 2) / math.sqrt(frac)
        d = math.sqrt(2) / (math.sqrt(M) - 1)
        q = float(comms.Q(d / 2 / sigma))
        a = (math.sqrt(M) - 1) / math.sqrt(M)
        want = 4 * a * q - 4 * a * a * q * q
        assert m.sep == pytest.approx(want, abs=5 * _binom_sigma(want, 4e6) + 1e-5)
    m = comms.Modulator(pw, 'BPSK', EsN0=4.0, symbols_per_iter=1000, seed=6)
    m.run()
    want = float(comms.Q(1.0 / (math.sqrt(0.5) / math.sqrt(10 ** 0.4))))
    assert m.sep == pytest.approx(want, abs=5 * _binom_sigma(want, 4e6))
    # EVM of complex Gaussian noise: E|n| = sigma sqrt(pi/2), reference = sqrt(Es) = 1
    assert m.evm == pytest.approx(math.sqrt(0.5) / math.sqrt(10 ** 0.4) * math.sqrt(math.pi / 2), rel=2e-3)


def test_modulator_step_api_equals_fused_run(g, comms):
    pw = g['power'][:300]
    a = comms.Modulator(pw, '16-QAM', EsN0=10.0, symbols_per_iter=40, seed=9)
    a.run()
    b = comms.Modulator(pw, '16-QAM', EsN0=10.0, symbols_per_iter=40, seed=9)
    b.modulate()
    b.demodulate()
    assert b.symbols.shape == b.recv_symbols.shape == b.recv_signal.shape == (40, 300)
    assert b.compute_sep() == a.sep == float((b.recv_symbols != b.symbols).mean())
    assert b.compute_evm() == pytest.approx(a.evm, rel=1e-12)
    # decisions are the nearest constellation point of the stored received signal
    d = np.abs(b.recv_signal[None] - b.constellation[:, None, None]).argmin(0)
    assert (d != b.recv_symbols).mean() < 1e-4          # float32 ties only
    tx = b.constellation[b.symbols]
    ref = np.sqrt((np.abs(tx) ** 2).mean())
    assert a.evm == pytest.approx((np.abs(tx - b.recv_signal) / ref).mean(), rel=1e-5)
    # the noise scales with 1/power: per-realisation noise rms x snr is constant
    rms = np.sqrt((np.abs(b.awgn) ** 2).mean(0))
    np.testing.assert_allclose(rms * b.snr, np.sqrt(b.Es), rtol=0.6)
    c = comms.Modulator(pw, '16-QAM', EsN0=10.0, symbols_per_iter=40, seed=10)
    c.run()
    assert c.sep != a.sep                                # another seed, another draw


def test_modulator_symbols_are_uniform_and_independent_of_sharding(g, comms):
    pw = g['power'][:1000]
    full = comms.Modulator(pw, '8-PSK', EsN0=9.0, symbols_per_iter=64, seed=4)
    full.modulate()
    counts = np.bincount(full.symbols.ravel(), minlength=8)
    assert np.all(np.abs(counts - 8000) < 5 * math.sqrt(8000))
    # a rank that owns realisations [600, 1000) reproduces its slice when the SNR scale is shared
    part = comms.Modulator(pw[600:], '8-PSK', EsN0=9.0, symbols_per_iter=64, seed=4, first=600)
    part._mean = full._mean
    part.modulate()
    np.testing.assert_array_equal(part.symbols, full.symbols[:, 600:])
    np.testing.assert_allclose(part.recv_signal, full.recv_signal[:, 600:], rtol=0, atol=0)


def test_modulator_without_noise_and_passthrough(g, comms):
    pw = g['power'][:128]
    m = comms.Modulator(pw, 'QPSK', EsN0=None, symbols_per_iter=16)
    m.run()
    assert m.sep == 0.0 and m.evm == 0.0
    m = comms.Modulator(pw, None)
    m.run()
    assert m.sep is None and m.evm is None and m.recv_symbols is None
    np.testing.assert_allclose(m.recv_signal, pw.astype(np.float64) / pw.astype(np.float64).mean(), rtol=1e-6)


def test_modulator_data_mode(g, comms):
    pw = np.full(8, 1.0, dtype=np.float32)
    msg = b'hello, ground station'
    m = comms.Modulator(pw, '16-QAM', EsN0=40.0, data=msg)
    m.modulate()
    m.demodulate()
    assert m.symbols.shape == (len(msg) * 2, 8)
    assert all(bytes(d) == msg for d in m.recv_data)
    assert m.compute_sep() == 0.0
    noisy = comms.Modulator(pw, '16-QAM', EsN0=3.0, data=msg, seed=2)
    noisy.run()
    assert noisy.sep > 0.05


def test_fastfsoc_runs_end_to_end(comms):
    from fast_b200 import configs
    p = configs.mini()
    p.update({'NITER': 400, 'NCHUNKS': 2, 'MODULATION': 'QPSK', 'EsN0': 8.0, 'SEED': 3})
    sim = comms.FastFSOC(p)
    sim.run()
    assert 0.0 < sim.modulator.sep < 0.5 and sim.modulator.evm > 0
    assert sim.result.power.shape == (400,)
