"""Pins the link-metrics oracle (oracle/comms_oracle.py) to outputs of the unmodified reference
fast/comms.py (tests/golden/comms.npz, written by oracle/make_golden_comms.py).  CPU only."""
import numpy as np
import pytest

from oracle import comms_oracle as co
from conftest import load_golden_arrays


@pytest.fixture(scope='module')
def g():
    return load_golden_arrays('comms')


def test_error_curves(g):
    power = g['power'].astype(np.float64)
    snr = g['snr_db']
    np.testing.assert_allclose([co.ber_ook(s, power) for s in snr], g['ber_ook'], rtol=1e-13)
    np.testing.assert_allclose([co.ber_ook(s) for s in snr], g['ber_ook_noatm'], rtol=1e-13)
    for M in g['qam_orders']:
        M = int(M)
        np.testing.assert_allclose([co.sep_qam(M, s, power) for s in snr], g[f'sep_qam_{M}'], rtol=1e-13)
        np.testing.assert_allclose([co.sep_qam(M, s) for s in snr], g[f'sep_qam_{M}_noatm'], rtol=1e-13)
        np.testing.assert_allclose([co.ber_qam(M, s, power) for s in snr], g[f'ber_qam_{M}'], rtol=1e-13)


def test_fade_statistics(g):
    series = g['series'].astype(np.float64)
    thr = g['fade_thresholds']
    np.testing.assert_array_equal([co.fade_prob(series, t) for t in thr], g['fade_prob'])
    np.testing.assert_allclose([co.fade_dur(series, t, dt=0.5) for t in thr], g['fade_dur'], rtol=1e-15)
    np.testing.assert_array_equal([co.fade_prob(series[:400], t, min_fades=5) for t in thr], g['fade_prob_min5'])
    np.testing.assert_allclose([co.fade_dur(series[:400], t, dt=2.0, min_fades=5) for t in thr],
                               g['fade_dur_min5'], rtol=1e-15)
    edges = series.copy()
    edges[:50] = 0.01
    edges[-70:] = 0.01          # a fade in progress at both ends: neither is a complete fade
    np.testing.assert_allclose([co.fade_dur(edges, t) for t in thr], g['fade_dur_edges'], rtol=1e-15)
    np.testing.assert_array_equal([co.fade_prob(edges, t) for t in thr], g['fade_prob_edges'])


def test_constellations_and_gray_maps(g):
    for s in list(g['mod_schemes']) + ['16-PSK', '4-QAM', '256-QAM']:
        np.testing.assert_allclose(co.constellation(str(s)), g[f'constellation_{s}'], rtol=0, atol=1e-15)
    for M in (4, 16, 64):
        np.testing.assert_array_equal(co.gray_map_qam(M), g[f'gray_{M}'])
    with pytest.raises(ValueError):
        co.constellation('8-QAM')
    with pytest.raises(ValueError):
        co.constellation('FSK')


def test_modulator_replays_the_reference_draws(g):
    pw = g['power'][:300].astype(np.float64)
    for s in g['mod_schemes']:
        s = str(s)
        np.random.seed(int(g['mod_seed']))
        r = co.modulator(pw, s, float(g['mod_esn0']), int(g['mod_symbols']))
        np.testing.assert_array_equal(r['symbols'], g[f'mod_{s}_symbols'])
        np.testing.assert_array_equal(r['awgn'], g[f'mod_{s}_awgn'])
        np.testing.assert_array_equal(r['recv_symbols'], g[f'mod_{s}_recv_symbols'])
        assert r['sep'] == float(g[f'mod_{s}_sep'])
        assert r['evm'] == pytest.approx(float(g[f'mod_{s}_evm']), rel=1e-14)
        assert r['Es'] == pytest.approx(float(g[f'mod_{s}_Es']), rel=1e-15)
    np.random.seed(int(g['mod_seed']))
    r = co.modulator(pw, 'QPSK', None, int(g['mod_symbols']))
    assert r['sep'] == float(g['mod_QPSK_nonoise_sep']) == 0.0
    assert r['evm'] == float(g['mod_QPSK_nonoise_evm']) == 0.0


@pytest.mark.parametrize('M', [4, 16])
def test_iq_histograms_and_information(g, M):
    field = g['field'].astype(np.complex128)
    npx, esn0 = int(g['iq_npxls']), float(g['iq_esn0'])
    for region in ('individual', 'full'):
        got = co.iq_histograms(field, M, npx, esn0, region=region)
        np.testing.assert_allclose(got, g[f'iq_{M}_{region}'], rtol=1e-12, atol=1e-300)
    got = co.iq_histograms(field, M, npx, esn0, region='full', shot=True)
    np.testing.assert_allclose(got, g[f'iq_{M}_full_shot'], rtol=1e-11, atol=1e-300)
    got = co.iq_histograms(field, M, npx, None, N0=0.02)
    np.testing.assert_allclose(got, g[f'iq_{M}_individual_N0'], rtol=1e-12, atol=1e-300)
    assert co.mutual_information_qam(field, M, npx, esn0) == pytest.approx(float(g[f'mi_{M}']), rel=1e-12)
    assert co.generalised_mutual_information_qam(field, M, npx, esn0) == pytest.approx(float(g[f'gmi_{M}']), rel=1e-12)
    assert co.mutual_information_qam(field, M, npx, 0.0) == pytest.approx(float(g[f'mi_{M}_lowsnr']), rel=1e-12)
    assert co.generalised_mutual_information_qam(field, M, npx, 0.0) == pytest.approx(float(g[f'gmi_{M}_lowsnr']), rel=1e-12)
