"""Generate tests/golden/comms.npz by running the UNMODIFIED reference fast/comms.py (through
oracle/shim) on small seeded inputs.  TEST INFRASTRUCTURE.

    python oracle/make_golden_comms.py

The inputs are synthetic per-realisation outputs (float32 powers, as the device produces them,
widened to float64 for the reference; a temporally correlated series for the fade statistics;
complex fields for the I-Q histograms) and are stored next to the reference's answers, so the
tests never need /root/reference.  The Monte-Carlo modulator uses numpy's legacy global RNG
(fast/comms.py:58,76-79): it is seeded here with numpy.random.seed and the draws themselves are
stored, so that the oracle restatement can be checked bit for bit.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'shim'))

import numpy as np  # noqa: E402

from fast import comms  # noqa: E402  (the reference)

OUT = os.path.join(ROOT, 'tests', 'golden', 'comms.npz')

SNR_DB = np.arange(0.0, 22.0, 2.0)
QAM_ORDERS = (4, 16, 64)
MOD_SCHEMES = ('OOK', 'BPSK', 'QPSK', 'QAM', '8-PSK', '16-QAM', '64-QAM')
MOD_SEED, MOD_ESN0, MOD_SYMBOLS = 20240, 12.0, 40
IQ_NPXLS, IQ_ESN0 = 32, 14.0


def inputs():
    rng = np.random.default_rng(777)
    power = np.exp(0.35 * rng.standard_normal(4096) - 0.06).astype(np.float32)
    # AR(1) log-normal series: long fades, > 30 of them at the thresholds below
    n = 30000
    w = rng.standard_normal(n)
    x = np.empty(n)
    x[0] = w[0]
    for i in range(1, n):
        x[i] = 0.97 * x[i - 1] + np.sqrt(1 - 0.97 ** 2) * w[i]
    series = np.exp(0.5 * x - 0.125).astype(np.float32)
    amp = np.sqrt(power[:2048].astype(np.float64))
    field = (amp * np.exp(1j * 0.2 * rng.standard_normal(2048))).astype(np.complex64)
    return power, series, field


def main():
    power32, series32, field64 = inputs()
    power = power32.astype(np.float64)
    series = series32.astype(np.float64)
    field = field64.astype(np.complex128)
    out = {'power': power32, 'series': series32, 'field': field64, 'snr_db': SNR_DB,
           'qam_orders': np.array(QAM_ORDERS), 'mod_schemes': np.array(MOD_SCHEMES),
           'mod_seed': MOD_SEED, 'mod_esn0': MOD_ESN0, 'mod_symbols': MOD_SYMBOLS,
           'iq_npxls': IQ_NPXLS, 'iq_esn0': IQ_ESN0}

    # closed-form error curves averaged over the samples (fast/comms.py:193-258)
    out['ber_ook'] = np.array([comms.ber_ook(s, power) for s in SNR_DB])
    out['ber_ook_noatm'] = np.array([comms.ber_ook(s) for s in SNR_DB])
    for M in QAM_ORDERS:
        out[f'sep_qam_{M}'] = np.array([comms.sep_qam(M, s, power) for s in SNR_DB])
        out[f'sep_qam_{M}_noatm'] = np.array([comms.sep_qam(M, s) for s in SNR_DB])
        out[f'ber_qam_{M}'] = np.array([comms.ber_qam(M, s, power) for s in SNR_DB])

    # fade statistics (fast/comms.py:171-191)
    thr = np.array([0.2, 0.4, 0.7, 1.0, 5.0, 0.001])
    out['fade_thresholds'] = thr
    out['fade_prob'] = np.array([comms.fade_prob(series, t) for t in thr])
    out['fade_dur'] = np.array([comms.fade_dur(series, t, dt=0.5) for t in thr])
    out['fade_prob_min5'] = np.array([comms.fade_prob(series[:400], t, min_fades=5) for t in thr])
    out['fade_dur_min5'] = np.array([comms.fade_dur(series[:400], t, dt=2.0, min_fades=5) for t in thr])
    # edge cases: series that starts / ends inside a fade
    lead = series.copy()
    lead[:50] = 0.01
    lead[-70:] = 0.01
    out['fade_dur_edges'] = np.array([comms.fade_dur(lead, t) for t in thr])
    out['fade_prob_edges'] = np.array([comms.fade_prob(lead, t) for t in thr])

    # constellations and Gray maps (fast/comms.py:417-506)
    for s in MOD_SCHEMES + ('16-PSK', '4-QAM', '256-QAM'):
        out[f'constellation_{s}'] = np.asarray(comms.define_constellation(s), dtype=np.complex128)
    for M in (4, 16, 64):
        out[f'gray_{M}'] = np.array([int(c, 2) for c in comms._bin2gray_qam(M)])

    # Monte-Carlo modulator (fast/comms.py:13-146): legacy global RNG, seeded; draws stored
    pw = power[:300]
    for s in MOD_SCHEMES:
        np.random.seed(MOD_SEED)
        m = comms.Modulator(pw, s, EsN0=MOD_ESN0, symbols_per_iter=MOD_SYMBOLS)
        m.run()
        out[f'mod_{s}_symbols'] = m.symbols.astype(np.int16)
        out[f'mod_{s}_awgn'] = np.asarray(m.awgn)
        out[f'mod_{s}_recv_symbols'] = np.asarray(m.recv_symbols).astype(np.int16)
        out[f'mod_{s}_sep'] = m.sep
        out[f'mod_{s}_evm'] = m.evm
        out[f'mod_{s}_Es'] = m.Es
    np.random.seed(MOD_SEED)
    m = comms.Modulator(pw, 'QPSK', EsN0=None, symbols_per_iter=MOD_SYMBOLS)
    m.run()
    out['mod_QPSK_nonoise_sep'] = m.sep
    out['mod_QPSK_nonoise_evm'] = m.evm

    # I-Q plane histograms with AWGN and the information measures (fast/comms.py:260-415)
    for M in (4, 16):
        for region in ('individual', 'full'):
            out[f'iq_{M}_{region}'] = comms.convolve_awgn_qam(field, M, IQ_NPXLS, IQ_ESN0, region_size=region)
        out[f'iq_{M}_full_shot'] = comms.convolve_awgn_qam(field, M, IQ_NPXLS, IQ_ESN0, region_size='full', shot=True)
        out[f'iq_{M}_individual_N0'] = comms.convolve_awgn_qam(field, M, IQ_NPXLS, None, N0=0.02)
        out[f'mi_{M}'] = comms.mutual_information_qam(field, M, IQ_NPXLS, IQ_ESN0)
        out[f'gmi_{M}'] = comms.generalised_mutual_information_qam(field, M, IQ_NPXLS, IQ_ESN0)
        out[f'mi_{M}_lowsnr'] = comms.mutual_information_qam(field, M, IQ_NPXLS, 0.0)
        out[f'gmi_{M}_lowsnr'] = comms.generalised_mutual_information_qam(field, M, IQ_NPXLS, 0.0)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KiB')


if __name__ == '__main__':
    main()
