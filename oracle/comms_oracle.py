"""CPU ORACLE for the link metrics (TEST INFRASTRUCTURE -- NOT PRODUCT CODE).

A float64 numpy restatement of the consumers of FAST's per-realisation output
(/root/reference/fast/comms.py): closed-form error curves averaged over the samples, fade
statistics, the Monte-Carlo modulator, the AWGN-convolved I-Q histograms and the (generalised)
mutual information.  Written as plain functions over arrays, independently of the reference's
code, so that each device kernel of fast_b200/csrc/link_metrics.cu has one function to be
compared with.

Who may import this: tests/ only.  The product package `fast_b200` never imports it.

Parity status: PINNED.  The reference's tests hold no golden vectors for this module, so the pin
is the reference itself run here: oracle/make_golden_comms.py imports the unmodified
fast/comms.py through oracle/shim and commits inputs and answers as tests/golden/comms.npz;
tests/test_oracle_vs_golden.py checks every function below against it (modulator: bit for bit,
with the reference's numpy.random draws replayed from the same seed).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erfc as _erfc


def q_function(x):
    """Gaussian tail probability (fast/comms.py:255-259)."""
    return 0.5 * _erfc(np.asarray(x, dtype=np.float64) / math.sqrt(2.0))


def ber_ook(ebn0_db, samples=None):
    """Mean of Q(s * sqrt(Eb/N0)) over mean-normalised samples s (fast/comms.py:193-217)."""
    snr = math.sqrt(10.0 ** (ebn0_db / 10.0))
    if samples is None:
        return float(q_function(snr))
    s = np.asarray(samples, dtype=np.float64)
    return float(q_function(s / s.mean() * snr).mean())


def _sep_qam_formula(M, esn0_frac):
    a = (math.sqrt(M) - 1.0) / math.sqrt(M)
    q = q_function(np.sqrt(3.0 / (M - 1.0) * esn0_frac))
    return 4.0 * (a * q - a * a * q * q)


def sep_qam(M, esn0_db, samples=None):
    """Square M-QAM symbol error probability, Es/N0 scaled by s^2 per sample (fast/comms.py:220-240)."""
    frac = 10.0 ** (esn0_db / 10.0)
    if samples is None:
        return float(_sep_qam_formula(M, frac))
    s = np.asarray(samples, dtype=np.float64)
    s = s / s.mean()
    return float(_sep_qam_formula(M, frac * s * s).mean())


def ber_qam(M, ebn0_db, samples=None):
    """One bit error per symbol error, Es = log2(M) Eb (fast/comms.py:243-253)."""
    bits = math.log2(M)
    return sep_qam(M, 10.0 * math.log10(bits) + ebn0_db, samples) / bits


def fade_counts(series, threshold):
    """(samples below threshold, complete fades, samples inside complete fades).

    A complete fade is a maximal run of below-threshold samples that starts at index >= 1 (a
    run already in progress at index 0 has no start) and ends before the last sample
    (fast/comms.py:180-187: splits at the 0->1 transitions, drops the chunk before the first
    one, keeps chunks whose last element is not fading)."""
    m = np.asarray(series) < threshold
    below = int(m.sum())
    n = len(m)
    fades = 0
    inside = 0
    i = 0
    while i < n:
        if m[i]:
            j = i
            while j < n and m[j]:
                j += 1
            if i >= 1 and j < n:
                fades += 1
                inside += j - i
            i = j
        else:
            i += 1
    return below, fades, inside


def fade_prob(series, threshold, min_fades=30):
    """Fraction of samples below threshold, NaN when fewer than min_fades samples (fast/comms.py:171-177)."""
    below, _, _ = fade_counts(series, threshold)
    return float('nan') if below < min_fades else below / len(series)


def fade_dur(series, threshold, dt=1, min_fades=30):
    """Mean length of the complete fades times dt, NaN when fewer than min_fades (fast/comms.py:180-191)."""
    _, fades, inside = fade_counts(series, threshold)
    return float('nan') if fades < min_fades else inside / fades * dt


def n_symbols(scheme):
    """Alphabet size per scheme name (fast/comms.py:38-50)."""
    if scheme in ('OOK', 'BPSK'):
        return 2
    if scheme in ('QPSK', 'QAM'):
        return 4
    parts = scheme.split('-')
    if len(parts) == 2:
        return int(parts[0])
    raise ValueError('Scheme not recognised')


def constellation(scheme):
    """Constellation points per scheme (fast/comms.py:417-470)."""
    if scheme == 'OOK':
        return np.array([0, 1])
    if scheme == 'BPSK':
        return np.exp(1j * np.pi * np.arange(2))
    if scheme in ('QPSK', 'QAM'):
        return np.exp(1j * (np.arange(4) * np.pi / 2 - np.pi / 4))
    if scheme.endswith('-PSK'):
        m = int(scheme[:-4])
        return np.exp(1j * (np.arange(m) * np.pi / (m / 2)))
    if scheme.endswith('-QAM'):
        m = int(scheme[:-4])
        side = int(round(math.sqrt(m)))
        if side * side != m:
            raise ValueError(f'{m}-QAM is not a square constellation')
        axis = np.linspace(-1, 1, side) / np.sqrt(2)
        xx, yy = np.meshgrid(axis, axis)
        return (xx + 1j * yy).flatten()
    raise ValueError(f'Modulation scheme {scheme} not supported')


def gray_map_qam(M):
    """Gray code of symbol index c, rows of the square alternately reversed (fast/comms.py:473-496).
    Returned as integers (bit i of the reference's string, from the left, is bit m-1-i)."""
    side = int(round(math.sqrt(M)))
    idx = np.arange(M)
    g = (idx ^ (idx >> 1)).reshape(side, side).copy()
    g[1::2] = g[1::2, ::-1]
    return g.flatten()


def modulator(power, scheme, esn0_db, symbols_per_iter, rng=np.random):
    """Monte-Carlo modulate / add AWGN / demodulate (fast/comms.py:13-146), drawing from `rng` in
    the reference's order: symbols, then the real and (coherent schemes) imaginary noise blocks."""
    p = np.asarray(power, dtype=np.float64)
    p = p / p.mean()
    n = len(p)
    pts = constellation(scheme)
    symbols = rng.randint(0, n_symbols(scheme), size=(symbols_per_iter, n))
    tx = pts[symbols]
    es = float((np.abs(pts) ** 2).mean())
    if esn0_db is None:
        noise = 0
    else:
        snr = math.sqrt(10.0 ** (esn0_db / 10.0)) * p
        if scheme == 'OOK':
            noise = rng.normal(0, es / snr, size=(symbols_per_iter, n))
        else:
            sd = math.sqrt(es / 2.0) / snr
            noise = rng.normal(0, sd, size=(symbols_per_iter, n)) + 1j * rng.normal(0, sd, size=(symbols_per_iter, n))
    rx = tx + noise
    if scheme == 'OOK':
        decided = (rx > 0.5).astype(int)
    elif scheme == 'BPSK':
        decided = (rx.real < 0).astype(int)
    else:
        decided = np.abs(rx[None] - pts[:, None, None]).argmin(0)
    ref = math.sqrt(float((tx.real ** 2 + tx.imag ** 2).mean()))
    return {'symbols': symbols, 'awgn': noise, 'recv_symbols': decided, 'Es': es,
            'sep': float((decided != symbols).mean()), 'evm': float((np.abs(tx - rx) / ref).mean())}


def _histogram_bins(values, edges):
    """numpy.histogramdd's bin rule: right-open bins, the last edge inclusive; -1 = outside."""
    idx = np.searchsorted(edges, values, side='right')
    idx[values == edges[-1]] -= 1
    idx = idx - 1
    idx[(idx < 0) | (idx >= len(edges) - 1)] = -1
    return idx


def _correlate_matrix(g, n):
    """G[i, j] = g[j - i + len(g)//2]: scipy.ndimage.correlate1d with mode='constant', cval=0."""
    k = len(g)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    t = j - i + k // 2
    ok = (t >= 0) & (t < k)
    return np.where(ok, g[np.clip(t, 0, k - 1)], 0.0)


def iq_geometry(samples, M, npxls, esn0_db, N0=None, region='individual'):
    """Region width, bin edges (before the per-symbol offset), AWGN variance in pixel units and
    the Gaussian taps (fast/comms.py:346-378)."""
    pts = constellation(f'{M}-QAM')
    mean_amp = float(np.mean(np.abs(samples)))
    if region == 'individual':
        width = 1.0 / (math.sqrt(M) - 1.0)
    elif region == 'full':
        width = 2.0
    else:
        raise ValueError("decision_region_size must be either 'full' or 'individual'")
    pts_norm = pts * mean_amp
    width *= mean_amp
    if N0 is None:
        N0 = float(np.mean(np.abs(pts_norm) ** 2)) / 10.0 ** (esn0_db / 10.0)
    if region == 'full':
        need = 2.0 * (mean_amp / math.sqrt(2.0) + 2.0 * math.sqrt(N0))
        width = max(width, need)
    dx = width / npxls
    sigma2 = max(N0 / (2.0 * dx * dx), 1.0)
    taps_x = np.linspace(-npxls / 2, npxls / 2, npxls + 1)
    taps = np.exp(-taps_x ** 2 / sigma2) / math.sqrt(math.pi * sigma2)
    edges = np.linspace(-width / 2, width / 2, npxls + 1)
    return pts, pts_norm, edges, sigma2, taps, mean_amp


def iq_histograms(samples, M, npxls, esn0_db, N0=None, region='individual', shot=False):
    """Per transmitted symbol: 2-D histogram of c*|sample| over the decision region, convolved
    with the AWGN Gaussian (fast/comms.py:306-414).  Returns (M, npxls, npxls)."""
    samples = np.asarray(samples)
    pts, pts_norm, edges, sigma2, taps, mean_amp = iq_geometry(samples, M, npxls, esn0_db, N0, region)
    amp = np.abs(samples).astype(np.float64)
    G = _correlate_matrix(taps, npxls)
    out = np.zeros((len(pts), npxls, npxls))
    for c, pt in enumerate(pts):
        ex, ey = edges.copy(), edges.copy()
        if region == 'individual':
            ex = ex + pts_norm[c].real
            ey = ey + pts_norm[c].imag
        z = pt * amp
        bx, by = _histogram_bins(z.real, ex), _histogram_bins(z.imag, ey)
        ok = (bx >= 0) & (by >= 0)
        h = np.zeros((npxls, npxls))
        np.add.at(h, (bx[ok], by[ok]), 1.0)
        h /= len(amp)
        if not shot:
            out[c] = G @ h @ G.T
        else:
            # signal-dependent noise: every occupied bin spreads as its own Gaussian whose
            # variance grows with mean_amp^2 / |bin position|^2 (fast/comms.py:399-408)
            ii, jj = np.nonzero(h)
            yy, xx = np.meshgrid(np.arange(npxls), np.arange(npxls), indexing='ij')
            acc = np.zeros((npxls, npxls))
            for i, j in zip(ii, jj):
                mult = mean_amp ** 2 / (ex[i] ** 2 + ey[j] ** 2)
                w = math.sqrt(sigma2 * mult / 2.0)
                blob = np.exp(-(((j - xx) / w) ** 2 + ((i - yy) / w) ** 2) / 2.0)
                acc += h[i, j] * blob / (math.pi * sigma2 * mult)
            out[c] = acc
    return out


def _xlog_ratio(f, fy):
    """f * (log2 f - log2 fy) with the masked-array rule of the reference: entries where f <= 0 or
    fy <= 0 drop out (numpy.ma.log2 masks them; a masked sum skips them, a masked product stored
    into a plain array leaves f = 0 there)."""
    ok = (f > 0) & (fy > 0)
    out = np.zeros_like(f)
    out[ok] = f[ok] * (np.log2(f[ok]) - np.log2(np.broadcast_to(fy, f.shape)[ok]))
    return out


def mutual_information_qam(samples, M, npxls, esn0_db, N0=None, shot=False):
    """Memoryless-receiver mutual information in bits/symbol (fast/comms.py:293-303)."""
    fyx = iq_histograms(samples, M, npxls, esn0_db, N0=N0, region='full', shot=shot)
    fy = fyx.mean(0)
    return float(_xlog_ratio(fyx, fy[None]).sum((-1, -2)).mean())


def generalised_mutual_information_qam(samples, M, npxls, esn0_db, N0=None, shot=False):
    """Bit-wise decoder GMI with the Gray map above (fast/comms.py:262-290)."""
    fyx = iq_histograms(samples, M, npxls, esn0_db, N0=N0, region='full', shot=shot)
    fy = fyx.mean(0)
    gray = gray_map_qam(M)
    m = int(round(math.log2(M)))
    total = 0.0
    for i in range(m):
        zero = ((gray >> (m - 1 - i)) & 1) == 0
        f0, f1 = fyx[zero].mean(0), fyx[~zero].mean(0)
        total += 0.5 * (_xlog_ratio(f0, fy).sum() + _xlog_ratio(f1, fy).sum())
    return float(total)
