"""Optical-communications metrics on FAST's per-realisation output: host mirror of the reference's
`fast/comms.py` (same names, arguments and return values) over the K5 kernels of libfastb
(include/fastb.h, fast_b200/csrc/link_metrics.cu).

Every function that walks the sample array runs on the GPU: inputs may be numpy arrays (copied
to the device as float32 -- the precision FAST's results are produced in) or torch CUDA tensors
(used in place, e.g. the device-resident output of `Fast.run`).  There is no CPU fallback.
Only scalar closed forms (no samples), constellation tables and the byte/bit packing helpers are
plain host code.
"""
import logging
import math

import numpy
import torch

from . import _lib
from .fast import Fast

logger = logging.getLogger(__name__)


# ---- device plumbing ------------------------------------------------------------------------
def _device():
    _lib.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def _real_samples(x):
    """float32 CUDA vector of real samples (powers, amplitudes)."""
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.to(_device())
        if t.is_complex():
            raise TypeError('real samples expected')
        return t.to(torch.float32).contiguous().view(-1)
    a = numpy.asarray(x)
    if numpy.iscomplexobj(a):
        raise TypeError('real samples expected')
    return torch.from_numpy(numpy.ascontiguousarray(a, dtype=numpy.float32).ravel()).to(_device())


def _field_samples(x):
    """(float32 CUDA buffer, is_complex, n) of field measurements or amplitudes."""
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.to(_device())
        if t.is_complex():
            t = torch.view_as_real(t.to(torch.complex64).contiguous().view(-1)).contiguous().view(-1)
            return t, True, t.numel() // 2
        t = t.to(torch.float32).contiguous().view(-1)
        return t, False, t.numel()
    a = numpy.asarray(x).ravel()
    if numpy.iscomplexobj(a):
        t = torch.from_numpy(numpy.ascontiguousarray(a, dtype=numpy.complex64).view(numpy.float32)).to(_device())
        return t, True, a.size
    return torch.from_numpy(numpy.ascontiguousarray(a, dtype=numpy.float32)).to(_device()), False, a.size


def _f64(values, dev):
    return torch.as_tensor(numpy.atleast_1d(numpy.asarray(values, dtype=numpy.float64)), device=dev)


# ---- closed-form error rates (fast/comms.py:193-258) -------------------------------------------
def Q(x):
    """Gaussian tail probability 1/2 erfc(x / sqrt 2)."""
    from scipy.special import erfc
    return 1 / 2 * erfc(x / numpy.sqrt(2))


def _curve(kind, M, snr_db, samples):
    d = _real_samples(samples)
    out = _lib.error_curve(d, kind, M, _f64(snr_db, d.device)).cpu().numpy()[:-1]
    return out if numpy.ndim(snr_db) else float(out[0])


def ber_ook(EbN0, samples=None):
    """Bit error rate of on-off keying at electrical Eb/N0 [dB], averaged over the mean-normalised
    received-power samples (no samples: no atmosphere).  EbN0 may be an array: one launch
    evaluates the whole curve."""
    if samples is None:
        return Q(numpy.sqrt(10 ** (numpy.asarray(EbN0, dtype=float) / 10)))[()]
    return _curve(_lib.CURVE_BER_OOK, 0, EbN0, samples)


def sep_qam(M, EsN0, samples=None):
    """Symbol error probability of square M-QAM at electrical Es/N0 [dB]."""
    if samples is None:
        a = (numpy.sqrt(M) - 1) / numpy.sqrt(M)
        q = Q(numpy.sqrt(3 / (M - 1) * 10 ** (numpy.asarray(EsN0, dtype=float) / 10)))
        return (4 * (a * q - a ** 2 * q ** 2))[()]
    return _curve(_lib.CURVE_SEP_QAM, int(M), EsN0, samples)


def ber_qam(M, EbN0, samples=None):
    """Bit error rate of Gray-coded square M-QAM: one bit error per symbol error."""
    return 1 / numpy.log2(M) * sep_qam(M, 10 * numpy.log10(numpy.log2(M)) + numpy.asarray(EbN0, dtype=float)[()], samples)


# ---- fade statistics (fast/comms.py:171-191) -----------------------------------------------------
def _fade_counts(I, threshold):
    d = _real_samples(I)
    return _lib.fade_stats(d, _f64(threshold, d.device)).cpu().numpy(), d.numel()


def fade_prob(I, threshold, min_fades=30):
    """Fraction of samples below `threshold`; NaN when fewer than `min_fades` samples are."""
    c, n = _fade_counts(I, threshold)
    p = numpy.where(c[:, 0] < min_fades, numpy.nan, c[:, 0] / n)
    return p if numpy.ndim(threshold) else float(p[0])


def fade_dur(I, threshold, dt=1, min_fades=30):
    """Mean duration of the fades that start and end inside the window; NaN when fewer than
    `min_fades` of them."""
    c, _ = _fade_counts(I, threshold)
    with numpy.errstate(divide='ignore', invalid='ignore'):
        d = numpy.where(c[:, 1] < min_fades, numpy.nan, c[:, 2] / c[:, 1] * dt)
    return d if numpy.ndim(threshold) else float(d[0])


# ---- constellations (fast/comms.py:417-506) ----------------------------------------------------
def define_constellation(modulation):
    """Complex constellation of OOK, BPSK, QPSK / QAM, M-PSK and square M-QAM."""
    if modulation == 'OOK':
        return numpy.array([0, 1])
    if modulation == 'BPSK':
        return numpy.exp(1j * numpy.arange(2) * numpy.pi)
    if modulation in ('QPSK', 'QAM'):
        return numpy.exp(1j * (numpy.arange(4) * numpy.pi / 2 - numpy.pi / 4))
    if modulation[-4:] == '-PSK':
        n = int(modulation[:-4])
        return numpy.exp(1j * (numpy.arange(n) * numpy.pi / (n / 2)))
    if modulation[-4:] == '-QAM':
        n = int(modulation[:-4])
        side = int(round(math.sqrt(n)))
        if side * side != n:
            raise ValueError(f"{n}-QAM not possible as {n} is not a perfect square, only square M-QAM "
                             "modulations supported")
        axis = numpy.linspace(-1, 1, side) / numpy.sqrt(2)
        xx, yy = numpy.meshgrid(axis, axis)
        return (xx + 1j * yy).flatten()
    raise ValueError(f"Modulation scheme {modulation} not supported")


def _gray_ints(M):
    side = int(round(math.sqrt(M)))
    idx = numpy.arange(M)
    g = (idx ^ (idx >> 1)).reshape(side, side).copy()
    g[1::2] = g[1::2, ::-1].copy()
    return g.flatten()


def _bin2gray_qam(M):
    """Gray code (bit strings) of the M-QAM symbols, every other constellation row reversed."""
    m = int(numpy.log2(M))
    return numpy.array([format(int(v), 'b').zfill(m) for v in _gray_ints(M)])


def _bit_at_index(code, index, bit):
    return numpy.array([c[index] == str(bit) for c in code], dtype=bool)


def _n_symbols(modulation):
    if modulation in ('OOK', 'BPSK'):
        return 2
    if modulation in ('QPSK', 'QAM'):
        return 4
    if len(modulation.split('-')) == 2:
        return int(modulation.split('-')[0])
    raise ValueError("Scheme not recognised")


# ---- Monte-Carlo modulator (fast/comms.py:13-146) ----------------------------------------------
class Modulator():
    '''
    Takes an array of optical powers and modulates / demodulates random symbols of a modulation
    scheme (OOK, BPSK, QPSK, M-PSK, M-QAM) with AWGN at an average symbol signal-to-noise ratio,
    for Monte-Carlo symbol error probability and error vector magnitude.

    Parameters:
        power (numpy.ndarray or CUDA tensor): array of optical powers
        modulation (string): modulation scheme (None: pass the powers through)
        EsN0 (float, optional): (average) symbol signal to noise ratio [dB]
        symbols_per_iter (int, optional): symbols per iteration of FAST. Defaults to 1000.
        data (bytes, optional): transmit these bytes (the same for every iteration) instead of
            random symbols
        seed (int, optional): Philox key of the device generator (default 0).  The reference draws
            from numpy's unseeded global generator, so results agree statistically.

    `run()` is one fused kernel (draw, add noise, decide, accumulate): nothing of size
    symbols_per_iter x len(power) is stored.  `modulate()` / `demodulate()` materialise
    `symbols`, `awgn`, `recv_signal`, `recv_symbols` from the same random stream.
    '''

    def __init__(self, power, modulation, EsN0=None, symbols_per_iter=1000, data=None, seed=0, first=0):
        self._d_power = _real_samples(power)
        amp, sums = _lib.amplitudes(self._d_power, False)
        self._mean = float(sums[0].item()) / self._d_power.numel()
        self.modulation = modulation
        self.symbols_per_iter = symbols_per_iter
        self.EsN0 = EsN0
        self.data = data
        self.seed = seed
        self.first = first
        self._sums = None

    power = property(lambda self: self._d_power.cpu().numpy().astype(numpy.float64) / self._mean)
    amplitude = property(lambda self: numpy.sqrt(self.power))

    @property
    def snr(self):
        if self.EsN0 is None:
            raise AttributeError('snr is defined only when EsN0 is given')
        return numpy.sqrt(10 ** (self.EsN0 / 10)) * self.power

    def generate_symbols(self):
        """Alphabet size, bits per symbol and (data mode) the symbol sequence to transmit."""
        self.nsymbols = _n_symbols(self.modulation)
        self.bits_per_symbol = numpy.log2(self.nsymbols).astype(int)
        self._d_tx = None
        if self.data is not None:
            s, self._pad_bits = _encode(self.data, self.bits_per_symbol)
            self.symbols_per_iter = len(s)
            self._d_tx = torch.from_numpy(numpy.ascontiguousarray(s, dtype=numpy.uint8)).to(self._d_power.device)

    def _launch(self, keep):
        self.generate_symbols()
        self.constellation = define_constellation(self.modulation)
        self.Es = (numpy.abs(self.constellation) ** 2).mean()
        dev = self._d_power.device
        pts = numpy.stack([numpy.real(self.constellation), numpy.imag(self.constellation)], axis=1)
        d_pts = torch.from_numpy(numpy.ascontiguousarray(pts, dtype=numpy.float32).ravel()).to(dev)
        mp = _lib.ModParams()
        mp.n, mp.first = self._d_power.numel(), int(self.first)
        mp.symbols_per_iter, mp.n_symbols = int(self.symbols_per_iter), int(self.nsymbols)
        mp.scheme = {'OOK': _lib.MOD_OOK, 'BPSK': _lib.MOD_BPSK}.get(self.modulation, _lib.MOD_NEAREST)
        mp.has_awgn = int(self.EsN0 is not None)
        mp.es = float(self.Es)
        mp.snr_scale = math.sqrt(10 ** (self.EsN0 / 10)) / self._mean if self.EsN0 is not None else 0.0
        mp.seed = int(self.seed)
        S, n = mp.symbols_per_iter, mp.n
        self._sums = torch.zeros(3, dtype=torch.float64, device=dev)
        sym = rx = dec = None
        if keep:
            sym = torch.empty((S, n), dtype=torch.uint8, device=dev)
            rx = torch.empty((S, n, 2), dtype=torch.float32, device=dev)
            dec = torch.empty((S, n), dtype=torch.uint8, device=dev)
        _lib.modulator_mc(mp, self._d_power, d_pts, self._sums, sym, rx, dec, tx_symbols=self._d_tx)
        self._count = S * n
        if keep:
            self.symbols = sym.cpu().numpy().astype(int)
            r = rx.cpu().numpy().astype(numpy.float64)
            tx = self.constellation[self.symbols]
            self.recv_signal = r[..., 0] if self.modulation == 'OOK' else r[..., 0] + 1j * r[..., 1]
            self.awgn = (self.recv_signal - tx) if self.EsN0 is not None else 0
            self._recv_symbols = dec.cpu().numpy().astype(int)

    def modulate(self):
        if self.modulation == None:  # noqa: E711
            self.recv_signal = self.power
            return self.recv_signal
        self._launch(keep=True)
        return self.recv_signal

    def demodulate(self):
        if self.modulation == None:  # noqa: E711
            self.recv_symbols = None
            return self.recv_symbols
        self.recv_symbols = self._recv_symbols
        if self.data is not None:
            # the reference's decode loop raises (shape / dtype mix-up, fast/comms.py:104-107);
            # this is its evident intent: the bytes each iteration received
            self.recv_data = [_decode(self.recv_symbols[:, i].astype(numpy.uint8), self.bits_per_symbol,
                                      self._pad_bits) for i in range(self.recv_symbols.shape[1])]
        return self.recv_symbols

    def compute_sep(self):
        '''Symbol error probability, from random bits'''
        self.sep = None if self.modulation == None else float(self._host_sums()[0] / self._count)  # noqa: E711
        return self.sep

    def compute_evm(self):
        '''Error Vector Magnitude (EVM), from random bits'''
        if self.modulation == None:  # noqa: E711
            self.evm = None
        else:
            s = self._host_sums()
            self.evm = float((s[1] / self._count) / math.sqrt(s[2] / self._count)) if s[1] > 0 else 0.0
        return self.evm

    def _host_sums(self):
        return self._sums.cpu().numpy()

    def run(self):
        if self.modulation == None:  # noqa: E711
            self.modulate()
            self.demodulate()
        else:
            self._launch(keep=False)
        self.compute_sep()
        self.compute_evm()


class FastFSOC(Fast):
    '''
    Fast simulation plus optical-comms post-processing: after `run()` the received powers are
    passed through a `Modulator` (params 'MODULATION', 'EsN0').
    '''

    def __init__(self, *args, **kwargs):
        super(FastFSOC, self).__init__(*args, **kwargs)
        self.modulation = self.params['MODULATION']
        self.EsN0 = self.params['EsN0']

    def run(self):
        super(FastFSOC, self).run()
        self.modulator = Modulator(self.result.power, self.modulation, self.EsN0)
        self.modulator.run()

    def make_header(self, params):
        hdr = super(FastFSOC, self).make_header(params)
        hdr['MODULATION'] = params['MODULATION']
        hdr['EsN0'] = self.EsN0
        return hdr


# ---- I-Q plane histograms and information measures (fast/comms.py:262-414) ----------------------
def _convolve_awgn_qam_device(samples, M, npxls, EsN0, N0=None, region_size="individual", shot=False):
    constellation = define_constellation(f"{M}-QAM")
    if region_size == "individual":
        region = 1 / (numpy.sqrt(M) - 1)
    elif region_size == "full":
        region = 2
    else:
        raise ValueError("decision_region_size must be either 'full' or 'individual'")
    d, is_complex, n = _field_samples(samples)
    amp, sums = _lib.amplitudes(d, is_complex)
    mean_amp = float(sums[0].item()) / n
    constellation_norm = constellation * mean_amp
    region_norm = region * mean_amp
    if N0 == None:  # noqa: E711
        N0 = numpy.mean(numpy.abs(constellation_norm) ** 2) / 10 ** (EsN0 / 10)
    if region_size == "full":
        # wide noise: grow the region to +-2 sigma around the outermost points
        need = 2 * (mean_amp / numpy.sqrt(2) + 2 * numpy.sqrt(N0))
        if need > region_norm:
            logger.debug("AWGN noise level too large for region, increasing region size")
            region_norm = need
    dx = region_norm / npxls
    sigma2 = max(N0 / (2 * dx ** 2), 1)           # variance in pixel units, at least one pixel
    x_g = numpy.linspace(-npxls / 2, npxls / 2, npxls + 1)
    taps = numpy.exp(-x_g ** 2 / sigma2) / numpy.sqrt(numpy.pi * sigma2)
    x = numpy.linspace(-region_norm / 2, region_norm / 2, npxls + 1)
    ex = numpy.tile(x, (len(constellation), 1))
    ey = ex.copy()
    if region_size == "individual":
        ex += constellation_norm.real[:, None]
        ey += constellation_norm.imag[:, None]
    dev = d.device
    pts = numpy.stack([constellation.real, constellation.imag], axis=1).ravel()
    d_ex, d_ey = _f64(ex.ravel(), dev), _f64(ey.ravel(), dev)
    counts = _lib.iq_histogram(amp, _f64(pts, dev), d_ex, d_ey, npxls)
    return _lib.iq_convolve(counts, n, _f64(taps, dev), shot=shot, sigma2=sigma2, mean_amp=mean_amp,
                            edges_x=d_ex, edges_y=d_ey)


def convolve_awgn_qam(samples, M, npxls, EsN0, N0=None, region_size="individual", shot=False):
    '''
    Received I-Q plane of M-ary QAM under AWGN from complex field measurements (or amplitudes):
    for every constellation point the samples are binned on npxls x npxls pixels over the
    decision region ("individual") or the whole plane ("full") and convolved with the noise
    Gaussian (shot=True: signal-dependent width per occupied bin).

    Returns:
        out (numpy.ndarray): (nsymbols x npxls x npxls) probability per pixel and symbol.
    '''
    return _convolve_awgn_qam_device(samples, M, npxls, EsN0, N0, region_size, shot).cpu().numpy()


def _information(samples, M, npxls, EsN0, N0, shot):
    f = _convolve_awgn_qam_device(samples, M, npxls, EsN0, N0=N0, region_size="full", shot=shot)
    gray = torch.from_numpy(_gray_ints(M).astype(numpy.int32)).to(f.device)
    return _lib.iq_information(f, gray, int(numpy.log2(M))).cpu().numpy()


def mutual_information_qam(samples, M, npxls, EsN0, N0=None, shot=False):
    '''Mutual information [bits/symbol] of a memoryless receiver (Alvarado et al 2016, eq. 16).'''
    return float(_information(samples, M, npxls, EsN0, N0, shot)[0])


def generalised_mutual_information_qam(samples, M, npxls, EsN0, N0=None, shot=False):
    '''Generalised mutual information [bits/symbol]: bit-wise decoder, Gray-coded square QAM.'''
    return float(_information(samples, M, npxls, EsN0, N0, shot)[1])


# ---- bytes <-> symbols (fast/comms.py:509-560) ----------------------------------------------
def _encode(bs, bps):
    """bytes -> symbols of `bps` bits (MSB first), zero-padded at the end; returns (symbols, pad)."""
    bits = numpy.unpackbits(numpy.frombuffer(bs, dtype=numpy.uint8))
    if bps == 1:
        return bits, 0
    pad = int((-len(bits)) % int(bps))
    if pad:
        bits = numpy.pad(bits, [0, pad])
    weights = 2 ** numpy.arange(bps, dtype=numpy.uint8)[::-1]
    return (bits.reshape(-1, bps) * weights).sum(1).flatten().astype(numpy.uint8), pad


def _decode(symbols, bps, pad_bits=0):
    if bps == 1:
        return numpy.packbits(symbols)
    bits = numpy.unpackbits(symbols).reshape(-1, 8)[:, -bps:].flatten()
    out = numpy.packbits(bits).tobytes()
    return out[:-1] if pad_bits > 0 else out      # the last byte holds only padding


def flip_bits(data, ber):
    """Flip each bit of a string / array independently with probability `ber`."""
    if isinstance(data, str):
        raw = data.encode("ascii")
    elif isinstance(data, numpy.ndarray):
        raw = data.tobytes()
    else:
        raise Exception("String or numpy array as data please")
    bits = numpy.unpackbits(numpy.frombuffer(raw, dtype=numpy.uint8))
    bits[numpy.random.rand(len(bits)) < ber] ^= 1
    packed = numpy.packbits(bits)
    if isinstance(data, str):
        return (packed % 128).tobytes().decode("ascii")
    return numpy.frombuffer(packed.tobytes(), dtype=data.dtype).reshape(data.shape)
