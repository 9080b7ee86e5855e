"""ctypes binding of libfastb.so (C ABI: include/fastb.h).

PyTorch is used only to own device memory and streams; every wrapper takes torch CUDA tensors,
passes their raw device pointers and the current stream, and raises `FastbError` on a non-zero
return code.  There is no CPU fallback: importing this module fails loudly when the shared
library has not been built (python build_fastb.py)."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FASTB_LIBRARY: tuning runs only (an alternative build of the same ABI, e.g. libfastb_tune.so)
LIB_PATH = os.environ.get('FASTB_LIBRARY') or os.path.join(_HERE, 'libfastb.so')

MAX_LAYERS = 32
AO_NOAO, AO_AO, AO_LGSAO = 0, 1, 2
ALGO_AUTO, ALGO_DIRECT, ALGO_RADIX, ALGO_RADIX_PAIR, ALGO_BLUESTEIN = 0, 1, 2, 3, 4
RUN_PREPARED, RUN_RNG_FAST = 1, 2


class FastbError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). '
        'Build it with `python build_fastb.py` (needs nvcc, targets sm_100a).')

lib = C.CDLL(LIB_PATH)


class PsdParams(C.Structure):
    _fields_ = [('n', C.c_int32), ('n_layers', C.c_int32), ('ao_mode', C.c_int32),
                ('alias', C.c_int32), ('lmax', C.c_int32), ('kmax', C.c_int32),
                ('reserved0', C.c_int32), ('reserved1', C.c_int32),
                ('df', C.c_double), ('k', C.c_double), ('wvl', C.c_double),
                ('L0', C.c_double), ('l0', C.c_double), ('dsubap', C.c_double),
                ('tloop', C.c_double), ('texp', C.c_double), ('noise_var', C.c_double),
                ('dtheta', C.c_double * 2),
                ('h', C.c_double * MAX_LAYERS), ('cn2', C.c_double * MAX_LAYERS),
                ('vx', C.c_double * MAX_LAYERS), ('vy', C.c_double * MAX_LAYERS)]


class PsdInputs(C.Structure):
    _fields_ = [('d_lf_mask', C.c_void_p), ('d_zfilter', C.c_void_p), ('d_pupil_filter', C.c_void_p)]


class PsdOutputs(C.Structure):
    _fields_ = [('d_powerspec', C.c_void_p), ('d_powerspec_per_layer', C.c_void_p),
                ('d_turb', C.c_void_p), ('d_g_ao', C.c_void_p), ('d_alias', C.c_void_p),
                ('d_noise', C.c_void_p), ('d_logamp', C.c_void_p), ('d_integrands', C.c_void_p),
                ('d_weight', C.c_void_p), ('d_weight_per_layer', C.c_void_p)]


class RunParams(C.Structure):
    _fields_ = [('n', C.c_int32), ('n_pup', C.c_int32), ('lo', C.c_int32), ('coherent', C.c_int32),
                ('algo', C.c_int32), ('flags', C.c_int32),
                ('n_pairs', C.c_int64), ('first_pair', C.c_int64), ('pairs_per_chunk', C.c_int64),
                ('seed', C.c_uint64), ('u_sum', C.c_double), ('sigma_chi', C.c_float),
                ('reserved_f', C.c_float)]


class RunBatch(C.Structure):
    _fields_ = [('n_items', C.c_int32), ('reserved', C.c_int32), ('pairs_per_item', C.c_int64),
                ('d_sigma_chi', C.c_void_p), ('d_seeds', C.c_void_p)]


class RunStats(C.Structure):
    _fields_ = [('db_lo', C.c_double), ('db_hi', C.c_double), ('nbins', C.c_int32), ('reserved', C.c_int32),
                ('d_sums', C.c_void_p), ('d_minmax', C.c_void_p), ('d_hist', C.c_void_p)]


class Subharm(C.Structure):
    _fields_ = [('d_weight', C.c_void_p), ('d_noise', C.c_void_p), ('d_ex', C.c_void_p),
                ('d_ey', C.c_void_p), ('d_mean', C.c_void_p)]


class TemporalParams(C.Structure):
    _fields_ = [('n', C.c_int32), ('n_pup', C.c_int32), ('n_layers', C.c_int32), ('coherent', C.c_int32),
                ('n_steps', C.c_int64), ('u_sum', C.c_double)]


class ZernikeParams(C.Structure):
    _fields_ = [('n', C.c_int32), ('noll_first', C.c_int32), ('noll_last', C.c_int32), ('gtilt', C.c_int32),
                ('clip_box', C.c_int32), ('reserved', C.c_int32), ('df', C.c_double), ('diameter', C.c_double),
                ('d_wfs', C.c_double), ('modal_mult', C.c_double)]


class ModParams(C.Structure):
    _fields_ = [('n', C.c_int64), ('first', C.c_int64), ('symbols_per_iter', C.c_int32),
                ('n_symbols', C.c_int32), ('scheme', C.c_int32), ('has_awgn', C.c_int32),
                ('es', C.c_double), ('snr_scale', C.c_double), ('seed', C.c_uint64)]


LAYER_PAIR_BASE = 1 << 62
CURVE_BER_OOK, CURVE_SEP_QAM = 0, 1
MOD_OOK, MOD_BPSK, MOD_NEAREST = 0, 1, 2

# every symbol include/fastb.h declares (tests/test_abi.py checks the list against the header)
_SIGS = {
    'fastb_psd_build': (C.c_int, [C.POINTER(PsdParams), C.POINTER(PsdInputs), C.POINTER(PsdOutputs), C.c_void_p]),
    'fastb_zernike_filter': (C.c_int, [C.POINTER(ZernikeParams), C.c_void_p, C.c_void_p]),
    'fastb_make_weight': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
    'fastb_simpson2d': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'fastb_pupil_filter': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'fastb_pupil_filter_workspace_bytes': (C.c_int64, [C.c_int32]),
    'fastb_screen_detect_workspace_bytes': (C.c_int64, [C.POINTER(RunParams)]),
    'fastb_screen_detect': (C.c_int, [C.POINTER(RunParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(Subharm), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p]),
    'fastb_screen_detect_batch_workspace_bytes': (C.c_int64, [C.POINTER(RunParams), C.c_int32]),
    'fastb_screen_detect_prepare': (C.c_int, [C.POINTER(RunParams), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int64, C.c_void_p]),
    'fastb_screen_detect_batch': (C.c_int, [C.POINTER(RunParams), C.POINTER(RunBatch), C.POINTER(RunStats),
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p]),
    'fastb_screens_crop': (C.c_int, [C.POINTER(RunParams), C.c_void_p, C.c_void_p, C.POINTER(Subharm), C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    'fastb_rng_dump': (C.c_int, [C.c_uint64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                 C.c_void_p, C.c_void_p]),
    'fastb_rng_dump_mode': (C.c_int, [C.c_uint64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p]),
    'fastb_noise_stride': (C.c_int32, [C.c_int32, C.c_int32]),
    'fastb_rng_dump_stride': (C.c_int, [C.c_uint64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                        C.c_int64, C.c_void_p, C.c_void_p]),
    'fastb_layer_screens_workspace_bytes': (C.c_int64, [C.c_int32, C.c_int32]),
    'fastb_layer_screens': (C.c_int, [C.c_int32, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]),
    'fastb_temporal_detect': (C.c_int, [C.POINTER(TemporalParams)] + [C.c_void_p] * 8 + [C.c_void_p]),
    'fastb_temporal_coords': (C.c_int, [C.c_int32] * 6 + [C.c_void_p] * 5 + [C.c_void_p]),
    'fastb_stats': (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int32, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    'fastb_error_curve': (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    'fastb_fade_stats': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'fastb_modulator_mc': (C.c_int, [C.POINTER(ModParams)] + [C.c_void_p] * 8),
    'fastb_amplitudes': (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'fastb_iq_histogram': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_void_p]),
    'fastb_iq_convolve_workspace_bytes': (C.c_int64, [C.c_int32, C.c_int32]),
    'fastb_iq_convolve': (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p]),
    'fastb_iq_information': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p]),
    'fastb_version': (C.c_int, []),
    'fastb_last_error': (C.c_char_p, []),
    'fastb_device_count': (C.c_int, []),
    'fastb_launch_count': (C.c_int64, []),
    'fastb_reset_launch_count': (None, []),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED = tuple(_SIGS)


def _check(rc, what):
    if rc != 0:
        raise FastbError(f'{what} failed (code {rc}): {lib.fastb_last_error().decode()}')


def _ptr(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise FastbError('expected a CUDA tensor (there is no CPU path)')
    if not t.is_contiguous():
        raise FastbError('expected a contiguous tensor')
    if dtype is not None and t.dtype != dtype:
        raise FastbError(f'expected dtype {dtype}, got {t.dtype}')
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    """The product path needs a GPU; fail loudly otherwise."""
    if not torch.cuda.is_available() or lib.fastb_device_count() < 1:
        raise FastbError('fast_b200 needs a CUDA device (B200 / sm_100a); no CPU fallback exists')


def version():
    return lib.fastb_version()


def launch_count():
    return int(lib.fastb_launch_count())


def reset_launch_count():
    lib.fastb_reset_launch_count()


def psd_build(params: PsdParams, outputs: dict, lf_mask=None, zfilter=None, pupil_filter=None):
    """outputs: name (field of PsdOutputs without the d_ prefix) -> tensor."""
    f64 = torch.float64
    ins = PsdInputs(_ptr(lf_mask, f64), _ptr(zfilter, f64), _ptr(pupil_filter, f64))
    outs = PsdOutputs()
    for name, t in outputs.items():
        want = torch.float32 if name.startswith('weight') else f64
        setattr(outs, 'd_' + name, _ptr(t, want))
    _check(lib.fastb_psd_build(C.byref(params), C.byref(ins), C.byref(outs), _stream()), 'fastb_psd_build')


def zernike_filter(n, df, device, noll_first=1, noll_last=0, diameter=0.0, d_wfs=0.0, modal_mult=1.0,
                   gtilt=False, clip_box=False):
    """Zernike squared filter / modal corrected-region mask on the n x n grid -> float64 [n, n]."""
    zp = ZernikeParams(n=int(n), noll_first=int(noll_first), noll_last=int(noll_last), gtilt=int(bool(gtilt)),
                       clip_box=int(bool(clip_box)), df=float(df), diameter=float(diameter), d_wfs=float(d_wfs),
                       modal_mult=float(modal_mult))
    out = torch.empty((n, n), dtype=torch.float64, device=device)
    _check(lib.fastb_zernike_filter(C.byref(zp), _ptr(out), _stream()), 'fastb_zernike_filter')
    return out


def make_weight(W, df, out=None):
    n = W.shape[-1]
    batch = W.numel() // (n * n)
    if out is None:
        out = torch.empty(W.shape, dtype=torch.float32, device=W.device)
    _check(lib.fastb_make_weight(_ptr(W, torch.float64), n, batch, float(df), _ptr(out, torch.float32),
                                 _stream()), 'fastb_make_weight')
    return out


def simpson2d(P, w):
    n = P.shape[-1]
    batch = P.numel() // (n * n)
    out = torch.empty(batch, dtype=torch.float64, device=P.device)
    _check(lib.fastb_simpson2d(_ptr(P, torch.float64), n, batch, _ptr(w, torch.float64), _ptr(out),
                               _stream()), 'fastb_simpson2d')
    return out


def pupil_filter(pm):
    n = pm.shape[-1]
    out = torch.empty_like(pm)
    nbytes = lib.fastb_pupil_filter_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=pm.device)
    _check(lib.fastb_pupil_filter(_ptr(pm, torch.float64), n, _ptr(out), _ptr(ws), nbytes, _stream()),
           'fastb_pupil_filter')
    return out


def screen_detect_workspace_bytes(rp: RunParams, n_items=1):
    nbytes = lib.fastb_screen_detect_batch_workspace_bytes(C.byref(rp), int(n_items))
    if nbytes < 0:
        raise FastbError('fastb_screen_detect_workspace_bytes: ' + lib.fastb_last_error().decode())
    return int(nbytes)


def screen_detect_prepare(rp: RunParams, weight, U, workspace, n_items=1):
    """Fill the workspace's derived tables once (transposed U, pre-scaled weight copies / chirp
    tables); later calls with RUN_PREPARED in rp.flags are then a single kernel launch."""
    f32 = torch.float32
    _check(lib.fastb_screen_detect_prepare(C.byref(rp), int(n_items), _ptr(weight, f32), _ptr(U, f32),
                                           _ptr(workspace), workspace.numel() * workspace.element_size(),
                                           _stream()), 'fastb_screen_detect_prepare')


def run_stats(db_lo, db_hi, nbins, sums, minmax, hist):
    return RunStats(db_lo=float(db_lo), db_hi=float(db_hi), nbins=int(nbins), d_sums=_ptr(sums, torch.float64),
                    d_minmax=_ptr(minmax, torch.float64), d_hist=_ptr(hist, torch.int64))


def screen_detect_batch(rp: RunParams, weight, U, out_a, out_b, workspace, batch=None, stats=None, chi=None):
    """K2 with the optional batch of configurations (batch: dict(n_items, pairs_per_item, sigma_chi f32
    tensor, seeds int64 tensor holding the uint64 bit patterns)) and the fused statistics
    (stats: a RunStats from run_stats())."""
    f32 = torch.float32
    b = None
    if batch is not None:
        b = C.byref(RunBatch(n_items=int(batch['n_items']), pairs_per_item=int(batch['pairs_per_item']),
                             d_sigma_chi=_ptr(batch['sigma_chi'], f32), d_seeds=_ptr(batch['seeds'], torch.int64)))
    s = C.byref(stats) if stats is not None else None
    _check(lib.fastb_screen_detect_batch(C.byref(rp), b, s, _ptr(weight, f32), _ptr(U, f32), _ptr(chi, f32),
                                         _ptr(out_a, f32), _ptr(out_b, f32), _ptr(workspace),
                                         workspace.numel() * workspace.element_size(), _stream()),
           'fastb_screen_detect_batch')


def screen_detect(rp: RunParams, weight, U, out_a, out_b, workspace, chi=None, noise=None, subharm=None):
    """subharm: None or dict(weight=(27,) f32, ex=(3,Pp,2) f32, ey=(3,Pp,2) f32, mean=(27,2) f32,
    noise=None | (n_pairs,27,2) f32) of CUDA tensors."""
    f32 = torch.float32
    sh = None
    if subharm is not None:
        sh = C.byref(Subharm(_ptr(subharm['weight'], f32), _ptr(subharm.get('noise'), f32),
                             _ptr(subharm['ex'], f32), _ptr(subharm['ey'], f32), _ptr(subharm['mean'], f32)))
    _check(lib.fastb_screen_detect(C.byref(rp), _ptr(weight, f32), _ptr(U, f32), _ptr(chi, f32),
                                   _ptr(noise), sh, _ptr(out_a, f32), _ptr(out_b, f32), _ptr(workspace),
                                   workspace.numel() * workspace.element_size(), _stream()),
           'fastb_screen_detect')


def screens_crop(rp: RunParams, weight, phs, workspace, noise=None, subharm=None):
    """Materialise the cropped screens of the pairs described by rp into phs (2*n_pairs, P, P)."""
    f32 = torch.float32
    sh = None
    if subharm is not None:
        sh = C.byref(Subharm(_ptr(subharm['weight'], f32), _ptr(subharm.get('noise'), f32),
                             _ptr(subharm['ex'], f32), _ptr(subharm['ey'], f32), _ptr(subharm['mean'], f32)))
    _check(lib.fastb_screens_crop(C.byref(rp), _ptr(weight, f32), _ptr(noise), sh, _ptr(phs, f32),
                                  _ptr(workspace), workspace.numel() * workspace.element_size(), _stream()),
           'fastb_screens_crop')


def noise_stride(n, n_pup):
    """Noise blocks per row of the K2 device RNG for an (n, n_pup) problem (include/fastb.h)."""
    s = lib.fastb_noise_stride(int(n), int(n_pup))
    if s < 1:
        raise FastbError(f'fastb_noise_stride: bad geometry n={n}, n_pup={n_pup}')
    return s


def rng_dump(seed, pair, n, device, chi_first=0, chi_count=0, want_tile=True, fast=False, n_pup=None):
    """Noise tile of one pair (and chi normals).  n_pup given: the K2 stride of that problem; otherwise
    ceil(n / 16) (K2 for powers of two, K4 layer screens)."""
    tile = torch.empty((n, n, 2), dtype=torch.float32, device=device) if want_tile else None
    chi = torch.empty(chi_count, dtype=torch.float32, device=device) if chi_count else None
    stride = noise_stride(n, n_pup) if n_pup is not None else (n + 15) // 16
    _check(lib.fastb_rng_dump_stride(int(seed), int(pair), n, stride, int(bool(fast)), _ptr(tile), int(chi_first),
                                     int(chi_count), _ptr(chi), _stream()), 'fastb_rng_dump_stride')
    return tile, chi


def stats(r, db_lo, db_hi, nbins, sums, minmax, hist):
    _check(lib.fastb_stats(_ptr(r, torch.float32), r.numel(), float(db_lo), float(db_hi), int(nbins),
                           _ptr(sums, torch.float64), _ptr(minmax, torch.float64), _ptr(hist, torch.int64),
                           _stream()), 'fastb_stats')


def layer_screens(weight_per_layer, seed, noise=None):
    """K4a: (L, N, N) signed per-layer weights -> (L, N, N) float32 real screens."""
    L, n = weight_per_layer.shape[0], weight_per_layer.shape[-1]
    out = torch.empty((L, n, n), dtype=torch.float32, device=weight_per_layer.device)
    nbytes = lib.fastb_layer_screens_workspace_bytes(n, L)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=out.device)
    nz = None if noise is None else torch.view_as_real(noise.contiguous())
    _check(lib.fastb_layer_screens(n, L, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(weight_per_layer, torch.float32),
                                   _ptr(nz), _ptr(out), _ptr(ws), nbytes, _stream()), 'fastb_layer_screens')
    return out


def temporal_coords(n, n_pup, lo, pixel_shifts, n_chunks):
    """K4c: (L, 2, J) float64 per-step pixel shifts -> xi, xf, yi, yf, each (L, n_chunks * J, n_pup), for every step
    of a run (the reference's chunk-after-chunk coordinate bookkeeping, fast/fast.py:617-635)."""
    L, _, J = pixel_shifts.shape
    dev = pixel_shifts.device
    shape = (L, int(n_chunks) * J, int(n_pup))
    xi = torch.empty(shape, dtype=torch.int32, device=dev)
    yi = torch.empty(shape, dtype=torch.int32, device=dev)
    xf = torch.empty(shape, dtype=torch.float32, device=dev)
    yf = torch.empty(shape, dtype=torch.float32, device=dev)
    _check(lib.fastb_temporal_coords(int(n), int(n_pup), int(lo), L, J, int(n_chunks), _ptr(pixel_shifts, torch.float64),
                                     _ptr(xi, torch.int32), _ptr(xf, torch.float32), _ptr(yi, torch.int32),
                                     _ptr(yf, torch.float32), _stream()), 'fastb_temporal_coords')
    return xi, xf, yi, yf


def temporal_detect(tp: TemporalParams, screens, xi, xf, yi, yf, U, chi, out):
    f32, i32 = torch.float32, torch.int32
    _check(lib.fastb_temporal_detect(C.byref(tp), _ptr(screens, f32), _ptr(xi, i32), _ptr(xf, f32),
                                     _ptr(yi, i32), _ptr(yf, f32), _ptr(U, f32), _ptr(chi, f32),
                                     _ptr(out, f32), _stream()), 'fastb_temporal_detect')


# ---- K5: link metrics (fast/comms.py consumers) --------------------------------------------
def error_curve(samples, kind, qam_order, snr_db):
    """-> float64 tensor [k + 1]: curve and the sample mean."""
    out = torch.empty(snr_db.numel() + 1, dtype=torch.float64, device=samples.device)
    _check(lib.fastb_error_curve(_ptr(samples, torch.float32), samples.numel(), int(kind), int(qam_order),
                                 _ptr(snr_db, torch.float64), snr_db.numel(), _ptr(out), _stream()),
           'fastb_error_curve')
    return out


def fade_stats(series, thresholds):
    """-> int64 tensor [k, 4]: below, complete fades, samples inside them, last index not below."""
    out = torch.empty((thresholds.numel(), 4), dtype=torch.int64, device=series.device)
    _check(lib.fastb_fade_stats(_ptr(series, torch.float32), series.numel(), _ptr(thresholds, torch.float64),
                                thresholds.numel(), _ptr(out), _stream()), 'fastb_fade_stats')
    return out


def modulator_mc(mp: ModParams, power, constellation, sums, symbols=None, recv=None, recv_symbols=None,
                 tx_symbols=None):
    _check(lib.fastb_modulator_mc(C.byref(mp), _ptr(power, torch.float32), _ptr(constellation, torch.float32),
                                  _ptr(sums, torch.float64), _ptr(symbols, torch.uint8),
                                  _ptr(recv, torch.float32), _ptr(recv_symbols, torch.uint8),
                                  _ptr(tx_symbols, torch.uint8), _stream()), 'fastb_modulator_mc')


def amplitudes(samples, is_complex):
    """-> (float64 |z| [n], float64 [2] = sum |z|, sum |z|^2)."""
    n = samples.numel() // (2 if is_complex else 1)
    amp = torch.empty(n, dtype=torch.float64, device=samples.device)
    sums = torch.empty(2, dtype=torch.float64, device=samples.device)
    _check(lib.fastb_amplitudes(_ptr(samples, torch.float32), int(bool(is_complex)), n, _ptr(amp), _ptr(sums),
                                _stream()), 'fastb_amplitudes')
    return amp, sums


def iq_histogram(amp, points, edges_x, edges_y, npxls):
    m = points.numel() // 2
    counts = torch.empty((m, npxls, npxls), dtype=torch.int32, device=amp.device)
    _check(lib.fastb_iq_histogram(_ptr(amp, torch.float64), amp.numel(), _ptr(points, torch.float64), m,
                                  _ptr(edges_x, torch.float64), _ptr(edges_y, torch.float64), int(npxls),
                                  _ptr(counts), _stream()), 'fastb_iq_histogram')
    return counts


def iq_convolve(counts, n, taps, shot=False, sigma2=1.0, mean_amp=1.0, edges_x=None, edges_y=None):
    m, npxls, _ = counts.shape
    nbytes = lib.fastb_iq_convolve_workspace_bytes(m, npxls)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=counts.device)
    out = torch.empty((m, npxls, npxls), dtype=torch.float64, device=counts.device)
    _check(lib.fastb_iq_convolve(_ptr(counts, torch.int32), int(n), m, npxls, _ptr(taps, torch.float64),
                                 int(bool(shot)), float(sigma2), float(mean_amp), _ptr(edges_x, torch.float64),
                                 _ptr(edges_y, torch.float64), _ptr(out), _ptr(ws), nbytes, _stream()),
           'fastb_iq_convolve')
    return out


def iq_information(f, gray, n_bits):
    """-> float64 [2]: mutual information, generalised mutual information (bits/symbol)."""
    m, npxls, _ = f.shape
    out = torch.empty(2, dtype=torch.float64, device=f.device)
    _check(lib.fastb_iq_information(_ptr(f, torch.float64), m, npxls, _ptr(gray, torch.int32), int(n_bits),
                                    _ptr(out), _stream()), 'fastb_iq_information')
    return out
