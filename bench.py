"""Benchmark of the FAST Monte-Carlo hot path (screen generation + detection, PSD already built).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5]

One "step" = one pass of the hot path over one batch of the named workload on every rank: ONE K2
launch (noise -> screens -> detector with the result statistics fused into its epilogue,
fastb_screen_detect_batch) followed, at N > 1, by ONE small collective that combines the ranks'
moments / extrema / histogram.  Weak scaling: every rank processes a full batch of its own
realisation range.  The batch is sized so that a step is >= 100 ms of GPU work (C2: 12 x the
config's 1e5 realisations).  Prints ONE JSON line on rank 0 (contract: task statement / DESIGN.md):

  value         device-timed throughput of the steps (CUDA events, L2 flushed between steps, max over ranks)
  e2e           the same workload through the public object: host weight/U -> device, Fast(p).run() ->
                host result.power, wall clock (includes the all-gather at N > 1)
  roofline      model-A HBM figure of the K2 kernel + the instruction-issue ceiling (secondary)
  per_workload  every configuration of BASELINE.json (C1 TEMPORAL, C1' N=164, C2, C3 batched sweep, C4, C5,
                C5 strong-scaled over the ranks), each with its own kernel time and roofline fraction
  check         run-time correctness: statistics of the timed steps; at N > 1 the sharded run is compared
                bit for bit with an unsharded one

--impl reference times the reference's CPU algorithm (the numpy oracle port of
fast/fast.py:115-140,589-668: /root/reference is pure Python and cannot travel to the GPU box)
on all host cores, on bounded samples of the same workload.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MC realizations/sec (screen+detect)"
UNIT = "realizations/s"

# name: (config factory, realisations per step per GPU, N, n_pup, description)
WORKLOADS = {
    'c2': ('c2', 1200000, 256, 82,
           "C2: GEO ground-station downlink, 256x256 grid, AO residual PSD + scintillation, SMF detection; "
           "1.2e6 realizations per step per GPU (12 x the config's 1e5, so that a step is >= 100 ms)"),
    'c4': ('c4', 300000, 512, 162, "C4: coherent detection, 512x512 grid, 3e5 realizations per step per GPU"),
    'c5': ('c5', 80000, 1024, 162, "C5 shard: 1024x1024 grid, 8e4 realizations per step per GPU"),
}


def algorithmic_bytes(N, coherent):
    """SURVEY.md 8(d) model A: one complex64 N x N round trip per pair = 8 N^2 bytes per
    realisation, + one fp32 scalar (complex: two)."""
    return 8 * N * N + (8 if coherent else 4)


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


_DIGEST_FILES = {
    # the sources that define the K2 radix kernels (device code, FFT, RNG, instance selection) ...
    'radix': ('fastb_common.cuh', 'fft_core.cuh', 'screen_detect_kernel.cuh', 'screen_detect_radix.cu'),
    # ... and the chirp-z kernel (C1')
    'bluestein': ('fastb_common.cuh', 'fft_core.cuh', 'screen_detect_kernel.cuh', 'bluestein.cuh',
                  'screen_detect_bluestein.cu', 'screen_detect_bluestein_m.cu'),
}


def kernel_digest(kind='radix'):
    """sha256 over the sources of one K2 kernel family: the ncu-derived counters in profiles/ are only used when
    they were captured from exactly these sources."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'fast_b200', 'csrc')
    for f in _DIGEST_FILES[kind]:
        with open(os.path.join(d, f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_counters(workload):
    """Per-pair counters of the K2 kernel from the committed ncu capture of this round
    (profiles/kernel_counters_r02.json, written by tools/ncu_counters.py): DRAM bytes and executed warp
    instructions.  Returns (dict | None, note): None when there is no capture of the current sources."""
    path = os.path.join(ROOT, 'profiles', 'kernel_counters_r02.json')
    try:
        rec = json.load(open(path))
    except Exception:
        return None, 'no ncu capture committed'
    w = rec.get('workloads', {}).get(workload)
    if not w:
        return None, 'no ncu capture for this workload'
    kind = 'bluestein' if workload == 'c1prime' else 'radix'
    if rec.get('kernel_digests', {}).get(kind) != kernel_digest(kind):
        return None, f"ncu capture is of other kernel sources ({rec.get('kernel_digests', {}).get(kind)}): not used"
    return w, 'ncu --set full capture of these sources (profiles/kernel_counters_r02.json)'


def config_of(workload, world):
    factory, n_real, N, P, desc = WORKLOADS[workload]
    return {"workload": desc, "N": N, "n_pup": P, "layers": 4, "realizations_per_step_per_gpu": n_real,
            "rng": "device Philox4x32-10 + Box-Muller",
            "l2": "flushed (256 MiB write) between timed steps",
            "parallelism": f"realization ranges sharded over {world} GPU(s), moments+histogram combined by one "
                           "collective per step"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.note = None

    def run(self):
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20,
                 'hw_thermal_slowdown': 0x40, 'hw_power_brake_slowdown': 0x80}
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_mask = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                getattr(pynvml, 'nvmlDeviceGetCurrentClocksThrottleReasons')
            while not self.stop_flag:
                self.samples.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    mask = int(get_mask(h))
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception as e:
                    self.note = f'reasons unavailable: {type(e).__name__}'
                time.sleep(0.05)
        except Exception as e:
            self.note = f'pynvml failed ({type(e).__name__}: {e}); nvidia-smi polling'
            q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
                'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
            keys = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            while not self.stop_flag:
                try:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits'], capture_output=True, text=True).stdout
                    f = [x.strip() for x in out.strip().split(',')]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for k, v in zip(keys, f[2:]):
                        if v.lower().startswith('active'):
                            self.reasons.add(k)
                except Exception:
                    break

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_min_mhz': s[0] if s else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(s),
                'note': self.note}


# ------------------------------------------------------------------------- CPU (oracle port)
_W = {}


def _cpu_worker_init(workload, seed_base):
    import numpy as np
    from oracle import configs, fast_oracle as fo
    factory = WORKLOADS[workload][0]
    p = getattr(configs, factory)(niter=2, nchunks=1)
    _W['init'] = fo.build(p)
    _W['fo'] = fo
    _W['rng'] = np.random.default_rng(seed_base + os.getpid())


def _cpu_worker_step(n_real):
    """n_real realisations in chunks of <= 50 pairs, exactly the reference's chunk loop."""
    fo, init, rng = _W['fo'], _W['init'], _W['rng']
    # chunk so that one process holds ~32 MiB of complex128 noise (the reference's NCHUNKS knob)
    chunk = max(2, 2 * ((1 << 21) // (init['N'] * init['N'])))
    done = 0
    while done < n_real:
        m = min(chunk, n_real - done)
        fo.run_mc(init, rng, niter=m, nchunks=1)
        done += m
    return done


def cpu_psd_build_seconds(workload):
    """Wall time of the oracle's whole init (PSD build dominated by the aliasing sum), 1 core."""
    from oracle import configs, fast_oracle as fo
    factory = WORKLOADS[workload][0]
    p = getattr(configs, factory)(niter=2, nchunks=1)
    t0 = time.perf_counter()
    fo.build(p)
    return time.perf_counter() - t0


def cpu_throughput(workload, n_proc, n_real_per_proc, steps=1, warmup=0):
    """realisations/s of the numpy oracle port on n_proc host processes."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Pool(n_proc, initializer=_cpu_worker_init, initargs=(workload, 1000)) as pool:
        for _ in range(warmup):
            pool.map(_cpu_worker_step, [n_real_per_proc] * n_proc)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_worker_step, [n_real_per_proc] * n_proc)
        dt = time.perf_counter() - t0
    return n_proc * n_real_per_proc * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    cores = min(os.cpu_count() or 1, 128)
    rate_guess = {'c2': 200.0, 'c4': 45.0, 'c5': 10.0}[args.workload]
    budget = 150.0 / max(1, args.steps + args.warmup)              # seconds per step
    per_proc = max(2, int(budget * rate_guess * 0.6) // 2 * 2)
    value, dt = cpu_throughput(args.workload, cores, per_proc, steps=args.steps, warmup=args.warmup)
    sample = f"{cores} processes x {per_proc} realizations per step (numpy oracle port of the reference chunk loop)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(args.workload, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- cuFFT comparator
def cufft_comparator(sim, n_real, steps, batch_pairs=256):
    """Library-built pipeline for the same work, timed in the same process as an FFT
    comparator (BASELINE.json north_star): cuRAND Philox normals (torch.randn) -> x weight ->
    cuFFT batched 2-D C2C (torch.fft.ifft2) -> crop -> exp / sum reductions (torch).  Not the
    product path; it materialises full N x N screens in HBM like the reference does."""
    import torch
    dev = sim.device
    N, P, lo = sim.Npxls, sim.Npxls_pup, sim._lo
    w_abs = sim._d['weight'].abs().to(torch.complex64) * (N * N)       # ifft2 normalises by 1/N^2
    w_shift = torch.fft.ifftshift(w_abs)                               # centred spectrum -> FFT order
    U = sim._d['U']
    inv = 1.0 / sim._u_sum
    sig = float(sim.logamp_var) ** 0.5
    n_pairs = n_real // 2

    def one_step():
        acc = []
        for p0 in range(0, n_pairs, batch_pairs):
            b = min(batch_pairs, n_pairs - p0)
            z = torch.randn((b, N, N, 2), device=dev)
            s = torch.view_as_complex(z) * w_shift
            scr = torch.fft.fftshift(torch.fft.ifft2(s), dim=(-1, -2))[:, lo:lo + P, lo:lo + P]
            chi = sig * torch.randn((2, b), device=dev)
            za = (U * torch.exp(1j * scr.real)).sum((1, 2)) * inv * torch.exp(chi[0])
            zb = (U * torch.exp(1j * scr.imag)).sum((1, 2)) * inv * torch.exp(chi[1])
            acc.append(torch.cat([za.abs() ** 2, zb.abs() ** 2]))
        return torch.cat(acc)

    r = one_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = one_step()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"kind": "cuRAND (torch.randn) + cuFFT (torch.fft.ifft2, batched C2C) + torch reductions, "
                    "full N x N screens in HBM", "value": n_real / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms, "mean_r": float(r.mean())}


# ------------------------------------------------------------------------- GPU
# ------------------------------------------------------------------------- K5 link metrics
def k5_link_metrics(with_cpu):
    """Secondary figures for the consumers of the per-realisation output:
    device time of each K5 entry point on synthetic samples (CUDA events, best of 5 after a
    warm-up) and the CPU oracle (numpy, 1 core) on a bounded sample of the same work."""
    import numpy as np
    import torch
    from fast_b200 import _lib, comms
    rng = np.random.default_rng(1)
    n = 4_000_000
    host = np.exp(0.35 * rng.standard_normal(n)).astype(np.float32)
    x = torch.from_numpy(host).cuda()
    snr = torch.linspace(0, 30, 32, dtype=torch.float64).cuda()
    thr = torch.tensor([0.1, 0.2, 0.3, 0.5, 0.7, 1.0, 1.5, 2.0], dtype=torch.float64).cuda()

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    out = {}
    ms = timed(lambda: _lib.error_curve(x, _lib.CURVE_SEP_QAM, 16, snr))
    out["error_curve"] = {"ms": ms, "samples": n, "snr_points": 32, "evaluations_per_s": n * 32 / (ms * 1e-3)}
    ms = timed(lambda: _lib.fade_stats(x, thr))
    out["fade_stats"] = {"ms": ms, "samples": n, "thresholds": 8, "GB_per_s": n * 4 * 8 / (ms * 1e6)}
    mod = comms.Modulator(x[:100_000], '16-QAM', EsN0=12.0, symbols_per_iter=1000, seed=1)
    ms = timed(lambda: mod.run(), reps=3)
    out["modulator_16qam"] = {"ms": ms, "symbols": 100_000 * 1000, "symbols_per_s": 1e8 / (ms * 1e-3),
                              "sep": mod.sep, "note": "fused draw + AWGN + nearest-point decision; includes host glue"}
    amp = torch.sqrt(x[:1_000_000]).contiguous()
    ms = timed(lambda: comms._information(amp, 16, 128, 14.0, None, False), reps=3)
    out["iq_information_16qam_128px"] = {"ms": ms, "samples": 1_000_000,
                                         "note": "amplitudes + histograms + AWGN convolution + MI/GMI"}
    if with_cpu:
        from oracle import comms_oracle as co
        sub = host[:250_000].astype(np.float64)
        t0 = time.perf_counter()
        for s_db in (0.0, 10.0, 20.0, 30.0):
            co.sep_qam(16, s_db, sub)
        dt = time.perf_counter() - t0
        out["error_curve"]["cpu_oracle_evaluations_per_s"] = sub.size * 4 / dt
        np.random.seed(1)
        t0 = time.perf_counter()
        co.modulator(host[:2000].astype(np.float64), '16-QAM', 12.0, 100)
        out["modulator_16qam"]["cpu_oracle_symbols_per_s"] = 2000 * 100 / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        co.mutual_information_qam(np.sqrt(sub[:100_000]), 16, 128, 14.0)
        out["iq_information_16qam_128px"]["cpu_oracle_ms_100k_samples_mi_only"] = 1e3 * (time.perf_counter() - t0)
    return out


# ------------------------------------------------------------------------- GPU
class DramSampler:
    """DRAM bandwidth utilisation of the timed region from NVML's GPM counters (Hopper and later),
    when the driver exposes them: a measured-in-run figure beside the ncu capture."""

    def __init__(self, index):
        self.ok, self.note = False, None
        try:
            import pynvml as nv
            self.nv = nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            sup = nv.nvmlGpmQueryDeviceSupport(self.h)
            if not sup.isSupportedDevice:
                raise RuntimeError('GPM not supported')
            self.s0, self.s1 = nv.nvmlGpmSampleAlloc(), nv.nvmlGpmSampleAlloc()
            self.ok = True
        except Exception as e:
            self.note = f'GPM unavailable ({type(e).__name__}: {e})'

    def start(self):
        if self.ok:
            try:
                self.nv.nvmlGpmSampleGet(self.h, self.s0)
            except Exception as e:
                self.ok, self.note = False, f'GPM sample failed ({type(e).__name__})'

    def stop(self):
        """-> percent of peak DRAM bandwidth over the region, or None."""
        if not self.ok:
            return None
        try:
            nv = self.nv
            nv.nvmlGpmSampleGet(self.h, self.s1)
            mg = nv.c_nvmlGpmMetricsGet_t()
            mg.version = nv.NVML_GPM_METRICS_GET_VERSION
            mg.numMetrics = 1
            mg.sample1, mg.sample2 = self.s0, self.s1
            mg.metrics[0].metricId = nv.NVML_GPM_METRIC_DRAM_BW_UTIL
            nv.nvmlGpmMetricsGet(mg)
            return float(mg.metrics[0].value)
        except Exception as e:
            self.note = f'GPM read failed ({type(e).__name__}: {e})'
            return None


def timed_launches(sim, n_real, reps, flush, sb=None, first_base=0):
    """Median device time [ms] of `reps` K2 launches of n_real realisations (L2 flushed before each)."""
    import torch
    n_pairs = n_real // 2
    sim.screen_detect(first_base, min(n_pairs, 20000), stats=sb)          # warm-up (+ table preparation)
    ms = []
    for r in range(reps):
        flush.fill_(r & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.screen_detect(first_base + (r + 1) * n_pairs, n_pairs, stats=sb)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def per_workload(world, rank, flush, peak, dev):
    """Every configuration of BASELINE.json on this box, outside the headline's timed region: a few
    launches each (CUDA events, L2 flushed), model-A roofline fraction of its own kernel time."""
    import torch
    import torch.distributed as td
    import fast_b200
    from fast_b200 import configs, dist, sweep
    out = {}

    def entry(name, n_real, ms, N, coherent, kernel, note=None, gpus=1):
        b_alg = algorithmic_bytes(N, coherent)
        v = n_real / (ms * 1e-3)
        counters, cnote = ncu_counters(name)
        e = {"value": v, "unit": UNIT, "n_gpus": gpus, "realizations_per_launch_per_gpu": n_real // gpus,
             "kernel": kernel, "kernel_ms": ms, "N": N,
             "roofline_frac": b_alg * v / gpus / 1e9 / peak, "algorithmic_bytes_per_realization": b_alg,
             "dram_bytes_per_launch": (counters['dram_bytes_per_pair'] * (n_real // gpus // 2)) if counters else None,
             "dram_bytes_source": cnote}
        if note:
            e["note"] = note
        out[name] = e

    def sync_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    if world == 1:
        # C1 verbatim: TEMPORAL frozen flow, N = 164 (auto), 100 time steps per run
        sim = fast_b200.Fast(configs.c1(seed=1))
        sim.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            sim.run()
        dt = (time.perf_counter() - t0) / reps
        out['c1_temporal'] = {"value": sim.Niter / dt, "unit": "time steps/s", "n_gpus": 1, "N": sim.Npxls,
                              "run_ms": 1e3 * dt,
                              "kernel": "layer_lines_kernel (chirp-z) + temporal_coords_kernel + temporal_detect_kernel",
                              "note": "test/test_params.py verbatim: wall time of Fast.run() -> host result (chi colouring on "
                                      "the host, layer screens, coordinate bookkeeping of all 10 chunks x 10 steps on the "
                                      "device, one detector launch, D2H), launch-latency-bound at this size"}
        # C1': same grid, TEMPORAL off -> the chirp-z kernel (N = 164 is not a power of two)
        n_real = 200000
        sim = fast_b200.Fast(configs.c1prime(niter=n_real, nchunks=1, seed=1))
        ms = timed_launches(sim, n_real, 3, flush)
        entry('c1prime', n_real, ms, sim.Npxls, False, 'screen_detect_bluestein<M=256, cell class 6>')
        # the O(N^2)-per-line direct kernel on the same grid, for scale
        from fast_b200 import _lib
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sim.screen_detect(0, 2000, algo=_lib.ALGO_DIRECT)
        e0.record()
        sim.screen_detect(0, 10000, algo=_lib.ALGO_DIRECT)
        e1.record()
        e1.synchronize()
        out['c1prime']['direct_dft_value'] = 20000 / (e0.elapsed_time(e1) * 1e-3)
        for name, factory, n_real, kern in (('c2', 'c2', 1200000, 'screen_detect_radix<N=256, window 2>'),
                                            ('c4', 'c4', 300000, 'screen_detect_radix<N=512, window 2>'),
                                            ('c5', 'c5', 80000, 'screen_detect_radix<N=1024, window 1>')):
            p = getattr(configs, factory)(niter=n_real, nchunks=1, seed=1)
            sim = fast_b200.Fast(dict(p))
            sb = dist.StatsBuffers(4096, dev)
            ms = timed_launches(sim, n_real, 3, flush, sb)
            entry(name, n_real, ms, sim.Npxls, bool(p['COHERENT']), kern)
            simf = fast_b200.Fast(dict(p, RNG='device-fast'))
            msf = timed_launches(simf, n_real, 3, flush, sb)
            entry(name + '_device_fast', n_real, msf, sim.Npxls, bool(p['COHERENT']), kern + ', RNG device-fast',
                  note="opt-in RNG='device-fast' (Philox4x32-7, 40 bits per complex sample)")
            del simf
            del sim
    # C3: 16 elevations x 1e4 realisations as ONE batched launch, (elevation x pair) sharded over the ranks
    ps = [configs.c3_elevation(e, niter=10000, nchunks=1, seed=100 + i) for i, e in enumerate(configs.C3_ELEVATIONS)]
    t_build = time.perf_counter()
    sims = sweep.build_sims(ps)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    sweep.run_sweep(sims)                                  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    runs = []
    for rep in range(3):
        flush.fill_(3 + rep)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        res = sweep.run_sweep(sims, stats=True)
        e1.record()
        e1.synchronize()
        runs.append((sync_max(e0.elapsed_time(e1)), time.perf_counter() - t0))
    runs.sort()
    ms, wall = runs[1]
    if rank == 0:
        entry('c3_sweep', 160000, ms, 256, False, 'screen_detect_radix<N=256, window 2>, 16-item batch',
              note="device time of sweep.run_sweep: stack weights, ONE K2 launch over 16 x 5000 pairs, statistics "
                   "all-gather, result all-gather, D2H; per-elevation mean dB_rel in mean_db", gpus=world)
        out['c3_sweep']['wall_ms'] = 1e3 * wall
        out['c3_sweep']['build_ms_per_sample'] = 1e3 * t_build / len(sims)    # Fast(p): host scalars + K1 on the device
        out['c3_sweep']['mean_db'] = [round(float(r.dB_rel.mean()), 3) for r in res]
    # C5 as BASELINE.json states it: 1e6 realisations in total, strong-scaled over the ranks, one collective
    total = 1000000
    p = configs.c5(niter=total, nchunks=1, seed=1)
    sim = fast_b200.Fast(dict(p))
    sb = dist.StatsBuffers(4096, dev)
    lo, hi = dist.shard_range(total // 2, rank, world)
    sim.screen_detect(lo, min(hi - lo, 2000), stats=sb)
    sb.reset()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    flush.fill_(5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sim.screen_detect(lo, hi - lo, stats=sb)
    sb.allreduce()
    e1.record()
    e1.synchronize()
    ms = sync_max(e0.elapsed_time(e1))
    if rank == 0:
        entry('c5_strong', total, ms, 1024, False, 'screen_detect_radix<N=1024, window 1>',
              note=f"1e6 realizations in total sharded over {world} GPU(s), fused statistics + one collective", gpus=world)
        st = sb.summary()
        out['c5_strong']['n_reduced'] = st['n']
        out['c5_strong']['mean_db'] = st['mean_dB']
    return out


def sharded_check(world, rank, dev):
    """N > 1: Fast.run() sharded over the ranks (+ all-gather) against the same global pair range
    computed unsharded on every rank: must be bit-identical (tests/multi_gpu_check.py, run here too so
    that the driver's scaling runs carry the evidence)."""
    import numpy as np
    import torch
    import torch.distributed as td
    import fast_b200
    from fast_b200 import configs, dist
    ok = True
    for factory, kw in (('mini', dict(niter=2000, nchunks=4, seed=4)),
                        ('mini', dict(niter=2002, nchunks=1, seed=4, COHERENT=True)),
                        ('c2', dict(niter=6000, nchunks=3, seed=8)),
                        ('c1prime', dict(niter=600, nchunks=2, seed=9))):
        p = getattr(configs, factory)(**kw)
        sim = fast_b200.Fast(dict(p))
        r_sharded = sim.run()._r
        ppc = sim.Niter_per_chunk // 2
        a, b = sim.screen_detect(0, sim.Nchunks * ppc)
        r_single = dist.assemble(a, b, sim.Nchunks, ppc).cpu().numpy()
        ok = ok and bool(np.array_equal(r_sharded, r_single.astype(r_sharded.dtype)))
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    td.all_reduce(t, op=td.ReduceOp.MIN)
    return bool(t.item())


def run_ours(args):
    import torch
    import torch.distributed as td
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group('nccl', device_id=torch.device('cuda', local))

    import fast_b200
    from fast_b200 import _lib, dist
    from fast_b200 import configs

    factory, n_real, _, _, desc = WORKLOADS[args.workload]
    # the public object: NITER = the whole job of one step (world x n_real), sharded by Fast.run()
    p = getattr(configs, factory)(niter=world * n_real, nchunks=1, seed=1)
    sim = fast_b200.Fast(dict(p))
    N, P = sim.Npxls, sim.Npxls_pup
    coherent = bool(p['COHERENT'])
    n_pairs = n_real // 2
    dev = sim.device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    nbins = 4096
    sb = dist.StatsBuffers(nbins, dev)

    def step(i, k2_events=None):
        # rank r, step i -> its own range of global pair indices (no overlap across ranks/steps)
        first = (i * world + rank) * n_pairs
        sb.reset()
        if k2_events is not None:
            k2_events[0].record()
        sim.screen_detect(first, n_pairs, stats=sb)          # ONE launch: K2 with the statistics fused
        if k2_events is not None:
            k2_events[1].record()
        sb.allreduce()                                       # ONE collective (no-op on a single rank)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()

    # K1 (once-per-config PSD build): wall time of Fast.compute_powerspec -- pupil-filter DFT,
    # fused PSD kernel, Simpson reductions and the small host<->device copies around them
    k1_ms = []
    for _ in range(3):
        t0 = time.perf_counter()
        sim.compute_powerspec()
        torch.cuda.synchronize()
        k1_ms.append(1e3 * (time.perf_counter() - t0))
    step(0)                                                  # tables re-prepared for the rebuilt weight
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    dram = DramSampler(local)
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    dram.start()
    t_region = time.perf_counter()
    step_ms, k2_ms = [], []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                                            # evict L2 between steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(args.warmup + i, (k0, k1))
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        k2_ms.append(k0.elapsed_time(k1))
    torch.cuda.synchronize()
    t_region = time.perf_counter() - t_region
    dram_util = dram.stop()
    if world > 1:
        td.barrier()
    launches = _lib.launch_count()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    stats_last = sb.summary()

    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(total_ms, op=td.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * n_real * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public object: pinned host weight / U in, host result.power out ----
    w_host = sim._d['weight'].cpu().pin_memory()
    u_host = sim._d['U'].cpu().pin_memory()
    width = 2 if coherent else 1
    h2d = w_host.numel() * 4 + u_host.numel() * 4
    # Fast.run() leaves the WHOLE result on every rank: two float64 arrays (result._r and I = result.power,
    # widened on the device) cross to pinned host memory per step
    d2h = world * n_real * width * 8 * 2

    def e2e_step():
        flush.fill_(7)                         # same cold L2 as the device-timed steps (the fill is inside the wall clock)
        sim._d['weight'].copy_(w_host, non_blocking=True)
        sim._d['U'].copy_(u_host, non_blocking=True)
        sim.run()                              # fast/fast.py:115-140 contract: FastResult on the host,
        return sim.I                           # sim.I = result.power (fast/fast.py:137), float64

    for _ in range(3):                         # untimed: the pinned result blocks of the host allocator exist afterwards
        power = e2e_step()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    for _ in range(e2e_steps):
        power = e2e_step()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(e2e_s, op=td.ReduceOp.MAX)
    e2e_value = world * n_real * e2e_steps / float(e2e_s.item())
    e2e_ok = bool(power.shape == (world * n_real,) and (abs(power) > 0).all())

    peak, peak_src = measured_peak()
    sharded_ok = sharded_check(world, rank, dev) if world > 1 else None
    pw = None if args.no_per_workload else per_workload(world, rank, flush, peak, dev)

    if rank == 0:
        b_alg = algorithmic_bytes(N, coherent)
        k2_avg_ms = sum(k2_ms) / len(k2_ms)
        achieved = b_alg * n_real / (k2_avg_ms * 1e-3) / 1e9
        counters, cnote = ncu_counters(args.workload)
        clocks = sampler.summary()
        sm_mhz = clocks['sm_mhz'] or 1965
        secondary = None
        if counters:
            ginst = counters['warp_inst_per_pair'] * n_pairs / (k2_avg_ms * 1e-3) / 1e9
            peak_inst = 148 * 4 * sm_mhz * 1e6 / 1e9
            secondary = {"bound": "issue", "achieved_ginst_s": ginst, "peak_ginst_s": peak_inst,
                         "frac": ginst / peak_inst, "warp_inst_per_pair": counters['warp_inst_per_pair'],
                         "peak_def": f"148 SMs x 4 warp-instructions/clk x {sm_mhz} MHz (clock sampled in this run)"}
        comparator = None
        if world == 1 and not args.no_comparator and not coherent:
            comparator = cufft_comparator(sim, 100000, 3)
        k5 = k5_link_metrics(not args.no_cpu) if world == 1 and not args.no_comparator else None
        cpu = None
        if world == 1 and not args.no_cpu:
            n_cpu = {'c2': 2000, 'c4': 500, 'c5': 120}[args.workload]
            v, dt = cpu_throughput(args.workload, 1, n_cpu)
            k1_cpu = cpu_psd_build_seconds(args.workload)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "psd_build_s": k1_cpu,
                   "sample": f"{n_cpu} realizations of the same workload, 1 process, numpy oracle port "
                             f"of the reference chunk loop ({dt:.1f} s)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(args.workload, world),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak,
                             "traffic": (counters['dram_bytes_per_pair'] * n_pairs) if counters else None,
                             "traffic_source": cnote,
                             "dram_bw_util_pct_gpm": dram_util, "dram_bw_util_note": dram.note or
                             "NVML GPM DRAM_BW_UTIL over the timed region (includes the L2-flush writes between steps)",
                             "peak_source": peak_src,
                             "kernel": "screen_detect_radix", "kernel_ms": k2_avg_ms,
                             "algorithmic_bytes_per_realization": b_alg,
                             "realizations_per_launch": n_real,
                             "secondary": secondary},
                "cpu_baseline": cpu,
                "comparator": comparator,
                "k1_psd_build": {"compute_powerspec_ms": min(k1_ms), "note": "K1 + Simpson + pupil filter incl. host glue; "
                                 "CPU counterpart is cpu_baseline.psd_build_s"},
                "k5_link_metrics": k5,
                "per_workload": pw,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps,
                        "what": "wall clock of: L2 flush, pinned host weight + U -> device, fast_b200.Fast(p).run() "
                                "(table preparation, K2 with the statistics fused, statistics and result all-gather at "
                                "N > 1, reference-order assembly, D2H through pinned memory, float64 host arrays) -> "
                                "result.power"},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "timed_region_s": t_region,
                "check": {"mean_r": stats_last['mean'], "mean_db": stats_last['mean_dB'],
                          "n_reduced": stats_last['n'], "e2e_result_ok": e2e_ok,
                          "sharded_bit_identical": sharded_ok}}
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-comparator', action='store_true', help='skip the cuFFT comparator and K5 legs')
    ap.add_argument('--no-per-workload', action='store_true', help='skip the per_workload block')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
