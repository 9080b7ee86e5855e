"""Benchmark of the FAST Monte-Carlo hot path (screen generation + detection, PSD already built).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5]

One "step" = one pass of the hot path over one batch of the named workload on every rank:
K2 (noise -> screens -> detector, fastb_screen_detect) + K3 (fastb_stats) + the all-reduce of
moments/histogram (N > 1).  Weak scaling: every rank processes a full batch of its own
realisation range.  Prints ONE JSON line on rank 0 (contract: see the task statement / DESIGN.md).

--impl reference times the reference's CPU algorithm (the numpy oracle port of
fast/fast.py:115-140,589-668: /root/reference is pure Python and cannot travel to the GPU box)
on all host cores, on bounded samples of the same workload.
"""
import argparse
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

WORKLOADS = {
    # name: (config factory, realisations per step per GPU, description)
    'c2': ('c2', 100000, "C2: GEO ground-station downlink, 256x256 grid, 1e5 realizations per step per GPU, "
                         "AO residual PSD + scintillation, SMF detection"),
    'c4': ('c4', 100000, "C4: coherent detection, 512x512 grid, 1e5 realizations per step per GPU"),
    'c5': ('c5', 20000, "C5 shard: 1024x1024 grid, 2e4 realizations per step per GPU"),
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
    factory, _, _ = WORKLOADS[workload]
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
    factory, _, _ = WORKLOADS[workload]
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
    factory, n_real, desc = WORKLOADS[args.workload]
    cores = min(os.cpu_count() or 1, 128)
    rate_guess = {'c2': 200.0, 'c4': 45.0, 'c5': 10.0}[args.workload]
    budget = 150.0 / max(1, args.steps + args.warmup)              # seconds per step
    per_proc = max(2, int(budget * rate_guess * 0.6) // 2 * 2)
    value, dt = cpu_throughput(args.workload, cores, per_proc, steps=args.steps, warmup=args.warmup)
    sample = f"{cores} processes x {per_proc} realizations per step (numpy oracle port of the reference chunk loop)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": desc, "sample": sample},
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

    factory, n_real, desc = WORKLOADS[args.workload]
    p = getattr(configs, factory)(niter=n_real, nchunks=1, seed=1)
    sim = fast_b200.Fast(dict(p))
    N, P = sim.Npxls, sim.Npxls_pup
    coherent = bool(p['COHERENT'])
    n_pairs = n_real // 2
    dev = sim.device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    nbins = 4096

    def step(i, k2_events=None):
        # rank r, step i -> its own range of global pair indices (no overlap across ranks/steps)
        first = (i * world + rank) * n_pairs
        if k2_events is not None:
            k2_events[0].record()
        a, b = sim.screen_detect(first, n_pairs)
        if k2_events is not None:
            k2_events[1].record()
        r = torch.cat([a, b])
        if r.is_complex():
            r = (r.real ** 2 + r.imag ** 2)
        sums, minmax, hist = dist.new_stats_buffers(nbins, dev)
        _lib.stats(r.contiguous(), -60.0, 3.0, nbins, sums, minmax, hist)
        dist.allreduce_stats(sums, minmax, hist)
        return sums

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

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    launches_torch = 0
    step_ms, k2_ms = [], []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                                            # evict L2 between steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sums = step(args.warmup + i, (k0, k1))
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        k2_ms.append(k0.elapsed_time(k1))
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    launches = _lib.launch_count()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(total_ms, op=td.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * n_real * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public object, host buffers in, host results out ----
    w_host = sim._d['weight'].cpu().pin_memory()
    u_host = sim._d['U'].cpu().pin_memory()
    width = 2 if coherent else 1
    out_host = torch.empty(n_real * width, dtype=torch.float32).pin_memory()
    h2d = w_host.numel() * 4 + u_host.numel() * 4
    d2h = out_host.numel() * 4

    def e2e_step(i):
        sim._d['weight'].copy_(w_host, non_blocking=True)
        sim._d['U'].copy_(u_host, non_blocking=True)
        first = (i * world + rank) * n_pairs
        a, b = sim.screen_detect(first, n_pairs)
        flat = dist.assemble(a, b, 1, n_pairs)
        src = torch.view_as_real(flat).reshape(-1) if flat.is_complex() else flat
        out_host.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step(0)
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(1000 + i)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(e2e_s, op=td.ReduceOp.MAX)
    e2e_value = world * n_real * args.steps / float(e2e_s.item())

    if rank == 0:
        peak, peak_src = measured_peak()
        b_alg = algorithmic_bytes(N, coherent)
        k2_avg_ms = sum(k2_ms) / len(k2_ms)
        achieved = b_alg * n_real / (k2_avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        comparator = None
        if world == 1 and not args.no_comparator and not coherent:
            comparator = cufft_comparator(sim, n_real, max(2, args.steps // 4))
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
                "config": {"workload": desc, "N": N, "n_pup": P, "layers": len(sim.h),
                           "realizations_per_step_per_gpu": n_real, "rng": "device Philox4x32-10 + Box-Muller",
                           "l2": "flushed (256 MiB write) between timed steps",
                           "parallelism": f"realization ranges sharded over {world} GPU(s), "
                                          "moments+histogram all-reduce per step"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "kernel": "screen_detect_radix", "kernel_ms": k2_avg_ms,
                             "algorithmic_bytes_per_realization": b_alg,
                             "realizations_per_launch": n_real},
                "cpu_baseline": cpu,
                "comparator": comparator,
                "k1_psd_build": {"compute_powerspec_ms": min(k1_ms), "note": "K1 + Simpson + pupil filter incl. host glue; "
                                 "CPU counterpart is cpu_baseline.psd_build_s"},
                "k5_link_metrics": k5,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "check": {"mean_r": float(sums[1] / sums[0]), "n_reduced": int(sums[0])}}
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
    ap.add_argument('--no-comparator', action='store_true', help='skip the cuFFT comparator leg')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
