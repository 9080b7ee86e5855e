"""Builds fast_b200/libfastb.so in-tree with nvcc for sm_100a (B200).

    python build_fastb.py [--force]

Lives outside the package on purpose: importing `fast_b200` loads the shared library (and fails
loudly when it is missing or stale), so the build must not depend on that import.
The .so is git-ignored but ships to the GPU box with the working tree."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, 'fast_b200')
CSRC = os.path.join(HERE, 'csrc')
# tuning builds (FASTB_TUNE=1) go to their own library so that the product .so is never replaced by one
TUNE = bool(os.environ.get('FASTB_TUNE'))
LIB = os.path.join(HERE, ('libfastb_tune%s.so' % os.environ.get('FASTB_TUNE_TAG', '')) if TUNE else 'libfastb.so')
STAMP = os.path.join(HERE, 'build', ('libfastb_tune%s.stamp' % os.environ.get('FASTB_TUNE_TAG', '')) if TUNE else 'libfastb.stamp')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']
# per-file extra flags: the PSD kernel follows numpy's operation order (no FMA contraction)
_DBG = ['-DFASTB_TUNE_DBG'] if os.environ.get('FASTB_TUNE_DBG') else []
if os.environ.get('FASTB_SPLIT'):              # tuning builds: FASTB_SPLIT=0 -> LineFFT for N >= 512
    _DBG += ['-DFASTB_SPLIT=' + os.environ['FASTB_SPLIT']]
if os.environ.get('FASTB_PF'):                 # tuning builds: pass-2 column prefetch through L1 (screen_detect_kernel.cuh)
    _DBG += ['-DFASTB_PF=' + os.environ['FASTB_PF']]
if os.environ.get('FASTB_SCALAR_STAGES'):      # tuning builds: which FFT stages use scalar FP32 (fft_core.cuh)
    _DBG += ['-DFASTB_SCALAR_STAGES=' + os.environ['FASTB_SCALAR_STAGES']]
# object name -> (source, extra flags).  The radix kernels of K2 are compiled once per grid size
# (N = 2^6 .. 2^11) so that the ~20 instances of each size build in parallel.
SOURCES = {
    'api': ('api.cu', []),
    'psd_build': ('psd_build.cu', ['-fmad=false']),
    'screen_detect': ('screen_detect.cu', ['-Xptxas', '-v'] + _DBG),
    'screen_detect_bluestein': ('screen_detect_bluestein.cu', []),
    'stats': ('stats.cu', []),
    'link_metrics': ('link_metrics.cu', ['-fmad=false']),
    'temporal': ('temporal.cu', []),
    'layer_screens_fft': ('layer_screens_fft.cu', ['-Xptxas', '-v']),
}
_BLUE = ['-DFASTB_BLUE_TWO=' + os.environ['FASTB_BLUE_TWO']] if os.environ.get('FASTB_BLUE_TWO') else []
for _k in range(6, 12):
    SOURCES[f'screen_detect_radix_{_k}'] = ('screen_detect_radix.cu', ['-Xptxas', '-v', f'-DFASTB_LOG2N={_k}'] + _DBG)
    # the chirp-z kernels likewise once per transform length M = 2^k
    SOURCES[f'screen_detect_bluestein_{_k}'] = ('screen_detect_bluestein_m.cu',
                                                ['-Xptxas', '-v', f'-DFASTB_LOG2M={_k}'] + _DBG + _BLUE)
if os.environ.get('FASTB_TUNE'):
    # tuning builds only: env-driven alternative kernel shapes / variants (never part of the product)
    SOURCES['screen_detect_tune'] = (os.path.join('tune', 'screen_detect_tune.cu'), ['-Xptxas', '-v'] + _DBG)


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(CSRC, 'tune'), os.path.join(ROOT, 'include')):
        for f in sorted(os.listdir(root)):
            if os.path.isdir(os.path.join(root, f)):
                continue
            with open(os.path.join(root, f), 'rb') as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(repr(SOURCES).encode())
    h.update(os.environ.get('FASTB_TUNE', '').encode())
    h.update(os.environ.get('FASTB_TUNE_DBG', '').encode())
    h.update(os.environ.get('FASTB_SCALAR_STAGES', '').encode())
    h.update(os.environ.get('FASTB_SPLIT', '').encode())
    h.update(os.environ.get('FASTB_BLUE_TWO', '').encode())
    h.update(os.environ.get('FASTB_PF', '').encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every CUDA source to an object and link the shared library.  No-op when the
    sources are unchanged since the last build."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objdir = os.path.join(HERE, 'build', ('tune' + os.environ.get('FASTB_TUNE_TAG', '')) if TUNE else 'product')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for name, (src, extra) in SOURCES.items():
        obj = os.path.join(objdir, name + '.o')
        cmd = [nvcc, *ARCH, *COMMON, *extra, '-c', os.path.join(CSRC, src), '-o', obj]
        log = open(os.path.join(objdir, name + '.log'), 'w')
        procs.append((name, cmd, log, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = None
    for name, cmd, log, pr in procs:
        pr.wait()
        log.close()
        out = open(log.name).read()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0 and failed is None:
            failed = f'nvcc failed for {name}: {" ".join(cmd)}'
    if failed:
        raise RuntimeError(failed)
    subprocess.run([nvcc, *ARCH, '-shared', '-o', LIB, *objs], check=True)
    with open(STAMP, 'w') as fh:
        fh.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
