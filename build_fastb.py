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
LIB = os.path.join(HERE, 'libfastb.so')
STAMP = os.path.join(HERE, 'build', 'libfastb.stamp')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']
# per-file extra flags: the PSD kernel follows numpy's operation order (no FMA contraction)
SOURCES = {
    'api.cu': [],
    'psd_build.cu': ['-fmad=false'],
    'screen_detect.cu': ['-Xptxas', '-v'] + (['-DFASTB_TUNE'] if os.environ.get('FASTB_TUNE') else []) + (['-DFASTB_TUNE_DBG'] if os.environ.get('FASTB_TUNE_DBG') else []),
    'stats.cu': [],
    'link_metrics.cu': ['-fmad=false'],
    'temporal.cu': [],
}


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(ROOT, 'include')):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), 'rb') as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(repr(SOURCES).encode())
    h.update(os.environ.get('FASTB_TUNE', '').encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every CUDA source to an object and link the shared library.  No-op when the
    sources are unchanged since the last build."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc, *ARCH, *COMMON, *extra, '-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                 stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        with open(os.path.join(objdir, src + '.log'), 'w') as fh:
            fh.write(out)
        if pr.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}: {" ".join(cmd)}')
    subprocess.run([nvcc, *ARCH, '-shared', '-o', LIB, *objs], check=True)
    with open(STAMP, 'w') as fh:
        fh.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
