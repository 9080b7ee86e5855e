"""Parameter sweeps over independent configurations, e.g. the elevation samples of a LEO pass
that the reference builds one `Fast` object at a time in `FAST_sat_orbit`
(fast/complete_orbit_simulation.py:204-232) and the user then runs in a Python loop.

`run_sweep` keeps that contract (one `Fast` per sample, same keys: L_SAT, DTHETA, ANISO_DL,
ZENITH_ANGLE, AZIMUT_SAT, ...) but batches the device work: every sample's PSD is built up front,
all screen+detect launches are enqueued back to back on the stream, and the host synchronises
once at the end instead of once per sample."""
import numpy
import torch

from . import dist
from .fast import Fast, FastResult


def build_sims(param_dicts):
    """One Fast per configuration (PSD built on the device for each: K1 is ~0.2 ms at 256^2)."""
    return [Fast(p) for p in param_dicts]


def run_sweep(sims):
    """Run every sim (non-temporal, RNG='device') with a single host synchronisation.
    Returns the list of FastResult, also stored on each sim as .result / .I."""
    pending = []
    for sim in sims:
        if sim.temporal or sim.rng_mode != 'device':
            sim.run()
            pending.append(None)
            continue
        sim.compute_logamp()
        ppc = sim.Niter_per_chunk // 2
        a, b = sim.screen_detect(0, sim.Nchunks * ppc)
        pending.append(dist.assemble(a, b, sim.Nchunks, ppc))
    torch.cuda.synchronize()
    out = []
    for sim, flat in zip(sims, pending):
        if flat is not None:
            sim._d['result'] = flat
            I = flat.cpu().numpy()
            I = I.astype(complex) if sim.params['COHERENT'] else I.astype(float)
            sim.result = FastResult(I, sim.diffraction_limit)
            sim.I = sim.result.power
        out.append(sim.result)
    return out


def summary_table(sims, keys=('ZENITH_ANGLE',)):
    """Per-sample mean / scintillation summary as a structured array (for quick inspection)."""
    rows = []
    for s in sims:
        r = numpy.abs(s.result._r) ** 2 if numpy.iscomplexobj(s.result._r) else s.result._r
        rows.append(tuple(s.params[k] for k in keys) + (10 * numpy.log10(r.mean()), (r / r.mean()).var()))
    return rows
