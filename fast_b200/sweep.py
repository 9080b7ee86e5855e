"""Parameter sweeps over independent configurations, e.g. the elevation samples of a LEO pass
that the reference builds one `Fast` object at a time in `FAST_sat_orbit`
(fast/complete_orbit_simulation.py:204-232) and the user then runs in a Python loop
(fast/complete_orbit_simulation.py:217-228).

`run_sweep` keeps that contract (one `Fast` per sample, same keys: L_SAT, DTHETA, ANISO_DL,
ZENITH_ANGLE, AZIMUT_SAT, ...) but runs every group of samples that share the grid and the pupil
as ONE batched launch (fastb_screen_detect_batch): the per-sample weight tables are stacked
(E, N, N), sigma_chi and the seeds are per-sample vectors, U is shared, and the flattened
(sample x pair) range is sharded over the ranks of torch.distributed.  Each sample's values are
bit-identical to running that `Fast` on its own."""
import math

import numpy
import torch

from . import _lib
from . import dist
from .fast import Fast, FastResult


def build_sims(param_dicts):
    """One Fast per configuration (PSD built on the device for each: K1 is ~0.2 ms at 256^2)."""
    return [Fast(dict(p)) for p in param_dicts]


def _group_key(sim):
    if sim.temporal or sim.rng_mode == 'numpy' or sim.subharmonics:
        return None
    U = sim._d['U']
    # the digest taken at construction stands while the device copy has not been written to since
    digest = sim._u_digest if U._version == sim._u_version else hash(U.cpu().numpy().tobytes())
    return (str(sim.device), sim.Npxls, sim.Npxls_pup, sim._lo, sim.Niter, sim.Nchunks,
            bool(sim.params['COHERENT']), sim.rng_mode, sim._u_sum, digest)


def run_batch(sims, stats=None):
    """One launch over the (sample x pair) range of sims that share grid, crop, U and run length.
    Returns the list of FastResult; `stats` (dist.StatsBuffers with n_items = len(sims)) receives
    the per-sample moments / histogram, all-reduced over the ranks."""
    lead = sims[0]
    E = len(sims)
    ppc = lead.Niter_per_chunk // 2
    ppi = lead.Nchunks * ppc
    for sim in sims:
        sim._run_index = sim._runs
        sim._runs += 1
        sim.compute_logamp()
    with torch.cuda.device(lead.device):
        dev = lead.device
        weights = torch.stack([sim._d['weight'] for sim in sims]).contiguous()
        sigma = torch.tensor([math.sqrt(sim.logamp_var) for sim in sims], dtype=torch.float32, device=dev)
        seeds = torch.from_numpy(numpy.array([sim._run_seed() for sim in sims], dtype=numpy.uint64).view(numpy.int64)).to(dev)
        rank, world = dist.rank_world()
        lo, hi = dist.shard_range(E * ppi, rank, world)
        rp = lead._run_params(hi - lo, lo)
        width = 2 if rp.coherent else 1
        out_a = torch.empty((hi - lo) * width, dtype=torch.float32, device=dev)
        out_b = torch.empty((hi - lo) * width, dtype=torch.float32, device=dev)
        if hi > lo:
            ws = torch.empty(_lib.screen_detect_workspace_bytes(rp, E), dtype=torch.uint8, device=dev)
            st = None
            if stats is not None:
                st = _lib.run_stats(stats.db_lo, stats.db_hi, stats.nbins, stats.sums, stats.minmax, stats.hist)
            _lib.screen_detect_batch(rp, weights, lead._d['U'], out_a, out_b, ws,
                                     batch=dict(n_items=E, pairs_per_item=ppi, sigma_chi=sigma, seeds=seeds), stats=st)
        if stats is not None:
            stats.allreduce()
        if rp.coherent:
            out_a = torch.view_as_complex(out_a.view(-1, 2))
            out_b = torch.view_as_complex(out_b.view(-1, 2))
        a, b = dist.gather_pairs(out_a, out_b, E * ppi, world)
        # per sample: the reference's order (chunk-major, Re half then Im half), then one D2H copy
        # (dist.assemble for every item at once: [item][chunk][Re half | Im half][pair])
        flat = torch.stack([a.reshape(E, lead.Nchunks, ppc), b.reshape(E, lead.Nchunks, ppc)], dim=2).reshape(E, -1)
        wide = flat.to(torch.complex128 if flat.is_complex() else torch.float64)
        dls = torch.tensor([sim.diffraction_limit for sim in sims], dtype=torch.float64, device=dev)
        host = lead._to_host(wide.reshape(-1)).reshape(E, -1)
        host_i = lead._to_host((wide * dls[:, None]).reshape(-1)).reshape(E, -1)
    out = []
    for e, sim in enumerate(sims):
        sim._d['result'] = flat[e]
        sim.result = FastResult(host[e], sim.diffraction_limit)
        sim.I = host_i[e]
        out.append(sim.result)
    return out


def run_sweep(sims, stats=False, db_lo=-60.0, db_hi=3.0, nbins=4096):
    """Run every sim; samples that share grid, crop, pupil and run length go through one batched
    launch per group.  Returns the list of FastResult in the order of `sims`, also stored on each
    sim as .result / .I.  stats=True additionally leaves the fused per-sample statistics
    (dist.summarise dict) on each batched sim as .stats."""
    groups, order = {}, []
    for i, sim in enumerate(sims):
        key = _group_key(sim)
        if key is None:
            order.append([i])
        elif key in groups:
            groups[key].append(i)
        else:
            groups[key] = [i]
            order.append(groups[key])
    results = [None] * len(sims)
    for idxs in order:
        members = [sims[i] for i in idxs]
        if len(members) == 1 and _group_key(members[0]) is None:
            results[idxs[0]] = members[0].run()
            continue
        sb = dist.StatsBuffers(nbins, members[0].device, n_items=len(members), db_lo=db_lo, db_hi=db_hi) if stats else None
        for i, r in zip(idxs, run_batch(members, sb)):
            results[i] = r
        if sb is not None:
            for e, sim in enumerate(members):
                sim.stats = sb.summary(e)
    return results


def summary_table(sims, keys=('ZENITH_ANGLE',)):
    """Per-sample mean / scintillation summary rows (for quick inspection).  Geometry keys are read
    from each sim's own state, not from `params` (callers may share one params dict between sims)."""
    own = {'ZENITH_ANGLE': lambda s: float(numpy.degrees(numpy.arccos(1.0 / s.zenith_correction))),
           'L_SAT': lambda s: s.L, 'DTHETA': lambda s: tuple(s.dtheta)}
    rows = []
    for s in sims:
        r = numpy.abs(s.result._r) ** 2 if numpy.iscomplexobj(s.result._r) else s.result._r
        vals = tuple(own[k](s) if k in own else s.params[k] for k in keys)
        rows.append(vals + (10 * numpy.log10(r.mean()), (r / r.mean()).var()))
    return rows
