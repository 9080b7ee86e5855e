"""Launched under torchrun by tests/test_gpu_multi.py: every rank runs the sharded Fast.run()
and rank 0 compares with an unsharded run of the same seed (must be bit-identical), then checks
the all-reduced statistics of per-rank shards against the full array."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    td.init_process_group('nccl', device_id=torch.device('cuda', local))
    import fast_b200
    from fast_b200 import configs, dist
    rank, world = dist.rank_world()
    for factory, kw in (('mini', dict(niter=2000, nchunks=4, seed=4)),
                        ('mini', dict(niter=2002, nchunks=1, seed=4, COHERENT=True)),
                        ('c2', dict(niter=600, nchunks=3, seed=8))):
        p = getattr(configs, factory)(**kw)
        sim = fast_b200.Fast(dict(p))
        r_sharded = sim.run()._r                       # sharded over the ranks + all-gather
        # unsharded on this rank: same global pair indices, one launch
        ppc = sim.Niter_per_chunk // 2
        a, b = sim.screen_detect(0, sim.Nchunks * ppc)
        r_single = dist.assemble(a, b, sim.Nchunks, ppc).cpu().numpy()
        assert np.array_equal(r_sharded, r_single.astype(r_sharded.dtype)), (factory, rank)
        # per-rank shard statistics + all-reduce == statistics of the whole
        lo, hi = dist.shard_range(sim.Nchunks * ppc, rank, world)
        mine = torch.cat([a[lo:hi], b[lo:hi]])
        if mine.is_complex():
            mine = (mine.real ** 2 + mine.imag ** 2)
        st = dist.reduced_stats(mine.contiguous(), -60, 3, 630)
        full = np.abs(r_single) ** 2 if np.iscomplexobj(r_single) else r_single
        assert st['n'] == full.size
        assert abs(st['mean'] - full.mean()) < 1e-6 * full.mean()
        assert abs(st['min'] - full.min()) <= 1e-6 * full.max() and abs(st['max'] - full.max()) <= 1e-6 * full.max()
        assert int(st['hist'].sum()) == full.size
    td.barrier()
    if rank == 0:
        print(f'MULTI_GPU_OK world={world}')
    td.destroy_process_group()


if __name__ == '__main__':
    main()
