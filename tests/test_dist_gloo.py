"""CPU tests of the multi-GPU host logic with torch.distributed (gloo, world_size 2): shard
ranges, gather + reference-order assembly, and the moments/histogram all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    td.init_process_group('gloo', rank=rank, world_size=world)
    from fast_b200 import dist
    try:
        nchunks, ppc = 3, 7
        total = nchunks * ppc
        lo, hi = dist.shard_range(total, rank, world)
        # fake per-pair results: value encodes (screen, global pair index)
        g = torch.arange(lo, hi, dtype=torch.float32)
        a, b = g + 1000.0, g + 2000.0
        fa, fb = dist.gather_pairs(a, b, total, world)
        flat = dist.assemble(fa, fb, nchunks, ppc)
        # complex variant
        ca, cb = torch.complex(a, -a), torch.complex(b, -b)
        fca, fcb = dist.gather_pairs(ca, cb, total, world)
        cflat = dist.assemble(fca, fcb, nchunks, ppc)
        # statistics all-reduce of per-rank partials
        sums, minmax, hist = dist.new_stats_buffers(8, torch.device('cpu'))
        sums[0], sums[1] = hi - lo, float(g.sum())
        minmax[0], minmax[1] = float(g.min()), float(g.max())
        hist[rank] = 5
        dist.allreduce_stats(sums, minmax, hist)
        # the packed form: 2 items, one collective
        sb = dist.StatsBuffers(8, torch.device('cpu'), n_items=2)
        sb.sums[0, 0], sb.sums[1, 1] = hi - lo, 10.0 * (rank + 1)
        sb.minmax[0, 0], sb.minmax[0, 1] = float(g.min()), float(g.max())
        sb.hist[1, rank] = 7
        sb.allreduce()
        packed = dict(sums=sb.sums.numpy().copy(), minmax=sb.minmax.numpy().copy(), hist=sb.hist.numpy().copy())
        sb.reset()
        packed['reset_ok'] = bool(sb.sums.abs().sum() == 0 and sb.hist.sum() == 0 and torch.isinf(sb.minmax).all())
        seed = dist.broadcast_seed(0xFEDCBA9876543210 + rank, torch.device('cpu'))
        out[rank] = dict(flat=flat.numpy(), cflat=cflat.numpy(), sums=sums.numpy(), minmax=minmax.numpy(),
                         hist=hist.numpy(), rw=dist.rank_world(), packed=packed, seed=seed)
    finally:
        td.destroy_process_group()


def test_shard_ranges_cover_everything():
    from fast_b200 import dist
    for total in (0, 1, 7, 50000, 50001):
        for world in (1, 2, 3, 8):
            r = [dist.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_assemble_is_the_reference_order():
    from fast_b200 import dist
    nchunks, ppc = 3, 4
    g = torch.arange(nchunks * ppc, dtype=torch.float32)
    flat = dist.assemble(g + 1000, g + 2000, nchunks, ppc).numpy()
    want = []
    for c in range(nchunks):          # chunk-major; Re half then Im half (fast/funcs.py:220-221)
        want += [1000 + c * ppc + i for i in range(ppc)] + [2000 + c * ppc + i for i in range(ppc)]
    np.testing.assert_array_equal(flat, want)


def test_world_size_2_gather_and_allreduce():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    total = 21
    g = np.arange(total, dtype=np.float32)
    want = []
    for c in range(3):
        want += list(1000 + g[c * 7:(c + 1) * 7]) + list(2000 + g[c * 7:(c + 1) * 7])
    for rank in range(world):
        o = out[rank]
        assert o['rw'] == (rank, world)
        np.testing.assert_array_equal(o['flat'], want)
        np.testing.assert_array_equal(o['cflat'].real, want)
        np.testing.assert_array_equal(o['cflat'].imag, [-x for x in want])
        assert o['sums'][0] == total and o['sums'][1] == g.sum()
        assert o['minmax'][0] == 0 and o['minmax'][1] == total - 1
        np.testing.assert_array_equal(o['hist'][:2], [5, 5])
        pk = o['packed']
        assert pk['sums'][0, 0] == total and pk['sums'][1, 1] == 30.0
        assert pk['minmax'][0, 0] == 0 and pk['minmax'][0, 1] == total - 1
        assert np.isinf(pk['minmax'][1]).all()
        np.testing.assert_array_equal(pk['hist'][1, :2], [7, 7])
        assert pk['hist'][0].sum() == 0 and pk['reset_ok']
        assert o['seed'] == 0xFEDCBA9876543210          # rank 0's draw on every rank


def test_flattened_item_pair_ranges_shard_like_a_batch():
    """C3: the (sample x pair) range of a batched sweep is split contiguously over the ranks;
    every flattened index maps to exactly one (item, pair) and items may straddle ranks."""
    from fast_b200 import dist
    E, ppi = 16, 5000
    for world in (1, 2, 4, 8, 3):
        seen = np.zeros(E * ppi, dtype=int)
        for rank in range(world):
            lo, hi = dist.shard_range(E * ppi, rank, world)
            q = np.arange(lo, hi)
            item, g = q // ppi, q % ppi
            assert (item * ppi + g == q).all() and item.max(initial=0) < E
            seen[lo:hi] += 1
        assert (seen == 1).all()
