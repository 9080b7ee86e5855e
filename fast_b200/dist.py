"""Multi-GPU plumbing: one process per GPU under torch.distributed (NCCL on the GPU box, gloo in
the CPU tests).  Realisation pairs are independent given the PSD (fast/fast.py:130-134), so the
path shards with NO data-path collective: rank g takes a contiguous range of global pair
indices, and because the device RNG is counter-based on the global pair index the results are
identical for any number of ranks.  The only exchange is at the end: either a tiny all-reduce
of moments + histogram (reduced_stats) or an all-gather of the per-realisation scalars
(gather_pairs) when the caller wants the whole `result.power` array on every rank."""
import torch
import torch.distributed as td


def rank_world():
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard_range(total, rank, world):
    """Balanced contiguous split of range(total): sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_pairs(a, b, total, world):
    """All-gather the per-rank result vectors into full-length ones (global pair order): ONE collective over
    equal-sized (padded) blocks, then the padding is dropped."""
    if world == 1:
        return a, b
    is_c = a.is_complex()
    if is_c:
        a, b = torch.view_as_real(a), torch.view_as_real(b)
    width = tuple(a.shape[1:])
    longest = shard_range(total, 0, world)[1]
    loc = torch.zeros((2, longest, *width), dtype=a.dtype, device=a.device)
    loc[0, :a.shape[0]] = a
    loc[1, :b.shape[0]] = b
    everything = torch.empty((world, 2, longest, *width), dtype=a.dtype, device=a.device)
    td.all_gather_into_tensor(everything.view(-1), loc.view(-1))
    sizes = [shard_range(total, r, world) for r in range(world)]
    if all(hi - lo == longest for lo, hi in sizes):
        fa = everything[:, 0].reshape(world * longest, *width)
        fb = everything[:, 1].reshape(world * longest, *width)
    else:
        fa = torch.cat([everything[r, 0, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
        fb = torch.cat([everything[r, 1, :hi - lo] for r, (lo, hi) in enumerate(sizes)])
    if is_c:
        fa, fb = torch.view_as_complex(fa.contiguous()), torch.view_as_complex(fb.contiguous())
    return fa, fb


def assemble(a, b, nchunks, ppc):
    """Global pair order -> the reference's realisation order: chunk-major, and inside a chunk
    the Re-screen results followed by the Im-screen results (fast/funcs.py:220-221 vstack;
    fast/fast.py:134-136 flatten)."""
    return torch.stack([a.reshape(nchunks, ppc), b.reshape(nchunks, ppc)], dim=1).reshape(-1)


class StatsBuffers:
    """Moments / extrema / dB histogram of `n_items` result sets in ONE contiguous buffer of 8-byte
    words, so that combining ranks is a single small collective:
        [ sums: n_items x 8 f64 | minmax: n_items x 2 f64 | hist: n_items x (nbins + 2) i64 ]
    The views `sums`, `minmax`, `hist` are what fastb_stats / the fused K2 epilogue accumulate into."""

    def __init__(self, nbins, device, n_items=1, db_lo=-60.0, db_hi=3.0):
        self.nbins, self.n_items, self.db_lo, self.db_hi = int(nbins), int(n_items), float(db_lo), float(db_hi)
        a, b = self.n_items * 8, self.n_items * 10
        words = b + self.n_items * (self.nbins + 2)
        self.raw = torch.zeros(words, dtype=torch.int64, device=device)
        self.sums = self.raw[:a].view(torch.float64).view(self.n_items, 8)
        self.minmax = self.raw[a:b].view(torch.float64).view(self.n_items, 2)
        self.hist = self.raw[b:].view(self.n_items, self.nbins + 2)
        self.minmax[:, 0] = float('inf')
        self.minmax[:, 1] = float('-inf')
        self._init = self.raw.clone()
        self._gathered = None

    def reset(self):
        self.raw.copy_(self._init)                 # one copy kernel

    def allreduce(self):
        """Combine the ranks in place with ONE collective: all-gather of the raw buffer (33 KB at
        4096 bins) followed by a local reduction -- SUM for moments and histogram, MIN / MAX for the
        extrema.  Latency-bound over NVLink; no-op on a single rank."""
        _, world = rank_world()
        if world == 1:
            return self
        if self._gathered is None or self._gathered.shape[0] != world:
            self._gathered = torch.empty((world, self.raw.numel()), dtype=torch.int64, device=self.raw.device)
        td.all_gather_into_tensor(self._gathered.view(-1), self.raw)
        a, b = self.n_items * 8, self.n_items * 10
        g = self._gathered
        self.sums.copy_(g[:, :a].view(torch.float64).sum(0).view(self.n_items, 8))
        mm = g[:, a:b].view(torch.float64).view(world, self.n_items, 2)
        self.minmax[:, 0] = mm[:, :, 0].min(0).values
        self.minmax[:, 1] = mm[:, :, 1].max(0).values
        self.hist.copy_(g[:, b:].sum(0).view(self.n_items, self.nbins + 2))
        return self

    def summary(self, item=0):
        return summarise(self.sums[item], self.minmax[item], self.hist[item], self.db_lo, self.db_hi)


def broadcast_seed(seed, device):
    """Rank 0's 64-bit seed on every rank (identity without torch.distributed)."""
    _, world = rank_world()
    if world == 1:
        return int(seed)
    dev = device if td.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([int(seed) - (1 << 64) if int(seed) >= (1 << 63) else int(seed)], dtype=torch.int64, device=dev)
    td.broadcast(t, src=0)
    return int(t.item()) & 0xFFFFFFFFFFFFFFFF


def new_stats_buffers(nbins, device):
    sb = StatsBuffers(nbins, device)
    return sb.sums[0], sb.minmax[0], sb.hist[0]


def allreduce_stats(sums, minmax, hist):
    """Combine separately allocated per-rank statistics in place (legacy three-buffer form; the
    bench and Fast use StatsBuffers.allreduce, which needs one collective instead of four)."""
    _, world = rank_world()
    if world > 1:
        td.all_reduce(sums, op=td.ReduceOp.SUM)
        td.all_reduce(hist, op=td.ReduceOp.SUM)
        lo, hi = minmax[0:1].clone(), minmax[1:2].clone()
        td.all_reduce(lo, op=td.ReduceOp.MIN)
        td.all_reduce(hi, op=td.ReduceOp.MAX)
        minmax[0], minmax[1] = lo[0], hi[0]
    return sums, minmax, hist


def summarise(sums, minmax, hist, db_lo, db_hi):
    """Host-side dict from reduced buffers: the FastResult summaries (fast/fast.py:965-983)
    without the per-realisation array."""
    s = sums.cpu().numpy()
    n = s[0]
    if n == 0:                       # nothing accumulated yet
        nan = float('nan')
        return {'n': 0, 'mean': nan, 'var': nan, 'scintillation_index': nan, 'mean_dB': nan, 'var_dB': nan,
                'min': float(minmax[0]), 'max': float(minmax[1]), 'n_nonpositive': 0,
                'hist': hist.cpu().numpy(), 'db_lo': db_lo, 'db_hi': db_hi}
    mean_r = s[1] / n
    var_r = s[2] / n - mean_r ** 2
    n_db = n - s[5]
    mean_db = s[3] / n_db if n_db else float('nan')
    return {'n': int(n), 'mean': mean_r, 'var': var_r, 'scintillation_index': var_r / mean_r ** 2,
            'mean_dB': mean_db, 'var_dB': (s[4] / n_db - mean_db ** 2) if n_db else float('nan'),
            'min': float(minmax[0]), 'max': float(minmax[1]), 'n_nonpositive': int(s[5]),
            'hist': hist.cpu().numpy(), 'db_lo': db_lo, 'db_hi': db_hi}


def reduced_stats(r_local, db_lo=-60.0, db_hi=3.0, nbins=4096, already_global=False):
    """fastb_stats on this rank's results, then the all-reduce (skipped when every rank
    already holds the full array)."""
    from . import _lib
    sb = StatsBuffers(nbins, r_local.device, db_lo=db_lo, db_hi=db_hi)
    _lib.stats(r_local.contiguous(), db_lo, db_hi, nbins, sb.sums[0], sb.minmax[0], sb.hist[0])
    if not already_global:
        sb.allreduce()
    return sb.summary()
