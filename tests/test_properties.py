"""Property tests (hypothesis, CPU only) of the host logic and of the oracle's combinatorial pieces:
sharding, result ordering, fade-run counting against the reference's own formulation, byte packing."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from fast_b200 import comms, dist
from oracle import comms_oracle as co


@given(total=st.integers(0, 10_000), world=st.integers(1, 64))
def test_shard_ranges_partition_the_pair_range(total, world):
    spans = [dist.shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes


@given(nchunks=st.integers(1, 7), ppc=st.integers(1, 9))
def test_assemble_matches_the_reference_vstack_order(nchunks, ppc):
    """Reference: per chunk, vstack([Re screens, Im screens]) then chunks concatenated
    (fast/funcs.py:220-221, fast/fast.py:134-136)."""
    n = nchunks * ppc
    a = torch.arange(n, dtype=torch.float32)                 # Re-screen result of global pair i
    b = torch.arange(n, dtype=torch.float32) + 1000          # Im-screen result of global pair i
    got = dist.assemble(a, b, nchunks, ppc).numpy()
    want = np.concatenate([np.concatenate([a[c * ppc:(c + 1) * ppc], b[c * ppc:(c + 1) * ppc]]) for c in range(nchunks)])
    np.testing.assert_array_equal(got, want)


def _reference_fade_formulation(mask):
    """fast/comms.py:180-187 restated literally: split at the 0->1 transitions, drop the first chunk,
    keep chunks whose last element is not fading; returns (number kept, their fading samples)."""
    starts = np.where(np.diff(mask.astype(int)) == 1)[0] + 1
    chunks = np.array_split(mask, starts)[1:]
    kept = [c for c in chunks if len(c) and c[-1] != True]  # noqa: E712
    return len(kept), int(sum(c.sum() for c in kept))


@settings(max_examples=300)
@given(bits=st.lists(st.booleans(), min_size=1, max_size=200))
def test_fade_run_counting_equals_the_reference_formulation(bits):
    mask = np.array(bits)
    series = np.where(mask, 0.0, 1.0)
    below, fades, inside = co.fade_counts(series, 0.5)
    assert below == int(mask.sum())
    assert (fades, inside) == _reference_fade_formulation(mask)


@given(data=st.binary(min_size=1, max_size=64), bps=st.sampled_from([1, 2, 3, 4, 5, 6, 7, 8]))
def test_encode_decode_round_trip_for_every_symbol_width(data, bps):
    sym, pad = comms._encode(data, bps)
    assert 0 <= pad < bps and sym.max(initial=0) < 2 ** bps
    out = comms._decode(np.asarray(sym, dtype=np.uint8), bps, pad)
    out = out.tobytes() if isinstance(out, np.ndarray) else out
    assert out == data


@given(M=st.sampled_from([4, 16, 64, 256]))
def test_gray_map_is_a_permutation_with_unit_hamming_steps(M):
    g = co.gray_map_qam(M)
    assert sorted(g) == list(range(M))
    side = int(np.sqrt(M))
    grid = g.reshape(side, side)
    assert all(bin(int(a) ^ int(b)).count('1') == 1 for row in grid for a, b in zip(row[:-1], row[1:]))
    assert [int(c, 2) for c in comms._bin2gray_qam(M)] == list(g)
