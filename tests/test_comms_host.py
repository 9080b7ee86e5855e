"""Host-side pieces of fast_b200.comms (no GPU): constellations, Gray maps, byte packing, the
no-atmosphere closed forms and argument errors, against tests/golden/comms.npz."""
import numpy as np
import pytest

from conftest import load_golden_arrays
from fast_b200 import comms


@pytest.fixture(scope='module')
def g():
    return load_golden_arrays('comms')


def test_constellations_match_reference(g):
    for s in list(g['mod_schemes']) + ['16-PSK', '4-QAM', '256-QAM']:
        np.testing.assert_allclose(comms.define_constellation(str(s)), g[f'constellation_{s}'], rtol=0, atol=1e-15)
    with pytest.raises(ValueError, match='perfect square'):
        comms.define_constellation('8-QAM')
    with pytest.raises(ValueError, match='not supported'):
        comms.define_constellation('FSK')


def test_gray_maps_match_reference(g):
    for M in (4, 16, 64):
        codes = comms._bin2gray_qam(M)
        assert [int(c, 2) for c in codes] == list(g[f'gray_{M}'])
        assert all(len(c) == int(np.log2(M)) for c in codes)
        # neighbours along a constellation row differ in exactly one bit
        side = int(np.sqrt(M))
        grid = np.array([int(c, 2) for c in codes]).reshape(side, side)
        assert all(bin(a ^ b).count('1') == 1 for row in grid for a, b in zip(row[:-1], row[1:]))
    np.testing.assert_array_equal(comms._bit_at_index(comms._bin2gray_qam(4), 0, 0), [True, True, False, False])


def test_no_atmosphere_closed_forms(g):
    snr = g['snr_db']
    np.testing.assert_allclose([comms.ber_ook(s) for s in snr], g['ber_ook_noatm'], rtol=1e-13)
    np.testing.assert_allclose(comms.ber_ook(snr), g['ber_ook_noatm'], rtol=1e-13)
    for M in (4, 16, 64):
        np.testing.assert_allclose([comms.sep_qam(M, s) for s in snr], g[f'sep_qam_{M}_noatm'], rtol=1e-13)
    assert comms.Q(0.0) == 0.5


@pytest.mark.parametrize('bps', [1, 2, 4, 6])
def test_encode_decode_round_trip(bps):
    msg = b'free-space optical link \x00\xff\x10'
    sym, pad = comms._encode(msg, bps)
    assert sym.max() < 2 ** bps
    out = comms._decode(sym.astype(np.uint8), bps, pad)
    out = out.tobytes() if isinstance(out, np.ndarray) else out
    assert out == msg


def test_flip_bits():
    a = np.arange(64, dtype=np.uint8)
    assert np.array_equal(comms.flip_bits(a, 0.0), a)
    assert np.array_equal(comms.flip_bits(a, 1.0), ~a)
    assert comms.flip_bits('hello', 0.0) == 'hello'
    with pytest.raises(Exception, match='String or numpy array'):
        comms.flip_bits(5, 0.1)


def test_symbol_alphabets():
    assert [comms._n_symbols(s) for s in ('OOK', 'BPSK', 'QPSK', 'QAM', '8-PSK', '64-QAM')] == [2, 2, 4, 4, 8, 64]
    with pytest.raises(ValueError, match='not recognised'):
        comms._n_symbols('FSK')


# ---- orbit-sweep driver (fast_b200/complete_orbit_simulation.py), host geometry ----------------
def test_fov_offsets_small_displacements():
    from fast_b200.complete_orbit_simulation import fov_offsets
    alt, az, d = np.radians(40.0), np.radians(120.0), np.radians(0.01)
    dx, dy = fov_offsets(alt, az, alt + d, az)               # higher by 0.01 deg: straight up in the field
    assert abs(dx) < 1e-9 and dy == pytest.approx(0.01, rel=1e-6)
    dx, dy = fov_offsets(alt, az, alt - d, az)
    assert abs(dx) < 1e-9 and dy == pytest.approx(-0.01, rel=1e-6)
    dx, dy = fov_offsets(alt, az, alt, az + d)               # azimuth step: shrinks with cos(altitude)
    assert dx == pytest.approx(0.01 * np.cos(alt), rel=1e-4) and abs(dy) < 1e-5
    dx, dy = fov_offsets(alt, az, alt, az - d)
    assert dx == pytest.approx(-0.01 * np.cos(alt), rel=1e-4)
    # separation is preserved: dx^2 + dy^2 = great-circle distance^2
    a1, z1 = alt + 0.7 * d, az + 1.3 * d
    dx, dy = fov_offsets(alt, az, a1, z1)
    sep = np.degrees(np.arccos(np.sin(alt) * np.sin(a1) + np.cos(alt) * np.cos(a1) * np.cos(z1 - az)))
    assert np.hypot(dx, dy) == pytest.approx(sep, rel=1e-6)
    # identical directions: 0/0 in the reference too (it then zeroes the NaN)
    assert all(np.isnan(v) or v == 0 for v in fov_offsets(alt, az, alt, az))


def test_orbit_geometry_needs_skyfield_but_fails_clearly():
    import importlib.util
    from fast_b200 import complete_orbit_simulation as cos
    if importlib.util.find_spec('skyfield') is not None:
        pytest.skip('skyfield is installed')
    with pytest.raises(ImportError, match='skyfield is required'):
        cos.get_satellite_obj('stations.tle')
