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
