"""Small integer helpers for headers: BCD and CRC.

BCD: baseband/base/utils.py:18-49.  CRC: plain polynomial division over
GF(2), used when *writing* Mark 5B (CRC-16, 0x18005) and Mark 4 (CRC-12,
0x180f) headers (baseband/base/utils.py:93-248).  Host side, a few words per
frame; not part of the GPU hot path.
"""
import numpy as np

__all__ = ['bcd_decode', 'bcd_encode', 'crc_remainder', 'crc_array',
           'crc_of_bits', 'lcm']


def lcm(a, b):
    import math
    return abs(a * b) // math.gcd(a, b)


def bcd_decode(value):
    """Binary-coded decimal -> int (scalar or integer array)."""
    if isinstance(value, np.ndarray):
        v = value.astype(np.int64)
        ndigit = 2 * value.dtype.itemsize
        nib = (v[..., None] >> (4 * np.arange(ndigit))) & 0xf
        if (nib > 9).any():
            raise ValueError('invalid BCD encoded value')
        return (nib * 10 ** np.arange(ndigit)).sum(-1)
    return int(format(int(value), 'x'))


def bcd_encode(value):
    if isinstance(value, np.ndarray):
        v = value.astype(np.int64)
        ndigit = 16
        digits = (v[..., None] // 10 ** np.arange(ndigit)) % 10
        return (digits << (4 * np.arange(ndigit))).sum(-1)
    return int(str(int(value)), 16)


def crc_remainder(value, polynomial, extend=True):
    """Remainder of ``value`` (an arbitrary-size int holding the message bits,
    most significant first) divided by ``polynomial``; with ``extend`` the
    message is first shifted by the CRC width, i.e. the CRC is *calculated*;
    without it a message that already ends in its CRC gives 0."""
    npol = polynomial.bit_length()
    if extend:
        value <<= npol - 1
    nbit = value.bit_length()
    while nbit >= npol:
        value ^= polynomial << (nbit - npol)
        nbit = value.bit_length()
    return value


def crc_array(values, nbits, polynomial):
    """CRC of every ``nbits``-bit message in an integer array (vectorised
    long division; nbits + CRC width must fit in 64 bits)."""
    ncrc = polynomial.bit_length() - 1
    assert nbits + ncrc <= 64
    work = np.asarray(values).astype(np.uint64) << np.uint64(ncrc)
    pol = np.uint64(polynomial)
    for bit in range(nbits + ncrc - 1, ncrc - 1, -1):
        top = (work >> np.uint64(bit)) & np.uint64(1)
        work ^= (pol << np.uint64(bit - ncrc)) * top
    return work.astype(np.int64)


def crc_of_bits(stream, polynomial):
    """CRC for parallel bit streams: ``stream[i]`` holds bit i of every
    stream (one per bit level of the integers).  Returns the ncrc words to
    append (baseband/base/utils.py:200-248 semantics)."""
    ncrc = polynomial.bit_length() - 1
    work = np.concatenate([stream, np.zeros((ncrc,) + stream.shape[1:],
                                            stream.dtype)])
    pol_bits = [(polynomial >> (ncrc - k)) & 1 for k in range(ncrc + 1)]
    for i in range(len(stream)):
        bits = work[i].copy()
        for k, pb in enumerate(pol_bits):
            if pb:
                work[i + k] ^= bits
    return work[-ncrc:]
