"""Oracle sample codecs (numpy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates, function by function, the arithmetic of

* baseband/base/encoding.py      (levels, generic quantisers, 8-bit codec)
* baseband/vdif/payload.py       (offset-binary 1/2/4-bit LUT codecs)
* baseband/mark5b/payload.py     (sign/magnitude 1/2-bit codecs)
* baseband/mark4/payload.py      (track reorder + five fan-out modes)
* baseband/guppi/payload.py, baseband/dada/payload.py (int8, layouts, MKBF)
* baseband/gsb/payload.py        (signed 4-bit nibbles, int8)

All arithmetic is done by numpy in the same dtype the reference uses, so the
results are bit-identical by construction; the golden tests verify that.
"""
import numpy as np

# --------------------------------------------------------------------------
# Constants and levels: baseband/base/encoding.py:14, :45-60
# --------------------------------------------------------------------------
OPTIMAL_2BIT_HIGH = 3.316505
TWO_BIT_1_SIGMA = 2.174564
FOUR_BIT_1_SIGMA = 2.95
EIGHT_BIT_1_SIGMA = 71.0 / 2.0

LEVELS = {
    1: np.array([-1.0, 1.0], dtype=np.float32),
    2: np.array([-OPTIMAL_2BIT_HIGH, -1.0, 1.0, OPTIMAL_2BIT_HIGH],
                dtype=np.float32),
    4: (np.arange(16, dtype=np.float32) - 8.0) / FOUR_BIT_1_SIGMA,
}

_TWO_SIGMA = 2 * TWO_BIT_1_SIGMA
_CLIP_LO, _CLIP_HI = -1.5 * TWO_BIT_1_SIGMA, 1.5 * TWO_BIT_1_SIGMA


# --------------------------------------------------------------------------
# Unpacked quantisers: baseband/base/encoding.py:63-128, :147-158
# --------------------------------------------------------------------------
def quantise_1bit(values):
    """``values >= 0`` as uint8 (encoding.py:63-74)."""
    out = np.empty(values.shape, np.uint8)
    return np.greater_equal(values, 0.0, out=out, casting='unsafe')


def quantise_2bit(values):
    """clip to +-1.5 sigma, shift by 2 sigma, floor-divide by sigma
    (encoding.py:77-102).  Arithmetic stays in the dtype of ``values``."""
    shifted = np.clip(values, _CLIP_LO, _CLIP_HI)
    shifted += _TWO_SIGMA
    out = np.empty(shifted.shape, np.uint8)
    return np.floor_divide(shifted, TWO_BIT_1_SIGMA, out=out,
                           casting='unsafe')


def quantise_4bit(values):
    """scale by 2.95, add 8.5, clip to [0, 15], truncate
    (encoding.py:105-128)."""
    scaled = values * FOUR_BIT_1_SIGMA
    scaled += 8.5
    return np.clip(scaled, 0.0, 15.0, out=scaled).astype(np.uint8)


def decode_8bit_offset(words):
    """(uint8 - 127.5) / 35.5 in float32 (encoding.py:131-144)."""
    vals = words.view(np.uint8).astype(np.float32)
    vals -= 127.5
    vals /= EIGHT_BIT_1_SIGMA
    return vals


def encode_8bit_offset(values):
    """clip(rint(v * 35.5 + 127.5), 0, 255) (encoding.py:147-158)."""
    return (np.clip(np.rint(values * EIGHT_BIT_1_SIGMA + 127.5), 0, 255)
            .astype(np.uint8))


# --------------------------------------------------------------------------
# VDIF: LSB-first offset binary.  baseband/vdif/payload.py:25-114
# --------------------------------------------------------------------------
def _byte_luts(levels, bps):
    """256-row table: sample i of byte b is levels[(b >> i*bps) & mask]
    (vdif/payload.py:53-63)."""
    byte = np.arange(256)[:, np.newaxis]
    shifts = np.arange(0, 8, bps)
    return levels[(byte >> shifts) & ((1 << bps) - 1)]


VDIF_LUT = {bps: _byte_luts(LEVELS[bps], bps) for bps in (1, 2, 4)}


def vdif_decode(words, bps):
    """Flat float32 decode of VDIF words (vdif/payload.py:69-71, :83-86,
    :100-103; 8 bit: encoding.py:131-144)."""
    if bps == 8:
        return decode_8bit_offset(words)
    return VDIF_LUT[bps].take(words.view(np.uint8), axis=0)


def vdif_encode(values, bps):
    """Packed uint8 from floats (vdif/payload.py:77-80, :92-97, :109-114)."""
    if bps == 1:
        bits = quantise_1bit(values.reshape(-1, 8))
        return np.packbits(bits[:, ::-1])
    if bps == 2:
        codes = quantise_2bit(values.reshape(-1, 4))
        codes <<= np.arange(0, 8, 2).astype(np.uint8)
        return np.bitwise_or.reduce(codes, axis=-1)
    if bps == 4:
        codes = quantise_4bit(values).reshape(-1, 2)
        codes <<= np.array([0, 4], np.uint8)
        return codes[:, 0] | codes[:, 1]
    if bps == 8:
        return encode_8bit_offset(values)
    raise ValueError("cannot encode data with {} bits".format(bps))


# --------------------------------------------------------------------------
# Mark 5B: sign on even bit, magnitude on odd.  mark5b/payload.py:27-106
# --------------------------------------------------------------------------
def _mark5b_luts():
    byte = np.arange(256)[:, np.newaxis]
    every = np.arange(8)
    # set bit -> -1 (mark5b/payload.py:62-65)
    lut1 = LEVELS[1][((byte >> every) & 1) ^ 1]
    sign = np.arange(0, 8, 2)
    mag = sign + 1
    index = (((byte >> sign) & 1) << 1) + ((byte >> mag) & 1)
    return lut1, LEVELS[2][index]


MARK5B_LUT1, MARK5B_LUT2 = _mark5b_luts()


def mark5b_decode(words, bps):
    """mark5b/payload.py:78-94."""
    lut = {1: MARK5B_LUT1, 2: MARK5B_LUT2}[bps]
    return lut.take(words.view(np.uint8), axis=0)


def mark5b_encode(values, bps):
    """mark5b/payload.py:86-106 (1 bit: signbit; 2 bit: codes 1<->2)."""
    if bps == 1:
        bits = np.signbit(values.reshape(-1, 8)).view(np.uint8)
        return np.packbits(bits[:, ::-1])
    if bps == 2:
        codes = quantise_2bit(values.reshape(-1, 4))
        np.array([0, 2, 1, 3], dtype=np.uint8).take(codes, out=codes)
        codes <<= np.arange(0, 8, 2).astype(np.uint8)
        return np.bitwise_or.reduce(codes, axis=-1)
    raise ValueError("cannot encode data with {} bits".format(bps))


MARK5B_FILL_WORD = 0x11223344   # mark5b/frame.py:62


def mark5b_payload_valid(words):
    """Frame is invalid iff all words equal the fill pattern
    (mark5b/frame.py:62-72)."""
    return bool((np.asarray(words) != MARK5B_FILL_WORD).any())


# --------------------------------------------------------------------------
# Mark 4.  mark4/payload.py:48-69 (little-endian branch), :88-115, :122-300
# --------------------------------------------------------------------------
def m4_reorder32(x):
    return ((x & 0xAA55AA55) | ((x & 0x55005500) >> 7)
            | ((x & 0x00AA00AA) << 7))


def m4_reorder64(x):
    return ((x & 0xAA55AA55AA55AA55) | ((x & 0x5500550055005500) >> 7)
            | ((x & 0x00AA00AA00AA00AA) << 7))


def m4_reorder64_ft(x):
    return ((x & 0xFFFFFAAFFFFFFAAF) | ((x & 0x0000050000000500) >> 4)
            | ((x & 0x0000005000000050) << 4))


def _mark4_luts():
    byte = np.arange(256)[:, np.newaxis]
    lut1 = LEVELS[1][((byte >> np.arange(8)) & 1) ^ 1]
    i = np.arange(4)

    def sm_lut(sign, mag):
        return LEVELS[2][2 * (byte >> sign & 1) + (byte >> mag & 1)]

    lut2a = sm_lut(i * 2, i * 2 + 1)               # s=0,2,4,6 m=s+1
    s = i + (i // 2) * 2
    lut2b = sm_lut(s, s + 2)                       # s=0,1,4,5 m=s+2
    lut2c = sm_lut(i, i + 4)                       # s=0..3   m=s+4
    return lut1, lut2a, lut2b, lut2c


M4_LUT1, M4_LUT2_1, M4_LUT2_2, M4_LUT2_3 = _mark4_luts()


def _m4_dec_2ch_f4(words):       # mark4/payload.py:122-135
    b = words.view(np.uint8).reshape(-1, 2)
    return M4_LUT2_3.take(b, axis=0).transpose(1, 0, 2).reshape(2, -1).T


def _m4_dec_4ch_f4(words):       # mark4/payload.py:152-163
    b = m4_reorder32(words.view(np.uint32)).view(np.uint8).reshape(-1, 4)
    b = b.take(np.array([0, 2, 1, 3]), axis=1)
    return M4_LUT2_1.take(b.T, axis=0).reshape(4, -1).T


def _m4_dec_8ch_f2(words):       # mark4/payload.py:178-197
    b = words.view(np.uint8).reshape(-1, 4)
    return (M4_LUT2_3.take(b, axis=0).reshape(-1, 4, 2, 2)
            .transpose(3, 1, 0, 2).reshape(8, -1).T)


def _m4_dec_16ch_f2_ft(words):   # mark4/payload.py:215-256
    b = m4_reorder64_ft(words.view(np.uint64)).view(np.uint8).reshape(-1, 8)
    return (M4_LUT2_3.take(b, axis=0).reshape(-1, 2, 4, 2, 2)
            .transpose(1, 4, 2, 0, 3).reshape(16, -1).T)


def _m4_dec_8ch_f4(words):       # mark4/payload.py:277-288
    b = m4_reorder64(words.view(np.uint64)).view(np.uint8).reshape(-1, 8)
    b = b.take(np.array([0, 2, 1, 3, 4, 6, 5, 7]), axis=1)
    return M4_LUT2_1.take(b.T, axis=0).reshape(8, -1).T


_SM_16 = np.array([0, 16, 1, 17], dtype=np.uint8)   # sign -> bit0, mag -> bit4
_SM_SWAP = np.array([0, 2, 1, 3], dtype=np.uint8)


def _m4_enc_2ch_f4(values):      # mark4/payload.py:138-149
    v = values.reshape(-1, 4, 2).transpose(0, 2, 1)
    codes = quantise_2bit(v)
    _SM_16.take(codes, out=codes)
    codes <<= np.array([0, 1, 2, 3], dtype=np.uint8)
    return np.bitwise_or.reduce(codes, axis=-1).ravel().view('<u2')


def _m4_enc_4ch_f4(values):      # mark4/payload.py:166-175
    v = values[:, np.array([0, 2, 1, 3])].reshape(-1, 4, 4).transpose(0, 2, 1)
    codes = quantise_2bit(v)
    _SM_SWAP.take(codes, out=codes)
    codes <<= np.array([0, 2, 4, 6], dtype=np.uint8)
    out = np.bitwise_or.reduce(codes, axis=-1).ravel().view(np.uint32)
    return m4_reorder32(out).view('<u4')


def _m4_enc_8ch_f2(values):      # mark4/payload.py:200-212
    v = (values.reshape(-1, 2, 2, 4).transpose(0, 3, 1, 2).reshape(-1, 4, 4))
    codes = quantise_2bit(v)
    _SM_16.take(codes, out=codes)
    codes <<= np.array([0, 1, 2, 3], dtype=np.uint8)
    return np.bitwise_or.reduce(codes, axis=-1).ravel().view('<u4')


def _m4_enc_16ch_f2_ft(values):  # mark4/payload.py:259-274
    v = (values.reshape(-1, 2, 2, 2, 4).transpose(0, 2, 4, 1, 3)
         .reshape(-1, 4))
    codes = quantise_2bit(v)
    _SM_16.take(codes, out=codes)
    codes <<= np.array([0, 1, 2, 3], dtype=np.uint8)
    out = np.bitwise_or.reduce(codes, axis=-1).ravel().view(np.uint64)
    return m4_reorder64_ft(out).view('<u8')


def _m4_enc_8ch_f4(values):      # mark4/payload.py:291-300
    order = np.array([0, 2, 1, 3, 4, 6, 5, 7])
    v = values[:, order].reshape(-1, 4, 8).transpose(0, 2, 1)
    codes = quantise_2bit(v)
    _SM_SWAP.take(codes, out=codes)
    codes <<= np.array([0, 2, 4, 6], dtype=np.uint8)
    out = np.bitwise_or.reduce(codes, axis=-1).ravel().view(np.uint64)
    return m4_reorder64(out).view('<u8')


M4_FT_MAGBITS = 0xf0faf050f0faf05    # mark4/payload.py:337, :342

# keyed (nchan, bps-or-magbit-mask, fanout) as mark4/payload.py:333-342
MARK4_DECODERS = {(2, 2, 4): _m4_dec_2ch_f4, (4, 2, 4): _m4_dec_4ch_f4,
                  (8, 2, 2): _m4_dec_8ch_f2, (8, 2, 4): _m4_dec_8ch_f4,
                  (16, M4_FT_MAGBITS, 2): _m4_dec_16ch_f2_ft}
MARK4_ENCODERS = {(2, 2, 4): _m4_enc_2ch_f4, (4, 2, 4): _m4_enc_4ch_f4,
                  (8, 2, 2): _m4_enc_8ch_f2, (8, 2, 4): _m4_enc_8ch_f4,
                  (16, M4_FT_MAGBITS, 2): _m4_enc_16ch_f2_ft}
MARK4_WORD_DTYPE = {8: '<u1', 16: '<u2', 32: '<u4', 64: '<u8'}  # header.py:26-29


def mark4_decode(words, nchan, fanout, ft=False):
    """(nsample, nchan) float32 from track words."""
    key = (nchan, M4_FT_MAGBITS if ft else 2, fanout)
    return MARK4_DECODERS[key](words)


def mark4_encode(values, nchan, fanout, ft=False):
    key = (nchan, M4_FT_MAGBITS if ft else 2, fanout)
    return MARK4_ENCODERS[key](values)


# --------------------------------------------------------------------------
# int8 formats (GUPPI, DADA, GSB 8 bit) and GSB signed nibbles
# guppi/payload.py:13-18, dada/payload.py:13-18, gsb/payload.py:24-53
# --------------------------------------------------------------------------
def int8_decode(words):
    return words.view(np.int8).astype(np.float32)


def int8_encode(values):
    return np.clip(np.rint(values), -128, 127).astype(np.int8)


def gsb4_decode(words):
    """Low nibble first, two's complement (gsb/payload.py:24-36)."""
    w = np.asarray(words).view(np.int8)
    split = np.left_shift(w[:, np.newaxis], np.array([4, 0], np.int8)).ravel()
    split >>= 4
    return split.astype(np.float32)


def gsb4_encode(values):
    """gsb/payload.py:45-49."""
    b = np.clip(np.around(values), -8, 7).astype(np.int8).reshape(-1, 2)
    b &= 0xf
    b <<= np.array([0, 4], np.int8)
    return b[:, 0] | b[:, 1]


# --------------------------------------------------------------------------
# Payload-level views: base/payload.py:314-330 and format overrides
# --------------------------------------------------------------------------
def as_samples(flat, sample_shape, complex_data):
    """``.view(dtype).reshape(-1, *sample_shape)`` (base/payload.py:315,
    :329-330): adjacent values pair up as (re, im)."""
    flat = np.ascontiguousarray(flat, dtype=np.float32).ravel()
    if complex_data:
        flat = flat.view(np.complex64)
    return flat.reshape((-1,) + tuple(sample_shape))


def as_reals(data):
    """Complex viewed as trailing (re, im) pairs (base/payload.py:323-324)."""
    data = np.asarray(data)
    if data.dtype.kind == 'c':
        data = data.view((data.real.dtype, (2,)))
    return data


def vdif_payload_decode(words, bps, sample_shape=(1,), complex_data=False,
                        mark5b=False):
    """VDIFPayload.data (vdif/payload.py:137-154; EDV 0xab uses the Mark 5B
    tables)."""
    flat = mark5b_decode(words, bps) if mark5b else vdif_decode(words, bps)
    return as_samples(flat, sample_shape, complex_data)


def vdif_payload_encode(data, bps, mark5b=False):
    """VDIFPayload.fromdata(...).words (base/payload.py:145-188, :317-348)."""
    reals = as_reals(data)
    enc = mark5b_encode(reals, bps) if mark5b else vdif_encode(reals, bps)
    return np.ascontiguousarray(enc).ravel().view('<u4')


def mark5b_payload_decode(words, bps=2, nchan=1):
    return as_samples(mark5b_decode(words, bps), (nchan,), False)


def mark5b_payload_encode(data, bps=2):
    return np.ascontiguousarray(mark5b_encode(np.asarray(data), bps)
                                ).ravel().view('<u4')


def dada_payload_decode(words, sample_shape, complex_data=True):
    """(time, pol, chan, re/im) int8 (dada/payload.py:21-51)."""
    return as_samples(int8_decode(np.asarray(words)), sample_shape,
                      complex_data)


def mkbf_payload_decode(words, sample_shape, complex_data=True):
    """Heaps (nheap, npol, nchan, 256[, re/im]) -> time-major
    (dada/payload.py:54-89)."""
    ncomp = 2 if complex_data else 1
    raw = np.asarray(words).view(np.int8).reshape(
        (-1,) + tuple(sample_shape) + (256, ncomp))
    ordered = np.moveaxis(raw, -2, 1)
    return as_samples(int8_decode(np.ascontiguousarray(ordered).ravel()),
                      sample_shape, complex_data)


def mkbf_payload_encode(data, sample_shape):
    reals = as_reals(np.asarray(data))
    ncomp = reals.shape[-1] if np.asarray(data).dtype.kind == 'c' else 1
    v = reals.reshape((-1, 256) + tuple(sample_shape) + (ncomp,))
    v = np.moveaxis(v, 1, -2)
    return int8_encode(v).ravel()


def guppi_payload_decode(words, npol, nchan, complex_data=True,
                         channels_first=True):
    """GUPPIPayload.data (guppi/payload.py:90-110).

    channels_first: stored (nchan, nsample, npol[, re/im]); otherwise
    (nsample, nchan, npol[, re/im]).  Result (nsample, npol, nchan).
    """
    flat = int8_decode(np.asarray(words))
    if complex_data:
        flat = flat.view(np.complex64)
    if channels_first:
        return flat.reshape(nchan, -1).T.reshape(-1, npol, nchan)
    return flat.reshape(-1, nchan, npol).transpose(0, 2, 1)


def guppi_payload_encode(data, channels_first=True):
    """guppi/payload.py:112-136 for a whole payload."""
    data = np.asarray(data)
    if channels_first:
        reals = as_reals(np.ascontiguousarray(data.transpose(2, 0, 1)))
    else:
        reals = as_reals(np.ascontiguousarray(data.transpose(0, 2, 1)))
    return int8_encode(reals).ravel()


def gsb_payload_decode(words, bps, sample_shape=(1,), complex_data=False):
    """gsb/payload.py:72-75 tables + base/payload.py:314-330."""
    w = np.asarray(words).view(np.int8)
    flat = gsb4_decode(w) if bps == 4 else int8_decode(w)
    return as_samples(flat, sample_shape, complex_data)


def gsb_payload_encode(data, bps):
    reals = as_reals(np.asarray(data))
    enc = gsb4_encode(reals) if bps == 4 else int8_encode(reals)
    return np.ascontiguousarray(enc).ravel().view(np.int8)


def gsb_interleave_files(parts, nthread, sample_nbytes):
    """Combine per-(thread, part) payload bytes into one word array
    (gsb/payload.py:115-131).  ``parts[thread][part]`` are int8 arrays."""
    npart = len(parts[0])
    nper = parts[0][0].size // sample_nbytes
    words = np.empty((npart, nper, nthread, sample_nbytes), np.int8)
    for thread_parts, view in zip(parts, words.transpose(2, 0, 1, 3)):
        for part, dest in zip(thread_parts, view):
            dest[:] = np.asarray(part).view(np.int8).reshape(-1,
                                                             sample_nbytes)
    return words.ravel()
