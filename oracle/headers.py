"""Oracle header bit-field extraction.  TEST INFRASTRUCTURE — see
oracle/__init__.py.

Restates the integer arithmetic of
* baseband/base/header.py:35-87      (``(word >> bit) & mask`` extractors)
* baseband/vdif/header.py:529-542, :557-559, :595-598, :701-702, :715-725,
  :762-770, :792-797, :293-364        (VDIF field tables + derived sizes)
* baseband/mark5b/header.py:60-68, :192-233; baseband/base/utils.py:18-34
  (Mark 5B fields, BCD); baseband/mark5b/base.py:206-213 (frame index)
* baseband/mark4/header.py:47-88, :116-142, :536-592 (track bit transpose,
  per-track fields, sizes); baseband/mark4/frame.py:78-87 (validity)
Time (astropy) arithmetic is outside the hot path and not restated.
"""
import numpy as np


def field(words, word, bit, nbits):
    """``(words[word] >> bit) & mask`` (base/header.py:35-87).  ``words`` may
    be a sequence of ints or an array whose first axis indexes words."""
    if nbits == 64:
        return int(words[word]) + (int(words[word + 1]) << 32)
    w = words[word]
    if isinstance(w, np.ndarray):
        w = w.astype(np.uint64)
        return ((w >> np.uint64(bit)) & np.uint64((1 << nbits) - 1)
                ).astype(np.int64)
    return (int(w) >> bit) & ((1 << nbits) - 1)


# (word, bit, nbits) -- vdif/header.py:529-542 (legacy = words 0..3)
VDIF_BASE_FIELDS = {
    'invalid_data': (0, 31, 1), 'legacy_mode': (0, 30, 1),
    'seconds': (0, 0, 30), 'ref_epoch': (1, 24, 6), 'frame_nr': (1, 0, 24),
    'vdif_version': (2, 29, 3), 'lg2_nchan': (2, 24, 5),
    'frame_length': (2, 0, 24), 'complex_data': (3, 31, 1),
    'bits_per_sample': (3, 26, 5), 'thread_id': (3, 16, 10),
    'station_id': (3, 0, 16)}
VDIF_EDV_FIELDS = {
    None: {},                                    # legacy
    0: {'edv': (4, 24, 8)},                      # header.py:557-559
    1: {'edv': (4, 24, 8), 'sampling_unit': (4, 23, 1),
        'sampling_rate': (4, 0, 23), 'sync_pattern': (5, 0, 32),
        'das_id': (6, 0, 64)},                   # :595-598, :701-702
    3: {'edv': (4, 24, 8), 'sampling_unit': (4, 23, 1),
        'sampling_rate': (4, 0, 23), 'sync_pattern': (5, 0, 32),
        'loif_tuning': (6, 0, 32), 'dbe_unit': (7, 24, 4),
        'if_nr': (7, 20, 4), 'subband': (7, 17, 3), 'sideband': (7, 16, 1),
        'major_rev': (7, 12, 4), 'minor_rev': (7, 8, 4),
        'personality': (7, 0, 8)},               # :715-725
    2: {'edv': (4, 24, 8), 'pol': (4, 0, 1), 'BL_quadrant': (4, 1, 2),
        'BL_correlator': (4, 3, 1), 'sync_pattern': (4, 4, 20),
        'PIC_status': (5, 0, 32), 'PSN': (6, 0, 64)},   # :762-770
    0xab: {'edv': (4, 24, 8), 'sync_pattern': (4, 0, 32),
           'user': (5, 16, 16), 'internal_tvg': (5, 15, 1),
           'mark5b_frame_nr': (5, 0, 15), 'bcd_jday': (6, 20, 12),
           'bcd_seconds': (6, 0, 20), 'bcd_fraction': (7, 16, 16),
           'crc': (7, 0, 16)},                   # :792-797
}
VDIF_SYNC = 0xACABFEED          # vdif/header.py:598


def vdif_parse(words):
    """All fields of one VDIF header (8 words, or 4 if legacy)."""
    out = {k: field(words, *v) for k, v in VDIF_BASE_FIELDS.items()}
    if out['legacy_mode']:
        edv = None
    else:
        edv = field(words, 4, 24, 8)
    for k, v in VDIF_EDV_FIELDS.get(edv, {'edv': (4, 24, 8)}).items():
        out[k] = field(words, *v)
    # derived: vdif/header.py:293-364
    out['header_nbytes'] = 16 if out['legacy_mode'] else 32
    out['frame_nbytes'] = out['frame_length'] * 8
    out['payload_nbytes'] = out['frame_nbytes'] - out['header_nbytes']
    out['bps'] = out['bits_per_sample'] + 1
    out['nchan'] = 2 ** out['lg2_nchan']
    values_per_word = (32 // out['bps']) // (2 if out['complex_data'] else 1)
    out['samples_per_frame'] = (out['payload_nbytes'] // 4 * values_per_word
                                // out['nchan'])
    return out


def vdif_parse_batch(raw, frame_nbytes, nframe):
    """Base fields for ``nframe`` equally spaced headers in a byte buffer.
    Returns dict of int64 arrays."""
    buf = np.frombuffer(raw, np.uint8, count=nframe * frame_nbytes)
    hw = buf.reshape(nframe, frame_nbytes)[:, :16].copy().view('<u4')
    return {k: field(hw.T, *v) for k, v in VDIF_BASE_FIELDS.items()}


def vdif_frame_index(seconds, frame_nr, seconds0, frame_nr0, frame_rate):
    """vdif/base.py:386-390."""
    return int(round((seconds - seconds0) * frame_rate + frame_nr
                     - frame_nr0))


# ---------------------------------------------------------------- Mark 5B
MARK5B_FIELDS = {                   # mark5b/header.py:60-68
    'sync_pattern': (0, 0, 32), 'user': (1, 16, 16),
    'internal_tvg': (1, 15, 1), 'frame_nr': (1, 0, 15),
    'bcd_jday': (2, 20, 12), 'bcd_seconds': (2, 0, 20),
    'bcd_fraction': (3, 16, 16), 'crc': (3, 0, 16)}
MARK5B_SYNC = 0xABADDEED


def bcd_decode(value):
    """BCD nibbles -> decimal (base/utils.py:18-34).  Scalar or array."""
    if isinstance(value, np.ndarray):
        value = value.astype(np.int64)
        digits = np.arange(16)
        nib = (value[..., np.newaxis] >> (4 * digits)) & 0xf
        if nib.max() > 9:
            raise ValueError("invalid BCD encoded value")
        return (nib * 10 ** digits).sum(-1)
    return int('{:x}'.format(int(value)))


def mark5b_parse(words):
    out = {k: field(words, *v) for k, v in MARK5B_FIELDS.items()}
    out['jday'] = bcd_decode(out['bcd_jday'])           # header.py:192-195
    out['seconds'] = bcd_decode(out['bcd_seconds'])     # :201-204
    ns = bcd_decode(out['bcd_fraction']) * 100000       # :210-225
    out['fraction_ns'] = 156250 * ((ns + 156249) // 156250)
    out['fraction'] = out['fraction_ns'] / 1e9
    return out


def mark5b_parse_batch(raw, nframe, frame_nbytes=10016):
    buf = np.frombuffer(raw, np.uint8, count=nframe * frame_nbytes)
    hw = buf.reshape(nframe, frame_nbytes)[:, :16].copy().view('<u4')
    out = {k: field(hw.T, *v) for k, v in MARK5B_FIELDS.items()}
    out['jday'] = bcd_decode(out['bcd_jday'])
    out['seconds'] = bcd_decode(out['bcd_seconds'])
    ns = bcd_decode(out['bcd_fraction']) * 100000
    out['fraction_ns'] = 156250 * ((ns + 156249) // 156250)
    return out


def mark5b_frame_index(hdr, hdr0, frame_rate, kday=0, kday0=0):
    """mark5b/base.py:206-213."""
    return int(round(frame_rate
                     * (hdr['seconds'] - hdr0['seconds']
                        + 86400 * (kday - kday0 + hdr['jday'] - hdr0['jday']))
                     + hdr['frame_nr'] - hdr0['frame_nr']))


# ---------------------------------------------------------------- Mark 4
MARK4_TRACK_FIELDS = {              # mark4/header.py:116-142
    'bcd_headstack1': (0, 0, 16), 'bcd_headstack2': (0, 16, 16),
    'headstack_id': (1, 30, 2), 'bcd_track_id': (1, 24, 6),
    'fan_out': (1, 22, 2), 'magnitude_bit': (1, 21, 1),
    'lsb_output': (1, 20, 1), 'converter_id': (1, 16, 4),
    'time_sync_error': (1, 15, 1), 'internal_clock_error': (1, 14, 1),
    'processor_time_out_error': (1, 13, 1), 'communication_error': (1, 12, 1),
    'track_roll_enabled': (1, 9, 1), 'sequence_suspended': (1, 8, 1),
    'system_id': (1, 0, 8), 'sync_pattern': (2, 0, 32),
    'bcd_unit_year': (3, 28, 4), 'bcd_day': (3, 16, 12),
    'bcd_hour': (3, 8, 8), 'bcd_minute': (3, 0, 8),
    'bcd_second': (4, 24, 8), 'bcd_fraction': (4, 12, 12), 'crc': (4, 0, 12)}
MARK4_HEADER_STEPS = 160            # 160 time steps of ntrack bits
MARK4_FRAME_STEPS = 20000           # mark4/header.py:32


def mark4_stream2words(stream):
    """(160,) track words -> (5, ntrack) uint32, MSB first within each word
    (mark4/header.py:47-63)."""
    stream = np.asarray(stream)
    ntrack = stream.dtype.itemsize * 8
    track = np.arange(ntrack, dtype=stream.dtype)
    sel = ((stream.reshape(-1, 32, 1) >> track) & 1).astype(np.uint32)
    sel <<= np.arange(31, -1, -1, dtype=np.uint32).reshape(-1, 1)
    return np.bitwise_or.reduce(sel, axis=1)


def mark4_parse(stream):
    """Per-track fields (arrays of length ntrack) from the 160 header
    words, plus the derived frame geometry and validity."""
    words = mark4_stream2words(stream)
    ntrack = words.shape[1]
    out = {k: field(words, *v) for k, v in MARK4_TRACK_FIELDS.items()}
    out['ntrack'] = ntrack
    out['fanout'] = int(out['fan_out'].max()) + 1        # header.py:558-563
    out['bps'] = 2 if out['magnitude_bit'].any() else 1  # payload.py:348
    out['nchan'] = ntrack // (out['bps'] * out['fanout'])
    out['header_nbytes'] = ntrack * MARK4_HEADER_STEPS // 8
    out['frame_nbytes'] = ntrack * MARK4_FRAME_STEPS // 8
    out['payload_nbytes'] = out['frame_nbytes'] - out['header_nbytes']
    out['samples_per_frame'] = (out['frame_nbytes'] * 8
                                // (ntrack // out['fanout']))   # :584-592
    out['valid'] = not bool(np.any(                      # frame.py:78-87
        out['time_sync_error'] | out['internal_clock_error']
        | out['processor_time_out_error'] | out['communication_error']))
    return out
