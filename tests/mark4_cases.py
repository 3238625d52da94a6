"""Shared Mark 4 cases (CPU emulation + GPU parity), oracle as the checker."""
import zlib

import numpy as np

from oracle import codec, headers

MODES = {'2_4': (2, 4, False), '4_4': (4, 4, False), '8_2': (8, 2, False),
         '8_4': (8, 4, False), '16_2ft': (16, 2, True)}

# (id, mode, nframe, invalid frames, sample_start, nsample or None, fill)
FRAME_CASES = [
    ('c3_8ch_f4', '8_4', 2, (), 0, None, 0.0),
    ('8ch_f4_partial', '8_4', 3, (1,), 79000, 90000, -7.0),
    ('8ch_f4_header_edge', '8_4', 1, (), 637, 9, 5.0),
    ('4ch_f4', '4_4', 2, (), 0, None, 0.0),
    ('4ch_f4_odd', '4_4', 2, (0,), 12345, 80001, 1.5),
    ('2ch_f4', '2_4', 2, (), 0, None, 0.0),
    ('2ch_f4_odd_start', '2_4', 2, (), 639, 1001, 3.0),
    ('8ch_f2', '8_2', 2, (1,), 100, 50000, -1.0),
    ('16ch_f2_ft', '16_2ft', 2, (), 0, None, 0.0),
    ('16ch_f2_ft_part', '16_2ft', 2, (), 319, 777, 2.0),
]


def make_frames(mode, nframe, invalid, seed):
    nchan, fanout, ft = MODES[mode]
    ntrack = nchan * 2 * fanout
    wb = ntrack // 8
    rng = np.random.default_rng(zlib.crc32(seed.encode()))
    frame_nbytes = 20000 * wb
    gap = 64                       # frames need not be contiguous
    raw = rng.integers(0, 256, nframe * (frame_nbytes + gap) + 64,
                       dtype=np.uint8)
    order = rng.permutation(nframe)
    start = order * (frame_nbytes + gap)
    truth = (start + 160 * wb).astype(np.int64)
    unit_offset = truth.copy()
    for f in invalid:
        unit_offset[f] = -1
    return dict(mode=mode, nchan=nchan, fanout=fanout, ft=ft, ntrack=ntrack,
                wb=wb, raw=raw, truth=truth, unit_offset=unit_offset,
                nframe=nframe, spf=20000 * fanout,
                payload_nbytes=(20000 - 160) * wb)


def oracle_frames(c, fill, start, count):
    spf, nchan = c['spf'], c['nchan']
    full = np.full((c['nframe'] * spf, nchan), np.float32(fill))
    dt = codec.MARK4_WORD_DTYPE[c['ntrack']]
    for f in range(c['nframe']):
        if c['unit_offset'][f] < 0:
            continue
        o = c['truth'][f]
        words = c['raw'][o:o + c['payload_nbytes']].view(dt)
        body = codec.mark4_decode(words, nchan, c['fanout'], c['ft'])
        full[f * spf + 160 * c['fanout']:(f + 1) * spf] = body
    if count is None:
        count = full.shape[0] - start
    return full[start:start + count]
