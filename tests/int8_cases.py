"""Shared int8-transpose cases (GUPPI channels-first, MKBF heaps)."""
import zlib

import numpy as np

from oracle import codec

# (id, nunit, nrow, ncol, item_nbytes, windows [(col_begin, col_end)] or None)
CASES = [
    ('guppi_512ch_2pol', 2, 512, 256 * 2, 2, None),
    ('guppi_overlap', 3, 64, 200 * 2, 2, [(0, 400), (32, 400), (32, 400)]),
    ('guppi_4ch_puppi_like', 4, 4, 1024 * 2, 2, [(0, 2048)] + [(128, 2048)] * 3),
    ('guppi_partial_read', 2, 96, 300, 2, [(7, 300), (0, 123)]),
    ('real_1byte_items', 2, 40, 333, 1, [(5, 333), (0, 300)]),
    ('mkbf_heaps', 6, 2 * 16, 256, 2, None),
    ('odd_rows', 1, 33, 130, 2, None),
    ('invalid_unit', 3, 32, 64, 2, None),
]


def make_case(case):
    cid, nunit, nrow, ncol, ib, windows = case
    rng = np.random.default_rng(zlib.crc32(cid.encode()))
    unit_nbytes = nrow * ncol * ib
    gap = 16
    raw = rng.integers(0, 256, nunit * (unit_nbytes + gap) + 16,
                       dtype=np.uint8)
    order = rng.permutation(nunit)
    truth = (order * (unit_nbytes + gap)).astype(np.int64)
    unit_offset = truth.copy()
    if cid == 'invalid_unit':
        unit_offset[1] = -1
    if windows is None:
        windows = [(0, ncol)] * nunit
    cb = np.array([w[0] for w in windows], np.int64)
    ce = np.array([w[1] for w in windows], np.int64)
    oc0 = np.concatenate([[0], np.cumsum(ce - cb)[:-1]]).astype(np.int64)
    return dict(id=cid, nunit=nunit, nrow=nrow, ncol=ncol, ib=ib, raw=raw,
                truth=truth, unit_offset=unit_offset, col_begin=cb,
                col_end=ce, out_col0=oc0, ncols_out=int((ce - cb).sum()),
                unit_nbytes=unit_nbytes)


def oracle_decode(c, fill=np.nan):
    """out[col][row][ib] float32; untouched (invalid) units keep ``fill``."""
    out = np.full((c['ncols_out'], c['nrow'], c['ib']), np.float32(fill))
    for u in range(c['nunit']):
        if c['unit_offset'][u] < 0:
            continue
        o = c['truth'][u]
        unit = c['raw'][o:o + c['unit_nbytes']].view(np.int8).reshape(
            c['nrow'], c['ncol'], c['ib'])
        dec = codec.int8_decode(unit.ravel()).reshape(unit.shape)
        cb, ce, oc0 = c['col_begin'][u], c['col_end'][u], c['out_col0'][u]
        out[oc0:oc0 + ce - cb] = dec[:, cb:ce].transpose(1, 0, 2)
    return out


def fuzz_cases(n, seed):
    """Random transposed-int8 geometries: odd/even rows (fast and generic
    kernels), ragged tiles, random column windows."""
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n):
        nunit = int(rng.integers(1, 4))
        nrow = int(rng.choice([2, 4, 30, 64, 66, 128, 130, 33, 7]))
        ncol = int(rng.integers(1, 400))
        ib = int(rng.choice([1, 2]))
        if rng.random() < 0.5:
            windows = None
        else:
            windows = []
            for _ in range(nunit):
                a = int(rng.integers(0, ncol))
                b = int(rng.integers(a + 1, ncol + 1))
                windows.append((a, b))
        cases.append(('fuzz%d_r%d_c%d_ib%d' % (i, nrow, ncol, ib), nunit, nrow,
                      ncol, ib, windows))
    return cases


# ---------------------------------------------------------------- time first
# (id, nunit, nsample, nchan, npol, item_nbytes, windows [(t_begin, t_end)])
TF_CASES = [
    ('tf_32ch_2pol_cplx', 2, 100, 32, 2, 2, None),
    ('tf_overlap_windows', 3, 64, 16, 2, 2, [(5, 64), (8, 64), (8, 40)]),
    ('tf_real_4pol', 2, 50, 8, 4, 1, [(0, 50), (3, 17)]),
    ('tf_odd_chan_scalar', 2, 33, 5, 2, 2, [(1, 33), (0, 33)]),
    ('tf_real_3ch_scalar', 1, 40, 3, 2, 1, None),
    ('tf_1pol', 2, 64, 12, 1, 2, None),
    ('tf_invalid_unit', 3, 16, 8, 2, 2, None),
    # units at odd byte offsets: the bytewise group load / store
    ('tf_shift4_2pol', 2, 20, 8, 2, 2, [(2, 20), (0, 11)]),
    ('tf_shift2_4pol_real', 2, 20, 8, 4, 1, None),
    ('tf_shift1_1pol', 2, 9, 4, 1, 2, None),
]


def make_tf_case(case):
    cid, nunit, nsample, nchan, npol, ib, windows = case
    rng = np.random.default_rng(zlib.crc32(cid.encode()))
    unit_nbytes = nsample * nchan * npol * ib
    gap = 16
    raw = rng.integers(0, 256, nunit * (unit_nbytes + gap) + 16,
                       dtype=np.uint8)
    order = rng.permutation(nunit)
    truth = (order * (unit_nbytes + gap)).astype(np.int64)
    if 'shift' in cid:
        truth += int(cid.split('shift')[1][0])
    unit_offset = truth.copy()
    if cid == 'tf_invalid_unit':
        unit_offset[1] = -1
    if windows is None:
        windows = [(0, nsample)] * nunit
    tb = np.array([w[0] for w in windows], np.int64)
    te = np.array([w[1] for w in windows], np.int64)
    t0 = np.concatenate([[0], np.cumsum(te - tb)[:-1]]).astype(np.int64)
    return dict(id=cid, nunit=nunit, nsample=nsample, nchan=nchan, npol=npol,
                ib=ib, raw=raw, truth=truth, unit_offset=unit_offset,
                t_begin=tb, t_end=te, out_t0=t0, nout=int((te - tb).sum()),
                unit_nbytes=unit_nbytes)


def oracle_tf_decode(c, fill=np.nan):
    """out[sample][pol][chan][ib] float32 (baseband/guppi/payload.py:97-102
    through the oracle's guppi_payload_decode, channels_first=False)."""
    out = np.full((c['nout'], c['npol'], c['nchan'], c['ib']),
                  np.float32(fill))
    for u in range(c['nunit']):
        if c['unit_offset'][u] < 0:
            continue
        o = c['truth'][u]
        words = c['raw'][o:o + c['unit_nbytes']].view(np.int8)
        dec = codec.guppi_payload_decode(words, c['npol'], c['nchan'],
                                         c['ib'] == 2, channels_first=False)
        if c['ib'] == 2:
            dec = np.stack([dec.real, dec.imag], -1)
        else:
            dec = dec[..., None]
        tb, te, t0 = c['t_begin'][u], c['t_end'][u], c['out_t0'][u]
        out[t0:t0 + te - tb] = dec[tb:te]
    return out


def tf_fuzz_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n):
        nunit = int(rng.integers(1, 4))
        nsample = int(rng.integers(1, 80))
        nchan = int(rng.choice([1, 2, 3, 4, 6, 8, 16, 64]))
        npol = int(rng.choice([1, 2, 4]))
        ib = int(rng.choice([1, 2]))
        if rng.random() < 0.5:
            windows = None
        else:
            windows = []
            for _ in range(nunit):
                a = int(rng.integers(0, nsample))
                b = int(rng.integers(a + 1, nsample + 1))
                windows.append((a, b))
        cases.append(('tffuzz%d' % i, nunit, nsample, nchan, npol, ib,
                      windows))
    return cases
