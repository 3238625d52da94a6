"""Shared case tables for the bit-field codec: used by the CPU emulation
tests and by the GPU parity tests, with the oracle as the checker."""
import zlib

import numpy as np

from baseband_b200 import levels
from oracle import codec


def _case(id, bps, nelem, nthread, nset, payload_nbytes, cplx=False,
          kind='vdif', start=0, count=None, invalid=(), shuffle=True,
          fill=0.0):
    return dict(id=id, bps=bps, nelem=nelem, nthread=nthread, nset=nset,
                payload_nbytes=payload_nbytes, complex=cplx, kind=kind,
                start=start, count=count, invalid=invalid, shuffle=shuffle,
                fill=fill)


DECODE_CASES = [
    # ROWGROUP4: VDIF multi-thread single channel (configs C1, C2)
    _case('c1_2bit_8thr', 2, 1, 8, 2, 5000),
    _case('c2_2bit_16thr', 2, 1, 16, 3, 8000),
    _case('2bit_16thr_partial', 2, 1, 16, 3, 800, start=3217, count=4001),
    _case('2bit_8thr_invalid', 2, 1, 8, 3, 400, invalid=(1, 9, 10, 23),
          fill=-999.0),
    # rows of four float4 (TILE modes when selected)
    _case('2bit_16thr_invalid_partial', 2, 1, 16, 4, 404, start=1619,
          count=4200, invalid=(0, 17, 18, 35, 63), fill=-999.0),
    _case('2bit_16thr_odd_words', 2, 1, 16, 3, 36, invalid=(40,), fill=3.0),
    _case('2bit_cplx_8thr', 2, 2, 8, 3, 260, cplx=True, start=77, count=700,
          invalid=(3, 12), fill=-5.0),
    _case('2bit_cplx_8thr_full', 2, 2, 8, 2, 800, cplx=True),
    _case('2bit_4thr_one_sample', 2, 1, 4, 2, 64, start=255, count=2),
    _case('1bit_8thr', 1, 1, 8, 2, 128, start=5, count=1999),
    _case('4bit_4thr', 4, 1, 4, 2, 256, invalid=(3,), fill=7.5),
    _case('8bit_12thr', 8, 1, 12, 2, 200, start=1, count=300),
    # WORDROW: rows that are exactly one float4 (4 real / 2 complex threads)
    _case('1bit_4thr_rowrun', 1, 1, 4, 3, 68, start=33, count=1200,
          invalid=(6,), fill=-2.0),
    _case('8bit_4thr_rowrun', 8, 1, 4, 2, 404, start=3, count=700),
    _case('4bit_cplx_2thr_rowrun', 4, 2, 2, 3, 100, cplx=True, start=1,
          count=250, invalid=(3,), fill=5.0),
    _case('1bit_cplx_2thr_wordrow', 1, 2, 2, 3, 68, cplx=True, start=17,
          count=700, invalid=(4,), fill=-7.0),
    _case('2bit_cplx_2thr_wordrow', 2, 2, 2, 3, 200, cplx=True, start=399,
          count=403),
    # ROWGROUP2: complex single channel / two channels
    _case('2bit_cplx_2thr', 2, 2, 2, 3, 512, cplx=True, invalid=(2,),
          fill=-3.0),
    _case('8bit_cplx_2thr_mwa', 8, 2, 2, 2, 1024, cplx=True, start=100,
          count=700),
    _case('4bit_2chan_6thr', 4, 2, 6, 2, 96, start=7, count=150),
    _case('1bit_2chan_4thr', 1, 2, 4, 2, 64),
    # RUN: one thread, any nelem; or nelem power of two >= 4
    _case('m5b_2bit_16ch', 2, 16, 1, 4, 10000, kind='mark5b',
          invalid=(2,), fill=-999.0),
    _case('m5b_2bit_8ch_partial', 2, 8, 1, 3, 10000, kind='mark5b',
          start=4999, count=7002),
    _case('m5b_1bit_4ch', 1, 4, 1, 2, 10000, kind='mark5b'),
    _case('2bit_1thr_1ch', 2, 1, 1, 3, 5000, start=4, count=59992),
    _case('2bit_1thr_3ch', 2, 3, 1, 2, 12 * 25),
    _case('1bit_16ch_bps1', 1, 16, 1, 2, 8000),
    _case('4bit_cplx_1024ch_aro', 4, 2048, 1, 2, 2048 * 5, cplx=True),
    _case('2bit_4thr_8ch', 2, 8, 4, 3, 640, invalid=(5,), fill=2.0,
          start=16, count=500),
    _case('8bit_2thr_cplx_4ch', 8, 8, 2, 2, 800, cplx=True, invalid=(0,),
          fill=-1.0),
    _case('4bit_3thr_4ch', 4, 4, 3, 2, 240),
    # RUNS: S = 2, 4, 8 samples per word, reads aligned to S rows (incl. one
    # that starts and ends inside frames), invalid units; and the same shapes
    # with an odd start, which must fall back to RUN
    _case('runs_2bit_4thr_8ch_full', 2, 8, 4, 3, 640, invalid=(5, 6)),
    _case('runs_2bit_8thr_4ch_partial', 2, 4, 8, 3, 404, start=4 * 33,
          count=4 * 201, invalid=(11,), fill=-3.0),
    _case('runs_1bit_2thr_4ch', 1, 4, 2, 2, 68, start=8, count=8 * 30),
    _case('runs_4bit_5thr_4ch', 4, 4, 5, 2, 240, start=2, count=2 * 77,
          invalid=(0, 9)),
    _case('runs_2bit_2thr_cplx_2ch', 2, 4, 2, 2, 320, cplx=True, start=12,
          count=4 * 50, invalid=(1,), fill=5.0),
    _case('runs_2bit_4thr_8ch_s2_only', 2, 8, 4, 3, 640, start=2 * 15,
          count=2 * 401, invalid=(2,)),            # aligned to 2, not to 4
    _case('runs_8bit_3thr_4ch', 8, 4, 3, 2, 400, start=8, count=4 * 40,
          invalid=(4,), fill=1.5),                 # S = 1: four words an item
    _case('runs_2bit_2thr_16ch', 2, 16, 2, 2, 320, invalid=(3,)),
    _case('runs_4bit_2thr_8ch_cplx', 4, 8, 2, 2, 256, cplx=True, start=4,
          count=4 * 13),
    _case('run_2bit_8thr_4ch_odd_start', 2, 4, 8, 3, 404, start=4 * 33 + 1,
          count=4 * 201),
    # SCALAR: odd geometries / unaligned row ranges
    _case('2bit_3thr_1ch', 2, 1, 3, 3, 100, invalid=(4,), fill=9.0),
    _case('2bit_1thr_1ch_unaligned', 2, 1, 1, 2, 500, start=3, count=1001),
    _case('2bit_5thr_cplx', 2, 2, 5, 2, 200, cplx=True, invalid=(7,),
          fill=-999.0),
    _case('4bit_2thr_6ch', 4, 6, 2, 2, 240, start=1, count=77),
    _case('8bit_1thr_3ch_unaligned', 8, 3, 1, 2, 300, start=1, count=150),
    # non-standard level tables (always the shared-memory table path)
    _case('custom8_4thr', 8, 1, 4, 2, 404, kind='custom', start=3, count=700),
    _case('custom8_1thr_cplx', 8, 2, 1, 2, 800, cplx=True, kind='custom'),
    _case('custom8_12thr', 8, 1, 12, 2, 200, kind='custom', invalid=(5,),
          fill=1.5),
    _case('custom2_16thr', 2, 1, 16, 2, 800, kind='custom'),
    _case('custom4_2thr_cplx', 4, 2, 2, 2, 100, cplx=True, kind='custom'),
    # signed-integer codecs
    _case('gsb_4bit_rawdump', 4, 1, 1, 2, 4096, kind='sint'),
    _case('dada_8bit_cplx_2pol', 8, 4, 1, 2, 6400, cplx=True, kind='sint'),
    _case('gsb_8bit_phased_2thr_cplx', 8, 1024, 2, 2, 4096, cplx=True,
          kind='sint', invalid=(1,), fill=0.5),
    _case('sint8_3thr', 8, 1, 3, 2, 120, kind='sint', start=5, count=100),
]


def custom_levels(bps):
    """A level table that is NOT the standard one (exercises the table path
    where the standard 8-bit levels are computed arithmetically)."""
    lv = np.ascontiguousarray(levels.offset_binary(bps), np.float32).copy()
    lv[::3] *= np.float32(1.25)
    lv[1] = np.float32(-77.0)
    return lv


def _levels(kind, bps):
    if kind == 'custom':
        return custom_levels(bps), 0
    if kind == 'vdif':
        return np.ascontiguousarray(levels.offset_binary(bps), np.float32), 0
    if kind == 'mark5b':
        return np.ascontiguousarray(levels.mark5b(bps), np.float32), 0
    return None, 1


def make_decode_case(case):
    rng = np.random.default_rng(zlib.crc32(case['id'].encode()))
    nunit = case['nset'] * case['nthread']
    nbytes = case['payload_nbytes']
    stride = nbytes + 32          # leave room for a (random) header
    order = rng.permutation(nunit) if case['shuffle'] else np.arange(nunit)
    raw = rng.integers(0, 256, nunit * stride + 64, dtype=np.uint8)
    unit_offset = (order * stride + 32).astype(np.int64)
    truth = unit_offset.copy()
    for u in case['invalid']:
        unit_offset[u] = -1
    spf = nbytes * 8 // (case['bps'] * case['nelem'])
    count = case['count']
    if count is None:
        count = case['nset'] * spf - case['start']
    lv, codec_id = _levels(case['kind'], case['bps'])
    return dict(case, raw=raw, unit_offset=unit_offset, truth=truth, spf=spf,
                sample_start=case['start'], nsample=count, levels=lv,
                codec=codec_id)


def _decode_unit(words, kind, bps):
    if kind == 'custom':
        codes = words.view(np.uint8)
        if bps < 8:
            shifts = np.arange(0, 8, bps, dtype=np.uint8)
            codes = (codes[:, None] >> shifts) & ((1 << bps) - 1)
        return custom_levels(bps)[codes.ravel()]
    if kind == 'vdif':
        return codec.vdif_decode(words, bps).ravel()
    if kind == 'mark5b':
        return codec.mark5b_decode(words, bps).ravel()
    b = words.view(np.int8)
    return (codec.gsb4_decode(b) if bps == 4 else codec.int8_decode(b)).ravel()


def oracle_decode(c):
    nset, nthread, nelem, spf = c['nset'], c['nthread'], c['nelem'], c['spf']
    full = np.empty((nset * spf, nthread, nelem), np.float32)
    for s in range(nset):
        for t in range(nthread):
            u = s * nthread + t
            blk = full[s * spf:(s + 1) * spf, t]
            if c['unit_offset'][u] < 0:
                blk[:] = c['fill']
                if c['complex']:
                    blk[:, 1::2] = 0.0
            else:
                o = c['truth'][u]
                words = c['raw'][o:o + c['payload_nbytes']].view('<u4')
                blk[:] = _decode_unit(words, c['kind'], c['bps']).reshape(
                    spf, nelem)
    return full[c['sample_start']:c['sample_start'] + c['nsample']]


def _ecase(id, bps, nelem, nthread, nset, payload_nbytes, quant='vdif',
           dtype='f4', invalid=(), offset0=32):
    return dict(id=id, bps=bps, nelem=nelem, nthread=nthread, nset=nset,
                payload_nbytes=payload_nbytes, quant=quant, dtype=dtype,
                invalid=invalid, offset0=offset0)


ENCODE_CASES = [
    _ecase('c2_2bit_16thr', 2, 1, 16, 3, 8000),
    _ecase('c2_2bit_16thr_f64', 2, 1, 16, 2, 800, dtype='f8'),
    _ecase('2bit_8thr_skip', 2, 1, 8, 3, 400, invalid=(1, 9)),
    _ecase('1bit_8thr', 1, 1, 8, 2, 128),
    _ecase('4bit_4thr', 4, 1, 4, 2, 256),
    _ecase('8bit_12thr', 8, 1, 12, 2, 200),
    _ecase('2bit_cplx_2thr', 2, 2, 2, 3, 512),
    _ecase('8bit_cplx_2thr_f64', 8, 2, 2, 2, 1024, dtype='f8'),
    _ecase('1bit_2chan_4thr', 1, 2, 4, 2, 64),
    _ecase('m5b_2bit_16ch', 2, 16, 1, 3, 10000, quant='mark5b'),
    _ecase('m5b_1bit_4ch_f64', 1, 4, 1, 2, 10000, quant='mark5b',
           dtype='f8'),
    _ecase('2bit_1thr_1ch', 2, 1, 1, 3, 5000),
    _ecase('2bit_1thr_3ch', 2, 3, 1, 2, 300),
    _ecase('4bit_cplx_1024ch', 4, 2048, 1, 2, 2048 * 5),
    _ecase('2bit_4thr_8ch', 2, 8, 4, 3, 640, invalid=(5,)),
    _ecase('4bit_3thr_4ch_f64', 4, 4, 3, 2, 240, dtype='f8'),
    _ecase('2bit_3thr_1ch', 2, 1, 3, 3, 100),
    _ecase('2bit_5thr_cplx', 2, 2, 5, 2, 200),
    _ecase('4bit_2thr_6ch', 4, 6, 2, 2, 240),
    # a row is one float4 (4 real / 2 complex threads): warp-cooperative
    # ROWWORD mode, chunks of 32 word positions crossing frame-set boundaries
    _ecase('1bit_4thr_rowword', 1, 1, 4, 3, 200),
    _ecase('2bit_4thr_rowword', 2, 1, 4, 3, 200, invalid=(2, 7)),
    _ecase('2bit_4thr_rowword_f64', 2, 1, 4, 2, 132, dtype='f8'),
    _ecase('8bit_4thr_rowword', 8, 1, 4, 3, 200),
    _ecase('1bit_cplx_2thr_rowword', 1, 2, 2, 3, 200),
    _ecase('4bit_cplx_2thr_rowword', 4, 2, 2, 3, 200, invalid=(0,)),
    _ecase('8bit_cplx_2thr_rowword', 8, 2, 2, 3, 200),
    _ecase('gsb_4bit', 4, 1, 1, 2, 4096, quant='sint'),
    _ecase('dada_8bit_cplx', 8, 4, 1, 2, 6400, quant='sint'),
    _ecase('gsb_8bit_2thr_f64', 8, 1024, 2, 2, 4096, quant='sint',
           dtype='f8'),
    # 8 bit in items of four words (RUNQ): one thread, or >= 16 elements per
    # thread row; payloads not at 16-byte offsets (four 4-byte stores); and
    # the shapes that must stay with one word per item
    _ecase('gsb_8bit_2thr', 8, 1024, 2, 3, 8192, quant='sint', invalid=(3,)),
    _ecase('vdif_8bit_1thr', 8, 1, 1, 3, 8000),
    _ecase('vdif_8bit_1thr_16ch_off8', 8, 16, 1, 3, 8000, offset0=40),
    _ecase('vdif_8bit_3thr_32ch_off4', 8, 32, 3, 2, 4096, offset0=36,
           invalid=(4,)),
    _ecase('dada_8bit_cplx_f64_off8', 8, 4, 1, 2, 6400, quant='sint',
           dtype='f8', offset0=8),
    _ecase('8bit_1thr_odd_words', 8, 2, 1, 2, 8008),
    _ecase('8bit_2thr_8ch', 8, 8, 2, 2, 4096),
]

_QUANT = {'vdif': 0, 'mark5b': 1, 'sint': 2}


def make_encode_case(case):
    rng = np.random.default_rng(zlib.crc32(case['id'].encode()))
    nset, nthread, nelem = case['nset'], case['nthread'], case['nelem']
    nbytes = case['payload_nbytes']
    spf = nbytes * 8 // (case['bps'] * nelem)
    scale = 40.0 if case['bps'] == 8 and case['quant'] == 'sint' else 2.5
    dtype = np.dtype(case['dtype'])
    data = (rng.standard_normal((nset * spf, nthread, nelem)) * scale
            ).astype(dtype)
    # sprinkle exact thresholds and half-integers
    flat = data.reshape(-1)
    marks = np.array([0.0, -0.0, 2.174564, -2.174564, 0.5, 1.5, 2.5, -0.5,
                      3.2618460000000002, 1e30, -1e30], dtype)
    flat[rng.integers(0, flat.size, marks.size * 4)] = np.tile(marks, 4)
    nunit = nset * nthread
    stride = nbytes + 32
    order = rng.permutation(nunit)
    unit_offset = (order * stride + case.get('offset0', 32)).astype(np.int64)
    truth = unit_offset.copy()
    for u in case['invalid']:
        unit_offset[u] = -1
    # 16-byte aligned input
    buf = np.zeros(data.size + 4, dtype)
    shift = (-buf.ctypes.data // dtype.itemsize) % (16 // dtype.itemsize)
    aligned = buf[shift:shift + data.size]
    aligned[:] = flat
    return dict(case, data=aligned, shaped=aligned.reshape(data.shape),
                unit_offset=unit_offset, truth=truth, spf=spf,
                dst_nbytes=nunit * stride + 64,
                dtype_code=0 if dtype.itemsize == 4 else 1,
                quantiser=_QUANT[case['quant']])


def oracle_encode(c):
    dst = np.full(c['dst_nbytes'], 0xEE, np.uint8)
    spf, nthread = c['spf'], c['nthread']
    for s in range(c['nset']):
        for t in range(nthread):
            u = s * nthread + t
            if c['unit_offset'][u] < 0:
                continue
            vals = np.ascontiguousarray(
                c['shaped'][s * spf:(s + 1) * spf, t]).ravel()
            with np.errstate(all='ignore'):
                if c['quant'] == 'vdif':
                    enc = codec.vdif_encode(vals.copy(), c['bps'])
                elif c['quant'] == 'mark5b':
                    enc = codec.mark5b_encode(vals.copy(), c['bps'])
                elif c['bps'] == 4:
                    enc = codec.gsb4_encode(vals)
                else:
                    enc = codec.int8_encode(vals)
            o = c['truth'][u]
            dst[o:o + c['payload_nbytes']] = np.ascontiguousarray(
                enc).ravel().view(np.uint8)
    return dst


def fuzz_cases(n, seed):
    """Seeded random decode / encode geometries covering every mode of the
    planner (thread counts 1-9, 1-16 elements, odd payload sizes, partial
    sample ranges, invalid units, all codecs)."""
    rng = np.random.default_rng(seed)
    dec, enc = [], []
    for i in range(n):
        kind = rng.choice(['vdif', 'vdif', 'mark5b', 'sint'])
        bps = int(rng.choice({'vdif': [1, 2, 4, 8], 'mark5b': [1, 2],
                              'sint': [4, 8]}[kind]))
        nthread = int(rng.choice([1, 1, 2, 2, 3, 4, 4, 5, 8, 9]))
        nelem = int(rng.choice([1, 1, 2, 2, 3, 4, 8, 16]))
        cplx = bool(nelem % 2 == 0 and rng.random() < 0.4)
        nset = int(rng.integers(1, 4))
        # payload: whole 32-bit words holding whole samples
        bits = bps * nelem
        unit = bits * 32 // np.gcd(bits, 32)          # lcm(bits, 32) bits
        payload = int(unit // 8 * rng.integers(1, 40))
        spf = payload * 8 // bits
        total = nset * spf
        if rng.random() < 0.5:
            start = int(rng.integers(0, total))
            count = int(rng.integers(1, total - start + 1))
        else:
            start, count = 0, None
        ninv = int(rng.integers(0, 3))
        invalid = tuple(int(v) for v in rng.choice(
            nset * nthread, size=min(ninv, nset * nthread), replace=False))
        fill = float(rng.choice([0.0, -999.0, 2.5]))
        cid = 'fuzz%d_%s_b%d_t%d_e%d' % (i, kind, bps, nthread, nelem)
        dec.append(_case(cid, bps, nelem, nthread, nset, payload, cplx=cplx,
                         kind=kind, start=start, count=count,
                         invalid=invalid, fill=fill))
        quant = {'vdif': 'vdif', 'mark5b': 'mark5b', 'sint': 'sint'}[kind]
        enc.append(_ecase(cid, bps, nelem, nthread, nset, payload,
                          quant=quant, invalid=invalid,
                          dtype='f8' if rng.random() < 0.3 else 'f4'))
    return dec, enc
