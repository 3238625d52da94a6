"""CPU emulation of the kernel bodies + launch planning against the oracle.

The per-thread bodies in baseband_b200/csrc/*.cuh are host/device code; here
they are compiled with g++ (tests/emu) and every item is run in a loop.  This
checks index arithmetic, edge paths and quantisers without a GPU; the same
cases run on the real kernels in tests/test_gpu_parity.py.
"""
import ctypes

import numpy as np
import pytest

from baseband_b200 import levels
from oracle import codec
import emu_build
from bitfield_cases import (DECODE_CASES, ENCODE_CASES, make_decode_case,
                            oracle_decode, make_encode_case, oracle_encode)


@pytest.fixture(scope='module')
def emu():
    return emu_build.load()


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize('case', DECODE_CASES, ids=lambda c: c['id'])
def test_decode_bitfield(emu, case):
    c = make_decode_case(case)
    want = oracle_decode(c)
    out = np.full(want.size + 8, np.float32(np.nan))
    base = out[(-out.ctypes.data // 4) % 4:][:want.size]   # 16-byte aligned
    assert base.ctypes.data % 16 == 0
    lv = c['levels']
    rc = emu.bb_decode_bitfield(
        _ptr(c['raw']), _ptr(c['unit_offset']), c['nset'], c['nthread'],
        c['payload_nbytes'], c['bps'], c['nelem'], int(c['complex']),
        c['codec'],
        lv.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if lv is not None
        else None, c['fill'], c['sample_start'], c['nsample'], _ptr(base),
        None)
    assert rc == 0, emu.bb_last_error()
    assert np.array_equal(base.view('u4'), want.ravel().view('u4'))


@pytest.mark.parametrize('case', ENCODE_CASES, ids=lambda c: c['id'])
def test_encode_bitfield(emu, case):
    c = make_encode_case(case)
    want = oracle_encode(c)
    dst = np.full(c['dst_nbytes'], 0xEE, np.uint8)
    rc = emu.bb_encode_bitfield(
        _ptr(c['data']), c['dtype_code'], _ptr(dst), _ptr(c['unit_offset']),
        c['nset'], c['nthread'], c['payload_nbytes'], c['bps'], c['nelem'],
        c['quantiser'], None)
    assert rc == 0, emu.bb_last_error()
    assert np.array_equal(dst, want)


def test_golden_encode_vectors(emu, codec_vectors):
    """The reference's own outputs on threshold-hugging inputs."""
    g = codec_vectors
    for tag, code in (('f32', 0), ('f64', 1)):
        vals = np.ascontiguousarray(g['enc_in_' + tag])
        finite = np.ascontiguousarray(g['enc_in_finite_' + tag])
        for quant, name, bpss, src in (
                (0, 'vdif_enc%d_', (1, 2, 4, 8), vals),
                (1, 'm5b_enc%d_', (1, 2), vals),
                (2, {4: 'gsb4_enc_', 8: 'int8_enc_'}, (4, 8), finite)):
            for bps in bpss:
                key = (name[bps] if isinstance(name, dict)
                       else name % bps) + tag
                want = g[key]
                dst = np.zeros(want.size, np.uint8)
                off = np.zeros(1, np.int64)
                rc = emu.bb_encode_bitfield(
                    _ptr(src), code, _ptr(dst), _ptr(off), 1, 1, dst.size,
                    bps, 1, quant, None)
                assert rc == 0, emu.bb_last_error()
                bad = np.nonzero(dst != want)[0]
                assert bad.size == 0, (key, bad[:5])


def test_golden_decode_vectors(emu, codec_vectors):
    g = codec_vectors
    words = np.ascontiguousarray(g['words32'])
    off = np.zeros(1, np.int64)
    for key, bps, codec_id, lv in (
            ('vdif_dec1', 1, 0, levels.offset_binary(1)),
            ('vdif_dec2', 2, 0, levels.offset_binary(2)),
            ('vdif_dec4', 4, 0, levels.offset_binary(4)),
            ('vdif_dec8', 8, 0, levels.offset_binary(8)),
            ('m5b_dec1', 1, 0, levels.mark5b(1)),
            ('m5b_dec2', 2, 0, levels.mark5b(2))):
        want = g[key]
        out = np.zeros(want.size + 4, np.float32)
        base = out[(-out.ctypes.data // 4) % 4:][:want.size]
        lv = np.ascontiguousarray(lv, np.float32)
        rc = emu.bb_decode_bitfield(
            _ptr(words), _ptr(off), 1, 1, words.nbytes, bps, 1, 0, codec_id,
            lv.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 0.0, 0,
            want.size, _ptr(base), None)
        assert rc == 0, emu.bb_last_error()
        assert np.array_equal(base.view('u4'), want.view('u4')), key
    b = np.ascontiguousarray(g['bytes'])
    for key, bps in (('gsb4_dec', 4), ('int8_dec', 8)):
        want = g[key]
        out = np.zeros(want.size + 4, np.float32)
        base = out[(-out.ctypes.data // 4) % 4:][:want.size]
        rc = emu.bb_decode_bitfield(_ptr(b), _ptr(off), 1, 1, b.nbytes, bps,
                                    1, 0, 1, None, 0.0, 0, want.size,
                                    _ptr(base), None)
        assert rc == 0
        assert np.array_equal(base, want), key
