"""GPU parity for the Mark 4, int8-transpose and header-scan kernels (through
the C ABI) against the oracle, the reference golden vectors and the
reference's sample files."""
import numpy as np
import pytest
import torch

from baseband_b200 import kernels, levels
from oracle import codec, headers, stream
from conftest import sample_bytes
from mark4_cases import MODES, FRAME_CASES, make_frames, oracle_frames
import int8_cases

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize('tag', sorted(MODES))
def test_mark4_words_golden(codec_vectors, tag):
    g = codec_vectors
    nchan, fanout, ft = MODES[tag]
    words = g['m4_words_' + tag]
    want = g['m4_dec_' + tag]
    out = kernels.mark4_decode_words(_t(words.view(np.uint8)), words.size,
                                     nchan, fanout, ft,
                                     levels.sign_magnitude())
    assert np.array_equal(out.cpu().numpy().view('u4'), want.view('u4'))
    for ftag in ('f32', 'f64'):
        vals = g['m4_enc_in_%s_%s' % (tag, ftag)]
        wantw = g['m4_enc_%s_%s' % (tag, ftag)]
        dst = torch.zeros(wantw.size, dtype=torch.uint8, device=DEV)
        kernels.mark4_encode_words(_t(vals), dst, vals.shape[0] // fanout,
                                   nchan, fanout, ft)
        assert np.array_equal(dst.cpu().numpy(), wantw), (tag, ftag)
    with pytest.raises(KeyError):       # no such decoder (fanout 1)
        kernels.mark4_decode_words(_t(words.view(np.uint8)), 4, 16, 1, False,
                                   levels.sign_magnitude())


@pytest.mark.parametrize('case', FRAME_CASES, ids=lambda c: c[0])
def test_mark4_frames(case):
    cid, mode, nframe, invalid, start, count, fill = case
    c = make_frames(mode, nframe, invalid, cid)
    want = oracle_frames(c, fill, start, count)
    raw, uo = _t(c['raw']), _t(c['unit_offset'])
    out = kernels.mark4_decode(raw, uo, nframe, c['nchan'], c['fanout'],
                               c['ft'], levels.sign_magnitude(), fill, start,
                               want.shape[0])
    assert np.array_equal(out.cpu().numpy().view('u4'), want.view('u4'))
    full = oracle_frames(c, 0.0, 0, None)
    dst = c['raw'].copy()
    for f in range(nframe):
        if c['unit_offset'][f] >= 0:
            o = c['truth'][f]
            dst[o:o + c['payload_nbytes']] = 0
    dst = _t(dst)
    kernels.mark4_encode(_t(full), dst, uo, nframe, c['nchan'], c['fanout'],
                         c['ft'])
    assert np.array_equal(dst.cpu().numpy(), c['raw'])


@pytest.mark.parametrize('name,ntrack', [
    ('sample.m4', 64), ('sample_32track.m4', 32),
    ('sample_32track_fanout2.m4', 32), ('sample_16track.m4', 16),
    ('sample_64track_fanout2_ft.m4', 64)])
def test_mark4_sample_files(sample_outputs, name, ntrack):
    """Frames located, header-parsed and decoded on the GPU == reference
    Mark4Frame.data (incl. the header-overwritten fill region)."""
    tag = name.replace('.', '_')
    raw = sample_bytes(name)
    off0 = int(sample_outputs[tag + '_offset0'])
    want = sample_outputs[tag + '_data']
    _, fanout, nchan, _, spf = [int(v) for v in sample_outputs[tag + '_geom']]
    nframe = want.shape[0] // spf
    frame_nbytes = ntrack * 2500
    dev = _t(raw[off0:off0 + nframe * frame_nbytes])     # re-based, aligned
    words5, uo = kernels.mark4_scan(dev, nframe, ntrack, track=0)
    assert bool((uo >= 0).all())
    ft = name.endswith('_ft.m4')
    out = kernels.mark4_decode(dev, uo, nframe, nchan, fanout, ft,
                               levels.sign_magnitude(), -7.0)
    assert np.array_equal(out.cpu().numpy(), want)
    # header words of track 0 against the oracle's stream2words
    dt = codec.MARK4_WORD_DTYPE[ntrack]
    for f in range(nframe):
        st = raw[off0 + f * frame_nbytes:][:ntrack * 20].view(dt)
        w = headers.mark4_stream2words(st)
        assert np.array_equal(words5[f].cpu().numpy().view(np.uint32),
                              w[:, 0])


def test_mark4_scan_invalid():
    c = make_frames('8_4', 3, (), 'scan')
    raw = np.zeros(3 * 160000, np.uint8)
    for f in range(3):
        o = c['truth'][f] - 1280
        raw[f * 160000:(f + 1) * 160000] = c['raw'][o:o + 160000]
    st = raw.view('<u8')
    st[:160] = 0                       # frame 0: no error flags
    st[20000:20160] = 0
    st[20000 + 50] = 1 << 37          # frame 1: an error flag on track 37
    st[40000:40160] = 0
    st[40000 + 52] = 1                 # frame 2: bit just outside the flags
    _, uo = kernels.mark4_scan(_t(raw), 3, 64)
    assert uo.cpu().tolist() == [1280, -1, 2 * 160000 + 1280]


@pytest.mark.parametrize('case', int8_cases.CASES, ids=lambda c: c[0])
def test_int8_transposed(case):
    c = int8_cases.make_case(case)
    want = int8_cases.oracle_decode(c)
    out = torch.full((want.size,), float('nan'), dtype=torch.float32,
                     device=DEV)
    kernels.decode_int8_transposed(
        _t(c['raw']), _t(c['unit_offset']), c['nunit'], c['nrow'], c['ncol'],
        c['ib'], _t(c['col_begin']), _t(c['col_end']), _t(c['out_col0']), out)
    assert np.array_equal(out.cpu().numpy().reshape(want.shape), want,
                          equal_nan=True)
    full = dict(c, col_begin=np.zeros(c['nunit'], np.int64),
                col_end=np.full(c['nunit'], c['ncol'], np.int64),
                out_col0=np.arange(c['nunit'], dtype=np.int64) * c['ncol'],
                ncols_out=c['nunit'] * c['ncol'])
    data = int8_cases.oracle_decode(full, fill=0.0)
    for dtype in (np.float32, np.float64):
        dst = c['raw'].copy()
        for u in range(c['nunit']):
            if c['unit_offset'][u] >= 0:
                dst[c['truth'][u]:c['truth'][u] + c['unit_nbytes']] = 0
        dst = _t(dst)
        kernels.encode_int8_transposed(_t(data.astype(dtype)), dst,
                                       _t(c['unit_offset']), c['nunit'],
                                       c['nrow'], c['ncol'], c['ib'])
        assert np.array_equal(dst.cpu().numpy(), c['raw'])


@pytest.mark.parametrize('case', int8_cases.TF_CASES
                         + int8_cases.tf_fuzz_cases(40, seed=78),
                         ids=lambda c: c[0])
def test_int8_timefirst(case):
    c = int8_cases.make_tf_case(case)
    want = int8_cases.oracle_tf_decode(c)
    out = torch.full((want.size,), float('nan'), dtype=torch.float32,
                     device=DEV)
    kernels.decode_int8_timefirst(
        _t(c['raw']), _t(c['unit_offset']), c['nunit'], c['nsample'],
        c['nchan'], c['npol'], c['ib'], _t(c['t_begin']), _t(c['t_end']),
        _t(c['out_t0']), out)
    assert np.array_equal(out.cpu().numpy().reshape(want.shape), want,
                          equal_nan=True)
    full = dict(c, t_begin=np.zeros(c['nunit'], np.int64),
                t_end=np.full(c['nunit'], c['nsample'], np.int64),
                out_t0=np.arange(c['nunit'], dtype=np.int64) * c['nsample'],
                nout=c['nunit'] * c['nsample'])
    data = int8_cases.oracle_tf_decode(full, fill=0.0)
    for dtype in (np.float32, np.float64):
        dst = c['raw'].copy()
        for u in range(c['nunit']):
            if c['unit_offset'][u] >= 0:
                dst[c['truth'][u]:c['truth'][u] + c['unit_nbytes']] = 0
        dst = _t(dst)
        kernels.encode_int8_timefirst(_t(data.astype(dtype)), dst,
                                      _t(c['unit_offset']), c['nunit'],
                                      c['nsample'], c['nchan'], c['npol'],
                                      c['ib'])
        assert np.array_equal(dst.cpu().numpy(), c['raw'])


def test_vdif_scan_sample(sample_outputs):
    raw = sample_bytes('sample.vdif')
    fields_want = sample_outputs['sample_vdif_fields']
    nframe = len(fields_want)
    slot = torch.full((1024,), -1, dtype=torch.int32, device=DEV)
    slot[:8] = torch.arange(8, dtype=torch.int32)
    fields, uo, bad = kernels.vdif_scan(_t(raw), nframe, 5032, 32, 8, slot, 8)
    f = fields.cpu().numpy()
    assert np.array_equal(f[:12].T, fields_want[:, :12])
    assert np.array_equal(f[12], fields_want[:, 12])
    assert int(bad.item()) == 0
    # file order of thread ids is 1,3,5,7,0,2,4,6 -> slots sorted by id
    tid = fields_want[:, 10]
    want_uo = np.empty(16, np.int64)
    for i in range(nframe):
        want_uo[(i // 8) * 8 + tid[i]] = i * 5032 + 32
    assert np.array_equal(uo.cpu().numpy(), want_uo)
    # decode through the table == reference VDIFFrameSet data
    out = kernels.decode_bitfield(_t(raw), uo, 2, 8, 5000, 2, 1, False, 0,
                                  levels.offset_binary(2))
    assert np.array_equal(out.cpu().numpy(), sample_outputs['sample_vdif_data'])
    # subset of threads, invalid frames, and an inconsistent set
    raw2 = raw.copy()
    raw2[5032 * 2 + 3] |= 0x80                 # frame 2 (thread 5) invalid
    slot2 = torch.full((1024,), -1, dtype=torch.int32, device=DEV)
    slot2[5], slot2[2] = 0, 1
    _, uo2, bad2 = kernels.vdif_scan(_t(raw2), nframe, 5032, 32, 8, slot2, 2)
    assert uo2.cpu().tolist() == [-1, 5 * 5032 + 32,
                                  10 * 5032 + 32, 13 * 5032 + 32]
    assert int(bad2.item()) == 0
    raw3 = raw.copy()
    raw3[5032 * 3 + 4] ^= 1                    # frame_nr of one frame differs
    _, _, bad3 = kernels.vdif_scan(_t(raw3), nframe, 5032, 32, 8, slot, 8)
    assert int(bad3.item()) == 1


def test_mark5b_scan_sample(sample_outputs):
    raw = sample_bytes('sample.m5b').copy()
    raw[10016 + 16:2 * 10016].view('<u4')[:] = 0x11223344   # frame 1 invalid
    raw[2 * 10016 + 16:3 * 10016].view('<u4')[:] = 0x11223344
    raw[2 * 10016 + 16 + 4 * 2499] ^= 0x10     # frame 2: last word differs
    fields, uo = kernels.mark5b_scan(_t(raw), 4)
    f = fields.cpu().numpy()
    want = sample_outputs['sample_m5b_fields']
    assert np.array_equal(f[:11].T.astype(np.int64) & 0xffffffff,
                          want & 0xffffffff)
    assert f[11].tolist() == [1, 0, 1, 1]
    assert uo.cpu().tolist() == [16, -1, 2 * 10016 + 16, 3 * 10016 + 16]
    out = kernels.decode_bitfield(_t(raw), uo, 4, 1, 10000, 2, 8, False, 0,
                                  levels.mark5b(2), fill_value=-999.)
    assert np.array_equal(out.cpu().numpy()[:, 0],
                          stream.mark5b_read(raw, 8, fill_value=-999.))


def test_scan_index_checks(sample_outputs):
    """Frame-index checks folded into the scan kernels == the reference's
    `_get_index` arithmetic (vdif/base.py:386-390, mark5b/base.py:206-213)
    applied by the oracle to every header."""
    # VDIF: sample.vdif has 2 frame sets of 8 threads, frame_nr 0 and 1
    raw = sample_bytes('sample.vdif')
    f = sample_outputs['sample_vdif_fields']
    sec0, fnr0 = int(f[0, 2]), int(f[0, 4])
    slot = torch.full((1024,), -1, dtype=torch.int32, device=DEV)
    slot[:8] = torch.arange(8, dtype=torch.int32)
    bad = kernels.new_counter(DEV)
    fields, _, _ = kernels.vdif_scan(_t(raw), 16, 5032, 32, 8, slot, 8,
                                     check=(0, sec0, fnr0, 1600), bad=bad,
                                     want_fields=False)
    assert fields is None and int(bad.item()) == 0
    # the counter accumulates: a wrong expected position flags both sets
    kernels.vdif_scan(_t(raw), 16, 5032, 32, 8, slot, 8,
                      check=(5, sec0, fnr0, 1600), bad=bad)
    assert int(bad.item()) == 2
    # second set one second later but same frame_nr: index = fps + 0, not 1
    raw2 = raw.copy()
    w = raw2[8 * 5032:].view('<u4')
    for k in range(8):
        w[k * 1258] += 1                        # seconds of set 1
        w[k * 1258 + 1] -= 1                    # frame_nr 1 -> 0
    bad2 = kernels.new_counter(DEV)
    kernels.vdif_scan(_t(raw2), 16, 5032, 32, 8, slot, 8,
                      check=(0, sec0, fnr0, 1600), bad=bad2)
    assert int(bad2.item()) == 1
    kernels.vdif_scan(_t(raw2), 16, 5032, 32, 8, slot, 8,
                      check=(0, sec0, fnr0, 1), bad=bad2)   # 1 frame / s: ok
    assert int(bad2.item()) == 1
    # Mark 5B: sample.m5b, frames in sequence
    raw = sample_bytes('sample.m5b')
    hd = headers.mark5b_parse_batch(raw, 4)
    jd0, s0, n0 = int(hd['jday'][0]), int(hd['seconds'][0]), \
        int(hd['frame_nr'][0])
    bad = kernels.new_counter(DEV)
    kernels.mark5b_scan(_t(raw), 4, check=(0, jd0, s0, n0, 6400), bad=bad,
                        want_fields=False)
    assert int(bad.item()) == 0
    raw2 = raw.copy()
    raw2[2 * 10016:2 * 10016 + 4] = 0            # sync word of frame 2 gone
    raw2[3 * 10016 + 4] ^= 2                     # frame_nr of frame 3 differs
    kernels.mark5b_scan(_t(raw2), 4, check=(0, jd0, s0, n0, 6400), bad=bad)
    assert int(bad.item()) == 2
    # jday wraps modulo 1000: header0 at day 999, frames at day 0
    kernels.mark5b_scan(_t(raw), 4, check=(
        86400 * 6400, (jd0 - 1) % 1000, s0, n0, 6400), bad=bad)
    assert int(bad.item()) == 2
    # Mark 4: sample.m4 (2014-06-16T07:38:12.47500, 2.5 ms per frame)
    raw = sample_bytes('sample.m4')[0xa88:][:2 * 160000]
    tick0 = ((7 * 60 + 38) * 60 + 12) * 4000 + 1900
    bad = kernels.new_counter(DEV)
    words5, uo = kernels.mark4_scan(_t(raw), 2, 64, check=(0, 56824, tick0, 10),
                                    bad=bad, want_words=False)
    assert words5 is None and int(bad.item()) == 0
    kernels.mark4_scan(_t(raw), 2, 64, check=(1, 56824, tick0, 10), bad=bad)
    assert int(bad.item()) == 2
    kernels.mark4_scan(_t(raw), 2, 64, check=(0, 56824 - 365, tick0, 10),
                       bad=bad)                  # a year earlier: year digit
    assert int(bad.item()) == 4


def test_fuzz_int8_transposed():
    for case in int8_cases.fuzz_cases(150, seed=78):
        test_int8_transposed(case)


def test_frames_assemble():
    """bb_frames_assemble == the numpy restatement used by the CPU backend:
    headers in place, unit offsets, Mark 5B-style payload fill of invalid
    frames (mark5b/frame.py:126-133) and several units per frame (MKBF)."""
    import cpu_backend
    rng = np.random.default_rng(5)
    for nframe, hn, frame, payload, per, ustride, use_valid in (
            (7, 16, 10016, 10000, 1, 0, True),
            (5, 32, 8032, 8000, 1, 0, False),
            (3, 4096, 4096 + 8192, 8192, 4, 2048, False),
            (2, 1280, 160000, 0, 1, 0, False),
            (4, 6, 6 + 40, 40, 1, 0, False)):          # unaligned header
        hdr = rng.integers(0, 256, (nframe, hn), dtype=np.uint8)
        valid = (rng.random(nframe) > 0.4).astype(np.uint8) \
            if use_valid else None
        want_f, want_uo = cpu_backend._frames_assemble(
            torch.from_numpy(hdr), frame, payload,
            None if valid is None else torch.from_numpy(valid),
            0x11223344, per, ustride)
        got_f, got_uo = kernels.frames_assemble(
            _t(hdr).view(nframe, hn), frame, payload,
            None if valid is None else _t(valid), 0x11223344, per, ustride)
        assert torch.equal(got_uo.cpu(), want_uo)
        # (only what the kernel is meant to write is read back: the payloads
        # of valid frames are left for the encode that follows)
        want = want_f.numpy()
        assert np.array_equal(got_f[:, :hn].cpu().numpy(), want[:, :hn])
        if valid is not None:
            bad = valid == 0
            assert np.array_equal(got_f[torch.from_numpy(bad).to(DEV)]
                                  .cpu().numpy(), want[bad])
