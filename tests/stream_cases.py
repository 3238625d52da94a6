"""Stream-level checks shared by the CPU-backend run (host logic) and the GPU
run (parity proper): each function takes nothing and asserts, comparing the
public API of baseband_b200 with the oracle / golden fixtures."""
import io
import os

import numpy as np
import pytest
import torch

import baseband_b200 as bb
from baseband_b200 import synthetic
from oracle import stream as ostream
from conftest import GOLDEN, sample_path

OUT = np.load(os.path.join(GOLDEN, 'sample_outputs.npz'))


def _same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape,
                                                       a.dtype, b.dtype)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


# ------------------------------------------------------------------ VDIF
def vdif_sample_read():
    """BASELINE config 1: sample.vdif through open(...,'rs').read()."""
    want = OUT['sample_vdif_data']
    with bb.vdif.open(sample_path('sample.vdif'), 'rs') as fh:
        assert fh.shape == (40000, 8)
        assert fh.sample_shape == (8,)
        assert fh.sample_shape.nthread == 8
        assert fh.samples_per_frame == 20000
        assert fh.bps == 2 and not fh.complex_data
        assert fh.sample_rate == 32e6
        assert fh.start_time.isot.startswith('2014-06-16T05:56:07.000000000')
        data = fh.read()
        assert fh.tell() == 40000
        _same(data, want[:, :, 0])
        # values the reference's own test asserts (test_vdif.py:930-931)
        assert np.all(data[:12, 0].astype(int)
                      == np.array([-1, -1, 3, -1, 1, -1, 3, -1, 1, 3, -1, 1]))
        fh.seek(19990)
        part = fh.read(20)
        _same(part, want[19990:20010, :, 0])
        fh.seek(-7, 2)
        _same(fh.read(), want[-7:, :, 0])
        try:
            fh.seek(-3, 2)
            fh.read(4)
        except EOFError:
            pass
        else:
            raise AssertionError('expected EOFError')
    with bb.vdif.open(sample_path('sample.vdif'), 'rs', squeeze=False) as fh:
        assert fh.sample_shape == (8, 1)
        _same(fh.read(), want)


def vdif_sample_subset():
    want = OUT['sample_vdif_data'][:, :, 0]
    for subset, pick in [(3, want[:, 3]), ([1, 6], want[:, [1, 6]]),
                         (slice(2, 7, 2), want[:, 2:7:2]),
                         ([5], want[:, [5]])]:
        with bb.vdif.open(sample_path('sample.vdif'), 'rs',
                          subset=subset) as fh:
            assert fh.shape == pick.shape
            fh.seek(10)
            _same(fh.read(30000), pick[10:30010])
    with bb.vdif.open(sample_path('sample.vdif'), 'rs', subset=(2, 0),
                      squeeze=False) as fh:
        _same(fh.read(5), OUT['sample_vdif_data'][:5, 2, 0])


def vdif_other_samples():
    for name in ('sample_vlbi.vdif', 'sample_mwa.vdif',
                 'sample_arochime.vdif', 'sample_bps1.vdif'):
        tag = name.replace('.', '_')
        want = OUT[tag + '_data']
        kw = {}
        if name in ('sample_mwa.vdif', 'sample_bps1.vdif',
                    'sample_arochime.vdif'):
            kw['sample_rate'] = 1e6     # legacy / EDV 0 headers hold no rate
        with bb.vdif.open(sample_path(name), 'rs', squeeze=False,
                          **kw) as fh:
            data = fh.read()
            _same(data, want[:data.shape[0]])
            assert data.shape[0] == want.shape[0]


def vdif_invalid_fill():
    want = OUT['sample_vdif_set0_invalid_1_4_7_fill_m999']
    raw = np.fromfile(sample_path('sample.vdif'), np.uint8).copy()
    frames = raw.reshape(16, 5032)
    # reference test marks the frames of threads 1, 4, 7 of set 0 invalid
    tid = (frames[:, 12:16].view('<u4')[:, 0] >> 16) & 0x3ff
    for i in range(8):
        if tid[i] in (1, 4, 7):
            frames[i, 3] |= 0x80
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', squeeze=False,
                      fill_value=-999.) as fh:
        _same(fh.read(20000), want)


def vdif_synthetic_chunked(nset=23):
    """Config 2 geometry, tiny chunks so the pipeline runs many stages."""
    raw = synthetic.vdif_stream(nset, 16, 8000, seed=5, invalid=[3, 77, 78])
    want = ostream.vdif_read(raw, fill_value=2.5)[:, :, 0]
    for chunk in (16 * 8032 * 3, 1 << 30):
        with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=64e6,
                          fill_value=2.5, chunk_nbytes=chunk) as fh:
            assert fh.shape == want.shape
            _same(fh.read(), want)
            fh.seek(31999)
            _same(fh.read(5 * 32000 + 3), want[31999:31999 + 5 * 32000 + 3])
            out = np.empty((70001, 16), np.float32)
            fh.seek(123)
            assert fh.read(out=out) is out
            _same(out, want[123:70124])


def vdif_device_output(dev):
    raw = synthetic.vdif_stream(7, 16, 8000, seed=9)
    want = ostream.vdif_read(raw)[:, :, 0]
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=64e6,
                      device=dev, chunk_nbytes=16 * 8032 * 2) as fh:
        data = fh.read()
        assert isinstance(data, torch.Tensor)
        assert str(data.device) == str(torch.device(dev))
        _same(data.cpu().numpy(), want)
        fh.seek(1001)
        out = torch.empty((64000, 16), dtype=torch.float32, device=dev)
        fh.read(out=out)
        _same(out.cpu().numpy(), want[1001:65001])


def vdif_write_roundtrip():
    """Stream writer: decode -> write -> bytes identical to the input."""
    raw = synthetic.vdif_stream(5, 4, 8000, seed=3, thread_order=[0, 1, 2, 3],
                                edv=1)
    # EDV 1 needs sample rate + sync pattern in the header
    w = raw.reshape(-1, 8032)[:, :32].view('<u4')
    w[:, 4] = (1 << 24) | (1 << 23) | 32        # 32 MHz complex => 64 MHz real
    w[:, 5] = 0xACABFEED
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs') as fr:
        assert fr.sample_rate == 64e6
        data = fr.read()
        header0 = fr.header0
    buf = io.BytesIO()
    fw = bb.vdif.open(buf, 'ws', header0=header0, nthread=4)
    fw.write(data[:1000])
    fw.write(data[1000:90000])
    fw.write(data[90000:])
    fw._flush(final=False)
    got = np.frombuffer(buf.getvalue(), np.uint8)
    fw.fh_raw = io.BytesIO()        # keep ``buf`` readable after close
    fw.close()
    _same(got, raw)
    # float64 input and invalid + partial last frame
    buf = io.BytesIO()
    import warnings
    fw = bb.vdif.open(buf, 'ws', nthread=4, edv=1, time='2014-06-16T05:56:07',
                      samples_per_frame=32000, bps=2, nchan=1,
                      sample_rate=64e6, station='me')
    fw.write(data[:32000].astype(np.float64))
    fw.write(data[32000:40000], valid=False)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        fw._flush(final=True)
    assert any('partial buffer' in str(r.message) for r in rec)
    got = np.frombuffer(buf.getvalue(), np.uint8).reshape(8, 8032)
    _same(got[:4, 32:], raw.reshape(-1, 8032)[:4, 32:])
    inv = (got[:, 3] >> 7).astype(bool)
    assert list(inv) == [False] * 4 + [True] * 4
    back = bb.vdif.open(io.BytesIO(buf.getvalue()), 'rs')
    assert back.header0.station == 'me'
    assert back.start_time.isot.startswith('2014-06-16T05:56:07.0')
    d2 = back.read()
    _same(d2[:32000], data[:32000])
    assert np.all(d2[32000:] == 0)


def vdif_frameset_api():
    want = OUT['sample_vdif_data']
    with bb.vdif.open(sample_path('sample.vdif'), 'rb') as fh:
        header = fh.read_header()
        assert header['thread_id'] == 1 and header.edv == 3
        assert header.samples_per_frame == 20000
        fh.seek(0)
        fs = fh.read_frameset()
        assert fs.shape == (20000, 8, 1)
        assert list(fs['thread_id']) == list(range(8))
        _same(fs.data, want[:20000])
        _same(fs[5:11, 3], want[5:11, 3])
        _same(fs[7], want[7])
        _same(fs[::1000, 2:4, 0], want[:20000:1000, 2:4, 0])
        fs.fill_value = -7.
        fs.frames[2].header.mutable = True
        fs.frames[2].valid = False
        d = fs.data
        assert np.all(d[:, 2] == -7.) and np.array_equal(d[:, 3], want[:20000, 3])
        fs.frames[2].valid = True
        fs2 = bb.vdif.VDIFFrameSet.fromdata(fs.data, fs.frames[0].header)
        for a, b in zip(fs.frames, fs2.frames):
            assert np.array_equal(a.payload.words, b.payload.words)
        fs3 = fh.read_frameset([3, 5])
        _same(fs3.data, want[20000:, [3, 5]])
    frame = bb.vdif.VDIFFrame.fromfile(open(sample_path('sample.vdif'), 'rb'))
    _same(frame.data, want[:20000, 1])
    _same(frame[11:17], want[11:17, 1])
    pl = frame.payload
    assert pl.shape == (20000, 1)
    pl2 = bb.vdif.VDIFPayload.fromdata(pl.data, frame.header)
    assert pl2 == pl
    pl2[10:13] = np.array([[3.316505], [1.], [-1.]], np.float32)
    assert np.array_equal(pl2[10:13].ravel(),
                          np.array([3.316505, 1., -1.], np.float32))
    assert np.array_equal(pl2[:10], pl[:10]) and np.array_equal(pl2[13:], pl[13:])


# ------------------------------------------------------------------ Mark 5B
def mark5b_sample_read():
    want = OUT['sample_m5b_data']
    with bb.mark5b.open(sample_path('sample.m5b'), 'rs', sample_rate=32e6,
                        kday=56000, nchan=8) as fh:
        assert fh.shape == (20000, 8)
        assert fh.sample_shape.nchan == 8
        assert fh.samples_per_frame == 5000
        assert fh.start_time.isot == '2014-06-13T05:30:01.000000000'
        assert fh.header0['frame_nr'] == 0 and fh.header0.jday == 821
        data = fh.read()
        _same(data, want)
        # values asserted by the reference (test_mark5b.py:172-175)
        assert np.all(data[:3].astype(int) == np.array(
            [[-3, -1, +1, -1, +3, -3, -3, +3],
             [-3, +3, -1, +3, -1, -1, -1, +1],
             [+3, -1, +3, +3, +1, -1, +3, -1]]))
        fh.seek(4990)
        _same(fh.read(30), want[4990:5020])
        assert fh.stop_time.isot == '2014-06-13T05:30:01.000625000'
    with bb.mark5b.open(sample_path('sample.m5b'), 'rs', sample_rate=32e6,
                        ref_time='2014-01-01T00:00:00', nchan=8,
                        subset=[1, 3], device=None) as fh:
        assert fh.header0.kday == 56000
        _same(fh.read(77), want[:77, [1, 3]])
    # file-level API
    with bb.mark5b.open(sample_path('sample.m5b'), 'rb', kday=56000,
                        nchan=8) as fb:
        frame = fb.read_frame()
        assert frame.valid and frame.shape == (5000, 8)
        _same(frame.data, want[:5000])
        assert frame.header['crc'] == 38749
        h = frame.header.copy()
        h.mutable = True
        h.update(time=frame.header.time, frame_rate=6400.)
        assert h == frame.header


def mark5b_invalid_frames():
    """Config 5: synthetic stream with fill-pattern frames."""
    raw, valid = synthetic.mark5b_stream(40, invalid_fraction=0.2, seed=77)
    mask = ostream.mark5b_valid_mask(raw)
    assert np.array_equal(mask, valid)
    assert 0 < (~mask).sum() < 40
    for fill in (0., -999.):
        want = ostream.mark5b_read(raw, 16, fill_value=fill)
        with bb.mark5b.open(io.BytesIO(raw.tobytes()), 'rs', nchan=16,
                            sample_rate=16e6, kday=56000, fill_value=fill,
                            chunk_nbytes=7 * 10016) as fh:
            assert fh.shape == want.shape
            _same(fh.read(), want)
            fh.seek(2499)
            _same(fh.read(2500 * 9 + 2), want[2499:2499 + 2500 * 9 + 2])


def mark5b_write_roundtrip():
    raw = np.fromfile(sample_path('sample.m5b'), np.uint8)
    with bb.mark5b.open(sample_path('sample.m5b'), 'rs', sample_rate=32e6,
                        kday=56000, nchan=8) as fh:
        data = fh.read()
        header0 = fh.header0
    buf = io.BytesIO()
    fw = bb.mark5b.open(buf, 'ws', header0=header0, sample_rate=32e6,
                        nchan=8)
    fw.write(data[:7000])
    fw.write(data[7000:])
    got = np.frombuffer(buf.getvalue(), np.uint8)
    # byte-for-byte the reference's rewrite test (test_mark5b.py:690-701),
    # apart from the 'user' header field which header0 carries for all frames
    _same(got, raw)
    # from keywords, one invalid frame -> fill pattern on disk
    buf = io.BytesIO()
    fw = bb.mark5b.open(buf, 'ws', time='2014-06-13T05:30:01', nchan=8,
                        sample_rate=32e6)
    fw.write(data[:5000])
    fw.write(data[5000:10000], valid=False)
    fw.write(data[10000:15000].astype(np.float64))
    got = np.frombuffer(buf.getvalue(), np.uint8).reshape(3, 10016)
    _same(got[0, 16:], raw[16:10016])
    assert np.all(got[1, 16:].view('<u4') == 0x11223344)
    _same(got[2, 16:], raw.reshape(4, 10016)[2, 16:])
    _same(got[:, 8:16], raw.reshape(4, 10016)[:3, 8:16])     # time code + CRC


# ------------------------------------------------------------------ Mark 4
M4_SAMPLES = [('sample.m4', 64), ('sample_32track.m4', 32),
              ('sample_32track_fanout2.m4', 32), ('sample_16track.m4', 16),
              ('sample_64track_fanout2_ft.m4', 64)]


def mark4_sample_read():
    for name, ntrack in M4_SAMPLES:
        tag = name.replace('.', '_')
        want = OUT[tag + '_data']
        for kw in ({'ntrack': ntrack}, {}):
            with bb.mark4.open(sample_path(name), 'rs', decade=2010,
                               fill_value=-7., **kw) as fh:
                assert fh.shape == want.shape, (name, fh.shape)
                assert fh.fh_raw.ntrack == ntrack
                assert fh._file_offset0 == int(OUT[tag + '_offset0'])
                data = fh.read()
                ref = want                  # golden made with fill_value -7
                _same(data, ref)
                spf = fh.samples_per_frame
                fh.seek(spf - 3)
                n = min(703, fh.shape[0] - (spf - 3))
                _same(fh.read(n), ref[spf - 3:spf - 3 + n])
    with bb.mark4.open(sample_path('sample.m4'), 'rs', ntrack=64,
                       ref_time='2013-01-01T00:00:00') as fh:
        assert fh.sample_rate == 32e6
        assert fh.start_time.isot == '2014-06-16T07:38:12.475000000'
        assert fh.header0.decade == 2010
        data = fh.read(80000)
        # values asserted by the reference (test_mark4.py:324-327, :757-765)
        assert np.all(data[640:642].astype(int) == np.array(
            [[-1, +1, +1, -3, -3, -3, +1, -1],
             [+1, +1, -3, +1, +1, -3, -1, -1]]))
        assert np.all(data[:640] == 0.)


def mark4_frame_api():
    want = OUT['sample_m4_data']
    with bb.mark4.open(sample_path('sample.m4'), 'rb', ntrack=64,
                       decade=2010) as fb:
        assert fb.locate_frame() == 0xa88
        frame = fb.read_frame()
    assert frame.shape == (80000, 8) and frame.valid
    assert np.all(frame[3] == 0.)
    frame.fill_value = -7.
    _same(frame.data, want[:80000])
    _same(frame[630:650], want[630:650])
    _same(frame[700:720, 3], want[700:720, 3])
    _same(frame[635::7][:5], want[635:80000:7][:5])
    frame2 = bb.mark4.Mark4Frame.fromdata(frame.data, frame.header)
    assert np.array_equal(frame2.payload.words, frame.payload.words)
    buf = io.BytesIO()
    frame2.tofile(buf)
    raw = np.fromfile(sample_path('sample.m4'), np.uint8)[0xa88:0xa88 + 160000]
    _same(np.frombuffer(buf.getvalue(), np.uint8), raw)
    frame.header.mutable = True
    frame.valid = False
    frame.fill_value = 9.
    assert np.all(frame[1000:1010] == 9.)


def mark4_synthetic_and_write():
    """Config 3 geometry: 64 tracks, fanout 4; invalid frame; byte-identical
    rewrite through the stream writer."""
    h0 = bb.mark4.Mark4Header.fromvalues(
        64, time='2014-06-16T07:38:12.475', bps=2, fanout=4, nsb=1,
        system_id=108)
    rng = np.random.default_rng(31)
    data = rng.choice(np.array([-3.316505, -1., 1., 3.316505], np.float32),
                      size=(7 * 80000, 8))
    buf = io.BytesIO()
    fw = bb.mark4.open(buf, 'ws', header0=h0, sample_rate=32e6)
    fw.write(data[:100000])
    fw.write(data[100000:3 * 80000])
    fw.write(data[3 * 80000:4 * 80000], valid=False)
    fw.write(data[4 * 80000:])
    raw = np.frombuffer(buf.getvalue(), np.uint8)
    assert raw.size == 7 * 160000
    # oracle reading of what was written
    want = ostream.mark4_read(raw, 64, fill_value=-2.)
    expect = data.copy()
    for f in range(7):
        expect[f * 80000:f * 80000 + 640] = -2.
    expect[3 * 80000:4 * 80000] = -2.
    _same(want, expect)
    with bb.mark4.open(io.BytesIO(raw.tobytes()), 'rs', ntrack=64,
                       decade=2010, fill_value=-2.,
                       chunk_nbytes=2 * 160000) as fh:
        assert fh.sample_rate == 32e6
        _same(fh.read(), expect)
        assert fh.stop_time.isot == '2014-06-16T07:38:12.492500000'
    # the reference's own sample, rewritten byte for byte
    src = np.fromfile(sample_path('sample.m4'), np.uint8)[0xa88:]
    with bb.mark4.open(sample_path('sample.m4'), 'rs', ntrack=64,
                       decade=2010) as fh:
        header0, sdata = fh.header0, fh.read()
    buf = io.BytesIO()
    fw = bb.mark4.open(buf, 'ws', header0=header0, sample_rate=32e6)
    fw.write(sdata)
    got = np.frombuffer(buf.getvalue(), np.uint8)
    _same(got, src[:got.size])
    assert got.size == 2 * 160000


def mark4_time_code_rollover():
    """Streams written across a year end and across 29 February are read
    back with verify: the time code the scan kernel computes for every frame
    position (year digit, day of year, h:m:s.ms) must be the one the writer
    generated (mark4/header.py:223-262)."""
    rng = np.random.default_rng(5)
    for start, decade, stop in (
            ('2020-12-31T23:59:59.9900', 2020, '2021-01-01T00:00:00.01'),
            ('2020-02-28T23:59:59.9875', 2020, '2020-02-29T00:00:00.0075'),
            ('2019-02-28T23:59:59.9950', 2010, '2019-03-01T00:00:00.0150'),
            ('2099-12-31T23:59:59.9900', 2090, '2100-01-01T00:00:00.01')):
        h0 = bb.mark4.Mark4Header.fromvalues(
            32, time=start, bps=2, fanout=4, nsb=1, system_id=108)
        data = rng.choice(np.array([-3.316505, -1., 1., 3.316505],
                                   np.float32), size=(8 * 80000, 4))
        buf = io.BytesIO()
        fw = bb.mark4.open(buf, 'ws', header0=h0, sample_rate=32e6)
        fw.write(data)
        raw = np.frombuffer(buf.getvalue(), np.uint8)
        want = ostream.mark4_read(raw, 32)
        with bb.mark4.open(io.BytesIO(raw.tobytes()), 'rs', ntrack=32,
                           decade=decade, chunk_nbytes=3 * 80000) as fh:
            assert fh.stop_time.isot.startswith(stop), fh.stop_time.isot
            _same(fh.read(), want)
        # a frame out of place is caught by the same check, and the stream
        # is then read through the GPU frame index (time codes across the
        # year end / leap day included)
        import warnings
        frames = raw.reshape(8, -1).copy()
        frames[[5, 6]] = frames[[6, 5]]
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter('always')
            with bb.mark4.open(io.BytesIO(frames.tobytes()), 'rs', ntrack=32,
                               decade=decade) as fh:
                _same(fh.read(), want)
                assert fh._index is not None


# ------------------------------------------------------------------ GUPPI
def _guppi_expected(frames, overlap, start, count):
    """Reference read loop on per-frame decoded arrays (full length incl.
    overlap): [start, len) of the first frame, [overlap, len) of later."""
    stride = frames.shape[1] - overlap
    total = stride * len(frames) + overlap
    normal_end = total - overlap
    pieces, pos, done = [], start, 0
    while done < count:
        if normal_end <= pos < total:
            index, local = divmod(normal_end - 1, stride)
            local += 1 + pos - normal_end
        else:
            index, local = divmod(pos, stride)
        n = min(count - done, frames.shape[1] - local)
        pieces.append(frames[index][local:local + n])
        done += n
        pos += n
    return np.concatenate(pieces)


def guppi_sample_read():
    frames = OUT['sample_puppi_frames']            # (4, 1024, 2, 4)
    with bb.guppi.open(sample_path('sample_puppi.raw'), 'rs') as fh:
        assert fh.shape == (3904, 2, 4)
        assert fh.sample_shape.npol == 2 and fh.sample_shape.nchan == 4
        assert fh.samples_per_frame == 960
        assert fh.sample_rate == 250.
        assert fh.complex_data and fh.bps == 8
        assert fh.start_time.isot == '2018-01-14T14:11:33.000000000'
        data = fh.read()
        _same(data, _guppi_expected(frames, 64, 0, 3904))
        for start, count in ((0, 1), (959, 3), (1000, 1500), (3839, 65),
                             (3850, 54), (30, 3000)):
            fh.seek(start)
            _same(fh.read(count), _guppi_expected(frames, 64, start, count))
        # values the reference asserts (test_guppi.py:236-249 region)
        assert data[0, 0, 0] == frames[0][0, 0, 0]
    with bb.guppi.open(sample_path('sample_puppi.raw'), 'rs',
                       subset=(0, [1, 3]), chunk_nbytes=1) as fh:
        fh.seek(900)
        _same(fh.read(200),
              _guppi_expected(frames, 64, 900, 200)[:, 0][:, [1, 3]])
    with bb.guppi.open(sample_path('sample_puppi.raw'), 'rb') as fb:
        frame = fb.read_frame()
        assert frame.shape == (1024, 2, 4)
        _same(frame.data, frames[0])
        _same(frame[100:200, 1], frames[0][100:200, 1])
        _same(frame.payload[37], frames[0][37])
        fb.seek(3 * 22784)
        f3 = fb.read_frame(memmap=False)
        _same(f3.data, frames[3])
        pl = bb.guppi.GUPPIPayload(f3.payload.words, sample_shape=(2, 4),
                                   bps=8, complex_data=True,
                                   channels_first=False)
        _same(pl.data, OUT['sample_puppi_frame3_timefirst'])
        for cf in (True, False):
            pl2 = bb.guppi.GUPPIPayload.fromdata(
                frames[2], bps=8, channels_first=cf)
            _same(pl2.data, frames[2])
            pl2[10:12, 1] = np.array([[1 - 2j] * 4, [3 + 4j] * 4],
                                     np.complex64)
            assert np.all(pl2[10, 1] == 1 - 2j) and np.all(pl2[11, 1] == 3 + 4j)
            _same(pl2[12:], frames[2][12:])
        pl3 = bb.guppi.GUPPIPayload.fromdata(frames[3], f3.header)
        assert np.array_equal(pl3.words, f3.payload.words)


def guppi_synthetic_and_write():
    """Config 4 geometry (scaled): 512 channels, 2 pol, overlap."""
    raw, truth = synthetic.guppi_stream(5, nchan=512, npol=2,
                                        samples_per_frame=128, overlap=16,
                                        seed=2)
    want = ostream.guppi_read(raw)
    cube = truth.astype(np.float32).transpose(1, 2, 0, 3)  # (t, pol, ch, 2)
    direct = (cube[..., 0] + 1j * cube[..., 1]).astype(np.complex64)
    _same(want, direct)
    with bb.guppi.open(io.BytesIO(raw.tobytes()), 'rs',
                       chunk_nbytes=2 * (raw.size // 5)) as fh:
        assert fh.shape == want.shape
        _same(fh.read(), want)
        fh.seek(100)
        _same(fh.read(400), want[100:500])
    # writer: no overlap; channels first and time first; read back
    rng = np.random.default_rng(8)
    data = (rng.integers(-128, 128, (3 * 64, 2, 32))
            + 1j * rng.integers(-128, 128, (3 * 64, 2, 32))).astype(
                np.complex64)
    for pktfmt in ('1SFA', 'SIMPLE'):
        buf = io.BytesIO()
        fw = bb.guppi.open(buf, 'ws', time='2018-01-14T14:11:33',
                           sample_rate=250., samples_per_frame=64,
                           sample_shape=(2, 32), pktfmt=pktfmt, pktsize=1024)
        fw.write(data[:100])
        fw.write(data[100:] + 0.25)     # rounds back to the integers
        raw2 = buf.getvalue()
        hdr = bb.guppi.GUPPIHeader.fromfile(io.BytesIO(raw2))
        assert hdr.channels_first == (pktfmt == '1SFA')
        assert len(raw2) == 3 * hdr.frame_nbytes
        _same(ostream.guppi_read(np.frombuffer(raw2, np.uint8)), data)
        with bb.guppi.open(io.BytesIO(raw2), 'rs') as fr:
            assert fr.start_time.isot == '2018-01-14T14:11:33.000000000'
            _same(fr.read(), data)
            h2 = bb.guppi.GUPPIHeader.fromfile(
                io.BytesIO(raw2[2 * hdr.frame_nbytes:]))
            assert fr._get_index(h2) == 2


def vdif_parallel_file_read():
    """Large chunks of plain files are copied out of the page cache by a pool
    of native threads
    (base/stream.py:_parallel_readinto): same bytes, same samples; other
    file-like objects fall back to readinto."""
    import tempfile
    from baseband_b200.base import stream
    raw = synthetic.vdif_stream(40, 4, 1000, seed=5)
    want = ostream.vdif_read(raw)[:, :, 0]
    calls = []
    saved = (stream.PARALLEL_READ_MIN_NBYTES, stream.PARALLEL_READ_THREADS,
             stream._parallel_readinto)

    def spy(fh, offset, view):
        got = saved[2](fh, offset, view)
        calls.append(got)
        return got

    stream.PARALLEL_READ_MIN_NBYTES = 1
    stream.PARALLEL_READ_THREADS = 3
    stream._parallel_readinto = spy
    try:
        with tempfile.NamedTemporaryFile(suffix='.vdif') as tmp:
            tmp.write(raw.tobytes())
            tmp.flush()
            with bb.vdif.open(tmp.name, 'rs', sample_rate=1e6,
                              chunk_nbytes=30000) as fh:
                _same(fh.read(), want)
                fh.seek(12345)
                _same(fh.read(7000), want[12345:19345])
        assert calls and all(c is not None and c > 0 for c in calls)
        calls.clear()
        with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs',
                          sample_rate=1e6) as fh:
            _same(fh.read(), want)
        assert calls and all(c is None for c in calls)
    finally:
        (stream.PARALLEL_READ_MIN_NBYTES, stream.PARALLEL_READ_THREADS,
         stream._parallel_readinto) = saved


def vdif_pageable_out_staged():
    """read(out=<large pageable numpy array>): chunks land in pinned staging
    and are copied on by threads (base/stream.py:_read_to_host)."""
    from baseband_b200.base import stream
    raw = synthetic.vdif_stream(30, 4, 1000, seed=9, invalid=[5])
    want = ostream.vdif_read(raw, fill_value=-1.)[:, :, 0]
    saved = stream.STAGED_HOST_OUT_MIN_NBYTES
    stream.STAGED_HOST_OUT_MIN_NBYTES = 1
    try:
        for chunk in (None, 9000, 40000):
            with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs',
                              sample_rate=1e6, fill_value=-1.,
                              chunk_nbytes=chunk) as fh:
                out = np.full(want.shape, np.nan, np.float32)
                assert fh.read(out=out) is out
                _same(out, want)
                fh.seek(777)
                part = np.full((50001, 4), np.nan, np.float32)
                fh.read(out=part)
                _same(part, want[777:777 + 50001])
                # subset: non-contiguous device piece
            with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs',
                              sample_rate=1e6, fill_value=-1.,
                              subset=[3, 1], chunk_nbytes=chunk) as fh:
                out = np.full((want.shape[0], 2), np.nan, np.float32)
                fh.read(out=out)
                _same(out, want[:, [3, 1]])
    finally:
        stream.STAGED_HOST_OUT_MIN_NBYTES = saved


# ------------------------------------------------------------------ DADA
def dada_sample_read():
    for name in ('sample.dada', 'sample_meerkat.dada', 'sample_mkbf.dada'):
        want = OUT[name.replace('.', '_') + '_data']
        for chunk in (None, 3000):
            with bb.dada.open(sample_path(name), 'rs', squeeze=False,
                              chunk_nbytes=chunk) as fh:
                assert fh.shape == want.shape, (name, fh.shape, want.shape)
                _same(fh.read(), want)
                fh.seek(250)
                _same(fh.read(777), want[250:1027]) if want.shape[0] > 1027 \
                    else _same(fh.read(5), want[250:255])
    with bb.dada.open(sample_path('sample.dada'), 'rs') as fh:
        assert fh.sample_shape == (2,) and fh.sample_shape.npol == 2
        assert fh.sample_rate == 16e6 and fh.complex_data
        assert fh.start_time.isot == '2013-07-02T01:39:20.000000000'
        data = fh.read(12)
        # values asserted by the reference (test_dada.py:180-183)
        assert np.all(data[:3] == np.array(
            [[-38 - 38j, -38 - 38j], [-38 - 38j, -40 + 0j],
             [-105 + 60j, 85 - 15j]], np.complex64))
        assert fh.stop_time.isot == '2013-07-02T01:39:20.001000000'
    with bb.dada.open(sample_path('sample.dada'), 'rb') as fb:
        frame = fb.read_frame(memmap=False)
        assert frame.shape == (16000, 2, 1)
        _same(frame.data, OUT['sample_dada_data'])
        _same(frame[100:110, 1], OUT['sample_dada_data'][100:110, 1])
        pl = bb.dada.DADAPayload.fromdata(frame.data, frame.header)
        assert np.array_equal(pl.words, frame.payload.words)
    # MKBF payload object: heaps decoded and re-encoded on the GPU
    want = OUT['sample_mkbf_dada_data']
    raw = np.fromfile(sample_path('sample_mkbf.dada'), np.uint8)[4096:]
    hdr = bb.dada.DADAHeader.fromfile(open(sample_path('sample_mkbf.dada'),
                                           'rb'))
    h2 = hdr.copy()
    h2.mutable = True
    h2.payload_nbytes = raw.size
    pl = bb.dada.DADAPayload(raw.view('<u4'), header=h2)
    assert type(pl) is bb.dada.MKBFPayload and pl.shape == want.shape
    _same(pl.data, want)
    _same(pl[100:300, 1, 5:9], want[100:300, 1, 5:9])
    pl2 = bb.dada.DADAPayload.fromdata(want, h2)
    assert np.array_equal(pl2.words, pl.words)


def dada_write_roundtrip():
    rng = np.random.default_rng(4)
    data = (rng.integers(-128, 128, (3000, 2, 4))
            + 1j * rng.integers(-128, 128, (3000, 2, 4))).astype(np.complex64)
    buf = io.BytesIO()
    fw = bb.dada.open(buf, 'ws', time='2013-07-02T01:39:20',
                      sample_rate=16e6, samples_per_frame=1000,
                      sample_shape=(2, 4), complex_data=True, bps=8)
    fw.write(data[:1500])
    fw.write(data[1500:])
    raw = buf.getvalue()
    assert len(raw) == 3 * (4096 + 16000)
    _same(ostream.dada_read(np.frombuffer(raw, np.uint8)), data)
    with bb.dada.open(io.BytesIO(raw), 'rs', chunk_nbytes=5000) as fr:
        assert fr.start_time.isot == '2013-07-02T01:39:20.000000000'
        assert fr.shape == (3000, 2, 4)
        _same(fr.read(), data)
        fr.seek(999)
        _same(fr.read(1003), data[999:2002])
    # truncated last frame
    cut = raw[:2 * 20096 + 4096 + 16 * 333 + 7]
    with bb.dada.open(io.BytesIO(cut), 'rs') as fr:
        assert fr.shape[0] == 2333
        _same(fr.read(), data[:2333])


def payload_reference_named_operators():
    """The operator-level names of the reference (base/encoding.py,
    <format>/payload.py module functions and look-up tables) against the
    golden vectors produced by the reference itself."""
    import os
    from baseband_b200.base import encoding
    from baseband_b200.vdif import payload as vp
    from baseband_b200.mark5b import payload as m5p
    from baseband_b200.mark4 import payload as m4p
    from baseband_b200.guppi import payload as gp
    from baseband_b200.dada import payload as dp
    from baseband_b200.gsb import payload as gsp
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden',
                             'codec_vectors.npz'))
    assert encoding.OPTIMAL_2BIT_HIGH == 3.316505
    assert encoding.TWO_BIT_1_SIGMA == 2.174564
    for bps in (1, 2, 4):
        _same(encoding.decoder_levels[bps], g['levels%d' % bps])
        _same(getattr(vp, 'lut%dbit' % bps), g['vdif_lut%d' % bps])
        _same(getattr(vp, 'decode_%dbit' % bps)(g['words32']).ravel(),
              g['vdif_dec%d' % bps].ravel())
    _same(encoding.decode_8bit(g['words32']).ravel(), g['vdif_dec8'].ravel())
    for tag in ('f32', 'f64'):
        vals = g['enc_in_' + tag]
        for bps in (1, 2, 4):
            packed = g['vdif_enc%d_%s' % (bps, tag)].view(np.uint8).ravel()
            _same(getattr(vp, 'encode_%dbit' % bps)(vals).view(
                np.uint8).ravel()[:packed.size], packed)
            # ..._base: one code per value = the packed codes, LSB first
            shifts = np.arange(0, 8, bps, dtype=np.uint8)
            codes = ((packed[:, None] >> shifts) & ((1 << bps) - 1)).ravel()
            got = getattr(encoding, 'encode_%dbit_base' % bps)(vals)
            assert got.dtype == np.uint8 and got.shape == vals.shape
            _same(got.ravel(), codes[:vals.size])
        _same(encoding.encode_8bit(vals).view(np.uint8).ravel(),
              g['vdif_enc8_' + tag].view(np.uint8).ravel())
        for bps in (1, 2):
            _same(getattr(m5p, 'encode_%dbit' % bps)(vals).view(
                np.uint8).ravel(),
                g['m5b_enc%d_%s' % (bps, tag)].view(np.uint8).ravel())
        vals = g['enc_in_finite_' + tag]
        _same(gp.encode_8bit(vals).view(np.uint8).ravel(),
              g['int8_enc_' + tag].view(np.uint8).ravel())
        _same(dp.encode_8bit(vals).view(np.uint8).ravel(),
              g['int8_enc_' + tag].view(np.uint8).ravel())
        _same(gsp.encode_4bit(vals).view(np.uint8).ravel(),
              g['gsb4_enc_' + tag].view(np.uint8).ravel())
    for bps in (1, 2):
        _same(getattr(m5p, 'decode_%dbit' % bps)(g['words32']).ravel(),
              g['m5b_dec%d' % bps].ravel())
    _same(gp.decode_8bit(g['bytes']).ravel(), g['int8_dec'].ravel())
    _same(gsp.decode_8bit(g['bytes']).ravel(), g['int8_dec'].ravel())
    for tag, name in (('2_4', '2chan_2bit_fanout4'),
                      ('4_4', '4chan_2bit_fanout4'),
                      ('8_2', '8chan_2bit_fanout2'),
                      ('8_4', '8chan_2bit_fanout4'),
                      ('16_2ft', '16chan_2bit_fanout2_ft')):
        _same(getattr(m4p, 'decode_' + name)(g['m4_words_' + tag]),
              g['m4_dec_' + tag])
        _same(getattr(m4p, 'encode_' + name)(g['m4_enc_in_%s_f32' % tag]
                                             ).view(np.uint8).ravel(),
              g['m4_enc_%s_f32' % tag].view(np.uint8).ravel())


def vdif_small_reads_window_cache():
    """Loops of small host reads are served from a decoded window of whole
    frames (base/stream.py:_read_small_cached), like the reference's frame
    cache (base/base.py:990-996): same values at every position, fill_value
    changes and large reads bypass / invalidate it."""
    raw = synthetic.vdif_stream(12, 4, 1000, seed=31, invalid=[9, 22])
    want = ostream.vdif_read(raw, fill_value=-5.)[:, :, 0]
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=1e6,
                      fill_value=-5.) as fh:
        spf = fh.samples_per_frame
        got = []
        while fh.tell() < fh.shape[0]:
            got.append(fh.read(min(777, fh.shape[0] - fh.tell())))
        _same(np.concatenate(got), want)
        assert fh._small_cache is not None
        # random small reads, some straddling frames, into out= as well
        rng = np.random.default_rng(3)
        for _ in range(40):
            start = int(rng.integers(0, want.shape[0] - 50))
            n = int(rng.integers(1, 50))
            fh.seek(start)
            if rng.random() < 0.5:
                _same(fh.read(n), want[start:start + n])
            else:
                out = np.empty((n, 4), np.float32)
                assert fh.read(out=out) is out
                _same(out, want[start:start + n])
            assert fh.tell() == start + n
        # a returned array is a copy: writing to it must not poison the cache
        fh.seek(10)
        a = fh.read(5)
        a[:] = 99.
        fh.seek(10)
        _same(fh.read(5), want[10:15])
        # the window spans whole frames around the read
        w0, w1 = fh._small_cache[:2]
        assert w0 % spf == 0 and w0 <= 10 < w1
        fh.seek(0)
        _same(fh.read(), want)                   # large read: direct path
        fh.seek(spf * 3 - 2)
        _same(fh.read(4), want[spf * 3 - 2:spf * 3 + 2])
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=1e6,
                      subset=[2, 0]) as fh:
        fh.seek(4000 * 2 + 17)                   # inside invalid frame 9
        _same(fh.read(30), ostream.vdif_read(raw)[8017:8047, :, 0][:, [2, 0]])


def vdif_many_small_writes():
    """Frames assembled from many small write() calls (host arrays, mixed
    valid flags) equal the frames of one big write."""
    rng = np.random.default_rng(17)
    data = (rng.standard_normal((5 * 4000, 4)) * 2.5).astype(np.float32)
    h0 = bb.vdif.VDIFHeader.fromvalues(
        edv=0, time='2020-01-01T00:00:00', nchan=1, bps=2,
        complex_data=False, thread_id=0, samples_per_frame=4000,
        station='bb', frame_nr=0)

    def written(pieces, valid=None):
        buf = io.BytesIO()
        fw = bb.vdif.open(buf, 'ws', header0=h0, nthread=4, sample_rate=1e6)
        pos = 0
        for i, n in enumerate(pieces):
            ok = True if valid is None else valid(pos, n)
            fw.write(data[pos:pos + n], valid=ok)
            pos += n
        assert pos == data.shape[0]
        return buf.getvalue()

    whole = written([data.shape[0]])
    sizes = [37] * (data.shape[0] // 37) + [data.shape[0] % 37]
    assert written(sizes) == whole
    assert written([3999, 1, 2, 7998, 8000]) == whole
    # a piece flagged invalid marks exactly the frames it touches
    bad = written(sizes, valid=lambda pos, n: not (pos <= 8100 < pos + n))
    frames = np.frombuffer(bad, np.uint8).reshape(5 * 4, 1032)
    invalid = (frames[:, 3] >> 7).reshape(5, 4)
    assert invalid[2].all() and not invalid[[0, 1, 3, 4]].any()


def vdif_stream_info_property():
    """fh.info on stream readers (base/file_info.py StreamReaderInfo)."""
    with bb.vdif.open(sample_path('sample.vdif'), 'rs') as fh:
        info = fh.info
        assert info and info.format == 'vdif' and info.readable
        assert info.shape == (40000, 8) and info.bps == 2
        assert info.sample_rate == 32e6 and not info.complex_data
        assert info.start_time == fh.start_time
        assert info.stop_time == fh.stop_time
    with bb.dada.open(sample_path('sample.dada'), 'rs') as fh:
        info = fh.info
        assert info.format == 'dada' and info.complex_data
        assert info.sample_shape == (2,)


def vdif_header_same_stream_and_mark5b():
    """VDIFHeader.same_stream (vdif/header.py:153-155) and
    VDIFHeader.from_mark5b_header (:246-288)."""
    with bb.vdif.open(sample_path('sample.vdif'), 'rb') as fh:
        h1 = fh.read_frame().header
        h2 = fh.read_frame().header
    assert h1.same_stream(h2) and h1 != h2
    other = h2.copy()
    other.mutable = True
    other['station_id'] = 1
    assert not h1.same_stream(other)
    with bb.mark5b.open(sample_path('sample.m5b'), 'rb', kday=56000,
                        nchan=8) as fb:
        fb.read_frame()
        m5f = fb.read_frame()
    vf = bb.vdif.VDIFFrame.from_mark5b_frame(m5f)
    vh = bb.vdif.VDIFHeader.from_mark5b_header(
        m5f.header, bps=m5f.payload.bps, nchan=m5f.payload.sample_shape[0])
    assert vh == vf.header and vh.edv == 0xab
    assert vh['frame_nr'] == m5f.header['frame_nr'] == 1
    assert vh['bcd_fraction'] == m5f.header['bcd_fraction']
    assert not h1.same_stream(vh)


def vdif_file_name_sequencer():
    """helpers.sequentialfile.FileNameSequencer: templates filled from a
    header, used for writing and reading a stream split over files
    (helpers/sequentialfile.py:18-83, tests/test_sequentialfile.py)."""
    import os
    import tempfile
    from baseband_b200.helpers import sequentialfile as sf
    assert sf.FileNameSequencer('a{file_nr:03d}.vdif')[10] == 'a010.vdif'
    raw = synthetic.vdif_stream(12, 4, 1000, seed=21)
    want = ostream.vdif_read(raw)[:, :, 0]
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=1e6) as fh:
        header0 = fh.header0
    tmp = tempfile.mkdtemp()
    fns = sf.FileNameSequencer(
        os.path.join(tmp, 'obs.edv{edv:d}.{file_nr:05d}.vdif'), header0)
    assert fns[3].endswith('obs.edv0.00003.vdif') and len(fns) == 0
    with bb.vdif.open(fns, 'ws', header0=header0, nthread=4,
                      sample_rate=1e6, file_size=4 * 4 * 1032) as fw:
        fw.write(want)
    assert len(fns) == 3 and fns[-1] == fns[2]
    assert os.path.getsize(fns[0]) == 4 * 4 * 1032
    with bb.vdif.open(fns, 'rs', sample_rate=1e6) as fh:
        _same(fh.read(), want)
    with bb.vdif.open([fns[i] for i in range(3)], 'rs',
                      sample_rate=1e6) as fh:
        fh.seek(5000)
        _same(fh.read(3000), want[5000:8000])
    try:
        sf.FileNameSequencer('x{nope}.vdif', header0)
    except KeyError:
        pass
    else:
        raise AssertionError('unknown template key should raise KeyError')


def dada_guppi_memmap_frame():
    """fw.memmap_frame(): header written at once, payload mapped and filled
    in pieces (dada/tests/test_dada.py:275-316, guppi likewise)."""
    import os
    import tempfile
    tmp = tempfile.mkdtemp()
    # DADA
    with bb.dada.open(sample_path('sample.dada'), 'rb') as fb:
        frame = fb.read_frame(memmap=False)
    name = os.path.join(tmp, 'a2.dada')
    with bb.dada.open(name, 'wb') as fw:
        frame4 = fw.memmap_frame(frame.header)
    assert frame4 != frame                  # nothing filled in yet
    with bb.dada.open(name, 'rb') as fr:
        assert fr.read_frame() != frame
    frame4[:20] = frame[:20]
    _same(frame4[:20], frame[:20])
    assert frame4 != frame
    frame4[20:] = frame[20:]
    assert frame4 == frame
    del frame4
    with bb.dada.open(name, 'rb') as fr:
        assert fr.read_frame() == frame
    # (the header text is re-formatted on writing; the payload bytes match)
    assert open(name, 'rb').read()[4096:] == open(
        sample_path('sample.dada'), 'rb').read()[4096:]
    name = os.path.join(tmp, 'a4.dada')
    with bb.dada.open(name, 'wb') as fw:
        frame8 = fw.memmap_frame(**frame.header)
        frame8[:] = frame.data
    assert frame8 == frame
    del frame8
    with bb.dada.open(name, 'rb') as fr:
        assert fr.read_frame() == frame
    # GUPPI
    with bb.guppi.open(sample_path('sample_puppi.raw'), 'rb') as fb:
        gframe = fb.read_frame(memmap=False)
    name = os.path.join(tmp, 'a2.raw')
    with bb.guppi.open(name, 'wb') as fw:
        g4 = fw.memmap_frame(gframe.header)
    assert g4 != gframe
    g4[:300] = gframe[:300]
    g4[300:] = gframe[300:]
    assert g4 == gframe
    del g4
    with bb.guppi.open(name, 'rb') as fr:
        assert fr.read_frame() == gframe


# ------------------------------------------------------------------ GSB
GSB = os.path.join(os.path.dirname(sample_path('x')), 'gsb')


def gsb_rawdump_read():
    want = ostream.gsb_rawdump_read(
        np.fromfile(os.path.join(GSB, 'sample_gsb_rawdump.dat'), np.uint8),
        payload_nbytes=4096, nframe=10)
    _same(want[:16384], OUT['gsb_rawdump_8192_data'])
    with bb.gsb.open(os.path.join(GSB, 'sample_gsb_rawdump.timestamp'), 'rs',
                     raw=os.path.join(GSB, 'sample_gsb_rawdump.dat'),
                     sample_rate=1e8 / 3 / 2 ** 10, payload_nbytes=4096,
                     squeeze=False, chunk_nbytes=3 * 4096) as fh:
        assert fh.header0.mode == 'rawdump'
        assert fh.samples_per_frame == 8192 and fh.bps == 4
        assert fh.shape == (10 * 8192, 1), fh.shape
        assert fh.start_time.isot == '2015-04-27T13:15:00.000000240'
        _same(fh.read(), want)
        fh.seek(8190)
        _same(fh.read(9000), want[8190:17190])
    with bb.gsb.open(os.path.join(GSB, 'sample_gsb_rawdump.timestamp'), 'rs',
                     raw=os.path.join(GSB, 'sample_gsb_rawdump.dat')) as fh:
        assert fh.samples_per_frame == 2 ** 23
        assert fh.payload_nbytes == 2 ** 22
        assert abs(fh.sample_rate - 1e8 / 3) < 1e-6
    with bb.gsb.open(os.path.join(GSB, 'sample_gsb_rawdump.timestamp'),
                     'rt') as ft:
        h = ft.read_timestamp()
        assert h['gps'] == '2015 04 27 18 45 00 0.000000240'
        h2 = bb.gsb.GSBHeader.fromvalues(mode='rawdump', time=h.time)
        assert h2 == h


def gsb_phased_read_write():
    frames = OUT['gsb_phased_8192_frames']            # (5, 16, 2, 512)
    want = frames.reshape(-1, 2, 512)
    raw = [[os.path.join(GSB, 'sample_gsb_phased.Pol-%s%d.dat' % (p, k))
            for k in (1, 2)] for p in 'LR']
    ts = os.path.join(GSB, 'sample_gsb_phased.timestamp')
    with bb.gsb.open(ts, 'rs', raw=raw, sample_rate=1e8 / 3 / 2 ** 19,
                     payload_nbytes=8192, chunk_nbytes=2 * 4 * 8192) as fh:
        assert fh.header0.mode == 'phased'
        assert fh.sample_shape == (2, 512)
        assert fh.sample_shape.nthread == 2
        assert fh.samples_per_frame == 16 and fh.complex_data
        assert fh.shape == (80, 2, 512)
        assert fh.header0['seq_nr'] == 9995
        assert fh.start_time.isot == '2013-07-27T21:23:55.324108800'
        data = fh.read()
        _same(data, want)
        fh.seek(15)
        _same(fh.read(33), want[15:48])
        header0 = fh.header0
    with bb.gsb.open(ts, 'rs', raw=raw[1], payload_nbytes=8192,
                     sample_rate=1e8 / 3 / 2 ** 19, squeeze=False) as fh:
        assert fh.shape == (80, 1, 512)
        _same(fh.read()[:, 0], want[:, 1])
    # write it back: raw files and timestamps identical to the sample's
    bufs = [[io.BytesIO(), io.BytesIO()], [io.BytesIO(), io.BytesIO()]]
    bts = io.StringIO()
    fw = bb.gsb.open(bts, 'ws', raw=bufs, header0=header0,
                     sample_rate=1e8 / 3 / 2 ** 19, payload_nbytes=8192)
    fw.write(data[:40])
    fw.write(data[40:])
    for group, names in zip(bufs, raw):
        for b, name in zip(group, names):
            _same(np.frombuffer(b.getvalue(), np.uint8),
                  np.fromfile(name, np.uint8)[:5 * 8192])
    lines = bts.getvalue().split('\n')
    with open(ts) as f:
        orig = [ln.strip() for ln in f.read().split('\n')]
    # GPS time, sequence number and memory block follow from the index
    for got, ref in zip(lines[:5], orig[:5]):
        assert got.split()[7:] == ref.split()[7:], (got, ref)


def vdif_host_buffer(dev):
    """Pinned-host source and sink (zero-copy staging)."""
    from baseband_b200.base.memory import HostBuffer
    raw = synthetic.vdif_stream(6, 16, 8000, seed=12, edv=0)
    want = ostream.vdif_read(raw)[:, :, 0]
    src = HostBuffer(raw)
    with bb.vdif.open(src, 'rs', sample_rate=64e6, device=dev,
                      chunk_nbytes=2 * 16 * 8032) as fh:
        data = fh.read()
        _same(data.cpu().numpy(), want)
        header0 = fh.header0
    with bb.vdif.open(HostBuffer(raw), 'rs', sample_rate=64e6) as fh:
        fh.seek(5)
        _same(fh.read(100000), want[5:100005])
    sink = HostBuffer(raw.size)
    fw = bb.vdif.open(sink, 'ws', header0=header0, nthread=16,
                      sample_rate=64e6, device=dev)
    fw.write(data)                       # device tensor in, no H2D of floats
    fw._flush(final=False)
    got = sink.getvalue().reshape(-1, 8032)
    ref = raw.reshape(-1, 8032)
    order = np.argsort(
        (ref[:, 12:16].view('<u4')[:, 0] >> 16 & 0x3ff).reshape(6, 16), 1)
    ref_sorted = ref.reshape(6, 16, 8032)[np.arange(6)[:, None], order]
    _same(got.reshape(6, 16, 8032)[:, :, 32:], ref_sorted[:, :, 32:])
    _same(got.reshape(6, 16, 8032)[:, :, :8], ref_sorted[:, :, :8])


def vdif_missing_frames():
    """Frame-level losses (dropped / duplicated frames): gaps are filled with
    fill_value, as the reference's verify='fix' recovery does one frame at a
    time (vdif/tests/test_corrupt_files.py)."""
    import warnings
    nset, nthread = 9, 8
    raw = synthetic.vdif_stream(nset, nthread, 5000, seed=41,
                                thread_order=np.arange(nthread))
    full = ostream.vdif_read(raw)[:, :, 0]
    frames = raw.reshape(nset * nthread, 5032)
    tid = (frames[:, 12:16].view('<u4')[:, 0] >> 16) & 0x3ff
    drop = [3, 20, 21, 22, 47]                 # physical frames lost
    keep = np.array([i for i in range(nset * nthread) if i not in drop])
    lossy = frames[keep]
    lossy = np.concatenate([lossy[:30], lossy[29:30], lossy[30:]])  # a dup
    want = full.copy()
    for i in drop:
        s = i // nthread
        want[s * 20000:(s + 1) * 20000, tid[i]] = -7.
    for chunk in (None, 2 * nthread * 5032):
        with warnings.catch_warnings(record=True):
            warnings.simplefilter('always')
            with bb.vdif.open(io.BytesIO(lossy.tobytes()), 'rs',
                              sample_rate=32e6, fill_value=-7.,
                              chunk_nbytes=chunk) as fh:
                assert fh.shape == want.shape
                _same(fh.read(), want)
                fh.seek(59990)
                _same(fh.read(20020), want[59990:80010])
    # a loss that keeps the file a whole number of sets long is found from
    # the time of the last frame
    drop8 = list(range(16, 24))                # one complete frame set
    lossy8 = frames[[i for i in range(nset * nthread) if i not in drop8]]
    want8 = full.copy()
    want8[2 * 20000:3 * 20000] = -7.
    with bb.vdif.open(io.BytesIO(lossy8.tobytes()), 'rs', sample_rate=32e6,
                      fill_value=-7.) as fh:
        assert fh._index is not None
        _same(fh.read(), want8)
    # two sets swapped: same length, same last frame -> caught by the GPU
    # consistency check during the read, which then re-reads via the index
    order = np.arange(nset * nthread).reshape(nset, nthread)
    order[[3, 4]] = order[[4, 3]]
    swapped = frames[order.reshape(-1)]
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter('always')
        with bb.vdif.open(io.BytesIO(swapped.tobytes()), 'rs',
                          sample_rate=32e6) as fh:
            assert fh._index is None
            _same(fh.read(), full)
            assert fh._index is not None
    assert any('missing or out-of-order' in str(r.message) for r in rec)
    # the same through a SMALL read (served from the decoded-window cache):
    # the window decoded before the index existed must not be reused
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.vdif.open(io.BytesIO(swapped.tobytes()), 'rs',
                          sample_rate=32e6) as fh:
            fh.seek(60000)
            _same(fh.read(1000), full[60000:61000])
            fh.seek(80000)
            _same(fh.read(1000), full[80000:81000])
            fh.seek(60500)
            _same(fh.read(100), full[60500:60600])


def vdif_duplicate_thread_selection():
    """A subset that names a thread twice returns it twice (as numpy indexing
    of the decoded frame set does in the reference, vdif/base.py:464-490)."""
    raw = synthetic.vdif_stream(3, 8, 5000, seed=5)
    full = ostream.vdif_read(raw)[:, :, 0]
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=32e6,
                      subset=[1, 5, 1, -1]) as fh:
        assert fh.sample_shape == (4,)
        _same(fh.read(), full[:, [1, 5, 1, 7]])
    one = synthetic.vdif_stream(4, 1, 5000, seed=6)
    want = ostream.vdif_read(one)[:, :, 0]
    with bb.vdif.open(io.BytesIO(one.tobytes()), 'rs', sample_rate=32e6,
                      squeeze=False, subset=[0, -1]) as fh:
        got = fh.read()
        assert got.shape == (want.shape[0], 2, 1)
        _same(got[:, 0, 0], want[:, 0])
        _same(got[:, 1, 0], want[:, 0])


def vdif_subset_multichunk_host_read():
    """Host read with a non-contiguous subset over many chunks (the gathered
    piece is a fresh tensor handed from the compute to the D2H stream)."""
    nset, nthread = 40, 8
    raw = synthetic.vdif_stream(nset, nthread, 5000, seed=77)
    full = ostream.vdif_read(raw)[:, :, 0]
    for subset in ([6, 1, 3], slice(1, 8, 3)):
        with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=32e6,
                          subset=subset, chunk_nbytes=2 * nthread * 5032) as fh:
            _same(fh.read(), full[:, subset])
            out = np.empty((300000,) + fh.sample_shape, np.float32)
            fh.seek(1234)
            fh.read(out=out)
            _same(out, full[1234:301234][:, subset])


def dada_file_sequence():
    """A list of files read as one stream; a template written as several
    files (helpers/sequentialfile)."""
    import tempfile
    rng = np.random.default_rng(6)
    data = (rng.integers(-128, 128, (4000, 2, 2))
            + 1j * rng.integers(-128, 128, (4000, 2, 2))).astype(np.complex64)
    with tempfile.TemporaryDirectory() as tmp:
        template = os.path.join(tmp, 'seq_{file_nr:02d}.dada')
        fw = bb.dada.open(template, 'ws', time='2013-07-02T01:39:20',
                          sample_rate=16e6, samples_per_frame=1000,
                          sample_shape=(2, 2), complex_data=True, bps=8,
                          file_size=4096 + 8000)
        fw.write(data)
        fw.close()
        names = sorted(os.path.join(tmp, n) for n in os.listdir(tmp))
        assert len(names) == 4
        assert all(os.path.getsize(n) == 4096 + 8000 for n in names)
        with bb.dada.open(names, 'rs', chunk_nbytes=3000) as fr:
            assert fr.shape == (4000, 2, 2)
            _same(fr.read(), data)
            fr.seek(1990)
            _same(fr.read(1020), data[1990:3010])
        with bb.dada.open(names[2], 'rs') as fr:
            assert fr.header0['OBS_OFFSET'] == 2 * 8000
            _same(fr.read(), data[2000:3000])


def vdif_legacy_headers():
    """Legacy VDIF: 16-byte headers (vdif/header.py:529-542 legacy_mode)."""
    nset, nthread, payload = 5, 2, 4000
    raw32 = synthetic.vdif_stream(nset, nthread, payload, seed=17, bps=4,
                                  nchan=2, thread_order=[1, 0])
    f32 = raw32.reshape(-1, payload + 32)
    legacy = np.concatenate([f32[:, :16], f32[:, 32:]], axis=1).copy()
    w = legacy[:, :16].view('<u4')
    w[:, 0] |= np.uint32(1 << 30)                       # legacy_mode
    w[:, 2] = (w[:, 2] & np.uint32(0xff000000)) | np.uint32(
        (payload + 16) // 8)
    want = ostream.vdif_read(legacy.reshape(-1))
    assert want.shape == (nset * 4000, 2, 2)
    with bb.vdif.open(io.BytesIO(legacy.tobytes()), 'rs', sample_rate=2e6,
                      squeeze=False) as fh:
        assert fh.header0.nbytes == 16 and fh.header0.edv is False
        assert fh.shape == want.shape
        _same(fh.read(), want)
        header0 = fh.header0
    buf = io.BytesIO()
    fw = bb.vdif.open(buf, 'ws', header0=header0, nthread=2, sample_rate=2e6)
    fw.write(want)
    got = np.frombuffer(buf.getvalue(), np.uint8).reshape(-1, payload + 16)
    # writer emits thread 0 first; the source had thread 1 first
    src = legacy.reshape(nset, 2, payload + 16)[:, ::-1].reshape(
        -1, payload + 16)
    _same(got[:, 16:], src[:, 16:])
    _same(got[:, :12], src[:, :12])


def mark5b_missing_frames():
    raw, valid = synthetic.mark5b_stream(30, invalid_fraction=0.1, seed=9)
    full = ostream.mark5b_read(raw, 8, fill_value=-1.)
    frames = raw.reshape(30, 10016)
    drop = [4, 5, 17]
    lossy = frames[[i for i in range(30) if i not in drop]]
    want = full.copy()
    for i in drop:
        want[i * 5000:(i + 1) * 5000] = -1.
    for chunk in (None, 4 * 10016):
        with bb.mark5b.open(io.BytesIO(lossy.tobytes()), 'rs', nchan=8,
                            sample_rate=32e6, kday=56000, fill_value=-1.,
                            chunk_nbytes=chunk) as fh:
            assert fh._index is not None
            assert fh.shape == want.shape
            _same(fh.read(), want)
            fh.seek(19990)
            _same(fh.read(10020), want[19990:30010])


def vdif_pickle_reader():
    """Readers can be sent to other processes (base/base.py:123-151): the
    file is re-opened by name, the offset kept, device state re-created."""
    import pickle
    want = OUT['sample_vdif_data'][:, :, 0]
    with bb.vdif.open(sample_path('sample.vdif'), 'rs') as fh:
        fh.seek(6)
        _same(fh.read(3), want[6:9])
        clone = pickle.loads(pickle.dumps(fh))
        assert clone.tell() == 9
        _same(clone.read(20001), want[9:20010])
        assert fh.tell() == 9
        _same(fh.read(2), want[9:11])
        clone.close()
    with bb.mark5b.open(sample_path('sample.m5b'), 'rs', sample_rate=32e6,
                        kday=56000, nchan=8) as fh:
        clone = pickle.loads(pickle.dumps(fh))
        _same(clone.read(), OUT['sample_m5b_data'])
        clone.close()


def vdif_on_device_hook(dev):
    """``read(on_device=...)``: every decoded chunk is handed to a GPU
    consumer (here the stream writer) before it is copied to the host."""
    from baseband_b200.base.memory import HostBuffer
    raw = synthetic.vdif_stream(12, 8, 5000, seed=23,
                                thread_order=np.arange(8))
    want = ostream.vdif_read(raw)[:, :, 0]
    sink = HostBuffer(raw.size)
    seen = []
    with bb.vdif.open(HostBuffer(raw), 'rs', sample_rate=32e6,
                      chunk_nbytes=3 * 8 * 5032) as fh:
        fw = bb.vdif.open(sink, 'ws', header0=fh.header0, nthread=8,
                          sample_rate=32e6, device=dev)

        def consume(piece):
            seen.append(tuple(piece.shape))
            fw.write(piece)

        out = np.empty(want.shape, np.float32)
        fh.read(out=out, on_device=consume)
        _same(out, want)
        fw._flush(final=False)
    assert len(seen) == 4 and sum(s[0] for s in seen) == want.shape[0]
    _same(sink.getvalue(), raw)
    with bb.vdif.open(HostBuffer(raw), 'rs', sample_rate=32e6, device=dev,
                      chunk_nbytes=5 * 8 * 5032) as fh:
        total = []
        data = fh.read(on_device=lambda p: total.append(p.shape[0]))
        assert sum(total) == want.shape[0] and len(total) == 3
        _same(data.cpu().numpy(), want)


def payload_todevice(dev):
    """``Payload.todevice()`` / ``FrameSet.todevice()``: decoded samples as a
    tensor on the GPU for every format's payload class."""
    with bb.vdif.open(sample_path('sample.vdif'), 'rb') as fh:
        fs = fh.read_frameset()
    t = fs.todevice(dev)
    assert isinstance(t, torch.Tensor) and tuple(t.shape) == (20000, 8, 1)
    _same(t.cpu().numpy(), OUT['sample_vdif_data'][:20000])
    _same(fs.frames[2].payload.todevice(dev).cpu().numpy(),
          OUT['sample_vdif_data'][:20000, 2])
    with bb.mark5b.open(sample_path('sample.m5b'), 'rb', kday=56000,
                        nchan=8) as fh:
        frame = fh.read_frame()
    _same(frame.payload.todevice(dev).cpu().numpy(),
          OUT['sample_m5b_data'][:5000])
    with bb.mark4.open(sample_path('sample.m4'), 'rb', ntrack=64,
                       decade=2010) as fh:
        fh.locate_frame()
        frame = fh.read_frame()
    _same(frame.payload.todevice(dev).cpu().numpy(),
          OUT['sample_m4_data'][640:80000])
    with bb.guppi.open(sample_path('sample_puppi.raw'), 'rb') as fh:
        frame = fh.read_frame(memmap=False)
    _same(frame.payload.todevice(dev).cpu().numpy(),
          OUT['sample_puppi_frames'][0])
    with bb.dada.open(sample_path('sample.dada'), 'rb') as fh:
        frame = fh.read_frame(memmap=False)
    _same(frame.payload.todevice(dev).cpu().numpy(), OUT['sample_dada_data'])


def vdif_mark5b_payload_edv_ab():
    """Mark 5B frames wrapped in VDIF EDV 0xab keep the Mark 5B codec
    (vdif/payload.py:151-154, tests/test_conversion.py in the reference)."""
    want = OUT['sample_m5b_data']
    buf = io.BytesIO()
    with bb.mark5b.open(sample_path('sample.m5b'), 'rb', kday=56000,
                        nchan=8) as fh:
        for _ in range(4):
            m5f = fh.read_frame()
            vf = bb.vdif.VDIFFrame.from_mark5b_frame(m5f, verify=False)
            assert vf.header.edv == 0xab and vf.header.nchan == 8
            vf.tofile(buf)
    raw = buf.getvalue()
    assert len(raw) == 4 * 10032
    with bb.vdif.open(io.BytesIO(raw), 'rs', sample_rate=32e6) as fv:
        assert fv.shape == (20000, 8)
        assert fv.header0['mark5b_frame_nr'] == 0
        _same(fv.read(), want)
    frame = bb.vdif.VDIFFrame.fromfile(io.BytesIO(raw))
    _same(frame.data, want[:5000])
    pl = bb.vdif.VDIFPayload.fromdata(want[:5000], frame.header)
    assert np.array_equal(pl.words, frame.payload.words)


# ------------------------------------------------------- consumers (tasks)
def _counts_from_decoded(data, lv, rows_per_bin):
    """Oracle for the state counts: how often each level occurs in the
    reference's decoded samples, per bin (data: (nsample, ...) float32 with
    NaN where frames are invalid)."""
    nbin = -(-data.shape[0] // rows_per_bin)
    out = np.zeros((nbin,) + data.shape[1:] + (len(lv),), np.int64)
    for b in range(nbin):
        block = data[b * rows_per_bin:(b + 1) * rows_per_bin]
        for c, level in enumerate(lv):
            out[b, ..., c] = (block == level).sum(0)
    return out


def vdif_task_state_counts():
    """tasks.state_counts / integrated_power == counting the levels in what
    the reference decodes (oracle), for the C2 geometry (register path, one
    real channel per thread), with invalid frames, bins that do not line up
    with the chunks, and a thread subset."""
    from baseband_b200 import levels, tasks
    raw = synthetic.vdif_stream(11, 16, 8000, seed=8, invalid=[5, 40, 41])
    lv = levels.offset_binary(2)
    decoded = ostream.vdif_read(raw, fill_value=np.nan)[:, :, 0]
    for chunk, sets_per_bin in ((16 * 8032 * 3, 2), (1 << 30, 4), (16 * 8032,
                                                                    11)):
        with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=64e6,
                          chunk_nbytes=chunk) as fh:
            got = tasks.state_counts(fh, sets_per_bin * 32000)
            assert fh.tell() == 11 * 32000
            want = _counts_from_decoded(decoded, lv, sets_per_bin * 32000)
            assert got.dtype == np.int64 and got.shape == want.shape
            assert np.array_equal(got, want)
            fh.seek(2 * 32000)
            part = tasks.state_counts(fh, 3 * 32000, count=6 * 32000)
            assert np.array_equal(part, _counts_from_decoded(
                decoded[2 * 32000:8 * 32000], lv, 3 * 32000))
            fh.seek(0)
            power = tasks.integrated_power(fh, sets_per_bin * 32000)
            x = decoded.astype(np.float64) ** 2
            nb = want.shape[0]
            ref = np.stack([np.nanmean(x[b * sets_per_bin * 32000:
                                         (b + 1) * sets_per_bin * 32000], 0)
                            for b in range(nb)])
            assert np.allclose(power, ref, rtol=1e-10, atol=0)   # float64 summation order
            try:
                fh.seek(5)
                tasks.state_counts(fh, 32000)
            except ValueError:
                pass
            else:
                raise AssertionError('partial frames must be refused')
    with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=64e6,
                      subset=[3, 12]) as fh:
        got = tasks.state_counts(fh)
        assert np.array_equal(got, _counts_from_decoded(
            decoded[:, [3, 12]], lv, 11 * 32000))


def vdif_task_state_counts_shapes():
    """Other payload shapes: complex, several channels per thread (register
    path with 2 and 4 elements per word; histogram path for 8 and 16
    channels, 4-bit), 1 bit."""
    from baseband_b200 import levels, tasks
    for bps, nthread, nchan, cplx, payload in (
            (2, 8, 1, True, 8000), (2, 2, 4, False, 4000),
            (2, 2, 8, False, 4000), (2, 1, 16, True, 8000),
            (1, 4, 1, False, 2000), (1, 2, 8, False, 2000),
            (1, 1, 64, False, 4000), (4, 2, 1, False, 2000),
            (4, 1, 4, True, 4000), (2, 1, 64, False, 8000)):
        nset = 5
        raw = synthetic.vdif_stream(nset, nthread, payload, bps=bps,
                                    nchan=nchan, complex_data=cplx,
                                    seed=100 + bps + nchan, invalid=[3])
        lv = levels.offset_binary(bps)
        with bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=1e6,
                          squeeze=False, fill_value=np.nan,
                          chunk_nbytes=2 * nthread * (payload + 32)) as fh:
            decoded = fh.read()
            spf = fh.samples_per_frame
            fh.seek(0)
            got = tasks.state_counts(fh, 2 * spf)
            fh.seek(0)
            power = tasks.integrated_power(fh, 2 * spf, average=False)
        if cplx:
            # an invalid complex frame reads as fill + 0j: blank both parts
            decoded = np.stack([decoded.real, np.where(
                np.isnan(decoded.real), np.nan, decoded.imag)], -1)
        want = _counts_from_decoded(decoded, lv, 2 * spf)
        assert got.shape == want.shape, (got.shape, want.shape)
        assert np.array_equal(got, want), (bps, nthread, nchan, cplx)
        x = np.nan_to_num(decoded.astype(np.float64)) ** 2
        if cplx:
            x = x.sum(-1)
        ref = np.stack([x[b * 2 * spf:(b + 1) * 2 * spf].sum(0)
                        for b in range(3)])
        assert np.allclose(power, ref, rtol=1e-10, atol=0)   # float64 summation order


def mark5b_task_state_counts():
    from baseband_b200 import levels, tasks
    raw, valid = synthetic.mark5b_stream(12, invalid_fraction=0.25, seed=9)
    lv = levels.mark5b(2)
    decoded = ostream.mark5b_read(raw, 16, fill_value=np.nan)
    with bb.mark5b.open(io.BytesIO(raw.tobytes()), 'rs', nchan=16,
                        sample_rate=16e6, kday=56000,
                        chunk_nbytes=5 * 10016) as fh:
        got = tasks.state_counts(fh, 4 * 2500)
    want = _counts_from_decoded(decoded, lv, 4 * 2500)
    assert got.shape == want.shape == (3, 16, 4)
    assert np.array_equal(got, want)
    assert got.sum() == valid.sum() * 2500 * 16


def mark4_task_state_counts():
    """tasks.state_counts on Mark 4 track words == counting the levels in what
    the reference decoded from its own sample files (golden vectors), for all
    five track layouts; header steps (fill) are not counted.  Plus the
    synthetic C3 stream with an invalid frame, in two bins."""
    from baseband_b200 import levels, tasks
    lv = levels.sign_magnitude()
    for name, ntrack in M4_SAMPLES:
        want_data = OUT[name.replace('.', '_') + '_data']
        with bb.mark4.open(sample_path(name), 'rs', ntrack=ntrack,
                           decade=2010, fill_value=-7.) as fh:
            spf = fh.samples_per_frame
            nframe = fh.shape[0] // spf
            got = tasks.state_counts(fh, spf, count=nframe * spf)
            assert fh.tell() == nframe * spf
            assert np.array_equal(tasks.state_levels(fh), lv)
        want = _counts_from_decoded(want_data[:nframe * spf], lv, spf)
        assert got.shape == want.shape == (nframe, want_data.shape[1], 4)
        assert np.array_equal(got, want), name
        fanout = spf // 20000
        assert np.all(got.sum(-1) == (20000 - 160) * fanout)
    h0 = bb.mark4.Mark4Header.fromvalues(
        64, time='2014-06-16T07:38:12.475', bps=2, fanout=4, nsb=1,
        system_id=108)
    rng = np.random.default_rng(77)
    data = rng.choice(lv, size=(6 * 80000, 8), p=[.1, .4, .3, .2])
    buf = io.BytesIO()
    fw = bb.mark4.open(buf, 'ws', header0=h0, sample_rate=32e6)
    fw.write(data[:2 * 80000])
    fw.write(data[2 * 80000:3 * 80000], valid=False)
    fw.write(data[3 * 80000:])
    decoded = ostream.mark4_read(np.frombuffer(buf.getvalue(), np.uint8), 64,
                                 fill_value=np.nan)
    with bb.mark4.open(io.BytesIO(buf.getvalue()), 'rs', ntrack=64,
                       decade=2010, chunk_nbytes=2 * 160000) as fh:
        fh.seek(80000)
        got = tasks.state_counts(fh, 3 * 80000, count=5 * 80000)
        fh.seek(80000)
        power = tasks.integrated_power(fh, 3 * 80000, count=5 * 80000,
                                       average=False)
    want = _counts_from_decoded(decoded[80000:], lv, 3 * 80000)
    assert got.shape == want.shape == (2, 8, 4)
    assert np.array_equal(got, want)
    assert got[0].sum() == 2 * 8 * (80000 - 640)      # one frame invalid
    assert np.allclose(power, (want * lv.astype(np.float64) ** 2).sum(-1),
                       rtol=1e-12)


def gsb_task_counts_and_moments():
    """Consumers on GSB streams: state counts of the 4-bit raw-voltage dump
    and moments / power of the 8-bit phased-array mode (two polarisations in
    two files each), against the reference's decoded sample files (golden
    vectors), in bins that cut across the chunks."""
    from baseband_b200 import tasks
    with bb.gsb.open(os.path.join(GSB, 'sample_gsb_rawdump.timestamp'), 'rs',
                     raw=os.path.join(GSB, 'sample_gsb_rawdump.dat'),
                     sample_rate=1e8 / 3 / 2 ** 10, payload_nbytes=4096,
                     squeeze=False, chunk_nbytes=3 * 4096) as fh:
        lv = tasks.state_levels(fh)
        assert np.array_equal(lv, np.r_[0:8, -8:0].astype(np.float32))
        fh.seek(8192)
        got = tasks.state_counts(fh, 4 * 8192, count=9 * 8192)
        fh.seek(8192)
        power = tasks.integrated_power(fh, 4 * 8192, count=9 * 8192)
    decoded = ostream.gsb_rawdump_read(
        np.fromfile(os.path.join(GSB, 'sample_gsb_rawdump.dat'), np.uint8),
        payload_nbytes=4096, nframe=10)[8192:]
    _same(decoded[:8192], OUT['gsb_rawdump_8192_data'][8192:])
    want = _counts_from_decoded(decoded, lv, 4 * 8192)
    assert got.shape == want.shape == (3, 1, 16)
    assert np.array_equal(got, want)
    for b in range(3):
        blk = decoded[b * 4 * 8192:(b + 1) * 4 * 8192].astype(np.float64)
        assert np.allclose(power[b], (blk ** 2).mean(0), rtol=1e-12)
    frames = OUT['gsb_phased_8192_frames']            # (5, 16, 2, 512)
    want = frames.reshape(-1, 2, 512)
    raw = [[os.path.join(GSB, 'sample_gsb_phased.Pol-%s%d.dat' % (p, k))
            for k in (1, 2)] for p in 'LR']
    ts = os.path.join(GSB, 'sample_gsb_phased.timestamp')
    with bb.gsb.open(ts, 'rs', raw=raw, sample_rate=1e8 / 3 / 2 ** 19,
                     payload_nbytes=8192, chunk_nbytes=2 * 4 * 8192) as fh:
        fh.seek(16)
        n, total, sq = tasks.moments(fh, 3 * 16, count=4 * 16)
        assert fh.tell() == 5 * 16
        fh.seek(16)
        power = tasks.integrated_power(fh, 3 * 16, count=4 * 16)
    parts = np.stack([want.real, want.imag], -1).astype(np.int64)[16:]
    assert n.shape == total.shape == sq.shape == (2, 2, 512, 2)
    for b in range(2):
        blk = parts[b * 48:(b + 1) * 48]
        assert np.all(n[b] == blk.shape[0])
        assert np.array_equal(total[b], blk.sum(0))
        assert np.array_equal(sq[b], (blk * blk).sum(0))
        assert np.allclose(power[b], (blk.astype(np.float64) ** 2)
                           .sum(-1).mean(0), rtol=1e-12)


def dada_task_moments():
    """tasks.moments / integrated_power on DADA streams, whose frames are cut
    into pieces by the chunk size, against the reference's decoded sample
    files (golden vectors)."""
    from baseband_b200 import tasks
    for name in ('sample.dada', 'sample_meerkat.dada'):
        want = OUT[name.replace('.', '_') + '_data']      # (n, npol, nchan)
        for chunk in (None, 4096, 10000):
            with bb.dada.open(sample_path(name), 'rs', squeeze=False,
                              chunk_nbytes=chunk) as fh:
                spf = fh.samples_per_frame
                n, total, sq = tasks.moments(fh)
                assert fh.tell() == fh.shape[0] // spf * spf
                fh.seek(0)
                power = tasks.integrated_power(fh)
            body = want[:want.shape[0] // spf * spf]
            if np.iscomplexobj(body):
                parts = np.stack([body.real, body.imag], -1).astype(np.int64)
                mean_power = (parts.astype(np.float64) ** 2).sum(-1).mean(0)
            else:
                parts = body.astype(np.int64)
                mean_power = (parts.astype(np.float64) ** 2).mean(0)
            assert n.shape == (1,) + parts.shape[1:], (n.shape, parts.shape)
            assert np.all(n[0] == parts.shape[0])
            assert np.array_equal(total[0], parts.sum(0))
            assert np.array_equal(sq[0], (parts * parts).sum(0))
            assert np.allclose(power[0], mean_power, rtol=1e-12)
    with bb.dada.open(sample_path('sample_meerkat.dada'), 'rs',
                      chunk_nbytes=4098) as fh:     # 2049 two-byte samples
        with pytest.raises(ValueError):       # pieces must be whole words
            tasks.moments(fh)
    with bb.dada.open(sample_path('sample_mkbf.dada'), 'rs') as fh:
        with pytest.raises(NotImplementedError):
            tasks.moments(fh)


# -------------------------------------------- byte-level damage (GPU index)
def vdif_byte_slip():
    """Bytes lost inside a frame and bytes inserted between frames: the GPU
    frame index (sync search on the stream's invariant header bits + check
    one frame on, `locate_frames` of base/base.py:181-335, then placement by
    header time) recovers every frame that is followed by a header where it
    should be; the others read as fill_value."""
    import warnings
    nset, nthread = 9, 4
    raw = synthetic.vdif_stream(nset, nthread, 5000, seed=43,
                                thread_order=np.arange(nthread))
    full = ostream.vdif_read(raw)[:, :, 0]
    rng = np.random.default_rng(2)
    for cut, extra in (((13, 700, 1234), None),       # cut inside frame 13
                       (None, (22, 7)),               # 7 bytes after frame 21
                       ((17, 100, 2), (30, 10001))):  # both; odd alignments
        pieces, lost = raw.copy(), []
        if extra is not None:
            at, n = extra
            pieces = np.concatenate([
                pieces[:at * 5032], rng.integers(0, 256, n, dtype=np.uint8),
                pieces[at * 5032:]])
            lost.append(at - 1)          # its check lands in the garbage
        if cut is not None:
            i, a, n = cut
            pieces = np.concatenate([pieces[:i * 5032 + a],
                                     pieces[i * 5032 + a + n:]])
            lost.append(i)               # truncated
        want = full.copy()
        for i in lost:
            s, t = divmod(i, nthread)
            want[s * 20000:(s + 1) * 20000, t] = -7.
        for chunk in (None, 3 * nthread * 5032):
            with warnings.catch_warnings(record=True):
                warnings.simplefilter('always')
                with bb.vdif.open(io.BytesIO(pieces.tobytes()), 'rs',
                                  sample_rate=32e6, fill_value=-7.,
                                  chunk_nbytes=chunk) as fh:
                    data = fh.read()
                    assert fh._index is not None
                    assert data.shape == want.shape, (data.shape, want.shape)
                    _same(data, want)
                    fh.seek(39990)
                    _same(fh.read(20020), want[39990:60010])


def mark5b_byte_slip():
    """The reference's own corrupted Mark 5B file (mark5b/tests/
    test_mark5b.py:508-524: bytes [10040, 20000) of sample.m5b cut out):
    frame 1 is gone, frames 2 and 3 sit 9960 bytes early."""
    import warnings
    m5 = np.fromfile(sample_path('sample.m5b'), np.uint8)
    bad = np.concatenate([m5[:10040], m5[20000:]])
    want = OUT['sample_m5b_data'].copy()
    want[5000:10000] = -3.
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.mark5b.open(io.BytesIO(bad.tobytes()), 'rs', nchan=8,
                            sample_rate=32e6, kday=56000,
                            fill_value=-3.) as fh:
            assert fh._index is not None
            assert fh._index[:, 0].tolist() == [0, -1, 10072, 20088]
            _same(fh.read(), want)


def mark4_missing_frames_and_byte_slip():
    """Irregular Mark 4 streams: frames dropped, swapped, a junk prefix and
    bytes cut out of a frame -- the GPU index (all-ones sync search with check
    one frame on, placement by the BCD time code of track 0) fills what is
    gone with fill_value.  Also across a year end (unit-year digit wraps)."""
    import warnings
    for start in ('2014-06-16T07:38:12.475', '2019-12-31T23:59:59.990'):
        h0 = bb.mark4.Mark4Header.fromvalues(
            32, time=start, bps=2, fanout=4, nsb=1, system_id=108)
        nframe, spf, fb = 11, 80000, 80000
        rng = np.random.default_rng(17)
        data = rng.choice(np.array([-3.316505, -1., 1., 3.316505],
                                   np.float32), size=(nframe * spf, 4))
        buf = io.BytesIO()
        fw = bb.mark4.open(buf, 'ws', header0=h0, sample_rate=32e6)
        fw.write(data)
        raw = np.frombuffer(buf.getvalue(), np.uint8)
        assert raw.size == nframe * fb
        full = ostream.mark4_read(raw, 32, fill_value=-2.)
        frames = raw.reshape(nframe, fb)

        def check(blob, lost, **kwargs):
            want = full.copy()
            for i in lost:
                want[i * spf:(i + 1) * spf] = -2.
            with warnings.catch_warnings(record=True):
                warnings.simplefilter('always')
                with bb.mark4.open(io.BytesIO(blob.tobytes()), 'rs',
                                   ntrack=32, decade=2010, sample_rate=32e6,
                                   fill_value=-2., **kwargs) as fh:
                    got = fh.read()
                    assert fh._index is not None
                    n = min(got.shape[0], want.shape[0])
                    assert n >= (max(set(range(nframe)) - set(lost)) + 1) * spf
                    _same(got[:n], want[:n])
        # frames 3 and 7 dropped, 5 and 6 swapped
        order = [0, 1, 2, 4, 6, 5, 8, 9, 10]
        check(frames[order].reshape(-1), [3, 7])
        check(frames[order].reshape(-1), [3, 7], chunk_nbytes=3 * fb)
        # 1001 bytes cut out of frame 4 (its successor's sync is not where it
        # should be: frame 4 reads as fill), junk before the first frame
        junk = rng.integers(0, 255, 77, dtype=np.uint8)
        blob = np.concatenate([junk, raw[:4 * fb + 30000],
                               raw[4 * fb + 31001:]])
        check(blob, [4])


def guppi_task_moments():
    """tasks.moments / integrated_power for 8-bit GUPPI streams, channels
    first (with overlap) and time first, against sums over the decoded
    samples (oracle), with bins that do not line up with the chunks."""
    from baseband_b200 import tasks
    for kwargs, fmt in ((dict(nchan=8, npol=2, samples_per_frame=64,
                              overlap=8), {}),
                        (dict(nchan=4, npol=2, samples_per_frame=96,
                              overlap=0), {}),
                        (dict(nchan=16, npol=1, samples_per_frame=32,
                              overlap=0), {})):
        for nframe, per_bin, chunk in ((7, 2, None), (5, 5, 1)):
            raw, _ = synthetic.guppi_stream(nframe, **kwargs)
            decoded = ostream.guppi_read(raw)          # (nsample, npol, nchan)
            with bb.guppi.open(io.BytesIO(raw.tobytes()), 'rs',
                               squeeze=False) as fh:
                spf = fh.samples_per_frame
                if chunk:
                    fh._chunk_nbytes = chunk * fh._frame_nbytes
                n, total, sq = tasks.moments(fh, per_bin * spf)
                assert fh.tell() == nframe * spf
                fh.seek(0)
                power = tasks.integrated_power(fh, per_bin * spf)
            body = decoded[:nframe * spf]
            parts = np.stack([body.real, body.imag], -1).astype(np.int64)
            nbin = -(-nframe // per_bin)
            for b in range(nbin):
                blk = parts[b * per_bin * spf:(b + 1) * per_bin * spf]
                assert np.array_equal(n[b], np.full(blk.shape[1:],
                                                    blk.shape[0]))
                assert np.array_equal(total[b], blk.sum(0))
                assert np.array_equal(sq[b], (blk * blk).sum(0))
                want = (blk.astype(np.float64) ** 2).sum(-1).mean(0)
                assert np.allclose(power[b], want, rtol=1e-12, atol=0)
