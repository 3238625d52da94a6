"""The reference's corrupt-file tests (baseband/vdif/tests/
test_corrupt_files.py) restated as data: the same damaged files, the same
expected arrays.  In the reference these go through the frame-at-a-time repair
of `_bad_frame` (vdif/base.py:536-755); here through the GPU frame index
(sync search + header times, base/stream.py:_build_index_on_device).  Shared
by the CPU-backend run and the GPU run.

Where the two differ it is stated at the case: the index keeps every frame
that is itself intact and followed by a header where one should be, so for one
kind of damage it returns more valid samples than the reference does.
"""
import io
import warnings

import numpy as np

import baseband_b200 as bb
from conftest import sample_path

HIGH = np.float32(3.316505)


# --------------------------------------- TestCorruptSampleCopy (:13-155)
def _sample_copy():
    """sample.vdif three times over, written through the stream writer
    (test_corrupt_files.py:14-33): 6 frame sets of 8 threads, 48 frames."""
    with bb.vdif.open(sample_path('sample.vdif'), 'rs') as fs:
        data = fs.read()
        header0 = fs.header0
        buf = io.BytesIO()
        with bb.vdif.open(buf, 'ws', header0=header0, nthread=8) as fw:
            for _ in range(3):
                fw.write(data)
            start_time, stop_time = fw.start_time, fw.tell('time')
            raw = buf.getvalue()
    return raw, np.concatenate([data, data, data]), start_time, stop_time


def _zero_frames(data, frames):
    """``data`` with the given frames (file order = set * 8 + thread) zeroed
    (test_corrupt_files.py:68-74)."""
    expected = (data.copy().reshape(-1, 20000, 8).transpose(0, 2, 1)
                .reshape(-1, 20000))
    expected[frames] = 0.
    return expected.reshape(-1, 8, 20000).transpose(0, 2, 1).reshape(-1, 8)


def _read(raw, **kwargs):
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.vdif.open(io.BytesIO(raw), 'rs', **kwargs) as fh:
            return fh.read(), fh


MISSING_FRAMES = (36, slice(46, 48), [30, 45], slice(8, 16), 0, slice(4, 12))


def sample_copy_missing_frames(missing):
    """test_missing_frames (:43-76): purely missing frames read as zeros."""
    raw, data, _, _ = _sample_copy()
    sample = np.frombuffer(raw, 'u1').reshape(-1, 5032)
    use = np.ones(len(sample), bool)
    use[missing] = False
    got, _ = _read(sample[use].tobytes())
    want = _zero_frames(data, missing)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)


MISSING_BYTES = (slice(5032 * 26, 5032 * 26 + 1),       # header byte of 26
                 slice(5032 * 26 + 50, 5032 * 26 + 60),  # payload of 26
                 slice(5032 * 27 + 50, 5032 * 29 + 700),  # parts of 27-29
                 slice(5032 * 31 + 10, 5032 * 31 + 20),  # header of 31
                 slice(5032 * 32, 5032 * 32 + 10),       # header of 32
                 slice(5032 * 48 - 1, 5032 * 48))        # last byte of all


def _expected_bad_frames(missing, frame_nbytes=5032):
    """test_corrupt_files.py:78-86: the frames the cut touches, and the one
    before if it starts inside a header."""
    (start_f, start_r), (stop_f, _) = [divmod(s, frame_nbytes)
                                       for s in (missing.start,
                                                 missing.stop - 1)]
    if start_r < 32:
        start_f -= 1
    return start_f, stop_f + 1


def sample_copy_missing_bytes(missing):
    """test_missing_bytes (:99-155): bytes cut out of the file; the frames
    the reference marks invalid are exactly those the index cannot verify."""
    raw, data, start_time, stop_time = _sample_copy()
    corrupted = raw[:missing.start] + raw[missing.stop:]
    bad_start, bad_stop = _expected_bad_frames(missing)
    got, fh = _read(corrupted)
    assert fh.start_time == start_time
    assert abs(fh.stop_time - stop_time) < 1e-9
    want = _zero_frames(data, slice(bad_start, bad_stop))
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)


# ------------------------------------------------ TestCorruptFile (:158-300)
def _fake():
    """16 frame sets of 2 threads x 2 channels, 16 samples per frame, EDV 1
    (test_corrupt_files.py:160-177)."""
    header0 = bb.vdif.VDIFHeader.fromvalues(
        edv=1, time='2010-11-12T13:14:15', nchan=2, bps=2,
        complex_data=False, thread_id=0, samples_per_frame=16, station='me',
        sample_rate=2000.)
    data = np.array([[[-1, 1], [-HIGH, HIGH]]] * 16, np.float32)
    buf = io.BytesIO()
    with bb.vdif.open(buf, 'ws', header0=header0, nthread=2) as fw:
        for _ in range(16):
            fw.write(data)
        raw = buf.getvalue()
    assert len(raw) == 16 * 80
    return raw, data


def _cut(raw, missing):
    return raw[:missing.start] + raw[missing.stop:]


MISSING_FRAMESET = (1, 3, 5, slice(7, 10))


def fake_missing_frameset(frame_nr):
    """test_missing_frameset (:190-205)."""
    if not isinstance(frame_nr, slice):
        frame_nr = slice(frame_nr, frame_nr + 1)
    raw, data = _fake()
    got, _ = _read(_cut(raw, slice(frame_nr.start * 80, frame_nr.stop * 80)))
    got = got.reshape((-1,) + data.shape)
    assert len(got) == 16
    assert np.all(got[:frame_nr.start] == data)
    assert np.all(got[frame_nr.stop:] == data)
    assert np.all(got[frame_nr] == 0.)


MISSING_THREAD = ((3, 0), (3, 1), (1, 1), (15, 1))


def fake_missing_thread(frame_nr, thread):
    """test_missing_thread (:207-225)."""
    raw, data = _fake()
    frame = frame_nr * 2 + thread
    got, _ = _read(_cut(raw, slice(frame * 40, (frame + 1) * 40)))
    got = got.reshape((-1,) + data.shape)
    assert len(got) == 16
    assert np.all(got[:frame_nr] == data)
    assert np.all(got[frame_nr + 1:] == data)
    assert np.all(got[frame_nr, :, thread] == 0.)
    assert np.all(got[frame_nr, :, 1 - thread] == data[:, 1 - thread])


MISSING_END = (slice(0, 80), slice(0, 40), slice(0, 32), slice(16, 32),
               slice(0, 16), slice(0, 1), slice(10, 11), slice(15, 16),
               slice(20, 21), slice(23, 24))


def fake_missing_end(missing_bytes):
    """test_missing_end (:227-252): damage to the first frame of the last
    frame set: the stream is just one frame set shorter (its length is set by
    the last good frame of the first header's thread, vdif/base.py:493-519)."""
    raw, data = _fake()
    missing = slice(missing_bytes.start + 15 * 80, missing_bytes.stop + 15 * 80)
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.vdif.open(io.BytesIO(_cut(raw, missing)), 'rs') as fr:
            assert fr.size == 15 * data.size, (fr.size, fr.shape)
            got = fr.read()
    got = got.reshape((-1,) + data.shape)
    assert len(got) == 15
    assert np.all(got == data)


# (bytes cut, samples expected to read as zero, samples the index ALSO
# recovers where the reference gives up on the whole frame set)
MISSING_MIDDLE = (
    (slice(240, 320), slice(48, 64), None),    # frame set 3 removed
    (slice(279, 281), slice(48, 64), None),    # end of thread 0 + header of 1
    (slice(280, 281), slice(48, 64), None),    # first header byte of thread 1
    # last payload byte of thread 0 of set 3: its successor is no longer where
    # it should be, so thread 0 reads as zero; thread 1 of the set is intact
    # and followed by a good header, so the index keeps it -- the reference
    # zeroes the whole set (slice(48, 64) for both threads)
    (slice(279, 280), slice(48, 64), 1),
    (slice(272, 365), slice(48, 80), None))    # frame sets 3 and 4


def fake_missing_middle(missing_bytes, missing_data, kept_thread):
    """test_missing_middle (:254-274)."""
    raw, data = _fake()
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.vdif.open(io.BytesIO(_cut(raw, missing_bytes)), 'rs') as fr:
            assert fr.size == 16 * data.size
            got = fr.read()
    expected = np.concatenate([data] * 16)
    zeroed = expected.copy()
    zeroed[missing_data] = 0.
    if kept_thread is not None:
        zeroed[missing_data, kept_thread] = expected[missing_data,
                                                     kept_thread]
    assert np.array_equal(got, zeroed)


# --------------------------------------- TestInvalidFrameHeaders (:303-350)
def fake_invalid_frame_headers():
    """Frame set 10 flagged invalid AND with corrupt seconds / frame_nr (CHIME
    ARO files, :303-350): skipped, reads as zeros."""
    header0 = bb.vdif.VDIFHeader.fromvalues(
        edv=1, time='2010-11-12T13:14:15', nchan=2, bps=2,
        complex_data=False, thread_id=0, samples_per_frame=16, station='me',
        sample_rate=2000.)
    data = np.array([[[-1, 1], [-HIGH, HIGH]]] * 16, np.float32)
    buf = io.BytesIO()
    with bb.vdif.open(buf, 'wb') as fw:
        for i in range(16):
            header = header0.copy()
            header.mutable = True
            if i != 10:
                header['frame_nr'] = i
            else:
                header['frame_nr'] = 0
                header['seconds'] = 0
                header['invalid_data'] = True
            fw.write_frameset(data, header=header)
        raw = buf.getvalue()
    got, _ = _read(raw)
    got = got.reshape((-1,) + data.shape)
    assert len(got) == 16
    assert np.all(got[:10] == data) and np.all(got[11:] == data)
    assert np.all(got[10] == 0.)


# ------------------------------- Mark 5B (mark5b/tests/test_corrupt_files.py)
def _m5b_expected_bad_frames(missing):
    """:24-33: frames touched, and the one before if the sync is touched."""
    (start_f, start_r), (stop_f, _) = [divmod(s, 10016)
                                       for s in (missing.start,
                                                 missing.stop - 1)]
    if start_r < 5:
        start_f -= 1
    return start_f, stop_f + 1


M5B_BAD_BYTES = ((slice(20032, 20033), b''), (slice(20096, 20100), b''),
                 (slice(12000, 22000), b''), (slice(30060, 30070), b''),
                 (slice(40063, 40064), b''),
                 (slice(20032, 20033), b'\xff'),   # corrupt sync of header 2
                 (slice(20032, 20036), b'\xff'),   # ... and wrong length
                 (slice(20040, 20041), b'\xff'))   # first time byte of hdr 2


def m5b_sample_bad_bytes(affected, replacement):
    """test_bad_bytes (:77-152): sample.m5b with bytes missing or replaced,
    four more (invalid) frames appended so that the end of the stream is
    well defined."""
    sample = open(sample_path('sample.m5b'), 'rb').read()
    with bb.mark5b.open(sample_path('sample.m5b'), 'rs', sample_rate=32e6,
                        kday=56000, nchan=8, bps=2) as fs:
        start_time, stop_time = fs.start_time, fs.stop_time
        frame_rate = fs._frame_rate
        data = fs.read()
    with bb.mark5b.open(sample_path('sample.m5b'), 'rb', kday=56000,
                        nchan=8, bps=2) as fb:
        fb.seek(3 * 10016)
        frame3 = fb.read_frame()
    corrupted = sample[:affected.start] + replacement + sample[affected.stop:]
    bad_start, bad_stop = _m5b_expected_bad_frames(affected)
    buf = io.BytesIO()
    buf.write(corrupted)
    for i in range(4, 8):
        header = frame3.header.copy()
        header.mutable = True
        header.set_time(start_time + i / frame_rate, frame_rate=frame_rate)
        header.update()
        type(frame3)(header, frame3.payload, valid=False).tofile(buf)
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.mark5b.open(io.BytesIO(buf.getvalue()), 'rs',
                            sample_rate=32e6, kday=56000, nchan=8,
                            bps=2) as fr:
            assert fr.start_time == start_time
            assert abs(fr.stop_time - stop_time - 4 / frame_rate) < 1e-9
            got = fr.read()
    assert got.shape == (40000, 8)
    expected = data.copy().reshape(-1, 5000, 8)
    expected[bad_start:bad_stop] = 0.
    expected = expected.reshape(-1, 8)
    expected = np.concatenate((expected, np.zeros_like(expected)))
    assert np.array_equal(got, expected)


def _m5b_fake():
    """:157-176: 16 frames of 2 channels."""
    header0 = bb.mark5b.Mark5BHeader.fromvalues(time='2010-11-12T13:14:15')
    data = np.repeat(np.array([[-1, 1], [-HIGH, HIGH]], np.float32), 10000,
                     axis=0)
    buf = io.BytesIO()
    with bb.mark5b.open(buf, 'ws', header0=header0, sample_rate=1e5,
                        nchan=2) as fw:
        for _ in range(16):
            fw.write(data)
        raw = buf.getvalue()
    assert len(raw) == 16 * 10016
    return raw, data


M5B_KW = dict(nchan=2, sample_rate=1e5, ref_time='2010-11-12T13:14:15')


def m5b_fake_missing_frames(frame_nr):
    """test_missing_frames (:192-207)."""
    if not isinstance(frame_nr, slice):
        frame_nr = slice(frame_nr, frame_nr + 1)
    raw, data = _m5b_fake()
    blob = _cut(raw, slice(frame_nr.start * 10016, frame_nr.stop * 10016))
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.mark5b.open(io.BytesIO(blob), 'rs', **M5B_KW) as fr:
            got = fr.read()
    got = got.reshape((-1,) + data.shape)
    assert len(got) == 16
    assert np.all(got[:frame_nr.start] == data)
    assert np.all(got[frame_nr.stop:] == data)
    assert np.all(got[frame_nr] == 0.)


M5B_MISSING_MIDDLE = ((slice(10016, 20032), slice(1, 2)),
                      (slice(20000, 20501), slice(1, 3)),
                      (slice(20032, 20048), slice(1, 3)))


def m5b_fake_missing_middle(missing_bytes, missing_frames):
    """test_missing_middle (:258-276)."""
    raw, data = _m5b_fake()
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.mark5b.open(io.BytesIO(_cut(raw, missing_bytes)), 'rs',
                            **M5B_KW) as fr:
            assert fr.size == 16 * data.size
            got = fr.read()
    got = got.reshape((-1,) + data.shape)
    expected = np.stack([data] * 16)
    expected[missing_frames] = 0.
    assert np.array_equal(got, expected)


# ----------------------------------- Mark 4 (mark4/tests/test_corrupt_files.py)
M4_MISSING_FRAMES = (1, 3, slice(3, 5))


def m4_fake_missing_frames(frame_nr):
    """test_missing_frames (:47-80): 8 frames, 16 tracks, 2 channels."""
    if not isinstance(frame_nr, slice):
        frame_nr = slice(frame_nr, frame_nr + 1)
    header0 = bb.mark4.Mark4Header.fromvalues(
        time='2010-11-12T13:14:15', ntrack=16, nchan=2, fanout=4)
    fb = header0.frame_nbytes
    data = np.zeros((2 * fb, 2), np.float32)
    data.reshape(-1, 4, 2)[160:] = [[-1, 1], [-HIGH, HIGH], [1, -1],
                                    [HIGH, -HIGH]]
    buf = io.BytesIO()
    with bb.mark4.open(buf, 'ws', header0=header0, sample_rate=1e5) as fw:
        for _ in range(8):
            fw.write(data)
        raw = buf.getvalue()
    assert len(raw) == 8 * fb
    blob = _cut(raw, slice(frame_nr.start * fb, frame_nr.stop * fb))
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        with bb.mark4.open(io.BytesIO(blob), 'rs', sample_rate=1e5,
                           ref_time='2010-11-12T13:14:15') as fr:
            got = fr.read()
    got = got.reshape((-1,) + data.shape)
    assert len(got) == 8
    assert np.all(got[:frame_nr.start] == data)
    assert np.all(got[frame_nr.stop:] == data)
    assert np.all(got[frame_nr] == 0.)
