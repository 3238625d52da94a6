"""Parametrised item tables for Payload, Frame and FrameSet of every format:
the reference's ``test_payload_getitem_setitem`` / ``test_frame*`` tables
(baseband/base/tests/test_base.py:264-340, vdif/tests/test_vdif.py:431-449 and
:601-692, mark4/tests/test_mark4.py:366-387, mark5b, dada, guppi likewise)
restated as data: (object kind) x (item).  Shared by the CPU-backend run
(tests/test_host_items.py, host logic of ``_item_to_slices`` & co.) and the
GPU run (tests/test_gpu_items.py, where every ``[...]`` decodes / encodes on
the device).

The check is the reference's: ``obj[item]`` equals ``obj.data[item]`` (numpy
semantics incl. negative indices, steps, channel and thread axes), setting an
item changes exactly that item, setting it back restores equality.
"""
import io

import numpy as np

import baseband_b200 as bb
from conftest import sample_path


# ------------------------------------------------------------ object makers
def _vdif_frame():
    with bb.vdif.open(sample_path('sample.vdif'), 'rb') as fh:
        return fh.read_frame()


def _vdif_payload():
    frame = _vdif_frame()
    return bb.vdif.VDIFPayload(frame.payload.words.copy(), frame.header)


def _level_data(shape, lv, seed, cplx=False):
    rng = np.random.default_rng(seed)
    data = np.asarray(lv, np.float32)[rng.integers(0, len(lv), shape)]
    if cplx:
        imag = np.asarray(lv, np.float32)[rng.integers(0, len(lv), shape)]
        data = (data + 1j * imag).astype(np.complex64)
    return data


def _vdif_payload_c4():
    # complex, 4 channels, 2 bit: one sample = 16 bits
    from baseband_b200 import levels
    data = _level_data((96, 4), levels.offset_binary(2), 3, cplx=True)
    return bb.vdif.VDIFPayload.fromdata(data, bps=2)


def _vdif_payload_8bit():
    # the geometry of the reference's base-class test: 8 bit, 2 channels
    words = np.arange(4 * 25, dtype=np.uint8).view('<u4').copy()
    return bb.vdif.VDIFPayload(words, sample_shape=(2,), bps=8,
                               complex_data=False)


def _vdif_payload_4bit():
    from baseband_b200 import levels
    data = _level_data((64, 2), levels.offset_binary(4), 4)
    return bb.vdif.VDIFPayload.fromdata(data, bps=4)


def _vdif_payload_1bit():
    from baseband_b200 import levels
    data = _level_data((128, 8), levels.offset_binary(1), 5)
    return bb.vdif.VDIFPayload.fromdata(data, bps=1)


def _mark5b_frame():
    with bb.mark5b.open(sample_path('sample.m5b'), 'rb', kday=56000,
                        nchan=8) as fh:
        return fh.read_frame()


def _mark5b_payload():
    frame = _mark5b_frame()
    return bb.mark5b.Mark5BPayload(frame.payload.words.copy(),
                                   sample_shape=(8,), bps=2)


def _mark4_frame():
    with bb.mark4.open(sample_path('sample.m4'), 'rb', ntrack=64,
                       decade=2010) as fh:
        fh.locate_frame()
        return fh.read_frame()


def _mark4_payload():
    frame = _mark4_frame()
    return bb.mark4.Mark4Payload(frame.payload.words.copy(), frame.header)


def _dada_frame():
    with bb.dada.open(sample_path('sample.dada'), 'rb') as fh:
        return fh.read_frame(memmap=False)


def _dada_payload():
    frame = _dada_frame()
    return bb.dada.DADAPayload(frame.payload.words.copy(),
                               header=frame.header)


def _guppi_frame():
    with bb.guppi.open(sample_path('sample_puppi.raw'), 'rb') as fh:
        return fh.read_frame(memmap=False)


def _guppi_payload():
    frame = _guppi_frame()
    return bb.guppi.GUPPIPayload(frame.payload.words.copy(),
                                 header=frame.header)


def _gsb_payload_4bit():
    # rawdump: real 4-bit samples, two per byte
    rng = np.random.default_rng(8)
    words = rng.integers(-128, 128, 256, dtype=np.int8)
    return bb.gsb.GSBPayload(words, sample_shape=(1,), bps=4,
                             complex_data=False)


def _gsb_payload_8bit():
    # phased: complex 8-bit, 2 polarisations x 8 channels
    rng = np.random.default_rng(9)
    words = rng.integers(-128, 128, 40 * 2 * 8 * 2, dtype=np.int8)
    return bb.gsb.GSBPayload(words, sample_shape=(2, 8), bps=8,
                             complex_data=True)


def _vdif_frameset():
    with bb.vdif.open(sample_path('sample.vdif'), 'rb') as fh:
        return fh.read_frameset()


MAKERS = {
    'vdif_payload': _vdif_payload, 'vdif_payload_c4': _vdif_payload_c4,
    'vdif_payload_8bit': _vdif_payload_8bit,
    'vdif_payload_4bit': _vdif_payload_4bit,
    'vdif_payload_1bit': _vdif_payload_1bit,
    'mark5b_payload': _mark5b_payload, 'mark4_payload': _mark4_payload,
    'dada_payload': _dada_payload, 'guppi_payload': _guppi_payload,
    'gsb_payload_4bit': _gsb_payload_4bit,
    'gsb_payload_8bit': _gsb_payload_8bit,
    'vdif_frame': _vdif_frame, 'mark5b_frame': _mark5b_frame,
    'mark4_frame': _mark4_frame, 'dada_frame': _dada_frame,
    'guppi_frame': _guppi_frame, 'vdif_frameset': _vdif_frameset,
}

# the reference's item tables (union over formats), by number of sample axes
ITEMS_1 = [2, (), -1, slice(1, 3), slice(2, 4), slice(-3, None),
           slice(2, None), slice(1, 1), slice(5, 40, 7), (slice(3, 9),)]
ITEMS_2 = ITEMS_1 + [(2, 1), (slice(None), 0), (slice(1, 3), 1),
                     (2, slice(0, 2)), (10, -1), (slice(None), slice(1, None)),
                     (slice(4, 20, 5), 0)]
ITEMS_3 = ITEMS_2 + [(15,), (slice(10, 20), slice(None), 0), (10, 1, 0),
                     (10, slice(None), slice(0, 1)), (slice(None), 1, 0),
                     (slice(None), slice(0, 2), slice(None))]
NAXES = {'vdif_payload': 2, 'vdif_payload_c4': 2, 'vdif_payload_8bit': 2,
         'vdif_payload_4bit': 2, 'vdif_payload_1bit': 2, 'mark5b_payload': 2,
         'mark4_payload': 2, 'dada_payload': 3, 'guppi_payload': 3,
         'gsb_payload_4bit': 2, 'gsb_payload_8bit': 3,
         'vdif_frame': 2, 'mark5b_frame': 2, 'mark4_frame': 2,
         'dada_frame': 3, 'guppi_frame': 3, 'vdif_frameset': 3}


def table():
    rows = []
    for kind, naxes in NAXES.items():
        for item in (ITEMS_1, ITEMS_2, ITEMS_3)[naxes - 1]:
            rows.append((kind, item))
    return rows


def _usable(item, shape):
    """Does the item address only axes/indices that exist for this shape?"""
    try:
        np.empty(shape, np.int8)[item]
    except IndexError:
        return False
    return True


def check_getitem(kind, item):
    obj = MAKERS[kind]()
    data = obj.data
    if not _usable(item, data.shape):
        try:
            obj[item]
        except IndexError:
            return
        raise AssertionError('expected IndexError for {!r}'.format(item))
    want = data[item]
    got = obj[item]
    assert got.shape == want.shape, (kind, item, got.shape, want.shape)
    assert got.dtype == want.dtype
    assert np.array_equal(got, want, equal_nan=True), (kind, item)


def check_setitem(kind, item):
    """Set the item to other representable values (its own values reversed,
    or those of the neighbouring block), check exactly that changed, set it
    back, check equality is restored."""
    make = MAKERS[kind]
    obj, ref = make(), make()
    data = ref.data
    if kind.endswith('_frame') or kind == 'vdif_frameset':
        # frames read from a file hold read-only words, as in the reference,
        # whose tests also set items on frames built with fromdata
        # (vdif/tests/test_vdif.py:637)
        header = ref.header0 if kind == 'vdif_frameset' else ref.header
        obj = type(ref).fromdata(data, header)
        ref = type(ref).fromdata(data, header)
    if not _usable(item, data.shape):
        try:
            obj[item] = 1.
        except IndexError:
            return
        raise AssertionError('expected IndexError for {!r}'.format(item))
    sel = data[item]
    if kind == 'mark4_frame':
        # the header-overwritten head of a Mark 4 frame is not stored
        # (mark4/frame.py:91-102): work beyond it
        return check_setitem_mark4_frame(obj, data, item)
    if np.ndim(sel) == 0:
        new = data.ravel()[7] if data.ravel()[7] != sel else data.ravel()[8]
    else:
        new = np.flip(sel).copy()
    check = data.copy()
    check[item] = new
    obj[item] = new
    assert np.array_equal(obj[item], np.asarray(new, data.dtype))
    assert np.array_equal(obj.data, check), (kind, item)
    if not np.array_equal(check, data):
        assert obj != ref
    obj[item] = sel
    assert np.array_equal(obj.data, data)
    assert obj == ref


def check_setitem_mark4_frame(obj, data, item):
    sel = data[item]
    new = np.flip(sel).copy() if np.ndim(sel) else data[700, 0]
    new = np.where(new == 0., np.float32(1.), new)   # fill is not a level
    obj[item] = new
    check = data.copy()
    check[item] = new
    check[:640] = data[:640]              # stays fill (invalid region)
    assert np.array_equal(obj.data, check)


# ------------------------------------------------------- error behaviour
def check_errors():
    """IndexError / TypeError / ValueError exactly where the reference raises
    them (base/tests/test_base.py:290-340)."""
    import pytest
    pl = _vdif_payload_8bit()                     # shape (50, 2)
    for item in (50, -51, (slice(None), 5), (0, 0, 0)):
        with pytest.raises(IndexError):
            pl[item]
        with pytest.raises(IndexError):
            pl[item] = 1
    with pytest.raises(TypeError):
        pl['l']
    with pytest.raises(TypeError):
        pl['l'] = 1
    for item, value in ((1, np.ones(10)), (1, np.ones((2, 2))),
                        (slice(1, 3), np.ones((2, 3))),
                        ((slice(1, 3), 0), np.ones((2, 2))),
                        ((slice(1, 3), slice(0, 1)), np.ones((1, 2)))):
        with pytest.raises(ValueError):
            pl[item] = value
    p11 = pl[1:1]
    assert p11.size == 0 and p11.shape == (0, 2) and p11.dtype == pl.dtype
    # a sample that does not start on a word boundary of a sub-byte payload
    # cannot be addressed when it would split a word between samples
    one = bb.vdif.VDIFPayload(np.zeros(5, '<u4'), sample_shape=(5,), bps=1,
                              complex_data=True)
    assert one.shape == (16, 5)
    with pytest.raises(TypeError):
        one[10:11]


def check_frameset_header_items():
    """FrameSet header access (vdif/tests/test_vdif.py:664-692)."""
    import pytest
    fs = _vdif_frameset()
    fs2 = bb.vdif.VDIFFrameSet.fromdata(fs.data, fs.header0)
    assert np.all(fs2['thread_id'] == [f.header['thread_id']
                                       for f in fs2.frames])
    assert fs2['frame_nr'] == fs2.header0['frame_nr']
    fs2['frame_nr'] = 25
    assert all(f.header['frame_nr'] == 25 for f in fs2.frames)
    fs2['thread_id'] = list(range(10, 18))
    assert all(fs2['thread_id'] == list(range(10, 18)))
    with pytest.raises(ValueError):
        fs2['thread_id'] = 0
    with pytest.raises(ValueError):
        fs2['thread_id'] = 0, 1, 2, 3, 4, 5, 6, 1
    with pytest.raises(ValueError):
        fs2['frame_nr'] = 0, 1, 0, 1, 0, 1, 0, 1
    assert fs2.valid
    mixed = True, True, False, False, True, True, False, False
    fs2.valid = mixed
    assert np.all(fs2.valid == mixed)
    fs2.valid = False
    assert not fs2.valid
    # fancy thread indices and broadcasting scalars (test_vdif.py:640-652)
    data = fs.data
    fs3 = bb.vdif.VDIFFrameSet.fromdata(data, fs.header0)
    fs3[()] = 1.
    assert np.all(fs3.data == 1.)
    fs3[:] = data
    fs3[10:20] = -1.
    assert np.all(fs3[10:20] == -1.)
    fs3[10:20:2] = data[10:20:2]
    assert np.all(fs3[10:20:2] == data[10:20:2])
    assert np.all(fs3[11:20:2] == -1.)
    fs3[:, [0, 4, 5, 6]] = data[:, :4]
    fs3[:, [1, 2, 3, 7]] = data[:, 4:]
    assert np.all(fs3[:, [0, 4, 5, 6, 1, 2, 3, 7]] == data)
    # written back in thread order: the file's frames (order 1,3,5,7,0,2,4,6)
    buf = io.BytesIO()
    fs.tofile(buf)
    raw = np.fromfile(sample_path('sample.vdif'),
                      np.uint8)[:8 * 5032].reshape(8, 5032)
    order = np.argsort([1, 3, 5, 7, 0, 2, 4, 6])
    assert np.array_equal(np.frombuffer(buf.getvalue(), np.uint8),
                          raw[order].ravel())
