"""Stream-level GPU parity: the public API (open / read / write / Frame /
Payload) against the oracle and the reference's golden sample outputs."""
import pytest

import stream_cases
from test_host_streams import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', CASES)
def test_case(name):
    fn = getattr(stream_cases, name)
    if 'dev' in fn.__code__.co_varnames[:fn.__code__.co_argcount]:
        fn('cuda:0')
    else:
        fn()


def test_empty_and_edge_reads():
    """Zero-length reads, reads ending exactly at EOF, EOFError past the end,
    reading after close, and wrong ``out`` shapes."""
    import io
    import numpy as np
    import baseband_b200 as bb
    from baseband_b200 import synthetic
    raw = synthetic.vdif_stream(3, 8, 5000, seed=2)
    fh = bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=32e6)
    assert fh.read(0).shape == (0, 8)
    fh.seek(0, 2)
    assert fh.tell() == 60000 and fh.read().shape == (0, 8)
    fh.seek(-1, 2)
    assert fh.read().shape == (1, 8)
    with pytest.raises(EOFError):
        fh.read(1)
    fh.seek(0)
    with pytest.raises(AssertionError):
        fh.read(out=np.empty((10, 7), np.float32))
    out = np.empty((5, 8), np.float64)            # wrong dtype: still filled
    fh.read(out=out)
    fh.seek(0)
    assert np.array_equal(out, fh.read(5).astype(np.float64))
    fh.close()
    with pytest.raises(ValueError):
        fh.read(1)
    with pytest.raises(ValueError):
        bb.vdif.open(io.BytesIO(raw.tobytes()), 'rx')
    dev = bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=32e6,
                       device='cuda:0')
    assert tuple(dev.read(0).shape) == (0, 8)
    with pytest.raises(ValueError):               # no CPU decode path
        bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=32e6,
                     device='cpu').read(1)


def test_writer_large_host_array_staged_upload(monkeypatch):
    """write() of a large host array goes to the GPU in slices through the
    threaded pinned staging of device.staged_upload: same bytes as one small
    write, for float32, float64 and complex input."""
    import io
    import numpy as np
    import baseband_b200 as bb
    from baseband_b200 import device as bbdev
    from baseband_b200.base import stream
    monkeypatch.setattr(bbdev, 'STAGED_UPLOAD_MIN_NBYTES', 1 << 10)
    monkeypatch.setattr(bbdev, 'STAGED_UPLOAD_PIECE_NBYTES', 40000)
    monkeypatch.setattr(bbdev, 'STAGED_UPLOAD_THREADS', 3)
    rng = np.random.default_rng(12)
    data = (rng.standard_normal((20 * 4000, 4)) * 2.5).astype(np.float32)
    h0 = bb.vdif.VDIFHeader.fromvalues(
        edv=0, time='2020-01-01T00:00:00', nchan=1, bps=2,
        complex_data=False, thread_id=0, samples_per_frame=4000,
        station='bb', frame_nr=0)

    def written(arr, slice_nbytes):
        monkeypatch.setattr(stream.StreamWriterBase,
                            'HOST_WRITE_SLICE_NBYTES', slice_nbytes)
        buf = io.BytesIO()
        fw = bb.vdif.open(buf, 'ws', header0=h0, nthread=4, sample_rate=1e6)
        fw.write(arr)                 # whole frames: all flushed by now
        return buf.getvalue()

    monkeypatch.setattr(bbdev, 'STAGED_UPLOAD_MIN_NBYTES', 1 << 60)
    want = written(data, 1 << 60)                 # plain torch copy
    monkeypatch.setattr(bbdev, 'STAGED_UPLOAD_MIN_NBYTES', 1 << 10)
    assert written(data, 1 << 60) == want         # staged, one slice
    assert written(data, 100000) == want          # staged, ragged slices
    assert written(data.astype(np.float64), 100000) == want
    # staged_upload itself, odd sizes and dtypes
    for dtype in (np.float32, np.float64, np.complex64, np.uint8):
        arr = (rng.standard_normal((1237, 3)) * 50).astype(dtype)
        got = bbdev.staged_upload(arr, 'cuda:0').cpu().numpy()
        assert got.dtype == arr.dtype and np.array_equal(got, arr)
