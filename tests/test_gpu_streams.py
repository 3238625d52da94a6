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
