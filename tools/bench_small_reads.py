"""Loop of small reads through the public API (the reference's frame-cache
usage pattern): read(100) over the whole of sample.vdif-like data."""
import io
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import synthetic  # noqa: E402
from baseband_b200.base import stream  # noqa: E402

raw = synthetic.vdif_stream(64, 8, 5000, seed=1).tobytes()   # 64 frame sets
for label, small in (('window cache (default)', 256 << 10),
                     ('GPU round trip per call', 0)):
    stream.StreamReaderBase.SMALL_READ_NBYTES = small
    fh = bb.vdif.open(io.BytesIO(raw), 'rs', sample_rate=32e6)
    fh.read(1)
    for n in (12, 100, 1000):
        fh.seek(0)
        t0 = time.perf_counter()
        calls = 0
        while fh.tell() + n <= fh.shape[0]:
            fh.read(n)
            calls += 1
        dt = time.perf_counter() - t0
        print('%-26s read(%4d) x %6d: %7.1f us per call, %6.2f Msamples/s'
              % (label, n, calls, dt / calls * 1e6,
                 calls * n * 8 / dt / 1e6))
