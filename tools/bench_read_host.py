"""Host-output read through the public API, by kind of destination."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import synthetic  # noqa: E402
from baseband_b200.base.memory import HostBuffer  # noqa: E402

nset = (128 << 20) // (16 * 8032)
src = HostBuffer(synthetic.vdif_stream(nset, 16, 8000, seed=1))
fh = bb.vdif.open(src, 'rs', sample_rate=64e6)
nsamp = nset * 32000 * 16
print('decoded output: %.2f GB float32' % (nsamp * 4 / 1e9))


def timed(label, fn, reps=4):
    ts = []
    for _ in range(reps):
        fh.seek(0)
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    print('%-44s first %6.2f  best %6.2f Gsamp/s' % (
        label, nsamp / ts[0] / 1e9, nsamp / min(ts) / 1e9))


timed('read() -> new array (out=None)', lambda: fh.read())
pageable = np.empty((nset * 32000, 16), np.float32)
pageable[:] = 0
timed('read(out=pageable numpy array)', lambda: fh.read(out=pageable))
pinned = torch.empty((nset * 32000, 16), dtype=torch.float32, pin_memory=True)
timed('read(out=pinned array)', lambda: fh.read(out=pinned.numpy()))
