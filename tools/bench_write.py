"""Public-API write throughput: a large host float32 array through
vdif.open(..., 'ws').write() (H2D, encode_2bit, D2H of the frames)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200.base.memory import HostBuffer  # noqa: E402

dev = 'cuda:0'
nthread, spf = 16, 32000
nset = 2000                                    # 4.1 GB of float32
rng = np.random.default_rng(1)
data = rng.standard_normal((nset * spf // 8, nthread)).astype(np.float32)
data = np.tile(data, (8, 1)) * 2.0
h0 = bb.vdif.VDIFHeader.fromvalues(
    edv=0, time='2020-01-01T00:00:00', nchan=1, bps=2, complex_data=False,
    thread_id=0, samples_per_frame=spf, station='bb', frame_nr=0)
print('host array: %.2f GB float32' % (data.nbytes / 1e9))
from baseband_b200 import device as bbdev  # noqa: E402
for label, min_nbytes, threads in (('torch pageable copy', 1 << 60, 1),
                                   ('staged upload, 1 thread', 32 << 20, 1),
                                   ('staged upload, 4 threads', 32 << 20, 4),
                                   ('staged upload, 8 threads', 32 << 20, 8)):
    bbdev.STAGED_UPLOAD_MIN_NBYTES = min_nbytes
    bbdev.STAGED_UPLOAD_THREADS = threads
    best = 1e9
    for rep in range(3):
        sink = HostBuffer(nset * nthread * 8032)
        fw = bb.vdif.open(sink, 'ws', header0=h0, nthread=nthread,
                          sample_rate=64e6, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fw.write(data)
        fw._flush(final=False)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print('%-28s %6.2f Gsamp/s  (%5.1f GB/s host->device)'
          % (label, data.size / best / 1e9, data.nbytes / best / 1e9))
# the frames decode back to the quantised data
sink.seek(0)
fr = bb.vdif.open(sink, 'rs', sample_rate=64e6, device=dev)
back = fr.read(spf * 4)
lv = np.array([-3.316505, -1., 1., 3.316505], np.float32)
assert set(np.unique(back.cpu().numpy()).tolist()) <= set(lv.tolist())
print('read back ok')
