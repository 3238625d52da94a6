import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import baseband_b200 as bb
from baseband_b200 import synthetic, kernels
from baseband_b200.base import stream
dev = 'cuda:0'
path = '/dev/shm/bb_bench.vdif'
nset = (1 << 30) // (16 * 8032)
raw = synthetic.vdif_stream(nset, 16, 8000, seed=1); raw.tofile(path); nbytes = raw.size
T = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            T[key] = T.get(key, 0) + time.perf_counter() - t0
    setattr(obj, name, g)
wrap(stream, 'read_file_into', 'read_file_into')
wrap(torch.cuda.Event, 'synchronize', 'event.sync')
wrap(kernels, 'vdif_scan', 'vdif_scan')
wrap(kernels, 'decode_bitfield', 'decode')
for threads in (4, 8):
    stream.PARALLEL_READ_THREADS = threads
    fh = bb.vdif.open(path, 'rs', sample_rate=64e6, device=dev)
    wrap(fh, '_upload', 'upload')
    for rep in range(3):
        T.clear()
        fh.seek(0); torch.cuda.synchronize(); t0 = time.perf_counter()
        data = fh.read(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(threads, 'total %.1f ms (host loop %.1f), %.1f GB/s' % ((t2-t0)*1e3, (t1-t0)*1e3, nbytes/(t2-t0)/1e9), {k: round(v*1e3,1) for k,v in T.items()})
    del data; fh.close()
os.remove(path)
