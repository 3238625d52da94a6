"""Experiment: can a page-cache file be DMA'd without the copy into pinned
staging, by page-locking its mapping (cudaHostRegister, read-only) piece by
piece?  Prints the rate of registering / unregistering pieces on 1..T
threads and of the H2D copy out of the registered mapping.

    python tools/probe_register.py [MiB] [piece MiB] [rw]

(rw: map the file shared and writable instead of read-only)
"""
import ctypes
import glob
import mmap
import os
import sys
import threading
import time

import numpy as np
import torch

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
piece = (int(sys.argv[2]) if len(sys.argv) > 2 else 16) << 20
n = mib << 20
path = '/dev/shm/bb_probe_register.bin'
with open(path, 'wb') as fh:
    blk = np.random.default_rng(1).integers(0, 256, 1 << 24, dtype=np.uint8)
    for _ in range(n >> 24):
        fh.write(blk.tobytes())

torch.cuda.init()
dev = torch.device('cuda:0')
out = torch.empty(n, dtype=torch.uint8, device=dev)
libdir = os.path.join(os.path.dirname(torch.__file__), 'lib')
cand = glob.glob(os.path.join(libdir, 'libcudart*.so*')) + \
    glob.glob('/usr/local/cuda/lib64/libcudart.so*')
rt = ctypes.CDLL(cand[0])
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t,
                                ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
READ_ONLY = 0x08

writable = len(sys.argv) > 3 and sys.argv[3] == 'rw'
fd = os.open(path, os.O_RDWR if writable else os.O_RDONLY)
mm = mmap.mmap(fd, n, access=mmap.ACCESS_WRITE if writable
               else mmap.ACCESS_READ)
base = ctypes.addressof(ctypes.c_char.from_buffer_copy(b'x'))  # placeholder
arr = np.frombuffer(mm, np.uint8)
base = arr.ctypes.data


def run(nthr, fn):
    offs = list(range(0, n, piece))
    errs = []

    def work(k):
        for o in offs[k::nthr]:
            rc = fn(o)
            if rc:
                errs.append(rc)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(nthr)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return n / dt / 1e9, errs[:1]


for flags in (READ_ONLY, 0):
    for nthr in (1, 2, 4, 8):
        r, e = run(nthr, lambda o: rt.cudaHostRegister(base + o, piece, flags))
        if e:
            print('flags', flags, 'threads', nthr, 'register failed rc', e)
            run(nthr, lambda o: rt.cudaHostUnregister(base + o))
            break
        s = torch.cuda.current_stream().cuda_stream
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rt.cudaMemcpyAsync(out.data_ptr(), base, n, 1, s)
        torch.cuda.synchronize()
        h2d = n / (time.perf_counter() - t0) / 1e9
        ok = bool((out[:1 << 20].cpu().numpy() == arr[:1 << 20]).all())
        u, _ = run(nthr, lambda o: rt.cudaHostUnregister(base + o))
        print('flags {} threads {}: register {:.1f} GB/s, h2d {:.1f} GB/s '
              '(ok {}), unregister {:.1f} GB/s'.format(flags, nthr, r, h2d,
                                                       ok, u))
os.unlink(path)
