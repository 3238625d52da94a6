"""File (page cache) -> device ingest vs pipeline chunk size and read threads,
and the H2D rate while host threads copy (do they share a bottleneck?)."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import synthetic  # noqa: E402
from baseband_b200.base import stream  # noqa: E402

dev = 'cuda:0'
path = '/dev/shm/bb_bench.vdif'
nset = (1 << 30) // (16 * 8032)
raw = synthetic.vdif_stream(nset, 16, 8000, seed=1)
raw.tofile(path)
nbytes = raw.size
try:
    for threads in (4, 8, 12):
        for mib in (2, 4, 8, 16, 32, 64, 256):
            stream.PARALLEL_READ_THREADS = threads
            stream.PARALLEL_READ_MIN_NBYTES = 1 << 20
            fh = bb.vdif.open(path, 'rs', sample_rate=64e6, device=dev,
                              chunk_nbytes=mib << 20)
            best = 1e9
            for rep in range(3):
                fh.seek(0)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                data = fh.read()
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            print('threads %2d chunk %3d MiB: %5.1f GB/s' % (
                threads, mib, nbytes / best / 1e9), flush=True)
            fh.close()
            del data
    # H2D alone and while T threads memcpy 1 GiB elsewhere
    GIB = 1 << 30
    h = torch.empty(GIB, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(GIB, dtype=torch.uint8, device=dev)
    a = np.frombuffer(raw.tobytes() + bytes(GIB - raw.size), np.uint8)
    b = torch.empty(GIB, dtype=torch.uint8, pin_memory=True).numpy()
    for nthr in (0, 2, 4, 8):
        stop = threading.Event()
        moved = [0] * max(nthr, 1)

        def work(i):
            lo = GIB // nthr * i
            while not stop.is_set():
                np.copyto(b[lo:lo + GIB // nthr], a[lo:lo + GIB // nthr])
                moved[i] += GIB // nthr
        th = [threading.Thread(target=work, args=(i,)) for i in range(nthr)]
        [t.start() for t in th]
        time.sleep(0.05)
        m0 = sum(moved)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(6):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        m1 = sum(moved)
        stop.set()
        [t.join() for t in th]
        print('H2D %5.1f GB/s while %d host threads copy at %5.1f GB/s'
              % (6 * GIB / dt / 1e9, nthr, (m1 - m0) / dt / 1e9), flush=True)
finally:
    os.remove(path)
