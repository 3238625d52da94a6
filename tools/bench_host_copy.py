"""The host-side ceiling of file ingest: how fast can this box's CPUs move
bytes from the page cache into pinned memory?  (a) os.preadv slices on T
threads (what base/stream.py does), (b) threaded memcpy from an mmap of the
same file (pages already faulted in), (c) plain threaded memcpy between
anonymous buffers (the memory system itself).  Payload GB/s."""
import mmap
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, '.')
GIB = 1 << 30
path = sys.argv[1] if len(sys.argv) > 1 else '/dev/shm/bb_copy_bench.bin'
src = np.random.default_rng(1).integers(0, 256, GIB, dtype=np.uint8)
src.tofile(path)
pinned = torch.empty(GIB, dtype=torch.uint8, pin_memory=True).numpy()
ncpu = len(os.sched_getaffinity(0))
print('host CPUs available:', ncpu)
fd = os.open(path, os.O_RDONLY)
mm = mmap.mmap(fd, GIB, prot=mmap.PROT_READ)
mapped = np.frombuffer(mm, np.uint8)
mapped[::4096].sum()                                   # fault every page in


def run(label, piece, threads):
    pool = ThreadPoolExecutor(threads)
    step = -(-GIB // threads) // 4096 * 4096 + 4096
    spans = [(lo, min(lo + step, GIB)) for lo in range(0, GIB, step)]
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        list(pool.map(piece, spans))
        best = min(best, time.perf_counter() - t0)
    pool.shutdown()
    print('%-34s %2d thread(s): %6.1f GB/s' % (label, threads,
                                               GIB / best / 1e9), flush=True)


def by_preadv(span):
    lo, hi = span
    mv = memoryview(pinned[lo:hi])
    got = 0
    while got < hi - lo:
        got += os.preadv(fd, [mv[got:]], lo + got)


def by_mmap(span):
    lo, hi = span
    np.copyto(pinned[lo:hi], mapped[lo:hi])


def by_memcpy(span):
    lo, hi = span
    np.copyto(pinned[lo:hi], src[lo:hi])


try:
    for threads in [t for t in (1, 2, 4, 8, 12, 16, 24, 32) if t <= ncpu]:
        run('preadv page cache -> pinned', by_preadv, threads)
        run('memcpy mmap(page cache) -> pinned', by_mmap, threads)
        run('memcpy anonymous -> pinned', by_memcpy, threads)
finally:
    os.close(fd)
    os.remove(path)
