"""A/B on one box: file (page cache) -> device ingest with two or three
pipeline stages, over read threads and chunk sizes."""
import sys
import time

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
import os  # noqa: E402
try:
    for rnd in range(2):
        for threads in (8, 12, 16):
            for mib in (8, 16, 32):
                line = 'threads %2d chunk %2d MiB:' % (threads, mib)
                for nstage in (2, 3):
                    stream.DEVICE_READ_STAGES = nstage
                    stream.PARALLEL_READ_THREADS = threads
                    fh = bb.vdif.open(path, 'rs', sample_rate=64e6,
                                      device=dev, chunk_nbytes=mib << 20)
                    best = 1e9
                    for rep in range(4):
                        fh.seek(0)
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        data = fh.read()
                        torch.cuda.synchronize()
                        best = min(best, time.perf_counter() - t0)
                    fh.close()
                    del data
                    line += '  %d stages %5.1f GB/s' % (nstage,
                                                        nbytes / best / 1e9)
                print(line, flush=True)
finally:
    os.remove(path)
