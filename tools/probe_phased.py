"""Experiment: does separating the reads from the writes in time lift a 1:16
expanding stream above its mixed-traffic ceiling?  Per slice: a pure-read
prefetch of the slice's input into L2 (evict_last), then the expand kernel,
which then finds its input in L2.  usage: probe_phased.py [out_gib]"""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from baseband_b200 import kernels, _lib  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
dev = torch.device('cuda:0')
n = int(gib * 2**30)
dst = torch.empty(n, dtype=torch.uint8, device=dev)
src = torch.randint(0, 256, (n // 16,), dtype=torch.uint8, device=dev)
lib = _lib.load()
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(slice_in_mib, prefetch, pattern=0):
    s_in = int(slice_in_mib * 2**20)
    nslice = (n // 16) // s_in

    def once():
        for i in range(nslice):
            sp = ctypes.c_void_p(src.data_ptr() + i * s_in)
            if prefetch:
                assert lib.bb_probe_prefetch(sp, s_in, stream) == 0
            assert lib.bb_probe_expand(
                ctypes.c_void_p(dst.data_ptr() + i * s_in * 16), s_in * 16,
                sp, pattern, stream) == 0
    once()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        once()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    total = nslice * s_in * 17
    print('slice %6.1f MiB in, %4d slices, prefetch %d: %8.1f GB/s'
          % (slice_in_mib, nslice, prefetch, total / best / 1e6), flush=True)


for mib in (n / 16 / 2**20, 64, 32, 16, 8, 4):
    for pf in (0, 1):
        run(mib, pf)
