"""Experiment: cluster the reads of a 1:16 expanding stream in time with
cp.async.bulk.prefetch.L2 issued by one thread per `lead` CTAs, `ahead` CTAs
before the data is used (bb_probe_expand pattern 4 vs pattern 0)."""
import os
import sys

import torch

sys.path.insert(0, '.')
from baseband_b200 import kernels  # noqa: E402
from tools.sweep_decode import timeit  # noqa: E402

DEV = 'cuda:0'
n = 8 << 30
dst = torch.empty(n, dtype=torch.uint8, device=DEV)
src = torch.randint(0, 256, (n // 16,), dtype=torch.uint8, device=DEV)
best, med = timeit(lambda: kernels.probe_expand(dst, src, 0))
print('pattern 0 (no prefetch):                 %7.1f GB/s' % (n * 17 / 16 / best / 1e6))
for lead in (64, 256, 512, 2048, 8192):
    for ahead in (1184, 2368, 4736, 16384):
        os.environ['BB_PROBE_LEAD'] = str(lead)
        os.environ['BB_PROBE_AHEAD'] = str(ahead)
        best, med = timeit(lambda: kernels.probe_expand(dst, src, 4))
        print('lead %5d CTAs (%5d KiB), ahead %5d:   %7.1f GB/s' % (
            lead, lead * 2, ahead, n * 17 / 16 / best / 1e6), flush=True)
