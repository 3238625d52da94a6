"""GUPPI transposed-decode timing for the current BB_I8_GROUP."""
import os, sys
import torch
sys.path.insert(0, '.')
from baseband_b200 import kernels
from tools.sweep_decode import timeit
DEV = 'cuda:0'
for spf, knock in ((65536, '0'), (65536, '1'), (65536, '2'), (60000, '0')):
    os.environ['BB_TUNE_KNOCK_I8'] = knock
    nchan, npol, ov = 512, 2, 512
    fbytes = nchan * spf * npol * 2
    nfr = 8
    raw = torch.randint(0, 256, (nfr * fbytes,), dtype=torch.uint8, device=DEV)
    off = torch.arange(nfr, dtype=torch.int64, device=DEV) * fbytes
    cb = torch.full((nfr,), ov * npol, dtype=torch.int64, device=DEV); cb[0] = 0
    ce = torch.full((nfr,), spf * npol, dtype=torch.int64, device=DEV)
    oc0 = torch.cumsum(ce - cb, 0) - (ce - cb)
    ncols = int((ce - cb).sum().item())
    out = torch.empty((ncols * nchan * 2,), dtype=torch.float32, device=DEV)
    best, med = timeit(lambda: kernels.decode_int8_transposed(
        raw, off, nfr, nchan, spf * npol, 2, cb, ce, oc0, out))
    nbytes = out.numel() * 5
    print('group %s spf %d knock %s (1 = no loads, 2 = no stores; algorithmic '
          'bytes): %7.1f GB/s best %7.1f med' % (
              os.environ.get('BB_I8_GROUP'), spf, knock, nbytes / best / 1e6,
              nbytes / med / 1e6))
os.environ['BB_TUNE_KNOCK_I8'] = '0'
