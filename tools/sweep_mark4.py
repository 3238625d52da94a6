"""Throughput of every Mark 4 mode (payload-word entry points)."""
import sys
import torch
sys.path.insert(0, '.')
from baseband_b200 import kernels, levels
from tools.sweep_decode import timeit
DEV = 'cuda:0'
lv = levels.sign_magnitude()
for nchan, fanout, ft in ((8, 4, False), (4, 4, False), (2, 4, False),
                          (8, 2, False), (16, 2, True)):
    ntrack = nchan * 2 * fanout
    wb = ntrack // 8
    nword = (512 << 20) // wb
    words = torch.randint(0, 256, (nword * wb,), dtype=torch.uint8, device=DEV)
    out = torch.empty((nword * fanout, nchan), dtype=torch.float32, device=DEV)
    best, med = timeit(lambda: kernels.mark4_decode_words(
        words, nword, nchan, fanout, ft, lv, out=out))
    nbytes = words.numel() + out.numel() * 4
    back = torch.zeros_like(words)
    beste, mede = timeit(lambda: kernels.mark4_encode_words(
        out, back, nword, nchan, fanout, ft))
    print('mark4 %2dch fanout%d%s (%2d tracks): DEC %7.1f GB/s  ENC %7.1f GB/s  '
          'round trip %s' % (nchan, fanout, ' ft' if ft else '   ', ntrack,
                             nbytes / med / 1e6, nbytes / mede / 1e6,
                             bool(torch.equal(back, words))))
