"""Device-resident throughput sweep of the bit-field decode/encode kernels
(development tool; bench.py is the contract benchmark)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from baseband_b200 import kernels, levels  # noqa: E402

DEV = 'cuda:0'


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[0], ts[len(ts) // 2]


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
    print(torch.cuda.get_device_name(0))
    # calibration: plain copy and fill
    n = int(8 * 2**30)
    a = torch.empty(n // 2, dtype=torch.uint8, device=DEV)
    b = torch.empty(n // 2, dtype=torch.uint8, device=DEV)
    best, med = timeit(lambda: b.copy_(a))
    print('copy   %6.1f GB/s (best) %6.1f (median)' % (n / best / 1e6,
                                                      n / med / 1e6))
    best, med = timeit(lambda: b.fill_(1))
    print('fill   %6.1f GB/s (best) %6.1f (median)' % (n / 2 / best / 1e6,
                                                      n / 2 / med / 1e6))
    del a, b
    configs = [
        ('C2 vdif 2bit 16thr x 8000B', 2, 16, 1, 8000, 32, 'vdif'),
        ('C1 vdif 2bit 8thr x 5000B', 2, 8, 1, 5000, 32, 'vdif'),
        ('vdif 2bit 1thr 1ch', 2, 1, 1, 8000, 32, 'vdif'),
        ('C5 mark5b 2bit 16ch', 2, 1, 16, 10000, 16, 'mark5b'),
        ('vdif 1bit 8thr', 1, 8, 1, 8000, 32, 'vdif'),
        ('vdif 4bit cplx 1024ch', 4, 1, 2048, 8192, 32, 'vdif'),
        ('vdif 8bit 2thr cplx', 8, 2, 2, 8000, 32, 'vdif'),
        ('vdif 2bit 4thr 8ch', 2, 4, 8, 8000, 32, 'vdif'),
        ('dada int8 cplx 2pol', 8, 1, 4, 1 << 20, 0, 'sint'),
        ('vdif 4bit 4thr 1ch', 4, 4, 1, 8000, 32, 'vdif'),
        ('vdif 8bit 4thr 1ch', 8, 4, 1, 8000, 32, 'vdif'),
        ('vdif 4bit 2thr cplx', 4, 2, 2, 8000, 32, 'vdif'),
        ('vdif 2bit 2thr cplx', 2, 2, 2, 8000, 32, 'vdif'),
        ('gsb phased 8bit 2pol 512ch', 8, 2, 1024, 1 << 22, 0, 'sint'),
        ('gsb rawdump 4bit', 4, 1, 1, 1 << 22, 0, 'sint'),
    ]
    if len(sys.argv) > 2:
        configs = [c for c in configs if sys.argv[2] in c[0]]
    for name, bps, nthread, nelem, payload, hdr, kind in configs:
        frame = payload + hdr
        nunit = int(gib * 2**30) // frame
        nset = nunit // nthread
        nunit = nset * nthread
        raw = torch.randint(0, 256, (nunit * frame,), dtype=torch.uint8,
                            device=DEV)
        off = (torch.arange(nunit, dtype=torch.int64, device=DEV) * frame
               + hdr)
        if kind == 'sint':
            lv, cd, q = None, 1, 2
        elif kind == 'mark5b':
            lv, cd, q = levels.mark5b(bps), 0, 1
        else:
            lv, cd, q = levels.offset_binary(bps), 0, 0
        spf = payload * 8 // (bps * nelem)
        out = torch.empty((nset * spf, nthread, nelem), dtype=torch.float32,
                          device=DEV)
        fn = lambda: kernels.decode_bitfield(raw, off, nset, nthread, payload,
                                             bps, nelem, False, cd, lv,
                                             out=out)
        best, med = timeit(fn)
        nbytes = nunit * frame + out.numel() * 4
        nsamp = out.numel()
        print('DEC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
              % (name, nbytes / best / 1e6, nbytes / med / 1e6,
                 nsamp / med / 1e6, med))
        back = torch.zeros_like(raw)
        fn = lambda: kernels.encode_bitfield(out, back, off, nset, nthread,
                                             payload, bps, nelem, q)
        best, med = timeit(fn)
        print('ENC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
              % (name, nbytes / best / 1e6, nbytes / med / 1e6,
                 nsamp / med / 1e6, med))
        ok = bool(torch.equal(back.view(-1, frame)[:, hdr:],
                              raw.view(-1, frame)[:, hdr:]))
        print('    round trip identical:', ok)
        del raw, off, out, back
        torch.cuda.empty_cache()


def extra(gib):
    # C3: Mark 4 64-track fanout 4
    nframe = int(gib * 2**30) // 160000
    raw = torch.randint(0, 256, (nframe * 160000,), dtype=torch.uint8,
                        device=DEV)
    off = torch.arange(nframe, dtype=torch.int64, device=DEV) * 160000 + 1280
    out = torch.empty((nframe * 80000, 8), dtype=torch.float32, device=DEV)
    lv = levels.sign_magnitude()
    best, med = timeit(lambda: kernels.mark4_decode(raw, off, nframe, 8, 4,
                                                    False, lv, out=out))
    nbytes = raw.numel() + out.numel() * 4
    print('DEC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
          % ('C3 mark4 64trk fanout4', nbytes / best / 1e6, nbytes / med / 1e6,
             out.numel() / med / 1e6, med))
    back = raw.clone()
    back.view(nframe, 160000)[:, 1280:] = 0
    best, med = timeit(lambda: kernels.mark4_encode(out, back, off, nframe, 8,
                                                    4, False))
    print('ENC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
          % ('C3 mark4 64trk fanout4', nbytes / best / 1e6, nbytes / med / 1e6,
             out.numel() / med / 1e6, med))
    print('    round trip identical:', bool(torch.equal(back, raw)))
    del raw, off, out, back
    torch.cuda.empty_cache()
    # C4: GUPPI 512 chan x 2 pol complex int8, channels first, overlap 512
    nchan, npol, spf, ov = 512, 2, 65536, 512
    fbytes = nchan * spf * npol * 2
    nfr = max(1, int(gib * 2**30) // fbytes)
    raw = torch.randint(0, 256, (nfr * fbytes,), dtype=torch.uint8, device=DEV)
    off = torch.arange(nfr, dtype=torch.int64, device=DEV) * fbytes
    cb = torch.full((nfr,), ov * npol, dtype=torch.int64, device=DEV)
    cb[0] = 0
    ce = torch.full((nfr,), spf * npol, dtype=torch.int64, device=DEV)
    oc0 = torch.cumsum(ce - cb, 0) - (ce - cb)
    ncols = int((ce - cb).sum().item())
    out = torch.empty((ncols * nchan * 2,), dtype=torch.float32, device=DEV)
    best, med = timeit(lambda: kernels.decode_int8_transposed(
        raw, off, nfr, nchan, spf * npol, 2, cb, ce, oc0, out))
    nbytes = out.numel() + out.numel() * 4
    print('DEC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
          % ('C4 guppi 512ch 2pol int8', nbytes / best / 1e6,
             nbytes / med / 1e6, out.numel() / med / 1e6, med))
    # encode: whole frames (writers have no overlap)
    del out
    cb0 = torch.zeros(nfr, dtype=torch.int64, device=DEV)
    oc = torch.arange(nfr, dtype=torch.int64, device=DEV) * (spf * npol)
    full = torch.empty((nfr * spf * npol * nchan * 2,), dtype=torch.float32,
                       device=DEV)
    kernels.decode_int8_transposed(raw, off, nfr, nchan, spf * npol, 2, cb0,
                                   ce, oc, full)
    back = torch.zeros_like(raw)
    best, med = timeit(lambda: kernels.encode_int8_transposed(
        full, back, off, nfr, nchan, spf * npol, 2))
    nbytes = full.numel() * 5
    print('ENC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
          % ('C4 guppi 512ch 2pol int8', nbytes / best / 1e6,
             nbytes / med / 1e6, full.numel() / med / 1e6, med))
    print('    round trip identical:', bool(torch.equal(back, raw)))
    del full, back
    torch.cuda.empty_cache()
    # GUPPI time first (PKTFMT other than 1SFA): [time][chan][pol] per frame
    nt = fbytes // (nchan * npol * 2)
    tb = torch.zeros(nfr, dtype=torch.int64, device=DEV)
    te = torch.full((nfr,), nt, dtype=torch.int64, device=DEV)
    t0 = torch.arange(nfr, dtype=torch.int64, device=DEV) * nt
    out = torch.empty((nfr * nt * npol * nchan * 2,), dtype=torch.float32,
                      device=DEV)
    best, med = timeit(lambda: kernels.decode_int8_timefirst(
        raw, off, nfr, nt, nchan, npol, 2, tb, te, t0, out))
    nbytes = out.numel() * 5
    print('DEC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
          % ('guppi time-first 512ch 2pol', nbytes / best / 1e6,
             nbytes / med / 1e6, out.numel() / med / 1e6, med))
    back = torch.zeros_like(raw)
    best, med = timeit(lambda: kernels.encode_int8_timefirst(
        out, back, off, nfr, nt, nchan, npol, 2))
    print('ENC %-28s %7.1f GB/s best %7.1f med  %7.1f Gsamp/s  (%.2f ms)'
          % ('guppi time-first 512ch 2pol', nbytes / best / 1e6,
             nbytes / med / 1e6, out.numel() / med / 1e6, med))
    print('    round trip identical:', bool(torch.equal(back, raw)))


if __name__ == '__main__':
    main()
    if len(sys.argv) <= 2:
        extra(float(sys.argv[1]) if len(sys.argv) > 1 else 0.5)
