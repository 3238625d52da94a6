"""Round-2 kernel variant sweep (development tool): C2 decode variants
(BB_TUNE_C2 / BB_TUNE_TILE_U), Mark 4 decode variants (BB_TUNE_M4) and the
bandwidth probes, each timed as a burst (best / median of 10 launches) and
sustained (back-to-back launches for ~1.5 s).  Variants must agree bit for
bit with variant 0.

    python tools/sweep_variants.py [chunk_gib] [sustain_s]
"""
import os
import sys
import time

import torch

sys.path.insert(0, '.')
from baseband_b200 import kernels, levels, synthetic  # noqa: E402

DEV = torch.device('cuda:0')


def burst(fn, n=10, warm=3):
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


def sustained(fn, seconds, ms_guess):
    n = max(10, int(seconds * 1e3 / ms_guess))
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def report(name, nbytes, fn, seconds):
    best, med = burst(fn)
    sus = sustained(fn, seconds, med)
    print('%-46s burst %7.1f GB/s (median %7.1f)  sustained %7.1f GB/s'
          % (name, nbytes / best / 1e6, nbytes / med / 1e6,
             nbytes / sus / 1e6), flush=True)
    time.sleep(1.0)            # let the power state settle between cases
    return nbytes / sus / 1e6


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
    quick = len(sys.argv) > 3 and sys.argv[3] == 'knock' 
    print(torch.cuda.get_device_name(0), 'chunk %.2f GiB packed' % gib)
    n = int(16 * gib * 2**30)
    a = torch.empty(n, dtype=torch.uint8, device=DEV)
    b = torch.empty(n, dtype=torch.uint8, device=DEV)
    report('torch copy_ (read+write)', 2 * n, lambda: b.copy_(a), seconds)
    report('bb_probe_copy (read+write)', 2 * n,
           lambda: kernels.probe_copy(b, a), seconds)
    report('torch fill_', n, lambda: b.fill_(1), seconds)
    report('bb_probe_fill contiguous', n, lambda: kernels.probe_fill(b, 0),
           seconds)
    report('bb_probe_fill rowgroup pattern', n,
           lambda: kernels.probe_fill(b, 1), seconds)
    # 1:16 expansion with ideal access patterns: what a 2 bit -> float32
    # stream can reach at all (bytes = read + written)
    for pat, what in ((0, 'contiguous input'), (1, '16 input streams')):
        report('bb_probe_expand 1:16, %s' % what, n + n // 16,
               lambda pat=pat: kernels.probe_expand(b, a, pat), seconds)
    report('bb_probe_expand 1:4 (8 bit -> float32), contiguous input',
           n + n // 4, lambda: kernels.probe_expand(b, a, 3), seconds)
    report('bb_probe_expand 1:16, input wrapped to 1 MiB (L2 hits)',
           n + n // 16, lambda: kernels.probe_expand(b, a, 2), seconds)
    del a, b
    torch.cuda.empty_cache()

    # ---- C2: VDIF 2 bit, 16 threads, 8032-byte frames
    for name, nthread, nelem, cplx in (('C2 16 thr real', 16, 1, False),
                                       ('8 thr complex', 8, 2, True)):
        payload, frame = 8000, 8032
        nset = int(gib * 2**30) // (nthread * frame)
        nunit = nset * nthread
        raw = torch.randint(0, 256, (nunit * frame,), dtype=torch.uint8,
                            device=DEV)
        perm = torch.arange(nunit, device=DEV).view(nset, nthread)
        perm = perm[:, torch.randperm(nthread, device=DEV)].reshape(-1)
        off = perm.to(torch.int64) * frame + 32
        spf = payload * 8 // (2 * nelem)
        out = torch.empty((nset * spf, nthread, nelem), dtype=torch.float32,
                          device=DEV)
        lv = levels.offset_binary(2)
        nbytes = nunit * frame + out.numel() * 4
        ref = None
        for c2 in ((0, 1) if quick else (0, 1, 2, 3, 4, 5)):
            for tu in ((1,) if c2 < 2 else (1, 2)):
                os.environ['BB_TUNE_C2'] = str(c2)
                os.environ['BB_TUNE_TILE_U'] = str(tu)
                out.zero_()

                def fn():
                    kernels.decode_bitfield(raw, off, nset, nthread, payload,
                                            2, nelem, cplx,
                                            kernels.CODEC_LEVELS, lv, out=out)
                fn()
                torch.cuda.synchronize()
                if ref is None:
                    ref = out.clone()
                    same = True
                else:
                    same = torch.equal(ref, out)
                report('%s DEC C2=%d U=%d %s' % (name, c2, tu,
                                                  'ok' if same else 'MISMATCH'),
                       nbytes, fn, seconds)
        os.environ.pop('BB_TUNE_C2', None)
        # where does the gap to the pure-write rate come from?  Same launch,
        # but every set reads the frames of the first `alias` sets, so the
        # packed input stays in L2 and DRAM sees writes only.
        for alias in (64, 1024):
            off_a = off.view(nset, nthread)[torch.arange(nset, device=DEV)
                                            % alias].reshape(-1).contiguous()

            def fn_alias():
                kernels.decode_bitfield(raw, off_a, nset, nthread, payload,
                                        2, nelem, cplx, kernels.CODEC_LEVELS,
                                        lv, out=out)
            report('%s DEC input aliased to %d sets (L2 hits)'
                   % (name, alias), nbytes, fn_alias, seconds)
        # knock-outs: which part of the kernel costs the gap to the fill rate
        for c2 in (0, 1):
            for knock in (1, 2):
                os.environ['BB_TUNE_C2'] = str(c2)
                os.environ['BB_TUNE_KNOCK'] = str(knock)

                def fn_k():
                    kernels.decode_bitfield(raw, off, nset, nthread, payload,
                                            2, nelem, cplx,
                                            kernels.CODEC_LEVELS, lv, out=out)
                report('%s DEC C2=%d knock-out %d (%s)' % (
                    name, c2, knock, 'no payload loads' if knock == 1
                    else 'no loads at all'), nbytes, fn_k, seconds)
        os.environ.pop('BB_TUNE_C2', None)
        os.environ['BB_TUNE_KNOCK'] = '0'
        del raw, out, ref
        torch.cuda.empty_cache()

    # ---- Mark 4: 64 tracks (C3) and 32 tracks, fan-out 4
    lv4 = levels.sign_magnitude()
    if quick:
        return
    for name, nchan in (('C3 mark4 64 trk', 8), ('mark4 32 trk', 4)):
        fbytes = nchan * 20000
        nframe = max(1, int(gib * 2**30) // fbytes)
        raw = torch.randint(0, 256, (nframe * fbytes,), dtype=torch.uint8,
                            device=DEV)
        uo = (torch.arange(nframe, dtype=torch.int64, device=DEV) * fbytes
              + nchan * 160)
        out = torch.empty((nframe * 80000, nchan), dtype=torch.float32,
                          device=DEV)
        nbytes = raw.numel() + out.numel() * 4
        ref = None
        for m4 in (0, 1, 2):
            os.environ['BB_TUNE_M4'] = str(m4)
            out.zero_()

            def fn():
                kernels.mark4_decode(raw, uo, nframe, nchan, 4, False, lv4,
                                     out=out)
            fn()
            torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
                same = True
            else:
                same = torch.equal(ref, out)
            report('%s DEC M4=%d %s' % (name, m4, 'ok' if same
                                         else 'MISMATCH'), nbytes, fn, seconds)
        os.environ['BB_TUNE_M4'] = '0'
        del raw, out, ref
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
