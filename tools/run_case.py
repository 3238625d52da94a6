"""Run one kernel case a few times (for ncu captures).
usage: python tools/run_case.py {guppi|guppitf|mark5b|mark4enc|vdif48|vdif44|
vdif22c|vdif84|c2|mark4} [gib]   (REPS=n launches per kernel, default 3)"""
import os
import sys

import torch

sys.path.insert(0, '.')
from baseband_b200 import kernels, levels  # noqa: E402

DEV = 'cuda:0'
case = sys.argv[1]
gib = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
reps = int(os.environ.get('REPS', 3))
if case == 'guppi':
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
    for _ in range(reps):
        kernels.decode_int8_transposed(raw, off, nfr, nchan, spf * npol, 2,
                                       cb, ce, oc0, out)
elif case == 'guppitf':
    nchan, npol, nt = 512, 2, 65536
    fbytes = nchan * nt * npol * 2
    nfr = max(1, int(gib * 2**30) // fbytes)
    raw = torch.randint(0, 256, (nfr * fbytes,), dtype=torch.uint8, device=DEV)
    off = torch.arange(nfr, dtype=torch.int64, device=DEV) * fbytes
    tb = torch.zeros(nfr, dtype=torch.int64, device=DEV)
    te = torch.full((nfr,), nt, dtype=torch.int64, device=DEV)
    t0 = torch.arange(nfr, dtype=torch.int64, device=DEV) * nt
    out = torch.empty((nfr * nt * npol * nchan * 2,), dtype=torch.float32,
                      device=DEV)
    back = torch.zeros_like(raw)
    for _ in range(reps):
        kernels.decode_int8_timefirst(raw, off, nfr, nt, nchan, npol, 2, tb,
                                      te, t0, out)
    for _ in range(reps):
        kernels.encode_int8_timefirst(out, back, off, nfr, nt, nchan, npol, 2)
elif case in ('mark5b', 'c2', 'vdif48', 'vdif44', 'vdif22c', 'vdif84',
              'vdif18', 'gsb', 'vdif82c'):
    bps, nthread, nelem, payload, hdr = {
        'mark5b': (2, 1, 16, 10000, 16), 'c2': (2, 16, 1, 8000, 32),
        'vdif48': (2, 4, 8, 8000, 32), 'vdif44': (4, 4, 1, 8000, 32),
        'vdif22c': (2, 2, 2, 8000, 32), 'vdif84': (8, 4, 1, 8000, 32),
        'vdif18': (1, 8, 1, 8000, 32), 'gsb': (8, 2, 1024, 1 << 22, 0),
        'vdif82c': (8, 2, 2, 8000, 32)}[case]
    frame = payload + hdr
    nset = int(gib * 2**30) // frame // nthread
    nunit = nset * nthread
    raw = torch.randint(0, 256, (nunit * frame,), dtype=torch.uint8,
                        device=DEV)
    uo = torch.arange(nunit, dtype=torch.int64, device=DEV) * frame + hdr
    lv = levels.mark5b(2) if case == 'mark5b' else levels.offset_binary(bps)
    codec = kernels.CODEC_SINT if case == 'gsb' else kernels.CODEC_LEVELS
    out = None
    for _ in range(reps):
        out = kernels.decode_bitfield(raw, uo, nset, nthread, payload, bps,
                                      nelem, False, codec, lv, out=out)
    back = torch.zeros_like(raw)
    for _ in range(reps):
        kernels.encode_bitfield(out, back, uo, nset, nthread, payload, bps,
                                nelem, kernels.QUANT_MARK5B if case == 'mark5b'
                                else kernels.QUANT_SINT if case == 'gsb'
                                else kernels.QUANT_OFFSET_BINARY)
elif case == 'mark4':
    # C3: 64 tracks, fan-out 4, 8 channels: header scan + decode
    nframe = int(gib * 2**30) // 160000
    from baseband_b200 import synthetic
    raw = synthetic.mark4_stream_device(nframe, torch.device(DEV))
    out = None
    for _ in range(reps):
        _, off = kernels.mark4_scan(raw, nframe, 64)
        out = kernels.mark4_decode(raw, off, nframe, 8, 4, False,
                                   levels.sign_magnitude(), out=out)
elif case == 'counts':
    # the state-count consumer on the C2 geometry
    payload, frame, nthread = 8000, 8032, 16
    nset = int(gib * 2**30) // frame // nthread
    raw = torch.randint(0, 256, (nset * nthread * frame,), dtype=torch.uint8,
                        device=DEV)
    uo = torch.arange(nset * nthread, dtype=torch.int64, device=DEV) * frame + 32
    acc = kernels.zeros((-(-nset // 200), nthread, 1, 4), torch.int64,
                        torch.device(DEV))
    for _ in range(reps):
        kernels.state_counts(raw, uo, nset, nthread, payload, 2, 1, acc,
                             sets_per_bin=200)
elif case == 'mark4ftenc':
    # 64 tracks, fan-out 2, 16 channels, Fortaleza layout: encode
    nframe = int(gib * 2**30) // 160000
    raw = torch.randint(0, 256, (nframe * 160000,), dtype=torch.uint8,
                        device=DEV)
    off = torch.arange(nframe, dtype=torch.int64, device=DEV) * 160000 + 1280
    out = kernels.mark4_decode(raw, off, nframe, 16, 2, True,
                               levels.sign_magnitude())
    back = raw.clone()
    for _ in range(reps):
        kernels.mark4_encode(out, back, off, nframe, 16, 2, True)
elif case == 'mark4enc':
    nframe = int(gib * 2**30) // 160000
    raw = torch.randint(0, 256, (nframe * 160000,), dtype=torch.uint8,
                        device=DEV)
    off = torch.arange(nframe, dtype=torch.int64, device=DEV) * 160000 + 1280
    out = kernels.mark4_decode(raw, off, nframe, 8, 4, False,
                               levels.sign_magnitude())
    back = raw.clone()
    for _ in range(reps):
        kernels.mark4_encode(out, back, off, nframe, 8, 4, False)
torch.cuda.synchronize()
print('done', case)
