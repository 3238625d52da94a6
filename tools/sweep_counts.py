"""Sweep of the state-count kernel's grid depth (BB_TUNE_COUNT_DEPTH = CTAs
per SM aimed at) on the C2 geometry, resident input."""
import os
import sys

import torch

sys.path.insert(0, '.')
from baseband_b200 import kernels  # noqa: E402

DEV = 'cuda:0'
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
payload, frame, nthread = 8000, 8032, 16
nset = int(gib * 2**30) // frame // nthread
raw = torch.randint(0, 256, (nset * nthread * frame,), dtype=torch.uint8,
                    device=DEV)
uo = torch.arange(nset * nthread, dtype=torch.int64, device=DEV) * frame + 32
for _ in range(3):
    ts = []
    for i in range(10):
        a, b = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        a.record()
        kernels.probe_read(raw[:raw.numel() // 16 * 16])
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[2:])
    print('pure read probe: %.1f GB/s median, %.1f best' % (
        raw.numel() / ts[len(ts) // 2] / 1e6, raw.numel() / ts[0] / 1e6))
for spb in (200, 2000):
    acc = kernels.zeros((-(-nset // spb), nthread, 1, 4), torch.int64,
                        torch.device(DEV))
    for depth in (8, 16, 20, 24, 28, 32, 40):
        os.environ['BB_TUNE_COUNT_DEPTH'] = str(depth)
        ts = []
        for i in range(12):
            a, b = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            a.record()
            kernels.state_counts(raw, uo, nset, nthread, payload, 2, 1, acc,
                                 sets_per_bin=spb)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[2:])
        print('sets/bin %5d depth %3d: %.1f GB/s median, %.1f best' % (
            spb, depth, nset * nthread * payload / ts[len(ts) // 2] / 1e6,
            nset * nthread * payload / ts[0] / 1e6))

# the 8-bit moments kernel on the C4 GUPPI geometry (512 channel rows per
# frame, 2 pol complex int8 = 4 elements)
del raw, uo
nchan, rowbytes = 512, 256 << 10
nfr = max(1, int(gib * 2**30) // (nchan * rowbytes))
raw = torch.randint(0, 256, (nfr * nchan * rowbytes,), dtype=torch.uint8,
                    device=DEV)
uo = torch.arange(nfr * nchan, dtype=torch.int64, device=DEV) * rowbytes
mom = kernels.zeros((nfr, nchan, 4, 3), torch.int64, torch.device(DEV))
for depth, mu in ((24, 2), (24, 4), (24, 8), (8, 8), (48, 4), (48, 8)):
    os.environ['BB_TUNE_COUNT_DEPTH'] = str(depth)
    os.environ['BB_TUNE_MOM_U'] = str(mu)
    ts = []
    for i in range(10):
        a, b = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        a.record()
        kernels.int8_moments(raw, uo, nfr, nchan, rowbytes, 4, mom,
                             sets_per_bin=1)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[2:])
    print('int8 moments depth %3d U %d: %.1f GB/s median, %.1f best' % (
        depth, mu, raw.numel() / ts[len(ts) // 2] / 1e6, raw.numel() / ts[0] / 1e6))

# the other paths: vertical counters (1 / 2 bit, many elements), shared-memory
# histogram (4 bit)
del raw, uo
os.environ['BB_TUNE_COUNT_DEPTH'] = '24'
payload, frame = 8192, 8224
for bps, nthread, nelem in ((4, 4, 1), (4, 1, 1024), (2, 1, 64), (2, 4, 8),
                            (2, 1, 16), (1, 1, 64), (1, 8, 1)):
    nset = int(gib * 2**30) // frame // nthread
    raw = torch.randint(0, 256, (nset * nthread * frame,), dtype=torch.uint8,
                        device=DEV)
    uo = torch.arange(nset * nthread, dtype=torch.int64, device=DEV) * frame + 32
    acc = kernels.zeros((-(-nset // 500), nthread, nelem, 1 << bps),
                        torch.int64, torch.device(DEV))
    for vd in ((8, 16, 24, 32, 48) if bps == 4 or nelem > 8 // bps else (16,)):
        os.environ['BB_TUNE_VERT_DEPTH'] = str(vd)
        ts = []
        for i in range(8):
            a, b = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            a.record()
            kernels.state_counts(raw, uo, nset, nthread, payload, bps, nelem,
                                 acc, sets_per_bin=500)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[2:])
        print('counts %d bit, %d threads x %d elements, depth %d: %.1f GB/s '
              'median' % (bps, nthread, nelem, vd,
                          nset * nthread * payload / ts[len(ts) // 2] / 1e6))
    del raw, uo, acc

# Mark 4 track words (C3: 64 tracks, fan-out 4; and the 16-track layout)
from baseband_b200 import synthetic  # noqa: E402
for nchan, fanout, ft in ((8, 4, False), (16, 2, True), (2, 4, False)):
    wordbytes = nchan * 2 * fanout // 8
    frame = 20000 * wordbytes
    nframe = int(gib * 2**30) // frame
    raw = torch.randint(0, 256, (nframe * frame,), dtype=torch.uint8,
                        device=DEV)
    uo = torch.arange(nframe, dtype=torch.int64, device=DEV) * frame \
        + 160 * wordbytes
    acc = kernels.zeros((-(-nframe // 50), nchan, 4), torch.int64,
                        torch.device(DEV))
    os.environ.pop('BB_TUNE_VERT_DEPTH', None)
    ts = []
    for i in range(8):
        a, b = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        a.record()
        kernels.mark4_state_counts(raw, uo, nframe, nchan, fanout, ft, acc,
                                   sets_per_bin=50)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[2:])
    print('mark4 counts %d ch fan-out %d%s: %.1f GB/s median (frame bytes)' % (
        nchan, fanout, ' ft' if ft else '',
        nframe * frame / ts[len(ts) // 2] / 1e6))
    del raw, uo, acc
