"""Public-API read throughput (pinned host source) for the stream readers."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import synthetic  # noqa: E402
from baseband_b200.base.memory import HostBuffer  # noqa: E402

dev = 'cuda:0'
nset = (512 << 20) // (16 * 8032)
src = HostBuffer(synthetic.vdif_stream(nset, 16, 8000, seed=1))
for chunk in (16 << 20, 64 << 20, 256 << 20):
    fh = bb.vdif.open(src, 'rs', sample_rate=64e6, device=dev,
                      chunk_nbytes=chunk)
    for rep in range(3):
        fh.seek(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data = fh.read()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print('device out chunk %3d MiB: %6.1f GB/s packed ingest, %6.1f Gsamp/s'
          % (chunk >> 20, src.size / dt / 1e9, data.numel() / dt / 1e9))
    del data
out = torch.empty((nset * 32000, 16), dtype=torch.float32, pin_memory=True)
for chunk in (16 << 20, 64 << 20):
    fh = bb.vdif.open(src, 'rs', sample_rate=64e6, chunk_nbytes=chunk)
    for rep in range(3):
        fh.seek(0)
        t0 = time.perf_counter()
        fh.read(out=out)
        dt = time.perf_counter() - t0
    print('host out   chunk %3d MiB: %6.1f GB/s D2H, %6.2f Gsamp/s'
          % (chunk >> 20, out.numel() * 4 / dt / 1e9, out.numel() / dt / 1e9))
small = bb.vdif.open(src, 'rs', sample_rate=64e6)
small.read(1)
t0 = time.perf_counter()
for i in range(200):
    small.seek(i * 1000)
    small.read(12)
print('12-sample read latency: %.0f us' % ((time.perf_counter() - t0) / 200 * 1e6))


def rate(fh, label, nbytes):
    best = 1e9
    for rep in range(3):
        fh.seek(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data = fh.read()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    comp = data.numel() * (2 if data.is_complex() else 1)
    print('%-34s %6.1f GB/s packed ingest, %6.1f Gsamp/s' % (
        label, nbytes / best / 1e9, comp / best / 1e9))


# Mark 5B (C5): 1 % invalid frames
raw, _ = synthetic.mark5b_stream(26000, invalid_fraction=0.01, seed=3)
fh = bb.mark5b.open(HostBuffer(raw), 'rs', nchan=16, sample_rate=16e6,
                    kday=56000, device=dev)
rate(fh, 'mark5b 2bit 16ch device out', raw.size)
# Mark 4 (C3): frames produced by our own writer from random levels
h0 = bb.mark4.Mark4Header.fromvalues(64, time='2014-06-16T07:38:12.475',
                                     bps=2, fanout=4, nsb=1)
nframe = 1600
lv = torch.tensor([-3.316505, -1., 1., 3.316505], device=dev)
vals = lv[torch.randint(0, 4, (nframe * 80000, 8), device=dev)]
sink = HostBuffer(nframe * 160000)
fw = bb.mark4.open(sink, 'ws', header0=h0, sample_rate=32e6, device=dev)
fw.write(vals)
fw._flush(final=False)
del vals
fh = bb.mark4.open(sink, 'rs', ntrack=64, decade=2010, device=dev)
rate(fh, 'mark4 64trk fanout4 device out', sink.size)
# GUPPI (C4, scaled): 512 chan x 2 pol, overlap
graw, _ = synthetic.guppi_stream(8, nchan=512, npol=2, samples_per_frame=8192,
                                 overlap=512, seed=5)
fh = bb.guppi.open(HostBuffer(graw), 'rs', device=dev)
rate(fh, 'guppi 512ch 2pol int8 device out', graw.size)
