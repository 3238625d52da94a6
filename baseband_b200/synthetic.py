"""Synthetic streams of the shapes BASELINE.json names.

Headers are built on the host with numpy (a few bytes per frame); payload
bytes are i.i.d. uniform and can be drawn directly on the GPU for the large
device-resident benchmark chunks.
"""
import numpy as np

VDIF_SEED = 20240601
MARK4_SEED = 20240602
GUPPI_SEED = 20240603
MARK5B_SEED = 20240604


def vdif_headers(nset, nthread, frame_nbytes, bps=2, nchan=1,
                 complex_data=False, first_set=0, frames_per_second=2000,
                 thread_order=None, seconds0=100, ref_epoch=48, edv=0,
                 station=0x4142, invalid=None):
    """(nset*nthread, 32) uint8: EDV 0 headers (baseband/vdif/header.py:
    529-559 field layout), threads in a fixed shuffled order per set."""
    if thread_order is None:
        thread_order = np.random.default_rng(VDIF_SEED).permutation(nthread)
    thread_order = np.asarray(thread_order)
    sets = first_set + np.arange(nset, dtype=np.int64)
    seconds = seconds0 + sets // frames_per_second
    frame_nr = sets % frames_per_second
    w = np.zeros((nset, nthread, 8), np.uint32)
    w[..., 0] = (seconds & 0x3fffffff)[:, None]
    w[..., 1] = ((ref_epoch & 0x3f) << 24) | (frame_nr & 0xffffff)[:, None]
    lg2 = int(np.log2(nchan))
    w[..., 2] = (1 << 29) | (lg2 << 24) | (frame_nbytes // 8)
    w[..., 3] = ((int(complex_data) << 31) | ((bps - 1) << 26)
                 | (thread_order[None, :].astype(np.uint32) << 16) | station)
    w[..., 4] = edv << 24
    w = w.reshape(nset * nthread, 8)
    if invalid is not None:
        w[np.asarray(invalid), 0] |= np.uint32(1 << 31)
    return w.view(np.uint8).reshape(nset * nthread, 32)


def vdif_stream(nset, nthread=16, payload_nbytes=8000, bps=2, nchan=1,
                complex_data=False, seed=VDIF_SEED, **kwargs):
    """Whole synthetic VDIF stream as a host uint8 array."""
    frame_nbytes = payload_nbytes + 32
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (nset * nthread, frame_nbytes),
                          dtype=np.uint8)
    frames[:, :32] = vdif_headers(nset, nthread, frame_nbytes, bps, nchan,
                                  complex_data, **kwargs)
    return frames.reshape(-1)


def vdif_stream_device(nset, nthread, payload_nbytes, device, seed=VDIF_SEED,
                       **kwargs):
    """Same layout, payload drawn on the GPU (for multi-GiB chunks)."""
    import torch
    frame_nbytes = payload_nbytes + 32
    g = torch.Generator(device=device).manual_seed(seed)
    frames = torch.randint(0, 256, (nset * nthread, frame_nbytes),
                           dtype=torch.uint8, device=device, generator=g)
    hdr = vdif_headers(nset, nthread, frame_nbytes, **kwargs)
    frames[:, :32] = torch.from_numpy(hdr).to(device)
    return frames.reshape(-1)


def bcd(value, ndigit):
    out = 0
    for d in range(ndigit):
        out |= ((value // 10 ** d) % 10) << (4 * d)
    return out


def bcd_array(values, ndigit):
    """Vectorised `bcd`."""
    values = np.asarray(values, np.int64)
    out = np.zeros(values.shape, np.int64)
    for d in range(ndigit):
        out |= ((values // 10 ** d) % 10) << (4 * d)
    return out


def mark5b_headers(nframe, frames_per_second=6400, jday=123, seconds0=3600):
    """(nframe, 4) uint32 Mark 5B header words (mark5b/header.py:60-75):
    sync, frame_nr, BCD jday/seconds, BCD fraction + CRC-16."""
    from .base.utils import crc_array
    idx = np.arange(nframe, dtype=np.int64)
    sec = seconds0 + idx // frames_per_second
    fnr = idx % frames_per_second
    w = np.empty((nframe, 4), np.uint32)
    w[:, 0] = 0xABADDEED
    w[:, 1] = fnr.astype(np.uint32)
    w[:, 2] = ((bcd(jday, 3) << 20) | bcd_array(sec, 5)).astype(np.uint32)
    frac = fnr * 10000 // frames_per_second
    w[:, 3] = bcd_array(frac, 4).astype(np.uint32) << 16
    w[:, 3] |= crc_array((w[:, 2].astype(np.uint64) << np.uint64(16))
                         | (w[:, 3] >> np.uint32(16)).astype(np.uint64),
                         48, 0x18005).astype(np.uint32)
    return w


def mark5b_stream(nframe, seed=MARK5B_SEED, invalid_fraction=0.01,
                  frames_per_second=6400, jday=123, seconds0=3600):
    """Mark 5B frames (16-byte header + 10000-byte payload); a Bernoulli
    fraction of frames carries the fill pattern (mark5b/frame.py:62)."""
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (nframe, 10016), dtype=np.uint8)
    bad = rng.random(nframe) < invalid_fraction
    w = frames.view('<u4').reshape(nframe, 2504)
    w[bad, 4:] = 0x11223344
    w[:, :4] = mark5b_headers(nframe, frames_per_second, jday, seconds0)
    return frames.reshape(-1), ~bad


def mark5b_stream_device(nframe, device, seed=MARK5B_SEED,
                         invalid_fraction=0.01, **kwargs):
    """Same layout with the payload drawn on the GPU (multi-GiB chunks);
    returns (uint8 CUDA tensor, host bool array of valid frames)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    frames = torch.randint(0, 256, (nframe, 10016), dtype=torch.uint8,
                           device=device, generator=g)
    bad = np.random.default_rng(seed).random(nframe) < invalid_fraction
    w = frames.view(torch.int32)
    w[torch.from_numpy(bad).to(device), 4:] = 0x11223344
    hdr = mark5b_headers(nframe, **kwargs).view(np.int32)
    w[:, :4] = torch.from_numpy(hdr).to(device)
    return frames.reshape(-1), ~bad


def mark4_stream(nframe, seed=MARK4_SEED):
    """64-track fan-out 4 frames: all-zero time fields, sync words set, no
    error flags; payload uniform random."""
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 2**63, (nframe, 20000), dtype=np.int64).view(
        np.uint64)
    frames[:, :160] = 0
    frames[:, 64:96] = np.uint64(0xffffffffffffffff)
    return frames.view(np.uint8).reshape(-1)


def mark4_stream_device(nframe, device, seed=MARK4_SEED):
    """`mark4_stream` with the payload drawn on the GPU."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    frames = torch.randint(0, 256, (nframe, 160000), dtype=torch.uint8,
                           device=device, generator=g)
    w = frames.view(torch.int64)
    w[:, :160] = 0
    w[:, 64:96] = -1
    return frames.reshape(-1)


def guppi_header(nchan, npol, blocsize, overlap, pktidx, tbin=1e-6):
    cards = [('BACKEND', "'GUPPI   '"), ('TELESCOP', "'SYNTH   '"),
             ('OBSNCHAN', nchan), ('NPOL', npol * 2), ('NBITS', 8),
             ('PKTFMT', "'1SFA    '"), ('PKTSIZE', 8192),
             ('BLOCSIZE', blocsize), ('OVERLAP', overlap),
             ('PKTIDX', pktidx), ('TBIN', tbin), ('OBSBW', 100.0),
             ('OBSFREQ', 1400.0), ('STT_IMJD', 58000), ('STT_SMJD', 0),
             ('STT_OFFS', 0)]
    text = ''.join('{:<8s}= {:>20s}'.format(k, str(v)).ljust(80)
                   for k, v in cards) + 'END'.ljust(80)
    return np.frombuffer(text.encode('ascii'), np.uint8)


def guppi_stream(nframe, nchan=512, npol=2, samples_per_frame=4096,
                 overlap=64, seed=GUPPI_SEED):
    """Channels-first complex int8 GUPPI frames whose overlap region repeats
    the head of the next frame."""
    rng = np.random.default_rng(seed)
    stride = samples_per_frame - overlap
    total = stride * nframe + overlap
    data = rng.integers(-128, 128, (nchan, total, npol, 2), dtype=np.int8)
    blocsize = nchan * samples_per_frame * npol * 2
    pieces = []
    for f in range(nframe):
        pieces.append(guppi_header(nchan, npol, blocsize, overlap,
                                   f * (blocsize - overlap * nchan * npol * 2)
                                   // 8192))
        block = data[:, f * stride:f * stride + samples_per_frame]
        pieces.append(np.ascontiguousarray(block).view(np.uint8).reshape(-1))
    return np.concatenate(pieces), data
