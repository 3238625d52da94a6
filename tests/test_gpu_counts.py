"""bb_state_counts (csrc/bb_counts.cu) against its numpy restatement
(tests/cpu_backend.py:_state_counts), through the C ABI: every template
instance, shuffled / invalid / unaligned units, bins that cut the call at
arbitrary sets, accumulation over calls."""
import numpy as np
import pytest
import torch

import cpu_backend
from baseband_b200 import kernels

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

# up to more word classes than a CTA has threads (several passes): 4 bit x
# 2048 elements (1024 complex channels) is 256 words per sample
CASES = [(bps, nelem) for bps in (1, 2, 4)
         for nelem in (1, 2, 4, 8, 16, 32, 64, 256, 2048, 8192)]


@pytest.mark.parametrize('bps,nelem', CASES)
def test_state_counts_fuzz(bps, nelem):
    rng = np.random.default_rng(1000 * bps + nelem)
    for trial in range(4):
        nthread = int(rng.choice([1, 2, 3, 16]))
        words_per_sample = max(1, nelem * bps // 32)
        nword = words_per_sample * int(rng.integers(1, 700))
        if trial == 3 and words_per_sample <= 4:
            # > 255 words per thread: the vertical counters flush in between
            nword = words_per_sample * int(rng.integers(40000, 50000))
        payload = nword * 4
        nset = int(rng.integers(1, 9))
        hdr = int(rng.choice([0, 4, 16, 32]))        # 4: unaligned for uint4
        frame = payload + hdr
        nunit = nset * nthread
        raw = rng.integers(0, 256, nunit * frame + 64, dtype=np.uint8)
        uo = (rng.permutation(nunit).astype(np.int64) * frame + hdr)
        uo[rng.random(nunit) < 0.15] = -1
        sets_per_bin = int(rng.integers(1, nset + 2))
        origin = int(rng.integers(0, 5))
        nbin = (origin + nset - 1) // sets_per_bin + 1
        shape = (nbin, nthread, nelem, 1 << bps)
        want = torch.zeros(shape, dtype=torch.int64)
        cpu_backend._state_counts(torch.from_numpy(raw), torch.from_numpy(uo),
                                  nset, nthread, payload, bps, nelem, want,
                                  origin, sets_per_bin)
        got = torch.zeros(shape, dtype=torch.int64, device=DEV)
        d_raw, d_uo = torch.from_numpy(raw).to(DEV), torch.from_numpy(uo).to(DEV)
        kernels.state_counts(d_raw, d_uo, nset, nthread, payload, bps, nelem,
                             got, origin, sets_per_bin)
        assert torch.equal(got.cpu(), want), (bps, nelem, trial)
        # accumulates
        kernels.state_counts(d_raw, d_uo, nset, nthread, payload, bps, nelem,
                             got, origin, sets_per_bin)
        assert torch.equal(got.cpu(), 2 * want)


def test_state_counts_large_and_errors():
    # C2 geometry, enough sets that the split over CTAs matters
    nset, nthread, payload = 300, 16, 8000
    g = torch.Generator(device=DEV).manual_seed(3)
    raw = torch.randint(0, 256, (nset * nthread * 8032,), dtype=torch.uint8,
                        device=DEV, generator=g)
    uo = torch.arange(nset * nthread, dtype=torch.int64, device=DEV) * 8032 + 32
    got = torch.zeros((3, nthread, 1, 4), dtype=torch.int64, device=DEV)
    kernels.state_counts(raw, uo, nset, nthread, payload, 2, 1, got, 0, 128)
    codes = raw.view(nset, nthread, 8032)[:, :, 32:]
    for c in range(4):
        per_set = sum(((codes >> (2 * k)) & 3) == c for k in range(4)).sum(-1)
        want = torch.stack([per_set[b * 128:(b + 1) * 128].sum(0)
                            for b in range(3)])
        assert torch.equal(got[:, :, 0, c], want)
    with pytest.raises(KeyError):                        # 8 bit: unsupported
        kernels.state_counts(raw, uo, 1, 1, 8000, 8, 1, torch.zeros(
            (1, 1, 1, 256), dtype=torch.int64, device=DEV))
    with pytest.raises(ValueError):                      # bins too few
        kernels.state_counts(raw, uo, nset, nthread, payload, 2, 1,
                             got[:1], 0, 128)
    with pytest.raises(ValueError):                      # nelem not 2**k
        kernels.state_counts(raw, uo, 1, 1, 8000 - 8000 % 12, 2, 3,
                             torch.zeros((1, 1, 3, 4), dtype=torch.int64,
                                         device=DEV))


@pytest.mark.parametrize('nelem', (1, 2, 4, 8, 16, 64, 256, 1024))
def test_int8_moments_fuzz(nelem):
    """bb_int8_moments against its numpy restatement: every element class,
    shuffled / invalid units, bins cutting the call, accumulation, values at
    the extremes (all -128; the overflow point of the 32-bit partial sums is
    covered by test_gpu_large.test_int8_moments_one_huge_unit)."""
    rng = np.random.default_rng(50 + nelem)
    for trial in range(4):
        nthread = int(rng.choice([1, 2, 5]))
        nsamp = int(rng.integers(1, 3000)) * max(1, 4 // nelem)
        payload = nsamp * nelem
        payload += (-payload) % 4
        if payload % nelem:
            payload = -(-payload // (4 * nelem)) * 4 * nelem
        nset = int(rng.integers(1, 7))
        hdr = int(rng.choice([0, 4, 16]))
        frame = payload + hdr
        nunit = nset * nthread
        raw = rng.integers(0, 256, nunit * frame + 16, dtype=np.uint8)
        if trial == 0:
            raw[:] = 0x80                                # all -128
        uo = rng.permutation(nunit).astype(np.int64) * frame + hdr
        uo[rng.random(nunit) < 0.2] = -1
        per_bin = int(rng.integers(1, nset + 2))
        origin = int(rng.integers(0, 4))
        nbin = (origin + nset - 1) // per_bin + 1
        shape = (nbin, nthread, nelem, 3)
        want = torch.zeros(shape, dtype=torch.int64)
        cpu_backend._int8_moments(torch.from_numpy(raw), torch.from_numpy(uo),
                                  nset, nthread, payload, nelem, want, origin,
                                  per_bin)
        got = torch.zeros(shape, dtype=torch.int64, device=DEV)
        d_raw, d_uo = torch.from_numpy(raw).to(DEV), torch.from_numpy(uo).to(DEV)
        for _ in range(2):
            kernels.int8_moments(d_raw, d_uo, nset, nthread, payload, nelem,
                                 got, origin, per_bin)
        assert torch.equal(got.cpu(), 2 * want), (nelem, trial)
    with pytest.raises(ValueError):
        kernels.int8_moments(d_raw, d_uo, 1, 1, 12, 3, torch.zeros(
            (1, 1, 3, 3), dtype=torch.int64, device=DEV))


M4_LAYOUTS = [(8, 4, False), (4, 4, False), (8, 2, False), (2, 4, False),
              (16, 2, True)]


@pytest.mark.parametrize('nchan,fanout,ft', M4_LAYOUTS)
def test_mark4_state_counts(nchan, fanout, ft):
    """bb_mark4_state_counts (track words -> two-bit code words -> vertical
    counters) against the decoder: decoding with the level table (0, 1, 2, 3)
    gives the index 2 * sign + magnitude of every sample (that decode is
    itself parity-tested against the oracle and the reference's golden
    vectors).  Random track words, shuffled / invalid frames, several bins,
    accumulation."""
    rng = np.random.default_rng(nchan * 10 + fanout)
    wordbytes = nchan * 2 * fanout // 8
    frame = 20000 * wordbytes
    for trial in range(3):
        nframe = int(rng.integers(1, 12))
        raw = rng.integers(0, 256, nframe * frame, dtype=np.uint8)
        uo = rng.permutation(nframe).astype(np.int64) * frame \
            + 160 * wordbytes
        if nframe > 2:
            uo[rng.random(nframe) < 0.25] = -1
        per_bin = int(rng.integers(1, nframe + 2))
        origin = int(rng.integers(0, 4))
        nbin = (origin + nframe - 1) // per_bin + 1
        d_raw, d_uo = torch.from_numpy(raw).to(DEV), torch.from_numpy(uo).to(DEV)
        codes = kernels.mark4_decode(d_raw, d_uo, nframe, nchan, fanout, ft,
                                     levels=np.arange(4, dtype=np.float32),
                                     fill_value=-1.0)
        codes = codes.cpu().numpy().reshape(nframe, -1, nchan)
        want = np.zeros((nbin, nchan, 4), np.int64)
        for i in range(nframe):
            for c in range(nchan):
                col = codes[i, :, c]
                want[(origin + i) // per_bin, c] += np.bincount(
                    col[col >= 0].astype(np.int64), minlength=4)
        got = torch.zeros((nbin, nchan, 4), dtype=torch.int64, device=DEV)
        for _ in range(2):
            kernels.mark4_state_counts(d_raw, d_uo, nframe, nchan, fanout, ft,
                                       got, origin, per_bin)
        assert np.array_equal(got.cpu().numpy(), 2 * want), (trial, nframe)
        valid = int((uo >= 0).sum())
        assert got.sum().item() == 2 * valid * (20000 - 160) * fanout * nchan
    with pytest.raises(KeyError):
        kernels.mark4_state_counts(d_raw, d_uo, 1, 3, 4, False, torch.zeros(
            (1, 3, 4), dtype=torch.int64, device=DEV))
