"""bb_locate_frames / bb_vdif_index / bb_mark5b_index against the oracle
(oracle/locate.py: the reference's `locate_frames`, base/base.py:181-335,
pinned to the reference's own known answers in test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from baseband_b200 import kernels, synthetic
from conftest import sample_bytes
from oracle import locate

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _gpu_locate(data, pattern, mask, frame_nbytes, offset=0, check=1,
                at_eof=True, own_stop=None, base=0):
    loc, cnt = kernels.locate_frames(
        torch.from_numpy(np.ascontiguousarray(data)).to(DEV),
        np.asarray(pattern, '<u4'),
        None if mask is None else np.asarray(mask, '<u4'), frame_nbytes,
        offset, own_stop=own_stop, check=check, at_eof=at_eof, base=base)
    n = int(cnt.item())
    assert n <= loc.numel()
    return sorted(loc[:n].cpu().tolist())


def test_locate_frames_samples():
    data = sample_bytes('sample.vdif')
    words = data[:32].view('<u4')
    inv = [0x40000000, 0, 0xffffffff, 0xfc00ffff, 0xffffffff, 0xffffffff,
           0, 0]
    every = [x * 5032 for x in range(16)]
    # whole file = the oracle searching forward from 0 over the file size
    want = locate.locate_frames(data, 0, words, mask=inv, frame_nbytes=5032,
                                maximum=data.size)
    assert want == every
    assert _gpu_locate(data, words, inv, 5032) == every
    # sync word only, at its offset in the header
    assert _gpu_locate(data, [0xACABFEED], None, 5032, offset=20) == every
    # without the check and without a frame size
    assert _gpu_locate(data, [0xACABFEED], None, 0, offset=20, check=0) \
        == locate.locate_frames(data, 0, 0xACABFEED, offset=20,
                                maximum=data.size)
    m5 = sample_bytes('sample.m5b')
    bad = np.concatenate([m5[:10040], m5[20000:]])
    for blob in (m5, bad, m5[:10018], m5[:30]):
        for check in (1, 0):
            want = locate.locate_frames(blob, 0, 0xABADDEED,
                                        frame_nbytes=10016,
                                        check=check or None,
                                        maximum=blob.size)
            got = _gpu_locate(blob, [0xABADDEED], None, 10016, check=check)
            if check == 0:
                # no check: the oracle still wants the frame to fit
                assert got == want
            else:
                assert got == want, (blob.size, check)


def test_locate_frames_fuzz_and_chunks():
    rng = np.random.default_rng(12)
    for trial in range(30):
        frame = int(rng.integers(40, 400))
        npat = int(rng.choice([1, 2, 3, 4, 5, 8, 13, 32]))
        offset = int(rng.integers(0, frame - npat))
        pat = rng.integers(0, 256, npat, dtype=np.uint8)
        mask = rng.integers(0, 256, npat, dtype=np.uint8)
        mask[rng.integers(0, npat)] |= 0x81
        nfr = int(rng.integers(2, 30))
        data = rng.integers(0, 4, nfr * frame + int(rng.integers(0, 50)),
                            dtype=np.uint8)              # many near-matches
        for f in range(nfr):
            if rng.random() < 0.8:
                at = f * frame + offset
                data[at:at + npat] = (data[at:at + npat] & ~mask) | (pat & mask)
        want = locate.locate_frames(data, 0, pat, mask=mask,
                                    frame_nbytes=frame, offset=offset,
                                    maximum=data.size)
        loc, cnt = kernels.locate_frames(
            torch.from_numpy(data).to(DEV), pat, mask, frame, offset)
        got = sorted(loc[:int(cnt.item())].cpu().tolist())
        assert got == want, trial
        # the same file in overlapping chunks, as the index builder does
        step, overlap = 3 * frame, frame + offset + npat
        pos, found = 0, []
        while pos < data.size:
            n = min(step + overlap, data.size - pos)
            eof = pos + n >= data.size
            loc, cnt = kernels.locate_frames(
                torch.from_numpy(data[pos:pos + n].copy()).to(DEV), pat, mask,
                frame, offset, own_stop=n if eof else step, at_eof=eof,
                base=pos)
            found += loc[:int(cnt.item())].cpu().tolist()
            if eof:
                break
            pos += step
        assert sorted(found) == want, ('chunked', trial)


def test_vdif_and_mark5b_index_tables():
    nset, nthread = 7, 4
    raw = synthetic.vdif_stream(nset, nthread, 1000, seed=3, invalid=[9],
                                frames_per_second=5)
    frames = raw.reshape(-1, 1032)
    order = np.array([0, 1, 2, 3, 7, 6, 5, 4, 8, 9, 9, 10, 11, 16, 17, 18,
                      19, 12, 13, 14, 27, 26, 25, 24])       # dup, swap, loss
    blob = frames[order].reshape(-1)
    words = blob[:32].view('<u4')

    def index_of(loc):
        w = blob[loc:loc + 16].view('<u4')
        return ((int(w[0]) & 0x3fffffff) - 100) * 5 + (int(w[1]) & 0xffffff), \
            (int(w[3]) >> 16) & 0x3ff

    locs = [i * 1032 for i in range(len(order))]
    want = locate.frame_table(blob, locs, index_of, nthread,
                              lambda loc: blob[loc + 3] >> 7)
    d = torch.from_numpy(blob).to(DEV)
    loc, cnt = kernels.locate_frames(d, words[:4], [0x40000000, 0, 0xffffffff,
                                                    0xfc00ffff], 1032)
    slot = torch.from_numpy(np.arange(1024, dtype=np.int32)).to(DEV)
    table = kernels.index_table(20 * nthread, DEV)
    stats = kernels.zeros(4, torch.int32, DEV)
    kernels.vdif_index(d, 0, loc, cnt, slot, nthread, 100, 0, 5, 20, table,
                       stats)
    got = kernels.index_table_finish(table).cpu().numpy().reshape(20, nthread)
    st = stats.cpu().numpy()
    assert st[0] + 1 == want.shape[0] and st[1] == 0
    assert np.array_equal(got[:want.shape[0]], want)
    assert (got[want.shape[0]:] == -1).all()
    # Mark 5B: frames shuffled, one dropped, jday rollover handled by wrap
    m5, _ = synthetic.mark5b_stream(12, invalid_fraction=0., seed=5,
                                    frames_per_second=4, jday=999,
                                    seconds0=86399)
    fr = m5.reshape(12, 10016)
    blob = fr[[0, 1, 3, 2, 4, 6, 7, 8, 9, 11, 10]].reshape(-1)
    d = torch.from_numpy(blob).to(DEV)
    loc, cnt = kernels.locate_frames(d, [0xABADDEED], None, 10016)
    table = kernels.index_table(40, DEV)
    stats = kernels.zeros(4, torch.int32, DEV)
    kernels.mark5b_index(d, 0, loc, cnt, 999, 86399, 0, 4, 40, table, stats)
    got = kernels.index_table_finish(table).cpu().numpy()
    want = np.full(40, -1, np.int64)
    for pos, f in enumerate([0, 1, 3, 2, 4, 6, 7, 8, 9, 11, 10]):
        want[f] = pos * 10016
    assert np.array_equal(got, want)
    assert stats.cpu().numpy().tolist() == [11, 0, 0, 0]


def test_mark4_index_table():
    """bb_mark4_index == its numpy restatement (tests/cpu_backend.py, built
    on the oracle's stream2words) on a written stream with frames shuffled
    and dropped, across a year end; sync candidates next to the true frame
    starts are rejected by the CRC-12."""
    import io
    import cpu_backend
    import baseband_b200 as bb
    h0 = bb.mark4.Mark4Header.fromvalues(
        64, time='2019-12-31T23:59:59.990', bps=2, fanout=4, nsb=1,
        system_id=108)
    rng = np.random.default_rng(3)
    data = rng.choice(np.array([-3.316505, -1., 1., 3.316505], np.float32),
                      size=(9 * 80000, 8))
    buf = io.BytesIO()
    fw = bb.mark4.open(buf, 'ws', header0=h0, sample_rate=32e6)
    fw.write(data)
    frames = np.frombuffer(buf.getvalue(), np.uint8).reshape(9, 160000)
    order = [0, 2, 1, 3, 5, 6, 8, 7]
    blob = np.concatenate([rng.integers(0, 255, 13, dtype=np.uint8),
                           frames[order].reshape(-1)])
    pat = np.full(256, 0xff, np.uint8)
    d = torch.from_numpy(blob).to(DEV)
    loc, cnt = kernels.locate_frames(d, pat, pat, 160000, 512)
    want_loc, want_cnt = cpu_backend._locate_frames(
        torch.from_numpy(blob), pat, pat, 160000, 512)
    n = int(cnt.item())
    assert n == int(want_cnt.item()) and n >= len(order)
    assert sorted(loc[:n].cpu().tolist()) == sorted(want_loc[:n].tolist())
    args = (64, 0, 2019, 365, 365, 365, 345599960, 10, 40)
    table = kernels.index_table(40, DEV)
    stats = kernels.zeros(4, torch.int32, DEV)
    kernels.mark4_index(d, 0, loc, cnt, *args, table, stats)
    got = kernels.index_table_finish(table).cpu().numpy()
    ref_table = cpu_backend._index_table(40, None)
    ref_stats = torch.zeros(4, dtype=torch.int32)
    cpu_backend._mark4_index(torch.from_numpy(blob), 0, want_loc, want_cnt,
                             *args, ref_table, ref_stats)
    want = cpu_backend._index_table_finish(ref_table).numpy()
    assert np.array_equal(got, want)
    expect = np.full(40, -1, np.int64)
    for pos, f in enumerate(order):
        expect[f] = 13 + pos * 160000
    assert np.array_equal(got, expect)
    assert stats.cpu().tolist() == ref_stats.tolist()
