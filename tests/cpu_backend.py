"""TEST INFRASTRUCTURE: run the host layer on a box without a GPU.

``install(monkeypatch)`` points ``baseband_b200``'s narrow device seam
(``_lib.load``, a few helpers of ``kernels`` and ``device``) at the CPU
emulation of the kernel bodies (tests/emu, compiled from the very same
``csrc/*.cuh`` bodies with g++) and at numpy restatements of the header-scan
kernels built on the oracle.  This exists only so that the *host logic*
(frame-range planning, chunking, squeeze/subset, header generation, file
I/O) is covered by ``pytest -m "not gpu"``; the shipped package has no such
path and refuses CPU tensors.  The same tests run against the real CUDA
library under ``-m gpu`` (tests/test_gpu_streams.py).
"""
import contextlib
import ctypes

import numpy as np
import torch

import emu_build
from oracle import headers as oheaders


class _NoStreams:
    def __init__(self, dev):
        self.dev = dev

    def use(self, i):
        return contextlib.nullcontext()

    def wait(self, i, j):
        pass

    after_caller = caller_after = lambda self, i: None

    def wait_event(self, i, ev):
        pass

    def keep_alive(self, tensor, i):
        pass

    def event(self, i):
        class _Ev:
            def synchronize(self):
                pass
        return _Ev()

    def synchronize(self):
        pass


def _vdif_scan(src, nframe, frame_stride, header_nbytes, frames_per_set,
               thread_slot, nthread, frame_offset=None, check=None, bad=None,
               want_fields=True):
    from baseband_b200 import kernels
    buf = src.numpy()
    fields = np.zeros((kernels.VDIF_NFIELD, nframe), np.int32)
    nset = nframe // frames_per_set
    uo = np.full(max(nset * nthread, 1), -2, np.int64)
    slots = thread_slot.numpy()
    counter = bad if bad is not None else torch.zeros(1, dtype=torch.int32)
    bad = 0
    names = ['invalid_data', 'legacy_mode', 'seconds', 'ref_epoch',
             'frame_nr', 'vdif_version', 'lg2_nchan', 'frame_length',
             'complex_data', 'bits_per_sample', 'thread_id', 'station_id']
    first = None
    for i in range(nframe):
        off = int(frame_offset[i]) if frame_offset is not None \
            else i * frame_stride
        w = buf[off:off + 32].view('<u4') if header_nbytes == 32 else \
            np.concatenate([buf[off:off + 16].view('<u4'),
                            np.zeros(4, '<u4')])
        for k, name in enumerate(names):
            fields[k, i] = oheaders.field(w, *oheaders.VDIF_BASE_FIELDS[name])
        fields[12, i] = (int(w[4]) >> 24) & 0xff
        fields[13:17, i] = w[4:8].view(np.int32)
        s = i // frames_per_set
        if s >= nset:
            continue
        if i % frames_per_set == 0:
            first = fields[4, i]
            if check is not None and check[3] > 0:
                index0, seconds0, frame_nr0, fps = check
                index = ((int(fields[2, i]) - seconds0) * fps
                         + int(fields[4, i]) - frame_nr0)
                bad += index != index0 + s
        elif fields[4, i] != first:
            bad += 1
        slot = slots[fields[10, i]]
        if 0 <= slot < nthread:
            if uo[s * nthread + slot] != -2:
                bad += 1
            uo[s * nthread + slot] = -1 if fields[0, i] else off + header_nbytes
    missing = uo[:nset * nthread] == -2
    bad += int(missing.sum())
    uo[:nset * nthread][missing] = -1
    counter += int(bad)
    return (torch.from_numpy(fields) if want_fields else None,
            torch.from_numpy(uo[:nset * nthread]), counter)


def _mark5b_scan(src, nframe, frame_stride=10016, frame_offset=None,
                 check=None, bad=None, want_fields=True):
    from baseband_b200 import kernels
    buf = src.numpy()
    fields = np.zeros((kernels.M5B_NFIELD, nframe), np.int32)
    uo = np.empty(nframe, np.int64)
    for i in range(nframe):
        off = int(frame_offset[i]) if frame_offset is not None \
            else i * frame_stride
        if off < 0:                       # absent according to the index
            uo[i] = -1
            continue
        try:
            h = oheaders.mark5b_parse(buf[off:off + 16].copy().view('<u4'))
        except ValueError:
            # a time code that is not BCD: the kernel's bcd() gives -1
            w = buf[off:off + 16].copy().view('<u4')
            h = {'sync_pattern': int(w[0]), 'user': int(w[1]) >> 16,
                 'internal_tvg': (int(w[1]) >> 15) & 1,
                 'frame_nr': int(w[1]) & 0x7fff, 'bcd_jday': int(w[2]) >> 20,
                 'bcd_seconds': int(w[2]) & 0xfffff,
                 'bcd_fraction': int(w[3]) >> 16, 'crc': int(w[3]) & 0xffff,
                 'jday': -1, 'seconds': -1, 'fraction_ns': -1}
        pl = buf[off + 16:off + 10016].copy().view('<u4')
        valid = not bool((pl == 0x11223344).all())
        row = [h['sync_pattern'], h['user'], h['internal_tvg'], h['frame_nr'],
               h['bcd_jday'], h['bcd_seconds'], h['bcd_fraction'], h['crc'],
               h['jday'], h['seconds'], h['fraction_ns'], int(valid)]
        fields[:, i] = np.array(row, np.int64).astype(np.uint32).view(np.int32)
        uo[i] = off + 16 if valid else -1
        if check is not None and check[4] > 0:
            index0, jday0, seconds0, frame_nr0, fps = check
            dday = (h['jday'] - jday0 + 1500) % 1000 - 500
            index = ((h['seconds'] - seconds0 + 86400 * dday) * fps
                     + h['frame_nr'] - frame_nr0)
            if (h['sync_pattern'] != 0xABADDEED or h['jday'] < 0
                    or h['seconds'] < 0 or index != index0 + i):
                bad += 1
    return (torch.from_numpy(fields) if want_fields else None,
            torch.from_numpy(uo))


def _mark4_time_words(mjd0, ticks):
    """Time-code words 3 and 4 (CRC bits zero) at ``ticks`` quarter
    milliseconds after 00:00 of MJD mjd0: restates mark4_time_words of
    csrc/bb_scan.cu with the datetime module."""
    import datetime
    day, tick = divmod(int(ticks), 86400 * 4000)
    date = datetime.date(1858, 11, 17) + datetime.timedelta(int(mjd0) + day)
    yday = date.timetuple().tm_yday
    sec, qms = divmod(tick, 4000)

    def bcd(v):
        return ((v // 100) << 8) | ((v // 10 % 10) << 4) | (v % 10)

    w3 = ((date.year % 10) << 28) | (bcd(yday) << 16) \
        | (bcd(sec // 3600) << 8) | bcd(sec // 60 % 60)
    w4 = (bcd(sec % 60) << 24) | (bcd(qms // 4) << 12)
    return w3, w4


def _mark4_scan(src, nframe, ntrack, frame_stride=None, track=0,
                frame_offset=None, check=None, bad=None, want_words=True):
    buf = src.numpy()
    dtype = {16: '<u2', 32: '<u4', 64: '<u8'}[ntrack]
    if frame_stride is None:
        frame_stride = ntrack * 2500
    words5 = np.zeros((nframe, 5), np.uint32)
    uo = np.empty(nframe, np.int64)
    for i in range(nframe):
        off = int(frame_offset[i]) if frame_offset is not None \
            else i * frame_stride
        if off < 0:                       # absent according to the index
            uo[i] = -1
            continue
        stream = buf[off:off + ntrack * 20].copy().view(dtype)
        h = oheaders.mark4_parse(stream)
        words5[i] = oheaders.mark4_stream2words(stream)[:, track]
        uo[i] = off + ntrack * 20 if h['valid'] else -1
        if check is not None and check[3] > 0:
            index0, mjd0, tick0, tick_step = check
            w3, w4 = _mark4_time_words(mjd0, tick0 + tick_step * (index0 + i))
            if int(words5[i, 3]) != w3 \
                    or (int(words5[i, 4]) & 0xfffff000) != w4:
                bad += 1
    return (torch.from_numpy(words5.view(np.int32)) if want_words else None,
            torch.from_numpy(uo))


def _frames_assemble(headers, frame_nbytes, payload_nbytes=0, valid=None,
                     fill_word=0, units_per_frame=1, unit_stride=0):
    """numpy restatement of k_frames_assemble (csrc/bb_scan.cu)."""
    nframe, hn = headers.shape
    frames = torch.zeros((nframe, frame_nbytes), dtype=torch.uint8)
    f = frames.numpy()
    f[:, :hn] = headers.numpy()
    uo = (np.arange(nframe, dtype=np.int64)[:, None] * frame_nbytes + hn
          + np.arange(units_per_frame, dtype=np.int64) * unit_stride)
    if valid is not None:
        bad = valid.numpy() == 0
        uo[bad] = -1
        f[bad, hn:hn + payload_nbytes] = np.full(
            payload_nbytes // 4, fill_word, '<u4').view(np.uint8)
    return frames, torch.from_numpy(uo.reshape(-1))


_EMPTY = -1            # 0xffff...ff as int64


def _locate_frames(src, pattern, mask=None, frame_nbytes=0, pattern_offset=0,
                   own_stop=None, check=1, at_eof=True, base=0,
                   max_locations=None, unverified=None):
    """k_locate_frames restated with numpy (csrc/bb_index.cu)."""
    buf = src.numpy()
    nbytes = buf.size
    own_stop = nbytes if own_stop is None else min(own_stop, nbytes)
    from baseband_b200.kernels import _host_bytes
    pat = _host_bytes(pattern)[0]
    msk = (_host_bytes(mask)[0] if mask is not None
           else np.full(pat.size, 0xff, np.uint8))
    n = pat.size
    if nbytes < n:
        hits = np.zeros(0, bool)
    else:
        win = np.lib.stride_tricks.sliding_window_view(buf, n)
        hits = (((win ^ pat) & msk) == 0).all(-1)       # index = pattern pos
    found = []
    for loc in range(own_stop):
        at = loc + pattern_offset
        if at + n > nbytes or not hits[at]:
            continue
        if frame_nbytes > 0:
            if at_eof and loc + frame_nbytes > nbytes:
                continue
            if check:
                c = at + check * frame_nbytes
                if c >= 0 and c + n <= nbytes:
                    if not hits[c]:
                        if unverified is not None:
                            u = int(unverified[1])
                            if u < unverified[0].numel():
                                unverified[0][u] = base + loc
                            unverified[1].add_(1)
                        continue
                elif not at_eof and c + n > nbytes:
                    continue
        found.append(base + loc)
    if max_locations is None:
        max_locations = (8 * (nbytes // frame_nbytes + 2) if frame_nbytes
                         else 4096)
    out = np.zeros(max(int(max_locations), 1), np.int64)
    keep = found[:max_locations]
    out[:len(keep)] = keep
    return torch.from_numpy(out), torch.tensor([len(found)], dtype=torch.int32)


def _index_table(nentry, device):
    return torch.full((int(nentry),), _EMPTY, dtype=torch.int64)


def _scatter(table, at, off, invalid):
    t = table.numpy().view(np.uint64)
    t[at] = min(int(t[at]), 2 * int(off) + int(invalid))


def _vdif_index(src, base, locations, count, thread_slot, nthread, seconds0,
                frame_nr0, fps, nset_max, table, stats, thread0=-1):
    buf, slots, st = src.numpy(), thread_slot.numpy(), stats.numpy()
    for off in locations.numpy()[:min(int(count), locations.numel())]:
        w = buf[off - base:off - base + 16].copy().view('<u4')
        index = ((int(w[0]) & 0x3fffffff) - seconds0) * fps \
            + (int(w[1]) & 0xffffff) - frame_nr0
        slot = slots[(int(w[3]) >> 16) & 0x3ff]
        if not 0 <= slot < nthread:
            continue
        if not 0 <= index < nset_max:
            st[1] += 1
            continue
        _scatter(table, index * nthread + slot, off, int(w[0]) >> 31)
        st[0] = max(st[0], index)
        if (int(w[3]) >> 16) & 0x3ff == thread0 and st.size > 3:
            st[3] = max(st[3], index)


def _mark5b_index(src, base, locations, count, jday0, seconds0, frame_nr0,
                  fps, nset_max, table, stats):
    buf, st = src.numpy(), stats.numpy()

    def bcd(v, nd):
        out = 0
        for d in range(nd):
            nib = (v >> (4 * d)) & 0xf
            if nib > 9:
                return -1
            out += nib * 10 ** d
        return out
    for off in locations.numpy()[:min(int(count), locations.numel())]:
        w = buf[off - base:off - base + 16].copy().view('<u4')
        jday, seconds = bcd(int(w[2]) >> 20, 3), bcd(int(w[2]), 5)
        if jday < 0 or seconds < 0:
            st[2] += 1
            continue
        dday = (jday - jday0 + 1500) % 1000 - 500
        index = (seconds - seconds0 + 86400 * dday) * fps \
            + (int(w[1]) & 0x7fff) - frame_nr0
        if not 0 <= index < nset_max:
            st[1] += 1
            continue
        _scatter(table, index, off, 0)
        st[0] = max(st[0], index)


def _mark4_index(src, base, locations, count, ntrack, track, year0, yday0,
                 days_year0, days_prev_year, tick0, tick_step, nset_max,
                 table, stats, check_crc=True):
    """k_mark4_index restated with the oracle's bit transpose."""
    buf, st = src.numpy(), stats.numpy()
    dtype = {16: '<u2', 32: '<u4', 64: '<u8'}[ntrack]

    def bcd(v, nd):
        out = 0
        for d in range(nd):
            nib = (v >> (4 * d)) & 0xf
            if nib > 9:
                return -1
            out += nib * 10 ** d
        return out
    for off in locations.numpy()[:min(int(count), locations.numel())]:
        stream = buf[off - base:off - base + ntrack * 20].copy().view(dtype)
        words = oheaders.mark4_stream2words(stream)[:, track]
        if check_crc:
            from baseband_b200.base.utils import crc_remainder
            message = 0
            for word in words:
                message = (message << 32) | int(word)
            if crc_remainder(message, 0x180f, extend=False):
                st[2] += 1
                continue
        w3, w4 = int(words[3]), int(words[4])
        y, doy, hh, mm = bcd(w3 >> 28, 1), bcd(w3 >> 16, 3), \
            bcd(w3 >> 8, 2), bcd(w3, 2)
        ss, ms = bcd(w4 >> 24, 2), bcd(w4 >> 12, 3)
        if (y < 0 or not 1 <= doy <= 366 or not 0 <= hh <= 23
                or not 0 <= mm <= 59 or not 0 <= ss <= 60 or ms < 0
                or ms % 5 == 4):
            st[2] += 1
            continue
        dy = (y - year0 % 10 + 15) % 10 - 5
        if dy == 0:
            ddays = doy - yday0
        elif dy == 1:
            ddays = doy + days_year0 - yday0
        elif dy == -1:
            ddays = doy - days_prev_year - yday0
        else:
            st[1] += 1
            continue
        ticks = ddays * 86400 * 4000 + (hh * 3600 + mm * 60 + ss) * 4000 \
            + ms * 4 + ms % 5 - tick0
        if ticks % tick_step:
            st[2] += 1
            continue
        index = ticks // tick_step
        if not 0 <= index < nset_max:
            st[1] += 1
            continue
        _scatter(table, index, off, 0)
        st[0] = max(st[0], index)


def _index_table_finish(table):
    t = table.numpy().view(np.uint64)
    out = np.where((t == np.uint64(0xffffffffffffffff)) | (t & np.uint64(1)).astype(bool),
                   np.int64(-1), (t >> np.uint64(1)).astype(np.int64))
    return torch.from_numpy(out.astype(np.int64))


def _state_counts(src, unit_offset, nset, nthread, payload_nbytes, bps, nelem,
                  counts, set_origin=0, sets_per_bin=None):
    """numpy restatement of bb_state_counts (csrc/bb_counts.cu)."""
    buf = src.numpy()
    uo = unit_offset.numpy().reshape(nset, nthread)
    if sets_per_bin is None:
        sets_per_bin = max(1, set_origin + nset)
    out = counts.numpy()
    shifts = np.arange(0, 32, bps, dtype=np.uint32)
    for s in range(nset):
        b = (set_origin + s) // sets_per_bin
        for t in range(nthread):
            if uo[s, t] < 0:
                continue
            w = buf[uo[s, t]:uo[s, t] + payload_nbytes].view('<u4')
            codes = ((w[:, None] >> shifts) & ((1 << bps) - 1)).reshape(-1)
            for e in range(nelem):
                out[b, t, e] += np.bincount(codes[e::nelem],
                                            minlength=1 << bps)
    return counts


def _mark4_state_counts(src, unit_offset, nframe, nchan, fanout, ft, counts,
                        set_origin=0, sets_per_bin=None):
    """bb_mark4_state_counts restated with the (emulated) decoder: decoding
    with the level table (0, 1, 2, 3) yields the index 2 * sign + magnitude
    of every sample; header steps and invalid frames decode to the fill."""
    from baseband_b200 import kernels
    if sets_per_bin is None:
        sets_per_bin = max(1, set_origin + nframe)
    out = counts.numpy()
    if nframe == 0:
        return counts
    codes = kernels.mark4_decode(
        src, unit_offset, nframe, nchan, fanout, ft,
        levels=np.arange(4, dtype=np.float32), fill_value=-1.0).numpy()
    codes = codes.reshape(nframe, -1, nchan)
    for i in range(nframe):
        b = (set_origin + i) // sets_per_bin
        for c in range(nchan):
            col = codes[i, :, c]
            out[b, c] += np.bincount(col[col >= 0].astype(np.int64),
                                     minlength=4)
    return counts


def _int8_moments(src, unit_offset, nset, nthread, payload_nbytes, nelem,
                  moments, set_origin=0, sets_per_bin=None):
    """numpy restatement of bb_int8_moments (csrc/bb_counts.cu)."""
    buf = src.numpy()
    uo = unit_offset.numpy().reshape(nset, nthread)
    if sets_per_bin is None:
        sets_per_bin = max(1, set_origin + nset)
    out = moments.numpy()
    for s in range(nset):
        b = (set_origin + s) // sets_per_bin
        for t in range(nthread):
            if uo[s, t] < 0:
                continue
            x = buf[uo[s, t]:uo[s, t] + payload_nbytes].view(np.int8) \
                .astype(np.int64).reshape(-1, nelem)
            out[b, t, :, 0] += x.shape[0]
            out[b, t, :, 1] += x.sum(0)
            out[b, t, :, 2] += (x * x).sum(0)
    return moments


def install(monkeypatch):
    from baseband_b200 import _lib, device, kernels
    emu = emu_build.load()
    monkeypatch.setattr(_lib, '_lib', emu)
    monkeypatch.setattr(_lib, 'load', lambda: emu)
    cpu = torch.device('cpu')
    monkeypatch.setattr(device, 'resolve', lambda d=None: cpu)
    monkeypatch.setattr(device, 'default_device', lambda: cpu)
    monkeypatch.setattr(
        device, 'upload', lambda raw, dev, stage=None, align=16:
        raw.view(torch.uint8).reshape(-1) if isinstance(raw, torch.Tensor)
        else torch.from_numpy(np.ascontiguousarray(
            np.frombuffer(raw, np.uint8) if isinstance(
                raw, (bytes, bytearray, memoryview)) else raw
        ).view(np.uint8).reshape(-1).copy()))
    monkeypatch.setattr(device, 'download', lambda t: t.numpy())
    monkeypatch.setattr(device, 'pinned_empty',
                        lambda shape, dtype: torch.empty(shape, dtype=dtype))
    monkeypatch.setattr(device, 'Streams', _NoStreams)
    monkeypatch.setattr(device, 'register_host', lambda arr: None)
    monkeypatch.setattr(device, 'current_stream_synchronize', lambda d: None)

    def _dev(t, name, dtype=None):
        assert isinstance(t, torch.Tensor) and t.is_contiguous(), name
        if dtype is not None:
            assert t.dtype == dtype, name
        return ctypes.c_void_p(t.data_ptr())

    monkeypatch.setattr(kernels, '_dev', _dev)
    monkeypatch.setattr(kernels, '_require_cuda', lambda *a: None)
    monkeypatch.setattr(kernels, '_stream_ptr', lambda d: None)
    monkeypatch.setattr(kernels, 'vdif_scan', _vdif_scan)
    monkeypatch.setattr(kernels, 'mark5b_scan', _mark5b_scan)
    monkeypatch.setattr(kernels, 'mark4_scan', _mark4_scan)
    monkeypatch.setattr(kernels, 'frames_assemble', _frames_assemble)
    monkeypatch.setattr(kernels, 'state_counts', _state_counts)
    monkeypatch.setattr(kernels, 'int8_moments', _int8_moments)
    monkeypatch.setattr(kernels, 'mark4_state_counts', _mark4_state_counts)
    monkeypatch.setattr(kernels, 'locate_frames', _locate_frames)
    monkeypatch.setattr(kernels, 'index_table', _index_table)
    monkeypatch.setattr(kernels, 'vdif_index', _vdif_index)
    monkeypatch.setattr(kernels, 'mark5b_index', _mark5b_index)
    monkeypatch.setattr(kernels, 'mark4_index', _mark4_index)
    monkeypatch.setattr(kernels, 'index_table_finish', _index_table_finish)
    monkeypatch.setattr(
        kernels, 'new_counter',
        lambda dev: torch.zeros(1, dtype=torch.int32))
    monkeypatch.setattr(
        kernels, 'zeros',
        lambda shape, dtype, device: torch.zeros(shape, dtype=dtype))
    monkeypatch.setattr(kernels, '_on', lambda d: contextlib.nullcontext())
    monkeypatch.setattr(device, 'is_device_tensor',
                        lambda t: isinstance(t, torch.Tensor))
    return emu
