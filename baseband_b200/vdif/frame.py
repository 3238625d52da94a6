"""VDIF frames and frame sets.

A frame = header + payload of one thread; a frame set = the frames of all
threads for one time span, decoded to ``(nsample, nthread, nchan)``
(baseband/vdif/frame.py:21-128, :131-512).  Decoding a frame set is ONE kernel
launch: the payloads are uploaded together and the decode kernel interleaves
the threads into axis 1 while it unpacks (what the reference does with a
Python loop of strided copies, :427-434); invalid frames become ``fill_value``
inside the kernel (:79-90, base/frame.py:191-199).
"""
import numpy as np
import torch

from .. import device as _device
from .. import kernels
from ..base.frame import FrameBase
from .header import VDIFHeader, BASE_FIELDS
from .payload import VDIFPayload

__all__ = ['VDIFFrame', 'VDIFFrameSet']


class VDIFFrame(FrameBase):
    _header_class = VDIFHeader
    _payload_class = VDIFPayload

    def __init__(self, header, payload, valid=None, verify=True):
        self.header = header
        self.payload = payload
        if valid is not None:
            self.valid = valid
        if verify:
            self.verify()

    def verify(self):
        assert isinstance(self.header, VDIFHeader)
        assert isinstance(self.payload, VDIFPayload)
        assert self.payload.nbytes == self.header.payload_nbytes
        assert tuple(self.payload.sample_shape) == (self.header.nchan,)

    @property
    def valid(self):
        return not self.header['invalid_data']

    @valid.setter
    def valid(self, valid):
        self.header['invalid_data'] = not valid

    @classmethod
    def fromfile(cls, fh, edv=None, verify=True):
        header = VDIFHeader.fromfile(fh, edv, verify)
        payload = VDIFPayload.fromfile(fh, header=header)
        return cls(header, payload, verify=verify)

    @classmethod
    def fromdata(cls, data, header=None, verify=True, **kwargs):
        if header is None:
            header = VDIFHeader.fromvalues(verify=verify, **kwargs)
        payload = VDIFPayload.fromdata(data, header=header)
        return cls(header, payload, verify=verify)

    @classmethod
    def from_mark5b_frame(cls, mark5b_frame, verify=True, **kwargs):
        """Wrap a Mark 5B frame as VDIF EDV 0xab (vdif/frame.py:104-128)."""
        from ..timeutil import Time
        m5h = mark5b_frame.header
        # whole seconds go into the VDIF time code; the position within the
        # second is the Mark 5B frame number (vdif/header.py:799-843)
        whole = Time(m5h.kday + m5h.jday, m5h.seconds)
        kwargs.update(edv=0xab, time=whole, bps=mark5b_frame.payload.bps,
                      nchan=mark5b_frame.payload.sample_shape[0],
                      complex_data=False)
        header = VDIFHeader.fromvalues(verify=False, **kwargs)
        header.mutable = True
        for key in ('user', 'internal_tvg', 'bcd_jday', 'bcd_seconds',
                    'bcd_fraction', 'crc'):
            header[key] = m5h[key]
        header['mark5b_frame_nr'] = m5h['frame_nr']
        header['frame_nr'] = m5h['frame_nr']
        payload = VDIFPayload(mark5b_frame.payload.words, header)
        return cls(header, payload, valid=mark5b_frame.valid, verify=verify)


def _sample_range(item, nsample):
    """Normalise the sample part of an index: (start, stop, local item)."""
    if isinstance(item, slice):
        start, stop, step = item.indices(nsample)
        assert step > 0, 'cannot deal with negative steps yet.'
        stop = max(stop, start)
        return start, stop, slice(None, None, None if step == 1 else step)
    index = int(item)
    if index < 0:
        index += nsample
    if not 0 <= index < nsample:
        raise IndexError('sample index out of range.')
    return index, index + 1, 0


class VDIFFrameSet:
    """Frames of several threads covering the same time span."""

    def __init__(self, frames, header0=None):
        self.frames = frames
        self.header0 = frames[0].header if header0 is None else header0

    # ---------------------------------------------------------------- I/O
    @classmethod
    def fromfile(cls, fh, thread_ids=None, edv=None, verify=True):
        """Read frames while the frame number stays the same and no thread
        repeats; keep the requested threads, ordered as requested (or by id)."""
        header0 = VDIFHeader.fromfile(fh, edv, verify)
        edv = header0.edv
        frame_nr = header0['frame_nr']
        found = {}
        header = header0
        while True:
            tid = header['thread_id']
            if header['frame_nr'] != frame_nr or tid in found:
                fh.seek(-header.nbytes, 1)      # belongs to the next set
                break
            if thread_ids is None or tid in thread_ids:
                payload = VDIFPayload.fromfile(fh, header=header)
                found[tid] = VDIFFrame(header, payload, verify=False)
            else:
                fh.seek(header.payload_nbytes, 1)
            try:
                header = VDIFHeader.fromfile(fh, edv, verify)
            except (EOFError, AssertionError):
                if thread_ids is None or len(found) == len(thread_ids):
                    break
                raise
        if thread_ids and len(found) < len(thread_ids):
            raise OSError('could not find all requested frames.')
        order = sorted(found) if thread_ids is None else thread_ids
        return cls([found[tid] for tid in order], header0)

    def tofile(self, fh):
        for frame in self.frames:
            frame.tofile(fh)

    @classmethod
    def fromdata(cls, data, headers=None, verify=True, **kwargs):
        """Encode ``data`` of shape (samples_per_frame, nthread, nchan)."""
        data = np.asanyarray(data)
        assert data.ndim == 3
        if not isinstance(headers, (list, tuple)):
            if headers is None:
                kwargs.setdefault('thread_id', 0)
                first = VDIFHeader.fromvalues(verify=verify, **kwargs)
            else:
                first = headers.copy()
            headers = []
            for tid in range(data.shape[1]):
                h = first.copy()
                h.mutable = True
                h['thread_id'] = tid
                headers.append(h)
        frames = [VDIFFrame.fromdata(np.ascontiguousarray(data[:, i]), h,
                                     verify=verify)
                  for i, h in enumerate(headers)]
        return cls(frames)

    # ------------------------------------------------------------ geometry
    @property
    def nbytes(self):
        return len(self.frames) * self.frames[0].nbytes

    @property
    def sample_shape(self):
        return (len(self.frames),) + tuple(self.frames[0].sample_shape)

    def __len__(self):
        return len(self.frames[0])

    @property
    def shape(self):
        return (len(self),) + self.sample_shape

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return self.frames[0].dtype

    @property
    def valid(self):
        valid = np.array([f.valid for f in self.frames])
        return valid[0] if len(np.unique(valid)) == 1 else valid

    @valid.setter
    def valid(self, valid):
        for f, v in zip(self.frames,
                        np.broadcast_to(valid, (len(self.frames),))):
            f.valid = v

    @property
    def fill_value(self):
        return self.frames[0].fill_value

    @fill_value.setter
    def fill_value(self, fill_value):
        for f in self.frames:
            f.fill_value = fill_value

    # -------------------------------------------------------------- decode
    def _decode(self, frames, start, stop, device=None):
        """Samples [start, stop) of ``frames`` in one launch ->
        float32/complex64 CUDA tensor (n, len(frames), nchan)."""
        dev = _device.resolve(device)
        pl0 = frames[0].payload
        nbytes = pl0.nbytes
        nelem = pl0._sample_size * (2 if pl0.complex_data else 1)
        if pl0._bpfs != pl0.bps * nelem:
            raise TypeError('cannot decode payloads whose words have unused '
                            'space (bps={})'.format(pl0.bps))
        host = np.empty((len(frames), nbytes), np.uint8)
        offsets = np.empty(len(frames), np.int64)
        for i, f in enumerate(frames):
            host[i] = np.ascontiguousarray(f.payload.words).view(np.uint8)
            offsets[i] = i * nbytes if f.valid else -1
        raw = _device.upload(host, dev)
        uo = torch.from_numpy(offsets).to(dev)
        fn = pl0._decoders[pl0._coder]
        out = kernels.decode_bitfield(
            raw, uo, 1, len(frames), nbytes, fn.bps, nelem, pl0.complex_data,
            fn.codec, fn.levels, float(self.fill_value), start, stop - start)
        if pl0.complex_data:
            return torch.view_as_complex(out.view(
                stop - start, len(frames), nelem // 2, 2))
        return out

    def _select(self, item):
        """-> frames, (start, stop), local index to apply after decoding."""
        if not isinstance(item, tuple):
            item = (item,)
        sample_item = item[0] if item else slice(None)
        start, stop, local = _sample_range(sample_item, len(self))
        if len(item) > 1:
            which = np.arange(len(self.frames))[item[1]]
            assert which.ndim <= 1
            frames = [self.frames[i] for i in np.atleast_1d(which)]
            thread_local = 0 if which.ndim == 0 else slice(None)
        else:
            frames, thread_local = self.frames, slice(None)
        return frames, start, stop, local, thread_local, tuple(item[2:])

    def __getitem__(self, item=()):
        if isinstance(item, str):
            if item == 'thread_id':
                return np.array([f.header[item] for f in self.frames])
            if item != 'invalid_data' and item in BASE_FIELDS:
                return self.header0[item]
            values = np.array([f.header[item] for f in self.frames])
            return values[0] if len(np.unique(values)) == 1 else values
        frames, start, stop, local, thread_local, rest = self._select(item)
        block = _device.download(self._decode(frames, start, stop))
        if rest:
            block = block[(slice(None), slice(None)) + rest]
        block = block[local]
        if thread_local == 0:
            block = block[0] if local == 0 else block[:, 0]
        return block

    data = property(__getitem__, doc='Full decoded frame set.')

    def todevice(self, device=None):
        """Decoded frame set as a CUDA tensor (nsample, nthread, nchan)."""
        return self._decode(self.frames, 0, len(self), device)

    def __setitem__(self, item, data):
        if isinstance(item, str):
            if isinstance(data, (int, np.integer)):
                values = [int(data)] * len(self.frames)
            elif isinstance(data, (tuple, list)) and all(
                    isinstance(d, (int, np.integer)) for d in data):
                values = list(data)
            else:
                raise ValueError('header items can only be set to integers.')
            distinct = len(set(values))
            if item == 'thread_id':
                if distinct != len(self.frames):
                    raise ValueError('all thread ids should be unique.')
            elif (item != 'invalid_data' and item in BASE_FIELDS
                  and distinct > 1):
                raise ValueError('base header keys should be identical.')
            for f, v in zip(self.frames, values):
                f.header[item] = v
            return
        frames, start, stop, local, thread_local, rest = self._select(item)
        data = np.asanyarray(data)
        whole = (start == 0 and stop == len(self) and local == slice(None)
                 and not rest)
        if whole:
            kind = np.complex64 if self.dtype.kind == 'c' else data.dtype
            block = np.empty((len(self), len(frames)) + self.sample_shape[1:],
                             dtype=kind if data.dtype.kind in 'fc'
                             else np.float64)
        else:
            block = _device.download(self._decode(frames, start, stop)).copy()
        if thread_local == 0:
            view = block[:, 0]
            view = view[(slice(None),) + rest] if rest else view
            view[local] = data
        else:
            view = block[(slice(None), slice(None)) + rest] if rest else block
            view[local] = data
        for i, f in enumerate(frames):
            f.payload[start:stop] = np.ascontiguousarray(block[:, i])

    # ------------------------------------------------------------- headers
    def keys(self):
        return self.header0.keys()

    def __contains__(self, key):
        return key in self.header0

    def __getattr__(self, attr):
        if attr.startswith('_') or attr in ('frames', 'header0'):
            raise AttributeError(attr)
        header0 = self.__dict__.get('header0')
        if header0 is not None and attr in header0._properties:
            values = [getattr(f.header, attr) for f in self.frames]
            if all(v == values[0] for v in values[1:]):
                return values[0]
            return np.array(values)
        raise AttributeError('{} has no attribute {!r}'.format(
            type(self).__name__, attr))

    def __eq__(self, other):
        return (type(self) is type(other)
                and len(self.frames) == len(other.frames)
                and self.header0 == other.header0
                and all(a == b for a, b in zip(self.frames, other.frames)))
