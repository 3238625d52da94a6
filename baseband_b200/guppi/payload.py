"""GUPPI payloads: int8 samples, stored channels first
``(nchan, nsample, npol, re/im)`` (default) or time first
``(nsample, nchan, npol, re/im)`` (``PKTFMT = 'SIMPLE'``); decoded to
``(nsample, npol, nchan)`` (baseband/guppi/payload.py:21-138).

The channels-first case is a batched 2-D transpose fused with the
int8 -> float32 conversion in one kernel (``bb_decode_int8_transposed``),
which replaces the reference's ``astype`` + ``.T.reshape`` copies.
"""
from collections import namedtuple

import numpy as np
import torch

from .. import device as _device
from .. import kernels
from ..base.payload import PayloadBase

__all__ = ['decode_8bit', 'encode_8bit', 'GUPPIPayload']


def _const(dev, *values):
    return torch.tensor(values, dtype=torch.int64, device=dev)


def decode_device(raw, offsets, nsample, npol, nchan, complex_data,
                  channels_first, col_begin=None, col_end=None,
                  out_col0=None, out=None):
    """Decode payloads at byte ``offsets`` (host list) inside the uint8 CUDA
    tensor ``raw`` -> float32 CUDA tensor laid out (time, pol, chan[, 2]).
    ``col_*`` select a time window per payload (in samples)."""
    dev = raw.device
    ib = 2 if complex_data else 1
    nunit = len(offsets)
    begin = np.zeros(nunit, np.int64) if col_begin is None \
        else np.asarray(col_begin, np.int64)
    end = np.full(nunit, nsample, np.int64) if col_end is None \
        else np.asarray(col_end, np.int64)
    first = (np.concatenate([[0], np.cumsum(end - begin)[:-1]])
             if out_col0 is None else np.asarray(out_col0, np.int64))
    total = int((end - begin).sum())
    if out is None:
        out = torch.empty(total * npol * nchan * ib, dtype=torch.float32,
                          device=dev)
    if channels_first:
        tables = torch.from_numpy(np.stack(
            [np.asarray(offsets, np.int64), begin * npol, end * npol,
             first * npol])).to(dev)
        kernels.decode_int8_transposed(raw, tables[0], nunit, nchan,
                                       nsample * npol, ib, tables[1],
                                       tables[2], tables[3], out)
        return out
    # time first: the (chan, pol) axes of every sample swap in the kernel
    tables = torch.from_numpy(np.stack(
        [np.asarray(offsets, np.int64), begin, end, first])).to(dev)
    kernels.decode_int8_timefirst(raw, tables[0], nunit, nsample, nchan, npol,
                                  ib, tables[1], tables[2], tables[3], out)
    return out


# codec callables under their reference names (guppi/payload.py:13-18)
def decode_8bit(words):
    from .. import codecs
    return codecs.INT8_DECODERS[8](words)


def encode_8bit(values):
    from .. import codecs
    return codecs.INT8_ENCODERS[8](values)


class GUPPIPayload(PayloadBase):
    _dtype_word = np.dtype('int8')
    _memmap = True
    _sample_shape_maker = namedtuple('SampleShape', 'npol, nchan')

    def __init__(self, words, *, header=None, sample_shape=(), bps=8,
                 complex_data=False, channels_first=True):
        from .. import codecs
        self._decoders = codecs.INT8_DECODERS
        self._encoders = codecs.INT8_ENCODERS
        super().__init__(words, header=header, sample_shape=sample_shape,
                         bps=bps, complex_data=complex_data)
        self.channels_first = (channels_first if header is None
                               else header.channels_first)

    @classmethod
    def fromdata(cls, data, header=None, bps=8, channels_first=True):
        data = np.asanyarray(data)
        complex_data = data.dtype.kind == 'c'
        sample_shape = data.shape[1:]
        if header is not None:
            if tuple(header.sample_shape) != tuple(sample_shape):
                raise ValueError('header is for sample_shape={} but data has '
                                 '{}'.format(tuple(header.sample_shape),
                                             sample_shape))
            if header.complex_data != complex_data:
                raise ValueError('header and data disagree on whether the '
                                 'data are complex.')
            kw = {'header': header}
            nbytes = header.payload_nbytes
        else:
            kw = {'sample_shape': sample_shape, 'bps': bps,
                  'complex_data': complex_data,
                  'channels_first': channels_first}
            nbytes = data.size * (2 if complex_data else 1) * bps // 8
        self = cls(np.empty(nbytes, cls._dtype_word), **kw)
        self[:] = data
        return self

    def __len__(self):
        return self.nbytes * 8 // self._bpfs

    def _decode_range(self, start, stop, device=None):
        if self.bps not in self._decoders:
            raise KeyError(self.bps)
        dev = _device.resolve(device)
        npol, nchan = self.sample_shape
        raw = _device.upload(self.words, dev)
        out = decode_device(raw, [0], len(self), npol, nchan,
                            self.complex_data, self.channels_first,
                            [start], [stop])
        if self.complex_data:
            return torch.view_as_complex(out.view(stop - start, npol, nchan,
                                                  2))
        return out.view(stop - start, npol, nchan)

    def _range(self, item):
        rest = ()
        if isinstance(item, tuple):
            item, rest = (item[0], item[1:]) if item else (slice(None), ())
        n = len(self)
        if isinstance(item, slice):
            start, stop, step = item.indices(n)
            assert step > 0, 'cannot deal with negative steps yet.'
            return start, max(start, stop), slice(None, None, step), rest
        import operator
        try:
            index = operator.index(item)
        except Exception:
            raise TypeError('{0} object can only be indexed or sliced.'
                            .format(type(self)))
        if index < 0:
            index += n
        if not 0 <= index < n:
            raise IndexError('{0} index out of range.'.format(type(self)))
        return index, index + 1, 0, rest

    def __getitem__(self, item=()):
        start, stop, local, rest = self._range(item)
        data = _device.download(self._decode_range(start, stop))[local]
        if not rest:
            return data
        return data[rest] if local == 0 else data[(slice(None),) + rest]

    data = property(__getitem__, doc='Full decoded payload.')

    def todevice(self, device=None):
        return self._decode_range(0, len(self), device)

    def __setitem__(self, item, data):
        start, stop, local, rest = self._range(item)
        data = np.asanyarray(data)
        npol, nchan = self.sample_shape
        whole = (local == slice(None, None, 1) and not rest
                 and data.shape == (stop - start, npol, nchan)
                 and data.dtype.kind == self.dtype.kind)
        if not whole:
            current = _device.download(self._decode_range(start, stop)).copy()
            if local == 0:
                current[(0,) + rest] = data
            else:
                current[(local,) + rest] = data
            data = current
        try:
            encoder = self._encoders[self.bps]
        except KeyError:
            raise ValueError('{} cannot encode data with {} bits'.format(
                type(self).__name__, self.bps)) from None
        del encoder
        dev = _device.resolve(None)
        if data.dtype not in (np.float32, np.complex64):
            data = data.astype(np.complex128 if data.dtype.kind == 'c'
                               else np.float64)
        t = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
        if t.is_complex():
            t = torch.view_as_real(t)
        n = stop - start
        ib = 2 if self.complex_data else 1
        words = self.words.view(np.uint8)
        if self.channels_first:
            packed = torch.empty(nchan * n * npol * ib, dtype=torch.uint8,
                                 device=dev)
            kernels.encode_int8_transposed(
                t.reshape(-1), packed, _const(dev, 0), 1, nchan, n * npol, ib)
            rows = _device.download(packed).reshape(nchan, n * npol * ib)
            words.reshape(nchan, -1)[:, start * npol * ib:
                                     stop * npol * ib] = rows
        else:
            packed = torch.empty(n * nchan * npol * ib, dtype=torch.uint8,
                                 device=dev)
            kernels.encode_int8_timefirst(t.reshape(-1), packed,
                                          _const(dev, 0), 1, n, nchan, npol,
                                          ib)
            per = nchan * npol * ib
            words[start * per:stop * per] = _device.download(packed)
