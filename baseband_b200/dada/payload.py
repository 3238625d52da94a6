"""DADA payloads: int8 samples ordered ``(time, pol, chan, re/im)``
(baseband/dada/payload.py:21-51); MeerKAT beamformer (MKBF) files store heaps
of 256 time samples as ``(heap, pol, chan, 256, re/im)`` (:54-89), undone on
the GPU by the transposing int8 kernel."""
from collections import namedtuple

import numpy as np
import torch

from .. import codecs, device as _device, kernels
from ..base.payload import PayloadBase

__all__ = ['decode_8bit', 'encode_8bit', 'DADAPayload', 'MKBFPayload',
           'decode_device']

HEAP = 256


def decode_device(raw, offset, nbytes, npol, nchan, complex_data, mkbf,
                  start, count, out=None):
    """Samples [start, start+count) of the payload at byte ``offset`` of the
    uint8 CUDA tensor ``raw`` -> float32 CUDA tensor (time, pol, chan[, 2])."""
    dev = raw.device
    ib = 2 if complex_data else 1
    nelem = npol * nchan * ib
    if out is None:
        out = torch.empty(count * nelem, dtype=torch.float32, device=dev)
    if count == 0:
        return out
    if not mkbf:
        uo = torch.tensor([offset], dtype=torch.int64, device=dev)
        kernels.decode_bitfield(raw, uo, 1, 1, nbytes, 8, nelem,
                                complex_data, kernels.CODEC_SINT, None, 0.,
                                start, count, out)
        return out
    heap_nbytes = nelem * HEAP
    h0, h1 = start // HEAP, -(-(start + count) // HEAP)
    n = h1 - h0
    heaps = np.arange(h0, h1, dtype=np.int64)
    begin = np.zeros(n, np.int64)
    end = np.full(n, HEAP, np.int64)
    begin[0] = start - h0 * HEAP
    end[-1] = start + count - (h1 - 1) * HEAP
    first = np.concatenate([[0], np.cumsum(end - begin)[:-1]])
    tables = torch.from_numpy(np.stack(
        [offset + heaps * heap_nbytes, begin, end, first])).to(dev)
    kernels.decode_int8_transposed(raw, tables[0], n, npol * nchan, HEAP, ib,
                                   tables[1], tables[2], tables[3], out)
    return out


# codec callables under their reference names (dada/payload.py:13-18)
def decode_8bit(words):
    from .. import codecs
    return codecs.INT8_DECODERS[8](words)


def encode_8bit(values):
    from .. import codecs
    return codecs.INT8_ENCODERS[8](values)


class DADAPayload(PayloadBase):
    _decoders = codecs.INT8_DECODERS
    _encoders = codecs.INT8_ENCODERS
    _memmap = True
    _sample_shape_maker = namedtuple('SampleShape', 'npol, nchan')
    _mkbf = False

    def __new__(cls, words, *, header=None, **kwargs):
        if header is not None and header.get('INSTRUMENT') == 'MKBF':
            cls = MKBFPayload
        return super().__new__(cls)

    def todevice(self, device=None):
        dev = _device.resolve(device)
        npol, nchan = self.sample_shape
        out = decode_device(_device.upload(self.words, dev), 0, self.nbytes,
                            npol, nchan, self.complex_data, self._mkbf, 0,
                            len(self))
        if self.complex_data:
            return torch.view_as_complex(out.view(len(self), npol, nchan, 2))
        return out.view(len(self), npol, nchan)


class MKBFPayload(DADAPayload):
    """Heaps of 256 samples; indexing decodes whole heaps on the GPU."""
    _mkbf = True

    def _heap_range(self, item):
        rest = ()
        if isinstance(item, tuple):
            item, rest = (item[0], item[1:]) if item else (slice(None), ())
        n = len(self)
        if isinstance(item, slice):
            start, stop, step = item.indices(n)
            assert step > 0, 'cannot deal with negative steps yet.'
            return start, max(start, stop), slice(None, None, step), rest
        import operator
        index = operator.index(item)
        if index < 0:
            index += n
        if not 0 <= index < n:
            raise IndexError('{0} index out of range.'.format(type(self)))
        return index, index + 1, 0, rest

    def __getitem__(self, item=()):
        if self.bps not in self._decoders:
            raise KeyError(self.bps)
        start, stop, local, rest = self._heap_range(item)
        dev = _device.resolve(None)
        npol, nchan = self.sample_shape
        out = decode_device(_device.upload(self.words, dev), 0, self.nbytes,
                            npol, nchan, self.complex_data, True, start,
                            stop - start)
        n = stop - start
        data = _device.download(out).reshape(
            (n, npol, nchan) + ((2,) if self.complex_data else ()))
        if self.complex_data:
            data = np.ascontiguousarray(data).view(np.complex64)[..., 0]
        data = data[local]
        if not rest:
            return data
        return data[rest] if local == 0 else data[(slice(None),) + rest]

    data = property(__getitem__, doc='Full decoded payload.')

    def __setitem__(self, item, data):
        start, stop, local, rest = self._heap_range(item)
        data = np.asanyarray(data)
        npol, nchan = self.sample_shape
        h0, h1 = start // HEAP, -(-stop // HEAP)
        a, b = h0 * HEAP, min(h1 * HEAP, len(self))
        whole = (start == a and stop == b and local == slice(None, None, 1)
                 and not rest and data.shape == (b - a, npol, nchan)
                 and data.dtype.kind == self.dtype.kind)
        if whole:
            block = data
        else:
            block = self[a:b].copy()
            view = block[start - a:stop - a]
            if local == 0:
                view[(0,) + rest] = data
            else:
                view[(local,) + rest] = data
        if self.bps not in self._encoders:
            raise ValueError('{} cannot encode data with {} bits'.format(
                type(self).__name__, self.bps))
        if block.dtype not in (np.float32, np.complex64):
            block = block.astype(np.complex128 if block.dtype.kind == 'c'
                                 else np.float64)
        dev = _device.resolve(None)
        t = torch.from_numpy(np.ascontiguousarray(block)).to(dev)
        if t.is_complex():
            t = torch.view_as_real(t)
        ib = 2 if self.complex_data else 1
        nheap = (b - a) // HEAP
        heap_nbytes = npol * nchan * HEAP * ib
        packed = torch.empty(nheap * heap_nbytes, dtype=torch.uint8,
                             device=dev)
        uo = torch.arange(nheap, dtype=torch.int64, device=dev) * heap_nbytes
        kernels.encode_int8_transposed(t.reshape(-1), packed, uo, nheap,
                                       npol * nchan, HEAP, ib)
        self.words.view(np.uint8)[h0 * heap_nbytes:
                                  h0 * heap_nbytes + nheap * heap_nbytes] = (
            _device.download(packed))
