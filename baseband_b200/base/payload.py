"""Payload container: packed words in host memory, decoded/encoded on the GPU.

Same public surface as the reference's ``PayloadBase``
(baseband/base/payload.py:18-360): ``words``, ``data``, ``[item]``,
``[item] = values``, ``fromdata``, ``fromfile``, ``tofile``, ``shape``,
``dtype``, ``nbytes``, ``len``.  The arithmetic behind ``_decoders`` /
``_encoders`` runs in the CUDA library (see baseband_b200/codecs.py); indexing
decodes only the minimal range of words that contains the requested samples,
exactly as the reference does (:226-312).

Addition: ``todevice()`` / ``data_device`` return a CUDA ``torch.Tensor``.
"""
import operator
from functools import reduce

import numpy as np

from .shape import ArrayLike

__all__ = ['PayloadBase']


class PayloadBase(ArrayLike):
    _nbytes = None                      # fixed payload size, if any
    _memmap = False
    _dtype_word = np.dtype('<u4')
    _encoders = {}
    _decoders = {}
    _sample_shape_maker = None

    def __init__(self, words, *, header=None, sample_shape=(), bps=2,
                 complex_data=False):
        if header is not None:
            sample_shape = header.sample_shape
            bps = header.bps
            complex_data = header.complex_data
            if self._nbytes is None:
                self._nbytes = header.payload_nbytes
            elif self._nbytes != header.payload_nbytes:
                raise ValueError('header payload size should be {0}'
                                 .format(self._nbytes))
        self.words = words
        maker = self._sample_shape_maker
        self.sample_shape = (maker(*sample_shape) if maker is not None
                             else sample_shape)
        self._sample_size = reduce(operator.mul, sample_shape, 1)
        self.bps = bps
        self.complex_data = complex_data
        self._bpfs = bps * (2 if complex_data else 1) * self._sample_size
        self._coder = bps
        if self._nbytes is not None and self._nbytes != words.nbytes:
            raise ValueError('encoded data should have length {0}'
                             .format(self._nbytes))
        if words.dtype != self._dtype_word:
            raise ValueError('encoded data should have dtype {0}'
                             .format(self._dtype_word))

    # ------------------------------------------------------------- I/O
    @classmethod
    def fromfile(cls, fh, header=None, *, payload_nbytes=None, dtype=None,
                 memmap=None, **kwargs):
        if header is not None:
            payload_nbytes = header.payload_nbytes
            kwargs['header'] = header
        elif payload_nbytes is None:
            payload_nbytes = cls._nbytes
            if payload_nbytes is None:
                raise ValueError(
                    'payload_nbytes or header should be passed in if no '
                    'default payload size is defined on the class.')
        dtype = cls._dtype_word if dtype is None else np.dtype(dtype)
        memmap = cls._memmap if memmap is None else memmap
        if memmap and hasattr(fh, 'memmap'):
            return cls(fh.memmap(dtype=dtype, shape=(
                payload_nbytes // dtype.itemsize,)), **kwargs)
        if memmap and hasattr(fh, 'fileno'):
            try:
                offset = fh.tell()
                # map with the file's own mode (base/payload.py:126-131): a
                # file opened for writing ('w+b') gives a writable map that
                # extends the file, which is what ``memmap_frame`` relies on
                mode = str(getattr(fh, 'mode', 'rb')).replace('b', '')
                if mode not in ('r', 'r+', 'w+'):
                    mode = 'r'
                words = np.memmap(fh, mode=mode, dtype=dtype, offset=offset,
                                  shape=(payload_nbytes // dtype.itemsize,))
                fh.seek(offset + words.nbytes)
                return cls(words, **kwargs)
            except (OSError, ValueError, AttributeError):
                fh.seek(offset)
        raw = fh.read(payload_nbytes)
        if len(raw) < payload_nbytes:
            raise EOFError('could not read full payload.')
        return cls(np.frombuffer(raw, dtype=dtype), **kwargs)

    def tofile(self, fh):
        return fh.write(self.words.tobytes())

    @classmethod
    def fromdata(cls, data, header=None, bps=2, **kwargs):
        """Encode ``data`` (trailing dimensions = sample shape)."""
        data = np.asanyarray(data)
        sample_shape = data.shape[1:]
        complex_data = data.dtype.kind == 'c'
        if header:
            bps = header.bps
            if tuple(header.sample_shape) != tuple(sample_shape):
                raise ValueError(
                    'header is for sample_shape={} but data has {}'.format(
                        tuple(header.sample_shape), sample_shape))
            if header.complex_data != complex_data:
                raise ValueError('header is for {0} data but data are {1}'
                                 .format(*(('complex' if c else 'real')
                                           for c in (header.complex_data,
                                                     complex_data))))
            payload_nbytes = header.payload_nbytes
            base = {'header': header}
        else:
            base = {'bps': bps, 'sample_shape': sample_shape,
                    'complex_data': complex_data}
            payload_nbytes = (data.size * (2 if complex_data else 1)
                              * bps // 8)
        nword = payload_nbytes // cls._dtype_word.itemsize
        self = cls(np.empty(nword, cls._dtype_word), **base, **kwargs)
        self[:] = data
        return self

    # ---------------------------------------------------------- geometry
    nbytes = property(lambda self: self.words.nbytes)

    def __len__(self):
        return self.words.nbytes * 8 // self._bpfs

    @property
    def dtype(self):
        return np.dtype(np.complex64 if self.complex_data else np.float32)

    def _item_to_slices(self, item):
        """Smallest slice of ``words`` holding ``item`` plus the index into
        its decoded samples (semantics of base/payload.py:226-312)."""
        rest = ()
        if isinstance(item, tuple):
            rest = item[1:]
            item = item[0] if item else slice(None)
        nsample = len(self)
        is_slice = isinstance(item, slice)
        if is_slice:
            start, stop, step = item.indices(nsample)
            assert step > 0, 'cannot deal with negative steps yet.'
            count = stop - start
            step = None if step == 1 else step
        else:
            try:
                item = operator.index(item)
            except Exception:
                raise TypeError('{0} object can only be indexed or sliced.'
                                .format(type(self)))
            if item < 0:
                item += nsample
            if not 0 <= item < nsample:
                raise IndexError('{0} index out of range.'.format(type(self)))
            start, stop, step, count = item, item + 1, 1, 1

        def pick(first=None, last=None):
            return slice(first, last, step) if is_slice else (first or 0)

        if count == nsample:
            return slice(None), (pick(),) + rest
        bits_word = 8 * self.words.itemsize
        bits_sample = self._bpfs
        if bits_sample % bits_word == 0:          # >= 1 word per sample
            per = bits_sample // bits_word
            return slice(start * per, stop * per), (pick(),) + rest
        if bits_word % bits_sample == 0:          # several samples per word
            per = bits_word // bits_sample
            w0, o0 = divmod(start, per)
            w1, o1 = divmod(stop, per)
            words_slice = slice(w0, w1 + 1 if o1 else w1)
            data_slice = pick(o0 if o0 else None, o0 + count if o1 else None)
            return words_slice, (data_slice,) + rest
        raise TypeError('do not know how to extract data when full samples '
                        'have {0} bits and words have {1} bits'
                        .format(bits_sample, bits_word))

    # ------------------------------------------------------- arithmetic
    def _decode(self, words):
        return self._decoders[self._coder](words).view(self.dtype)

    def _encode(self, data):
        try:
            encoder = self._encoders[self._coder]
        except KeyError:
            raise ValueError('{} cannot encode data with {} bits'.format(
                self.__class__.__name__, self._coder)) from None
        if data.dtype.kind == 'c':
            data = data.view((data.real.dtype, (2,)))
        return encoder(data)

    def __getitem__(self, item=()):
        words_slice, data_slice = self._item_to_slices(item)
        decoded = self._decode(self.words[words_slice])
        return decoded.reshape((-1,) + tuple(self.sample_shape))[data_slice]

    def __setitem__(self, item, data):
        words_slice, data_slice = self._item_to_slices(item)
        data = np.asanyarray(data)
        nshape = len(self.sample_shape)
        whole = (data_slice == (slice(None),)
                 and data.shape[data.ndim - nshape:] == tuple(
                     self.sample_shape)
                 and data.dtype.kind == self.dtype.kind)
        if not whole:
            # read-modify-write of the touched words
            current = self._decode(self.words[words_slice])
            current = current.reshape((-1,) + tuple(self.sample_shape)).copy()
            current[data_slice] = data
            data = current
        encoded = np.ascontiguousarray(self._encode(data)).ravel().view(
            self._dtype_word)
        self.words[words_slice] = encoded

    data = property(__getitem__, doc='Full decoded payload.')

    # ------------------------------------------------ device-output option
    def todevice(self, device=None):
        """Decoded payload as a CUDA ``torch.Tensor`` of shape ``.shape``
        (float32, or complex64 for complex data)."""
        import torch
        from .. import device as _device
        dev = _device.resolve(device)
        flat = self._decode_device(_device.upload(self.words, dev))
        if self.complex_data:
            flat = torch.view_as_complex(flat.reshape(-1, 2))
        return flat.reshape(self.shape)

    def _decode_device(self, raw):
        from .. import codecs
        fn = self._decoders[self._coder]
        return codecs.decode_flat_device(raw, fn.bps, fn.levels, fn.codec)

    def __eq__(self, other):
        return (type(self) is type(other)
                and self.shape == other.shape
                and self.dtype == other.dtype
                and (self.words is other.words
                     or np.all(self.words == other.words)))

    def __ne__(self, other):
        return not self.__eq__(other)
