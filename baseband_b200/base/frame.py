"""Frame = header + payload, with validity and fill_value.

Behaviour of baseband/base/frame.py:58-241: indexing a frame with a string
reads the header, anything else reads samples; an invalid frame yields
``fill_value`` without touching the payload (:191-199).
"""
import numpy as np

from .shape import ArrayLike

__all__ = ['FrameBase']


class FrameBase(ArrayLike):
    _header_class = None
    _payload_class = None
    _fill_value = 0.
    _valid = True

    def __init__(self, header, payload, valid=None, verify=True):
        self.header = header
        self.payload = payload
        if valid is not None:
            self.valid = valid
        if verify:
            self.verify()

    def verify(self):
        assert isinstance(self.header, self._header_class)
        assert isinstance(self.payload, self._payload_class)
        nbytes = getattr(self.header, 'payload_nbytes', None)
        if nbytes is not None:
            assert self.payload.nbytes == nbytes

    @property
    def valid(self):
        return self._valid

    @valid.setter
    def valid(self, valid):
        self._valid = bool(valid)

    @classmethod
    def fromfile(cls, fh, memmap=None, valid=None, verify=True, **kwargs):
        header = cls._header_class.fromfile(fh, verify=verify)
        payload = cls._payload_class.fromfile(fh, header=header,
                                              memmap=memmap, **kwargs)
        return cls(header, payload, valid=valid, verify=verify)

    def tofile(self, fh):
        self.header.tofile(fh)
        self.payload.tofile(fh)

    @classmethod
    def fromdata(cls, data, header=None, *, valid=None, verify=True,
                 **kwargs):
        if header is None:
            header = cls._header_class.fromvalues(verify=verify, **kwargs)
        payload = cls._payload_class.fromdata(data, header=header)
        return cls(header, payload, valid=valid, verify=verify)

    sample_shape = property(lambda self: self.payload.sample_shape)
    dtype = property(lambda self: self.payload.dtype)
    nbytes = property(lambda self: self.header.nbytes + self.payload.nbytes)

    def __len__(self):
        return len(self.payload)

    @property
    def fill_value(self):
        return self._fill_value

    @fill_value.setter
    def fill_value(self, fill_value):
        self._fill_value = fill_value

    def __getitem__(self, item=()):
        if isinstance(item, str):
            return self.header[item]
        if self.valid:
            return self.payload[item]
        shape = np.empty(self.shape, dtype=bool)[item].shape
        return np.full(shape, self.fill_value, dtype=self.dtype)

    data = property(__getitem__, doc='Full decoded frame.')

    def __setitem__(self, item, value):
        if isinstance(item, str):
            self.header[item] = value
        else:
            self.payload[item] = value

    def keys(self):
        return self.header.keys()

    def __contains__(self, key):
        return key in self.keys()

    def __getattr__(self, attr):
        # header properties (time, sample_rate, ...) show through the frame
        if attr.startswith('_') or attr in ('header', 'payload'):
            raise AttributeError(attr)
        header = self.__dict__.get('header')
        if header is not None and (
                attr in getattr(header, '_properties', ())
                or attr in ('get_time', 'set_time', 'update')):
            return getattr(header, attr)
        raise AttributeError('{} has no attribute {!r}'.format(
            type(self).__name__, attr))

    def __eq__(self, other):
        return (type(self) is type(other) and self.valid == other.valid
                and self.header == other.header
                and self.payload == other.payload)
