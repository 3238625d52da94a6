"""Bit-field headers.

A header is a short list of 32-bit words; a field table maps names to
``(word, first bit, number of bits[, default])``.  Reading a field is
``(words[word] >> bit) & mask`` — the same arithmetic the GPU scan kernels do
in bulk (csrc/bb_scan.cu) — and this host class is used for the handful of
headers a reader inspects at open time and for the headers a writer emits.

Mirrors the behaviour of baseband/base/header.py:35-144 (parser / setter
semantics incl. 64-bit fields, `True` = all ones, `None` = default) and
:278-500 (dict-like access, copy, equality, fromvalues/update ordering).
"""
import struct
from copy import copy

import numpy as np

__all__ = ['FieldTable', 'BitFieldHeader', 'four_word_struct',
           'eight_word_struct']

four_word_struct = struct.Struct('<4I')
eight_word_struct = struct.Struct('<8I')


class FieldTable(dict):
    """``{name: (word, bit, nbits[, default])}`` with merge via ``|``."""

    def __or__(self, other):
        new = FieldTable(self)
        new.update(other)
        return new

    def default(self, key):
        spec = self[key]
        return spec[3] if len(spec) > 3 else None

    @property
    def defaults(self):
        return {k: self.default(k) for k in self}

    def get_field(self, words, key):
        word, bit, nbits = self[key][:3]
        if nbits == 64:
            return int(words[word]) + (int(words[word + 1]) << 32)
        value = (words[word] >> bit) & ((1 << nbits) - 1)
        if nbits == 1:
            return value != 0
        return value

    def set_field(self, words, key, value):
        word, bit, nbits = self[key][:3]
        mask = (1 << nbits) - 1
        if value is None:
            value = self.default(key)
            if value is None:
                raise ValueError("no default value so cannot set to 'None'.")
        if value is True:
            value = mask
        elif isinstance(value, (np.integer, np.bool_, bool)):
            value = int(value)
        if np.any((value & mask) != value):
            raise ValueError('{0} cannot be represented with {1} bits'
                             .format(value, nbits))
        if nbits == 64:
            words[word] = value & 0xffffffff
            words[word + 1] = value >> 32
        else:
            words[word] = (words[word] & ~(mask << bit)
                           & 0xffffffff) | (value << bit)


class BitFieldHeader:
    """Header built on a :class:`FieldTable` (class attribute ``_fields``)."""

    _fields = FieldTable()
    _struct = None
    _properties = ('payload_nbytes', 'frame_nbytes', 'time')

    def __init__(self, words, verify=True):
        if words is None:
            self.words = [0] * (self._struct.size // 4)
        else:
            self.words = words
        if verify:
            self.verify()

    def verify(self):
        pass

    # -- mutability follows the container type, as in the reference --------
    @property
    def mutable(self):
        return isinstance(self.words, (list, np.ndarray))

    @mutable.setter
    def mutable(self, mutable):
        if isinstance(self.words, np.ndarray):
            self.words.flags['WRITEABLE'] = bool(mutable)
        elif mutable:
            self.words = list(self.words)
        else:
            self.words = tuple(self.words)

    def copy(self, **kwargs):
        kwargs.setdefault('verify', False)
        words = (self.words.copy() if isinstance(self.words, np.ndarray)
                 else list(self.words))
        return self.__class__(words, **kwargs)

    __copy__ = copy

    # -- what does not change within a stream -----------------------------
    def invariants(self):
        """Keys of the header parts shared by all headers of a stream
        (baseband/base/header.py:564-586): the sync pattern unless a class
        says more."""
        return {'sync_pattern'} if 'sync_pattern' in self._fields else set()

    def invariant_pattern(self, invariants=None):
        """``(pattern, mask)``: the header words and, for each, the bits that
        belong to invariant fields -- what a sync search can match on
        (baseband/base/header.py:588-647, instance form)."""
        if invariants is None:
            invariants = self.invariants()
        if not invariants:
            raise ValueError('cannot create an invariant_mask without some '
                             'invariants')
        nword = len(self.words)
        mask = [0] * nword
        for key in invariants:
            word, bit, nbits = self._fields[key][:3]
            if nbits == 64:
                mask[word] = mask[word + 1] = 0xffffffff
            else:
                mask[word] |= (((1 << nbits) - 1) << bit) & 0xffffffff
        pattern = [int(w) & 0xffffffff for w in self.words]
        return pattern, mask

    # -- dict-like ------------------------------------------------------------
    def keys(self):
        return self._fields.keys()

    def __contains__(self, key):
        return key in self._fields

    def __getitem__(self, key):
        try:
            return self._fields.get_field(self.words, key)
        except KeyError:
            raise KeyError('{0} header does not contain {1}'
                           .format(type(self).__name__, key))

    def __setitem__(self, key, value):
        if key not in self._fields:
            raise KeyError('{0} header does not contain {1}'
                           .format(type(self).__name__, key))
        if not self.mutable or (isinstance(self.words, np.ndarray)
                                and not self.words.flags['WRITEABLE']):
            raise TypeError("header is immutable. Set '.mutable` attribute "
                            "or make a copy.")
        self._fields.set_field(self.words, key, value)

    def __eq__(self, other):
        return (type(self) is type(other)
                and np.all(np.asarray(self.words) == np.asarray(other.words)))

    @property
    def nbytes(self):
        return self._struct.size

    # -- I/O -----------------------------------------------------------------
    @classmethod
    def fromfile(cls, fh, *args, **kwargs):
        raw = fh.read(cls._struct.size)
        if len(raw) != cls._struct.size:
            raise EOFError('could not read full header.')
        return cls(cls._struct.unpack(raw), *args, **kwargs)

    def tofile(self, fh):
        return fh.write(self._struct.pack(*[int(w) for w in self.words]))

    @classmethod
    def fromkeys(cls, *args, **kwargs):
        verify = kwargs.pop('verify', True)
        self = cls(None, *args, verify=False)
        for key in self.keys():
            self[key] = kwargs.pop(key)
        if kwargs:
            raise KeyError('unknown header keys: {}'.format(sorted(kwargs)))
        if verify:
            self.verify()
        return self

    @classmethod
    def fromvalues(cls, *args, **kwargs):
        """Defaults first, then explicit keys, then properties in class
        order (baseband/base/header.py:394-430)."""
        verify = kwargs.pop('verify', True)
        self = cls(None, *args, verify=False)
        for key in self.keys():
            default = self._fields.default(key)
            if default is not None and key not in kwargs:
                self[key] = default
        self.update(verify=verify, **kwargs)
        return self

    def update(self, *, verify=True, **kwargs):
        for key in [k for k in kwargs if k in self._fields]:
            self[key] = kwargs.pop(key)
        for prop in self._properties:
            if prop in kwargs:
                setattr(self, prop, kwargs.pop(prop))
        if kwargs:
            raise KeyError('{} does not know how to set {}'.format(
                type(self).__name__, sorted(kwargs)))
        if verify:
            self.verify()

    def __repr__(self):
        return '<{} {}>'.format(type(self).__name__, ', '.join(
            '{}: {}'.format(k, self[k]) for k in self.keys()
            if not k.startswith('_')))
