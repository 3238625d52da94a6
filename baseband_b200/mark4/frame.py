"""Mark 4 frames.  The header overwrites the first 160 time steps of every
track, so the first ``160 * fanout`` samples of a frame do not exist; they
read as ``fill_value`` and are ignored on writing (baseband/mark4/frame.py:
148-293).  A frame is valid if no track has an error flag set (:78-97)."""
import operator

import numpy as np

from ..base.frame import FrameBase
from .header import Mark4Header
from .payload import Mark4Payload

__all__ = ['Mark4Frame']

_ERRORS = ('time_sync_error', 'internal_clock_error',
           'processor_time_out_error', 'communication_error')


class Mark4Frame(FrameBase):
    _header_class = Mark4Header
    _payload_class = Mark4Payload

    def __init__(self, header, payload, valid=None, verify=True):
        self.header = header
        self.payload = payload
        if valid is not None:
            self.valid = valid
        if verify:
            self.verify()

    def verify(self):
        assert isinstance(self.header, Mark4Header)
        assert isinstance(self.payload, Mark4Payload)
        assert self.payload.nbytes == self.header.payload_nbytes

    @property
    def valid(self):
        flags = self.header[_ERRORS[0]]
        for key in _ERRORS[1:]:
            flags = flags | self.header[key]
        return not np.any(flags)

    @valid.setter
    def valid(self, valid):
        if valid:
            for key in _ERRORS:
                self.header[key] = False
        else:
            self.header['communication_error'] = True

    @classmethod
    def fromfile(cls, fh, ntrack, decade=None, ref_time=None, verify=True):
        header = Mark4Header.fromfile(fh, ntrack, decade=decade,
                                      ref_time=ref_time, verify=verify)
        payload = Mark4Payload.fromfile(fh, header=header)
        return cls(header, payload, verify=verify)

    @classmethod
    def fromdata(cls, data, header=None, verify=True, **kwargs):
        if header is None:
            header = Mark4Header.fromvalues(verify=verify, **kwargs)
        data = np.asanyarray(data)
        assert data.shape[0] == header.samples_per_frame
        start = header.nbytes * 8 // (header.ntrack // header.fanout)
        payload = Mark4Payload.fromdata(data[start:], header=header)
        return cls(header, payload, verify=verify)

    def __len__(self):
        return self.header.samples_per_frame

    @property
    def _nhidden(self):
        return len(self) - len(self.payload)

    def _split(self, item):
        """-> (start, stop, step or None for an int, rest of the index)."""
        rest = ()
        if isinstance(item, tuple):
            item, rest = (item[0], item[1:]) if item else (slice(None), ())
        if isinstance(item, slice):
            start, stop, step = item.indices(len(self))
            assert step > 0, 'cannot deal with negative steps yet.'
            return start, max(start, stop), step, rest
        try:
            index = operator.index(item)
        except Exception:
            raise TypeError('{0} object can only be indexed or sliced.'
                            .format(type(self)))
        if index < 0:
            index += len(self)
        if not 0 <= index < len(self):
            raise IndexError('{0} index out of range.'.format(type(self)))
        return index, index + 1, None, rest

    def __getitem__(self, item=()):
        if isinstance(item, str):
            return self.header[item]
        start, stop, step, rest = self._split(item)
        hidden = self._nhidden
        count = len(range(start, stop, step or 1))
        data = np.full((count,) + tuple(self.sample_shape), self.fill_value,
                       self.dtype)
        if self.valid and stop > hidden:
            # first requested sample that exists in the payload
            skip = 0 if start >= hidden else -(-(hidden - start) // (step or 1))
            first = start + skip * (step or 1)
            if first < stop:
                data[skip:] = self.payload[first - hidden:stop - hidden:
                                           step or 1]
        if step is None:
            data = data[0]
        return data[(Ellipsis,) + rest] if rest else data

    data = property(__getitem__,
                    doc='Full decoded frame, with header part filled in.')

    def __setitem__(self, item, value):
        if isinstance(item, str):
            self.header[item] = value
            return
        start, stop, step, rest = self._split(item)
        hidden = self._nhidden
        if stop <= hidden:
            return
        value = np.asanyarray(value)
        assert value.ndim <= 2
        stride = step or 1
        skip = 0 if start >= hidden else -(-(hidden - start) // stride)
        first = start + skip * stride
        if first >= stop:
            return
        if skip:
            sample_ndim = (len(self.sample_shape) if not rest else
                           np.empty(self.sample_shape)[rest].ndim)
            if value.ndim == 1 + sample_ndim:
                value = value[skip:]
        if step is None:
            target = first - hidden
        else:
            target = slice(first - hidden, stop - hidden, stride)
        self.payload[(target,) + rest if rest else target] = value
