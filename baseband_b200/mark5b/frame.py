"""Mark 5B frames.  A frame is invalid iff its whole payload equals the fill
pattern 0x11223344 (baseband/mark5b/frame.py:62-72); writing an invalid frame
emits that pattern (:126-133)."""
import numpy as np

from ..base.frame import FrameBase
from .header import Mark5BHeader
from .payload import Mark5BPayload

__all__ = ['Mark5BFrame', 'FILL_PATTERN']

FILL_PATTERN = 0x11223344


class Mark5BFrame(FrameBase):
    _header_class = Mark5BHeader
    _payload_class = Mark5BPayload

    def __init__(self, header, payload, valid=None, verify=True):
        if valid is None:
            w = payload.words
            valid = not (w[0] == FILL_PATTERN and w[1] == FILL_PATTERN
                         and w[2] == FILL_PATTERN
                         and bool(np.all(w[3:] == FILL_PATTERN)))
        super().__init__(header, payload, valid=valid, verify=verify)

    def verify(self):
        assert isinstance(self.header, Mark5BHeader)
        assert isinstance(self.payload, Mark5BPayload)
        assert self.payload.nbytes == 10000

    @classmethod
    def fromfile(cls, fh, kday=None, ref_time=None, nchan=1, bps=2,
                 valid=None, verify=True):
        header = Mark5BHeader.fromfile(fh, kday=kday, ref_time=ref_time,
                                       verify=verify)
        payload = Mark5BPayload.fromfile(fh, sample_shape=(nchan,), bps=bps)
        return cls(header, payload, valid=valid, verify=verify)

    def tofile(self, fh):
        self.header.tofile(fh)
        if self.valid:
            self.payload.tofile(fh)
        else:
            fh.write(np.full(2500, FILL_PATTERN, '<u4').tobytes())

    @classmethod
    def fromdata(cls, data, header=None, bps=2, valid=True, verify=True,
                 **kwargs):
        if header is None:
            header = Mark5BHeader.fromvalues(verify=verify, **kwargs)
        payload = Mark5BPayload.fromdata(data, bps=bps)
        return cls(header, payload, valid=valid, verify=verify)
