"""GUPPI frames (header + payload); always valid on reading
(baseband/guppi/frame.py:24-28)."""
from ..base.frame import FrameBase
from .header import GUPPIHeader
from .payload import GUPPIPayload

__all__ = ['GUPPIFrame']


class GUPPIFrame(FrameBase):
    _header_class = GUPPIHeader
    _payload_class = GUPPIPayload

    def verify(self):
        assert isinstance(self.header, GUPPIHeader)
        assert isinstance(self.payload, GUPPIPayload)
        assert self.payload.nbytes == self.header.payload_nbytes

    @classmethod
    def fromfile(cls, fh, memmap=True, valid=True, verify=True):
        header = GUPPIHeader.fromfile(fh, verify=verify)
        payload = GUPPIPayload.fromfile(fh, header=header, memmap=memmap)
        return cls(header, payload, valid=valid, verify=verify)

    @classmethod
    def fromdata(cls, data, header=None, *, valid=True, verify=True,
                 **kwargs):
        if header is None:
            header = GUPPIHeader.fromvalues(verify=verify, **kwargs)
        payload = GUPPIPayload.fromdata(data, header=header)
        return cls(header, payload, valid=valid, verify=verify)
