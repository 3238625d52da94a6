"""DADA frames; always valid on reading (baseband/dada/frame.py:24-28)."""
from ..base.frame import FrameBase
from .header import DADAHeader
from .payload import DADAPayload

__all__ = ['DADAFrame']


class DADAFrame(FrameBase):
    _header_class = DADAHeader
    _payload_class = DADAPayload

    def verify(self):
        assert isinstance(self.header, DADAHeader)
        assert isinstance(self.payload, DADAPayload)
        assert self.payload.nbytes == self.header.payload_nbytes

    @classmethod
    def fromfile(cls, fh, memmap=True, valid=True, verify=True):
        header = DADAHeader.fromfile(fh, verify=verify)
        payload = DADAPayload.fromfile(fh, header=header, memmap=memmap)
        return cls(header, payload, valid=valid, verify=verify)

    @classmethod
    def fromdata(cls, data, header=None, *, valid=True, verify=True,
                 **kwargs):
        if header is None:
            header = DADAHeader.fromvalues(verify=verify, **kwargs)
        payload = DADAPayload.fromdata(data, header=header)
        return cls(header, payload, valid=valid, verify=verify)
