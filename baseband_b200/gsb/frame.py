"""GSB frames: a timestamp line plus payload bytes from one or more raw
files; always valid on reading (baseband/gsb/frame.py:30-34)."""
from ..base.frame import FrameBase
from .header import GSBHeader
from .payload import GSBPayload

__all__ = ['GSBFrame']


class GSBFrame(FrameBase):
    _header_class = GSBHeader
    _payload_class = GSBPayload

    def verify(self):
        assert isinstance(self.header, GSBHeader)
        assert isinstance(self.payload, GSBPayload)

    @classmethod
    def fromfile(cls, fh_ts, fh_raw, payload_nbytes=1 << 24, sample_shape=(1,),
                 bps=4, complex_data=False, valid=True, verify=True):
        header = GSBHeader.fromfile(fh_ts, verify=verify)
        payload = GSBPayload.fromfile(fh_raw, payload_nbytes=payload_nbytes,
                                      sample_shape=sample_shape, bps=bps,
                                      complex_data=complex_data)
        return cls(header, payload, valid=valid, verify=verify)

    def tofile(self, fh_ts, fh_raw):
        self.header.tofile(fh_ts)
        self.payload.tofile(fh_raw)

    @classmethod
    def fromdata(cls, data, header=None, bps=4, valid=True, verify=True,
                 **kwargs):
        if header is None:
            header = GSBHeader.fromvalues(verify=verify, **kwargs)
        payload = GSBPayload.fromdata(data, bps=bps)
        return cls(header, payload, valid=valid, verify=verify)

    @property
    def nbytes(self):
        return self.payload.nbytes
