"""Mark 4 payloads: one ``ntrack``-bit word per time step, tracks carrying
sign and magnitude bits of ``nchan`` channels fanned out over 1, 2 or 4
tracks.  Class surface of baseband/mark4/payload.py:303-406; codec table keyed
``(nchan, bps | packed magnitude bits, fanout)`` (:333-342), CUDA backed."""
from collections import namedtuple

import numpy as np

from .. import codecs
from ..base.payload import PayloadBase
from .header import MARK4_DTYPES

__all__ = ['decode_2chan_2bit_fanout4', 'decode_4chan_2bit_fanout4',
           'decode_8chan_2bit_fanout2', 'decode_8chan_2bit_fanout4',
           'decode_16chan_2bit_fanout2_ft', 'encode_2chan_2bit_fanout4',
           'encode_4chan_2bit_fanout4', 'encode_8chan_2bit_fanout2',
           'encode_8chan_2bit_fanout4', 'encode_16chan_2bit_fanout2_ft',
           'Mark4Payload']


# codec callables under their reference names (mark4/payload.py:122-300)
def _export_codecs():
    from .. import codecs
    for table in (codecs.MARK4_DECODERS, codecs.MARK4_ENCODERS):
        for fn in table.values():
            globals()[fn.__name__] = fn


_export_codecs()


class Mark4Payload(PayloadBase):
    _dtype_word = None
    _decoders = codecs.MARK4_DECODERS
    _encoders = codecs.MARK4_ENCODERS
    _sample_shape_maker = namedtuple('SampleShape', 'nchan')

    def __init__(self, words, header=None, *, sample_shape=(1,), bps=2,
                 fanout=1, magnitude_bit=None, complex_data=False):
        if header is not None:
            magbit = header['magnitude_bit']
            bps = 2 if magbit.any() else 1
            ta = header.track_assignment
            if bps == 1 or np.all(magbit[ta] == [False, True]):
                magnitude_bit = None          # standard layout
            else:
                magnitude_bit = int(np.packbits(magbit).view(
                    header.stream_dtype)[0])
            ntrack = header.ntrack
            fanout = header.fanout
            sample_shape = (ntrack // (bps * fanout),)
            self._nbytes = header.payload_nbytes
        else:
            ntrack = sample_shape[0] * bps * fanout
            magnitude_bit = None
        self._dtype_word = MARK4_DTYPES[ntrack]
        self.fanout = fanout
        if complex_data:
            raise ValueError('Mark4 format does not support complex data.')
        super().__init__(words, sample_shape=sample_shape, bps=bps,
                         complex_data=False)
        self._coder = (self.sample_shape.nchan,
                       self.bps if magnitude_bit is None else magnitude_bit,
                       self.fanout)

    @classmethod
    def fromfile(cls, fh, header=None, **kwargs):
        nbytes = header.payload_nbytes
        raw = fh.read(nbytes)
        if len(raw) < nbytes:
            raise EOFError('could not read full payload.')
        return cls(np.frombuffer(raw, dtype=header.stream_dtype), header)

    @classmethod
    def fromdata(cls, data, header):
        data = np.asanyarray(data)
        if data.dtype.kind == 'c':
            raise ValueError('Mark4 format does not support complex data.')
        if tuple(header.sample_shape) != data.shape[1:]:
            raise ValueError('header is for {0} channels but data has {1}'
                             .format(header.nchan, data.shape[-1]))
        words = np.empty(header.payload_nbytes
                         // header.stream_dtype.itemsize,
                         header.stream_dtype)
        self = cls(words, header)
        self[:] = data
        return self

    def todevice(self, device=None):
        """Decoded payload as a CUDA tensor (nsample, nchan) float32."""
        from .. import device as _device
        from .. import kernels, levels
        dev = _device.resolve(device)
        try:
            nchan, fanout, ft = codecs._M4_MODES[self._coder]
        except KeyError:
            raise KeyError(self._coder) from None
        raw = _device.upload(self.words, dev)
        return kernels.mark4_decode_words(raw, self.words.size, nchan, fanout,
                                          ft, levels.sign_magnitude())
