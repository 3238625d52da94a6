"""GSB payloads: headerless int8 bytes, or packed signed nibbles (low nibble
first) for 4-bit rawdump data (baseband/gsb/payload.py:24-53); phased data
are spread over several files per polarisation that interleave in time
blocks (:88-144)."""
from collections import namedtuple

import numpy as np

from .. import codecs
from ..base.payload import PayloadBase

__all__ = ['decode_4bit', 'decode_8bit', 'encode_4bit', 'encode_8bit',
           'GSBPayload']

_Shape1 = namedtuple('SampleShape', 'nchan')
_ShapeN = namedtuple('SampleShape', 'nthread, nchan')


# codec callables under their reference names (gsb/payload.py:17-53)
def decode_4bit(words):
    from .. import codecs
    return codecs.GSB_DECODERS[4](words)


def decode_8bit(words):
    from .. import codecs
    return codecs.GSB_DECODERS[8](words)


def encode_4bit(values):
    from .. import codecs
    return codecs.GSB_ENCODERS[4](values)


def encode_8bit(values):
    from .. import codecs
    return codecs.GSB_ENCODERS[8](values)


class GSBPayload(PayloadBase):
    _decoders = codecs.GSB_DECODERS
    _encoders = codecs.GSB_ENCODERS
    _dtype_word = np.dtype('int8')

    @staticmethod
    def _sample_shape_maker(*args):
        return _Shape1(*args) if len(args) == 1 else _ShapeN(*args)

    @classmethod
    def fromfile(cls, fh, *, payload_nbytes=1 << 22, sample_shape=(1,),
                 bps=4, complex_data=False):
        """``fh``: one file, or ``fh[thread][part]`` for phased data."""
        if hasattr(fh, 'read'):
            raw = fh.read(payload_nbytes)
            if len(raw) < payload_nbytes:
                raise EOFError('could not read full payload.')
            return cls(np.frombuffer(raw, cls._dtype_word),
                       sample_shape=sample_shape, bps=bps,
                       complex_data=complex_data)
        nthread, npart = len(fh), len(fh[0])
        assert nthread == sample_shape[0]
        sample_nbytes, extra = divmod(
            bps * (2 if complex_data else 1) * int(np.prod(sample_shape[1:])),
            8)
        assert extra == 0, ('Full samples do not fit in integer number of '
                            'bytes')
        words = np.empty((npart, payload_nbytes // sample_nbytes, nthread,
                          sample_nbytes), cls._dtype_word)
        for t, parts in enumerate(fh):
            for p, part in enumerate(parts):
                raw = part.read(payload_nbytes)
                if len(raw) < payload_nbytes:
                    raise EOFError('could not read full payload.')
                words[p, :, t] = np.frombuffer(raw, cls._dtype_word).reshape(
                    -1, sample_nbytes)
        return cls(words.ravel(), sample_shape=sample_shape, bps=bps,
                   complex_data=complex_data)

    def tofile(self, fh):
        if hasattr(fh, 'write'):
            return fh.write(self.words.tobytes())
        nthread = len(fh)
        assert nthread == self.sample_shape[0]
        words = self.words.reshape(len(fh[0]), -1, nthread,
                                   self._bpfs // nthread // 8)
        for t, parts in enumerate(fh):
            for p, part in enumerate(parts):
                part.write(np.ascontiguousarray(words[p, :, t]).tobytes())
