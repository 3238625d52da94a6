"""GUPPI raw-file headers: 80-column ASCII cards.

Derived quantities follow baseband/guppi/header.py:216-352: payload size
``BLOCSIZE``; ``NPOL`` counts real and imaginary parts, so complex data
(``OBSNCHAN != 1``) have ``npol = NPOL // 2``; bits per complete sample
``OBSNCHAN * NPOL * NBITS``; ``OVERLAP`` samples at the end of each frame
repeat the start of the next; ``PKTFMT != 'SIMPLE'`` means the payload is
stored channels first; ``DIRECTIO`` headers are padded to 512 bytes; the time
of a frame is ``STT_IMJD/STT_SMJD/STT_OFFS`` plus ``PKTIDX`` packets.
"""
import operator
from fractions import Fraction

from ..base.cards import CardHeader
from ..timeutil import Time, as_time

__all__ = ['GUPPIHeader']


class GUPPIHeader(CardHeader):
    _properties = ('payload_nbytes', 'frame_nbytes', 'bps', 'complex_data',
                   'sample_shape', 'sample_rate', 'sideband',
                   'samples_per_frame', 'overlap', 'offset', 'start_time',
                   'time')
    _defaults = [('BACKEND', 'GUPPI'), ('BLOCSIZE', 0), ('PKTIDX', 0),
                 ('STT_OFFS', 0), ('OVERLAP', 0), ('SRC_NAME', 'unset'),
                 ('TELESCOP', 'unset'), ('PKTFMT', '1SFA'),
                 ('PKTSIZE', 8192), ('NBITS', 8), ('NPOL', 1),
                 ('OBSNCHAN', 1)]

    def __init__(self, cards=None, verify=True, mutable=True):
        super().__init__(cards)
        self.mutable = mutable
        if len(self) and verify:
            self.verify()

    def verify(self):
        assert all(key in self for key in ('BLOCSIZE', 'PKTIDX'))

    def __setitem__(self, key, value):
        if not self.mutable:
            raise TypeError('immutable {0} does not support assignment.'
                            .format(type(self).__name__))
        super().__setitem__(key, value)

    def copy(self):
        new = type(self)(verify=False)
        new._values = dict(self._values)
        new._text = dict(self._text)
        return new

    __copy__ = copy

    def __eq__(self, other):
        return all(self.get(k) == other.get(k)
                   for k in set(self.keys()) | set(other.keys()))

    # ------------------------------------------------------------------ I/O
    @classmethod
    def fromfile(cls, fh, verify=True):
        start = fh.tell()
        ncard = 0
        while True:
            card = fh.read(80)
            if len(card) < 80:
                raise EOFError('could not read full GUPPI header.')
            ncard += 1
            if card[:3] == b'END':
                break
            if card[8:9] not in (b'=', b' '):
                raise OSError('not a GUPPI header card: {!r}'.format(card))
        fh.seek(start)
        text = fh.read(80 * ncard).decode('ascii')
        self = cls.parse(text)
        self.mutable = True
        if verify:
            self.verify()
        fh.seek(start + self.nbytes)        # skip DIRECTIO padding
        self.mutable = False
        return self

    def tofile(self, fh):
        raw = self.tostring().encode('ascii')
        raw += b'\0' * (self.nbytes - len(raw))
        return fh.write(raw)

    @classmethod
    def fromkeys(cls, *args, verify=True, mutable=True, **kwargs):
        self = cls(verify=False, mutable=True)
        for key, value in kwargs.items():
            self[key] = value
        self.mutable = mutable
        if verify:
            self.verify()
        return self

    @classmethod
    def fromvalues(cls, verify=True, mutable=True, **kwargs):
        self = cls(cls._defaults, verify=False, mutable=True)
        self.update(verify=verify, **kwargs)
        self.mutable = mutable
        return self

    def update(self, *, verify=True, **kwargs):
        extras = [(p, kwargs.pop(p)) for p in self._properties
                  if p in kwargs]
        if 'sample_shape' in dict(extras):
            # complex-ness depends on nchan: set the shape first
            extras.sort(key=lambda kv: kv[0] != 'sample_shape')
        for key, value in kwargs.items():
            self[key] = value
        for attr, value in extras:
            setattr(self, attr, value)
        if verify:
            self.verify()

    # ------------------------------------------------------------- geometry
    @property
    def nbytes(self):
        nbytes = (len(self) + 1) * 80
        if int(self.get('DIRECTIO', 0)) and nbytes % 512:
            nbytes += 512 - nbytes % 512
        return nbytes

    @property
    def payload_nbytes(self):
        return int(self['BLOCSIZE'])

    @payload_nbytes.setter
    def payload_nbytes(self, nbytes):
        self['BLOCSIZE'] = int(nbytes)

    @property
    def frame_nbytes(self):
        return self.nbytes + self.payload_nbytes

    @frame_nbytes.setter
    def frame_nbytes(self, nbytes):
        self.payload_nbytes = nbytes - self.nbytes

    @property
    def bps(self):
        return int(self['NBITS'])

    @bps.setter
    def bps(self, bps):
        self['NBITS'] = int(bps)

    @property
    def complex_data(self):
        return int(self['OBSNCHAN']) != 1

    @property
    def npol(self):
        return int(self['NPOL']) // (2 if self.complex_data else 1)

    @npol.setter
    def npol(self, npol):
        self['NPOL'] = int(npol) * (2 if self.complex_data else 1)

    @property
    def nchan(self):
        return int(self['OBSNCHAN'])

    @nchan.setter
    def nchan(self, nchan):
        self['OBSNCHAN'] = operator.index(nchan)

    @property
    def sample_shape(self):
        return self.npol, self.nchan

    @sample_shape.setter
    def sample_shape(self, sample_shape):
        self.nchan = sample_shape[1]
        self.npol = sample_shape[0]

    @property
    def _bpcs(self):
        return int(self['OBSNCHAN']) * int(self['NPOL']) * self.bps

    @property
    def sample_rate(self):
        return 1. / float(self['TBIN'])

    @sample_rate.setter
    def sample_rate(self, sample_rate):
        to_value = getattr(sample_rate, 'to_value', None)
        rate = float(to_value('Hz')) if to_value else float(sample_rate)
        self['TBIN'] = 1. / abs(rate)
        self['OBSBW'] = (rate / 1e6 * int(self['OBSNCHAN'])
                         / (1 if self.complex_data else 2))

    @property
    def sideband(self):
        return float(self['OBSBW']) > 0

    @sideband.setter
    def sideband(self, sideband):
        self['OBSBW'] = (1 if sideband else -1) * abs(self['OBSBW'])

    @property
    def channels_first(self):
        return self['PKTFMT'] != 'SIMPLE'

    @channels_first.setter
    def channels_first(self, channels_first):
        self['PKTFMT'] = '1SFA' if channels_first else 'SIMPLE'

    @property
    def samples_per_frame(self):
        return self.payload_nbytes * 8 // self._bpcs

    @samples_per_frame.setter
    def samples_per_frame(self, samples_per_frame):
        old = self.payload_nbytes
        self.payload_nbytes = (samples_per_frame * self._bpcs + 7) // 8
        if self.samples_per_frame != samples_per_frame:
            nearest = self.samples_per_frame
            self.payload_nbytes = old
            raise ValueError('header cannot store {} samples per frame. '
                             'Nearest is {}.'.format(samples_per_frame,
                                                     nearest))

    @property
    def overlap(self):
        return int(self['OVERLAP'])

    @overlap.setter
    def overlap(self, overlap):
        self['OVERLAP'] = operator.index(overlap)

    # ----------------------------------------------------------------- time
    def _tbin(self):
        return Fraction(float(self['TBIN'])).limit_denominator(10**15)

    @property
    def offset(self):
        """Seconds since `start_time` (PKTIDX counts non-overlap packets)."""
        nsample = int(self['PKTIDX']) * int(self['PKTSIZE']) * 8 // self._bpcs
        return nsample * self._tbin()

    @offset.setter
    def offset(self, offset):
        self['PKTIDX'] = int(round(
            Fraction(offset) / self._tbin() / int(self['PKTSIZE'])
            * ((self._bpcs + 7) // 8)))

    @property
    def start_time(self):
        return Time(int(self['STT_IMJD']), Fraction(int(self['STT_SMJD']))
                    + Fraction(float(self['STT_OFFS'])).limit_denominator(
                        10**12))

    @start_time.setter
    def start_time(self, start_time):
        start_time = as_time(start_time)
        whole = int(start_time.sec)
        self['STT_IMJD'] = start_time.mjd
        self['STT_SMJD'] = whole
        frac = start_time.sec - whole
        self['STT_OFFS'] = float(frac) if frac else 0

    @property
    def time(self):
        return self.start_time + self.offset

    @time.setter
    def time(self, time):
        time = as_time(time)
        if 'STT_IMJD' not in self:
            self.start_time = time - self.offset
        else:
            self.offset = time - self.start_time
