"""Mark 4 headers: per-track 160-bit headers stored bit-interleaved.

The first 160 time steps of a frame hold, for every track, five 32-bit header
words, most significant bit first (baseband/mark4/header.py:47-88); the
per-track fields are those of :116-142.  This class keeps the header as a
``(5, ntrack)`` uint32 array; every field access is vectorised over tracks.
Track assignments follow tables 10-14 of the Mark 4 memo 230.3 (as listed at
:306-328 of the reference).
"""
from fractions import Fraction

import numpy as np

from ..base.header import FieldTable
from ..base.utils import bcd_decode, bcd_encode, crc_of_bits
from ..timeutil import Time, as_time

__all__ = ['Mark4Header', 'stream2words', 'words2stream', 'MARK4_DTYPES',
           'HEADER_STEPS', 'FRAME_STEPS', 'CRC12']

MARK4_DTYPES = {8: np.dtype('<u1'), 16: np.dtype('<u2'), 32: np.dtype('<u4'),
                64: np.dtype('<u8')}
HEADER_STEPS = 160
FRAME_STEPS = 20000
CRC12 = 0x180f

TRACK_FIELDS = FieldTable((
    ('bcd_headstack1', (0, 0, 16, 0x3344)),
    ('bcd_headstack2', (0, 16, 16, 0x1122)),
    ('headstack_id', (1, 30, 2)),
    ('bcd_track_id', (1, 24, 6)),
    ('fan_out', (1, 22, 2)),
    ('magnitude_bit', (1, 21, 1)),
    ('lsb_output', (1, 20, 1)),
    ('converter_id', (1, 16, 4)),
    ('time_sync_error', (1, 15, 1, False)),
    ('internal_clock_error', (1, 14, 1, False)),
    ('processor_time_out_error', (1, 13, 1, False)),
    ('communication_error', (1, 12, 1, False)),
    ('_1_11_1', (1, 11, 1, False)),
    ('_1_10_1', (1, 10, 1, False)),
    ('track_roll_enabled', (1, 9, 1, False)),
    ('sequence_suspended', (1, 8, 1, False)),
    ('system_id', (1, 0, 8)),
    ('_1_0_1_sync', (1, 0, 1, 0)),
    ('sync_pattern', (2, 0, 32, 0xffffffff)),
    ('bcd_unit_year', (3, 28, 4)),
    ('bcd_day', (3, 16, 12)),
    ('bcd_hour', (3, 8, 8)),
    ('bcd_minute', (3, 0, 8)),
    ('bcd_second', (4, 24, 8)),
    ('bcd_fraction', (4, 12, 12)),
    ('crc', (4, 0, 12))))


def stream2words(stream, track=None):
    """Track words of 160 (or any multiple of 32) time steps -> uint32 header
    words per track, shape (nword, ntrack); bit 31 comes first in time."""
    stream = np.asarray(stream)
    ntrack = stream.dtype.itemsize * 8
    tracks = (np.arange(ntrack, dtype=stream.dtype) if track is None
              else np.atleast_1d(np.asarray(track, dtype=stream.dtype)))
    bits = ((stream.reshape(-1, 32, 1) >> tracks) & 1).astype(np.uint32)
    bits <<= np.arange(31, -1, -1, dtype=np.uint32).reshape(32, 1)
    words = np.bitwise_or.reduce(bits, axis=1)
    return words if track is None or np.ndim(track) else words[:, 0]


def words2stream(words):
    """Inverse of `stream2words`: (..., nword, ntrack) uint32 -> (..., nword
    * 32) track words."""
    words = np.asarray(words, dtype=np.uint32)
    ntrack = words.shape[-1]
    dtype = MARK4_DTYPES[ntrack]
    shifts = np.arange(31, -1, -1, dtype=np.uint32).reshape(32, 1)
    bits = ((words[..., np.newaxis, :] >> shifts) & 1).astype(dtype)
    bits <<= np.arange(ntrack, dtype=dtype)
    out = np.bitwise_or.reduce(bits, axis=-1)
    return out.reshape(words.shape[:-2] + (-1,))


# (bps, fanout) -> tracks of (fanout, channel, [sign, magnitude]), 32 tracks,
# memo numbering minus 2.
_ASSIGN = {
    (2, 4): np.array([[2, 10, 3, 11, 18, 26, 19, 27],
                      [4, 12, 5, 13, 20, 28, 21, 29],
                      [6, 14, 7, 15, 22, 30, 23, 31],
                      [8, 16, 9, 17, 24, 32, 25, 33]]).reshape(4, 4, 2) - 2,
    (1, 4): np.array([[2, 3, 10, 11, 18, 19, 26, 27],
                      [4, 5, 12, 13, 20, 21, 28, 29],
                      [6, 7, 14, 15, 22, 23, 30, 31],
                      [8, 9, 16, 17, 24, 25, 32, 33]]).reshape(4, 8, 1) - 2,
    (2, 2): np.array([[2, 6, 3, 7, 10, 14, 11, 15, 18, 22, 19, 23, 26, 30,
                       27, 31],
                      [4, 8, 5, 9, 12, 16, 13, 17, 20, 24, 21, 25, 28, 32,
                       29, 33]]).reshape(2, 8, 2) - 2,
    (1, 2): np.array([[2, 3, 6, 7, 10, 11, 14, 15, 18, 19, 22, 23, 26, 27,
                       30, 31],
                      [4, 5, 8, 9, 12, 13, 16, 17, 20, 21, 24, 25, 28, 29,
                       32, 33]]).reshape(2, 16, 1) - 2,
    (2, 1): np.array([[2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28,
                       30, 32, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25,
                       27, 29, 31, 33]]).reshape(1, 16, 2) - 2}


def track_assignment(ntrack, bps, fanout):
    try:
        ta = _ASSIGN[(bps, fanout)]
    except KeyError:
        raise ValueError('Mark 4 reader does not support bps={0}, fanout={1}; '
                         'supported are {2}'.format(bps, fanout,
                                                    list(_ASSIGN)))
    if ntrack == 64:
        return np.concatenate((ta, ta + 32), axis=1)
    if ntrack == 32:
        return ta
    if ntrack == 16:
        return ta[:, ::2, :] // 2
    raise ValueError('have Mark 4 track assignments only for ntrack=16, 32 '
                     'or 64, not {0}'.format(ntrack))


class Mark4Header:
    """All track headers of one frame, as a (5, ntrack) uint32 array."""

    _fields = TRACK_FIELDS
    _properties = ('decade', 'track_id', 'fraction', 'time', 'fanout',
                   'samples_per_frame', 'bps', 'complex_data', 'nchan',
                   'sample_shape', 'nsb', 'converters')
    decade = None
    complex_data = False

    def __init__(self, words, ntrack=None, decade=None, ref_time=None,
                 verify=True):
        if words is None:
            words = np.zeros((5, ntrack), np.uint32)
            verify = False
        self.words = np.asarray(words, dtype=np.uint32)
        if decade is not None:
            self.decade = decade
        if verify:
            self.verify()
        if decade is None and ref_time is not None:
            self.infer_decade(ref_time)

    def verify(self):
        assert self.words.shape[0] == 5
        assert np.all(self['sync_pattern'] == 0xffffffff)
        assert np.all((self['bcd_fraction'] & 0xf) % 5 != 4)
        if self.decade is not None:
            assert 1950 < self.decade < 3000
            assert self.decade % 10 == 0, 'decade must end in zero'
        assert set(self['fan_out'].tolist()) == set(range(self.fanout))
        assert len(set(zip(self['converter_id'].tolist(),
                           self['lsb_output'].tolist()))) == self.nchan

    # ------------------------------------------------------- dict interface
    def keys(self):
        return self._fields.keys()

    def __contains__(self, key):
        return key in self._fields

    def __getitem__(self, key):
        try:
            word, bit, nbits = self._fields[key][:3]
        except KeyError:
            raise KeyError('Mark4Header header does not contain {}'.format(
                key))
        value = (self.words[word] >> np.uint32(bit)) & np.uint32(
            (1 << nbits) - 1)
        return value != 0 if nbits == 1 else value

    def __setitem__(self, key, value):
        if key not in self._fields:
            raise KeyError('Mark4Header header does not contain {}'.format(
                key))
        if not self.mutable:
            raise TypeError("header is immutable. Set '.mutable` attribute "
                            "or make a copy.")
        word, bit, nbits = self._fields[key][:3]
        mask = (1 << nbits) - 1
        if value is None:
            value = self._fields.default(key)
        if value is True:
            value = mask
        value = np.asarray(value).astype(np.int64)
        if np.any((value & mask) != value):
            raise ValueError('{0} cannot be represented with {1} bits'
                             .format(value, nbits))
        keep = np.uint32(~(mask << bit) & 0xffffffff)
        self.words[word] = (self.words[word] & keep) | (
            value.astype(np.uint32) << np.uint32(bit))

    @property
    def mutable(self):
        return self.words.flags['WRITEABLE']

    @mutable.setter
    def mutable(self, mutable):
        self.words.flags['WRITEABLE'] = bool(mutable)

    def copy(self):
        return self.__class__(self.words.copy(), decade=self.decade,
                              verify=False)

    __copy__ = copy

    def __eq__(self, other):
        return (type(self) is type(other)
                and np.array_equal(self.words, other.words))

    # ------------------------------------------------------------------ I/O
    @classmethod
    def fromfile(cls, fh, ntrack, decade=None, ref_time=None, verify=True):
        dtype = MARK4_DTYPES[ntrack]
        nbytes = ntrack * HEADER_STEPS // 8
        raw = fh.read(nbytes)
        if len(raw) != nbytes:
            raise EOFError('could not read full Mark 4 Header.')
        self = cls(stream2words(np.frombuffer(raw, dtype)), decade=decade,
                   ref_time=ref_time, verify=verify)
        self.mutable = False
        return self

    def tofile(self, fh):
        return fh.write(words2stream(self.words).tobytes())

    @classmethod
    def fromvalues(cls, ntrack, decade=None, ref_time=None, *, verify=True,
                   **kwargs):
        """Header from keywords; needs at least ``time``, ``bps`` and
        ``fanout`` (or ``samples_per_frame``)."""
        if ntrack == 64:
            kwargs.setdefault('headstack_id', np.repeat(np.arange(2), 32))
            kwargs.setdefault('track_id', np.tile(np.arange(2, 34), 2))
        elif ntrack == 32:
            kwargs.setdefault('headstack_id', np.zeros(32, int))
            kwargs.setdefault('track_id', np.arange(2, 34))
        elif ntrack == 16:
            kwargs.setdefault('headstack_id', np.zeros(16, int))
            kwargs.setdefault('track_id', np.arange(2, 34, 2))
        if not any(k in kwargs for k in ('lsb_output', 'converter_id',
                                         'converters')):
            kwargs.setdefault('nsb', 1)
        self = cls(None, ntrack=ntrack, decade=decade, ref_time=ref_time)
        for key in self.keys():
            default = self._fields.default(key)
            if default is not None and key not in kwargs:
                self[key] = default
        self.update(verify=verify, **kwargs)
        return self

    def update(self, *, crc=None, verify=True, **kwargs):
        for key in [k for k in kwargs if k in self._fields]:
            self[key] = kwargs.pop(key)
        for prop in self._properties:
            if prop in kwargs:
                setattr(self, prop, kwargs.pop(prop))
        if kwargs:
            raise KeyError('Mark4Header does not know how to set {}'.format(
                sorted(kwargs)))
        if crc is None:
            stream = words2stream(self.words)
            stream[-12:] = crc_of_bits(stream[:-12], CRC12)
            self.words = stream2words(stream)
        else:
            self['crc'] = crc
        if verify:
            self.verify()

    # ------------------------------------------------------------- geometry
    @property
    def ntrack(self):
        return self.words.shape[1]

    @property
    def stream_dtype(self):
        return MARK4_DTYPES[self.ntrack]

    @property
    def nbytes(self):
        return self.ntrack * HEADER_STEPS // 8

    @property
    def frame_nbytes(self):
        return self.ntrack * FRAME_STEPS // 8

    @property
    def payload_nbytes(self):
        return self.frame_nbytes - self.nbytes

    @property
    def fanout(self):
        return int(np.max(self['fan_out'])) + 1

    @fanout.setter
    def fanout(self, fanout):
        if fanout not in (1, 2, 4):
            raise ValueError('Mark 4 data only supports fanout=1, 2, or 4, '
                             'not {0}.'.format(fanout))
        if self.ntrack == 16:
            self['fan_out'] = np.tile(np.arange(fanout),
                                      self.ntrack // fanout)
        else:
            self['fan_out'] = np.tile(np.repeat(np.arange(fanout), 2),
                                      self.ntrack // 2 // fanout)

    @property
    def samples_per_frame(self):
        return self.frame_nbytes * 8 // (self.ntrack // self.fanout)

    @samples_per_frame.setter
    def samples_per_frame(self, samples_per_frame):
        fanout, extra = divmod(samples_per_frame * self.ntrack,
                               8 * self.frame_nbytes)
        if extra or fanout not in (1, 2, 4):
            raise ValueError('header cannot store {} samples per frame.'
                             .format(samples_per_frame))
        self.fanout = int(fanout)

    @property
    def bps(self):
        return 2 if self['magnitude_bit'].any() else 1

    @bps.setter
    def bps(self, bps):
        if bps == 1:
            self['magnitude_bit'] = False
        elif bps == 2:
            ta = track_assignment(self.ntrack, bps, self.fanout)
            magbit = np.empty(self.ntrack, bool)
            magbit[ta] = [False, True]
            self['magnitude_bit'] = magbit
        else:
            raise ValueError('Mark 4 data can only have bps=1 or 2, not {0}'
                             .format(bps))

    @property
    def nchan(self):
        return self.ntrack // (self.fanout * self.bps)

    @nchan.setter
    def nchan(self, nchan):
        self.bps = self.ntrack // (self.fanout * nchan)

    @property
    def sample_shape(self):
        return (self.nchan,)

    @sample_shape.setter
    def sample_shape(self, sample_shape):
        self.nchan, = sample_shape

    @property
    def track_assignment(self):
        return track_assignment(self.ntrack, self.bps, self.fanout)

    @property
    def nsb(self):
        sb = self['lsb_output']
        return 1 if (sb == sb[0]).all() else 2

    @nsb.setter
    def nsb(self, nsb):
        if nsb == 1:
            self['lsb_output'] = True
        elif nsb == 2:
            self['lsb_output'] = np.tile([False, True], self.ntrack // 2)
        else:
            raise ValueError('number of sidebands can only be 1 or 2.')
        nconverter = self.ntrack // (self.fanout * self.bps * self.nsb)
        converters = np.arange(nconverter)
        if nconverter > 2:
            converters = converters.reshape(-1, 2, 2).transpose(
                0, 2, 1).ravel()
        self.converters = converters

    @property
    def converters(self):
        """Per channel: structured array of 'converter' id and 'lsb'."""
        ta = self.track_assignment[0, :, 0]
        out = np.empty(len(ta), [('converter', int), ('lsb', bool)])
        out['converter'] = self['converter_id'][ta]
        out['lsb'] = self['lsb_output'][ta]
        return out

    @converters.setter
    def converters(self, converters):
        ta = self.track_assignment
        nchan = ta.shape[1]
        if isinstance(converters, dict) or getattr(
                getattr(converters, 'dtype', None), 'names', None):
            conv = np.asarray(converters['converter'])
            lsb = np.asarray(converters['lsb'])
        else:
            conv = np.asarray(converters)
            lsb = self['lsb_output'][ta[0, :, 0]]
            if conv.size * 2 == nchan:
                conv = np.repeat(conv, 2)
        conv = np.broadcast_to(conv, (nchan,))
        lsb = np.broadcast_to(lsb, (nchan,))
        cid = np.empty(self.ntrack, int)
        sb = np.empty(self.ntrack, bool)
        cid[ta] = conv[:, np.newaxis]
        sb[ta] = lsb[:, np.newaxis]
        self['converter_id'] = cid
        self['lsb_output'] = sb

    # ----------------------------------------------------------------- time
    def infer_decade(self, ref_time):
        ref = as_time(ref_time)
        year = ref.year + (ref.yday - 1) / 365.25
        decades = np.around(year - self['bcd_unit_year'].astype(float),
                            decimals=-1).astype(int)
        assert np.all(decades == decades[0])
        self.decade = int(decades[0])

    @property
    def track_id(self):
        return bcd_decode(self['bcd_track_id'])

    @track_id.setter
    def track_id(self, track_id):
        self['bcd_track_id'] = bcd_encode(np.asarray(track_id))

    @property
    def ms(self):
        """Milliseconds within the second as exact Fractions per track: the
        last BCD digit d stands for d * 1.25 (0, 5) pattern of memo 230.3."""
        ms = bcd_decode(self['bcd_fraction'])
        return ms * 4 + ms % 5          # in units of 0.25 ms

    @property
    def fraction(self):
        return self.ms / 4000.

    @fraction.setter
    def fraction(self, fraction):
        ms = np.asarray(fraction, float) * 1000.
        if np.any(np.abs(ms / 1.25 - np.around(ms / 1.25)) > 1e-6):
            raise ValueError('{0} ms is not a multiple of 1.25 ms'
                             .format(ms))
        self['bcd_fraction'] = bcd_encode(np.floor(ms + 1e-6).astype(
            np.int64))

    def get_time(self):
        if self.decade is None:
            raise ValueError('need decade or ref_time to get a full time.')
        t0 = lambda k: int(np.atleast_1d(self[k])[0])  # noqa: E731
        for key in ('bcd_unit_year', 'bcd_day', 'bcd_hour', 'bcd_minute',
                    'bcd_second', 'bcd_fraction'):
            v = np.atleast_1d(self[key])
            assert np.all(v == v[0]), 'tracks disagree on ' + key
        year = self.decade + bcd_decode(t0('bcd_unit_year'))
        sec = (bcd_decode(t0('bcd_hour')) * 3600
               + bcd_decode(t0('bcd_minute')) * 60
               + bcd_decode(t0('bcd_second'))
               + Fraction(int(np.atleast_1d(self.ms)[0]), 4000))
        return Time.from_yday(year, bcd_decode(t0('bcd_day')), sec)

    def set_time(self, time):
        time = as_time(time)
        year = time.year
        whole = int(time.sec)
        self.fraction = float(time.sec - whole)
        self.decade = year // 10 * 10
        self['bcd_unit_year'] = bcd_encode(year % 10)
        self['bcd_day'] = bcd_encode(time.yday)
        self['bcd_hour'] = bcd_encode(whole // 3600)
        self['bcd_minute'] = bcd_encode(whole // 60 % 60)
        self['bcd_second'] = bcd_encode(whole % 60)

    time = property(get_time, set_time)

    def __repr__(self):
        return '<Mark4Header ntrack={} fanout={} bps={} time={}>'.format(
            self.ntrack, self.fanout, self.bps,
            self.time.isot if self.decade is not None else '?')
