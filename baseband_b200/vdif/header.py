"""VDIF headers: 4 (legacy) or 8 little-endian 32-bit words.

Field layout and derived quantities follow baseband/vdif/header.py:529-542
(base words 0-3), :557-559 (edv), :595-598 (sample rate + sync pattern,
EDV 1/3), :701-702, :715-725, :762-770, :792-797, and the properties at
:293-364.  Times use baseband_b200.timeutil (no astropy).
"""
from ..base.header import (BitFieldHeader, FieldTable, four_word_struct,
                           eight_word_struct)
from ..timeutil import Time, as_time, ymd_to_mjd

__all__ = ['VDIFHeader', 'VDIFLegacyHeader', 'VDIFBaseHeader', 'VDIFHeader0',
           'VDIFHeader1', 'VDIFHeader2', 'VDIFHeader3', 'VDIFMark5BHeader',
           'VDIF_HEADER_CLASSES']

BASE_FIELDS = FieldTable((
    ('invalid_data', (0, 31, 1, False)),
    ('legacy_mode', (0, 30, 1, True)),
    ('seconds', (0, 0, 30)),
    ('_1_30_2', (1, 30, 2, 0x0)),
    ('ref_epoch', (1, 24, 6)),
    ('frame_nr', (1, 0, 24, 0x0)),
    ('vdif_version', (2, 29, 3, 0x1)),
    ('lg2_nchan', (2, 24, 5)),
    ('frame_length', (2, 0, 24, 0x80)),
    ('complex_data', (3, 31, 1)),
    ('bits_per_sample', (3, 26, 5)),
    ('thread_id', (3, 16, 10, 0x0)),
    ('station_id', (3, 0, 16))))

VDIF_HEADER_CLASSES = {}


def header_class_for(edv, default=None):
    """Class for an extended-data version; ``False`` means a legacy header
    (looked up by identity: ``False == 0`` would clash with EDV 0)."""
    if edv is False:
        return VDIF_HEADER_CLASSES['legacy']
    return VDIF_HEADER_CLASSES.get(int(edv), default)


def ref_epoch_start(ref_epoch):
    """Half-year epochs counted from 2000-01-01 (VDIF spec section 5)."""
    return Time(ymd_to_mjd(2000 + ref_epoch // 2, 1 if ref_epoch % 2 == 0
                           else 7, 1))


class VDIFHeader(BitFieldHeader):
    """Generic VDIF header; instantiating it picks the class for the EDV
    found in the words (as baseband/vdif/header.py:125-143)."""

    _fields = BASE_FIELDS
    _struct = eight_word_struct
    _edv = None
    _properties = ('frame_nbytes', 'payload_nbytes', 'bps', 'nchan',
                   'samples_per_frame', 'station', 'time')

    def __new__(cls, words=None, edv=None, verify=True, **kwargs):
        if cls is VDIFHeader or cls is VDIFBaseHeader:
            if edv is None:
                if words is None:
                    raise ValueError('need words or edv to pick a class')
                edv = (False if (words[0] >> 30) & 1
                       else (words[4] >> 24) & 0xff)
            cls = header_class_for(edv, VDIFBaseHeader)
        return super().__new__(cls)

    def __init__(self, words=None, edv=None, verify=True, **kwargs):
        if words is None:
            self.words = [0] * (self._struct.size // 4)
        else:
            self.words = words
        if edv is not None and self._edv is not None:
            pass
        if verify:
            self.verify()

    def copy(self, **kwargs):
        kwargs.setdefault('verify', False)
        return self.__class__(list(self.words), **kwargs)

    # ---------------------------------------------------------- factories
    @classmethod
    def fromfile(cls, fh, edv=None, verify=True):
        raw = fh.read(16)
        if len(raw) != 16:
            raise EOFError('could not read full header.')
        words = four_word_struct.unpack(raw)
        if not (words[0] >> 30) & 1:              # not legacy: 4 more words
            more = fh.read(16)
            if len(more) != 16:
                raise EOFError('could not read full header.')
            words = words + four_word_struct.unpack(more)
            found = (words[4] >> 24) & 0xff
        else:
            found = False
        if edv is not None and verify:
            assert edv == found, 'unexpected EDV {} (wanted {})'.format(
                found, edv)
        return cls(words, edv=found, verify=verify)

    @classmethod
    def fromvalues(cls, edv=False, *, verify=True, **kwargs):
        klass = header_class_for(edv)
        if klass is None:
            raise ValueError('no VDIF header class for EDV {}'.format(edv))
        self = klass(None, verify=False)
        for key in self.keys():
            default = self._fields.default(key)
            if default is not None and key not in kwargs:
                self[key] = default
        if edv is not False:
            self['edv'] = edv
        # sizes can be given in several equivalent ways
        kwargs.setdefault('nchan', 1)
        time = kwargs.pop('time', None)
        sample_rate = kwargs.pop('sample_rate', None)
        frame_rate = kwargs.pop('frame_rate', None)
        ordered = ['bps', 'complex_data', 'nchan', 'frame_nbytes',
                   'payload_nbytes', 'samples_per_frame', 'station']
        for key in [k for k in list(kwargs) if k in self._fields]:
            self[key] = kwargs.pop(key)
        for key in ordered:
            if key in kwargs:
                setattr(self, key, kwargs.pop(key))
        if sample_rate is not None and hasattr(type(self), 'sample_rate'):
            self.sample_rate = sample_rate
        if time is not None:
            if frame_rate is None and sample_rate is not None:
                frame_rate = sample_rate / self.samples_per_frame
            self.set_time(time, frame_rate=frame_rate)
        if kwargs:
            raise KeyError('cannot use {} to set up a VDIF header'.format(
                sorted(kwargs)))
        if verify:
            self.verify()
        return self

    # --------------------------------------------------------- properties
    @property
    def edv(self):
        return self._edv

    @property
    def nbytes(self):
        return self._struct.size

    @property
    def frame_nbytes(self):
        return self['frame_length'] * 8

    @frame_nbytes.setter
    def frame_nbytes(self, nbytes):
        assert nbytes % 8 == 0
        self['frame_length'] = int(nbytes) // 8

    @property
    def payload_nbytes(self):
        return self.frame_nbytes - self.nbytes

    @payload_nbytes.setter
    def payload_nbytes(self, nbytes):
        self.frame_nbytes = nbytes + self.nbytes

    @property
    def bps(self):
        return self['bits_per_sample'] + 1

    @bps.setter
    def bps(self, bps):
        assert bps % 1 == 0
        self['bits_per_sample'] = int(bps) - 1

    @property
    def complex_data(self):
        return bool(self['complex_data'])

    @complex_data.setter
    def complex_data(self, value):
        self['complex_data'] = bool(value)

    @property
    def nchan(self):
        return 2 ** self['lg2_nchan']

    @nchan.setter
    def nchan(self, nchan):
        lg2 = int(nchan).bit_length() - 1
        assert 2 ** lg2 == nchan, 'nchan must be a power of two'
        self['lg2_nchan'] = lg2

    @property
    def sample_shape(self):
        return (self.nchan,)

    @property
    def samples_per_frame(self):
        values_per_word = 32 // self.bps // (2 if self.complex_data else 1)
        return self.payload_nbytes // 4 * values_per_word // self.nchan

    @samples_per_frame.setter
    def samples_per_frame(self, samples_per_frame):
        values_per_word = 32 // self.bps // (2 if self.complex_data else 1)
        values = samples_per_frame * self.nchan
        assert values % values_per_word == 0
        self.payload_nbytes = values // values_per_word * 4

    @property
    def station(self):
        sid = self['station_id']
        hi, lo = sid >> 8, sid & 0xff
        if 48 <= hi < 128 and 48 <= lo < 128:
            return chr(hi) + chr(lo)
        return sid

    @station.setter
    def station(self, station):
        if isinstance(station, str):
            assert len(station) == 2
            station = (ord(station[0]) << 8) + ord(station[1])
        self['station_id'] = station

    # --------------------------------------------------------------- time
    def get_time(self, frame_rate=None):
        frame_nr = self['frame_nr']
        if frame_nr == 0:
            offset = 0
        else:
            if frame_rate is None:
                frame_rate = getattr(self, 'frame_rate', None)
            if frame_rate is None:
                raise ValueError('this header does not provide a frame rate; '
                                 'pass it in explicitly.')
            from fractions import Fraction
            offset = Fraction(frame_nr) / Fraction(frame_rate
                                                   ).limit_denominator(10**9)
        return ref_epoch_start(self['ref_epoch']) + self['seconds'] + offset

    def set_time(self, time, frame_rate=None):
        time = as_time(time)
        year = time.year
        ref_epoch = 2 * (year - 2000) + (1 if time.yday > (
            ymd_to_mjd(year, 7, 1) - ymd_to_mjd(year, 1, 1)) else 0)
        seconds = time - ref_epoch_start(ref_epoch)
        whole = int(seconds)
        frac = seconds - whole
        if frac == 0:
            frame_nr = 0
        else:
            if frame_rate is None:
                frame_rate = getattr(self, 'frame_rate', None)
            if frame_rate is None:
                raise ValueError('cannot set a fractional-second time '
                                 'without a frame rate.')
            frame_nr = int(round(float(frac) * float(frame_rate)))
            if frame_nr == int(round(float(frame_rate))):
                whole, frame_nr = whole + 1, 0
        self['ref_epoch'] = ref_epoch
        self['seconds'] = whole
        self['frame_nr'] = frame_nr

    time = property(get_time, set_time)

    def verify(self):
        pass

    # Header parts that do not change within one stream
    # (vdif/header.py:109-115, :561-566, :600-605).
    _stream_invariants = ('legacy_mode', 'vdif_version', 'lg2_nchan',
                          'frame_length', 'complex_data', 'bits_per_sample',
                          'station_id', 'edv', 'sync_pattern',
                          'sampling_unit', 'sampling_rate')

    def invariants(self):
        return {key for key in self._stream_invariants if key in self.keys()}

    def same_stream(self, other):
        """Whether ``other`` can be a header of the same stream
        (vdif/header.py:153-155)."""
        return all(key in other.keys() and self[key] == other[key]
                   for key in self.invariants())

    @classmethod
    def from_mark5b_header(cls, mark5b_header, bps, nchan, **kwargs):
        """Mark 5B-over-VDIF header (EDV 0xab) for a Mark 5B header
        (vdif/header.py:246-288): the time code carries the whole seconds,
        ``frame_nr`` and the BCD fraction are taken over unchanged."""
        assert 'time' not in kwargs, 'Time is inferred from Mark 5B Header.'
        from ..timeutil import Time
        for key in mark5b_header.keys():
            kwargs.setdefault(key, mark5b_header[key])
        kwargs.pop('sync_pattern', None)
        frame_nr = kwargs.pop('frame_nr')
        whole = Time(mark5b_header.kday + mark5b_header.jday,
                     mark5b_header.seconds)
        self = cls.fromvalues(edv=0xab, bps=bps, nchan=nchan,
                              complex_data=False, time=whole, **kwargs)
        self.mutable = True
        self['frame_nr'] = frame_nr
        self['mark5b_frame_nr'] = frame_nr
        self['bcd_fraction'] = mark5b_header['bcd_fraction']
        return self


class VDIFLegacyHeader(VDIFHeader):
    _struct = four_word_struct
    _edv = False

    def verify(self):
        assert self['legacy_mode']
        assert len(self.words) == 4
        assert self['frame_length'] >= 2


class VDIFBaseHeader(VDIFHeader):
    _fields = BASE_FIELDS | FieldTable((
        ('legacy_mode', (0, 30, 1, False)),
        ('edv', (4, 24, 8))))

    @property
    def edv(self):
        return self._edv if self._edv is not None else self['edv']

    def verify(self):
        assert not self['legacy_mode']
        assert self._edv is None or self._edv == self['edv']
        assert len(self.words) == 8
        assert self['frame_length'] >= 4
        sync = self._fields.default('sync_pattern') \
            if 'sync_pattern' in self._fields else None
        if sync is not None:
            assert self['sync_pattern'] == sync


class VDIFHeader0(VDIFBaseHeader):
    _edv = 0

    def verify(self):
        super().verify()
        assert all(w == 0 for w in self.words[4:])


class _SampleRateMixin:
    """EDV 1 and 3: per-channel complex sample rate in word 4."""

    @property
    def sample_rate(self):
        rate = self['sampling_rate'] * (1 if self['complex_data'] else 2)
        return rate * (1e6 if self['sampling_unit'] else 1e3)

    @sample_rate.setter
    def sample_rate(self, sample_rate):
        sample_rate = float(sample_rate)
        assert sample_rate % 1 == 0
        complex_rate = sample_rate / (1 if self['complex_data'] else 2)
        mhz = complex_rate % 1e6 == 0
        self['sampling_unit'] = mhz
        if mhz:
            self['sampling_rate'] = int(complex_rate // 1e6)
        else:
            assert complex_rate % 1e3 == 0
            self['sampling_rate'] = int(complex_rate // 1e3)

    @property
    def frame_rate(self):
        return self.sample_rate / self.samples_per_frame

    @frame_rate.setter
    def frame_rate(self, frame_rate):
        self.sample_rate = frame_rate * self.samples_per_frame


_RATE_FIELDS = VDIFBaseHeader._fields | FieldTable((
    ('sampling_unit', (4, 23, 1)),
    ('sampling_rate', (4, 0, 23)),
    ('sync_pattern', (5, 0, 32, 0xACABFEED))))


class VDIFHeader1(_SampleRateMixin, VDIFBaseHeader):
    _edv = 1
    _fields = _RATE_FIELDS | FieldTable((('das_id', (6, 0, 64, 0x0)),))
    _properties = VDIFHeader._properties[:-1] + ('sample_rate', 'frame_rate',
                                                 'time')


class VDIFHeader3(_SampleRateMixin, VDIFBaseHeader):
    _edv = 3
    _fields = _RATE_FIELDS | FieldTable((
        ('frame_length', (2, 0, 24, 629)),
        ('loif_tuning', (6, 0, 32, 0x0)),
        ('_7_28_4', (7, 28, 4, 0x0)),
        ('dbe_unit', (7, 24, 4, 0x0)),
        ('if_nr', (7, 20, 4, 0x0)),
        ('subband', (7, 17, 3, 0x0)),
        ('sideband', (7, 16, 1, False)),
        ('major_rev', (7, 12, 4, 0x0)),
        ('minor_rev', (7, 8, 4, 0x0)),
        ('personality', (7, 0, 8))))
    _properties = VDIFHeader1._properties

    def verify(self):
        super().verify()
        assert self['frame_length'] in (129, 629)


class VDIFHeader2(VDIFBaseHeader):
    _edv = 2
    _fields = VDIFBaseHeader._fields | FieldTable((
        ('complex_data', (3, 31, 1, 0x0)),
        ('bits_per_sample', (3, 26, 5, 0x1)),
        ('pol', (4, 0, 1)),
        ('BL_quadrant', (4, 1, 2)),
        ('BL_correlator', (4, 3, 1)),
        ('sync_pattern', (4, 4, 20, 0xa5ea5)),
        ('PIC_status', (5, 0, 32)),
        ('PSN', (6, 0, 64))))


class VDIFMark5BHeader(VDIFBaseHeader):
    """Mark 5B frames wrapped in VDIF (EDV 0xab): words 4-7 are the Mark 5B
    header, the payload uses the Mark 5B codec."""
    _edv = 0xab
    _fields = VDIFBaseHeader._fields | FieldTable((
        ('frame_length', (2, 0, 24, 1254)),
        ('sync_pattern', (4, 0, 32, 0xABADDEED)),
        ('user', (5, 16, 16)),
        ('internal_tvg', (5, 15, 1)),
        ('mark5b_frame_nr', (5, 0, 15)),
        ('bcd_jday', (6, 20, 12)),
        ('bcd_seconds', (6, 0, 20)),
        ('bcd_fraction', (7, 16, 16)),
        ('crc', (7, 0, 16))))

    @property
    def edv(self):
        return 0xab

    def verify(self):
        assert not self['legacy_mode']
        assert len(self.words) == 8
        assert self['sync_pattern'] == 0xABADDEED
        assert self['frame_length'] == 1254


VDIF_HEADER_CLASSES.update({'legacy': VDIFLegacyHeader, 0: VDIFHeader0,
                            1: VDIFHeader1, 2: VDIFHeader2, 3: VDIFHeader3,
                            0xab: VDIFMark5BHeader})
