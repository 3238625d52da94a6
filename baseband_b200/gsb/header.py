"""GSB timestamp-file headers: one text line per frame.

Rawdump lines hold one time (``YYYY MM DD HH MM SS 0.SSSSSSSSS``); phased
lines hold the PC time, the GPS time, a sequence number and a memory block
number (baseband/gsb/header.py:23-75, :243-361).  Times in the file are
Indian Standard Time: ``utc_offset`` (5.5 h) is subtracted.
"""
from fractions import Fraction

from ..timeutil import Time, as_time, ymd_to_mjd

__all__ = ['GSBHeader', 'GSBRawdumpHeader', 'GSBPhasedHeader']


def parse_gsb_time(words):
    y, mo, d, h, mi, s = (int(w) for w in words[:6])
    frac = Fraction(words[6])
    return Time(ymd_to_mjd(y, mo, d), h * 3600 + mi * 60 + s + frac)


def format_gsb_time(time, precision):
    from ..timeutil import mjd_to_ymd
    scaled = int(round(time.sec * 10 ** precision))
    mjd = time.mjd
    if scaled >= 86400 * 10 ** precision:
        scaled -= 86400 * 10 ** precision
        mjd += 1
    whole, frac = divmod(scaled, 10 ** precision)
    y, mo, d = mjd_to_ymd(mjd)
    return '{:04d} {:02d} {:02d} {:02d} {:02d} {:02d} 0.{:0{p}d}'.format(
        y, mo, d, whole // 3600, whole // 60 % 60, whole % 60, frac,
        p=precision).split()


class GSBHeader:
    _mode = None
    _classes = {}
    _gps_precision = 9
    _pc_precision = 6

    def __new__(cls, words=None, mode=None, nbytes=None,
                utc_offset=5.5 * 3600, verify=True):
        if cls is GSBHeader:
            if mode is None:
                if words is None:
                    raise TypeError('cannot construct a GSB header without '
                                    'knowing the mode.')
                mode = 'rawdump' if len(words) == 7 else 'phased'
            cls = cls._classes[mode]
        return super().__new__(cls)

    def __init__(self, words=None, mode=None, nbytes=None,
                 utc_offset=5.5 * 3600, verify=True):
        if words is None:
            self.words = [''] * self._nwords
        else:
            self.words = words
        self._nbytes = nbytes
        self.utc_offset = utc_offset
        if verify and words is not None:
            self.verify()

    def verify(self):
        assert len(self.words) == self._nwords

    mode = property(lambda self: self._mode)

    @property
    def mutable(self):
        return isinstance(self.words, list)

    @mutable.setter
    def mutable(self, mutable):
        self.words = list(self.words) if mutable else tuple(self.words)

    def copy(self):
        return type(self)(list(self.words), nbytes=self._nbytes,
                          utc_offset=self.utc_offset, verify=False)

    @property
    def nbytes(self):
        if self._nbytes is None:
            return len(' '.join(self.words)) + 1
        return self._nbytes

    def __eq__(self, other):
        return type(self) is type(other) and list(self.words) == list(
            other.words)

    # ------------------------------------------------------------------ I/O
    @classmethod
    def fromfile(cls, fh, **kwargs):
        start = fh.tell()
        line = fh.readline()
        if isinstance(line, bytes):
            line = line.decode('ascii')
        if line == '':
            raise EOFError
        return cls(tuple(line.split()), nbytes=fh.tell() - start, **kwargs)

    def tofile(self, fh):
        text = ' '.join(self.words) + '\n'
        try:
            return fh.write(text)
        except TypeError:
            return fh.write(text.encode('ascii'))

    @classmethod
    def fromvalues(cls, mode=None, nbytes=None, **kwargs):
        if mode is None and cls._mode is None:
            if set(kwargs) & {'pc', 'pc_time', 'seq_nr', 'mem_block'}:
                mode = 'phased'
            else:
                raise TypeError('cannot construct a GSB header from values '
                                'without knowing the mode.')
        self = cls(None, mode=mode, nbytes=nbytes)
        self.words = list(self.words)
        if self.mode == 'phased':
            kwargs.setdefault('seq_nr', 0)
            kwargs.setdefault('mem_block', 0)
        self.update(**kwargs)
        return self

    def update(self, *, verify=True, **kwargs):
        for key in [k for k in kwargs if k in self.keys()]:
            self[key] = kwargs.pop(key)
        for prop in ('time', 'pc_time', 'gps_time'):
            if prop in kwargs:
                setattr(self, prop, kwargs.pop(prop))
        if kwargs:
            raise KeyError('GSB header cannot set {}'.format(sorted(kwargs)))
        if verify:
            self.verify()

    # --------------------------------------------------------- dict access
    _layout = {}

    def keys(self):
        return self._layout.keys()

    def __contains__(self, key):
        return key in self._layout

    def __getitem__(self, key):
        index, length, kind = self._layout[key]
        if length > 1:
            return ' '.join(self.words[index:index + length])
        return kind(self.words[index])

    def __setitem__(self, key, value):
        if not self.mutable:
            raise TypeError("header is immutable. Set '.mutable` attribute "
                            "or make a copy.")
        index, length, kind = self._layout[key]
        if length > 1:
            self.words[index:index + length] = str(value).split()
        else:
            self.words[index] = str(value)

    def seek_offset(self, n, nbytes=None):
        return n * (self.nbytes if nbytes is None else nbytes)


class GSBRawdumpHeader(GSBHeader):
    _mode = 'rawdump'
    _nwords = 7
    _layout = {'gps': (0, 7, str)}
    _properties = ('gps_time', 'time')

    @property
    def gps_time(self):
        return parse_gsb_time(self['gps'].split()) - self.utc_offset

    @gps_time.setter
    def gps_time(self, time):
        self['gps'] = ' '.join(format_gsb_time(
            as_time(time) + self.utc_offset, self._gps_precision))

    time = gps_time


class GSBPhasedHeader(GSBRawdumpHeader):
    _mode = 'phased'
    _nwords = 16
    _layout = {'pc': (0, 7, str), 'gps': (7, 7, str), 'seq_nr': (14, 1, int),
               'mem_block': (15, 1, int)}
    _properties = ('time', 'pc_time', 'gps_time')

    @property
    def pc_time(self):
        return parse_gsb_time(self['pc'].split()) - self.utc_offset

    @pc_time.setter
    def pc_time(self, time):
        self['pc'] = ' '.join(format_gsb_time(
            as_time(time) + self.utc_offset, self._pc_precision))

    @property
    def time(self):
        return self.gps_time

    @time.setter
    def time(self, time):
        self.gps_time = time
        self.pc_time = time

    def seek_offset(self, n, nbytes=None):
        """Lines grow with the number of digits of the sequence number
        (gsb/header.py:316-357)."""
        nbytes = self.nbytes if nbytes is None else nbytes
        guess = n * nbytes
        seq = self['seq_nr']
        target = seq + n
        ndseq, ndtarget = len(str(seq)), len(str(target))
        while ndseq != ndtarget:
            if n > 0:
                guess += target - 10 ** ndseq
                ndseq += 1
            else:
                guess += 10 ** (ndseq - 1) - target
                ndseq -= 1
        return guess


GSBHeader._classes.update(rawdump=GSBRawdumpHeader, phased=GSBPhasedHeader)
