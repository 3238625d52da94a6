"""DADA headers: ``KEY value # comment`` lines in a block of ``HDR_SIZE``
(4096) bytes (baseband/dada/header.py:117-200); sizes and times derived as
in :289-440 (``FILE_SIZE`` payload bytes, ``OBS_OFFSET`` bytes since
``UTC_START``/``MJD_START``, ``TSAMP`` in microseconds)."""
import io
from fractions import Fraction

from ..timeutil import Time, as_time

__all__ = ['DADAHeader']

_INT_KEYS = ('FILE_SIZE', 'FILE_NUMBER', 'HDR_SIZE', 'OBS_OFFSET',
             'OBS_OVERLAP', 'NBIT', 'NDIM', 'NPOL', 'NCHAN', 'RESOLUTION',
             'DSB')
_FLOAT_KEYS = ('FREQ', 'BW', 'TSAMP')


class DADAHeader(dict):
    """Ordered mapping of DADA keywords; ``comments`` holds the comments.
    Lines without a keyword (blank or comment only) are kept under keys
    ``_<line number>`` so a header is written back unchanged."""

    _properties = ('payload_nbytes', 'frame_nbytes', 'bps', 'complex_data',
                   'sample_shape', 'sample_rate', 'sideband', 'tsamp',
                   'samples_per_frame', 'offset', 'start_time', 'time')
    _defaults = (('HEADER', 'DADA'), ('HDR_VERSION', '1.0'),
                 ('HDR_SIZE', 4096), ('DADA_VERSION', '1.0'),
                 ('OBS_ID', 'unset'), ('PRIMARY', 'unset'),
                 ('SECONDARY', 'unset'), ('FILE_NAME', 'unset'),
                 ('FILE_NUMBER', 0), ('FILE_SIZE', 0), ('OBS_OFFSET', 0),
                 ('OBS_OVERLAP', 0), ('SOURCE', 'unset'),
                 ('TELESCOPE', 'unset'), ('INSTRUMENT', 'unset'),
                 ('RECEIVER', 'unset'), ('NBIT', 8), ('NDIM', 1), ('NPOL', 1),
                 ('NCHAN', 1), ('RESOLUTION', 1), ('DSB', 1))

    def __init__(self, items=(), verify=True, mutable=True, **kwargs):
        super().__init__()
        self.mutable = True
        self.comments = {}
        if isinstance(items, str):
            items = self._fromlines(items.split('\n'))
        for key, value in (items.items() if hasattr(items, 'items')
                           else items):
            self[key] = value
        for key, value in kwargs.items():
            self[key] = value
        self.mutable = mutable
        if verify and len(self):
            self.verify()

    def verify(self):
        known = {k for k, _ in self._defaults}
        assert len(known.intersection(self.keys())) > 10

    def copy(self):
        new = type(self)(verify=False)
        for key in self:
            dict.__setitem__(new, key, self[key])
        new.comments = dict(self.comments)
        return new

    __copy__ = copy

    def __setitem__(self, key, value):
        if not self.mutable:
            raise TypeError('immutable {0} does not support assignment.'
                            .format(type(self).__name__))
        if isinstance(value, tuple):
            value, comment = value
            self.comments[key.upper()] = comment
        super().__setitem__(key.upper(), value)

    @staticmethod
    def _fromlines(lines):
        items = []
        for number, line in enumerate(lines):
            head, _, comment = line.strip().partition('#')
            comment = comment.strip() or None
            parts = head.split()
            key = parts[0] if parts else '_{0:d}'.format(number)
            value = parts[1] if len(parts) > 1 else None
            if key in _INT_KEYS:
                value = int(value)
            elif key in _FLOAT_KEYS:
                value = float(value)
            items.append((key, (value, comment)))
        return items

    def _tolines(self):
        lines = []
        for key, value in self.items():
            comment = self.comments.get(key)
            if value is not None:
                line = '{0} {1}'.format(key, value)
                if comment is not None:
                    line += ' # {0}'.format(comment)
            else:
                line = '# {0}'.format(comment) if comment is not None else ''
            lines.append(line)
        return lines

    @classmethod
    def fromfile(cls, fh, verify=True):
        start = fh.tell()
        block = fh.read(4096)
        if len(block) == 0:
            raise EOFError('could not read DADA header.')
        text = block.split(b'\x00')[0].decode('ascii')
        size = 4096
        for line in text.split('\n'):
            if line.startswith('HDR_SIZE'):
                size = int(line.split()[1])
        if size > 4096:
            fh.seek(start)
            text = fh.read(size).split(b'\x00')[0].decode('ascii')
        lines = []
        for line in text.split('\n'):
            if line[:1] == '#' and 'end of header' in line:
                break
            lines.append(line)
        else:
            if lines and lines[-1] == '':
                lines.pop()
        fh.seek(start + size)
        return cls(cls._fromlines(lines), verify=verify, mutable=False)

    def tofile(self, fh):
        with io.BytesIO() as s:
            for line in self._tolines():
                s.write((line + '\n').encode('ascii'))
            s.write(b'# end of header\n')
            extra = self.nbytes - s.tell()
            if extra < 0:
                raise ValueError('cannot write header in allocated size of '
                                 '{0}'.format(self.nbytes))
            return fh.write(s.getvalue() + b'\0' * extra)

    @classmethod
    def fromkeys(cls, *args, **kwargs):
        if not args:
            kwargs.setdefault('HEADER', 'DADA')
        return cls(*args, **kwargs)

    @classmethod
    def fromvalues(cls, **kwargs):
        self = cls(cls._defaults, verify=False)
        self.update(**kwargs)
        return self

    def update(self, *, verify=True, **kwargs):
        extras = [(p, kwargs.pop(p)) for p in self._properties
                  if p in kwargs]
        for key, value in kwargs.items():
            self[key] = value
        for attr, value in extras:
            setattr(self, attr, value)
        if verify:
            self.verify()

    # ------------------------------------------------------------- geometry
    @property
    def nbytes(self):
        return self['HDR_SIZE']

    @property
    def payload_nbytes(self):
        return self['FILE_SIZE']

    @payload_nbytes.setter
    def payload_nbytes(self, nbytes):
        self['FILE_SIZE'] = int(nbytes)

    @property
    def frame_nbytes(self):
        return self.nbytes + self.payload_nbytes

    @frame_nbytes.setter
    def frame_nbytes(self, nbytes):
        self.payload_nbytes = nbytes - self.nbytes

    @property
    def bps(self):
        return self['NBIT']

    @bps.setter
    def bps(self, bps):
        self['NBIT'] = int(bps)

    @property
    def complex_data(self):
        return self['NDIM'] == 2

    @complex_data.setter
    def complex_data(self, complex_data):
        self['NDIM'] = 2 if complex_data else 1

    @property
    def sample_shape(self):
        return self['NPOL'], self['NCHAN']

    @sample_shape.setter
    def sample_shape(self, sample_shape):
        self['NPOL'], self['NCHAN'] = (int(d) for d in sample_shape)

    @property
    def _bits_per_sample(self):
        return self['NBIT'] * self['NDIM'] * self['NPOL'] * self['NCHAN']

    @property
    def sample_rate(self):
        """Complete samples per second (TSAMP is in microseconds)."""
        return 1e6 / self['TSAMP']

    @sample_rate.setter
    def sample_rate(self, sample_rate):
        to_value = getattr(sample_rate, 'to_value', None)
        mhz = (float(to_value('Hz')) if to_value else float(sample_rate)) / 1e6
        self['TSAMP'] = 1. / abs(mhz)
        bw = mhz * self['NCHAN'] / (1 if self.complex_data else 2)
        self['BW'] = (-1 if self.get('BW', bw) < 0 else 1) * bw

    tsamp = property(lambda self: self['TSAMP'])

    @property
    def sideband(self):
        return self['BW'] > 0

    @sideband.setter
    def sideband(self, sideband):
        self['BW'] = (1 if sideband else -1) * abs(self['BW'])

    @property
    def samples_per_frame(self):
        return self.payload_nbytes * 8 // self._bits_per_sample

    @samples_per_frame.setter
    def samples_per_frame(self, samples_per_frame):
        old = self.payload_nbytes
        self.payload_nbytes = (samples_per_frame * self._bits_per_sample
                               + 7) // 8
        if self.samples_per_frame != samples_per_frame:
            nearest = self.samples_per_frame
            self.payload_nbytes = old
            raise ValueError('header cannot store {} samples per frame. '
                             'Nearest is {}.'.format(samples_per_frame,
                                                     nearest))

    # ----------------------------------------------------------------- time
    def _tsamp(self):
        return Fraction(self['TSAMP']).limit_denominator(10**12) / 10**6

    @property
    def offset(self):
        return (self['OBS_OFFSET'] * 8 // self._bits_per_sample) \
            * self._tsamp()

    @offset.setter
    def offset(self, offset):
        self['OBS_OFFSET'] = (int(round(Fraction(offset) / self._tsamp()))
                              * ((self._bits_per_sample + 7) // 8))

    @property
    def start_time(self):
        if 'MJD_START' in self:
            whole, frac = str(self['MJD_START']).split('.')
            sec = Fraction(int(frac), 10 ** len(frac)) * 86400
            # stored to ~1e-15 day: round to the nearest nanosecond
            sec = Fraction(int(round(sec * 10**9)), 10**9)
            return Time(int(whole), sec)
        t0 = self['UTC_START']
        return Time.from_isot(t0[:10] + 'T' + t0[11:])

    @start_time.setter
    def start_time(self, start_time):
        t = as_time(start_time)
        isot = t.isot
        self['UTC_START'] = isot.replace('T', '-').replace('.000000000', '')
        frac = '{0:17.15f}'.format(float(t.sec / 86400))[1:]
        self['MJD_START'] = '{0:05d}'.format(t.mjd) + frac

    @property
    def time(self):
        return self.start_time + self.offset

    @time.setter
    def time(self, time):
        time = as_time(time)
        if 'MJD_START' not in self:
            self.start_time = time - self.offset
        else:
            self.offset = time - self.start_time

    def __eq__(self, other):
        keys = {k for k in set(self.keys()) | set(other.keys())
                if not k.startswith('_') and k != 'MJD_START'}
        return (all(self.get(k) == other.get(k) for k in keys)
                and float(self.get('MJD_START', 0.))
                == float(other.get('MJD_START', 0.)))

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return '{0}("""{1}""")'.format(type(self).__name__,
                                       '\n'.join(self._tolines()))
