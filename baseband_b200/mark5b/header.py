"""Mark 5B headers: four 32-bit words, BCD time code, CRC-16.

Field layout of baseband/mark5b/header.py:60-68; BCD properties :192-233
(incl. the "unrounding" of the 0.1 ms fraction to the 156250 ns grid);
time :235-321; CRC over the 48 time-code bits on update :150-163.  The MJD is
only stored modulo 1000, so ``kday`` (or ``ref_time``) completes it.
"""
from fractions import Fraction

from ..base.header import BitFieldHeader, FieldTable, four_word_struct
from ..base.utils import bcd_decode, bcd_encode, crc_remainder
from ..timeutil import Time, as_time

__all__ = ['Mark5BHeader', 'CRC16']

CRC16 = 0x18005


def crc16(value):
    return crc_remainder(int(value), CRC16)


class Mark5BHeader(BitFieldHeader):
    _fields = FieldTable((
        ('sync_pattern', (0, 0, 32, 0xABADDEED)),
        ('user', (1, 16, 16)),
        ('internal_tvg', (1, 15, 1)),
        ('frame_nr', (1, 0, 15)),
        ('bcd_jday', (2, 20, 12)),
        ('bcd_seconds', (2, 0, 20)),
        ('bcd_fraction', (3, 16, 16)),
        ('crc', (3, 0, 16))))
    _struct = four_word_struct
    _properties = ('payload_nbytes', 'frame_nbytes', 'complex_data', 'kday',
                   'jday', 'seconds', 'fraction', 'time')
    kday = None
    payload_nbytes = 10000
    frame_nbytes = 10016
    complex_data = False

    def __init__(self, words, kday=None, ref_time=None, verify=True):
        if kday is not None:
            self.kday = kday
        super().__init__(words, verify=verify)
        if kday is None and ref_time is not None:
            self.infer_kday(ref_time)

    def verify(self):
        assert len(self.words) == 4
        assert self['sync_pattern'] == 0xABADDEED
        assert self.kday is None or 33000 < self.kday < 400000
        if self.kday is not None:
            assert self.kday % 1000 == 0, 'kday must be thousands of MJD.'

    def copy(self, **kwargs):
        return super().copy(kday=self.kday, **kwargs)

    @classmethod
    def fromfile(cls, fh, kday=None, ref_time=None, verify=True):
        raw = fh.read(16)
        if len(raw) != 16:
            raise EOFError('could not read full header.')
        return cls(four_word_struct.unpack(raw), kday=kday,
                   ref_time=ref_time, verify=verify)

    def update(self, *, time=None, frame_rate=None, crc=None, verify=True,
               **kwargs):
        super().update(verify=False, **kwargs)
        if time is not None:
            self.set_time(time, frame_rate=frame_rate)
        if crc is None:
            stream = ((((self['bcd_jday'] << 20) + self['bcd_seconds']) << 16)
                      + self['bcd_fraction'])
            crc = crc16(stream)
        self['crc'] = crc
        if verify:
            self.verify()

    def infer_kday(self, ref_time):
        """``kday`` such that the header time is within 500 days of
        ``ref_time``."""
        mjd = as_time(ref_time).mjd
        self.kday = int(round((mjd - self.jday) / 1000.)) * 1000

    @property
    def jday(self):
        return bcd_decode(self['bcd_jday'])

    @jday.setter
    def jday(self, jday):
        self['bcd_jday'] = bcd_encode(jday)

    @property
    def seconds(self):
        return bcd_decode(self['bcd_seconds'])

    @seconds.setter
    def seconds(self, seconds):
        self['bcd_seconds'] = bcd_encode(seconds)

    @property
    def fraction_ns(self):
        ns = bcd_decode(self['bcd_fraction']) * 100000
        return 156250 * ((ns + 156249) // 156250)

    @property
    def fraction(self):
        return self.fraction_ns / 1e9

    @fraction.setter
    def fraction(self, fraction):
        ns = int(round(float(fraction) * 1e9))
        self['bcd_fraction'] = bcd_encode(ns // 100000)     # truncated

    def get_time(self, frame_rate=None):
        frame_nr = self['frame_nr']
        if frame_nr == 0:
            fraction = Fraction(0)
        elif frame_rate is None:
            if self.fraction_ns == 0:
                raise ValueError('header does not provide correct fractional '
                                 'second (it is zero for non-zero frame '
                                 'number). Please pass in a frame_rate.')
            fraction = Fraction(self.fraction_ns, 10**9)
        else:
            fraction = Fraction(frame_nr) / Fraction(
                frame_rate).limit_denominator(10**9)
        if self.kday is None:
            raise ValueError('need kday or ref_time to get a full time.')
        return Time(self.kday + self.jday, self.seconds + fraction)

    def set_time(self, time, frame_rate=None):
        time = as_time(time)
        self.kday = time.mjd // 1000 * 1000
        self.jday = time.mjd - self.kday
        int_sec = int(time.sec)
        fraction = time.sec - int_sec
        ns = Fraction(1, 10**9)
        if abs(fraction) < ns:
            frame_nr, frac = 0, 0.
        elif abs(1 - fraction) < ns:
            int_sec, frame_nr, frac = int_sec + 1, 0, 0.
        else:
            if frame_rate is None:
                raise ValueError('cannot calculate frame rate. Pass it in '
                                 'explicitly.')
            rate = Fraction(frame_rate).limit_denominator(10**9)
            frame_nr = int(round(fraction * rate))
            fraction = frame_nr / rate
            if abs(fraction - 1) < ns:
                int_sec, frame_nr, frac = int_sec + 1, 0, 0.
            else:
                frac = float(fraction)
        self.seconds = int_sec
        self.fraction = frac
        self['frame_nr'] = frame_nr

    time = property(get_time, set_time)
