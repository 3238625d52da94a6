"""Minimal UTC time arithmetic so headers can be read and written without
astropy.

A `Time` is an integer MJD plus seconds within that day kept as an exact
`fractions.Fraction`, which is all the VLBI header formats need: VDIF counts
seconds from a half-year epoch, Mark 5B stores MJD mod 1000 + BCD seconds,
Mark 4 stores year digit / day of year / h:m:s.ms.  Leap seconds are ignored,
as the reference's integer index arithmetic does
(baseband/mark5b/base.py:206-213).
"""
import datetime
from fractions import Fraction

_MJD0 = datetime.date(1858, 11, 17).toordinal()


def ymd_to_mjd(year, month, day):
    return datetime.date(year, month, day).toordinal() - _MJD0


def mjd_to_ymd(mjd):
    d = datetime.date.fromordinal(int(mjd) + _MJD0)
    return d.year, d.month, d.day


class Time:
    __slots__ = ('mjd', 'sec')

    def __init__(self, mjd, sec=0):
        sec = Fraction(sec).limit_denominator(10**12)
        extra, sec = divmod(sec, 86400)
        self.mjd = int(mjd) + int(extra)
        self.sec = sec

    @classmethod
    def from_isot(cls, text):
        date, _, clock = text.strip().partition('T')
        y, m, d = (int(v) for v in date.split('-'))
        sec = Fraction(0)
        if clock:
            h, mi, s = clock.split(':')
            sec = int(h) * 3600 + int(mi) * 60 + Fraction(s)
        return cls(ymd_to_mjd(y, m, d), sec)

    @classmethod
    def from_mjd(cls, mjd):
        whole = int(mjd // 1)
        return cls(whole, Fraction(mjd - whole) * 86400)

    @classmethod
    def from_yday(cls, year, yday, sec=0):
        return cls(ymd_to_mjd(year, 1, 1) + yday - 1, sec)

    @property
    def year(self):
        return mjd_to_ymd(self.mjd)[0]

    @property
    def yday(self):
        y = self.year
        return self.mjd - ymd_to_mjd(y, 1, 1) + 1

    @property
    def isot(self):
        # round to whole nanoseconds first, so that a fraction that rounds
        # up to 10^9 ns carries into the seconds (and the day)
        total_ns = int(round(self.sec * 10**9))
        mjd = self.mjd
        if total_ns >= 86400 * 10**9:
            total_ns -= 86400 * 10**9
            mjd += 1
        y, m, d = mjd_to_ymd(mjd)
        whole, ns = divmod(total_ns, 10**9)
        h, rem = divmod(whole, 3600)
        mi, s = divmod(rem, 60)
        return '{:04d}-{:02d}-{:02d}T{:02d}:{:02d}:{:02d}.{:09d}'.format(
            y, m, d, h, mi, s, ns)

    def __add__(self, seconds):
        return Time(self.mjd, self.sec + Fraction(seconds).limit_denominator(
            10**12))

    def __sub__(self, other):
        if isinstance(other, Time):
            return (self.mjd - other.mjd) * 86400 + (self.sec - other.sec)
        return self + (-other)

    def __eq__(self, other):
        return (isinstance(other, Time) and self.mjd == other.mjd
                and self.sec == other.sec)

    def __lt__(self, other):
        return (self.mjd, self.sec) < (other.mjd, other.sec)

    def __le__(self, other):
        return (self.mjd, self.sec) <= (other.mjd, other.sec)

    def __hash__(self):
        return hash((self.mjd, self.sec))

    def __repr__(self):
        return "<Time '{}'>".format(self.isot)


def as_time(value):
    """Accept a Time, an ISO string, an astropy-like Time (``.jd1``/``.jd2``)
    or anything with ``.isot`` or ``.mjd``."""
    if value is None or isinstance(value, Time):
        return value
    if isinstance(value, str):
        return Time.from_isot(value)
    # astropy-like objects: the two-double Julian date keeps sub-nanosecond
    # resolution (``isot`` prints only 3 decimals by default, ``mjd`` is one
    # double: ~1 us); rounded to whole nanoseconds
    utc = getattr(value, 'utc', value)
    if hasattr(utc, 'jd1') and hasattr(utc, 'jd2'):
        days = (Fraction(float(utc.jd1)) + Fraction(float(utc.jd2))
                - Fraction(24000005, 10))
        whole = days.numerator // days.denominator
        ns = int(round((days - whole) * 86400 * 10**9))
        return Time(whole, Fraction(ns, 10**9))
    if hasattr(value, 'isot'):
        return Time.from_isot(str(value.isot))
    if hasattr(value, 'mjd'):
        return Time.from_mjd(float(value.mjd))
    raise TypeError('cannot interpret {!r} as a time'.format(value))
