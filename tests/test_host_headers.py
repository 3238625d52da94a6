"""Host header classes (bit fields, derived sizes, times, CRCs) against the
golden field values produced by the unmodified reference, and round trips
through ``fromvalues`` / ``tofile`` / ``fromfile``.  Pure host code: no GPU."""
import io

import numpy as np
import pytest

from baseband_b200.vdif.header import VDIFHeader
from baseband_b200.mark5b.header import Mark5BHeader
from baseband_b200.mark4.header import Mark4Header, stream2words, words2stream
from baseband_b200.base.utils import (bcd_decode, bcd_encode, crc_remainder,
                                      crc_array, crc_of_bits)
from baseband_b200.timeutil import Time
from conftest import sample_path

VDIF_KEYS = ('invalid_data', 'legacy_mode', 'seconds', 'ref_epoch',
             'frame_nr', 'vdif_version', 'lg2_nchan', 'frame_length',
             'complex_data', 'bits_per_sample', 'thread_id', 'station_id')


@pytest.mark.parametrize('name', ['sample.vdif', 'sample_vlbi.vdif',
                                  'sample_mwa.vdif', 'sample_arochime.vdif',
                                  'sample_bps1.vdif'])
def test_vdif_header_fields(sample_outputs, name):
    want = sample_outputs[name.replace('.', '_') + '_fields']
    with open(sample_path(name), 'rb') as fh:
        for row in want:
            h = VDIFHeader.fromfile(fh)
            got = [int(h[k]) for k in VDIF_KEYS] + [
                h.edv if h.edv else 0, h.payload_nbytes, h.samples_per_frame]
            assert got == list(row)
            # write it back unchanged
            buf = io.BytesIO()
            h.tofile(buf)
            fh.seek(-h.nbytes, 1)
            assert buf.getvalue() == fh.read(h.nbytes)
            fh.seek(h.payload_nbytes, 1)


def test_vdif_header_fromvalues_and_time():
    with open(sample_path('sample.vdif'), 'rb') as fh:
        h = VDIFHeader.fromfile(fh)
    assert h.edv == 3 and h.station == 65532 and h.bps == 2 and h.nchan == 1
    assert h.sample_rate == 32e6 and h.frame_rate == 1600.
    assert h.time.isot == '2014-06-16T05:56:07.000000000'
    h2 = VDIFHeader.fromvalues(
        edv=3, time=h.time, samples_per_frame=20000, station=65532, bps=2,
        nchan=1, complex_data=False, thread_id=1, sample_rate=32e6,
        loif_tuning=h['loif_tuning'], dbe_unit=h['dbe_unit'],
        if_nr=h['if_nr'], subband=h['subband'], sideband=h['sideband'],
        major_rev=h['major_rev'], minor_rev=h['minor_rev'],
        personality=h['personality'], _7_28_4=h['_7_28_4'])
    assert h2 == h
    # times off the second boundary need the frame rate
    h3 = h.copy()
    h3.mutable = True
    h3.set_time(h.time + 0.5, frame_rate=1600.)
    assert h3['frame_nr'] == 800 and h3['seconds'] == h['seconds']
    assert h3.get_time(frame_rate=1600.) - h.time == 0.5
    # legacy and EDV 1 round trips
    legacy = VDIFHeader.fromvalues(edv=False, time='2015-01-01T00:00:00',
                                   payload_nbytes=8000, bps=2, nchan=4,
                                   station=65)
    assert legacy.nbytes == 16 and legacy['legacy_mode']
    assert legacy.samples_per_frame == 8000
    buf = io.BytesIO()
    legacy.tofile(buf)
    buf.seek(0)
    assert VDIFHeader.fromfile(buf) == legacy
    with pytest.raises(TypeError):
        h['frame_nr'] = 5                      # read from file: immutable
    with pytest.raises(KeyError):
        h['no_such_key']


def test_mark5b_header(sample_outputs):
    want = sample_outputs['sample_m5b_fields']
    with open(sample_path('sample.m5b'), 'rb') as fh:
        for row in want:
            h = Mark5BHeader.fromfile(fh, kday=56000)
            got = [int(h[k]) for k in ('sync_pattern', 'user', 'internal_tvg',
                                       'frame_nr', 'bcd_jday', 'bcd_seconds',
                                       'bcd_fraction', 'crc')]
            got += [h.jday, h.seconds, int(round(h.fraction * 1e9))]
            assert got == list(row)
            # CRC and BCD time regenerate exactly
            h2 = h.copy()
            h2.mutable = True
            h2.update(time=h.time, frame_rate=6400.)
            assert h2 == h
            fh.seek(10000, 1)
    assert h.time.isot == '2014-06-13T05:30:01.000468750'
    h3 = Mark5BHeader.fromvalues(time=Time.from_isot('2014-06-13T05:30:01'),
                                 user=3901, internal_tvg=False)
    assert h3.kday == 56000 and h3.jday == 821 and h3['frame_nr'] == 0
    hr = Mark5BHeader(h.words, ref_time='2014-01-01T00:00:00')
    assert hr.kday == 56000


@pytest.mark.parametrize('name,ntrack', [
    ('sample.m4', 64), ('sample_32track.m4', 32),
    ('sample_32track_fanout2.m4', 32), ('sample_16track.m4', 16),
    ('sample_64track_fanout2_ft.m4', 64)])
def test_mark4_header(sample_outputs, name, ntrack):
    tag = name.replace('.', '_')
    off0 = int(sample_outputs[tag + '_offset0'])
    want = sample_outputs[tag + '_track_fields']
    _, fanout, nchan, bps, spf = [int(v) for v in sample_outputs[tag + '_geom']]
    # the golden fields are those of the LAST frame of the sample
    nframe = sample_outputs[tag + '_data'].shape[0] // spf
    start = off0 + (nframe - 1) * ntrack * 2500
    with open(sample_path(name), 'rb') as fh:
        fh.seek(start)
        raw = fh.read(ntrack * 20)
        fh.seek(start)
        h = Mark4Header.fromfile(fh, ntrack, decade=2010)
    keys = ('fan_out', 'magnitude_bit', 'lsb_output', 'converter_id',
            'bcd_unit_year', 'bcd_day', 'bcd_hour', 'bcd_minute',
            'bcd_second', 'bcd_fraction', 'crc', 'sync_pattern')
    got = np.array([np.asarray(h[k]).astype(np.int64) for k in keys])
    assert np.array_equal(got, want)
    assert (h.fanout, h.nchan, h.bps, h.samples_per_frame) == (fanout, nchan,
                                                               bps, spf)
    # bit transpose and CRC-12 regenerate the bytes on disk
    stream = np.frombuffer(raw, h.stream_dtype)
    assert np.array_equal(words2stream(stream2words(stream)), stream)
    h2 = h.copy()
    h2.update(time=h.time)
    buf = io.BytesIO()
    h2.tofile(buf)
    assert buf.getvalue() == raw
    assert Mark4Header(h.words, ref_time=h.time + 86400 * 400).decade == 2010


def test_bcd_and_crc():
    assert bcd_decode(0x1234) == 1234 and bcd_encode(8765) == 0x8765
    arr = np.array([0x0821, 0x0999], np.uint32)
    assert list(bcd_decode(arr)) == [821, 999]
    assert list(bcd_encode(np.array([821, 999]))) == [0x821, 0x999]
    with pytest.raises(ValueError):
        bcd_decode(np.array([0x1a], np.uint32))
    # CRC-16 of the Mark 5B sample time code (mark5b/tests: crc 38749)
    stream = (((0x821 << 20) + 0x19801) << 16) + 0
    assert crc_remainder(stream, 0x18005) == 38749
    assert crc_array(np.array([stream]), 48, 0x18005)[0] == 38749
    assert crc_remainder((stream << 16) | 38749, 0x18005, extend=False) == 0
    # bit-parallel CRC agrees with the scalar one on every bit lane
    rng = np.random.default_rng(2)
    bits = rng.integers(0, 2, (40, 8))
    lanes = (bits << np.arange(8)).sum(1).astype(np.uint8)
    crc = crc_of_bits(lanes, 0x180f)
    for lane in range(8):
        value = int(''.join(str(b) for b in bits[:, lane]), 2)
        want = crc_remainder(value, 0x180f)
        got = int(''.join(str((int(c) >> lane) & 1) for c in crc), 2)
        assert got == want


def test_guess_format():
    import baseband_b200 as bb
    expected = {'sample.vdif': 'vdif', 'sample_vlbi.vdif': 'vdif',
                'sample_mwa.vdif': 'vdif', 'sample_arochime.vdif': 'vdif',
                'sample_bps1.vdif': 'vdif', 'sample.m5b': 'mark5b',
                'sample.m4': 'mark4', 'sample_32track.m4': 'mark4',
                'sample_16track.m4': 'mark4', 'sample_puppi.raw': 'guppi',
                'sample.dada': 'dada', 'sample_mkbf.dada': 'dada',
                'gsb/sample_gsb_rawdump.dat': None}
    for name, fmt in expected.items():
        assert bb.guess_format(sample_path(name)) == fmt, name


def test_plugin_protocol():
    """Every format module offers what the reference's ``baseband.io`` entry
    points need: ``open`` and (except GSB, which has no file signature)
    ``info`` that is truthy only for its own format."""
    import importlib
    import baseband_b200 as bb
    for fmt in bb.FORMATS:
        module = importlib.import_module('baseband_b200.' + fmt)
        assert callable(module.open)
    assert bb.vdif.info(sample_path('sample.vdif'))
    assert not bb.vdif.info(sample_path('sample.m5b'))
    assert not bb.mark5b.info(sample_path('sample.vdif'))
    m5 = bb.mark5b.info(sample_path('sample.m5b'))
    assert m5 and m5.format == 'mark5b' and m5.readable is False  # needs nchan
    assert bb.dada.info(sample_path('sample.dada')).format == 'dada'


def test_as_time_two_double_julian_date():
    """astropy-like times are taken from jd1/jd2 (sub-microsecond), not from
    the 3-decimal ``isot``; a fraction that rounds to 10^9 ns carries."""
    from fractions import Fraction
    from baseband_b200.timeutil import Time, as_time

    class FakeAstropy:
        # 2014-06-16T05:56:07.000123456 UTC: MJD 56824
        jd1 = 2456824.5
        jd2 = (5 * 3600 + 56 * 60 + 7 + 0.000123456) / 86400.
        isot = '2014-06-16T05:56:07.000'
        mjd = 56824.24730324217

    t = as_time(FakeAstropy())
    assert t.mjd == 56824
    assert abs(t.sec - Fraction(21367000123456, 10**9)) <= Fraction(2, 10**9)
    assert t.isot.startswith('2014-06-16T05:56:07.00012345')
    # carry: 0.9999999996 s prints as the next second, not '.1000000000'
    t = Time(56824, Fraction(86399) + Fraction(9999999996, 10**10))
    assert t.isot == '2014-06-17T00:00:00.000000000'
    t = Time(56824, Fraction(59) + Fraction(9999999996, 10**10))
    assert t.isot == '2014-06-16T00:01:00.000000000'


def test_file_reader_locate_frames_and_find_header():
    """`locate_frames` / `find_header` of the binary file readers: the
    answers the reference's own tests assert (vdif/tests/test_vdif.py:695-760,
    mark5b/tests/test_mark5b.py:489-560, mark4 sample at 0xa88), and
    agreement with the oracle on every case."""
    import io
    import numpy as np
    import pytest
    import baseband_b200 as bb
    from baseband_b200.base.locate import HeaderNotFoundError
    from oracle import locate as olocate
    from conftest import sample_bytes
    data = sample_bytes('sample.vdif')
    with bb.vdif.open(sample_path('sample.vdif'), 'rb') as fh:
        header0 = fh.read_header()
        fh.seek(0)
        every = [x * 5032 for x in range(16)]
        assert fh.locate_frames(header0['sync_pattern'], offset=20) == every
        fh.seek(0, 2)
        assert fh.locate_frames(header0['sync_pattern'], offset=20,
                                forward=False) == every[::-1]
        fh.seek(10)
        mask = [0, 0, 0xffffffff, 0xfc00ffff, 0xffffffff, 0, 0, 0]
        assert fh.locate_frames(header0.words, mask=mask,
                                frame_nbytes=5032) == [5032, 10064]
        assert fh.tell() == 10
        for pos, fwd, want in ((5000, True, [5032, 10064]),
                               (15000, True, [15096, 20128]),
                               (20128, True, [20128, 25160]),
                               (16, False, [0]),
                               (data.size - 10000, False, [70448, 65416]),
                               (data.size - 5000, False, [75480, 70448]),
                               (data.size - 20, True, []),
                               (40254, True, [40256, 45288]),
                               (40254, False, [35224, 30192])):
            fh.seek(pos)
            assert fh.locate_frames(header0, forward=fwd) == want
            pat, msk = header0.invariant_pattern()
            assert want == olocate.locate_frames(
                data, pos, pat, mask=msk, frame_nbytes=5032, forward=fwd)
        fh.seek(5000)
        header = fh.find_header(header0)
        assert fh.tell() == 5032 and header['frame_nr'] == 0
        fh.seek(data.size - 20)
        with pytest.raises(HeaderNotFoundError):
            fh.find_header(header0)
        with pytest.raises(TypeError):
            fh.locate_frames()
    m5 = sample_bytes('sample.m5b')
    with bb.mark5b.open(sample_path('sample.m5b'), 'rb', kday=56000,
                        nchan=8) as fh:
        assert fh.locate_frames() == [0, 10016]
        assert fh.locate_frames(forward=False) == [0]
        fh.seek(10000)
        assert fh.locate_frames() == [10016, 20032]
        assert fh.locate_frames(forward=False) == [0]
        assert fh.find_header()['frame_nr'] == 1 and fh.tell() == 10016
        fh.seek(-10000, 2)
        assert fh.locate_frames(forward=False) == [30048, 20032]
        fh.seek(-30, 2)
        assert fh.locate_frames() == []
    bad = np.concatenate([m5[:10040], m5[20000:]])
    with bb.mark5b.open(io.BytesIO(bad.tobytes()), 'rb', kday=56000,
                        nchan=8) as fh:
        shifted = 2 * 10016 - 9960
        assert fh.locate_frames() == [0, shifted]
        assert fh.locate_frames(check=None) == [0, 10016, shifted]
        fh.seek(10000)
        assert fh.locate_frames() == [shifted, shifted + 10016]
    with bb.mark4.open(sample_path('sample.m4'), 'rb', ntrack=64,
                       decade=2010) as fh:
        assert fh.locate_frames(maximum=10000)[0] == 0xa88
        assert fh.find_header(maximum=10000).ntrack == 64
        assert fh.tell() == 0xa88
