"""The oracle against (a) vectors produced by the unmodified reference
(tests/golden/make_golden.py) and (b) the known-answer values the reference's
own tests assert (mostly mark5access output quoted there).  CPU only."""
import numpy as np
import pytest

from oracle import codec, headers, stream
from conftest import sample_bytes


def test_levels_and_luts(codec_vectors):
    g = codec_vectors
    for bps in (1, 2, 4):
        assert np.array_equal(codec.LEVELS[bps].view('u4'),
                              g['levels%d' % bps].view('u4'))
        assert np.array_equal(codec.VDIF_LUT[bps], g['vdif_lut%d' % bps])
    assert np.array_equal(codec.MARK5B_LUT1, g['m5b_lut1'])
    assert np.array_equal(codec.MARK5B_LUT2, g['m5b_lut2'])
    assert np.array_equal(codec.M4_LUT1, g['m4_lut1'])
    assert np.array_equal(codec.M4_LUT2_1, g['m4_lut2_1'])
    assert np.array_equal(codec.M4_LUT2_2, g['m4_lut2_2'])
    assert np.array_equal(codec.M4_LUT2_3, g['m4_lut2_3'])
    # reference known answers: vdif/tests/test_vdif.py:320-345
    assert np.all(codec.VDIF_LUT[2][0b10100101] == [-1., -1., 1., 1.])
    assert codec.LEVELS[2][3] == np.float32(3.316505)


@pytest.mark.parametrize('bps', [1, 2, 4, 8])
def test_vdif_decode(codec_vectors, bps):
    got = codec.vdif_decode(codec_vectors['words32'], bps).ravel()
    assert np.array_equal(got.view('u4'),
                          codec_vectors['vdif_dec%d' % bps].view('u4'))


@pytest.mark.parametrize('bps', [1, 2])
def test_mark5b_decode(codec_vectors, bps):
    got = codec.mark5b_decode(codec_vectors['words32'], bps).ravel()
    assert np.array_equal(got, codec_vectors['m5b_dec%d' % bps])


@pytest.mark.parametrize('tag', ['f32', 'f64'])
@pytest.mark.parametrize('bps', [1, 2, 4, 8])
def test_vdif_encode(codec_vectors, bps, tag):
    vals = codec_vectors['enc_in_' + tag]
    with np.errstate(all='ignore'):
        got = codec.vdif_encode(vals.copy(), bps).ravel().view(np.uint8)
    assert np.array_equal(got, codec_vectors['vdif_enc%d_%s' % (bps, tag)])


@pytest.mark.parametrize('tag', ['f32', 'f64'])
@pytest.mark.parametrize('bps', [1, 2])
def test_mark5b_encode(codec_vectors, bps, tag):
    vals = codec_vectors['enc_in_' + tag]
    got = codec.mark5b_encode(vals.copy(), bps).ravel().view(np.uint8)
    assert np.array_equal(got, codec_vectors['m5b_enc%d_%s' % (bps, tag)])


@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_int8_gsb4_encode(codec_vectors, tag):
    vals = codec_vectors['enc_in_finite_' + tag]
    assert np.array_equal(codec.int8_encode(vals).view(np.uint8),
                          codec_vectors['int8_enc_' + tag])
    assert np.array_equal(codec.gsb4_encode(vals).view(np.uint8),
                          codec_vectors['gsb4_enc_' + tag])


def test_int8_gsb4_decode(codec_vectors):
    b = codec_vectors['bytes']
    assert np.array_equal(codec.int8_decode(b), codec_vectors['int8_dec'])
    assert np.array_equal(codec.gsb4_decode(b), codec_vectors['gsb4_dec'])
    # gsb/tests/test_gsb.py:233-248: byte 0x7f -> (-1, 7); 0x80 -> (0, -8)
    assert np.all(codec.gsb4_decode(np.array([0x7f, 0x80], 'u1').view('i1'))
                  == [-1., 7., 0., -8.])


def test_nan_encode(codec_vectors):
    nanv = codec_vectors['nan_in']
    with np.errstate(invalid='ignore'):
        assert np.array_equal(codec.vdif_encode(nanv.copy(), 2),
                              codec_vectors['vdif_enc2_nan'])
        assert np.array_equal(codec.vdif_encode(nanv.copy(), 1),
                              codec_vectors['vdif_enc1_nan'])
        assert np.array_equal(codec.mark5b_encode(nanv.copy(), 1),
                              codec_vectors['m5b_enc1_nan'])


def test_mark4_reorder(codec_vectors):
    g = codec_vectors
    # mark4/tests/test_mark4.py:302-308 (C code known answer)
    x = np.array([738811025863578102], np.uint64)
    assert codec.m4_reorder64(x)[0] == 738829572664316278
    assert list(codec.m4_reorder64(x).view(np.uint8)) == [
        118, 209, 53, 244, 148, 217, 64, 10]
    w = g['words32']
    assert np.array_equal(codec.m4_reorder32(w.view(np.uint32)),
                          g['m4_reorder32'])
    assert np.array_equal(codec.m4_reorder64(w.view(np.uint64)),
                          g['m4_reorder64'])
    assert np.array_equal(codec.m4_reorder64_ft(w.view(np.uint64)),
                          g['m4_reorder64_ft'])


M4_MODES = {'2_4': (2, 4, False), '4_4': (4, 4, False), '8_2': (8, 2, False),
            '8_4': (8, 4, False), '16_2ft': (16, 2, True)}


@pytest.mark.parametrize('tag', sorted(M4_MODES))
def test_mark4_codec(codec_vectors, tag):
    g = codec_vectors
    nchan, fanout, ft = M4_MODES[tag]
    got = codec.mark4_decode(g['m4_words_' + tag], nchan, fanout, ft)
    assert np.array_equal(got, g['m4_dec_' + tag])
    for ftag in ('f32', 'f64'):
        vals = g['m4_enc_in_%s_%s' % (tag, ftag)]
        enc = codec.mark4_encode(vals.copy(), nchan, fanout, ft)
        assert np.array_equal(np.ascontiguousarray(enc).view(np.uint8),
                              g['m4_enc_%s_%s' % (tag, ftag)])


def test_mark4_known_decode():
    # mark4/payload.py:75-86 comment (decode of the reorder test word)
    d = codec.mark4_decode(np.array([738811025863578102], '<u8'), 8, 4)
    want = np.array([[-1, 1, 3, 1], [1, 1, 3, -3], [1, -3, 1, 3],
                     [-3, 1, 3, 3], [-3, 1, 1, -1], [-3, -3, -3, 1],
                     [1, -1, 1, 3], [-1, -1, -3, -3]]).T
    assert np.array_equal(np.round(d).astype(int),
                          np.where(abs(want) == 3, want, want))


# ------------------------------------------------------------ sample files
VDIF_FIELD_NAMES = ('invalid_data', 'legacy_mode', 'seconds', 'ref_epoch',
                    'frame_nr', 'vdif_version', 'lg2_nchan', 'frame_length',
                    'complex_data', 'bits_per_sample', 'thread_id',
                    'station_id')


@pytest.mark.parametrize('name', ['sample.vdif', 'sample_vlbi.vdif',
                                  'sample_mwa.vdif', 'sample_arochime.vdif',
                                  'sample_bps1.vdif'])
def test_vdif_samples(sample_outputs, name):
    tag = name.replace('.', '_')
    raw = sample_bytes(name)
    h0, frames = stream.vdif_scan(raw)
    fields = sample_outputs[tag + '_fields']
    assert len(frames) == len(fields)
    for h, row in zip(frames, fields):
        for k, v in zip(VDIF_FIELD_NAMES, row):
            assert h[k] == v, k
        assert (h.get('edv') or 0) == row[12]
        assert h['payload_nbytes'] == row[13]
        assert h['samples_per_frame'] == row[14]
    data = stream.vdif_read(raw)
    want = sample_outputs[tag + '_data']
    assert data.shape == want.shape and data.dtype == want.dtype
    assert np.array_equal(data, want)


def test_vdif_sample_known_values():
    # vdif/tests/test_vdif.py:930-931 (mark5access m5d values)
    data = stream.vdif_read(sample_bytes('sample.vdif'))[:, :, 0]
    assert data.shape == (40000, 8)
    assert np.all(data[:12, 0].astype(int)
                  == [-1, -1, 3, -1, 1, -1, 3, -1, 1, 3, -1, 1])
    # vdif/tests/test_vdif.py:21-27: thread 1 payload starts 2a 0a 7c
    assert np.all(np.round(data[:12, 1]).astype(int)
                  == [1, 1, 1, -3, 1, 1, -3, -3, -3, 3, 3, -1])
    # thread order in the file is 1,3,5,7,0,2,4,6 (test_vdif.py:823)
    _, frames = stream.vdif_scan(sample_bytes('sample.vdif'))
    assert [f['thread_id'] for f in frames[:8]] == [1, 3, 5, 7, 0, 2, 4, 6]


def test_vdif_invalid_fill(sample_outputs):
    raw = sample_bytes('sample.vdif').copy()
    _, frames = stream.vdif_scan(raw)
    for h in frames[:8]:
        if h['thread_id'] in (1, 4, 7):
            raw[h['offset'] + 3] |= 0x80       # invalid_data = word 0 bit 31
    data = stream.vdif_read(raw, fill_value=-999., count=20000)
    assert np.array_equal(
        data, sample_outputs['sample_vdif_set0_invalid_1_4_7_fill_m999'])


def test_mark5b_sample(sample_outputs):
    raw = sample_bytes('sample.m5b')
    data = stream.mark5b_read(raw, nchan=8)
    assert np.array_equal(data, sample_outputs['sample_m5b_data'])
    assert np.array_equal(stream.mark5b_valid_mask(raw),
                          sample_outputs['sample_m5b_valid'])
    # mark5b/tests/test_mark5b.py:172-175
    assert np.all(data[:3].astype(int) == [[-3, -1, 1, -1, 3, -3, -3, 3],
                                           [-3, 3, -1, 3, -1, -1, -1, 1],
                                           [3, -1, 3, 3, 1, -1, 3, -1]])
    names = ('sync_pattern', 'user', 'internal_tvg', 'frame_nr', 'bcd_jday',
             'bcd_seconds', 'bcd_fraction', 'crc', 'jday', 'seconds',
             'fraction_ns')
    for i, row in enumerate(sample_outputs['sample_m5b_fields']):
        h = headers.mark5b_parse(raw[i * 10016:i * 10016 + 16].view('<u4'))
        for k, v in zip(names, row):
            assert h[k] == v, k
    batch = headers.mark5b_parse_batch(raw, 4)
    assert np.array_equal(batch['frame_nr'],
                          sample_outputs['sample_m5b_fields'][:, 3])
    assert np.array_equal(batch['fraction_ns'],
                          sample_outputs['sample_m5b_fields'][:, 10])


def test_mark5b_fill_frame(sample_outputs):
    raw = sample_bytes('sample.m5b')[:10016].copy()
    raw[16:].view('<u4')[:] = 0x11223344
    assert not sample_outputs['sample_m5b_fillframe_valid']
    data = stream.mark5b_read(raw, nchan=8, fill_value=-999.)
    assert np.array_equal(data, sample_outputs['sample_m5b_fillframe_data'])
    raw[16 + 4 * 1234] ^= 1       # one differing word -> valid again
    assert stream.mark5b_valid_mask(raw)[0]


@pytest.mark.parametrize('name,ntrack', [
    ('sample.m4', 64), ('sample_32track.m4', 32),
    ('sample_32track_fanout2.m4', 32), ('sample_16track.m4', 16),
    ('sample_64track_fanout2_ft.m4', 64)])
def test_mark4_samples(sample_outputs, name, ntrack):
    tag = name.replace('.', '_')
    raw = sample_bytes(name)
    off = int(sample_outputs[tag + '_offset0'])
    data = stream.mark4_read(raw, ntrack, fill_value=-7., offset0=off)
    assert np.array_equal(data, sample_outputs[tag + '_data'])
    geom = sample_outputs[tag + '_geom']
    dt = codec.MARK4_WORD_DTYPE[ntrack]
    # the golden track fields are those of the last complete frame
    last = off + (data.shape[0] // int(geom[4]) - 1) * ntrack * 2500
    hdr = headers.mark4_parse(raw[last:last + ntrack * 20].view(dt))
    assert [hdr[k] for k in ('ntrack', 'fanout', 'nchan', 'bps',
                             'samples_per_frame')] == list(geom)
    names = ('fan_out', 'magnitude_bit', 'lsb_output', 'converter_id',
             'bcd_unit_year', 'bcd_day', 'bcd_hour', 'bcd_minute',
             'bcd_second', 'bcd_fraction', 'crc', 'sync_pattern')
    for k, row in zip(names, sample_outputs[tag + '_track_fields']):
        assert np.array_equal(hdr[k], row), k


def test_mark4_known_values():
    # mark4/tests/test_mark4.py:324-327: first valid samples of sample.m4
    raw = sample_bytes('sample.m4')
    data = stream.mark4_read(raw, 64, offset0=0xa88, count=642)
    assert np.all(data[:640] == 0.)
    assert np.all(data[640:642].astype(int) == [
        [-1, 1, 1, -3, -3, -3, 1, -1], [1, 1, -3, 1, 1, -3, -1, -1]])


def test_guppi_sample(sample_outputs):
    raw = sample_bytes('sample_puppi.raw')
    frames = stream.guppi_scan(raw)
    geom = sample_outputs['sample_puppi_geom']
    h = frames[0]
    assert [len(frames), h['header_nbytes'], h['payload_nbytes'], h['npol'],
            h['nchan'], h['overlap'], h['samples_per_frame']] == list(geom)
    want = sample_outputs['sample_puppi_frames']
    for i, f in enumerate(frames):
        assert np.array_equal(stream.guppi_decode_frame(raw, f), want[i])
    # stream read with overlap: guppi/tests/test_guppi.py:471-495 semantics
    full = stream.guppi_read(raw)
    spf, ov = h['samples_per_frame'], h['overlap']
    assert full.shape[0] == (spf - ov) * len(frames) + ov
    assert np.array_equal(full[:spf], want[0])
    assert np.array_equal(full[spf:spf + spf - ov], want[1][ov:])
    # a read started inside frame 1 sees frame 1's own data throughout
    part = stream.guppi_read(raw, offset=(spf - ov) + 3, count=spf - 3)
    assert np.array_equal(part, want[1][3:])
    tf = codec.guppi_payload_decode(
        raw[frames[-1]['offset'] + h['header_nbytes']:][:h['payload_nbytes']]
        .view(np.int8), h['npol'], h['nchan'], True, channels_first=False)
    assert np.array_equal(tf, sample_outputs['sample_puppi_frame3_timefirst'])


@pytest.mark.parametrize('name', ['sample.dada', 'sample_meerkat.dada',
                                  'sample_mkbf.dada'])
def test_dada_samples(sample_outputs, name):
    tag = name.replace('.', '_')
    data = stream.dada_read(sample_bytes(name))
    want = sample_outputs[tag + '_data']
    assert data.shape == want.shape
    assert np.array_equal(data, want)


def test_dada_known_values():
    # dada/tests/test_dada.py:180-183
    data = stream.dada_read(sample_bytes('sample.dada'))
    assert data.shape == (16000, 2, 1)
    assert np.all(data[:3, :, 0] == np.array(
        [[-38 - 38j, -38 - 38j], [-38 - 38j, -40 + 0j], [-105 + 60j, 85 - 15j]],
        dtype=np.complex64))


def test_gsb_samples(sample_outputs):
    raw = sample_bytes('gsb/sample_gsb_rawdump.dat')
    data = stream.gsb_rawdump_read(raw, payload_nbytes=8192, nframe=1)
    assert np.array_equal(data, sample_outputs['gsb_rawdump_8192_data'])
    files = [[sample_bytes('gsb/sample_gsb_phased.Pol-%s%d.dat' % (p, i))
              for i in (1, 2)] for p in 'LR']
    got = stream.gsb_phased_read(files, nframe=5, payload_nbytes=8192)
    want = sample_outputs['gsb_phased_8192_frames']
    assert np.array_equal(got.reshape(want.shape), want)


def test_encode_round_trips(sample_outputs):
    """decode -> encode reproduces the payload bytes
    (vdif/tests/test_vdif.py:392-393, mark5b :188-189, mark4 :338-342)."""
    raw = sample_bytes('sample.vdif')
    words = raw[32:5032].view('<u4')
    d = codec.vdif_payload_decode(words, 2)
    assert np.array_equal(codec.vdif_payload_encode(d, 2), words)
    raw = sample_bytes('sample.m5b')
    words = raw[16:10016].view('<u4')
    d = codec.mark5b_payload_decode(words, 2, 8)
    assert np.array_equal(codec.mark5b_payload_encode(d, 2), words)
    raw = sample_bytes('sample.m4')
    words = raw[0xa88 + 1280:0xa88 + 160000].view('<u8')
    d = codec.mark4_decode(words, 8, 4)
    assert np.array_equal(codec.mark4_encode(d, 8, 4), words)


def test_locate_frames_oracle_known_answers():
    """oracle.locate.locate_frames against the answers the reference's own
    tests assert (vdif/tests/test_vdif.py:695-760, mark5b/tests/
    test_mark5b.py:489-506)."""
    from oracle import locate
    from conftest import sample_bytes
    data = sample_bytes('sample.vdif')
    size = data.size
    words = data[:32].view('<u4')
    sync = int(words[5])
    every = [x * 5032 for x in range(16)]
    assert sync == 0xACABFEED
    assert locate.locate_frames(data, 0, sync, offset=20) == every
    assert locate.locate_frames(data, size, sync, offset=20,
                                forward=False) == every[::-1]
    mask = [0, 0, 0xffffffff, 0xfc00ffff, 0xffffffff, 0, 0, 0]
    assert locate.locate_frames(data, 10, words, mask=mask,
                                frame_nbytes=5032) == [5032, 10064]
    # the header's invariants (what passing header0 means)
    inv = [0x40000000, 0, 0xffffffff, 0xfc00ffff, 0xffffffff, 0xffffffff,
           0, 0]
    for pos, fwd, want in ((5000, True, [5032, 10064]),
                           (15000, True, [15096, 20128]),
                           (20128, True, [20128, 25160]),
                           (16, False, [0]),
                           (size - 10000, False, [14 * 5032, 13 * 5032]),
                           (size - 5000, False, [15 * 5032, 14 * 5032]),
                           (size - 20, True, []),
                           (40254, True, [8 * 5032, 9 * 5032]),
                           (40254, False, [7 * 5032, 6 * 5032])):
        assert locate.locate_frames(data, pos, words, mask=inv,
                                    frame_nbytes=5032, forward=fwd) == want
    m5 = sample_bytes('sample.m5b')
    s5 = 0xABADDEED
    assert locate.locate_frames(m5, 0, s5, frame_nbytes=10016) == [0, 10016]
    assert locate.locate_frames(m5, 0, s5, frame_nbytes=10016,
                                forward=False) == [0]
    assert locate.locate_frames(m5, 10000, s5, frame_nbytes=10016) \
        == [10016, 20032]
    assert locate.locate_frames(m5, 16, s5, frame_nbytes=10016,
                                forward=False) == [0]
    assert locate.locate_frames(m5, m5.size - 20, s5, frame_nbytes=10016) == []
    assert locate.locate_frames(m5, m5.size - 10000, s5, frame_nbytes=10016,
                                forward=False) == [3 * 10016, 2 * 10016]
    assert locate.locate_frames(m5, m5.size - 30, s5, frame_nbytes=10016) == []
    # the reference's corrupted file (test_mark5b.py:508-524): bytes
    # [10040, 20000) cut out
    bad = np.concatenate([m5[:10040], m5[20000:]])
    shifted = 2 * 10016 - 9960
    assert locate.locate_frames(bad, 0, s5, frame_nbytes=10016) == [0, shifted]
    assert locate.locate_frames(bad, 0, s5, frame_nbytes=10016,
                                check=None) == [0, 10016, shifted]
    assert locate.locate_frames(bad, 10000, s5, frame_nbytes=10016) \
        == [shifted, shifted + 10016]
    assert locate.locate_frames(bad, 10000, s5, frame_nbytes=10016,
                                check=None) == [10016, shifted,
                                                shifted + 10016]
    short = m5[:10018]
    assert locate.locate_frames(short, 10, s5, frame_nbytes=10016) == []
    assert locate.locate_frames(short, 10, s5, frame_nbytes=10016,
                                forward=False) == [0]
