#!/usr/bin/env python
"""Generate the golden vectors in this directory from the UNMODIFIED reference.

Run in the build container (``/root/reference`` mounted, numpy only):

    python tests/golden/make_golden.py

Outputs (all committed):

* ``samples/``            the reference's small recorded sample files (inputs)
* ``codec_vectors.npz``   seeded inputs + outputs of every reference codec
                          function on the hot path (decode and encode, float32
                          and float64, values placed on and next to every
                          quantiser threshold)
* ``sample_outputs.npz``  decoded arrays, validity and header fields obtained
                          by running the reference Payload/Frame/Header classes
                          on the sample files

The reference modules are imported by path with a stand-in for the few
``astropy.utils`` helpers they need (``ref_loader.py``); no reference source is
copied.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

SAMPLES = [
    'sample.vdif', 'sample_vlbi.vdif', 'sample_mwa.vdif',
    'sample_arochime.vdif', 'sample_bps1.vdif', 'sample.m5b', 'sample.m4',
    'sample_32track.m4', 'sample_32track_fanout2.m4', 'sample_16track.m4',
    'sample_64track_fanout2_ft.m4', 'sample_puppi.raw', 'sample.dada',
    'sample_meerkat.dada', 'sample_mkbf.dada',
    'gsb/sample_gsb_rawdump.dat', 'gsb/sample_gsb_rawdump.timestamp',
    'gsb/sample_gsb_phased.Pol-L1.dat', 'gsb/sample_gsb_phased.Pol-L2.dat',
    'gsb/sample_gsb_phased.Pol-R1.dat', 'gsb/sample_gsb_phased.Pol-R2.dat',
    'gsb/sample_gsb_phased.timestamp']


def threshold_values(dtype):
    """Values on, just below and just above every quantiser threshold."""
    enc = ref.encoding
    s = enc.TWO_BIT_1_SIGMA
    marks = [0.0, -0.0, s, -s, 1.5 * s, -1.5 * s, 2 * s, -2 * s, 1.0, -1.0,
             enc.OPTIMAL_2BIT_HIGH, -enc.OPTIMAL_2BIT_HIGH, 1e-30, -1e-30,
             1e30, -1e30, np.inf, -np.inf]
    marks += [(k - 8 + h) / enc.FOUR_BIT_1_SIGMA for k in range(17)
              for h in (0.0, 0.5, -0.5)]
    marks += [(k - 127.5 + h) / enc.EIGHT_BIT_1_SIGMA
              for k in (0, 1, 2, 127, 128, 254, 255, 256) for h in (0., .5)]
    marks += [k + h for k in range(-130, 131, 1) for h in (0.0, 0.5)]
    marks = np.array(marks, dtype=dtype)
    out = [marks]
    up, dn = marks.copy(), marks.copy()
    for _ in range(3):
        up = np.nextafter(up, dtype(np.inf))
        dn = np.nextafter(dn, dtype(-np.inf))
        out += [up.copy(), dn.copy()]
    # The float32 thresholds seen from float64 and vice versa.
    other = np.float64 if dtype is np.float32 else np.float32
    out.append(np.array(marks, dtype=other).astype(dtype))
    return np.concatenate(out)


def encode_inputs(dtype, rng, n):
    vals = np.concatenate([
        threshold_values(dtype),
        (rng.standard_normal(n) * 2.5).astype(dtype),
        (rng.standard_normal(n) * 40).astype(dtype),
        rng.uniform(-9, 9, n).astype(dtype)])
    pad = (-vals.size) % 64
    return np.concatenate([vals, np.zeros(pad, dtype)])


def codec_vectors():
    rng = np.random.default_rng(20240531)
    out = {}
    enc, vp, m5, m4 = (ref.encoding, ref.vdif_payload, ref.mark5b_payload,
                       ref.mark4_payload)
    out['levels1'], out['levels2'], out['levels4'] = (
        enc.decoder_levels[1], enc.decoder_levels[2], enc.decoder_levels[4])
    out['vdif_lut1'], out['vdif_lut2'], out['vdif_lut4'] = (
        vp.lut1bit, vp.lut2bit, vp.lut4bit)
    out['m5b_lut1'], out['m5b_lut2'] = m5.lut1bit, m5.lut2bit
    out['m4_lut1'], out['m4_lut2_1'], out['m4_lut2_2'], out['m4_lut2_3'] = (
        m4.lut1bit, m4.lut2bit1, m4.lut2bit2, m4.lut2bit3)
    # Every byte value in every position, then random words.
    words = np.concatenate([
        np.arange(256, dtype=np.uint8).repeat(4).view('<u4'),
        np.tile(np.arange(256, dtype=np.uint8), 4).view('<u4'),
        rng.integers(0, 2**32, 512, dtype=np.uint64).astype('<u4')])
    out['words32'] = words
    for bps, fn in ((1, vp.decode_1bit), (2, vp.decode_2bit),
                    (4, vp.decode_4bit), (8, enc.decode_8bit)):
        out['vdif_dec%d' % bps] = np.ascontiguousarray(fn(words)).ravel()
    for bps, fn in ((1, m5.decode_1bit), (2, m5.decode_2bit)):
        out['m5b_dec%d' % bps] = np.ascontiguousarray(fn(words)).ravel()
    for dtype, tag in ((np.float32, 'f32'), (np.float64, 'f64')):
        vals = encode_inputs(dtype, rng, 4096)
        out['enc_in_' + tag] = vals
        finite = np.where(np.isfinite(vals), vals, dtype(0.))
        out['enc_in_finite_' + tag] = finite
        for bps, fn in ((1, vp.encode_1bit), (2, vp.encode_2bit),
                        (4, vp.encode_4bit), (8, enc.encode_8bit)):
            out['vdif_enc%d_%s' % (bps, tag)] = np.ascontiguousarray(
                fn(vals.copy())).ravel().view(np.uint8)
        for bps, fn in ((1, m5.encode_1bit), (2, m5.encode_2bit)):
            out['m5b_enc%d_%s' % (bps, tag)] = np.ascontiguousarray(
                fn(vals.copy())).ravel().view(np.uint8)
        out['int8_enc_' + tag] = ref.guppi_payload.encode_8bit(
            finite.copy()).view(np.uint8)
        out['gsb4_enc_' + tag] = ref.gsb_payload.encode_4bit(
            finite.copy()).view(np.uint8)
    # NaN handling of the 2/4/8-bit encoders (implementation-defined cast;
    # recorded as observed on x86-64 numpy).
    nanv = np.array([np.nan, 1.0, -1.0, np.nan] * 8, np.float32)
    out['nan_in'] = nanv
    with np.errstate(invalid='ignore'):
        out['vdif_enc2_nan'] = vp.encode_2bit(nanv.copy())
        out['vdif_enc1_nan'] = vp.encode_1bit(nanv.copy())
        out['m5b_enc1_nan'] = m5.encode_1bit(nanv.copy())
    # Mark 4.
    out['m4_reorder64_in'] = np.array([738811025863578102], np.uint64)
    out['m4_reorder64_out'] = m4.reorder64(out['m4_reorder64_in'])
    out['m4_reorder32'] = m4.reorder32(words.view(np.uint32))
    out['m4_reorder64'] = m4.reorder64(words.view(np.uint64))
    out['m4_reorder64_ft'] = m4.reorder64_Ft(words.view(np.uint64))
    modes = {'2_4': (2, 4, '<u2', m4.decode_2chan_2bit_fanout4,
                     m4.encode_2chan_2bit_fanout4),
             '4_4': (4, 4, '<u4', m4.decode_4chan_2bit_fanout4,
                     m4.encode_4chan_2bit_fanout4),
             '8_2': (8, 2, '<u4', m4.decode_8chan_2bit_fanout2,
                     m4.encode_8chan_2bit_fanout2),
             '8_4': (8, 4, '<u8', m4.decode_8chan_2bit_fanout4,
                     m4.encode_8chan_2bit_fanout4),
             '16_2ft': (16, 2, '<u8', m4.decode_16chan_2bit_fanout2_ft,
                        m4.encode_16chan_2bit_fanout2_ft)}
    for tag, (nchan, fanout, dt, dec, encf) in modes.items():
        w = words.view(dt)
        # one-hot words pin every bit position
        nbit = np.dtype(dt).itemsize * 8
        onehot = (np.uint64(1) << np.arange(nbit, dtype=np.uint64)).astype(dt)
        w = np.concatenate([onehot, w])
        out['m4_words_' + tag] = w
        d = np.ascontiguousarray(dec(w))
        out['m4_dec_' + tag] = d
        for dtype, ftag in ((np.float32, 'f32'), (np.float64, 'f64')):
            vals = encode_inputs(dtype, rng, 1024)
            vals = vals[:vals.size // (nchan * 4) * nchan * 4].reshape(
                -1, nchan)
            out['m4_enc_in_%s_%s' % (tag, ftag)] = vals
            out['m4_enc_%s_%s' % (tag, ftag)] = np.ascontiguousarray(
                encf(vals.copy())).ravel().view(np.uint8)
    # int8 / nibble decoders.
    b = np.concatenate([np.arange(256, dtype=np.uint8),
                        rng.integers(0, 256, 768, dtype=np.uint8)]
                       ).view(np.int8)
    out['bytes'] = b
    out['int8_dec'] = ref.guppi_payload.decode_8bit(b)
    out['gsb4_dec'] = ref.gsb_payload.decode_4bit(b)
    out['gsb8_dec'] = ref.gsb_payload.decode_8bit(b)
    np.savez_compressed(os.path.join(HERE, 'codec_vectors.npz'), **out)
    print('codec_vectors.npz: %d arrays' % len(out))


class _Hdr:
    """Just the attributes PayloadBase.__init__ reads from a header."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def get(self, key, default=None):
        return self.__dict__.get(key, default)


def find_mark4_frame(buf, ntrack):
    """First byte offset whose header words 64..95 are all ones."""
    wsize = ntrack // 8
    run = 32 * wsize
    ones = (buf == 0xff).astype(np.int32)
    csum = np.concatenate([[0], np.cumsum(ones)])
    full = np.nonzero(csum[run:] - csum[:-run] == run)[0]
    for start in full:
        off = start - 64 * wsize
        if off >= 0:
            return int(off)
    raise ValueError('no Mark 4 sync found')


def sample_outputs():
    from oracle import stream as ostream      # only its text-header parsers
    out = {}
    vf, vh = ref.vdif_frame, ref.vdif_header
    # ---- VDIF ------------------------------------------------------------
    for name in ('sample.vdif', 'sample_vlbi.vdif', 'sample_mwa.vdif',
                 'sample_arochime.vdif', 'sample_bps1.vdif'):
        tag = name.replace('.', '_')
        size = os.path.getsize(ref_loader.sample(name))
        datas, fields = [], []
        with open(ref_loader.sample(name), 'rb') as fh:
            # per-frame header fields, in file order
            while fh.tell() < size:
                h = vh.VDIFHeader.fromfile(fh)
                fields.append([int(h[k]) for k in (
                    'invalid_data', 'legacy_mode', 'seconds', 'ref_epoch',
                    'frame_nr', 'vdif_version', 'lg2_nchan', 'frame_length',
                    'complex_data', 'bits_per_sample', 'thread_id',
                    'station_id')] + [h.edv if h.edv else 0,
                                      h.payload_nbytes, h.samples_per_frame])
                fh.seek(h.payload_nbytes, 1)
            fh.seek(0)
            while fh.tell() < size:
                try:
                    fs = vf.VDIFFrameSet.fromfile(fh)
                except EOFError:
                    break
                datas.append(fs.data)
                tids = fs['thread_id']
        out[tag + '_fields'] = np.array(fields, dtype=np.int64)
        out[tag + '_data'] = np.concatenate(datas)
        out[tag + '_thread_ids'] = np.asarray(tids)
    # invalid-frame fill through the reference frame set
    with open(ref_loader.sample('sample.vdif'), 'rb') as fh:
        fs = vf.VDIFFrameSet.fromfile(fh)
    for f in fs.frames[1::3]:
        f.header.mutable = True
        f.valid = False
    fs.fill_value = -999.
    out['sample_vdif_set0_invalid_1_4_7_fill_m999'] = fs.data
    # ---- Mark 5B ---------------------------------------------------------
    mf = ref.mark5b_frame
    datas, fields, valid = [], [], []
    with open(ref_loader.sample('sample.m5b'), 'rb') as fh:
        for _ in range(4):
            fr = mf.Mark5BFrame.fromfile(fh, kday=56000, sample_shape=(8,),
                                         bps=2)
            datas.append(fr.data)
            valid.append(fr.valid)
            h = fr.header
            fields.append([int(h[k]) for k in (
                'sync_pattern', 'user', 'internal_tvg', 'frame_nr',
                'bcd_jday', 'bcd_seconds', 'bcd_fraction', 'crc')]
                + [h.jday, h.seconds, int(round(h.fraction * 1e9))])
    out['sample_m5b_data'] = np.concatenate(datas)
    out['sample_m5b_valid'] = np.array(valid)
    out['sample_m5b_fields'] = np.array(fields, dtype=np.int64)
    # a frame made invalid by the fill pattern
    raw = np.fromfile(ref_loader.sample('sample.m5b'), np.uint8)[:10016].copy()
    raw[16:].view('<u4')[:] = 0x11223344
    import io
    fr = mf.Mark5BFrame.fromfile(io.BytesIO(raw.tobytes()), kday=56000,
                                 sample_shape=(8,), bps=2)
    fr.fill_value = -999.
    out['sample_m5b_fillframe_valid'] = np.array(fr.valid)
    out['sample_m5b_fillframe_data'] = fr.data
    # ---- Mark 4 ----------------------------------------------------------
    m4f = ref.mark4_frame
    for name, ntrack in (('sample.m4', 64), ('sample_32track.m4', 32),
                         ('sample_32track_fanout2.m4', 32),
                         ('sample_16track.m4', 16),
                         ('sample_64track_fanout2_ft.m4', 64)):
        tag = name.replace('.', '_')
        buf = np.fromfile(ref_loader.sample(name), np.uint8)
        off = find_mark4_frame(buf, ntrack)
        frame_nbytes = ntrack * 2500
        nframe = (buf.size - off) // frame_nbytes
        datas = []
        with open(ref_loader.sample(name), 'rb') as fh:
            fh.seek(off)
            for _ in range(nframe):
                fr = m4f.Mark4Frame.fromfile(fh, ntrack=ntrack, decade=2010)
                fr.fill_value = -7.
                datas.append(fr.data)
        out[tag + '_offset0'] = np.array(off)
        out[tag + '_data'] = np.concatenate(datas)
        h = fr.header
        out[tag + '_geom'] = np.array([ntrack, h.fanout, h.nchan, h.bps,
                                       h.samples_per_frame])
        out[tag + '_track_fields'] = np.array(
            [np.asarray(h[k]).astype(np.int64) for k in (
                'fan_out', 'magnitude_bit', 'lsb_output', 'converter_id',
                'bcd_unit_year', 'bcd_day', 'bcd_hour', 'bcd_minute',
                'bcd_second', 'bcd_fraction', 'crc', 'sync_pattern')])
    # ---- GUPPI -----------------------------------------------------------
    buf = np.fromfile(ref_loader.sample('sample_puppi.raw'), np.uint8)
    frames = ostream.guppi_scan(buf)
    datas = []
    for h in frames:
        hdr = _Hdr(sample_shape=(h['npol'], h['nchan']), bps=h['bps'],
                   complex_data=h['complex_data'],
                   payload_nbytes=h['payload_nbytes'],
                   channels_first=h['channels_first'])
        p0 = h['offset'] + h['header_nbytes']
        pl = ref.guppi_payload.GUPPIPayload(
            buf[p0:p0 + h['payload_nbytes']].view(np.int8), header=hdr)
        datas.append(pl.data)
    out['sample_puppi_frames'] = np.array(datas)
    out['sample_puppi_geom'] = np.array(
        [len(frames), frames[0]['header_nbytes'], frames[0]['payload_nbytes'],
         frames[0]['npol'], frames[0]['nchan'], frames[0]['overlap'],
         frames[0]['samples_per_frame']])
    # time-first variant of the same bytes
    hdr.channels_first = False
    out['sample_puppi_frame3_timefirst'] = ref.guppi_payload.GUPPIPayload(
        buf[p0:p0 + h['payload_nbytes']].view(np.int8), header=hdr).data
    # ---- DADA ------------------------------------------------------------
    for name in ('sample.dada', 'sample_meerkat.dada', 'sample_mkbf.dada'):
        tag = name.replace('.', '_')
        buf = np.fromfile(ref_loader.sample(name), np.uint8)
        h = ostream.dada_parse_header(buf)
        p0 = h['header_nbytes']
        # short last frame: dada/base.py:277-332
        nbytes = min(h['payload_nbytes'], buf.size - p0)
        block = np.lcm(4, h['bps'] * (2 if h['complex_data'] else 1)
                       * h['npol'] * h['nchan'] // 8)
        h['payload_nbytes'] = int(nbytes // block * block)
        hdr = _Hdr(sample_shape=(h['npol'], h['nchan']), bps=h['bps'],
                   complex_data=h['complex_data'],
                   payload_nbytes=h['payload_nbytes'],
                   INSTRUMENT=h.get('INSTRUMENT'))
        pl = ref.dada_payload.DADAPayload(
            buf[p0:p0 + h['payload_nbytes']].view('<u4'), header=hdr)
        out[tag + '_data'] = pl.data
        out[tag + '_geom'] = np.array([h['header_nbytes'],
                                       h['payload_nbytes'], h['npol'],
                                       h['nchan'], int(h['complex_data']),
                                       int(type(pl).__name__ == 'MKBFPayload')
                                       ])
    # ---- GSB -------------------------------------------------------------
    gp = ref.gsb_payload.GSBPayload
    with open(ref_loader.sample('gsb/sample_gsb_rawdump.dat'), 'rb') as fh:
        pl = gp.fromfile(fh, payload_nbytes=8192, sample_shape=(1,), bps=4)
        out['gsb_rawdump_8192_data'] = pl.data
    names = [['gsb/sample_gsb_phased.Pol-L1.dat',
              'gsb/sample_gsb_phased.Pol-L2.dat'],
             ['gsb/sample_gsb_phased.Pol-R1.dat',
              'gsb/sample_gsb_phased.Pol-R2.dat']]
    fhs = [[open(ref_loader.sample(n), 'rb') for n in pair] for pair in names]
    frames = []
    for _ in range(5):
        pl = gp.fromfile(fhs, payload_nbytes=8192, sample_shape=(2, 512),
                         bps=8, complex_data=True)
        frames.append(pl.data)
    out['gsb_phased_8192_frames'] = np.array(frames)
    for pair in fhs:
        for fh in pair:
            fh.close()
    np.savez_compressed(os.path.join(HERE, 'sample_outputs.npz'), **out)
    print('sample_outputs.npz: %d arrays' % len(out))


def copy_samples():
    dest = os.path.join(HERE, 'samples')
    os.makedirs(os.path.join(dest, 'gsb'), exist_ok=True)
    for name in SAMPLES:
        shutil.copyfile(ref_loader.sample(name), os.path.join(dest, name))
    print('copied %d sample files' % len(SAMPLES))


if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    ref = ref_loader.load_reference()
    copy_samples()
    codec_vectors()
    sample_outputs()
