"""Oracle frame assembly: raw bytes -> decoded sample arrays.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates the per-frame Python loops of the reference stream readers for
*clean, contiguous* files (corrupt-file recovery is host control flow and out
of scope, SURVEY.md section 8(a) footnote):

* baseband/base/base.py:919-1018          read loop, _get_frame
* baseband/base/frame.py:191-199          invalid frame -> fill_value
* baseband/vdif/frame.py:175-243, :402-434  frame-set grouping, thread axis
* baseband/vdif/base.py:172-215, :441-490   thread ids
* baseband/mark5b/frame.py:62-72          validity from the fill pattern
* baseband/mark4/frame.py:148-263         header-overwritten samples -> fill
* baseband/guppi/base.py:203-221, :270-278  overlap
* baseband/dada/base.py:277-332           short last frame
* baseband/gsb/base.py:373-387, gsb/payload.py:88-131  multi-file interleave

All sample arithmetic goes through oracle.codec.
"""
import numpy as np

from . import codec, headers


def _as_bytes(raw):
    if isinstance(raw, (bytes, bytearray, memoryview)):
        return np.frombuffer(raw, np.uint8)
    return np.ascontiguousarray(raw).view(np.uint8).ravel()


def _fill(shape, fill_value, complex_data):
    return np.full(shape, fill_value,
                   np.complex64 if complex_data else np.float32)


# ------------------------------------------------------------------ VDIF
def vdif_scan(raw):
    """Parse every frame header of a clean VDIF byte stream.

    Returns (header0 dict, list of per-frame dicts with 'offset')."""
    buf = _as_bytes(raw)
    h0 = headers.vdif_parse(buf[:32].view('<u4'))
    frames = []
    pos = 0
    while pos + h0['header_nbytes'] <= buf.size:
        hw = buf[pos:pos + h0['header_nbytes']].view('<u4')
        h = headers.vdif_parse(hw)
        if pos + h['frame_nbytes'] > buf.size:
            break
        h['offset'] = pos
        frames.append(h)
        pos += h['frame_nbytes']
    return h0, frames


def vdif_framesets(frames, thread_ids=None):
    """Group frames as VDIFFrameSet.fromfile does (vdif/frame.py:201-243):
    consecutive frames with the same frame_nr and not-yet-seen thread_id;
    threads outside ``thread_ids`` are skipped; order by ``thread_ids`` or
    sorted id."""
    sets = []
    i = 0
    while i < len(frames):
        frame_nr = frames[i]['frame_nr']
        seen = {}
        while i < len(frames):
            h = frames[i]
            tid = h['thread_id']
            if h['frame_nr'] != frame_nr or tid in seen:
                break
            seen[tid] = h
            i += 1
        wanted = sorted(seen) if thread_ids is None else list(thread_ids)
        if all(t in seen for t in wanted):
            sets.append([seen[t] for t in wanted])
    return sets


def vdif_decode_frame(buf, h, fill_value=0.0):
    """One VDIFFrame.data: (samples_per_frame, nchan)."""
    start = h['offset'] + h['header_nbytes']
    words = buf[start:start + h['payload_nbytes']].view('<u4')
    shape = (h['samples_per_frame'], h['nchan'])
    if h['invalid_data']:                       # vdif/frame.py:79-90
        return _fill(shape, fill_value, h['complex_data'])
    return codec.vdif_payload_decode(
        words, h['bps'], (h['nchan'],), bool(h['complex_data']),
        mark5b=(h.get('edv') == 0xab))


def vdif_read(raw, thread_ids=None, fill_value=0.0, offset=0, count=None):
    """``vdif.open(..., 'rs', squeeze=False).read()``: (nsample, nthread,
    nchan).  ``thread_ids`` mimics ``subset`` on the thread axis."""
    buf = _as_bytes(raw)
    h0, frames = vdif_scan(buf)
    if thread_ids is None:
        # vdif/base.py:172-215: ids present in the first frame sets, sorted.
        first = vdif_framesets(frames[:len({f['thread_id'] for f in frames})
                                      * 2])
        thread_ids = sorted({h['thread_id'] for s in first[:1] for h in s})
    sets = vdif_framesets(frames, thread_ids)
    spf = h0['samples_per_frame']
    total = len(sets) * spf
    if count is None:
        count = total - offset
    if offset + count > total:
        raise EOFError("cannot read from beyond end of input.")
    cplx = bool(h0['complex_data'])
    out = np.empty((count, len(thread_ids), h0['nchan']),
                   np.complex64 if cplx else np.float32)
    done = 0
    while done < count:                        # base/base.py:957-967
        index, start = divmod(offset + done, spf)
        n = min(count - done, spf - start)
        for slot, h in enumerate(sets[index]):  # vdif/frame.py:427-434
            out[done:done + n, slot] = vdif_decode_frame(
                buf, h, fill_value)[start:start + n]
        done += n
    return out


# ------------------------------------------------------------------ Mark 5B
M5B_FRAME = 10016
M5B_HEADER = 16


def mark5b_read(raw, nchan, bps=2, fill_value=0.0, offset=0, count=None):
    """``mark5b.open(..., 'rs', nchan=..., squeeze=False).read()``."""
    buf = _as_bytes(raw)
    nframe = buf.size // M5B_FRAME
    spf = 10000 * 8 // (bps * nchan)
    total = nframe * spf
    if count is None:
        count = total - offset
    if offset + count > total:
        raise EOFError("cannot read from beyond end of input.")
    out = np.empty((count, nchan), np.float32)
    done = 0
    while done < count:
        index, start = divmod(offset + done, spf)
        n = min(count - done, spf - start)
        p0 = index * M5B_FRAME + M5B_HEADER
        words = buf[p0:p0 + 10000].view('<u4')
        if codec.mark5b_payload_valid(words):
            data = codec.mark5b_payload_decode(words, bps, nchan)
        else:
            data = _fill((spf, nchan), fill_value, False)
        out[done:done + n] = data[start:start + n]
        done += n
    return out


def mark5b_valid_mask(raw):
    buf = _as_bytes(raw)
    nframe = buf.size // M5B_FRAME
    return np.array([codec.mark5b_payload_valid(
        buf[i * M5B_FRAME + 16:(i + 1) * M5B_FRAME].view('<u4'))
        for i in range(nframe)])


# ------------------------------------------------------------------ Mark 4
def mark4_frame_decode(frame_bytes, ntrack, fill_value=0.0):
    """One Mark4Frame.data (mark4/frame.py:239-263): samples overwritten by
    the header are fill; an invalid frame is all fill."""
    dtype = codec.MARK4_WORD_DTYPE[ntrack]
    stream = np.ascontiguousarray(frame_bytes).view(dtype)
    hdr = headers.mark4_parse(stream[:headers.MARK4_HEADER_STEPS])
    nchan, fanout, spf = hdr['nchan'], hdr['fanout'], hdr['samples_per_frame']
    out = _fill((spf, nchan), fill_value, False)
    if hdr['valid']:
        ft = _mark4_is_ft(hdr)
        body = codec.mark4_decode(stream[headers.MARK4_HEADER_STEPS:],
                                  nchan, fanout, ft)
        out[spf - body.shape[0]:] = body
    return out, hdr


def _mark4_is_ft(hdr):
    """Non-standard magnitude-bit layout -> Fortaleza decoder
    (mark4/payload.py:349-356, :337)."""
    if hdr['bps'] != 2:
        return False
    packed = int(np.packbits(hdr['magnitude_bit'].astype(bool)).view(
        codec.MARK4_WORD_DTYPE[hdr['ntrack']])[0])
    return packed == codec.M4_FT_MAGBITS


def mark4_read(raw, ntrack, fill_value=0.0, offset0=0, offset=0, count=None):
    """``mark4.open(..., 'rs', ntrack=..., squeeze=False).read()`` for a file
    whose first frame starts at byte ``offset0``."""
    buf = _as_bytes(raw)[offset0:]
    frame_nbytes = ntrack * headers.MARK4_FRAME_STEPS // 8
    nframe = buf.size // frame_nbytes
    first, hdr0 = mark4_frame_decode(buf[:frame_nbytes], ntrack, fill_value)
    spf = hdr0['samples_per_frame']
    total = nframe * spf
    if count is None:
        count = total - offset
    if offset + count > total:
        raise EOFError("cannot read from beyond end of input.")
    out = np.empty((count, hdr0['nchan']), np.float32)
    done = 0
    while done < count:
        index, start = divmod(offset + done, spf)
        n = min(count - done, spf - start)
        data, _ = mark4_frame_decode(
            buf[index * frame_nbytes:(index + 1) * frame_nbytes], ntrack,
            fill_value)
        out[done:done + n] = data[start:start + n]
        done += n
    return out


# ------------------------------------------------------------------ GUPPI
def guppi_parse_header(buf, pos=0):
    """80-char cards up to END (guppi/header.py:168-186); returns (dict,
    header_nbytes) with the DIRECTIO padding of guppi/header.py:216-224."""
    cards = {}
    ncard = 0
    while True:
        line = bytes(buf[pos + 80 * ncard:pos + 80 * (ncard + 1)]).decode(
            'ascii')
        if line[:3] == 'END':
            break
        if len(line) < 80:
            raise EOFError
        key = line[:8].strip()
        val = line[9:].split('/')[0].strip() if line[8] == '=' else ''
        if val.startswith("'"):
            val = val.strip("'").strip()
        cards[key] = val
        ncard += 1
    nbytes = (ncard + 1) * 80
    if int(cards.get('DIRECTIO', '0')) and nbytes % 512:
        nbytes += 512 - nbytes % 512
    h = dict(cards)
    h['header_nbytes'] = nbytes
    h['payload_nbytes'] = int(cards['BLOCSIZE'])
    h['bps'] = int(cards['NBITS'])
    h['nchan'] = int(cards['OBSNCHAN'])
    h['complex_data'] = h['nchan'] != 1                 # header.py:276-278
    h['npol'] = int(cards['NPOL']) // (2 if h['complex_data'] else 1)
    h['bpcs'] = h['nchan'] * int(cards['NPOL']) * h['bps']
    h['samples_per_frame'] = h['payload_nbytes'] * 8 // h['bpcs']
    h['overlap'] = int(cards.get('OVERLAP', 0))
    h['channels_first'] = cards.get('PKTFMT', '1SFA') != 'SIMPLE'
    return h


def guppi_scan(raw):
    buf = _as_bytes(raw)
    frames = []
    pos = 0
    while pos < buf.size:
        h = guppi_parse_header(buf, pos)
        h['offset'] = pos
        if pos + h['header_nbytes'] + h['payload_nbytes'] > buf.size:
            break
        frames.append(h)
        pos += h['header_nbytes'] + h['payload_nbytes']
    return frames


def guppi_decode_frame(buf, h):
    p0 = h['offset'] + h['header_nbytes']
    words = buf[p0:p0 + h['payload_nbytes']].view(np.int8)
    return codec.guppi_payload_decode(words, h['npol'], h['nchan'],
                                      h['complex_data'], h['channels_first'])


def guppi_read(raw, offset=0, count=None):
    """``guppi.open(..., 'rs', squeeze=False).read()`` incl. the overlap
    rule: a read takes [start, len) of its first frame, then
    [overlap, len) of each later one (base/base.py:957-967 with
    guppi/base.py:203-206, :270-278)."""
    buf = _as_bytes(raw)
    frames = guppi_scan(buf)
    h0, hl = frames[0], frames[-1]
    stride = h0['samples_per_frame'] - h0['overlap']
    total = stride * (len(frames) - 1) + hl['samples_per_frame']
    if count is None:
        count = total - offset
    if offset + count > total:
        raise EOFError("cannot read from beyond end of input.")
    cplx = h0['complex_data']
    out = np.empty((count, h0['npol'], h0['nchan']),
                   np.complex64 if cplx else np.float32)
    normal_end = total - hl['overlap']
    pos, done = offset, 0
    cache = (None, None)
    while done < count:
        if normal_end <= pos < total:          # guppi/base.py:270-278
            index, start = divmod(normal_end - 1, stride)
            start += 1 + pos - normal_end
        else:
            index, start = divmod(pos, stride)
        if cache[0] != index:
            cache = (index, guppi_decode_frame(buf, frames[index]))
        data = cache[1]
        n = min(count - done, data.shape[0] - start)
        out[done:done + n] = data[start:start + n]
        done += n
        pos = offset + done
    return out


# ------------------------------------------------------------------ DADA
def dada_parse_header(buf, pos=0):
    """ASCII key/value lines (dada/header.py:117-200)."""
    hdr_size = 4096
    text = bytes(buf[pos:pos + hdr_size]).split(b'\x00')[0].decode('ascii')
    h = {}
    for line in text.split('\n'):
        if line.startswith('#') and 'end of header' in line:
            break
        body = line.split('#')[0].split()
        if len(body) >= 2:
            h[body[0]] = body[1]
    out = dict(h)
    out['header_nbytes'] = int(h.get('HDR_SIZE', 4096))
    out['payload_nbytes'] = int(h['FILE_SIZE'])
    out['bps'] = int(h['NBIT'])
    out['complex_data'] = int(h['NDIM']) == 2
    out['npol'] = int(h['NPOL'])
    out['nchan'] = int(h['NCHAN'])
    out['mkbf'] = h.get('INSTRUMENT') == 'MKBF'          # dada/payload.py:46-50
    return out


def dada_read(raw, offset=0, count=None):
    """``dada.open(..., 'rs', squeeze=False).read()`` for one file, with a
    possibly short last frame (dada/base.py:277-332)."""
    buf = _as_bytes(raw)
    h = dada_parse_header(buf)
    bytes_per_sample = (h['bps'] * (2 if h['complex_data'] else 1)
                        * h['npol'] * h['nchan']) // 8
    frame_nbytes = h['header_nbytes'] + h['payload_nbytes']
    nfull, rest = divmod(buf.size, frame_nbytes)
    pieces = []
    for i in range(nfull + (1 if rest > h['header_nbytes'] else 0)):
        p0 = i * frame_nbytes + h['header_nbytes']
        nbytes = min(h['payload_nbytes'], buf.size - p0)
        block = np.lcm(4, bytes_per_sample)
        nbytes = nbytes // block * block
        words = buf[p0:p0 + nbytes].view(np.int8)
        shape = (h['npol'], h['nchan'])
        if h['mkbf']:
            pieces.append(codec.mkbf_payload_decode(words, shape,
                                                    h['complex_data']))
        else:
            pieces.append(codec.dada_payload_decode(words, shape,
                                                    h['complex_data']))
    data = np.concatenate(pieces)
    if count is None:
        count = data.shape[0] - offset
    if offset + count > data.shape[0]:
        raise EOFError("cannot read from beyond end of input.")
    return data[offset:offset + count]


# ------------------------------------------------------------------ GSB
def gsb_rawdump_read(raw, payload_nbytes=1 << 22, bps=4, nframe=None,
                     offset=0, count=None):
    """GSB rawdump: one headerless file of packed nibbles
    (gsb/base.py:373-387): (nsample, 1) float32."""
    buf = _as_bytes(raw)
    if nframe is not None:
        buf = buf[:nframe * payload_nbytes]
    data = codec.gsb_payload_decode(buf.view(np.int8), bps, (1,), False)
    if count is None:
        count = data.shape[0] - offset
    return data[offset:offset + count]


def gsb_phased_read(raw_files, nframe, payload_nbytes, nchan=512, bps=8,
                    offset=0, count=None):
    """GSB phased: ``raw_files[pol][part]`` byte arrays; per frame the parts
    are interleaved as in gsb/payload.py:115-131; result (nsample, npol,
    nchan) complex64."""
    nthread = len(raw_files)
    sample_nbytes = nchan * 2 * bps // 8
    pieces = []
    for i in range(nframe):
        parts = [[_as_bytes(f)[i * payload_nbytes:(i + 1) * payload_nbytes]
                  .view(np.int8) for f in pol] for pol in raw_files]
        words = codec.gsb_interleave_files(parts, nthread, sample_nbytes)
        pieces.append(codec.gsb_payload_decode(words, bps, (nthread, nchan),
                                               True))
    data = np.concatenate(pieces)
    if count is None:
        count = data.shape[0] - offset
    return data[offset:offset + count]
