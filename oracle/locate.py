"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's sync
search and of the frame index it implies.  Only tests/, smoke() and bench.py's
cpu_baseline leg may import this.

`locate_frames` follows baseband/base/base.py:181-335 step by step, with the
file replaced by a byte array and the file pointer by ``position``.  It is
pinned against the known answers the reference's own tests assert for
sample.vdif and sample.m5b (vdif/tests/test_vdif.py:695-760,
mark5b/tests/test_mark5b.py:489-533) in tests/test_oracle_golden.py.
"""
import numpy as np


def byte_array(pattern):
    """baseband/base/utils.py:251-270 (`byte_array`)."""
    if isinstance(pattern, np.ma.MaskedArray):
        raise TypeError('pass mask separately')
    if isinstance(pattern, (bytes, bytearray)):
        return np.frombuffer(pattern, 'u1')
    arr = np.asarray(pattern)
    if arr.dtype.kind in 'iu' and arr.dtype.itemsize != 1:
        arr = np.atleast_1d(arr).astype('<u4')
    return np.atleast_1d(arr).view('u1')


def locate_frames(data, position, pattern, mask=None, frame_nbytes=None,
                  offset=0, forward=True, maximum=None, check=1):
    """baseband/base/base.py:227-335."""
    data = np.asarray(data, 'u1')
    file_size = data.size
    pattern = byte_array(pattern)
    if mask is not None:                                   # :232-238
        mask = byte_array(mask)
        useful = np.nonzero(mask)[0]
        sl = slice(useful[0], useful[-1] + 1)
        mask, pattern = mask[sl], pattern[sl]
        offset += sl.start
    if maximum is None:                                    # :240-241
        maximum = (2 * frame_nbytes if frame_nbytes else 1000000) - 1
    if check is None or frame_nbytes is None:              # :243-252
        check = np.array([], dtype=int)
        check_min = check_max = 0
    else:
        check = np.atleast_1d(check) * frame_nbytes
        check_min = min(check.min(), 0)
        check_max = max(check.max(), 0)
    if frame_nbytes is None:                               # :254-257
        frame_nbytes = offset + pattern.size
    seek_start = position if forward else position - maximum   # :261-264
    start = max(seek_start + offset + check_min, 0)        # :272-274
    stop = max(seek_start + maximum + 1 + check_max + frame_nbytes, start)
    block = data[start:min(stop, file_size)]               # :276-277
    stop = start + block.size                              # :281
    size = min(maximum + 1 + check_max - check_min,
               stop - start - pattern.size)                # :284-285
    if size <= 0:
        return []
    block = block[:size + pattern.size]
    # :295-309, all pattern bytes at once
    windows = np.lib.stride_tricks.sliding_window_view(
        block, pattern.size)[:size]
    if mask is None:
        match = (windows == pattern).all(-1)
    else:
        match = (((windows ^ pattern) & mask) == 0).all(-1)
    matches = np.nonzero(match)[0]
    if not forward:                                        # :311-313
        matches = matches[::-1]
    matches = (matches + start - offset).tolist()          # :316-317
    loc_start = max(seek_start, 0)                         # :325-326
    loc_stop = min(seek_start + maximum + 1, stop - frame_nbytes + 1)
    check_start = start                                    # :328-329
    check_stop = stop - offset - pattern.size
    found = set(matches)
    return [loc for loc in matches                         # :330-333
            if loc_start <= loc < loc_stop
            and all(c in found for c in loc + check
                    if check_start <= c < check_stop)]


def frame_table(data, locations, index_of, nslot, invalid_of=None):
    """The stream's frame table from located headers: entry (index, slot) =
    byte offset of the first frame in the file with that index and slot, -1
    where none exists or that frame is flagged invalid.  ``index_of(loc)``
    returns ``(index, slot)`` or None (cf. the one-frame-at-a-time bookkeeping
    of baseband/vdif/base.py:536-755 and baseband/base/offsets.py:6-126)."""
    entries = {}
    for loc in sorted(locations):
        key = index_of(loc)
        if key is None or key[0] < 0:
            continue
        entries.setdefault(key, loc)
    nset = max((k[0] for k in entries), default=-1) + 1
    table = np.full((nset, nslot), -1, np.int64)
    for (index, slot), loc in entries.items():
        if invalid_of is None or not invalid_of(loc):
            table[index, slot] = loc
    return table
