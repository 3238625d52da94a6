"""Sync-pattern search in binary files (`locate_frames`, `find_header` of the
reference's file readers, baseband/base/base.py:181-376).

This is the small-window, host-side form used when a file is opened or
inspected by hand: a few frames' worth of bytes around the file pointer.
Whole files are searched on the GPU instead (csrc/bb_index.cu,
`StreamReaderBase._build_index_on_device`), with the same matching rule.
"""
import numpy as np

__all__ = ['HeaderNotFoundError', 'pattern_bytes', 'locate_frames',
           'find_header']


class HeaderNotFoundError(LookupError):
    """No header was found (baseband/base/base.py:30-33)."""


def pattern_bytes(pattern):
    """Byte view of a pattern or mask: bytes stay bytes, integers are
    unsigned 32-bit little-endian words (baseband/base/utils.py:251-270)."""
    if isinstance(pattern, (bytes, bytearray, memoryview)):
        return np.frombuffer(pattern, np.uint8)
    arr = np.atleast_1d(np.asarray(pattern))
    if arr.dtype.kind in 'iu' and arr.dtype.itemsize != 1:
        arr = arr.astype('<u4')
    return np.ascontiguousarray(arr).view(np.uint8).reshape(-1)


def locate_frames(fh, pattern, *, mask=None, frame_nbytes=None, offset=0,
                  forward=True, maximum=None, check=1):
    """File positions near the current one at which ``pattern`` occurs
    ``offset`` bytes into a frame, nearest first; the file pointer is left
    where it was.  Arguments and result as in the reference
    (baseband/base/base.py:181-225): ``pattern`` may be a header (its
    invariant bits and frame size are used), ``mask`` selects the bits that
    count, ``maximum`` bounds the distance searched (default two frames),
    ``check`` lists the frame offsets at which the pattern must recur if that
    position lies inside the file."""
    if hasattr(pattern, 'invariant_pattern'):
        if frame_nbytes is None:
            frame_nbytes = pattern.frame_nbytes
        pattern, mask = pattern.invariant_pattern()
    pat = pattern_bytes(pattern)
    if mask is not None:
        msk = pattern_bytes(mask)
        used = np.flatnonzero(msk)
        first, last = int(used[0]), int(used[-1]) + 1
        pat, msk = pat[first:last], msk[first:last]
        offset += first
    else:
        msk = None
    if maximum is None:
        maximum = (2 * frame_nbytes if frame_nbytes else 1000000) - 1
    if check is None or frame_nbytes is None:
        checks = np.zeros(0, np.int64)
    else:
        checks = np.atleast_1d(check).astype(np.int64) * frame_nbytes
    lowest = min(int(checks.min()), 0) if checks.size else 0
    highest = max(int(checks.max()), 0) if checks.size else 0
    frame = frame_nbytes if frame_nbytes is not None else offset + pat.size

    here = fh.tell()
    origin = here if forward else here - maximum
    start = max(origin + offset + lowest, 0)
    want = max(origin + maximum + 1 + highest + frame, start) - start
    fh.seek(start)
    block = np.frombuffer(fh.read(want), np.uint8)
    fh.seek(here)
    stop = start + block.size                    # where the file let us get to
    nplace = min(maximum + 1 + highest - lowest, block.size - pat.size)
    if nplace <= 0:
        return []
    windows = np.lib.stride_tricks.sliding_window_view(
        block[:nplace + pat.size], pat.size)[:nplace]
    # cheap rejection on the first byte, then the rest for the survivors
    if msk is None:
        cand = np.flatnonzero(windows[:, 0] == pat[0])
        cand = cand[(windows[cand] == pat).all(-1)]
    else:
        cand = np.flatnonzero(((windows[:, 0] ^ pat[0]) & msk[0]) == 0)
        cand = cand[(((windows[cand] ^ pat) & msk) == 0).all(-1)]
    places = (cand + start - offset).tolist()
    found = set(places)
    if not forward:
        places.reverse()
    lo, hi = max(origin, 0), min(origin + maximum + 1, stop - frame + 1)
    can_check = (start, stop - offset - pat.size)
    return [loc for loc in places
            if lo <= loc < hi and all(
                int(loc + c) in found for c in checks
                if can_check[0] <= loc + c < can_check[1])]


def find_header(reader, *args, **kwargs):
    """The nearest header that can be read, with the file pointer left at its
    start (baseband/base/base.py:337-376).  ``reader`` needs
    ``locate_frames`` and ``read_header``."""
    for loc in reader.locate_frames(*args, **kwargs):
        reader.fh_raw.seek(loc)
        try:
            header = reader.read_header()
        except Exception:
            continue
        reader.fh_raw.seek(loc)
        return header
    raise HeaderNotFoundError('could not locate a nearby frame.')
