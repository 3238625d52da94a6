"""baseband_b200 — B200-native (sm_100a) sample codec for radio baseband data.

Same user-facing API as ``baseband`` for the data-parallel hot path:
``baseband_b200.vdif.open(name, 'rs').read()`` etc., decoded/encoded by
hand-written CUDA kernels behind a C ABI (include/baseband_b200.h).  There is
no CPU fallback.
"""
import importlib

__version__ = '0.1'
FORMATS = ('vdif', 'mark5b', 'mark4', 'guppi', 'dada', 'gsb')
__all__ = list(FORMATS) + ['open', 'guess_format', 'set_default_device']


def __getattr__(name):
    if name in FORMATS or name in ('kernels', 'codecs', 'levels', 'synthetic',
                                   'device', 'parallel', 'timeutil', 'helpers',
                                   'base'):
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)


def set_default_device(device):
    from . import device as _device
    _device.set_default_device(device)


def guess_format(name):
    """Format of a file from its first bytes (a light-weight stand-in for
    ``baseband.file_info``, baseband/io/__init__.py:107-177): 'dada', 'guppi',
    'mark5b', 'vdif', 'mark4' or None.  GSB data have no signature."""
    import io
    import numpy as np
    with io.open(name, 'rb') as fh:
        head = fh.read(1 << 20)
    first = head[:4096]
    if b'HDR_SIZE' in first and b'NBIT' in first and b'NPOL' in first:
        return 'dada'
    if len(head) >= 80 and head[8:10] == b'= ' and b'BLOCSIZE' in head[:8000]:
        return 'guppi'
    words = np.frombuffer(head[:len(head) // 4 * 4], '<u4')
    if words.size >= 4 and words[0] == 0xABADDEED:
        return 'mark5b'
    if words.size >= 8:
        # VDIF: a second header one frame further with the same shape
        nbytes = int(words[2] & 0xffffff) * 8
        version = int(words[2] >> 29)
        if 32 <= nbytes <= 1 << 20 and version <= 1 \
                and len(head) >= nbytes + 16:
            nxt = np.frombuffer(head[nbytes:nbytes + 16], '<u4')
            if (nxt[2] == words[2] and (nxt[3] & 0xffff) == (words[3] & 0xffff)
                    and (nxt[1] >> 24) == (words[1] >> 24)):
                return 'vdif'
    from .mark4.base import Mark4FileReader
    with io.open(name, 'rb') as fh:
        try:
            Mark4FileReader(fh).determine_ntrack(maximum=400000)
            return 'mark4'
        except Exception:
            pass
    return None


def open(name, mode='rs', format=None, **kwargs):
    """Open a baseband file; ``format`` is one of FORMATS, or is inferred
    from the file contents (readers) or the extension."""
    if format is None and isinstance(name, str) and mode[0] == 'r':
        try:
            format = guess_format(name)
        except OSError:
            format = None
    if format is None:
        ext = str(name).rsplit('.', 1)[-1].lower()
        format = {'vdif': 'vdif', 'm5b': 'mark5b', 'm4': 'mark4',
                  'raw': 'guppi', 'dada': 'dada'}.get(ext)
        if format is None:
            raise ValueError('cannot infer the format of {!r}; pass '
                             'format='.format(name))
    return importlib.import_module('.' + format, __name__).open(
        name, mode, **kwargs)
