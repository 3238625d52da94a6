"""baseband_b200 — B200-native (sm_100a) sample codec for radio baseband data.

Same user-facing API as ``baseband`` for the data-parallel hot path:
``baseband_b200.vdif.open(name, 'rs').read()`` etc., decoded/encoded by
hand-written CUDA kernels behind a C ABI (include/baseband_b200.h).  There is
no CPU fallback.
"""
import importlib

__version__ = '0.1'
FORMATS = ('vdif', 'mark5b', 'mark4', 'guppi', 'dada', 'gsb')
__all__ = list(FORMATS) + ['open', 'set_default_device']


def __getattr__(name):
    if name in FORMATS or name in ('kernels', 'codecs', 'levels', 'synthetic',
                                   'device', 'parallel', 'timeutil'):
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)


def set_default_device(device):
    from . import device as _device
    _device.set_default_device(device)


def open(name, mode='rs', format=None, **kwargs):
    """Open a baseband file; ``format`` is one of FORMATS (or inferred from
    the file extension)."""
    if format is None:
        ext = str(name).rsplit('.', 1)[-1].lower()
        format = {'vdif': 'vdif', 'm5b': 'mark5b', 'm4': 'mark4',
                  'raw': 'guppi', 'dada': 'dada'}.get(ext)
        if format is None:
            raise ValueError('cannot infer the format of {!r}; pass '
                             'format='.format(name))
    return importlib.import_module('.' + format, __name__).open(
        name, mode, **kwargs)
