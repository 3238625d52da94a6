"""Load the arithmetic modules of the *unmodified* reference checkout.

Tooling for ``make_golden.py`` only.  It is used in the build container, where
``/root/reference`` is mounted, to produce the golden vectors committed under
``tests/golden/``; nothing in the product, the tests or the benchmark imports
it at run time (the GPU box has no reference checkout).

The reference package needs ``astropy`` which is not installed here.  The
sample arithmetic (payload codecs, header bit-fields, frame validity) does not
depend on it, so a minimal stand-in for the handful of ``astropy.utils``
helpers those modules import is registered before loading them by path.  The
stand-ins only cover import-time needs; any time/unit arithmetic raises.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('BASEBAND_REFERENCE', '/root/reference')


class _classproperty(property):
    """Stand-in for astropy.utils.classproperty (getter on the class)."""

    def __new__(cls, fget=None, doc=None, lazy=False):
        if fget is None:
            def wrapper(func):
                return cls(func, lazy=lazy)
            return wrapper
        return super().__new__(cls)

    def __init__(self, fget, doc=None, lazy=False):
        fget = self._wrap_fget(fget)
        super().__init__(fget=fget, doc=doc)

    def __get__(self, obj, objtype):
        return self.fget.__wrapped__(objtype)

    @staticmethod
    def _wrap_fget(orig_fget):
        if isinstance(orig_fget, classmethod):
            orig_fget = orig_fget.__func__

        def fget(obj):
            return orig_fget(obj.__class__)
        fget.__wrapped__ = orig_fget
        fget.__name__ = getattr(orig_fget, '__name__', 'fget')
        return fget


class _sharedmethod(classmethod):
    """Stand-in for astropy.utils.sharedmethod."""

    def __get__(self, obj, objtype=None):
        if obj is None:
            mcls = type(objtype)
            clsmeth = getattr(mcls, self.__func__.__name__, None)
            func = clsmeth if callable(clsmeth) else self.__func__
            return types.MethodType(func, objtype)
        return types.MethodType(self.__func__, obj)


class _lazyproperty(property):
    def __init__(self, fget, fset=None, fdel=None, doc=None):
        super().__init__(fget, fset, fdel, doc)
        self._key = fget.__name__

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        try:
            return obj.__dict__[self._key]
        except KeyError:
            val = self.fget(obj)
            obj.__dict__[self._key] = val
            return val

    def __set__(self, obj, val):
        obj.__dict__[self._key] = val

    def __delete__(self, obj):
        obj.__dict__.pop(self._key, None)


def _isiterable(obj):
    try:
        iter(obj)
        return True
    except TypeError:
        return False


class _NoTime:
    """Placeholder for astropy.time.Time: constructible, but inert."""

    def __init__(self, *args, **kwargs):
        self.args = args

    @classmethod
    def now(cls):
        self = cls()
        # Only used to size the VDIF ref_epoch table at import.
        self.jyear = 2040.0
        return self

    def __getattr__(self, name):
        raise RuntimeError("time arithmetic is not available in the "
                           "astropy stand-in (attribute {!r})".format(name))


class _Unit:
    def __getattr__(self, name):
        return _Unit()

    def __rmul__(self, other):
        return self

    def __mul__(self, other):
        return self

    def __truediv__(self, other):
        return self

    def __rtruediv__(self, other):
        return self

    def __pow__(self, other):
        return self


def _install_astropy_stub():
    if 'astropy' in sys.modules:
        return
    astropy = types.ModuleType('astropy')
    utils = types.ModuleType('astropy.utils')
    utils.classproperty = _classproperty
    utils.sharedmethod = _sharedmethod
    utils.lazyproperty = _lazyproperty
    utils.isiterable = _isiterable
    time = types.ModuleType('astropy.time')
    time.Time = _NoTime
    time.TimeDelta = _NoTime
    time.TimeString = object
    units = _UnitsModule('astropy.units')
    astropy.utils = utils
    astropy.time = time
    astropy.units = units
    astropy.__stub__ = True
    sys.modules.update({'astropy': astropy, 'astropy.utils': utils,
                        'astropy.time': time, 'astropy.units': units})


class _UnitsModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Unit()


_PKG = 'baseband'


def _stub_package(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def _load(name):
    """Load ``baseband.<name>`` by file path without running package inits."""
    full = _PKG + '.' + name
    if full in sys.modules:
        return sys.modules[full]
    path = os.path.join(REFERENCE_ROOT, _PKG, *name.split('.')) + '.py'
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    parent = sys.modules[full.rsplit('.', 1)[0]]
    setattr(parent, name.rsplit('.', 1)[-1], mod)
    return mod


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, _PKG))


def load_reference():
    """Return a namespace with the reference's arithmetic modules."""
    if not available():
        raise RuntimeError("reference checkout not found at "
                           + REFERENCE_ROOT)
    _install_astropy_stub()
    base = os.path.join(REFERENCE_ROOT, _PKG)
    if _PKG not in sys.modules or not hasattr(sys.modules[_PKG], '__stubbed__'):
        root = _stub_package(_PKG, base)
        root.__stubbed__ = True
        for sub in ('base', 'vdif', 'mark5b', 'mark4', 'guppi', 'dada',
                    'gsb'):
            setattr(root, sub, _stub_package(_PKG + '.' + sub,
                                             os.path.join(base, sub)))
    ns = types.SimpleNamespace()
    ns.encoding = _load('base.encoding')
    ns.utils = _load('base.utils')
    ns.base_header = _load('base.header')
    ns.base_payload = _load('base.payload')
    ns.base_frame = _load('base.frame')
    ns.mark5b_header = _load('mark5b.header')
    ns.mark5b_payload = _load('mark5b.payload')
    # `from ..mark5b import Mark5BPayload` inside vdif.payload needs these on
    # the (stub) package.
    sys.modules[_PKG + '.mark5b'].Mark5BPayload = ns.mark5b_payload.Mark5BPayload
    sys.modules[_PKG + '.mark5b'].Mark5BHeader = ns.mark5b_header.Mark5BHeader
    ns.mark5b_frame = _load('mark5b.frame')
    ns.vdif_header = _load('vdif.header')
    ns.vdif_payload = _load('vdif.payload')
    ns.vdif_frame = _load('vdif.frame')
    ns.mark4_header = _load('mark4.header')
    ns.mark4_payload = _load('mark4.payload')
    ns.mark4_frame = _load('mark4.frame')
    ns.guppi_payload = _load('guppi.payload')
    ns.dada_payload = _load('dada.payload')
    ns.gsb_payload = _load('gsb.payload')
    return ns


def sample(name):
    return os.path.join(REFERENCE_ROOT, _PKG, 'data', name)
