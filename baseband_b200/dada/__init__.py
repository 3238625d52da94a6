"""DADA format reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from ..base.opener import make_info as _make_info

info = _make_info('dada')
from .header import DADAHeader  # noqa: F401
from .payload import DADAPayload, MKBFPayload  # noqa: F401
from .frame import DADAFrame  # noqa: F401
