"""GUPPI raw format reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from ..base.opener import make_info as _make_info

info = _make_info('guppi')
from .header import GUPPIHeader  # noqa: F401
from .payload import GUPPIPayload  # noqa: F401
from .frame import GUPPIFrame  # noqa: F401
