"""VDIF (VLBI Data Interchange Format) reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from ..base.opener import make_info as _make_info

info = _make_info('vdif')
from .header import VDIFHeader  # noqa: F401
from .payload import VDIFPayload  # noqa: F401
from .frame import VDIFFrame, VDIFFrameSet  # noqa: F401
