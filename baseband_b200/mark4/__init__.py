"""Mark 4 VLBI format reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from ..base.opener import make_info as _make_info

info = _make_info('mark4')
from .header import Mark4Header  # noqa: F401
from .payload import Mark4Payload  # noqa: F401
from .frame import Mark4Frame  # noqa: F401
