"""Mark 5B VLBI format reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from ..base.opener import make_info as _make_info

info = _make_info('mark5b')
from .header import Mark5BHeader  # noqa: F401
from .payload import Mark5BPayload  # noqa: F401
from .frame import Mark5BFrame  # noqa: F401
