"""GUPPI raw format reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from .header import GUPPIHeader  # noqa: F401
from .payload import GUPPIPayload  # noqa: F401
from .frame import GUPPIFrame  # noqa: F401
