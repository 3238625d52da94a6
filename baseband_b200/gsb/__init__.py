"""GMRT Software Backend (GSB) format reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from .header import GSBHeader  # noqa: F401
from .payload import GSBPayload  # noqa: F401
from .frame import GSBFrame  # noqa: F401
