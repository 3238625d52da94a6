"""VDIF (VLBI Data Interchange Format) reader/writer, decoded on the GPU."""
from .base import open  # noqa: F401
from .header import VDIFHeader  # noqa: F401
from .payload import VDIFPayload  # noqa: F401
from .frame import VDIFFrame, VDIFFrameSet  # noqa: F401
