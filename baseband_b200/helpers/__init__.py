"""Helpers around the stream classes."""
from . import sequentialfile  # noqa: F401
