"""A file-like object over pinned host memory.

``HostBuffer`` lets the stream readers/writers work on frames that already
sit in page-locked host memory (e.g. filled by a NIC or a capture thread):
the reader then skips its own staging copy and starts the H2D transfer
straight from the buffer, and the writer lets the D2H copy of encoded frames
land in it.  It otherwise behaves like ``io.BytesIO`` (read, readinto, write,
seek, tell), so it can be handed to any ``open`` of this package.
"""
import numpy as np
import torch

from .. import device as _device

__all__ = ['HostBuffer']


class HostBuffer:
    def __init__(self, source, name='<pinned host buffer>'):
        """``source``: number of bytes to allocate, or a uint8 torch tensor /
        numpy array / bytes to wrap (copied into pinned memory unless it
        already is a pinned tensor)."""
        if isinstance(source, int):
            self.tensor = _device.pinned_empty(source, torch.uint8)
            self.size = 0
        else:
            if isinstance(source, torch.Tensor):
                t = source.view(torch.uint8).reshape(-1)
            else:
                arr = (np.frombuffer(source, np.uint8) if isinstance(
                    source, (bytes, bytearray, memoryview))
                    else np.ascontiguousarray(source).view(np.uint8).ravel())
                t = torch.from_numpy(arr if arr.flags.writeable
                                     else arr.copy())
            if not t.is_pinned():
                pinned = _device.pinned_empty(t.numel(), torch.uint8)
                pinned.copy_(t)
                t = pinned
            self.tensor = t
            self.size = t.numel()
        self.pos = 0
        self.name = name
        self.closed = False
        self._pending = []           # events of device copies still landing

    # ------------------------------------------------------------ file API
    def seek(self, offset, whence=0):
        self.pos = (offset if whence == 0 else self.pos + offset
                    if whence == 1 else self.size + offset)
        return self.pos

    def tell(self):
        return self.pos

    # Writers let their device-to-host copies land here without blocking the
    # host (`defer`); whoever looks at the bytes waits for them first.
    def defer(self, event):
        """Register an event after which the bytes reserved so far are in
        place."""
        self._pending.append(event)
        if len(self._pending) > 64:
            self._pending.pop(0).synchronize()

    def wait(self):
        """Block until every deferred copy into this buffer has landed."""
        while self._pending:
            self._pending.pop().synchronize()

    def readinto(self, target):
        self.wait()
        view = np.frombuffer(target, np.uint8)
        n = max(0, min(view.size, self.size - self.pos))
        view[:n] = self.tensor.numpy()[self.pos:self.pos + n]
        self.pos += n
        return n

    def read(self, n=-1):
        if n is None or n < 0:
            n = self.size - self.pos
        n = max(0, min(n, self.size - self.pos))
        self.wait()
        out = self.tensor.numpy()[self.pos:self.pos + n].tobytes()
        self.pos += n
        return out

    def write(self, data):
        view = np.frombuffer(data, np.uint8)
        self.reserve(view.size).numpy()[:] = view
        return view.size

    def close(self):
        self.wait()
        self.closed = True

    def getvalue(self):
        self.wait()
        return self.tensor.numpy()[:self.size]

    # ------------------------------------------------- zero-copy extensions
    def pinned_view(self, offset, nbytes):
        """Pinned tensor over [offset, offset + nbytes) or None past EOF."""
        if offset + nbytes > self.size:
            return None
        self.wait()
        return self.tensor[offset:offset + nbytes]

    def reserve(self, nbytes):
        """Pinned tensor for the next ``nbytes`` bytes to be written (the
        position advances as if they had been written)."""
        if self.pos + nbytes > self.tensor.numel():
            raise OSError('HostBuffer of {} bytes is full'.format(
                self.tensor.numel()))
        out = self.tensor[self.pos:self.pos + nbytes]
        self.pos += nbytes
        self.size = max(self.size, self.pos)
        return out
