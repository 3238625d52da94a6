"""Device selection and host<->device staging for the host layer.

All sample arithmetic runs on a CUDA device; this module decides which one
and moves bytes.  Streams of frames are staged through pinned host buffers so
the H2D copy is asynchronous on the current torch stream.
"""
import os

import numpy as np
import torch

_default = None


def default_device():
    """Device used when the caller does not name one: ``cuda:LOCAL_RANK``
    under torchrun, else ``BASEBAND_B200_DEVICE`` or ``cuda:0``."""
    global _default
    if _default is None:
        name = os.environ.get('BASEBAND_B200_DEVICE')
        if name is None:
            name = 'cuda:{}'.format(int(os.environ.get('LOCAL_RANK', 0)))
        _default = torch.device(name)
    return _default


def set_default_device(device):
    global _default
    _default = torch.device(device)


def resolve(device=None):
    dev = default_device() if device is None else torch.device(device)
    if dev.type != 'cuda':
        raise ValueError('baseband_b200 decodes on CUDA devices only (no CPU '
                         'fallback); got {!r}'.format(str(dev)))
    if not torch.cuda.is_available():
        raise RuntimeError('baseband_b200 needs a CUDA device; none is '
                           'visible (there is no CPU fallback).')
    return dev


class PinnedStage:
    """A reusable pinned host buffer (grows on demand)."""

    def __init__(self):
        self._buf = None

    def get(self, nbytes):
        if self._buf is None or self._buf.numel() < nbytes:
            size = max(nbytes, 1 << 20)
            self._buf = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        return self._buf[:nbytes]


_stage = PinnedStage()


def upload(raw, device, stage=None, align=16):
    """Copy host bytes to a uint8 CUDA tensor through pinned memory.

    ``raw`` may be bytes, a numpy array (any dtype) or a uint8 torch tensor.
    Device allocations are at least 256-byte aligned, which covers every
    alignment the kernels need for offset 0.
    """
    if isinstance(raw, torch.Tensor):
        if raw.is_cuda:
            return raw.view(torch.uint8).reshape(-1)
        return raw.view(torch.uint8).reshape(-1).to(device, non_blocking=True)
    if isinstance(raw, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(raw, np.uint8)
    else:
        arr = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    nbytes = arr.size
    stage = stage or _stage
    pinned = stage.get(nbytes)
    pinned.numpy()[:] = arr
    out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    out.copy_(pinned, non_blocking=True)
    # the stage is reused by the next call: make the copy complete first
    torch.cuda.current_stream(device).synchronize()
    return out


def download(tensor):
    """CUDA tensor -> numpy (synchronous)."""
    return tensor.cpu().numpy()


def register_host(arr):
    """Page-lock a numpy array in place (``cudaHostRegister`` through the C
    ABI) so a D2H copy can land in it asynchronously at full PCIe rate.
    Returns a token for `unregister_host`, or None if pinning failed."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    ptr = ctypes.c_void_p(arr.ctypes.data)
    if lib.bb_host_register(ptr, arr.nbytes) != 0:
        return None
    return ptr


def unregister_host(token):
    from . import _lib
    _lib.load().bb_host_unregister(token)


def pinned_empty(shape, dtype):
    """Page-locked host tensor (D2H / H2D copies to it are asynchronous)."""
    return torch.empty(shape, dtype=dtype, pin_memory=True)


class Streams:
    """The three CUDA streams of the read pipeline: 0 = H2D copies,
    1 = kernels, 2 = D2H copies.  All CUDA stream/event plumbing of the host
    layer goes through this class."""

    def __init__(self, dev):
        self.dev = dev
        self.streams = [torch.cuda.Stream(dev) for _ in range(3)]

    def use(self, i):
        return torch.cuda.stream(self.streams[i])

    def wait(self, i, j):
        """Stream i waits for everything queued so far on stream j."""
        self.streams[i].wait_stream(self.streams[j])

    def after_caller(self, i):
        self.streams[i].wait_stream(torch.cuda.current_stream(self.dev))

    def caller_after(self, i):
        torch.cuda.current_stream(self.dev).wait_stream(self.streams[i])

    def event(self, i):
        ev = torch.cuda.Event()
        ev.record(self.streams[i])
        return ev

    def wait_event(self, i, ev):
        """Stream i waits for one recorded event only."""
        if ev is not None:
            self.streams[i].wait_event(ev)

    def synchronize(self):
        for s in self.streams:
            s.synchronize()


def current_stream_synchronize(dev):
    torch.cuda.current_stream(dev).synchronize()


def is_device_tensor(t):
    return isinstance(t, torch.Tensor) and t.is_cuda


def bind_host_to_device(index=None):
    """Pin this process to the CPU cores that are local (same NUMA node) to
    CUDA device ``index`` so that pinned staging buffers allocated afterwards
    are first-touched next to the GPU's PCIe root.  With one process per GPU
    this keeps every rank's H2D/D2H traffic on its own socket.  Returns the
    CPU list used, or None if the topology could not be determined."""
    try:
        if index is None:
            index = default_device().index or 0
        props = torch.cuda.get_device_properties(index)
        bus = '{:04x}:{:02x}:{:02x}.0'.format(props.pci_domain_id,
                                              props.pci_bus_id,
                                              props.pci_device_id)
        with open('/sys/bus/pci/devices/{}/local_cpulist'.format(bus)) as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except (OSError, AttributeError, ValueError, RuntimeError):
        return None
