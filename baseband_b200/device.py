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


# Pageable host array -> device.  torch stages a pageable copy through its
# own pinned buffer on one thread (~11 GB/s); page-locking the array in place
# costs more than it saves (cudaHostRegister pins ~10 GB/s).  Here the array is
# copied in pieces into two reusable pinned buffers by the library's native
# thread pool while the previous piece crosses PCIe.
STAGED_UPLOAD_MIN_NBYTES = 32 << 20
STAGED_UPLOAD_PIECE_NBYTES = 32 << 20
STAGED_UPLOAD_THREADS = max(1, min(8, len(os.sched_getaffinity(0))
                                   if hasattr(os, 'sched_getaffinity')
                                   else (os.cpu_count() or 1)))
_upload_stages = None


def _threaded_copy(dst, src):
    """dst[:] = src for 1-D uint8 numpy arrays, split over the library's pool
    of native threads (bb_host_copy; a numpy copy moves ~13 GB/s on one core,
    four threads 40-50)."""
    n = src.size
    if STAGED_UPLOAD_THREADS < 2 or n < (4 << 20) \
            or not (dst.flags.c_contiguous and src.flags.c_contiguous):
        dst[:] = src
        return
    import ctypes
    from ._lib import host_io
    rc = host_io().bb_host_copy(ctypes.c_void_p(dst.ctypes.data),
                                ctypes.c_void_p(src.ctypes.data), n,
                                STAGED_UPLOAD_THREADS)
    if rc != 0:
        dst[:] = src


def staged_upload(arr, device):
    """C-contiguous numpy array -> CUDA tensor of the same shape and dtype,
    through double-buffered pinned staging filled by several threads."""
    global _upload_stages
    t_dtype = torch.from_numpy(arr[:0]).dtype
    out = torch.empty(arr.shape, dtype=t_dtype, device=device)
    src = arr.reshape(-1).view(np.uint8)
    dst = out.view(-1).view(torch.uint8)
    piece = STAGED_UPLOAD_PIECE_NBYTES
    if _upload_stages is None or _upload_stages[0][0].numel() != piece:
        _upload_stages = [[pinned_empty(piece, torch.uint8), None]
                          for _ in range(2)]
    stream = torch.cuda.current_stream(device)
    for k, lo in enumerate(range(0, src.size, piece)):
        stage = _upload_stages[k % 2]
        if stage[1] is not None:
            stage[1].synchronize()              # its last copy has left
        n = min(piece, src.size - lo)
        _threaded_copy(stage[0].numpy()[:n], src[lo:lo + n])
        dst[lo:lo + n].copy_(stage[0][:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        stage[1] = ev
    for stage in _upload_stages:
        if stage[1] is not None:
            stage[1].synchronize()
            stage[1] = None
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


class _UseStream:
    """``with`` block that makes ``stream`` the current one: what
    ``torch.cuda.stream`` does, minus its per-entry Python overhead (an
    uncached small read enters three of these; they were a quarter of its
    cost)."""
    __slots__ = ('ident', 'index', 'prev')

    def __init__(self, stream):
        self.ident = (stream.stream_id, stream.device_index,
                      stream.device_type)
        self.index = stream.device_index
        self.prev = None

    def __enter__(self):
        self.prev = torch._C._cuda_getCurrentStream(self.index)
        torch._C._cuda_setStream(stream_id=self.ident[0],
                                 device_index=self.ident[1],
                                 device_type=self.ident[2])

    def __exit__(self, *exc):
        prev = self.prev
        torch._C._cuda_setStream(stream_id=prev[0], device_index=prev[1],
                                 device_type=prev[2])
        return False


class Streams:
    """The three CUDA streams of the read pipeline: 0 = H2D copies,
    1 = kernels, 2 = D2H copies.  All CUDA stream/event plumbing of the host
    layer goes through this class."""

    def __init__(self, dev):
        self.dev = dev
        self.streams = [torch.cuda.Stream(dev) for _ in range(3)]
        self._use = [_UseStream(s) for s in self.streams]

    def use(self, i):
        return self._use[i]

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

    def keep_alive(self, tensor, i):
        """``tensor`` (allocated under another stream) is about to be used on
        stream i: the caching allocator must not hand its block out again
        before that use has finished."""
        tensor.record_stream(self.streams[i])

    def synchronize(self):
        for s in self.streams:
            s.synchronize()


def record_event(dev):
    """Event recorded on the current stream of ``dev``."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    return ev


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
