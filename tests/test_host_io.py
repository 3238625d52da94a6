"""The library's host-side ingest helpers (bb_host_copy / bb_host_pread:
native thread pool, no GPU involved) and the mmap-based file reads built on
them."""
import ctypes
import os

import numpy as np

from baseband_b200._lib import host_io


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def test_host_copy_and_pread(tmp_path):
    lib = host_io()
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, (9 << 20) + 123, dtype=np.uint8)
    for threads in (1, 3, 8, 64, 1000):
        dst = np.zeros_like(src)
        assert lib.bb_host_copy(_ptr(dst), _ptr(src), src.size, threads) == 0
        assert np.array_equal(dst, src)
    assert lib.bb_host_copy(None, None, 0, 4) == 0
    assert lib.bb_host_copy(None, _ptr(src), 10, 4) != 0
    path = tmp_path / 'blob.bin'
    src.tofile(path)
    fd = os.open(path, os.O_RDONLY)
    try:
        nread = ctypes.c_int64(-1)
        dst = np.zeros(src.size - 1000, np.uint8)
        assert lib.bb_host_pread(fd, _ptr(dst), dst.size, 1000, 5,
                                 ctypes.byref(nread)) == 0
        assert nread.value == dst.size and np.array_equal(dst, src[1000:])
        # short read at the end of the file: the bytes before the first
        # short slice are reported
        dst = np.zeros(4 << 20, np.uint8)
        assert lib.bb_host_pread(fd, _ptr(dst), dst.size,
                                 src.size - (3 << 20), 4,
                                 ctypes.byref(nread)) == 0
        assert nread.value == 3 << 20
        assert np.array_equal(dst[:3 << 20], src[-(3 << 20):])
    finally:
        os.close(fd)


def test_host_pool_survives_fork():
    lib = host_io()
    a = np.arange(1 << 23, dtype=np.uint8)
    b = np.empty_like(a)
    assert lib.bb_host_copy(_ptr(b), _ptr(a), a.size, 4) == 0
    pid = os.fork()
    if pid == 0:                       # the child has none of the threads
        b[:] = 0
        rc = lib.bb_host_copy(_ptr(b), _ptr(a), a.size, 4)
        os._exit(0 if rc == 0 and np.array_equal(a, b) else 1)
    _, status = os.waitpid(pid, 0)
    assert os.waitstatus_to_exitcode(status) == 0


def test_read_file_into_mmap_and_fallbacks(tmp_path, monkeypatch):
    """read_file_into: mmap copy for plain files (any offset, file growing
    after it was mapped), pread when mmap is switched off, readinto for
    objects that are not files."""
    import io
    from baseband_b200.base import stream
    monkeypatch.setattr(stream, 'PARALLEL_READ_MIN_NBYTES', 1)
    rng = np.random.default_rng(1)
    blob = rng.integers(0, 256, 3_000_001, dtype=np.uint8)
    path = tmp_path / 'data.bin'
    blob.tofile(path)
    with open(path, 'rb') as fh:
        for use_mmap in (True, False):
            monkeypatch.setattr(stream, 'PARALLEL_READ_MMAP', use_mmap)
            for offset, n in ((0, 4096), (17, 1_000_003), (2_999_000, 1001)):
                view = np.zeros(n, np.uint8)
                assert stream.read_file_into(fh, offset, view) == n
                assert np.array_equal(view, blob[offset:offset + n])
                assert fh.tell() == offset + n
            view = np.zeros(5000, np.uint8)          # runs past the end
            got = stream.read_file_into(fh, blob.size - 1000, view)
            assert got == 1000
            assert np.array_equal(view[:1000], blob[-1000:])
        monkeypatch.setattr(stream, 'PARALLEL_READ_MMAP', True)
        more = rng.integers(0, 256, 50_000, dtype=np.uint8)
        with open(path, 'ab') as fw:
            fw.write(more.tobytes())
        view = np.zeros(50_000, np.uint8)            # beyond the old mapping
        assert stream.read_file_into(fh, blob.size, view) == 50_000
        assert np.array_equal(view, more)
    view = np.zeros(100, np.uint8)
    assert stream.read_file_into(io.BytesIO(blob.tobytes()), 5, view) == 100
    assert np.array_equal(view, blob[5:105])
