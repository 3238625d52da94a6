"""Treat a sequence of files as one contiguous byte stream.

Functional counterpart of baseband/helpers/sequentialfile.py:199-416 for the
two things the stream classes need: a binary reader over a list of files
(``read``, ``readinto``, ``seek``, ``tell``) and a writer that starts a new
file every ``file_size`` bytes.  File plumbing only — the bytes go through
the same pinned-buffer pipeline as a single file.
"""
import io
import os
import string

__all__ = ['FileNameSequencer', 'SequentialFileReader',
           'SequentialFileWriter', 'open']


class FileNameSequencer:
    """List-like source of file names made from a template
    (baseband/helpers/sequentialfile.py:18-83): items in curly brackets are
    filled from ``header`` (keys are case sensitive), ``{file_nr}`` with the
    index.  ``len()`` counts the files that exist for file_nr = 0, 1, ...

    >>> FileNameSequencer('a{file_nr:03d}.vdif')[10]
    'a010.vdif'
    """

    def __init__(self, template, header={}):
        self.template = template
        self.items = {}
        for _, field, _, _ in string.Formatter().parse(template):
            if field and field != 'file_nr':
                key = field.split('.')[0].split('[')[0]
                self.items[key] = header[key]

    def __getitem__(self, file_nr):
        if file_nr < 0:
            file_nr += len(self)
            if file_nr < 0:
                raise IndexError('file number out of range.')
        return self.template.format(file_nr=file_nr, **self.items)

    def __len__(self):
        file_nr = 0
        while os.path.isfile(self[file_nr]):
            file_nr += 1
        return file_nr


class SequentialFileReader:
    def __init__(self, files, mode='rb'):
        if mode != 'rb':
            raise ValueError("can only read sequences in 'rb' mode.")
        # a FileNameSequencer: the files that exist, in order
        self.files = ([files[i] for i in range(len(files))]
                      if isinstance(files, FileNameSequencer)
                      else list(files))
        if not self.files:
            raise ValueError('need at least one file.')
        self._sizes = [os.path.getsize(f) for f in self.files]
        self._starts = [0]
        for size in self._sizes:
            self._starts.append(self._starts[-1] + size)
        self._fh = None
        self._file_nr = None
        self._pos = 0
        self.closed = False
        self.name = self.files[0]

    def _locate(self, pos):
        """File number holding byte ``pos`` and the offset inside it."""
        lo, hi = 0, len(self.files) - 1
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if self._starts[mid] <= pos:
                lo = mid
            else:
                hi = mid - 1
        return lo, pos - self._starts[lo]

    def _open(self, file_nr):
        if file_nr != self._file_nr:
            if self._fh is not None:
                self._fh.close()
            self._fh = io.open(self.files[file_nr], 'rb')
            self._file_nr = file_nr

    def tell(self):
        return self._pos

    def seek(self, offset, whence=0):
        if whence == 0:
            self._pos = offset
        elif whence == 1:
            self._pos += offset
        elif whence == 2:
            self._pos = self._starts[-1] + offset
        else:
            raise ValueError("invalid 'whence'; should be 0, 1, or 2.")
        return self._pos

    def readinto(self, target):
        view = memoryview(target).cast('B')
        done = 0
        while done < len(view) and self._pos < self._starts[-1]:
            file_nr, offset = self._locate(self._pos)
            self._open(file_nr)
            self._fh.seek(offset)
            got = self._fh.readinto(view[done:])
            if not got:
                break
            done += got
            self._pos += got
        return done

    def read(self, count=-1):
        if count is None or count < 0:
            count = max(0, self._starts[-1] - self._pos)
        buf = bytearray(count)
        got = self.readinto(buf)
        return bytes(buf[:got])

    def close(self):
        if self._fh is not None:
            self._fh.close()
            self._fh = None
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class SequentialFileWriter:
    """``files``: a list of names, or a template such as ``'x_{file_nr:03d}'``
    (formatted with ``file_nr`` and any extra keywords)."""

    def __init__(self, files, mode='wb', file_size=None, **fmt):
        if mode != 'wb':
            raise ValueError("can only write sequences in 'wb' mode.")
        if file_size is None:
            raise ValueError('need file_size to split a stream over files.')
        self.files, self.file_size, self._fmt = files, int(file_size), fmt
        self._file_nr = -1
        self._fh = None
        self._room = 0
        self._pos = 0
        self.closed = False
        self.names = []

    def _next(self):
        if self._fh is not None:
            self._fh.close()
        self._file_nr += 1
        if isinstance(self.files, str):
            name = self.files.format(file_nr=self._file_nr, **self._fmt)
        else:                       # list of names or a FileNameSequencer
            name = self.files[self._file_nr]
        self.names.append(name)
        self._fh = io.open(name, 'wb')
        self._room = self.file_size

    def write(self, data):
        view = memoryview(data).cast('B')
        done = 0
        while done < len(view):
            if self._room == 0:
                self._next()
            n = min(self._room, len(view) - done)
            self._fh.write(view[done:done + n])
            done += n
            self._room -= n
        self._pos += done
        return done

    def tell(self):
        return self._pos

    def close(self):
        if self._fh is not None:
            self._fh.close()
            self._fh = None
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open(name, mode='rb', file_size=None, **kwargs):
    if 'r' in mode:
        return SequentialFileReader(name, mode)
    return SequentialFileWriter(name, mode, file_size=file_size, **kwargs)
