"""Array-like conveniences shared by payloads, frames and frame sets: given
``__len__``, ``sample_shape`` and ``dtype`` a class gets ``shape``, ``size``,
``ndim`` and numpy array conversion (decoding on the GPU via ``.data``)."""
import numpy as np

__all__ = ['ArrayLike']


class ArrayLike:
    @property
    def shape(self):
        return (len(self),) + tuple(self.sample_shape)

    @property
    def size(self):
        n = 1
        for dim in self.shape:
            n *= dim
        return n

    @property
    def ndim(self):
        return 1 + len(self.sample_shape)

    def __array__(self, dtype=None, copy=None):
        data = self.data
        if dtype is None or np.dtype(dtype) == data.dtype:
            return data
        return data.astype(dtype)
