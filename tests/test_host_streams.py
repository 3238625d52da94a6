"""Host logic of the stream layer on the CPU emulation backend (see
tests/cpu_backend.py): frame-range planning, chunked pipeline, squeeze /
subset, header generation, file I/O.  The same cases run on the real CUDA
library in tests/test_gpu_streams.py."""
import pytest

import cpu_backend
import stream_cases


@pytest.fixture(autouse=True)
def backend(monkeypatch):
    cpu_backend.install(monkeypatch)


CASES = [n for n in dir(stream_cases)
         if n.split('_')[0] in ('vdif', 'mark5b', 'mark4', 'guppi', 'dada',
                                'gsb', 'shard', 'payload')
         and callable(getattr(stream_cases, n))]


@pytest.mark.parametrize('name', CASES)
def test_case(name):
    fn = getattr(stream_cases, name)
    if 'dev' in fn.__code__.co_varnames[:fn.__code__.co_argcount]:
        fn('cpu')
    else:
        fn()

