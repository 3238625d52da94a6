import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
SAMPLES = os.path.join(GOLDEN, 'samples')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line(
        'markers', 'gpu2: needs two CUDA devices (NCCL; gpurun --gpus 2)')


def pytest_collection_modifyitems(config, items):
    # A '-m gpu' run on a box without a device must fail loudly, not skip.
    return


@pytest.fixture(scope='session')
def codec_vectors():
    return np.load(os.path.join(GOLDEN, 'codec_vectors.npz'))


@pytest.fixture(scope='session')
def sample_outputs():
    return np.load(os.path.join(GOLDEN, 'sample_outputs.npz'))


def sample_path(name):
    return os.path.join(SAMPLES, name)


def sample_bytes(name):
    return np.fromfile(sample_path(name), np.uint8)
