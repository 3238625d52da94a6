"""Item tables of Payload / Frame / FrameSet (tests/item_cases.py) on the CPU
emulation backend: the host logic of item -> word/sample ranges.  The same
table runs against the CUDA library in tests/test_gpu_items.py."""
import pytest

import cpu_backend
import item_cases


@pytest.fixture(autouse=True)
def backend(monkeypatch):
    cpu_backend.install(monkeypatch)


TABLE = item_cases.table()
IDS = ['{}-{}'.format(k, str(i).replace(' ', '')) for k, i in TABLE]


@pytest.mark.parametrize('kind,item', TABLE, ids=IDS)
def test_getitem(kind, item):
    item_cases.check_getitem(kind, item)


@pytest.mark.parametrize('kind,item', TABLE, ids=IDS)
def test_setitem(kind, item):
    item_cases.check_setitem(kind, item)


def test_errors():
    item_cases.check_errors()


def test_frameset_header_items():
    item_cases.check_frameset_header_items()
