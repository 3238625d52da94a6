"""Item tables of Payload / Frame / FrameSet of every format on the GPU
(reference: base/tests/test_base.py:264-340, vdif/tests/test_vdif.py:431-449,
:601-692, mark4/tests/test_mark4.py:366-387)."""
import pytest

import item_cases
from test_host_items import TABLE, IDS

pytestmark = pytest.mark.gpu


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
