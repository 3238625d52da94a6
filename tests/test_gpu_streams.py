"""Stream-level GPU parity: the public API (open / read / write / Frame /
Payload) against the oracle and the reference's golden sample outputs."""
import pytest

import stream_cases
from test_host_streams import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', CASES)
def test_case(name):
    fn = getattr(stream_cases, name)
    if 'dev' in fn.__code__.co_varnames[:fn.__code__.co_argcount]:
        fn('cuda:0')
    else:
        fn()
