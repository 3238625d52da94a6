"""CPU oracle for the baseband hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's (mhvk/baseband) sample codecs, header
bit-field extraction and frame-assembly loops.  Every function cites the
reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and only as
the checker (or the timed CPU baseline) — never as part of the shipped decode
or encode path.  ``baseband_b200`` never imports it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against vectors produced by running the unmodified reference modules
(``tests/golden/make_golden.py``, run where ``/root/reference`` is mounted) and
against the known-answer values the reference's own tests assert
(``mark5access`` values quoted in the reference test-suite).
"""
from . import codec, headers, stream  # noqa: F401
