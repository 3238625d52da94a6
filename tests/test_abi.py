"""The C-ABI library loads and exports every symbol include/baseband_b200.h
declares (no compute calls: runs without a GPU)."""
import os
import re

import pytest

from baseband_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'baseband_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bb_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_everything():
    lib = _lib.load()          # raises ImportError listing missing symbols
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.bb_abi_version() == 2
    assert lib.bb_device_count() >= 0
    assert isinstance(lib.bb_last_error(), bytes)


def test_no_cpu_fallback():
    import numpy as np
    import pytest
    import torch
    from baseband_b200 import kernels, levels
    raw = torch.zeros(64, dtype=torch.uint8)
    off = torch.zeros(1, dtype=torch.int64)
    with pytest.raises(TypeError):
        kernels.decode_bitfield(raw, off, 1, 1, 64, 2, 1, False, 0,
                                levels.offset_binary(2))
    assert not hasattr(kernels, 'oracle')
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, 'baseband_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src


def test_plain_c_client_compiles(tmp_path):
    """The C ABI is usable from plain C: the client in tests/abi_c compiles
    against include/baseband_b200.h and resolves every symbol it needs."""
    import subprocess
    exe = tmp_path / 'abi_smoke'
    subprocess.check_call(['gcc', '-O1', '-Wall', '-Werror', '-I',
                           os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'abi_c', 'abi_smoke.c'),
                           '-ldl', '-o', str(exe)])
    res = subprocess.run([str(exe), _lib.LIB_PATH], capture_output=True,
                         text=True)
    # without a GPU it stops at the device check (4); with one it must pass
    assert res.returncode in (0, 4), res.stderr + res.stdout


@pytest.mark.gpu
def test_plain_c_client_runs(tmp_path):
    import subprocess
    exe = tmp_path / 'abi_smoke'
    subprocess.check_call(['gcc', '-O1', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'abi_c', 'abi_smoke.c'),
                           '-ldl', '-o', str(exe)])
    res = subprocess.run([str(exe), _lib.LIB_PATH], capture_output=True,
                         text=True)
    assert res.returncode == 0, res.stderr + res.stdout
    assert '0 mismatches' in res.stdout
