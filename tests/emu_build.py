"""Build/load the CPU emulation of the kernels (test infrastructure)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, 'emu')
EMU_LIB = os.path.join(EMU_DIR, 'libbb_emu.so')
ROOT = os.path.dirname(HERE)


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def load():
    from baseband_b200 import _lib
    srcs = [os.path.join(EMU_DIR, f) for f in os.listdir(EMU_DIR)
            if f.endswith('.cpp')]
    deps = srcs + [os.path.join(ROOT, 'baseband_b200', 'csrc', f)
                   for f in os.listdir(os.path.join(ROOT, 'baseband_b200',
                                                    'csrc'))
                   if f.endswith(('.cuh', '.h'))]
    if not os.path.exists(EMU_LIB) or os.path.getmtime(EMU_LIB) < _newest(deps):
        subprocess.check_call(
            ['g++', '-O1', '-std=c++17', '-shared', '-fPIC',
             '-ffp-contract=off', '-Wno-unknown-pragmas', '-o', EMU_LIB]
            + srcs)
    lib = ctypes.CDLL(EMU_LIB)
    _lib.bind(lib, required=())
    return lib
