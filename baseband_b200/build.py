"""Build libbaseband_b200.so (sm_100a) in-tree with nvcc.

    python -m baseband_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to ``baseband_b200/csrc/_obj``;
the shared library lands next to this file so it travels with the source tree.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
LIB = os.path.join(HERE, 'libbaseband_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    # numpy rounds after every ufunc: never fuse a*b+c.
    '-fmad=false',
    '-Xcompiler', '-fPIC,-O2,-Wall,-Wno-unknown-pragmas',
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cuh', '.h')):
                newest = max(newest, os.path.getmtime(os.path.join(root, f)))
    return newest


def _compile(src, verbose, extra):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    extra = extra + os.environ.get('NVCC_EXTRA', '').split()
    cmd = [NVCC] + NVCC_FLAGS + extra + ['-c', os.path.join(CSRC, src),
                                         '-o', obj]
    if verbose:
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}\n{}'.format(
            src, res.stdout, res.stderr))
    if verbose and (res.stdout or res.stderr):
        print(res.stdout, res.stderr)
    return obj


def build(force=False, verbose=False, extra=()):
    os.makedirs(OBJ, exist_ok=True)
    hdr_mtime = _deps_mtime()
    todo, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, src[:-3] + '.o')
        objs.append(obj)
        stale = (force or not os.path.exists(obj)
                 or os.path.getmtime(obj) < max(
                     os.path.getmtime(os.path.join(CSRC, src)), hdr_mtime))
        if stale:
            todo.append(src)
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as pool:
            list(pool.map(lambda s: _compile(s, verbose, list(extra)), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + [
            '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        if verbose:
            print(' '.join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    lib = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv,
                extra=['-Xptxas', '-v'] if '--ptxas' in sys.argv else [])
    print(lib)
