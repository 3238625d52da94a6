"""The drop-in boundary as the reference's plug-in loader sees it.

``baseband.io`` (baseband/io/__init__.py:35-94) resolves a format name
through an ``importlib.metadata.EntryPoint`` of group 'baseband.io' that
points to a module, loads it, and calls ``module.open(name, mode=...,
**kwargs)`` (:236-237) and ``module.info(name, **kwargs)`` (:163-164).  Here
the reference's own loader module is loaded by path (it imports only the
standard library at module level), given fake entry points for the
``baseband_b200`` format modules exactly as
baseband/tests/test_entry_points.py:50-92 does, and files are opened and read
through it.  /root/reference exists only in the build container, so that
part is skipped elsewhere; the GPU variant exercises the same protocol
(EntryPoint.load() -> module.open / module.info) without the loader file.
"""
import importlib.util
import os
import sys
from importlib.metadata import EntryPoint

import numpy as np
import pytest

from conftest import GOLDEN, sample_path

REF_IO = '/root/reference/baseband/io/__init__.py'
OUT = np.load(os.path.join(GOLDEN, 'sample_outputs.npz'))
FORMATS = ('vdif', 'mark5b', 'mark4', 'dada', 'guppi', 'gsb')


def _load_reference_io():
    spec = importlib.util.spec_from_file_location('ref_baseband_io', REF_IO)
    module = importlib.util.module_from_spec(spec)
    sys.modules['ref_baseband_io'] = module       # its __self__ lookup
    spec.loader.exec_module(module)
    return module


def _read_all_through(open_, tag=''):
    """Open and read every sample format through ``open_(name, mode,
    format, **kwargs)``; compare with the reference's outputs (golden)."""
    with open_(sample_path('sample.vdif'), 'rs', 'vdif' + tag) as fh:
        data = fh.read()
    assert np.array_equal(data, OUT['sample_vdif_data'][:, :, 0])
    with open_(sample_path('sample.m5b'), 'rs', 'mark5b' + tag, nchan=8,
               kday=56000, sample_rate=32e6) as fh:
        assert np.array_equal(fh.read(), OUT['sample_m5b_data'])
    with open_(sample_path('sample.m4'), 'rs', 'mark4' + tag, ntrack=64,
               decade=2010, sample_rate=32e6, fill_value=-7.) as fh:
        assert np.array_equal(fh.read(), OUT['sample_m4_data'])
    with open_(sample_path('sample.dada'), 'rs', 'dada' + tag,
               squeeze=False) as fh:
        assert np.array_equal(fh.read(), OUT['sample_dada_data'])
    with open_(sample_path('sample.vdif'), 'rb', 'vdif' + tag) as fb:
        assert fb.read_frame().shape == (20000, 1)


@pytest.mark.skipif(not os.path.exists(REF_IO),
                    reason='reference tree only exists in the build container')
def test_reference_loader_resolves_b200_formats(monkeypatch):
    import cpu_backend
    cpu_backend.install(monkeypatch)
    import baseband_b200 as bb
    bio = _load_reference_io()
    try:
        dir(bio)                                  # creates FORMATS, _entries
        for fmt in FORMATS:
            name = fmt + '_b200'
            bio._entries[name] = EntryPoint(name, 'baseband_b200.' + fmt, '')
            bio.FORMATS.append(name)
        assert 'vdif_b200' in dir(bio) and 'vdif_b200' in bio.FORMATS
        assert bio.vdif_b200 is bb.vdif          # EntryPoint.load() -> module
        assert 'vdif_b200' in bio.__dict__
        for fmt in FORMATS:
            module = getattr(bio, fmt + '_b200')
            assert module is getattr(bb, fmt)
            assert callable(module.open)
        # bio.open(name, mode, format=...) -> module.open(name, mode=mode, **kw)
        _read_all_through(
            lambda name, mode, fmt, **kw: bio.open(name, mode, format=fmt,
                                                   **kw), tag='_b200')
        # what bio.file_info(name, format) does for one format (:163-164; the
        # function itself imports baseband.base.file_info, i.e. astropy)
        info = bio.vdif_b200.info(sample_path('sample.vdif'))
        assert info and info.format == 'vdif'
        assert not bio.mark5b_b200.info(sample_path('sample.vdif'))
        # a bad entry behaves as in the reference's own test (:75-92)
        bio._entries['bad'] = EntryPoint('bad', 'really_bad', '')
        bio.FORMATS.append('bad')
        with pytest.raises(AttributeError, match='not loadable'):
            bio.bad
        assert 'bad' not in bio.FORMATS
    finally:
        sys.modules.pop('ref_baseband_io', None)


def test_entry_point_protocol_cpu(monkeypatch):
    import cpu_backend
    cpu_backend.install(monkeypatch)
    _entry_point_protocol()


@pytest.mark.gpu
def test_entry_point_protocol_gpu():
    _entry_point_protocol()


def _entry_point_protocol():
    """EntryPoint('<fmt>', 'baseband_b200.<fmt>', '').load() is a module with
    the ``open`` / ``info`` the loader calls; the package's own pyproject
    declares the same entry points."""
    modules = {fmt: EntryPoint(fmt, 'baseband_b200.' + fmt, '').load()
               for fmt in FORMATS}
    _read_all_through(lambda name, mode, fmt, **kw:
                      modules[fmt].open(name, mode=mode, **kw))
    info = modules['vdif'].info(sample_path('sample.vdif'))
    assert info and info.format == 'vdif'
    text = open(os.path.join(os.path.dirname(GOLDEN), '..',
                             'pyproject.toml')).read()
    assert 'baseband.io' in text
    for fmt in FORMATS:
        assert 'baseband_b200.' + fmt in text


REF_TASKS = '/root/reference/baseband/tasks/__init__.py'


@pytest.mark.skipif(not os.path.exists(REF_TASKS),
                    reason='reference tree only exists in the build container')
def test_reference_tasks_loader_finds_b200_consumers(tmp_path, monkeypatch):
    """The analysis plug-in point (baseband/tasks/__init__.py:25-62): the
    reference's own loader, given the entry points pyproject.toml declares
    (written to a dist-info as baseband/tests/test_entry_points.py:103-110
    does), exposes the GPU consumers, and they run."""
    import cpu_backend
    cpu_backend.install(monkeypatch)
    import baseband_b200 as bb
    from baseband_b200 import tasks as b200_tasks
    info = tmp_path / 'b200_tasks-0.1.dist-info'
    info.mkdir()
    (info / 'entry_points.txt').write_text(
        '[baseband.tasks]\n'
        'state_counts = baseband_b200.tasks:state_counts\n'
        'integrated_power = baseband_b200.tasks:integrated_power\n'
        '_ = baseband_b200.tasks:__all__\n')
    monkeypatch.syspath_prepend(str(tmp_path))
    spec = importlib.util.spec_from_file_location('ref_baseband_tasks',
                                                  REF_TASKS)
    tasks = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tasks)
    assert tasks._bad_entries == []
    assert tasks.state_counts is b200_tasks.state_counts
    assert tasks.integrated_power is b200_tasks.integrated_power
    assert tasks.state_levels is b200_tasks.state_levels     # via __all__
    with bb.vdif.open(sample_path('sample.vdif'), 'rs') as fh:
        counts = tasks.state_counts(fh, 20000)
        data = OUT['sample_vdif_data'][:, :, 0]
        lv = tasks.state_levels(fh)
        want = np.stack([np.stack([(data[b * 20000:(b + 1) * 20000] == v
                                    ).sum(0) for v in lv], -1)
                         for b in range(2)])
        assert np.array_equal(counts, want)
        fh.seek(0)
        power = tasks.integrated_power(fh, 20000)
        ref = np.stack([(data[b * 20000:(b + 1) * 20000].astype(np.float64)
                         ** 2).mean(0) for b in range(2)])
        assert np.allclose(power, ref, rtol=1e-10, atol=0)
