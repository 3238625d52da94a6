"""World-size-2 test of the sharded read path on CPU (gloo): shard planning,
per-rank reads of disjoint frame ranges and the optional all-gather.  The
decode itself runs on the CPU emulation backend (tests/cpu_backend.py); on
the GPU box the same code runs under NCCL (bench.py --gpus N)."""
import io
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    from _pytest.monkeypatch import MonkeyPatch
    import cpu_backend
    patch = MonkeyPatch()
    cpu_backend.install(patch)
    import baseband_b200 as bb
    from baseband_b200 import parallel, synthetic
    from oracle import stream as ostream
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        raw = synthetic.vdif_stream(7, 16, 8000, seed=21, invalid=[5, 40])
        want = ostream.vdif_read(raw)[:, :, 0]
        fh = bb.vdif.open(io.BytesIO(raw.tobytes()), 'rs', sample_rate=64e6,
                          device='cpu')
        data, (a, b) = parallel.read_sharded(fh)
        assert (a, b) == parallel.shard_samples(fh, rank, world)
        ok = np.array_equal(data.numpy(), want[a:b])
        whole, span = parallel.read_sharded(fh, gather=True)
        ok = ok and span == (0, want.shape[0]) and np.array_equal(
            whole.numpy(), want)
        # GUPPI: the last rank also owns the trailing overlap
        graw, _ = synthetic.guppi_stream(5, nchan=8, npol=2,
                                         samples_per_frame=64, overlap=8)
        gwant = ostream.guppi_read(graw)
        gfh = bb.guppi.open(io.BytesIO(graw.tobytes()), 'rs', device='cpu')
        gdata, (ga, gb) = parallel.read_sharded(gfh, gather=True)
        ok = ok and np.array_equal(gdata.numpy(), gwant)
        # the packed-byte consumer: bins dealt out over the ranks, one
        # all-reduce of the count tables (7 bins; with 5-set bins a single
        # bin, so that ranks without bins take part too)
        from baseband_b200 import levels, tasks
        fh.seek(0)
        single = tasks.state_counts(fh, 32000)
        for per_bin in (32000, 5 * 32000):
            fh.seek(0)
            counts, (b0, b1) = parallel.state_counts_sharded(fh, per_bin)
            nbin = 7 * 32000 // per_bin
            ok = ok and (b0, b1) == parallel.shard_bounds(nbin, rank, world)
            ok = ok and counts.shape == (nbin, 16, 4)
            ok = ok and np.array_equal(
                counts, single[:nbin * (per_bin // 32000)].reshape(
                    nbin, -1, 16, 4).sum(1))
            ok = ok and fh.tell() == nbin * per_bin
        lv = levels.offset_binary(2)
        ok = ok and all(np.array_equal(single[..., c], (want.reshape(
            7, 32000, 16) == lv[c]).sum(1)) for c in range(4))
        results[rank] = (bool(ok), a, b)
    finally:
        dist.destroy_process_group()
        patch.undo()


def test_shard_bounds():
    from baseband_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 534731):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


@pytest.mark.timeout(300)
@pytest.mark.parametrize('world', [2, 3])
def test_world_size_n_gloo(world):
    # world 3 over 7 frame sets: gathered blocks of 3, 3 and 1 sets
    port = _free_port()
    ctx = mp.get_context('spawn')
    with ctx.Manager() as manager:
        results = manager.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, results))
                 for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(240)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode
                                                     for p in procs]
        out = dict(results)
    assert all(out[r][0] for r in range(world))
    assert all(out[r][2] == out[r + 1][1]      # contiguous shards
               for r in range(world - 1))
