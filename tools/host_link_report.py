"""Which host resource bounds the end-to-end path when N GPUs run at once?

    torchrun --nproc-per-node N tools/host_link_report.py

Every probe starts on a barrier, so all ranks hit the host at the same
time; rates are per GPU, the slowest rank's.  Probes: D2H alone, H2D
alone, both at once (what the e2e pipeline does), transfer size, pinned memory
from cudaHostAlloc vs cudaHostRegister of a (huge-page advised) anonymous
mapping, buffers first touched on the GPU-local CPUs vs on a far CPU, and the
host's own memory bandwidth (STREAM-style copy with all ranks' threads).
Rank 0 also prints the box's topology."""
import ctypes
import mmap
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from baseband_b200 import _lib, device as bb_device  # noqa: E402

rank = int(os.environ.get('RANK', 0))
local = int(os.environ.get('LOCAL_RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
GIB = 1 << 30


def barrier():
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)


def slowest(x):
    if world == 1:
        return x
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def say(label, gbs):
    g = slowest(gbs)
    if rank == 0:
        print('%-66s %7.1f GB/s per GPU  (x%d = %7.1f GB/s)'
              % (label, g, world, g * world), flush=True)


def timed_copies(pairs, reps=4):
    """pairs: [(dst, src, stream)]; all queued together, ``reps`` times."""
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for dst, src, stream in pairs:
            with torch.cuda.stream(stream):
                dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return reps / dt


if rank == 0:
    for cmd in (['nvidia-smi', 'topo', '-m'], ['lscpu'],
                ['cat', '/proc/meminfo']):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True,
                                 timeout=20).stdout
            keep = out if cmd[0] == 'nvidia-smi' else '\n'.join(
                ln for ln in out.splitlines() if any(k in ln for k in (
                    'Model name', 'Socket', 'NUMA', 'CPU(s):', 'Thread',
                    'MemTotal', 'Hugepagesize', 'HugePages_Total',
                    'AnonHugePages')))
            print('$ ' + ' '.join(cmd) + '\n' + keep, flush=True)
        except Exception as exc:
            print(cmd, 'failed:', exc)
    print('ranks: %d, affinity of rank 0 before binding: %d CPUs'
          % (world, len(os.sched_getaffinity(0))), flush=True)

cpus = bb_device.bind_host_to_device(local)
s_out, s_in = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
d_big = torch.empty(GIB, dtype=torch.uint8, device=dev)
d_big2 = torch.empty(GIB, dtype=torch.uint8, device=dev)
h_out = torch.empty(GIB, dtype=torch.uint8, pin_memory=True)
h_in = torch.empty(GIB, dtype=torch.uint8, pin_memory=True)
h_out.fill_(1)
h_in.fill_(2)
if rank == 0:
    print('\nrank 0 bound to %s CPUs local to its GPU; 1 GiB transfers unless '
          'stated' % (len(cpus) if cpus else 'no'), flush=True)
say('D2H alone (cudaHostAlloc pinned)',
    timed_copies([(h_out, d_big, s_out)]))
say('H2D alone (cudaHostAlloc pinned)',
    timed_copies([(d_big2, h_in, s_in)]))
both = timed_copies([(h_out, d_big, s_out), (d_big2, h_in, s_in)])
say('D2H and H2D at once: each direction', both)
# e2e mix: 16 bytes out per byte in
say('D2H 1 GiB + H2D 64 MiB at once (the e2e mix): D2H',
    timed_copies([(h_out, d_big, s_out),
                  (d_big2[:GIB // 16], h_in[:GIB // 16], s_in)]))
for mib in (4, 64):
    n = mib << 20
    reps = GIB // n
    pairs = [(h_out[i * n:(i + 1) * n], d_big[i * n:(i + 1) * n], s_out)
             for i in range(reps)]
    say('D2H in %d MiB pieces' % mib, timed_copies(pairs, reps=2) * 1.0)

# cudaHostRegister of an anonymous mapping, with and without huge pages
lib = _lib.load()
for huge in (False, True):
    mm = mmap.mmap(-1, GIB, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    if huge and hasattr(mmap, 'MADV_HUGEPAGE'):
        mm.madvise(mmap.MADV_HUGEPAGE)
    arr = np.frombuffer(mm, np.uint8)
    arr[::4096] = 1                                   # first touch, bound CPU
    ptr = arr.ctypes.data
    rc = lib.bb_host_register(ctypes.c_void_p(ptr), GIB)
    if rc == 0:
        t = torch.from_numpy(arr)
        say('D2H into cudaHostRegister(mmap%s)'
            % (', MADV_HUGEPAGE' if huge else ''),
            timed_copies([(t, d_big, s_out)]))
        lib.bb_host_unregister(ctypes.c_void_p(ptr))
        del t
    elif rank == 0:
        print('cudaHostRegister failed:', lib.bb_last_error())
    del arr
    mm.close()

# buffers first touched far from the GPU (only matters with > 1 NUMA node)
all_cpus = sorted(os.sched_getaffinity(0) | set(range(os.cpu_count() or 1)))
try:
    os.sched_setaffinity(0, {all_cpus[-1 - local]})
    h_far = torch.empty(GIB, dtype=torch.uint8, pin_memory=True)
    h_far.fill_(3)
    say('D2H into pinned memory first touched on a far CPU',
        timed_copies([(h_far, d_big, s_out)]))
    del h_far
except OSError as exc:
    if rank == 0:
        print('could not move to a far CPU:', exc)
if cpus:
    os.sched_setaffinity(0, set(cpus))

# host memory bandwidth: every rank copies 1 GiB with T threads at once
src_np, dst_np = h_in.numpy(), h_out.numpy()
for nthr in (1, 4):
    def work(i, nthr=nthr):
        a = GIB // nthr * i
        np.copyto(dst_np[a:a + GIB // nthr], src_np[a:a + GIB // nthr])
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        th = [threading.Thread(target=work, args=(i,)) for i in range(nthr)]
        [x.start() for x in th]
        [x.join() for x in th]
    dt = (time.perf_counter() - t0) / 3
    say('host memcpy, %d thread(s) per rank (read + write bytes)' % nthr,
        2.0 / dt)
# host memcpy while the D2H runs (do they share a bottleneck?)
barrier()
t0 = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s_out):
        h_out.copy_(d_big, non_blocking=True)
half = GIB // 2
np.copyto(h_in.numpy()[:half], h_in.numpy()[half:])
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize(dev)
dt = time.perf_counter() - t0
say('D2H while one host thread per rank copies memory', 4 / dt)
if world > 1:
    dist.destroy_process_group()
