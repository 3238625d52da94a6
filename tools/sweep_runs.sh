python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for R in 0 1; do BB_TUNE_RUNS=$R python - <<PY
import os, sys, torch
sys.path.insert(0, '.')
from baseband_b200 import kernels, levels
from tools.sweep_decode import timeit
DEV = 'cuda:0'
for bps, nthread, nelem in ((2, 4, 8), (2, 8, 4), (1, 4, 8), (4, 4, 4), (2, 2, 4), (2, 2, 16), (8, 2, 4), (4, 4, 8)):
    payload, frame = 8000, 8032
    nset = (1 << 30) // frame // nthread
    raw = torch.randint(0, 256, (nset * nthread * frame,), dtype=torch.uint8, device=DEV)
    uo = torch.arange(nset * nthread, dtype=torch.int64, device=DEV) * frame + 32
    spf = payload * 8 // (bps * nelem)
    out = torch.empty((nset * spf, nthread, nelem), dtype=torch.float32, device=DEV)
    lv = levels.offset_binary(bps)
    best, med = timeit(lambda: kernels.decode_bitfield(raw, uo, nset, nthread, payload, bps, nelem, False, kernels.CODEC_LEVELS, lv, out=out))
    nbytes = raw.numel() + out.numel() * 4
    print('BB_TUNE_RUNS=%s %d bit %d thr x %d ch: %7.1f GB/s best %7.1f med' % (os.environ['BB_TUNE_RUNS'], bps, nthread, nelem, nbytes / best / 1e6, nbytes / med / 1e6), flush=True)
PY
done
