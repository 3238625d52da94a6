"""Where does an uncached small host read spend its time?  cProfile of
read(100) with the window cache off."""
import cProfile
import io
import pstats
import sys
import time

sys.path.insert(0, '.')
import baseband_b200 as bb  # noqa: E402
from baseband_b200 import synthetic  # noqa: E402
from baseband_b200.base import stream  # noqa: E402

raw = synthetic.vdif_stream(64, 8, 5000, seed=1).tobytes()
stream.StreamReaderBase.SMALL_READ_NBYTES = 0
fh = bb.vdif.open(io.BytesIO(raw), 'rs', sample_rate=32e6)
fh.read(1)


def loop(n=2000):
    for k in range(n):
        fh.seek((k * 137) % 100000)
        fh.read(100)


loop(200)
t0 = time.perf_counter()
loop()
print('%.1f us per uncached read(100)' % ((time.perf_counter() - t0) / 2000 * 1e6))
pr = cProfile.Profile()
pr.enable()
loop()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
