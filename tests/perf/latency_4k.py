"""The 4K single-frame `Flow.combine_with(g, 3)` of tests/perf/latency.py in isolation (round 1 measured 49 ms there
against 0.07 ms at 1080p): per-call times with a synchronisation after every call, which kernels ran, and the same
call through the batch API. Fails if a device-resident 4K combination takes longer than 0.5 ms."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oflibnumpy_b200 as of
from oflibnumpy_b200 import _lib
of.device.require_gpu()
h, w = 2160, 3840
f = of.Flow.from_transforms([['rotation', w / 2, h / 2, 10]], (h, w), 't')
g = of.Flow.from_transforms([['scaling', 300, 200, 0.9]], (h, w), 't')
for name, a, b in (('rotation 10 deg . scaling 0.9 about (300, 200)', f, g), ('rotation . rotation', f, f),
                   ('scaling . rotation', g, f)):
    c0 = [_lib.call('ofk_rt_path_count', k) for k in range(6)]
    times = []
    for k in range(12):
        of.device.synchronize()
        t0 = time.perf_counter()
        r = a.combine_with(b, 3)
        of.device.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    c1 = [_lib.call('ofk_rt_path_count', k) for k in range(6)]
    print('%-50s ms per call: %s' % (name, ' '.join('%.3f' % t for t in times)))
    print('   paths (combine3 TMA, combine3 gather, warp TMA, warp gather, tiles with global taps c3 / warp):',
          [y - x for x, y in zip(c0, c1)])
    steady = min(times[2:])
    assert steady < 0.5, 'device-resident 4K combine_with took %.3f ms' % steady
t0 = time.perf_counter()
for _ in range(20):
    r = f.combine_with(g, 3)
of.device.synchronize()
print('20 calls back to back, no synchronisation in between: %.3f ms per call' % ((time.perf_counter() - t0) / 20 * 1e3))
