"""combine_with modes 1 / 2 on a few 1080p frame pairs of configuration 4: the command line profiled with ncu.

    python tools/modes_prof.py <mode> <ref> [N]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import golden_inputs as gi  # noqa: E402
import oflibnumpy_b200 as of  # noqa: E402
from oflibnumpy_b200 import _lib  # noqa: E402
from oflibnumpy_b200.device import Event, Stream  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ref = sys.argv[2] if len(sys.argv) > 2 else 't'
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H, W = 1080, 1920
of.device.require_gpu()
st = Stream()
of.device.set_stream(st)
rng = np.random.default_rng(0)
fa = of.FlowBatch.from_transforms([gi.cfg4_transforms(i) for i in range(N)], (H, W), ref,
                                  rng.random((N, H, W)) > 0.02)
fb = of.FlowBatch.from_transforms([gi.cfg4_transforms(i + 100000) for i in range(N)], (H, W), ref,
                                  rng.random((N, H, W)) > 0.02)
c0 = [_lib.call('ofk_rt_path_count', k) for k in range(6, 12)]
for _ in range(2):
    e0, e1 = Event(), Event()
    e0.record(st)
    res = fa.combine_with(fb, mode)
    e1.record(st)
    st.synchronize()
    print('mode', mode, ref, 'N', N, 'ms', e0.elapsed_ms(e1))
c1 = [_lib.call('ofk_rt_path_count', k) for k in range(6, 12)]
print('instr (coop steps, -, -, sites):', [_lib.call('ofk_rt_path_count', k) for k in range(14, 18)])
print(dict(zip(('located', 'outside_by_search', 'failed', 'hull_rejected', 'pocket_pixels', 'pocket_items'),
               [(b - a) // 2 for a, b in zip(c0, c1)])))
