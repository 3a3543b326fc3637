"""One configuration of ofk_forward_s, a few calls: the command line that gets profiled with ncu.

    python tools/fwd_prof.py <rot|rotmask|smooth> [N] [calls]
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
from oflibnumpy_b200.device import DeviceArray, Stream  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'rot'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H, W = 1080, 1920
of.device.require_gpu()
st = Stream()
of.device.set_stream(st)
rng = np.random.default_rng(0)
if mode == 'smooth':
    fl = of.FlowBatch(np.ascontiguousarray(np.broadcast_to(gi.smooth_field(H, W)[None], (N, H, W, 2))), 's')
else:
    fl = of.FlowBatch.from_transforms([gi.cfg4_transforms(i) for i in range(N)], (H, W), 's')
mask = None
if mode == 'rotmask':
    mask = DeviceArray.from_numpy(np.ascontiguousarray(np.broadcast_to(rng.random((1, H, W)) > 0.02, (N, H, W))).view(np.uint8))
o_v = DeviceArray.empty((N, H, W, 2), np.float32)
o_m = DeviceArray.empty((N, H, W), np.uint8)
ws_bytes = _lib.call('ofk_forward_s_workspace', N, H, W)
ws = DeviceArray.empty((max(ws_bytes, 16),), np.uint8)
pm = mask if mask is not None else fl.masks
for _ in range(calls):
    _lib.call('ofk_forward_s', fl.vecs.ptr, 2, fl.vecs.ptr, 1.0, pm.ptr, mask.ptr if mask is not None else None, o_v.ptr,
              o_m.ptr, _lib.RULE_STRICT, N, H, W, ws.ptr, ws_bytes, st.handle)
st.synchronize()
print('ok', mode, N, calls, [_lib.call('ofk_rt_path_count', k) for k in range(6, 12)], [_lib.call('ofk_rt_path_count', k) for k in range(14, 18)])
