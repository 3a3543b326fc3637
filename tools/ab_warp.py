"""A/B timing of the image warp (uint8 x3 + valid area, cfg-4 transforms) and the mode-3 composition for builds of the
library selected with OFK_LIB_PATH:  python tools/ab_warp.py [N] [reps]   (run once per variant inside one job)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import golden_inputs as gi  # noqa: E402
import oflibnumpy_b200 as of  # noqa: E402
from oflibnumpy_b200 import _lib  # noqa: E402
from oflibnumpy_b200.device import DeviceArray, Event, Stream  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
H, W = 1080, 1920
of.device.require_gpu()
st = Stream()
of.device.set_stream(st)
rng = np.random.default_rng(0)
ta, tb = zip(*[(gi.cfg4_transforms(i), gi.cfg4_transforms(i + 100000)) for i in range(N)])
fa = of.FlowBatch.from_transforms(list(ta), (H, W), 't')
fb = of.FlowBatch.from_transforms(list(tb), (H, W), 't')
m1 = (rng.random((1, H, W)) > 0.02).astype(np.uint8)
img1 = rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
imgs = DeviceArray.empty((N, H, W, 3), np.uint8)
for fl in (fa, fb):
    for i in range(N):
        d = fl.masks.frames(i, i + 1)
        _lib.call('ofk_rt_memcpy_h2d', d.ptr, m1.ctypes.data, d.nbytes, st.handle)
for i in range(N):
    _lib.call('ofk_rt_memcpy_h2d', imgs.frames(i, i + 1).ptr, img1.ctypes.data, img1.nbytes, st.handle)
st.synchronize()
o_img = DeviceArray.empty((N, H, W, 3), np.uint8)
o_m = DeviceArray.empty((N, H, W), np.uint8)
o_v = DeviceArray.empty((N, H, W, 2), np.float32)
flags = DeviceArray.empty((N, 2), np.int32)
s = st.handle
c = _lib.call


def timeit(fn):
    for _ in range(3):
        fn()
    st.synchronize()
    ts = []
    for _ in range(REPS):
        e0, e1 = Event(), Event()
        e0.record(st)
        fn()
        e1.record(st)
        st.synchronize()
        ts.append(e0.elapsed_ms(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


px = N * H * W
for name, bpp, fn in (
        ("warp u8x3 + valid", 16, lambda: c('ofk_warp_t', imgs.ptr, _lib.U8, 3, _lib.ARITH_RINT, fa.vecs.ptr, -1.0, None,
                                            fa.masks.ptr, o_img.ptr, o_m.ptr, _lib.RULE_GT_HALF, N, H, W, H, W, 0, 0, 1, s)),
        ("combine3 't'", 27, lambda: c('ofk_combine3', fa.vecs.ptr, fa.masks.ptr, fb.vecs.ptr, fb.masks.ptr, ord('t'), 0.0,
                                       o_v.ptr, o_m.ptr, flags.ptr, N, H, W, s))):
    best, med = timeit(fn)
    print("%-8s %-20s best %.4f ms  median %.4f ms  %.0f GB/s (best)" % (
        os.path.basename(os.environ.get('OFK_LIB_PATH', 'default')), name, best, med, px * bpp / best / 1e6))
