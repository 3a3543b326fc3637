"""Dev micro-benchmarks of the streaming kernels (ceiling for the I/O mix of the hot kernels)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oflibnumpy_b200 as of
from oflibnumpy_b200 import _lib, _ops
from oflibnumpy_b200.device import DeviceArray, Event, Stream

N, H, W = int(os.environ.get('MB_N', 64)), 1080, 1920
st = Stream(); of.device.set_stream(st)
a = DeviceArray.zeros((N, H, W, 2), np.float32); b = DeviceArray.zeros((N, H, W, 2), np.float32)
am = DeviceArray.zeros((N, H, W), np.uint8); bm = DeviceArray.zeros((N, H, W), np.uint8)
o = DeviceArray.empty((N, H, W, 2), np.float32); om = DeviceArray.empty((N, H, W), np.uint8)
_lib.call('ofk_rt_memset', am.ptr, 1, am.nbytes, st.handle); _lib.call('ofk_rt_memset', bm.ptr, 1, bm.nbytes, st.handle)
px = N * H * W

def timeit(name, fn, bytes_px, reps=10):
    for _ in range(3): fn()
    e0, e1 = Event(), Event()
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); st.synchronize()
    ms = e0.elapsed_ms(e1) / reps
    print("%-28s %7.3f ms  %7.1f GB/s  %8.1f Mpx/s" % (name, ms, px * bytes_px / ms / 1e6, px / ms / 1e3))

timeit('addsub (27 B/px)', lambda: _lib.call('ofk_addsub', _lib.OP_ADD, a.ptr, am.ptr, b.ptr, bm.ptr, o.ptr, om.ptr, N, H, W, st.handle), 27)
timeit('d2d copy flows (16 B/px)', lambda: _lib.call('ofk_rt_memcpy_d2d', o.ptr, a.ptr, a.nbytes, st.handle), 16)
timeit('valid_geom_t (10 B/px)', lambda: _lib.call('ofk_valid_geom_t', a.ptr, -1.0, am.ptr, om.ptr, N, H, W, st.handle), 10)
flags = DeviceArray.empty((N, 2), np.int32)
timeit('combine3 zero flows (27)', lambda: _lib.call('ofk_combine3', a.ptr, am.ptr, b.ptr, bm.ptr, ord('t'), 0.0, o.ptr, om.ptr, None, N, H, W, st.handle), 27)
timeit('combine3 no masks zero (25)', lambda: _lib.call('ofk_combine3', a.ptr, None, b.ptr, None, ord('t'), 0.0, o.ptr, om.ptr, None, N, H, W, st.handle), 25)
img = DeviceArray.zeros((N, H, W, 3), np.uint8); oimg = DeviceArray.empty((N, H, W, 3), np.uint8)
timeit('warp u8x3 rint+valid (16)', lambda: _lib.call('ofk_warp_t', img.ptr, _lib.U8, 3, _lib.ARITH_RINT, a.ptr, -1.0, None, am.ptr, oimg.ptr, om.ptr, _lib.RULE_GT_HALF, N, H, W, H, W, 0, 0, 1, st.handle), 16)
timeit('warp u8x3 fixed novalid (14)', lambda: _lib.call('ofk_warp_t', img.ptr, _lib.U8, 3, _lib.ARITH_NATIVE, a.ptr, -1.0, None, None, oimg.ptr, None, _lib.RULE_STRICT, N, H, W, H, W, 0, 0, 1, st.handle), 14)
timeit('warp flow f32x2 + masks (27)', lambda: _lib.call('ofk_warp_t', b.ptr, _lib.F32, 2, _lib.ARITH_NATIVE, a.ptr, -1.0, bm.ptr, am.ptr, o.ptr, om.ptr, _lib.RULE_STRICT, N, H, W, H, W, 0, 0, 1, st.handle), 27)
timeit('from_matrix (8)', lambda: _ops.from_matrix(np.tile(np.eye(3)[None], (N, 1, 1)), (H, W), -1.0), 8, reps=3)
