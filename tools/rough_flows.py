"""How the TMA kernels degrade on flows whose 32x32 tiles do not fit a 48x48 source box (noise, motion boundaries),
against the gather kernels (OFK_C3_WS=0 OFK_WARP_WS=0 in the environment selects them)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oflibnumpy_b200 as of
from oflibnumpy_b200 import _lib
from oflibnumpy_b200.device import DeviceArray, Event, Stream
of.device.require_gpu()
st = Stream(); of.device.set_stream(st)
N, H, W = 16, 1080, 1920
rng = np.random.default_rng(0)
yy, xx = np.mgrid[:H, :W].astype(np.float32)


def field(kind):
    if kind.startswith('from_transforms'):
        return np.ascontiguousarray(of.from_transforms([['rotation', W / 2, H / 2, 5]], (H, W), 't'))
    if kind.startswith('smooth rotation'):
        ang = float(kind.split()[2])
        c, s = np.float32(np.cos(np.deg2rad(ang))), np.float32(np.sin(np.deg2rad(ang)))
        u = (1 - c) * (xx - W / 2) + s * (yy - H / 2); v = -s * (xx - W / 2) + (1 - c) * (yy - H / 2)
    elif kind == 'motion boundaries (50 px blocks, +-40 px)':
        bu = rng.uniform(-40, 40, (H // 50 + 1, W // 50 + 1)).astype(np.float32)
        bv = rng.uniform(-40, 40, (H // 50 + 1, W // 50 + 1)).astype(np.float32)
        u = np.kron(bu, np.ones((50, 50), np.float32))[:H, :W]; v = np.kron(bv, np.ones((50, 50), np.float32))[:H, :W]
    elif kind == 'motion boundaries (200 px blocks, +-40 px)':
        bu = rng.uniform(-40, 40, (H // 200 + 1, W // 200 + 1)).astype(np.float32)
        bv = rng.uniform(-40, 40, (H // 200 + 1, W // 200 + 1)).astype(np.float32)
        u = np.kron(bu, np.ones((200, 200), np.float32))[:H, :W]; v = np.kron(bv, np.ones((200, 200), np.float32))[:H, :W]
    else:   # white noise
        u = rng.uniform(-30, 30, (H, W)).astype(np.float32); v = rng.uniform(-30, 30, (H, W)).astype(np.float32)
    return np.ascontiguousarray(np.stack([u, v], -1), dtype=np.float32)


img = DeviceArray.from_numpy(np.tile(rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8), (N, 1, 1, 1)))
mask = DeviceArray.from_numpy(np.ones((N, H, W), np.uint8))
o_img = DeviceArray.empty((N, H, W, 3), np.uint8); o_v = DeviceArray.empty((N, H, W, 2), np.float32)
o_m = DeviceArray.empty((N, H, W), np.uint8)
print("kernels:", "gather" if os.environ.get('OFK_C3_WS') == '0' else "TMA")
for kind in ('smooth rotation 5 deg', 'smooth rotation 5 deg', 'from_transforms rotation 5', 'motion boundaries (200 px blocks, +-40 px)', 'motion boundaries (50 px blocks, +-40 px)',
             'white noise +-30 px'):
    f = DeviceArray.from_numpy(np.tile(field(kind)[None], (N, 1, 1, 1)))
    res = []
    for name, fn in (('warp u8x3+valid', lambda: _lib.call('ofk_warp_t', img.ptr, _lib.U8, 3, _lib.ARITH_RINT, f.ptr, -1.0, None, mask.ptr, o_img.ptr, o_m.ptr, _lib.RULE_GT_HALF, N, H, W, H, W, 0, 0, 1, st.handle)),
                     ('combine3', lambda: _lib.call('ofk_combine3', f.ptr, mask.ptr, f.ptr, mask.ptr, ord('t'), 0.0, o_v.ptr, o_m.ptr, None, N, H, W, st.handle))):
        for _ in range(20): fn()
        st.synchronize()
        which = 5 if name.startswith('warp') else 4
        c0 = _lib.call('ofk_rt_path_count', which)
        e0, e1 = Event(), Event(); e0.record(st)
        for _ in range(5): fn()
        e1.record(st); st.synchronize()
        glob = (_lib.call('ofk_rt_path_count', which) - c0) / 5 / (N * ((H + 31) // 32) * ((W + 31) // 32) * 8)
        res.append("%s %7.3f ms (%6.1f Gpx/s, %4.1f %% global taps)" % (name, e0.elapsed_ms(e1) / 5, N * H * W / (e0.elapsed_ms(e1) / 5) / 1e6, 100 * glob))
    print("%-46s %s" % (kind, "   ".join(res)))
