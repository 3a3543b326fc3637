"""SURVEY section 8d, "CPU baseline beside it": the single-frame configurations 1-3 (and one 1080p forward warp) timed
through the drop-in API on one B200 and through the CPU oracle port (cv2.remap / scipy griddata, the reference's own
native calls) on the same host, same inputs, results compared. Wall clock, numpy in / numpy out.

    python tests/perf/config_table.py [--skip-1080p-griddata]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import golden_inputs as gi  # noqa: E402
import oflibnumpy_b200 as of  # noqa: E402
from oracle import flowref as R  # noqa: E402

of.device.require_gpu()


def gpu_time(fn, reps=10):
    fn()
    fn()
    of.device.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        of.device.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, r


def cpu_time(fn, reps=1):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, r


def row(name, g_ms, c_ms, agree):
    print("%-66s %9.3f ms %11.1f ms %9.0fx   %s" % (name, g_ms, c_ms, c_ms / g_ms, agree))


def vec_agree(a, b, tol=1e-3):
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    if d.max() <= tol:
        return "max |diff| %.1e" % float(d.max())
    # forward path: cells with exactly co-circular corners (similarity transforms) have no unique Delaunay diagonal
    # (DESIGN.md 2 / 4.3); only a payload that is not locally linear shows it
    bad = d.reshape(d.shape[0], d.shape[1], -1).max(-1) > tol if d.ndim >= 2 else d > tol
    return "%.2f %% of pixels beyond %g (documented: co-circular cells of a similarity transform, either diagonal is a Delaunay triangulation)" % (100 * bad.mean(), tol)


print("%-66s %12s %14s %10s   %s" % ("operation (numpy in / numpy out)", "B200", "CPU oracle", "ratio", "agreement"))

# ---- cfg 1: 375 x 1242, rotation 10 deg, ref 't'
flow1, img1 = gi.cfg1_full()
f1, r1 = of.Flow(flow1, 't'), R.make(flow1, 't')
g, w = gpu_time(lambda: of.apply_flow(flow1, img1, 't'))
c, wr = cpu_time(lambda: R.warp_array(flow1, img1, 't'), 5)
row("cfg1 apply_flow(flow, img u8x3, 't')", g, c, "identical" if np.array_equal(w, wr) else "DIFFERENT")
imgf = img1.astype(np.float32)
g, w = gpu_time(lambda: of.apply_flow(flow1, imgf, 't'))
c, wr = cpu_time(lambda: R.warp_array(flow1, imgf, 't'), 5)
row("cfg1 apply_flow(flow, img f32x3, 't')", g, c, vec_agree(w, wr))
g, w = gpu_time(lambda: f1.apply(img1, return_valid_area=True))
c, wr = cpu_time(lambda: R.apply(r1, img1, return_valid_area=True), 5)
row("cfg1 Flow.apply(img u8x3, return_valid_area=True)", g, c,
    "identical" if np.array_equal(w[0], wr[0]) and np.array_equal(w[1], wr[1]) else "DIFFERENT")

# ---- cfg 2: 436 x 1024, rotation -8 deg (+) scaling 0.9, masks 5 % invalid, mode 3
a, am, b, bm = gi.cfg2_full()
fa, fb, ra, rb = of.Flow(a, 't', am), of.Flow(b, 't', bm), R.make(a, 't', am), R.make(b, 't', bm)
g, w = gpu_time(lambda: fa.combine_with(fb, 3).vecs)
c, wr = cpu_time(lambda: R.combine(ra, rb, 3), 5)
row("cfg2 Flow.combine_with(mode=3) 't' (result read back)", g, c,
    "identical" if np.array_equal(w, wr.vecs) else "DIFFERENT")
g, w = gpu_time(lambda: of.combine_flows(a, b, 3, 't'))
c, wr = cpu_time(lambda: R.combine(R.make(a, 't'), R.make(b, 't'), 3).vecs, 5)
row("cfg2 combine_flows(a, b, 3, 't')", g, c, "identical" if np.array_equal(w, wr) else "DIFFERENT")

# ---- cfg 3: 436 x 1024, 's' flows (griddata in the reference)
h, wd = 436, 1024
s_rot = R.from_transforms([['rotation', 512, 218, -8]], (h, wd), 's')
s_smooth = gi.smooth_field(h, wd)
img3 = np.random.default_rng(6).random((h, wd, 3)).astype(np.float32) * 255          # white noise: the harshest payload
yy, xx = np.mgrid[:h, :wd].astype(np.float32)
img3s = np.stack([xx * 0.2 + yy * 0.1, np.sin(xx / 30) * 100 + 100, yy * 0.5], -1).astype(np.float32)   # smooth payload
for name, v in (("rotation -8 deg", s_rot), ("smooth non-affine field", s_smooth)):
    fs, rs = of.Flow(v, 's'), R.make(v, 's')
    g, w = gpu_time(lambda: fs.apply(img3))
    c, wr = cpu_time(lambda: R.apply(rs, img3))
    row("cfg3 Flow.apply(noise img f32x3), 's', %s" % name, g, c, vec_agree(w, wr))
    g, w = gpu_time(lambda: fs.apply(img3s))
    c, wr = cpu_time(lambda: R.apply(rs, img3s))
    row("cfg3 Flow.apply(smooth img f32x3), 's', %s" % name, g, c, vec_agree(w, wr))
    g, w = gpu_time(lambda: fs.invert().vecs)
    c, wr = cpu_time(lambda: R.invert(rs))
    m_same = np.mean(fs.invert().mask == wr.mask)
    row("cfg3 invert() s -> s, %s" % name, g, c, vec_agree(w[wr.mask & fs.invert().mask], wr.vecs[wr.mask & fs.invert().mask]) +
        ", masks equal on %.4f %%" % (100 * m_same))
ft, rt = of.Flow(R.from_transforms([['rotation', 512, 218, -8]], (h, wd), 't'), 't'), None
rt = R.make(ft.vecs, 't')
g, w = gpu_time(lambda: ft.switch_ref().vecs)
c, wr = cpu_time(lambda: R.switch_ref(rt))
both = wr.mask & ft.switch_ref().mask
row("cfg3 switch_ref() t -> s, rotation -8 deg", g, c, vec_agree(w[both], wr.vecs[both]))

# ---- one 1080p forward warp (the reference needs about a minute)
if '--skip-1080p-griddata' not in sys.argv:
    hh, ww = 1080, 1920
    v = R.from_transforms([['rotation', 960, 540, 5], ['scaling', 960, 540, 0.95]], (hh, ww), 's')
    fs, rs = of.Flow(v, 's'), R.make(v, 's')
    g, w = gpu_time(lambda: fs.invert().vecs)
    c, wr = cpu_time(lambda: R.invert(rs))
    both = wr.mask & fs.invert().mask
    row("1080p invert() s -> s, rotation 5 deg + scaling 0.95", g, c, vec_agree(w[both], wr.vecs[both]))
