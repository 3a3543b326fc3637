"""Single-frame latency through the drop-in Flow API (numpy in / numpy out), the way a user of the reference calls it."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oflibnumpy_b200 as of
of.device.require_gpu()
rng = np.random.default_rng(0)

def t(name, fn, reps=20):
    fn(); fn()
    of.device.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    of.device.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("%-70s %8.3f ms" % (name, dt * 1e3))
    return r

for (h, w, label) in ((375, 1242, 'cfg1 375x1242'), (436, 1024, 'cfg2 436x1024'), (1080, 1920, '1080p'), (2160, 3840, '4K')):
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    tr = [['rotation', w / 2, h / 2, 10]]
    f = t(label + ': Flow.from_transforms', lambda: of.Flow.from_transforms(tr, (h, w), 't'))
    g = of.Flow.from_transforms([['scaling', 300, 200, 0.9]], (h, w), 't')
    t(label + ': apply_flow(vecs, img, "t")', lambda: of.apply_flow(f.vecs, img, 't'))
    t(label + ': Flow.apply(img, return_valid_area=True)', lambda: f.apply(img, return_valid_area=True))
    t(label + ': Flow.combine_with(g, 3) -> Flow (device resident)', lambda: f.combine_with(g, 3))
    t(label + ': combine_flows(a, b, 3, "t") numpy -> numpy', lambda: of.combine_flows(f.vecs, g.vecs, 3, 't'))
    t(label + ': valid_target()', lambda: f.valid_target())
    if h <= 1080:
        t(label + ': invert() (t -> t, forward resampling)', lambda: f.invert(), reps=5)

# cfg 5 of BASELINE.json: 4K chain, device resident (from_transforms -> invert -> combine x4 -> apply to an image)
import golden_inputs as gi
img5 = gi.cfg5_image()


def chain(cross_ref_invert):
    f = of.Flow.from_transforms(gi.CFG5_TRANSFORMS, gi.CFG5_SHAPE, 't')
    if cross_ref_invert:
        g_ = f.invert('s')
        g_ = of.Flow(g_.vecs_device, 't', g_.mask_device)
    else:
        g_ = f.invert()                     # same reference: forward resampling (griddata in the reference)
    acc = f
    for i in range(4):
        acc = acc.combine_with(g_ if i % 2 == 0 else f, 3)
    return acc.apply(img5, return_valid_area=True)


t("cfg5 4K chain, invert('s') relabelled (as tests/golden digests)", lambda: chain(True), reps=5)
t("cfg5 4K chain, same-ref invert (forward resampling)", lambda: chain(False), reps=5)
from oracle import flowref as R
t0 = time.perf_counter()
f = R.make(R.from_transforms(gi.CFG5_TRANSFORMS, gi.CFG5_SHAPE, 't'), 't')
g_ = R.invert(f, 's'); g_ = R.make(g_.vecs, 't', g_.mask)
acc = f
for i in range(4):
    acc = R.combine(acc, g_ if i % 2 == 0 else f, 3)
R.apply(acc, img5, return_valid_area=True)
print("%-70s %8.1f ms  (CPU oracle port, cv2.remap on all threads; the same-ref variant needs griddata: minutes)" %
      ("cfg5 4K chain, invert('s') relabelled, reference path on the host", (time.perf_counter() - t0) * 1e3))
