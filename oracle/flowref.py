"""ORACLE (test infrastructure, never imported by the product package).

CPU restatement, in plain functional numpy, of the flow-field hot path of oflibnumpy v1.1.1
(/root/reference/src/oflibnumpy/{utils,flow_class}.py). A flow is the triple ``(vecs, ref, mask)`` held in the
small record :class:`F`; every function cites the reference lines it restates.

The two native calls of the reference are kept at the same call sites:

* target-referenced sampling: ``cv2.remap`` (utils.py:236) -- or, with ``REMAP_IMPL = 'q32'``, the pure-numpy
  restatement in :mod:`oracle.remap_q32` (asserted bit-equal to cv2.remap in tests/test_oracle_remap.py);
* source-referenced resampling: ``scipy.interpolate.griddata(method='linear')`` (utils.py:253,
  flow_class.py:1407) -- third-party (SciPy -> Qhull Delaunay + barycentric interpolation; docs pin 1.6.0, this
  image 1.18.1). Qhull is not restated; the oracle calls it exactly where the reference does.

Parity pin: ``oracle/gen_golden.py`` ran the UNMODIFIED reference (imported from /root/reference/src in the build
container) on seeded inputs and stored its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
asserts this module reproduces them (masks bit-exact; values exact on the remap paths, 1e-9 on griddata paths),
together with the reference's own 7x7 golden masks (tests/test_flow_class.py:852-980 in the reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""
import math
from typing import NamedTuple, Optional

import numpy as np

from . import remap_q32

try:  # the reference's own native dependencies; present in this image
    import cv2
except Exception:  # pragma: no cover
    cv2 = None
try:
    from scipy.interpolate import griddata
except Exception:  # pragma: no cover
    griddata = None

REMAP_IMPL = 'cv2' if cv2 is not None else 'q32'
THRESHOLD = 1e-3  # utils.py:22


class F(NamedTuple):
    """(vecs float32 (H,W,2), ref 's'|'t', mask bool (H,W)) -- flow_class.py:29-171."""
    vecs: np.ndarray
    ref: str
    mask: np.ndarray

    @property
    def shape(self):
        return self.vecs.shape[:2]


def make(vecs, ref='t', mask=None) -> F:
    """Constructor semantics of flow_class.py:65-81,142-161: float32 copy, bool copy, default all-valid mask."""
    v = np.asarray(vecs)
    if v.ndim != 3 or v.shape[2] != 2:
        raise ValueError("flow must be (H,W,2)")
    if not np.isfinite(v).all():
        raise ValueError("flow contains non-finite values")
    v = v.astype(np.float32)
    m = np.ones(v.shape[:2], bool) if mask is None else np.asarray(mask).astype(bool)
    return F(v, 't' if ref is None else ref, m)


# ----------------------------------------------------------------------------- reductions (utils.py:298-316,527-544)
def threshold(vecs, thr=THRESHOLD):
    out = np.array(vecs, copy=True)
    out[(vecs < thr) & (vecs > -thr)] = 0
    return out


def is_zero_array(vecs, thresholded=True) -> bool:
    v = threshold(vecs) if thresholded else vecs
    return bool(np.all(v == 0))


def is_zero(f: F, thresholded=True, masked=True) -> bool:
    """flow_class.py:1230-1245 -- an empty mask counts as zero."""
    v = f.vecs[f.mask] if masked else f.vecs
    return is_zero_array(v, thresholded)


# ----------------------------------------------------------------------------- generators (utils.py:91-158,319-423)
def matrix_from_transform(name, vals):
    m = np.identity(3)
    if name == 'translation':
        m[0, 2], m[1, 2] = vals[0], vals[1]
        return m
    to_origin = matrix_from_transform('translation', [-vals[0], -vals[1]])
    back = matrix_from_transform('translation', [vals[0], vals[1]])
    if name == 'scaling':
        m[0, 0] = m[1, 1] = vals[2]
    elif name == 'rotation':  # y axis points down, hence the sign pattern (utils.py:155-156)
        r = math.radians(vals[2])
        m[0:2, 0:2] = [[math.cos(r), math.sin(r)], [-math.sin(r), math.cos(r)]]
    else:
        raise ValueError("unknown transform " + str(name))
    return back @ m @ to_origin


def matrix_from_transforms(transforms):
    m = np.identity(3)
    for t in reversed(transforms):
        m = m @ matrix_from_transform(t[0], t[1:])
    return m


def _fma(a, b, c):
    """Vectorised float64 fused multiply-add ``round(a*b + c)`` from error-free transformations (Dekker product
    split + Knuth two-sum); exact up to rare double-rounding cases. numpy has no fma ufunc."""
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64))
    p = a * b
    k = 134217729.0  # 2**27 + 1
    t = k * a
    ah = t - (t - a)
    al = a - ah
    t = k * b
    bh = t - (t - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl      # a*b == p + e exactly
    s = p + c
    v = s - p
    f = (p - (s - v)) + (c - v)                            # p + c == s + f exactly
    return s + (e + f)


def flow_from_matrix(matrix, shape):
    """utils.py:91-111. float32 homogeneous grid, float64 ``np.matmul(M, grid[..., None])``, perspective divide,
    minus grid, cast to float32.

    The (3,3)@(3,1) product is evaluated by numpy through OpenBLAS ``dgemv`` (scipy-openblas 0.3.30, Haswell-class
    kernel in the build container), whose rounding sequence was identified against np.matmul with exact rational
    arithmetic: ``t = m1*y; t = fma(m0, x, t); r = t + m2*1``. That sequence is restated here (and in the CUDA
    generator) so results do not depend on the BLAS build of the machine running the tests.
    """
    h, w = shape
    gx = np.arange(w, dtype=np.float32)[None, :].astype(np.float64)
    gy = np.arange(h, dtype=np.float32)[:, None].astype(np.float64)
    m = np.asarray(matrix, dtype=np.float64)
    tx = _fma(m[0, 0], gx, m[0, 1] * gy) + m[0, 2]
    ty = _fma(m[1, 0], gx, m[1, 1] * gy) + m[1, 2]
    tz = _fma(m[2, 0], gx, m[2, 1] * gy) + m[2, 2]
    out = np.empty((h, w, 2), np.float32)
    out[..., 0] = (tx / tz - gx).astype(np.float32)
    out[..., 1] = (ty / tz - gy).astype(np.float32)
    return out


def from_matrix(matrix, shape, ref='t'):
    if ref == 's':
        return flow_from_matrix(matrix, shape)
    return -flow_from_matrix(np.linalg.pinv(matrix), shape)  # utils.py:343


def from_transforms(transforms, shape, ref='t'):
    return from_matrix(matrix_from_transforms(transforms), shape, ref)


# ----------------------------------------------------------------------------- the warp primitive (utils.py:199-261)
def _remap(payload, mx, my):
    if REMAP_IMPL == 'cv2':
        return cv2.remap(payload, np.dstack([mx, my]), None, cv2.INTER_LINEAR)
    return remap_q32.remap_linear(payload, mx, my)


def warp_array(vecs, target, ref, mask=None):
    """``apply_flow`` (utils.py:199-261): ref 't' -> remap at grid - flow; ref 's' -> griddata from grid + flow."""
    vecs = np.asarray(vecs, dtype=np.float32)
    if is_zero_array(vecs, thresholded=True):          # utils.py:215-216 (returns the target itself)
        return target
    h, w = vecs.shape[:2]
    if ref == 't':
        mx, my = remap_q32.backward_map(vecs, -1.0)     # utils.py:231-235
        out = _remap(target, mx, my)
    else:
        rows, cols = np.mgrid[:h, :w]
        pts = np.stack([rows.ravel() + vecs[..., 1].ravel().astype(np.float64),
                        cols.ravel() + vecs[..., 0].ravel().astype(np.float64)], axis=1)   # utils.py:239-242
        vals = target.reshape(h * w, -1) if target.ndim == 3 else target.ravel()
        if mask is not None:                             # utils.py:249-251
            keep = mask.ravel()
            pts, vals = pts[keep], vals[keep]
        out = griddata(pts, vals, (rows, cols), method='linear')                            # utils.py:253
        out = np.nan_to_num(out)
        if np.issubdtype(target.dtype, np.integer):
            out = np.round(out)
        out = out.astype(target.dtype)
    if out.shape != target.shape:
        out = out[:, :, None]
    return out


def pad(f: F, padding, mode='constant') -> F:
    """flow_class.py:508-526. padding = [top, bottom, left, right]; mask padded False."""
    t, b, l, r = padding
    v = np.pad(f.vecs, ((t, b), (l, r), (0, 0)), mode=mode)
    m = np.pad(f.mask, ((t, b), (l, r)))
    return F(v, f.ref, m)


def apply(f: F, target, target_mask=None, return_valid_area=False, consider_mask=True, padding=None, cut=True):
    """``Flow.apply`` (flow_class.py:528-695). ``target`` is an :class:`F` or an ndarray (H,W[,C])."""
    target_is_flow = isinstance(target, F)
    two_d = False
    if target_is_flow:
        payload, tmask, out_dtype = target.vecs, target.mask, None
    else:
        two_d = target.ndim == 2
        payload = target[..., None] if two_d else target
        tmask = np.ones(payload.shape[:2], 'b') if target_mask is None else target_mask   # int8 vs bool: :615,:626
        out_dtype = target.dtype
    with_mask = target_is_flow or return_valid_area
    fh, fw = f.shape
    if with_mask:
        if f.ref == 's':                                  # :634-643, AND before warping
            if tmask.shape != f.shape:
                win = (slice(padding[0], padding[0] + fh), slice(padding[2], padding[2] + fw))
                inner = tmask[win] & f.mask
                tmask = np.zeros_like(tmask)              # (the reference overwrites the caller's array in place)
                tmask[win] = inner
            else:
                tmask = tmask & f.mask
        payload = np.concatenate((payload, tmask[..., None]), axis=-1)                       # :644, dtype promotion
    if padding is None:
        warped = warp_array(f.vecs, payload, f.ref, f.mask if consider_mask else None)
    else:
        g = pad(f, padding, 'constant' if f.ref == 't' else 'edge')                           # :652-660
        warped = warp_array(g.vecs, payload, g.ref, g.mask if consider_mask else None)
        if cut:
            warped = warped[padding[0]:padding[0] + fh, padding[2]:padding[2] + fw]
    out_mask = None
    if with_mask:
        out_mask = warped[..., -1] == 1                                                       # :668
        if f.ref == 't':
            if out_mask.shape != f.mask.shape:
                win = (slice(padding[0], padding[0] + fh), slice(padding[2], padding[2] + fw))
                inner = out_mask[win] & f.mask
                out_mask = np.zeros_like(out_mask)
                out_mask[win] = inner
            else:
                out_mask = out_mask & f.mask
    if target_is_flow:
        return F(warped[:, :, :2].astype(np.float32), target.ref, out_mask)
    if return_valid_area:
        warped = warped[:, :, :-1]
    if np.issubdtype(out_dtype, np.integer):
        warped = np.round(warped)
    if two_d:
        warped = warped[:, :, 0]
    warped = warped.astype(out_dtype)
    return (warped, out_mask) if return_valid_area else warped


# ----------------------------------------------------------------------------- algebra (flow_class.py:310-375,479-489)
def add(a: F, b: F) -> F:
    return F(a.vecs + b.vecs, a.ref, a.mask & b.mask)


def sub(a: F, b: F) -> F:
    return F(a.vecs - b.vecs, a.ref, a.mask & b.mask)


def neg(a: F) -> F:
    return F((a.vecs * float(-1)).astype(np.float32), a.ref, a.mask)


def relabel(a: F, ref) -> F:
    return F(a.vecs, ref, a.mask)


# ----------------------------------------------------------------------------- switch_ref / invert (:697-753)
def switch_ref(f: F) -> F:
    if is_zero(f, thresholded=False):
        return relabel(f, 't' if f.ref == 's' else 's')
    if f.ref == 's':
        return relabel(apply(f, f), 't')
    g = relabel(f, 's')
    return apply(neg(g), g)


def invert(f: F, ref=None) -> F:
    ref = f.ref if ref is None else ref
    if f.ref == 's':
        return apply(f, neg(f)) if ref == 's' else F(-f.vecs, 't', f.mask)
    if ref == 's':
        return F(-f.vecs, 's', f.mask)
    return switch_ref(invert(f, 's'))


# ----------------------------------------------------------------------------- combine (:1247-1424)
def combine(a: F, b: F, mode: int, thresholded=False) -> F:
    """``a.combine_with(b, mode)``; returns ``b`` / ``a`` themselves on the early exits (:1338-1354)."""
    if is_zero(a, thresholded=thresholded):
        return b
    if is_zero(b, thresholded=thresholded):
        return a if mode == 3 else invert(a)
    ref = a.ref
    if mode == 3:
        if ref == 's':
            return add(a, apply(invert(a, 't'), b))                                            # :1418
        return add(b, apply(b, a))                                                             # :1422
    if mode == 2:
        if ref == 's':
            return apply(a, sub(b, a))                                                         # :1390
        h, w = a.shape                                                                          # :1398-1410
        cols = np.arange(w, dtype=np.float32)[None, :]     # in-place float32 adds in the reference
        rows = np.arange(h, dtype=np.float32)[:, None]
        src = np.stack([(-a.vecs[..., 0] + cols).ravel(), (-a.vecs[..., 1] + rows).ravel()], axis=1)
        vals = np.concatenate((a.vecs, a.mask[..., None]), axis=-1).reshape(-1, 3)
        qx = -b.vecs[..., 0] + cols
        qy = -b.vecs[..., 1] + rows
        res = griddata(src, vals, (qx, qy), method='linear', fill_value=0)
        return sub(b, make(res[..., :2], 't', res[..., 2] > .99))
    # mode 1
    if ref == 's':
        bi = invert(b, 't')                                                                    # :1369-1370
        return sub(b, apply(add(bi, apply(bi, switch_ref(a))), a))
    a_s = switch_ref(a)                                                                         # :1383-1385
    inner = add(a_s, apply(invert(a_s, 't'), invert(b, 's')))
    return switch_ref(sub(switch_ref(b), apply(inner, a_s)))


# ----------------------------------------------------------------------------- valid areas (:1113-1195)
def valid_target(f: F, consider_mask=True):
    if f.ref == 's':
        area = warp_array(f.vecs, f.mask.astype('f'), 's', f.mask if consider_mask else None)
        return area == 1
    area = warp_array(f.vecs, np.ones(f.shape), 't') == 1
    return area & f.mask


def valid_source(f: F, consider_mask=True):
    if f.ref == 's':
        area = warp_array(-f.vecs, np.ones(f.shape), 't') == 1
        return area & f.mask
    area = warp_array(-f.vecs, f.mask.astype('f'), 's', f.mask if consider_mask else None)
    return area == 1


def points_inside_area(pts, shape):
    """utils.py:283-295."""
    p = np.asarray(pts)
    if np.issubdtype(p.dtype, np.floating):
        p = np.round(p).astype('i')
    return (p[..., 0] >= 0) & (p[..., 0] <= shape[0] - 1) & (p[..., 1] >= 0) & (p[..., 1] <= shape[1] - 1)


def get_padding(f: F):
    """flow_class.py:1197-1228 -> [top, bottom, left, right]."""
    v = threshold(f.vecs)
    if f.ref == 's':
        v = v * -1
    h, w = f.shape
    rows, cols = np.mgrid[:h, :w]
    ex = -(v[..., 0] - cols.astype(np.float32))           # all in-place float32 in the reference
    ey = -(v[..., 1] - rows.astype(np.float32))
    m = f.mask
    p = [max(-np.min(ey[m]), 0), max(np.max(ey[m]) - (h - 1), 0),
         max(-np.min(ex[m]), 0), max(np.max(ex[m]) - (w - 1), 0)]
    return [int(np.ceil(x)) for x in p]


# ----------------------------------------------------------------------------- "next" rows: resize / track
def resize(f: F, scale) -> F:
    """flow_class.py:491-506 + utils.py:493-524 (cv2.resize INTER_LINEAR, per-axis vector scaling)."""
    sc = [scale, scale] if isinstance(scale, (int, float)) else list(scale)
    v = cv2.resize(f.vecs, None, fx=sc[1], fy=sc[0])
    v[..., 0] *= sc[1]
    v[..., 1] *= sc[0]
    m = cv2.resize(f.mask.astype('f'), None, fx=sc[1], fy=sc[0])
    return make(v, f.ref, np.round(m))


def bilinear_at(data, pts):
    """utils.py:161-196 (exact float64 bilinear at (row, col) points; IndexError outside)."""
    r, c = pts[:, 0], pts[:, 1]
    h, w = data.shape[:2]
    if np.any(~((0 <= r) & (r <= h - 1)) | ~((0 <= c) & (c <= w - 1))):
        raise IndexError("Some points are outside of the data area.")
    r0, c0 = np.floor(r).astype(int), np.floor(c).astype(int)
    r1, c1 = np.clip(r0 + 1, 0, h - 1), np.clip(c0 + 1, 0, w - 1)
    r0, c0 = np.clip(r0, 0, h - 1), np.clip(c0, 0, w - 1)
    wa = (r1 - r) * (c1 - c)
    wb = (r1 - r) * (c - c0)
    wc = (r - r0) * (c1 - c)
    wd = (r - r0) * (c - c0)
    return wa[:, None] * data[r0, c0] + wb[:, None] * data[r1, c0] + wc[:, None] * data[r0, c1] + \
        wd[:, None] * data[r1, c1]


def track(f: F, pts, int_out=False, s_exact_mode=False):
    """utils.py:547-622; pts (N,2) as (row, col)."""
    vecs = f.vecs
    if is_zero_array(vecs, True):
        out = pts
    else:
        h, w = f.shape
        rc = vecs[..., ::-1]
        if f.ref == 's':
            if np.issubdtype(pts.dtype, np.integer):
                d = rc[pts[:, 0], pts[:, 1]]
            elif s_exact_mode:
                rows, cols = np.mgrid[:h, :w]
                grid = np.stack([rows.ravel(), cols.ravel()], 1)
                d = griddata(grid, rc.reshape(-1, 2), (pts[:, 0], pts[:, 1]), method='linear')
            else:
                d = bilinear_at(rc, pts)
            out = pts + d
        else:
            rows, cols = np.mgrid[:h, :w]
            grid = np.stack([rows.ravel(), cols.ravel()], 1)
            flat = rc.reshape(-1, 2)
            d = griddata(grid - flat, flat, (pts[:, 0], pts[:, 1]), method='linear')
            out = pts + d
        bad = np.isnan(out).any(axis=1)
        out[bad] = 0
    if int_out:
        out = np.round(out).astype('i')
    return out


# ----------------------------------------------------------------------------- visualisation (flow_class.py:869-951)
def visualise(f: F, mode: str, show_mask: bool = False, show_mask_borders: bool = False, range_max=None):
    """Flow.visualise: hue from cv2.cartToPolar's angle, saturation from the magnitude scaled to `range_max`
    (default: 99th percentile, numpy.percentile), value 255 (180 on invalid pixels with show_mask), mask outline via
    cv2.findContours / drawContours, HSV -> RGB in numpy. Third-party calls are kept at the reference's call sites."""
    fl = threshold(f.vecs)                                                   # :895
    hsv = np.zeros((fl.shape[0], fl.shape[1], 3), 'f')
    mag, ang = cv2.cartToPolar(fl[..., 0], fl[..., 1], angleInDegrees=True)  # :900
    hsv[..., 0] = np.mod(ang, 360) / 2
    hsv[..., 2] = 255
    if show_mask:
        hsv[np.invert(f.mask), 2] = 180                                      # :906
    if range_max is None:                                                    # :909-915
        if np.percentile(mag, 99) > 0:
            range_max = float(np.percentile(mag, 99))
        elif np.max(mag):
            range_max = float(np.max(mag))
        else:
            range_max = 1
    hsv[..., 1] = np.clip(mag * 255 / range_max, 0, 255)                     # :920
    if show_mask_borders:                                                    # :923-926
        contours, _ = cv2.findContours((255 * f.mask).astype('uint8'), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        cv2.drawContours(hsv, contours, -1, (0, 0, 0), 1)
    if mode == 'hsv':
        return np.round(hsv).astype('uint8')
    h = hsv[..., 0] / 180                                                    # :931-945
    s = hsv[..., 1] / 255
    v = hsv[..., 2] / 255
    i = np.int_(h * 6.)
    fr = h * 6. - i
    i = np.ravel(i)
    t = np.ravel(1. - fr)
    fr = np.ravel(fr)
    i %= 6
    c_list = (1 - np.ravel(s) * np.vstack([np.zeros_like(fr), np.ones_like(fr), fr, t])) * np.ravel(v)
    order = np.array([[0, 3, 1], [2, 0, 1], [1, 0, 3], [1, 2, 0], [3, 1, 0], [0, 1, 2]])
    rgb = c_list[order[i], np.arange(np.prod(h.shape))[:, None]].reshape(*h.shape, 3)
    rgb = np.round(rgb * 255).astype('uint8')
    return rgb[..., ::-1] if mode == 'bgr' else rgb
