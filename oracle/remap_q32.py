"""ORACLE (test infrastructure, never imported by the product package).

Pure-numpy restatement of ``cv2.remap(src, map_f32xy, None, cv2.INTER_LINEAR)`` with the default
``BORDER_CONSTANT`` / value 0, which is the native call behind the reference's target-referenced warp
(/root/reference/src/oflibnumpy/utils.py:231-236).

The algorithm is not in /root/reference: it lives in the third-party dependency ``opencv-python``
(un-pinned in /root/reference/setup.py:52-56, docs pin 4.2.0.34; 4.13.0 in this image). What is restated
here is OpenCV's long-standing fixed-point sampling scheme (``INTER_BITS = 5``,
``INTER_REMAP_COEF_BITS = 15``; modules/imgproc/src/imgwarp.cpp, functions ``remap`` / ``remapBilinear`` /
``initInterTab2D``):

* a float32 map coordinate X is quantised to 1/32 px: ``sx = rint(X * 32)`` (round-half-even, ``cvRound``),
  integer part ``ix = sx >> 5`` (saturated to int16), fraction ``a = sx & 31``;
* the four taps (iy,ix) (iy,ix+1) (iy+1,ix) (iy+1,ix+1) read 0 outside the source;
* float32 / int16 / uint16 / float64 payloads use float32 weights ``(32-a)(32-b)/1024`` etc. and accumulate
  ``t00*w00 + t01*w01 + t10*w10 + t11*w11`` left to right (float32 accumulator for float32/int16/uint16
  payloads, float64 accumulator for float64 payloads); integer payloads finish with round-half-even and
  saturation;
* uint8 payloads use 15-bit integer weights ``w*32768`` and finish with ``(sum + 2**14) >> 15``.

Parity pin: ``tests/test_oracle_remap.py`` asserts this module equal to ``cv2.remap`` itself (bit-exact for
integer payloads; <=1 ulp-level for float payloads, where OpenCV's SIMD build may contract to FMA) on random
and adversarial coordinates, so the pin is "outputs of the reference's own native dependency run here".
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS            # 32
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS                 # 32768


def quantise_coords(map_x, map_y):
    """float32 absolute sampling coordinates -> (ix, iy, a, b) as in cv::remap's CV_32FC2 branch."""
    mx = np.asarray(map_x, dtype=np.float32)
    my = np.asarray(map_y, dtype=np.float32)
    # product in float32 (exact: power-of-two scale), rint = round-half-even, like cvRound
    sx = np.rint(mx * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    sy = np.rint(my * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    # cvRound saturates at the int32 range through the SSE conversion; irrelevant for finite image-scale inputs
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    a = sx & (INTER_TAB_SIZE - 1)
    b = sy & (INTER_TAB_SIZE - 1)
    return ix, iy, a, b


def _taps(src, ix, iy):
    """Gather the 4 taps with zero fill outside. src (H,W,C). Returns 4 arrays (...,C) and 4 in-bounds masks."""
    h, w = src.shape[:2]
    out_v, out_in = [], []
    for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
        yy = iy + dy
        xx = ix + dx
        inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        v = np.where(inside[..., None], v, np.zeros((), src.dtype))
        out_v.append(v)
        out_in.append(inside)
    return out_v, out_in


def float_weights(a, b):
    """The 4 float32 bilinear weights, exact multiples of 1/1024."""
    fa = a.astype(np.float32) / np.float32(32)
    fb = b.astype(np.float32) / np.float32(32)
    one = np.float32(1)
    return ((one - fb) * (one - fa), (one - fb) * fa, fb * (one - fa), fb * fa)


def int_weights(a, b):
    """The 4 integer (15-bit) weights of the uint8 path; they sum to 32768."""
    return ((32 - b) * (32 - a) * 32, (32 - b) * a * 32, b * (32 - a) * 32, b * a * 32)


def remap_linear(src, map_x, map_y):
    """cv2.remap(src, dstack(map_x, map_y), None, INTER_LINEAR) with BORDER_CONSTANT 0.

    :param src: (H,W) or (H,W,C) array of dtype uint8, int16, uint16, float32 or float64
    :param map_x, map_y: (Ho,Wo) absolute sampling coordinates (horizontal, vertical)
    """
    src = np.asarray(src)
    squeeze = src.ndim == 2
    s = src[..., None] if squeeze else src
    ix, iy, a, b = quantise_coords(map_x, map_y)
    taps, _ = _taps(s, ix, iy)
    if s.dtype == np.uint8:
        w = int_weights(a, b)
        acc = np.zeros(ix.shape + (s.shape[2],), np.int64)
        for t, wi in zip(taps, w):
            acc += t.astype(np.int64) * wi[..., None]
        out = ((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS)
        out = np.clip(out, 0, 255).astype(np.uint8)
    elif s.dtype in (np.int16, np.uint16, np.float32):
        w = float_weights(a, b)
        acc = taps[0].astype(np.float32) * w[0][..., None]
        for t, wi in zip(taps[1:], w[1:]):
            acc = acc + t.astype(np.float32) * wi[..., None]
        if s.dtype == np.float32:
            out = acc
        else:
            info = np.iinfo(s.dtype)
            out = np.clip(np.rint(acc), info.min, info.max).astype(s.dtype)
    elif s.dtype == np.float64:
        w = float_weights(a, b)
        acc = taps[0] * w[0][..., None].astype(np.float64)
        for t, wi in zip(taps[1:], w[1:]):
            acc = acc + t * wi[..., None].astype(np.float64)
        out = acc
    else:
        raise TypeError("remap_linear: unsupported payload dtype {}".format(s.dtype))
    return out[..., 0] if squeeze else out


def valid_weight_sum(mask, map_x, map_y):
    """Sum (in 1/1024 units, integer) of the bilinear weights falling on in-bounds taps where ``mask`` is True.

    This is the quantity all three of the reference's warped-mask rules are functions of:
    float payload ``== 1``  <=> sum == 1024 (STRICT); int16 payload <=> sum > 512; uint8 payload <=> sum >= 512.
    """
    mask = np.asarray(mask).astype(bool)
    ix, iy, a, b = quantise_coords(map_x, map_y)
    taps, _ = _taps(mask[..., None], ix, iy)
    w = ((32 - b) * (32 - a), (32 - b) * a, b * (32 - a), b * a)
    tot = np.zeros(ix.shape, np.int64)
    for t, wi in zip(taps, w):
        tot += t[..., 0].astype(np.int64) * wi
    return tot


def backward_map(flow, sign=-1.0):
    """Absolute float32 sampling coordinates for a flow: X = f32(col + sign*u), Y = f32(row + sign*v).

    sign = -1 restates /root/reference/src/oflibnumpy/utils.py:231-235 (``field *= -1; field += grid``).
    """
    f = np.asarray(flow, dtype=np.float32)
    h, w = f.shape[:2]
    s = np.float32(sign)
    mx = (s * f[..., 0] + np.arange(w, dtype=np.float32)[None, :]).astype(np.float32)
    my = (s * f[..., 1] + np.arange(h, dtype=np.float32)[:, None]).astype(np.float32)
    return mx, my
