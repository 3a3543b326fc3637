"""Pins oracle/remap_q32.py (the numpy restatement of cv2.remap's fixed-point bilinear sampling) against cv2.remap
itself -- the native dependency behind /root/reference/src/oflibnumpy/utils.py:236. CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import remap_q32 as Q


def _coords(kind, rng, h, w):
    if kind == 'rand':
        return (rng.uniform(-3, w + 3, (h, w)).astype(np.float32), rng.uniform(-3, h + 3, (h, w)).astype(np.float32))
    if kind == 'ties':  # exact multiples of 1/64: exercises round-half-even of the 1/32 quantisation
        return ((rng.integers(-3 * 64, (w + 3) * 64, (h, w)) / 64).astype(np.float32),
                (rng.integers(-3 * 64, (h + 3) * 64, (h, w)) / 64).astype(np.float32))
    if kind == 'int':
        return (rng.integers(-3, w + 3, (h, w)).astype(np.float32), rng.integers(-3, h + 3, (h, w)).astype(np.float32))
    yy, xx = np.mgrid[:h, :w]
    f = np.zeros((h, w, 2), np.float32)
    f[..., 0] = 0.1 * (yy - 40)
    f[..., 1] = -0.1 * (xx - 60)
    return Q.backward_map(f)


@pytest.mark.parametrize('kind', ['rand', 'ties', 'int', 'rot'])
@pytest.mark.parametrize('dtype', [np.uint8, np.int16, np.uint16, np.float32, np.float64])
def test_remap_q32_equals_cv2(kind, dtype):
    rng = np.random.default_rng(7)
    h, w = 61, 83
    mx, my = _coords(kind, rng, h, w)
    for c in (1, 2, 3, 4):
        if np.issubdtype(dtype, np.integer):
            info = np.iinfo(dtype)
            src = rng.integers(info.min, info.max + 1, (h, w, c)).astype(dtype)
        else:
            src = rng.uniform(0, 255, (h, w, c)).astype(dtype)
        if c == 1:
            src = src[..., 0]
        want = cv2.remap(src, np.dstack([mx, my]), None, cv2.INTER_LINEAR)
        got = Q.remap_linear(src, mx, my)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)  # bit-exact, float payloads included


def test_mask_rules_follow_valid_weight_sum():
    """The three warped-mask rules of the reference as functions of the valid-weight sum."""
    rng = np.random.default_rng(8)
    h, w = 50, 70
    mx, my = _coords('rand', rng, h, w)
    mask = rng.random((h, w)) > 0.2
    s = Q.valid_weight_sum(mask, mx, my)
    m = np.dstack([mx, my])
    as_f32 = cv2.remap(mask.astype(np.float32), m, None, cv2.INTER_LINEAR) == 1
    as_i16 = cv2.remap(mask.astype(np.int16), m, None, cv2.INTER_LINEAR) == 1
    as_u8 = cv2.remap(mask.astype(np.uint8), m, None, cv2.INTER_LINEAR) == 1
    np.testing.assert_array_equal(as_f32, s == 1024)
    np.testing.assert_array_equal(as_i16, s > 512)
    np.testing.assert_array_equal(as_u8, s >= 512)
