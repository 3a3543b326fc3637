"""Seeded synthetic inputs shared by oracle/gen_golden.py (which feeds them to the real reference) and by the tests
(which feed the identical arrays to the oracle restatement and to the CUDA path). SURVEY.md section 8d.

Flows built from transforms use the oracle's generator, itself pinned bit-exact against the reference
(tests/golden/generators.npz, digests.json: cfg1_flow_sha / cfg5_from_transforms_*).
"""
import numpy as np



class _LazyOracle(object):
    """The oracle is imported only when an input builder actually needs it: bench.py takes its workload parameters from
    this module (cfg4_transforms, plain lists) and must not pull the oracle into its GPU leg."""

    def __getattr__(self, name):
        from oracle import flowref
        return getattr(flowref, name)


R = _LazyOracle()

RESIZE_SCALES = [0.5, 2, [1.5, 0.75]]
CFG5_SHAPE = (2160, 3840)
CFG5_TRANSFORMS = [['rotation', 1920, 1080, 5], ['scaling', 500, 300, 0.95]]


def smooth_field(h, w, scale=1.0):
    """Non-affine smooth field of SURVEY section 8d-2."""
    y, x = np.mgrid[:h, :w].astype(np.float64)
    f = np.empty((h, w, 2), np.float32)
    f[..., 0] = scale * 8 * np.sin(x / 40) * np.cos(y / 55)
    f[..., 1] = scale * 6 * np.cos(x / 35 + 1) * np.sin(y / 45)
    return f


def _images(rng, h, w):
    d = {}
    d['img_u8c3'] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    d['img_u8c1'] = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)
    d['img_u8c4'] = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    d['img_f32c3'] = (rng.random((h, w, 3)) * 255).astype(np.float32)
    d['img_f32c1'] = (rng.random((h, w, 1)) * 255).astype(np.float32)
    d['img_f64c2'] = rng.random((h, w, 2)) * 255
    d['img_i16c3'] = rng.integers(-3000, 3000, (h, w, 3)).astype(np.int16)
    d['img_u16c2'] = rng.integers(0, 65536, (h, w, 2)).astype(np.uint16)
    d['img_u8_2d'] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    d['img_f32_2d'] = (rng.random((h, w)) * 255).astype(np.float32)
    return d


def warp_t_inputs():
    rng = np.random.default_rng(101)
    h, w = 47, 155                                           # cfg 1 aspect (375x1242) / 8, odd sizes
    d = _images(rng, h, w)
    d['flow'] = R.from_transforms([['rotation', w / 2, h / 2, 10]], (h, w), 't')
    d['flow'][5:9, 7:30] *= np.float32(0.0)                # a patch of exact zeros (identity taps)
    d['flow'][20, 40:60] = np.float32(1) / 64               # exact half-quantum ties
    d['flow_mask'] = rng.random((h, w)) > 0.05
    d['target_mask'] = rng.random((h, w)) > 0.1
    d['flow2'] = smooth_field(h, w, 0.5)
    d['flow2_mask'] = rng.random((h, w)) > 0.05
    return d


def padded_inputs():
    rng = np.random.default_rng(102)
    h, w = 31, 50
    pad = [3, 5, 7, 2]                                      # top, bottom, left, right
    H, W = h + pad[0] + pad[1], w + pad[2] + pad[3]
    d = {'padding': np.array(pad)}
    d['flow'] = R.from_transforms([['rotation', 10, 12, -12], ['translation', 2.5, -1.25]], (h, w), 't')
    d['flow_mask'] = rng.random((h, w)) > 0.1
    d['img_u8c3'] = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    d['img_f32c3'] = (rng.random((H, W, 3)) * 255).astype(np.float32)
    d['target_mask'] = rng.random((H, W)) > 0.1
    d['big_flow'] = smooth_field(H, W, 0.3)
    d['big_flow_mask'] = rng.random((H, W)) > 0.1
    return d


def combine_inputs():
    h, w = 55, 129                                          # cfg 2 aspect (436x1024) / 8, odd sizes
    d = {}
    d['aff_1'] = R.from_transforms([['rotation', w / 2, h / 2, -8]], (h, w), 't')
    d['aff_2'] = R.from_transforms([['scaling', 37.5, 25, 0.9]], (h, w), 't')
    d['smooth_1'] = smooth_field(h, w, 0.6)
    d['smooth_2'] = smooth_field(h, w, -0.4)[::-1, ::-1].copy()
    d['mask_1'] = np.random.default_rng(2).random((h, w)) > 0.05
    d['mask_2'] = np.random.default_rng(3).random((h, w)) > 0.05
    z = np.zeros((h, w, 2), np.float32)
    z[3, 4] = (5.0, -2.0)                                   # hidden by the mask -> is_zero(masked) is True
    zm = np.ones((h, w), bool)
    zm[3, 4] = False
    d['zero_where_valid'] = z
    d['zero_mask'] = zm
    d['tiny'] = (np.random.default_rng(5).random((h, w, 2)).astype(np.float32) - np.float32(0.5)) * np.float32(1e-3)
    return d


def forward_inputs():
    rng = np.random.default_rng(103)
    h, w = 40, 64
    d = {}
    d['aff'] = R.from_transforms([['rotation', w / 2, h / 2, -8]], (h, w), 's')
    d['smooth'] = smooth_field(h, w, 0.5)
    d['img_f32c3'] = (rng.random((h, w, 3)) * 255).astype(np.float32)
    d['img_u8c3'] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    d['mask'] = rng.random((h, w)) > 0.05
    return d


def combine12_inputs():
    h, w = 48, 60
    t = [['rotation', 30.5, 24.5, -20], ['scaling', 12, 10, 0.85]]
    d = {}
    for r in ('s', 't'):
        d['f1_' + r] = R.from_transforms(t[0:1], (h, w), r)
        d['f2_' + r] = R.from_transforms(t[1:2], (h, w), r)
        d['f3_' + r] = R.from_transforms(t, (h, w), r)
    return d


def generator_specs():
    proj = [[1.02, 0.03, -4.0], [-0.02, 0.97, 3.0], [1e-4, -2e-4, 1.0]]
    aff = [[0.9, -0.2, 12.5], [0.15, 1.1, -7.25], [0.0, 0.0, 1.0]]
    return [
        ('transforms', [['rotation', 10, 50, -30]], (41, 67), 's'),
        ('transforms', [['rotation', 10, 50, -30]], (41, 67), 't'),
        ('transforms', [['translation', 10, 20]], (33, 70), 't'),
        ('transforms', [['scaling', 20, 10, 1.3], ['rotation', 5, 6, 17.5], ['translation', -3, 4.5]], (64, 48), 's'),
        ('transforms', [['scaling', 20, 10, 1.3], ['rotation', 5, 6, 17.5], ['translation', -3, 4.5]], (64, 48), 't'),
        ('matrix', proj, (50, 81), 's'),
        ('matrix', proj, (50, 81), 't'),
        ('matrix', aff, (1, 1), 't'),
        ('matrix', aff, (3, 257), 's'),
    ]


def next_inputs():
    rng = np.random.default_rng(104)
    h, w = 36, 52
    d = {}
    d['flow'] = R.from_transforms([['rotation', 0, 0, 12]], (h, w), 's')
    m = np.ones((h, w), bool)
    m[:, 40:] = False
    d['mask'] = m
    pts = np.stack([rng.uniform(0, h - 1, 40), rng.uniform(0, w - 1, 40)], 1)
    d['pts_f'] = pts
    d['pts_i'] = np.round(pts).astype(np.int64)
    return d


# ---------------------------------------------------------------------------------------- full-size configurations
def cfg1_full():
    h, w = 375, 1242
    flow = R.from_transforms([['rotation', 621, 187.5, 10]], (h, w), 't')
    img = np.random.default_rng(1).integers(0, 256, (h, w, 3), dtype=np.uint8)
    return flow, img


def cfg2_full():
    h, w = 436, 1024
    a = R.from_transforms([['rotation', 512, 218, -8]], (h, w), 't')
    b = R.from_transforms([['scaling', 300, 200, 0.9]], (h, w), 't')
    am = np.random.default_rng(2).random((h, w)) > 0.05
    bm = np.random.default_rng(3).random((h, w)) > 0.05
    return a, am, b, bm


def cfg3_full():
    """cfg 3 of BASELINE.json at full size (436x1024): 's'-labelled rotation and smooth non-affine fields of SURVEY
    section 8d-3, a float32 x3 image and the 5 % random flow mask that is reported separately."""
    h, w = 436, 1024
    rng = np.random.default_rng(33)
    d = {}
    d['aff'] = R.from_transforms([['rotation', 512, 218, -8]], (h, w), 's')
    d['smooth'] = smooth_field(h, w)
    d['img_f32c3'] = (rng.random((h, w, 3)) * 255).astype(np.float32)
    d['mask'] = rng.random((h, w)) > 0.05
    return d


def sample_pixels(h, w, seed=77, frac=0.08, band=6):
    """Flat pixel indices at which full-size goldens keep VALUES (masks are kept complete, bit-packed): a seeded
    random subset plus every pixel of a border band (hull effects live there)."""
    rng = np.random.default_rng(seed)
    pick = rng.random((h, w)) < frac
    pick[:band] = pick[-band:] = True
    pick[:, :band] = pick[:, -band:] = True
    return np.flatnonzero(pick.ravel())


def cfg4_pair_s(idx, h=1080, w=1920):
    """cfg4_frame's two flows in the other direction ('s' generators of the same transforms), same masks."""
    rng = np.random.default_rng(1000 + idx)
    fa = R.from_transforms(cfg4_transforms(idx), (h, w), 's')
    fb = R.from_transforms(cfg4_transforms(idx + 100000), (h, w), 's')
    fam = rng.random((h, w)) > 0.02
    fbm = rng.random((h, w)) > 0.02
    return fa, fam, fb, fbm


def cfg4_transforms(idx):
    """Per-frame rotation U(-10,10) deg about centre, scaling U(0.9,1.1), translation U(-20,20) px; seed 4+idx."""
    rng = np.random.default_rng(4 + idx)
    ang, sc, tx, ty = rng.uniform(-10, 10), rng.uniform(0.9, 1.1), rng.uniform(-20, 20), rng.uniform(-20, 20)
    return [['rotation', 960, 540, float(ang)], ['scaling', 960, 540, float(sc)],
            ['translation', float(tx), float(ty)]]


def cfg4_frame(idx, h=1080, w=1920):
    """One frame of the batched 1080p workload: two flows + 2 % invalid masks + uint8x3 image."""
    rng = np.random.default_rng(1000 + idx)
    fa = R.from_transforms(cfg4_transforms(idx), (h, w), 't')
    fb = R.from_transforms(cfg4_transforms(idx + 100000), (h, w), 't')
    fam = rng.random((h, w)) > 0.02
    fbm = rng.random((h, w)) > 0.02
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    return fa, fam, fb, fbm, img


def cfg5_image():
    h, w = CFG5_SHAPE
    return np.random.default_rng(5).integers(0, 256, (h, w, 3), dtype=np.uint8)
