"""ORACLE tooling: generate golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz and tests/golden/MANIFEST.json

Each ``.npz`` holds the seeded inputs (``in_*``) and the reference's outputs (``out_*``) for one scenario, at
sizes small enough to commit. Full-size configurations of BASELINE.json are pinned by SHA-256 digests of the
reference's outputs on inputs regenerated from a seed (``tests/golden/digests.json``); see
``tests/golden_inputs.py`` for the shared input generators (imported by this script and by the tests, so both
sides build identical arrays).
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, '/root/reference/src')

import oflibnumpy as ref  # noqa: E402  the real reference
import golden_inputs as gi  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def flow_out(prefix, fl, d):
    d['out_' + prefix + '_vecs'] = fl.vecs
    d['out_' + prefix + '_mask'] = fl.mask
    d['out_' + prefix + '_ref'] = np.array(fl.ref)


def case_warp_t():
    """cfg 1 shape family: apply_flow / Flow.apply ref 't' over dtypes, channel counts and mask plumbing."""
    d = {}
    inp = gi.warp_t_inputs()
    d.update({'in_' + k: v for k, v in inp.items()})
    flow, fmask = inp['flow'], inp['flow_mask']
    fl = ref.Flow(flow, 't', fmask)
    fl_nomask = ref.Flow(flow, 't')
    for name in ('img_u8c3', 'img_u8c1', 'img_u8c4', 'img_f32c3', 'img_f32c1', 'img_f64c2', 'img_i16c3',
                 'img_u16c2', 'img_u8_2d', 'img_f32_2d'):
        img = inp[name]
        d['out_applyflow_' + name] = ref.apply_flow(flow, img, 't')
        d['out_apply_' + name] = fl.apply(img)
        try:
            w, m = fl.apply(img, return_valid_area=True)
            d['out_applyva_' + name] = w
            d['out_applyva_' + name + '_valid'] = m
        except Exception as e:  # e.g. uint16 || int8 promotes to int32, which cv2.remap rejects
            d['err_applyva_' + name] = np.array(type(e).__name__)
        if name in ('img_u8c3', 'img_f32c3', 'img_u8_2d', 'img_f64c2'):
            w, m = fl.apply(img, target_mask=inp['target_mask'], return_valid_area=True)
            d['out_applyvatm_' + name] = w
            d['out_applyvatm_' + name + '_valid'] = m
            w, m = fl_nomask.apply(img, target_mask=inp['target_mask'], return_valid_area=True)
            d['out_applyvatm_nofm_' + name] = w
            d['out_applyvatm_nofm_' + name + '_valid'] = m
    # warping a flow with a flow
    tgt = ref.Flow(inp['flow2'], 't', inp['flow2_mask'])
    flow_out('apply_flowobj', fl.apply(tgt), d)
    tgt_s = ref.Flow(inp['flow2'], 's', inp['flow2_mask'])
    flow_out('apply_flowobj_s', fl.apply(tgt_s), d)
    # valid areas, is_zero, padding
    d['out_valid_target'] = fl.valid_target()
    d['out_valid_source_of_s'] = ref.Flow(flow, 's', fmask).valid_source()
    d['out_get_padding_t'] = np.array(fl.get_padding())
    d['out_get_padding_s'] = np.array(ref.Flow(flow, 's', fmask).get_padding())
    return d


def case_warp_t_padded():
    """Flow smaller than the target: padding / cut variants (flow_class.py:651-664,671-680)."""
    d = {}
    inp = gi.padded_inputs()
    d.update({'in_' + k: v for k, v in inp.items()})
    pad = [int(x) for x in inp['padding']]
    fl = ref.Flow(inp['flow'], 't', inp['flow_mask'])
    for cut in (True, False):
        tag = 'cut' if cut else 'nocut'
        d['out_apply_u8_' + tag] = fl.apply(inp['img_u8c3'], padding=pad, cut=cut)
        w, m = fl.apply(inp['img_u8c3'], return_valid_area=True, padding=pad, cut=cut)
        d['out_applyva_u8_' + tag] = w
        d['out_applyva_u8_' + tag + '_valid'] = m
        w, m = fl.apply(inp['img_f32c3'], target_mask=inp['target_mask'].copy(), return_valid_area=True,
                        padding=pad, cut=cut)
        d['out_applyvatm_f32_' + tag] = w
        d['out_applyvatm_f32_' + tag + '_valid'] = m
        tgt = ref.Flow(inp['big_flow'], 't', inp['big_flow_mask'])
        flow_out('apply_flowobj_' + tag, fl.apply(tgt, padding=pad, cut=cut), d)
    d['out_pad_constant_vecs'] = fl.pad(pad).vecs
    d['out_pad_constant_mask'] = fl.pad(pad).mask
    d['out_pad_edge_vecs'] = fl.pad(pad, 'edge').vecs
    d['out_pad_symmetric_vecs'] = fl.pad(pad, 'symmetric').vecs
    return d


def case_combine3():
    """cfg 2 shape family: combine_with(mode=3), both refs, affine and smooth non-affine fields, early exits."""
    d = {}
    inp = gi.combine_inputs()
    d.update({'in_' + k: v for k, v in inp.items()})
    for r in ('t', 's'):
        for pair in ('aff', 'smooth'):
            a = ref.Flow(inp[pair + '_1'], r, inp['mask_1'])
            b = ref.Flow(inp[pair + '_2'], r, inp['mask_2'])
            flow_out('c3_{}_{}'.format(r, pair), a.combine_with(b, 3), d)
            d['out_cf3_{}_{}'.format(r, pair)] = ref.combine_flows(inp[pair + '_1'], inp[pair + '_2'], 3, r)
        # early exits: A zero where valid (mask hides a non-zero vector); B zero
        a = ref.Flow(inp['zero_where_valid'], r, inp['zero_mask'])
        b = ref.Flow(inp['aff_2'], r, inp['mask_2'])
        flow_out('c3_{}_Azero'.format(r), a.combine_with(b, 3), d)
        flow_out('c3_{}_Bzero'.format(r), b.combine_with(a, 3), d)
        # thresholded early exit
        a = ref.Flow(inp['tiny'], r, inp['mask_1'])
        flow_out('c3_{}_tiny_thr'.format(r), a.combine_with(b, 3, thresholded=True), d)
        flow_out('c3_{}_tiny_nothr'.format(r), a.combine_with(b, 3, thresholded=False), d)
    return d


def case_forward():
    """cfg 3 shape family: ref 's' resampling (griddata): apply, invert, switch_ref, valid areas."""
    d = {}
    inp = gi.forward_inputs()
    d.update({'in_' + k: v for k, v in inp.items()})
    for pair in ('aff', 'smooth'):
        fs = ref.Flow(inp[pair], 's')
        ft = ref.Flow(inp[pair], 't')
        d['out_apply_s_f32c3_' + pair] = fs.apply(inp['img_f32c3'])
        w, m = fs.apply(inp['img_f32c3'], return_valid_area=True)
        d['out_applyva_s_f32c3_' + pair] = w
        d['out_applyva_s_f32c3_' + pair + '_valid'] = m
        d['out_apply_s_u8c3_' + pair] = fs.apply(inp['img_u8c3'])
        flow_out('invert_ss_' + pair, fs.invert(), d)
        flow_out('invert_st_' + pair, fs.invert('t'), d)
        flow_out('invert_tt_' + pair, ft.invert(), d)
        flow_out('invert_ts_' + pair, ft.invert('s'), d)
        flow_out('switch_s_' + pair, fs.switch_ref(), d)
        flow_out('switch_t_' + pair, ft.switch_ref(), d)
        d['out_valid_target_s_' + pair] = fs.valid_target()
        d['out_valid_source_t_' + pair] = ft.valid_source()
        d['out_valid_source_s_' + pair] = fs.valid_source()
        d['out_valid_target_t_' + pair] = ft.valid_target()
        # masked variants (reported separately: hard part 3 of SURVEY section 7)
        fsm = ref.Flow(inp[pair], 's', inp['mask'])
        d['out_valid_target_s_masked_nocm_' + pair] = fsm.valid_target(consider_mask=False)
        d['out_valid_target_s_masked_cm_' + pair] = fsm.valid_target(consider_mask=True)
        flow_out('invert_ss_masked_nocm_' + pair, fsm.apply(-fsm, consider_mask=False), d)
    return d


def case_combine12():
    """modes 1 and 2, both refs (griddata chains)."""
    d = {}
    inp = gi.combine12_inputs()
    d.update({'in_' + k: v for k, v in inp.items()})
    for r in ('s', 't'):
        f1 = ref.Flow(inp['f1_' + r], r)
        f2 = ref.Flow(inp['f2_' + r], r)
        f3 = ref.Flow(inp['f3_' + r], r)
        flow_out('m1_' + r, f2.combine_with(f3, 1), d)
        flow_out('m2_' + r, f1.combine_with(f3, 2), d)
        flow_out('m3_' + r, f1.combine_with(f2, 3), d)
    return d


def case_generators():
    """from_matrix / from_transforms over shapes, refs and a projective matrix."""
    d = {}
    specs = gi.generator_specs()
    for i, (kind, arg, shape, r) in enumerate(specs):
        if kind == 'transforms':
            d['out_gen_%d' % i] = ref.from_transforms(arg, shape, r)
        else:
            d['out_gen_%d' % i] = ref.from_matrix(np.array(arg, dtype=np.float64), shape, r)
    return d


def case_small_masks():
    """The reference's own 7x7 scenario (tests/test_flow_class.py:852-980), outputs recomputed by the reference."""
    d = {}
    transforms = [['rotation', 0, 0, 45]]
    shape = (7, 7)
    ms = np.ones(shape, bool); ms[4:, :3] = False
    mt = np.ones(shape, bool); mt[:3, 4:] = False
    d['in_mask_s'] = ms
    d['in_mask_t'] = mt
    fs = ref.Flow.from_transforms(transforms, shape, 's')
    ft = ref.Flow.from_transforms(transforms, shape, 't')
    fsm = ref.Flow.from_transforms(transforms, shape, 's', ms)
    ftm = ref.Flow.from_transforms(transforms, shape, 't', mt)
    d['in_vecs_s'] = fs.vecs
    d['in_vecs_t'] = ft.vecs
    d['out_vt_s'] = fs.valid_target(); d['out_vt_t'] = ft.valid_target()
    d['out_vt_s_masked_cm'] = fsm.valid_target(); d['out_vt_s_masked'] = fsm.valid_target(False)
    d['out_vt_t_masked'] = ftm.valid_target()
    d['out_vs_s'] = fs.valid_source(); d['out_vs_t'] = ft.valid_source()
    d['out_vs_s_masked'] = fsm.valid_source()
    d['out_vs_t_masked_cm'] = ftm.valid_source(); d['out_vs_t_masked'] = ftm.valid_source(False)
    return d


def case_next_rows():
    """'next' rows of SURVEY section 8f: track, resize, get_padding."""
    d = {}
    inp = gi.next_inputs()
    d.update({'in_' + k: v for k, v in inp.items()})
    for r in ('s', 't'):
        fl = ref.Flow(inp['flow'], r, inp['mask'])
        d['out_track_f_' + r] = fl.track(inp['pts_f'])
        p, st = fl.track(inp['pts_f'], get_valid_status=True)
        d['out_track_status_' + r] = st
        d['out_track_int_' + r] = fl.track(inp['pts_f'], int_out=True)
        d['out_track_ipts_' + r] = fl.track(inp['pts_i'])
        for j, sc in enumerate(gi.RESIZE_SCALES):
            rz = fl.resize(sc)
            d['out_resize_%d_%s_vecs' % (j, r)] = rz.vecs
            d['out_resize_%d_%s_mask' % (j, r)] = rz.mask
        d['out_padding_' + r] = np.array(fl.get_padding())
    d['out_track_exact_s'] = ref.Flow(inp['flow'], 's').track(inp['pts_f'], s_exact_mode=True)
    return d


def digests():
    """Full-size configurations: SHA-256 of the reference's outputs (bit-exact paths only)."""
    out = {}
    # cfg 1 at full size, uint8 x3 image (integer output -> hashable)
    flow, img = gi.cfg1_full()
    t0 = time.time()
    out['cfg1_flow_sha'] = sha(flow)
    out['cfg1_apply_flow_u8c3'] = sha(ref.apply_flow(flow, img, 't'))
    w, m = ref.Flow(flow, 't').apply(img, return_valid_area=True)
    out['cfg1_apply_va_u8c3'] = sha(w)
    out['cfg1_apply_va_valid'] = sha(m)
    # cfg 2 full size: masks of combine mode 3 are hashable (bit-exact), vecs sampled
    a, am, b, bm = gi.cfg2_full()
    res = ref.Flow(a, 't', am).combine_with(ref.Flow(b, 't', bm), 3)
    out['cfg2_c3_t_mask'] = sha(res.mask)
    out['cfg2_c3_t_vecs'] = sha(res.vecs)
    res = ref.Flow(a, 's', am).combine_with(ref.Flow(b, 's', bm), 3)
    out['cfg2_c3_s_mask'] = sha(res.mask)
    out['cfg2_c3_s_vecs'] = sha(res.vecs)
    # cfg 4: one 1080p frame (frame index 0 and 5 of the seeded batch)
    for idx in (0, 5):
        fa, fam, fb, fbm, img = gi.cfg4_frame(idx)
        w, m = ref.Flow(fa, 't', fam).apply(img, return_valid_area=True)
        out['cfg4_f%d_apply_va_u8c3' % idx] = sha(w)
        out['cfg4_f%d_apply_va_valid' % idx] = sha(m)
        res = ref.Flow(fa, 't', fam).combine_with(ref.Flow(fb, 't', fbm), 3)
        out['cfg4_f%d_c3_t_vecs' % idx] = sha(res.vecs)
        out['cfg4_f%d_c3_t_mask' % idx] = sha(res.mask)
    # cfg 5 generators at 4K
    for r in ('t', 's'):
        out['cfg5_from_transforms_' + r] = sha(ref.from_transforms(gi.CFG5_TRANSFORMS, gi.CFG5_SHAPE, r))
    # cfg 5 chain at 4K: from_transforms -> invert('s'/'t' cheap legs) -> combine x4 (mode 3) -> apply u8x3
    f = ref.Flow.from_transforms(gi.CFG5_TRANSFORMS, gi.CFG5_SHAPE, 't')
    g = f.invert('s')
    g = ref.Flow(g.vecs, 't', g.mask)  # relabelled negated field: the cheap "inverse" used as second operand
    acc = f
    for i in range(4):
        acc = acc.combine_with(g if i % 2 == 0 else f, 3)
    img = gi.cfg5_image()
    w, m = acc.apply(img, return_valid_area=True)
    out['cfg5_chain_vecs'] = sha(acc.vecs)
    out['cfg5_chain_mask'] = sha(acc.mask)
    out['cfg5_chain_img'] = sha(w)
    out['cfg5_chain_valid'] = sha(m)
    out['_seconds'] = round(time.time() - t0, 1)
    return out


def case_datasets():
    """Dataset formats (utils.py:426-490): small synthetic KITTI / Sintel files are written next to the vectors
    (tests/golden/files/) and read back by the reference's own loaders and Flow constructors."""
    import cv2
    files = os.path.join(OUT, 'files')
    os.makedirs(files, exist_ok=True)
    rng = np.random.default_rng(31)
    h, w = 24, 40
    kitti = rng.integers(0, 65536, (h, w, 3), dtype=np.uint16)            # BGR on disk: B = valid, G = v, R = u
    kitti[..., 0] = rng.integers(0, 2, (h, w))                            # valid channel: 0 / 1 as in the dataset
    kitti[3, 5] = (7, 32768, 32768)                                       # zero flow, "valid" stored as 7
    kpath = os.path.join(files, 'kitti_sample.png')
    assert cv2.imwrite(kpath, kitti)
    flo = ((rng.random((h, w, 2)) - 0.5) * 40).astype('<f4')
    spath = os.path.join(files, 'sintel_sample.flo')
    with open(spath, 'wb') as fh:
        fh.write(b'PIEH')
        fh.write(int(w).to_bytes(4, 'little'))
        fh.write(int(h).to_bytes(4, 'little'))
        fh.write(flo.tobytes())
    inv = (rng.random((h, w)) > 0.8).astype(np.uint8) * 255
    ipath = os.path.join(files, 'sintel_invalid.png')
    assert cv2.imwrite(ipath, inv)
    d = {}
    d['out_load_kitti'] = ref.utils.load_kitti(kpath)
    flow_out('kitti_valid', ref.Flow.from_kitti(kpath), d)
    flow_out('kitti_novalid', ref.Flow.from_kitti(kpath, load_valid=False), d)
    d['out_load_sintel'] = ref.utils.load_sintel(spath)
    d['out_load_sintel_mask'] = ref.utils.load_sintel_mask(ipath)
    flow_out('sintel', ref.Flow.from_sintel(spath), d)
    flow_out('sintel_masked', ref.Flow.from_sintel(spath, ipath), d)
    return d


def visualise_inputs(seed=41, h=45, w=64):
    """Flows for the visualisation fixtures: smooth + noise (all hues), a zero flow, a flow whose 99th percentile is 0."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[:h, :w].astype(np.float32)
    swirl = np.stack([(yy - h / 2) * 0.4 + rng.standard_normal((h, w)) * 0.3,
                      -(xx - w / 2) * 0.3 + rng.standard_normal((h, w)) * 0.3], -1).astype(np.float32)
    swirl[5:9, 7:12] = 1e-4                                                # below the threshold -> exactly zero
    mask = rng.random((h, w)) > 0.15
    mask[10:20, 30:50] = False                                             # a hole with a clean outline
    mask[0, :5] = False
    sparse = np.zeros((h, w, 2), np.float32)
    sparse.reshape(-1, 2)[rng.integers(0, h * w, 20)] = rng.uniform(-9, 9, (20, 2))   # p99 == 0, max > 0
    return {'swirl': swirl, 'sparse': sparse, 'zero': np.zeros((h, w, 2), np.float32)}, mask


def case_visualise():
    """Flow.visualise (flow_class.py:869-951) of the unmodified reference: all modes, mask options, range_max."""
    flows, mask = visualise_inputs()
    d = {'in_mask': mask}
    for name, v in flows.items():
        d['in_' + name] = v
        fl = ref.Flow(v, 't', mask)
        for mode in ('hsv', 'rgb', 'bgr'):
            d['out_%s_%s' % (name, mode)] = fl.visualise(mode)
        d['out_%s_rgb_mask' % name] = fl.visualise('rgb', show_mask=True)
        d['out_%s_hsv_borders' % name] = fl.visualise('hsv', show_mask=True, show_mask_borders=True)
        d['out_%s_bgr_range' % name] = fl.visualise('bgr', show_mask_borders=True, range_max=7.5)
    return d


def pack(mask):
    return np.packbits(np.ascontiguousarray(mask, dtype=bool).ravel())


def flow_out_sampled(prefix, fl, idx, d):
    """Full-size outputs: the mask complete (bit-packed), the vectors at the sampled pixels only."""
    d['out_' + prefix + '_vecs_s'] = fl.vecs.reshape(-1, 2)[idx]
    d['out_' + prefix + '_maskbits'] = pack(fl.mask)
    d['out_' + prefix + '_ref'] = np.array(fl.ref)


def case_cfg3_full():
    """cfg 3 at 436x1024 through the unmodified reference: apply 's' of a float32 x3 image, same-reference invert
    (s->s and t->t), switch_ref, the 's'-resampled valid areas; full masks and the 5 % random mask (consider_mask
    True and False)."""
    d = {}
    inp = gi.cfg3_full()
    h, w = inp['mask'].shape
    idx = gi.sample_pixels(h, w, frac=0.02, band=2)
    for pair in ('aff', 'smooth'):
        for tag, m in (('full', None), ('m5', inp['mask'])):
            t0 = time.time()
            fs = ref.Flow(inp[pair], 's', m)
            ft = ref.Flow(inp[pair], 't', m)
            k = pair + '_' + tag
            img, va = fs.apply(inp['img_f32c3'], return_valid_area=True)
            d['out_applyva_' + k + '_img_s'] = img.reshape(-1, 3)[idx]
            d['out_applyva_' + k + '_validbits'] = pack(va)
            flow_out_sampled('invert_ss_' + k, fs.invert(), idx, d)
            flow_out_sampled('invert_tt_' + k, ft.invert(), idx, d)
            flow_out_sampled('switch_s_' + k, fs.switch_ref(), idx, d)
            flow_out_sampled('switch_t_' + k, ft.switch_ref(), idx, d)
            d['out_valid_target_s_' + k + '_bits'] = pack(fs.valid_target())
            d['out_valid_source_t_' + k + '_bits'] = pack(ft.valid_source())
            if m is not None:
                d['out_valid_target_s_' + k + '_nocm_bits'] = pack(fs.valid_target(consider_mask=False))
                flow_out_sampled('invert_ss_' + k + '_nocm', fs.apply(-fs, consider_mask=False), idx, d)
            print('cfg3_full', k, round(time.time() - t0, 1), 's', flush=True)
    return d


def case_modes12_1080p():
    """cfg 4 as written: one 1080p frame pair (frame 0 of the seeded batch, 2 % invalid masks) through
    combine_with modes 1 and 2 of the unmodified reference, both references."""
    d = {}
    fa, fam, fb, fbm, _ = gi.cfg4_frame(0)
    sa, sam, sb, sbm = gi.cfg4_pair_s(0)
    idx = gi.sample_pixels(1080, 1920, frac=0.02)
    for r, (a, am, b, bm) in (('t', (fa, fam, fb, fbm)), ('s', (sa, sam, sb, sbm))):
        for mode in (2, 1):
            t0 = time.time()
            res = ref.Flow(a, r, am).combine_with(ref.Flow(b, r, bm), mode)
            flow_out_sampled('m%d_%s' % (mode, r), res, idx, d)
            print('modes12_1080p', mode, r, round(time.time() - t0, 1), 's', flush=True)
    return d


def case_cfg5_true():
    """cfg 5 as written at 3840x2160: from_transforms -> invert() (same reference, t->t: one griddata) -> 4 x
    combine_with(mode 3) -> apply to a uint8 x3 image with valid area."""
    d = {}
    t0 = time.time()
    f = ref.Flow.from_transforms(gi.CFG5_TRANSFORMS, gi.CFG5_SHAPE, 't')
    g = f.invert()
    print('cfg5_true invert', round(time.time() - t0, 1), 's', flush=True)
    h, w = gi.CFG5_SHAPE
    idx = gi.sample_pixels(h, w, frac=0.01)
    flow_out_sampled('g', g, idx, d)
    acc = f
    for i in range(4):
        acc = acc.combine_with(g if i % 2 == 0 else f, 3)
    flow_out_sampled('chain', acc, idx, d)
    img = gi.cfg5_image()
    wimg, m = acc.apply(img, return_valid_area=True)
    d['out_chain_img_s'] = wimg.reshape(-1, 3)[idx]
    d['out_chain_validbits'] = pack(m)
    print('cfg5_true', round(time.time() - t0, 1), 's', flush=True)
    return d


CASES = {
    'cfg3_full': case_cfg3_full,
    'modes12_1080p': case_modes12_1080p,
    'cfg5_true': case_cfg5_true,
    'visualise': case_visualise,
    'datasets': case_datasets,
    'warp_t': case_warp_t,
    'warp_t_padded': case_warp_t_padded,
    'combine3': case_combine3,
    'forward': case_forward,
    'combine12': case_combine12,
    'generators': case_generators,
    'small_masks': case_small_masks,
    'next_rows': case_next_rows,
}


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    manifest = {}
    mpath = os.path.join(OUT, 'MANIFEST.json')
    if os.path.exists(mpath):
        manifest = json.load(open(mpath))
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        d = fn()
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **d)
        manifest[name] = {'arrays': len(d), 'bytes': os.path.getsize(path), 'seconds': round(time.time() - t0, 1)}
        print(name, manifest[name])
    if not only or 'digests' in only:
        dg = digests()
        json.dump(dg, open(os.path.join(OUT, 'digests.json'), 'w'), indent=1, sort_keys=True)
        print('digests', dg['_seconds'], 's')
    import cv2
    import scipy
    manifest['_versions'] = {'numpy': np.__version__, 'cv2': cv2.__version__, 'scipy': scipy.__version__,
                             'reference': 'oflibnumpy 1.1.1 (/root/reference, unmodified)'}
    json.dump(manifest, open(mpath, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
