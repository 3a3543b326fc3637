"""Dev probe: forward-resampling ('s') parity statistics of the CUDA path against the golden vectors / oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden  # noqa: E402
import oflibnumpy_b200 as of  # noqa: E402
from oracle import flowref as R  # noqa: E402


def stat(name, got_v, got_m, want_v, want_m):
    mm = int((got_m != want_m).sum())
    both = got_m & want_m
    err = np.abs(got_v.astype(np.float64) - want_v.astype(np.float64))
    e_valid = float(err[both].max()) if both.any() else 0.0
    where = np.argwhere(got_m != want_m)
    print("%-34s mask mismatches %5d / %d   max |dv| on valid %.3e   p99 %.3e  %s" % (
        name, mm, got_m.size, e_valid, float(np.percentile(err[both], 99)) if both.any() else 0,
        where[:6].tolist() if mm else ''))


g = load_golden('forward')
for pair in ('aff', 'smooth'):
    fs = of.Flow(g['in_' + pair], 's')
    ft = of.Flow(g['in_' + pair], 't')
    for key, fl in (('invert_ss_', fs.invert()), ('invert_tt_', ft.invert()), ('switch_s_', fs.switch_ref()),
                    ('switch_t_', ft.switch_ref())):
        stat(key + pair, fl.vecs, fl.mask, g['out_' + key + pair + '_vecs'], g['out_' + key + pair + '_mask'])
    for key, got in (('valid_target_s_', fs.valid_target()), ('valid_source_t_', ft.valid_source())):
        want = g['out_' + key + pair]
        print("%-34s mask mismatches %5d / %d" % (key + pair, int((got != want).sum()), want.size))
    w, m = fs.apply(g['in_img_f32c3'], return_valid_area=True)
    stat('applyva_s_f32c3_' + pair, w, m[..., None] & np.ones(3, bool), g['out_applyva_s_f32c3_' + pair],
         g['out_applyva_s_f32c3_' + pair + '_valid'][..., None] & np.ones(3, bool))
    fsm = of.Flow(g['in_' + pair], 's', g['in_mask'])
    for key, got in (('valid_target_s_masked_nocm_', fsm.valid_target(False)),
                     ('valid_target_s_masked_cm_', fsm.valid_target(True))):
        want = g['out_' + key + pair]
        print("%-34s mask mismatches %5d / %d" % (key + pair, int((got != want).sum()), want.size))
    fl = fsm.apply(-fsm, consider_mask=False)
    stat('invert_ss_masked_nocm_' + pair, fl.vecs, fl.mask, g['out_invert_ss_masked_nocm_' + pair + '_vecs'],
         g['out_invert_ss_masked_nocm_' + pair + '_mask'])

g = load_golden('small_masks')
fs, ft = of.Flow(g['in_vecs_s'], 's'), of.Flow(g['in_vecs_t'], 't')
fsm, ftm = of.Flow(g['in_vecs_s'], 's', g['in_mask_s']), of.Flow(g['in_vecs_t'], 't', g['in_mask_t'])
for key, got in (('vt_s', fs.valid_target()), ('vs_t', ft.valid_source()), ('vt_s_masked', fsm.valid_target(False)),
                 ('vs_t_masked', ftm.valid_source(False)), ('vt_s_masked_cm', fsm.valid_target()),
                 ('vs_t_masked_cm', ftm.valid_source())):
    print("7x7 %-20s mismatches %d" % (key, int((got != g['out_' + key]).sum())))

g = load_golden('combine12')
for r in ('s', 't'):
    f1, f2, f3 = (of.Flow(g['in_f%d_%s' % (i, r)], r) for i in (1, 2, 3))
    for key, fn in (('m1_', lambda: f2.combine_with(f3, 1)), ('m2_', lambda: f1.combine_with(f3, 2))):
        try:
            fl = fn()
            stat(key + r, fl.vecs, fl.mask, g['out_' + key + r + '_vecs'], g['out_' + key + r + '_mask'])
        except NotImplementedError as e:
            print(key + r, 'not implemented:', e)

# identities
rng = np.random.default_rng(0)
img = rng.random((30, 41, 3)).astype(np.float32)
for sh in ((0, 0), (3, -2)):
    flow = np.zeros((30, 41, 2), np.float32)
    flow[..., 0], flow[..., 1] = sh[0], sh[1]
    out = of.apply_flow(flow, img, 's')
    want = R.warp_array(flow, img, 's')
    print('translation', sh, 'max diff', float(np.abs(out - want).max()))
