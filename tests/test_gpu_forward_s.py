"""GPU parity of the source-referenced path (forward resampling, invert / switch_ref, combine modes 1 and 2, track,
resize) against outputs of the unmodified reference (tests/golden) and the CPU oracle.

Tolerances: flow and warped values within 1e-3 on pixels valid on both sides (north_star). Masks are bit-exact where
the reference's triangulation is determined by the data; the two documented exceptions are asserted as such:
  * pockets between the displaced image border and its convex hull (Qhull fills them with long triangles, the
    rasteriser leaves them invalid) -- a few border pixels on non-affine fields;
  * `consider_mask=True` with removed points (Qhull bridges the gaps, the rasteriser leaves the touched cells empty).
"""
import numpy as np
import pytest

import golden_inputs as gi
from conftest import load_golden
from oracle import flowref as R

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope='module')
def of():
    import oflibnumpy_b200 as of
    of.device.require_gpu()
    return of


def close_on_valid(got_v, got_m, want_v, want_m, tol=TOL):
    both = got_m & want_m
    assert both.any()
    err = np.abs(got_v.astype(np.float64) - want_v.astype(np.float64))[both]
    assert err.max() <= tol, err.max()


def border_band(shape, width):
    m = np.ones(shape, bool)
    m[width:-width, width:-width] = False
    return m


def check_flow(g, key, fl, exact_mask, band=4, max_frac=0.01):
    want_v, want_m = g['out_' + key + '_vecs'], g['out_' + key + '_mask']
    assert fl.ref == str(g['out_' + key + '_ref'])
    close_on_valid(fl.vecs, fl.mask, want_v, want_m)
    diff = fl.mask != want_m
    if exact_mask:
        assert not diff.any(), int(diff.sum())
    else:
        # hull pockets only: reference valid, ours invalid, close to the border, few
        assert not (diff & fl.mask).any()
        assert not (diff & ~border_band(diff.shape, band)).any()
        assert diff.mean() <= max_frac


def test_affine_field_bit_exact_masks(of):
    g = load_golden('forward')
    fs, ft = of.Flow(g['in_aff'], 's'), of.Flow(g['in_aff'], 't')
    check_flow(g, 'invert_ss_aff', fs.invert(), True)
    check_flow(g, 'invert_tt_aff', ft.invert(), True)
    check_flow(g, 'invert_st_aff', fs.invert('t'), True)
    check_flow(g, 'invert_ts_aff', ft.invert('s'), True)
    check_flow(g, 'switch_s_aff', fs.switch_ref(), True)
    check_flow(g, 'switch_t_aff', ft.switch_ref(), True)
    np.testing.assert_array_equal(fs.valid_target(), g['out_valid_target_s_aff'])
    np.testing.assert_array_equal(ft.valid_source(), g['out_valid_source_t_aff'])
    np.testing.assert_array_equal(fs.valid_source(), g['out_valid_source_s_aff'])
    np.testing.assert_array_equal(ft.valid_target(), g['out_valid_target_t_aff'])
    # functional wrappers
    np.testing.assert_allclose(of.invert_flow(g['in_aff'], 's')[g['out_invert_ss_aff_mask']],
                               g['out_invert_ss_aff_vecs'][g['out_invert_ss_aff_mask']], rtol=0, atol=TOL)
    np.testing.assert_allclose(of.switch_flow_ref(g['in_aff'], 't')[g['out_switch_t_aff_mask']],
                               g['out_switch_t_aff_vecs'][g['out_switch_t_aff_mask']], rtol=0, atol=TOL)
    np.testing.assert_array_equal(of.valid_target(g['in_aff'], 's'), g['out_valid_target_s_aff'])


def test_smooth_field_values_and_hull_pockets(of):
    g = load_golden('forward')
    fs, ft = of.Flow(g['in_smooth'], 's'), of.Flow(g['in_smooth'], 't')
    check_flow(g, 'invert_ss_smooth', fs.invert(), False)
    check_flow(g, 'invert_tt_smooth', ft.invert(), False)
    check_flow(g, 'switch_s_smooth', fs.switch_ref(), False)
    check_flow(g, 'switch_t_smooth', ft.switch_ref(), False)
    # image payload: on a non-degenerate field the Delaunay diagonal is determined, values agree everywhere
    w, m = fs.apply(g['in_img_f32c3'], return_valid_area=True)
    close_on_valid(w, m[..., None] & np.ones(3, bool), g['out_applyva_s_f32c3_smooth'],
                   g['out_applyva_s_f32c3_smooth_valid'][..., None] & np.ones(3, bool))
    w8 = fs.apply(g['in_img_u8c3'])
    both = m & g['out_applyva_s_f32c3_smooth_valid']
    assert np.abs(w8.astype(int) - g['out_apply_s_u8c3_smooth'].astype(int))[both].max() <= 1  # rounding of x.5 ties


def test_image_payload_on_cocircular_cells_is_bounded(of):
    """Pure rotations leave ~7 % of the cells exactly co-circular in float64 (|in-circle| < 1e-13): Qhull's diagonal
    there is an artefact of its merge order. Everywhere else values agree; masks agree everywhere."""
    g = load_golden('forward')
    fs = of.Flow(g['in_aff'], 's')
    w, m = fs.apply(g['in_img_f32c3'], return_valid_area=True)
    np.testing.assert_array_equal(m, g['out_applyva_s_f32c3_aff_valid'])
    err = np.abs(w - g['out_applyva_s_f32c3_aff']).max(axis=-1)[m]
    assert (err <= TOL).mean() >= 0.85


def test_reference_7x7_golden_masks_s_side(of):
    """tests/test_flow_class.py:852-980 of the reference, the cases resampled in 's' direction."""
    g = load_golden('small_masks')
    fs, ft = of.Flow(g['in_vecs_s'], 's'), of.Flow(g['in_vecs_t'], 't')
    fsm, ftm = of.Flow(g['in_vecs_s'], 's', g['in_mask_s']), of.Flow(g['in_vecs_t'], 't', g['in_mask_t'])
    np.testing.assert_array_equal(fs.valid_target(), g['out_vt_s'])
    np.testing.assert_array_equal(ft.valid_source(), g['out_vs_t'])
    np.testing.assert_array_equal(fsm.valid_target(False), g['out_vt_s_masked'])
    np.testing.assert_array_equal(ftm.valid_source(False), g['out_vs_t_masked'])
    # consider_mask=True: the reference marks the convex hull of the remaining points valid (documented artefact,
    # docs/usage.rst:353-361 of the reference); the rasteriser keeps removed cells invalid -> subset relation
    for got, key in ((fsm.valid_target(), 'out_vt_s_masked_cm'), (ftm.valid_source(), 'out_vs_t_masked_cm')):
        assert not (got & ~g[key]).any()
        assert int((got != g[key]).sum()) <= 6


def test_combine_modes_1_and_2(of):
    g = load_golden('combine12')
    for r in ('s', 't'):
        f1, f2, f3 = (of.Flow(g['in_f%d_%s' % (i, r)], r) for i in (1, 2, 3))
        check_flow(g, 'm1_' + r, f2.combine_with(f3, 1), r == 's', band=6, max_frac=0.02)
        check_flow(g, 'm2_' + r, f1.combine_with(f3, 2), r == 's', band=6, max_frac=0.02)
        check_flow(g, 'm3_' + r, f1.combine_with(f2, 3), True)
        np.testing.assert_allclose(of.combine_flows(g['in_f1_' + r], g['in_f3_' + r], 2, r)[g['out_m2_%s_mask' % r]],
                                   g['out_m2_%s_vecs' % r][g['out_m2_%s_mask' % r]], rtol=0, atol=TOL)


def test_combine_against_analytic_composition(of):
    """The reference's own criterion (tests/test_flow_class.py:1020-1057): combined flows equal the flow of the
    composed transforms within 5e-2 on the valid area, all modes, both references."""
    shape = (96, 96)
    transforms = [['rotation', 47.5, 47.5, -30], ['scaling', 20, 20, 0.8]]
    for ref in ('s', 't'):
        f1 = of.Flow.from_transforms(transforms[0:1], shape, ref)
        f2 = of.Flow.from_transforms(transforms[1:2], shape, ref)
        f3 = of.Flow.from_transforms(transforms, shape, ref)
        for mode, (a, b, want) in ((1, (f2, f3, f1)), (2, (f1, f3, f2)), (3, (f1, f2, f3))):
            got = a.combine_with(b, mode)
            assert isinstance(got, of.Flow) and got.ref == ref
            comb = got.mask & want.mask
            assert comb.mean() > 0.2
            np.testing.assert_allclose(got.vecs[comb], want.vecs[comb], atol=5e-2)


def test_switch_ref_and_invert_against_analytic(of):
    """tests/test_flow_class.py:509-573 of the reference: rtol = atol = 1e-3 against the analytic opposite flows."""
    shape = (128, 128)
    tr = [['rotation', 64, 64, 30]]
    tr_inv = [['rotation', 64, 64, -30]]
    f_s, f_t = of.Flow.from_transforms(tr, shape, 's'), of.Flow.from_transforms(tr, shape, 't')
    b_s, b_t = of.Flow.from_transforms(tr_inv, shape, 's'), of.Flow.from_transforms(tr_inv, shape, 't')
    for got, want in ((f_s.switch_ref(), f_t), (f_t.switch_ref(), f_s), (f_s.invert(), b_s), (f_t.invert(), b_t),
                      (f_s.invert('t'), b_t), (f_t.invert('s'), b_s)):
        assert got.ref == want.ref
        np.testing.assert_allclose(got.vecs[got.mask], want.vecs[got.mask], rtol=1e-3, atol=1e-3)
    z = of.Flow.zero(shape, 's')
    assert z.switch_ref().ref == 't' and z.switch_ref(mode='invalid').ref == 't'
    with pytest.raises(ValueError):
        z.switch_ref('test')


def test_identity_and_integer_translation_exact(of):
    """tests/test_utils.py:277-283 of the reference demands exact equality with an integer shift, both refs."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (40, 53, 3), dtype=np.uint8)
    imgf = rng.random((40, 53)).astype(np.float32)
    for ref in ('s', 't'):
        for sh in ((0, 0), (10, 0), (0, -7), (4, 9)):
            flow = np.zeros((40, 53, 2), np.float32)
            flow[..., 0], flow[..., 1] = sh
            for im in (img, imgf):
                got = of.apply_flow(flow, im, ref)
                want = R.warp_array(flow, im, ref)
                np.testing.assert_array_equal(got, want)


def test_next_rows_track_resize_padding(of):
    g = load_golden('next_rows')
    for r in ('s', 't'):
        f = of.Flow(g['in_flow'], r, g['in_mask'])
        np.testing.assert_allclose(f.track(g['in_pts_f']), g['out_track_f_' + r], rtol=0, atol=1e-3 if r == 't' else 1e-9)
        pts, st = f.track(g['in_pts_f'], get_valid_status=True)
        np.testing.assert_array_equal(st, g['out_track_status_' + r])
        got_int = f.track(g['in_pts_f'], int_out=True)
        assert got_int.dtype == g['out_track_int_' + r].dtype
        assert np.abs(got_int - g['out_track_int_' + r]).max() <= (1 if r == 't' else 0)
        np.testing.assert_allclose(f.track(g['in_pts_i']), g['out_track_ipts_' + r], rtol=0, atol=1e-3)
        for j, sc in enumerate(gi.RESIZE_SCALES):
            rz = f.resize(sc)
            want_v, want_m = g['out_resize_%d_%s_vecs' % (j, r)], g['out_resize_%d_%s_mask' % (j, r)]
            assert rz.vecs.shape == want_v.shape
            np.testing.assert_allclose(rz.vecs, want_v, rtol=0, atol=1e-5)
            np.testing.assert_array_equal(rz.mask, want_m)
            np.testing.assert_allclose(of.resize_flow(g['in_flow'], sc), want_v, rtol=0, atol=1e-5)
        assert f.get_padding() == list(g['out_padding_' + r])
    fs = of.Flow(g['in_flow'], 's')
    np.testing.assert_allclose(fs.track(g['in_pts_f'], s_exact_mode=True), g['out_track_exact_s'], rtol=0, atol=1e-3)
    np.testing.assert_allclose(of.track_pts(g['in_flow'], 's', g['in_pts_f']), g['out_track_f_s'], rtol=0, atol=1e-9)
    with pytest.raises(IndexError):
        fs.track(np.array([[-1.0, 3.0]]))
    with pytest.raises(TypeError):
        fs.track(g['in_pts_f'], True, get_valid_status='test')
    # the reference's own known answers (tests/test_utils.py:543-571): rotation 30 deg about the origin
    f_s = of.from_transforms([['rotation', 0, 0, 30]], (512, 512), 's')
    f_t = of.from_transforms([['rotation', 0, 0, 30]], (512, 512), 't')
    pts = np.array([[20.5, 10.5], [8.3, 7.2], [120.4, 160.2]])
    desired = np.array([[12.5035207776, 19.343266740], [3.58801085141, 10.385382907], [24.1694586156, 198.93726969]])
    np.testing.assert_allclose(of.track_pts(f_s, 's', pts), desired, atol=1e-1, rtol=1e-2)
    np.testing.assert_allclose(of.track_pts(f_s, 's', pts, s_exact_mode=True), desired, atol=1e-5)
    np.testing.assert_allclose(of.track_pts(f_t, 't', pts), desired, atol=1e-5, rtol=1e-6)
    got = of.track_pts(f_t, 't', pts, int_out=True)
    np.testing.assert_array_equal(got, np.round(desired))
    assert np.issubdtype(got.dtype, np.integer)
    f = of.from_transforms([['translation', 10, 20]], (512, 512), 's')
    np.testing.assert_array_equal(of.track_pts(f, 's', np.array([[20, 10], [8, 7]])), [[40, 20], [28, 17]])
