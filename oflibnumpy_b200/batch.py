"""Batched extension of the API: N independent flow fields of one shape processed by single kernel launches.

The reference has no batch axis; ``FlowBatch`` is how batches of frame / flow pairs are sharded
over GPUs (one process per GPU, contiguous split of the batch axis, no data-path collective). Frame ``i`` of every
result equals what the single-frame :class:`Flow` method returns for frame ``i`` of the inputs.
"""
import numpy as np

from . import _lib
from . import _ops
from . import device as dev
from .device import DeviceArray
from .flow import Flow
from .validation import get_valid_ref, validate_shape, validate_transform_list, DEFAULT_THRESHOLD

__all__ = ['FlowBatch', 'shard_range', 'apply_flow_host', 'combine_flows_host', 'apply_combine_host']


def shard_range(n, rank, world):
    """Contiguous split of ``n`` frames over ``world`` ranks -> [start, stop) of ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _to_device(x, dtype=None):
    if isinstance(x, np.ndarray):
        return DeviceArray.from_numpy(x, dtype)
    return dev.as_device(x, dtype)


class FlowBatch(object):
    def __init__(self, vecs, ref=None, masks=None, validate=True):
        """:param vecs: (N,H,W,2) float32 numpy or device array; :param masks: (N,H,W) bool/uint8 or None"""
        self.ref = get_valid_ref(ref)
        if isinstance(vecs, np.ndarray):
            if vecs.ndim != 4 or vecs.shape[3] != 2:
                raise ValueError("Error setting flow vectors: batch input needs shape (N,H,W,2)")
        v = _to_device(vecs, np.float32) if not isinstance(vecs, np.ndarray) else \
            DeviceArray.from_numpy(vecs, np.float32)
        if v.ndim != 4 or v.shape[3] != 2:
            raise ValueError("Error setting flow vectors: batch input needs shape (N,H,W,2)")
        if validate and not _ops.all_finite(v):
            raise ValueError("Error setting flow vectors: Input contains NaN, Inf or -Inf values")
        self.vecs = v
        if masks is None:
            m = DeviceArray.empty(v.shape[:3], np.uint8)
            _lib.call('ofk_rt_memset', m.ptr, 1, m.nbytes, dev.current_stream())
        elif isinstance(masks, np.ndarray):
            if masks.shape != v.shape[:3]:
                raise ValueError("Error setting flow mask: Input has a different shape than the flow vectors")
            if masks.dtype != np.bool_ and ((masks != 0) & (masks != 1)).any():
                raise ValueError("Error setting flow mask: Values must be 0 or 1")
            m = DeviceArray.from_numpy(masks.astype(np.bool_).view(np.uint8))
        else:
            d = dev.as_device(masks)
            if d.shape != v.shape[:3] or d.dtype not in (np.uint8, np.bool_):
                raise ValueError("Error setting flow mask: device mask needs shape (N,H,W) and dtype bool/uint8")
            m = DeviceArray(d.ptr, d.shape, np.uint8, owner=d)
        self.masks = m

    @classmethod
    def _wrap(cls, vecs, ref, masks):
        b = cls.__new__(cls)
        b.vecs, b.ref, b.masks = vecs, ref, masks
        return b

    # ------------------------------------------------------------------------------------------------ basics
    def __len__(self):
        return self.vecs.shape[0]

    @property
    def shape(self):
        return tuple(self.vecs.shape[1:3])

    def __getitem__(self, i):
        """Frame ``i`` as a single :class:`Flow` (device view, no copy); a slice gives a FlowBatch view."""
        n = len(self)
        if isinstance(i, slice):
            start, stop, step = i.indices(n)
            if step != 1:
                raise IndexError("FlowBatch slices need unit step")
            return FlowBatch._wrap(self.vecs.frames(start, stop), self.ref, self.masks.frames(start, stop))
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("frame index out of range")
        return Flow._wrap(self.vecs.frames(i, i + 1), self.ref, self.masks.frames(i, i + 1))

    def shard(self, rank, world):
        start, stop = shard_range(len(self), rank, world)
        return self[start:stop]

    def numpy(self):
        return self.vecs.numpy(), self.masks.numpy().view(np.bool_)

    @classmethod
    def from_transforms(cls, transform_lists, shape, ref=None, masks=None):
        """One transform list per frame (Flow.from_transforms semantics, evaluated by one launch)."""
        from .ops import matrix_from_transforms
        validate_shape(shape)
        ref = get_valid_ref(ref)
        mats = []
        for tl in transform_lists:
            validate_transform_list(tl)
            m = matrix_from_transforms(tl)
            mats.append(m if ref == 's' else np.linalg.pinv(m))
        v = _ops.from_matrix(np.stack(mats), shape, 1.0 if ref == 's' else -1.0)
        b = cls._wrap(v, ref, None)
        if masks is None:
            m = DeviceArray.empty(v.shape[:3], np.uint8)
            _lib.call('ofk_rt_memset', m.ptr, 1, m.nbytes, dev.current_stream())
            b.masks = m
        else:
            b.masks = cls(v, ref, masks, validate=False).masks
        return b

    # ------------------------------------------------------------------------------------------------ hot path
    def apply(self, targets, target_masks=None, return_valid_area=False, consider_mask=True):
        """Warp images (N,H,W,C) [numpy or device] frame by frame with this batch, like ``Flow.apply`` (ref 't':
        ofk_warp_t, ref 's': ofk_forward_s). Returns device arrays (``.numpy()`` to download): images in the dtype of
        the targets (ref 's' resamples in float32 and rounds integer targets like the reference), and the valid areas
        if requested."""
        payload = _to_device(targets)
        if payload.ndim == 3:
            payload = payload.reshape(payload.shape + (1,))
        if payload.shape[:3] != self.vecs.shape[:3]:
            raise ValueError("Error applying flow: Flow shape does not match target shape")
        if self.ref == 's':
            return self._apply_s(payload, target_masks, return_valid_area, consider_mask)
        pmask = None
        if return_valid_area:
            arith, rule = _ops.promoted_rule(payload.dtype, target_masks is not None)
            if target_masks is not None:
                pmask = _to_device(target_masks.view(np.uint8) if isinstance(target_masks, np.ndarray) else
                                   target_masks)
        else:
            arith, rule = _lib.ARITH_NATIVE, _lib.RULE_STRICT
        out, omask = _ops.warp_t(self.vecs, -1.0, payload, pmask, self.masks if return_valid_area else None,
                                 return_valid_area, arith, rule)
        return (out, omask) if return_valid_area else out

    def _zero_flags(self):
        """Device int32 [N]: 0 where the flow is zero below the threshold (apply_flow passes its target through)."""
        return _ops.nonzero_flags_device(self.vecs, None, DEFAULT_THRESHOLD)

    def _apply_s(self, payload, target_masks, return_valid_area, consider_mask):
        out_dtype = payload.dtype
        is_int = np.issubdtype(out_dtype, np.integer)
        pay32 = payload if out_dtype == np.float32 else _ops.cast(payload, np.float32)
        pmask = None
        if return_valid_area:
            pmask = self.masks
            if target_masks is not None:
                tm = _to_device(target_masks.view(np.uint8) if isinstance(target_masks, np.ndarray) else target_masks)
                pmask = _ops.mask_and(DeviceArray(tm.ptr, tm.shape, np.uint8, owner=tm), self.masks)
        rule = _lib.RULE_GT_HALF if is_int else _lib.RULE_STRICT
        out, omask = _ops.forward_s(self.vecs, 1.0, pay32, pmask, self.masks if consider_mask else None,
                                    return_valid_area, rule, flow_nonzero=self._zero_flags())
        if out_dtype != np.float32:
            out = _ops.cast(out, out_dtype, round_ints=True)
        return (out, omask) if return_valid_area else out

    def apply_to_flows(self, other, consider_mask=True):
        """``self[i].apply(other[i])`` for every frame: warp a batch of flows (vectors and masks) with this batch."""
        if self.ref != 't':
            pm = _ops.mask_and(other.masks, self.masks)
            out, omask = _ops.forward_s(self.vecs, 1.0, other.vecs, pm, self.masks if consider_mask else None, True,
                                        _lib.RULE_STRICT, flow_nonzero=self._zero_flags())
            return FlowBatch._wrap(out, other.ref, omask)
        out, omask = _ops.warp_t(self.vecs, -1.0, other.vecs, other.masks, self.masks, True, _lib.ARITH_NATIVE,
                                 _lib.RULE_STRICT)
        return FlowBatch._wrap(out, other.ref, omask)

    def combine_with(self, other, mode, thresholded=False, return_flags=False):
        """Frame-wise ``self[i].combine_with(other[i], mode)``; mode 3 is one fused launch for the whole batch, the
        zero-flow early exits of the reference are resolved on the device (no host round trip)."""
        if not isinstance(other, FlowBatch):
            raise TypeError("Error combining flows: Flow need to be of type 'FlowBatch'")
        if self.vecs.shape != other.vecs.shape:
            raise ValueError("Error combining flows: Flow fields need to have the same shape")
        if self.ref != other.ref:
            raise ValueError("Error combining flows: Flow fields need to have the same reference")
        if mode not in (1, 2, 3):
            raise ValueError("Error combining flows: Mode needs to be 1, 2 or 3")
        thr = DEFAULT_THRESHOLD if thresholded else 0.0
        if mode == 3:
            v, m, flags = _ops.combine3(self.vecs, self.masks, other.vecs, other.masks, self.ref, thr)
            res = FlowBatch._wrap(v, self.ref, m)
            return (res, flags) if return_flags else res
        # modes 1 and 2: one device-resident chain for the whole batch (ofk_combine12). The reference's early exits
        # (self zero -> flow, flow zero -> self.invert(); flow_class.py:1338-1354) are rare: the two zero tests are
        # read back once and those frames redone through the single-frame path
        v, m = _ops.combine12(mode, self.ref, self.vecs, self.masks, other.vecs, other.masks)
        a_nz = _ops.nonzero_flags(self.vecs, self.masks, thr)
        b_nz = _ops.nonzero_flags(other.vecs, other.masks, thr)
        res = FlowBatch._wrap(v, self.ref, m)
        for i in np.flatnonzero((a_nz == 0) | (b_nz == 0)):
            one = self[int(i)].combine_with(other[int(i)], mode, thresholded)
            res.vecs.frames(int(i), int(i) + 1).copy_from(one._vd())
            res.masks.frames(int(i), int(i) + 1).copy_from(one._md())
        flags = np.stack([a_nz, b_nz], 1)
        return (res, flags) if return_flags else res

    def valid_target(self, consider_mask=True):
        """Device uint8 (N,H,W): Flow.valid_target frame by frame (flow_class.py:1113-1151)."""
        if self.ref == 't':
            return _ops.valid_geom_t(self.vecs, -1.0, self.masks)
        _, area = _ops.forward_s(self.vecs, 1.0, None, self.masks, self.masks if consider_mask else None,
                                 flow_nonzero=self._zero_flags())
        return area

    def valid_source(self, consider_mask=True):
        if self.ref == 's':
            return _ops.valid_geom_t(self.vecs, 1.0, self.masks)
        _, area = _ops.forward_s(self.vecs, -1.0, None, self.masks, self.masks if consider_mask else None,
                                 flow_nonzero=self._zero_flags())
        return area

    def _negated(self):
        return _ops.scale(_lib.OP_MUL, self.vecs, -1.0, -1.0, False)

    def switch_ref(self, mode='valid'):
        """Frame-wise Flow.switch_ref (flow_class.py:697-733): one forward resampling of vectors and mask; frames whose
        flow is exactly zero on its valid pixels, or zero below the threshold everywhere, are only relabelled (the two
        tests run on the device)."""
        other = 't' if self.ref == 's' else 's'
        if mode == 'invalid':
            return FlowBatch._wrap(self.vecs, other, self.masks)
        if mode != 'valid':
            raise ValueError("Error switching flow reference: Mode not recognised, should be 'valid' or 'invalid'")
        active = _ops.and_flags(self._zero_flags(), _ops.nonzero_flags_device(self.vecs, self.masks, 0.0))
        out, omask = _ops.forward_s(self.vecs, 1.0 if self.ref == 's' else -1.0, self.vecs, self.masks, self.masks,
                                    True, _lib.RULE_STRICT, flow_nonzero=active)
        return FlowBatch._wrap(out, other, omask)

    def invert(self, ref=None):
        """Frame-wise Flow.invert (flow_class.py:735-753): cross-reference = negate + relabel, same reference = one
        forward resampling (s -> s: self.apply(-self); t -> t: self.invert('s').switch_ref())."""
        ref = self.ref if ref is None else get_valid_ref(ref)
        if ref != self.ref:
            return FlowBatch._wrap(self._negated(), ref, self.masks)
        if self.ref == 's':
            neg = self._negated()
            out, omask = _ops.forward_s(self.vecs, 1.0, neg, self.masks, self.masks, True, _lib.RULE_STRICT,
                                        flow_nonzero=self._zero_flags())
            return FlowBatch._wrap(out, 's', omask)
        return FlowBatch._wrap(self._negated(), 's', self.masks).switch_ref()


# ---------------------------------------------------------------------------------------------------- host-buffer calls
def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


def _check_out(arr, shape, dtype, name):
    """Caller-supplied output buffers become raw pointers for the copy engines: shape, dtype and layout must be exact."""
    if not isinstance(arr, np.ndarray) or arr.shape != tuple(shape) or arr.dtype != np.dtype(dtype) or \
            not arr.flags.c_contiguous or not arr.flags.writeable:
        raise ValueError("{} needs to be a writeable C-contiguous numpy array of shape {} and dtype {}".format(
            name, tuple(shape), np.dtype(dtype)))


def apply_flow_host(flows, images, flow_masks=None, target_masks=None, return_valid_area=False, out=None,
                    out_valid=None, device=None):
    """Batched ``Flow(flows[i], 't', flow_masks[i]).apply(images[i], target_masks[i], return_valid_area)`` on HOST
    arrays through ofh_warp_t: frames stream through a device ring, host<->device copies overlap the kernels.
    Pass pinned arrays (device.pinned_empty) for full-rate copies. Returns numpy arrays."""
    flows = _c(flows, np.float32)
    images = _c(images)
    if flows.ndim != 4 or flows.shape[3] != 2:
        raise ValueError("Error applying flow: flows need shape (N,H,W,2)")
    if images.ndim == 3:
        images = images[..., None]
    n, h, w = flows.shape[:3]
    if images.shape[:3] != (n, h, w):
        raise ValueError("Error applying flow: Flow shape does not match target shape")
    code = _ops.dtype_code(images.dtype)
    if return_valid_area:
        arith, rule = _ops.promoted_rule(images.dtype, target_masks is not None)
    else:
        arith, rule = _lib.ARITH_NATIVE, _lib.RULE_STRICT
    if out is None:
        out = np.empty_like(images)
    else:
        _check_out(out, images.shape if out.ndim == 4 else images.shape[:3], images.dtype, 'out')
    pm = fm = om = None
    if return_valid_area:
        if out_valid is None:
            out_valid = np.empty((n, h, w), np.bool_)
        else:
            _check_out(out_valid, (n, h, w), np.bool_, 'out_valid')
        om = out_valid.ctypes.data
        if target_masks is not None:
            target_masks = _c(target_masks, np.bool_)
            pm = target_masks.ctypes.data
        if flow_masks is not None:
            flow_masks = _c(flow_masks, np.bool_)
            fm = flow_masks.ctypes.data
    device = dev.get_device() if device is None else device
    _lib.call('ofh_warp_t', images.ctypes.data, code, images.shape[3], arith, flows.ctypes.data, -1.0, pm, fm,
              out.ctypes.data, om, rule, n, h, w, device)
    return (out, out_valid) if return_valid_area else out


def combine_flows_host(flows_1, flows_2, mode, ref, masks_1=None, masks_2=None, thresholded=False, out=None,
                       out_masks=None, device=None):
    """Batched ``combine_flows(flows_1[i], flows_2[i], mode, ref)`` / ``Flow.combine_with`` on HOST arrays. Mode 3
    streams through ofh_combine3 (pinned ring, copies overlapped with the kernel); modes 1 / 2 upload the batch, run the
    device chain of FlowBatch.combine_with and download. Returns (vecs (N,H,W,2) float32, masks (N,H,W) bool)."""
    if mode not in (1, 2, 3):
        raise ValueError("Error combining flows: Mode needs to be 1, 2 or 3")
    ref = get_valid_ref(ref)
    if mode != 3:
        res = FlowBatch(flows_1, ref, masks_1).combine_with(FlowBatch(flows_2, ref, masks_2), mode, thresholded)
        v, m = res.numpy()
        if out is not None:
            _check_out(out, v.shape, np.float32, 'out')
            out[...] = v
            v = out
        if out_masks is not None:
            _check_out(out_masks, m.shape, np.bool_, 'out_masks')
            out_masks[...] = m
            m = out_masks
        return v, m
    a, b = _c(flows_1, np.float32), _c(flows_2, np.float32)
    if a.shape != b.shape or a.ndim != 4 or a.shape[3] != 2:
        raise ValueError("Error combining flows: Flow fields need to have the same shape (N,H,W,2)")
    n, h, w = a.shape[:3]
    am = None if masks_1 is None else _c(masks_1, np.bool_)
    bm = None if masks_2 is None else _c(masks_2, np.bool_)
    if out is None:
        out = np.empty_like(a)
    else:
        _check_out(out, a.shape, np.float32, 'out')
    if out_masks is None:
        out_masks = np.empty((n, h, w), np.bool_)
    else:
        _check_out(out_masks, (n, h, w), np.bool_, 'out_masks')
    flags = np.empty((n, 2), np.int32)
    device = dev.get_device() if device is None else device
    _lib.call('ofh_combine3', a.ctypes.data, None if am is None else am.ctypes.data, b.ctypes.data,
              None if bm is None else bm.ctypes.data, ord(ref), DEFAULT_THRESHOLD if thresholded else 0.0,
              out.ctypes.data, out_masks.ctypes.data, flags.ctypes.data, n, h, w, device)
    return out, out_masks


def apply_combine_host(flows_1, flows_2, images, ref='t', masks_1=None, masks_2=None, thresholded=False, out_images=None,
                       out_valid=None, out=None, out_masks=None, device=None):
    """The pair of calls of a frame pipeline on HOST arrays in one pass (ofh_apply_combine3):
    ``Flow(flows_1[i], ref, masks_1[i]).apply(images[i], return_valid_area=True)`` and
    ``....combine_with(Flow(flows_2[i], ref, masks_2[i]), 3)``. ``flows_1`` and ``masks_1`` are uploaded once for both
    kernels (apply_flow_host + combine_flows_host upload them twice). Returns (images, valid areas, vecs, masks)."""
    ref = get_valid_ref(ref)
    a, b = _c(flows_1, np.float32), _c(flows_2, np.float32)
    images = _c(images)
    if a.shape != b.shape or a.ndim != 4 or a.shape[3] != 2:
        raise ValueError("Error combining flows: Flow fields need to have the same shape (N,H,W,2)")
    if images.ndim == 3:
        images = images[..., None]
    n, h, w = a.shape[:3]
    if images.shape[:3] != (n, h, w):
        raise ValueError("Error applying flow: Flow shape does not match target shape")
    code = _ops.dtype_code(images.dtype)
    arith, rule = _ops.promoted_rule(images.dtype, False)
    am = None if masks_1 is None else _c(masks_1, np.bool_)
    bm = None if masks_2 is None else _c(masks_2, np.bool_)
    for name, arr, shape, dt in (('out_images', out_images, images.shape, images.dtype),
                                 ('out_valid', out_valid, (n, h, w), np.bool_), ('out', out, a.shape, np.float32),
                                 ('out_masks', out_masks, (n, h, w), np.bool_)):
        if arr is not None:
            _check_out(arr, shape, dt, name)
    out_images = np.empty_like(images) if out_images is None else out_images
    out_valid = np.empty((n, h, w), np.bool_) if out_valid is None else out_valid
    out = np.empty_like(a) if out is None else out
    out_masks = np.empty((n, h, w), np.bool_) if out_masks is None else out_masks
    flags = np.empty((n, 2), np.int32)
    device = dev.get_device() if device is None else device
    _lib.call('ofh_apply_combine3', images.ctypes.data, code, images.shape[3], arith, rule, a.ctypes.data,
              None if am is None else am.ctypes.data, b.ctypes.data, None if bm is None else bm.ctypes.data, ord(ref),
              DEFAULT_THRESHOLD if thresholded else 0.0, out_images.ctypes.data, out_valid.ctypes.data, out.ctypes.data,
              out_masks.ctypes.data, flags.ctypes.data, n, h, w, device)
    return out_images, out_valid, out, out_masks
