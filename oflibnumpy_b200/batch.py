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

__all__ = ['FlowBatch', 'shard_range', 'apply_flow_host', 'combine_flows_host']


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
    def apply(self, targets, target_masks=None, return_valid_area=False):
        """Warp images (N,H,W,C) [numpy or device] frame by frame with this batch (ref 't'), like ``Flow.apply``.
        Returns device arrays (``.numpy()`` to download): images, and the valid areas if requested."""
        if self.ref != 't':
            raise NotImplementedError("FlowBatch.apply: batched warping is built for ref 't' flows")
        payload = _to_device(targets)
        if payload.ndim == 3:
            payload = payload.reshape(payload.shape + (1,))
        if payload.shape[:3] != self.vecs.shape[:3]:
            raise ValueError("Error applying flow: Flow shape does not match target shape")
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

    def apply_to_flows(self, other):
        """``self[i].apply(other[i])`` for every frame: warp a batch of flows (vectors and masks) with this batch."""
        if self.ref != 't':
            raise NotImplementedError("FlowBatch.apply_to_flows: built for ref 't' flows")
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
        if mode != 3:
            raise NotImplementedError("FlowBatch.combine_with: batched combination is built for mode 3")
        thr = DEFAULT_THRESHOLD if thresholded else 0.0
        v, m, flags = _ops.combine3(self.vecs, self.masks, other.vecs, other.masks, self.ref, thr)
        res = FlowBatch._wrap(v, self.ref, m)
        return (res, flags) if return_flags else res

    def valid_target(self):
        if self.ref != 't':
            raise NotImplementedError("FlowBatch.valid_target: built for ref 't' flows")
        return _ops.valid_geom_t(self.vecs, -1.0, self.masks)

    def valid_source(self):
        if self.ref != 's':
            raise NotImplementedError("FlowBatch.valid_source: built for ref 's' flows")
        return _ops.valid_geom_t(self.vecs, 1.0, self.masks)

    def invert(self, ref):
        """Cross-reference inversion (negate + relabel, flow_class.py:748,751)."""
        ref = get_valid_ref(ref)
        if ref == self.ref:
            raise NotImplementedError("FlowBatch.invert: same-reference inversion is per-frame (Flow.invert)")
        return FlowBatch._wrap(_ops.scale(_lib.OP_MUL, self.vecs, -1.0, -1.0, False), ref, self.masks)


# ---------------------------------------------------------------------------------------------------- host-buffer calls
def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


def apply_flow_host(flows, images, flow_masks=None, target_masks=None, return_valid_area=False, out=None,
                    out_valid=None, device=None):
    """Batched ``Flow(flows[i], 't', flow_masks[i]).apply(images[i], target_masks[i], return_valid_area)`` on HOST
    arrays through ofh_warp_t: frames stream through a device ring, host<->device copies overlap the kernels.
    Pass pinned arrays (device.pinned_empty) for full-rate copies. Returns numpy arrays."""
    flows = _c(flows, np.float32)
    images = _c(images)
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
    out = np.empty_like(images) if out is None else out
    pm = fm = om = None
    if return_valid_area:
        out_valid = np.empty((n, h, w), np.bool_) if out_valid is None else out_valid
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
    """Batched ``combine_flows(flows_1[i], flows_2[i], 3, ref)`` / ``Flow.combine_with`` on HOST arrays (ofh_combine3).
    Returns (vecs (N,H,W,2) float32, masks (N,H,W) bool)."""
    if mode != 3:
        raise NotImplementedError("combine_flows_host: built for mode 3")
    ref = get_valid_ref(ref)
    a, b = _c(flows_1, np.float32), _c(flows_2, np.float32)
    if a.shape != b.shape or a.ndim != 4 or a.shape[3] != 2:
        raise ValueError("Error combining flows: Flow fields need to have the same shape (N,H,W,2)")
    n, h, w = a.shape[:3]
    am = None if masks_1 is None else _c(masks_1, np.bool_)
    bm = None if masks_2 is None else _c(masks_2, np.bool_)
    out = np.empty_like(a) if out is None else out
    out_masks = np.empty((n, h, w), np.bool_) if out_masks is None else out_masks
    flags = np.empty((n, 2), np.int32)
    device = dev.get_device() if device is None else device
    _lib.call('ofh_combine3', a.ctypes.data, None if am is None else am.ctypes.data, b.ctypes.data,
              None if bm is None else bm.ctypes.data, ord(ref), DEFAULT_THRESHOLD if thresholded else 0.0,
              out.ctypes.data, out_masks.ctypes.data, flags.ctypes.data, n, h, w, device)
    return out, out_masks
