"""Device memory, streams and events for the Python shim -- thin wrappers over the ofk_rt_* entry points.

:class:`DeviceArray` is the only device container: a typed, C-contiguous block from the library's stream-ordered pool.
It implements ``__cuda_array_interface__`` (so ``torch.as_tensor(arr, device='cuda')`` is zero-copy) and
:func:`as_device` wraps any object exposing that interface (torch / cupy tensors) without copying.
"""
import atexit
import ctypes as C

import numpy as np

from . import _lib

_current_stream = None  # cudaStream_t as int, None = legacy default stream
_closing = False        # set at interpreter exit: objects die in arbitrary order then, and the process teardown frees
                        # every device resource anyway, so finalisers stop calling into the library


def _mark_closing():
    global _closing
    _closing = True


atexit.register(_mark_closing)


def device_count():
    n = C.c_int(0)
    try:
        _lib.call('ofk_rt_device_count', C.byref(n))
    except _lib.OflibCudaError:
        return 0
    return n.value


def set_device(index):
    _lib.call('ofk_rt_set_device', int(index))


def get_device():
    d = C.c_int(0)
    _lib.call('ofk_rt_get_device', C.byref(d))
    return d.value


def device_info(index=None):
    index = get_device() if index is None else index
    sm, mj, mn = C.c_int(), C.c_int(), C.c_int()
    l2, mem = C.c_size_t(), C.c_size_t()
    _lib.call('ofk_rt_device_info', index, C.byref(sm), C.byref(mj), C.byref(mn), C.byref(l2), C.byref(mem))
    return {'sm_count': sm.value, 'cc': (mj.value, mn.value), 'l2_bytes': l2.value, 'total_mem': mem.value}


def require_gpu():
    if device_count() < 1:
        raise _lib.OflibCudaError("oflibnumpy_b200 needs a CUDA device (B200, sm_100a); none is visible and there is "
                                  "no CPU fallback: " + _lib.last_error())


def set_stream(stream):
    """Stream used by every subsequent call (int / object with .cuda_stream / None for the default stream).

    One stream is current per process. Device buffers are returned to the stream-ordered pool on whatever stream is
    current when they are collected, so a switch orders the new stream after everything queued on the old one (an
    event on the old stream, a wait on the new one): a buffer whose kernels are still running on the old stream
    cannot be handed to a new allocation on the new stream before they finish."""
    global _current_stream
    if stream is None:
        new = None
    elif hasattr(stream, 'cuda_stream'):
        new = int(stream.cuda_stream)
    elif isinstance(stream, Stream):
        new = stream.handle
    else:
        new = int(stream)
    if new != _current_stream and device_count() > 0:
        ev = C.c_void_p()
        _lib.call('ofk_rt_event_create', C.byref(ev))
        try:
            _lib.call('ofk_rt_event_record', ev.value, _current_stream)
            _lib.call('ofk_rt_stream_wait_event', new, ev.value)
        finally:
            _lib.call('ofk_rt_event_destroy', ev.value)
    _current_stream = new


def current_stream():
    return _current_stream


def synchronize():
    _lib.call('ofk_rt_device_sync')


class Stream:
    def __init__(self):
        h = C.c_void_p()
        _lib.call('ofk_rt_stream_create', C.byref(h))
        self.handle = h.value

    def synchronize(self):
        _lib.call('ofk_rt_stream_sync', self.handle)

    def __del__(self):
        global _current_stream
        try:
            if self.handle and not _closing:
                if _current_stream == self.handle:      # later frees must not be ordered on a destroyed stream
                    _current_stream = None
                _lib.call('ofk_rt_stream_destroy', self.handle)
        except Exception:
            pass


class Event:
    def __init__(self):
        h = C.c_void_p()
        _lib.call('ofk_rt_event_create', C.byref(h))
        self.handle = h.value

    def record(self, stream=None):
        s = _current_stream if stream is None else (stream.handle if isinstance(stream, Stream) else stream)
        _lib.call('ofk_rt_event_record', self.handle, s)

    def synchronize(self):
        _lib.call('ofk_rt_event_sync', self.handle)

    def elapsed_ms(self, end):
        ms = C.c_float()
        _lib.call('ofk_rt_event_elapsed_ms', self.handle, end.handle, C.byref(ms))
        return ms.value

    def __del__(self):
        try:
            if self.handle and not _closing:
                _lib.call('ofk_rt_event_destroy', self.handle)
        except Exception:
            pass


class DeviceArray:
    """C-contiguous typed device buffer."""
    __slots__ = ('ptr', 'shape', 'dtype', '_owner', '_owned')

    def __init__(self, ptr, shape, dtype, owner=None, owned=False):
        self.ptr = ptr
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self._owner = owner   # keeps the exporting object (torch tensor, parent array) alive
        self._owned = owned

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def ndim(self):
        return len(self.shape)

    @classmethod
    def empty(cls, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        p = C.c_void_p()
        _lib.call('ofk_rt_malloc', C.byref(p), nbytes, _current_stream)
        return cls(p.value, shape, dtype, owned=True)

    @classmethod
    def zeros(cls, shape, dtype):
        a = cls.empty(shape, dtype)
        _lib.call('ofk_rt_memset', a.ptr, 0, a.nbytes, _current_stream)
        return a

    @classmethod
    def from_numpy(cls, arr, dtype=None):
        h = np.ascontiguousarray(arr, dtype=dtype)
        a = cls.empty(h.shape, h.dtype)
        _lib.call('ofk_rt_memcpy_h2d', a.ptr, h.ctypes.data, h.nbytes, _current_stream)
        if _current_stream is not None:  # pageable source: make sure the copy has consumed it before h can die
            _lib.call('ofk_rt_stream_sync', _current_stream)
        return a

    def numpy(self, out=None):
        """Copy to the host (synchronous)."""
        if out is None:
            out = np.empty(self.shape, self.dtype)
        _lib.call('ofk_rt_memcpy_d2h', out.ctypes.data, self.ptr, self.nbytes, _current_stream)
        _lib.call('ofk_rt_stream_sync', _current_stream)
        return out

    def copy(self):
        a = DeviceArray.empty(self.shape, self.dtype)
        _lib.call('ofk_rt_memcpy_d2d', a.ptr, self.ptr, self.nbytes, _current_stream)
        return a

    def copy_from(self, other):
        """Device-to-device copy of `other` (same size in bytes) into this array."""
        if other.nbytes != self.nbytes:
            raise ValueError("copy_from: size mismatch ({} vs {} bytes)".format(other.nbytes, self.nbytes))
        _lib.call('ofk_rt_memcpy_d2d', self.ptr, other.ptr, self.nbytes, _current_stream)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError("cannot reshape device array of size {} into {}".format(self.size, shape))
        return DeviceArray(self.ptr, shape, self.dtype, owner=self, owned=False)

    def frames(self, start, stop):
        """View of frames [start, stop) along axis 0."""
        per = self.nbytes // self.shape[0] if self.shape[0] else 0
        return DeviceArray(self.ptr + start * per, (stop - start,) + self.shape[1:], self.dtype, owner=self)

    @property
    def __cuda_array_interface__(self):
        return {'shape': self.shape, 'typestr': self.dtype.str, 'data': (self.ptr, False), 'version': 3,
                'strides': None}

    def __del__(self):
        if self._owned and self.ptr and not _closing:
            try:
                _lib.call('ofk_rt_free', self.ptr, _current_stream)
            except Exception:
                pass
            self.ptr = None


def as_device(obj, dtype=None):
    """Wrap a DeviceArray / any ``__cuda_array_interface__`` exporter (torch, cupy) without copying."""
    if isinstance(obj, DeviceArray):
        if dtype is not None and obj.dtype != np.dtype(dtype):
            raise TypeError("device array has dtype {}, expected {}".format(obj.dtype, np.dtype(dtype)))
        return obj
    cai = getattr(obj, '__cuda_array_interface__', None)
    if cai is None:
        raise TypeError("object does not expose __cuda_array_interface__")
    if cai.get('strides') is not None:
        item = np.dtype(cai['typestr']).itemsize
        expect, acc = [], item
        for s in reversed(cai['shape']):
            expect.append(acc)
            acc *= s
        if tuple(cai['strides']) != tuple(reversed(expect)):
            raise ValueError("device array must be C-contiguous")
    dt = np.dtype(cai['typestr'])
    if dtype is not None and dt != np.dtype(dtype):
        raise TypeError("device array has dtype {}, expected {}".format(dt, np.dtype(dtype)))
    return DeviceArray(cai['data'][0], cai['shape'], dt, owner=obj)


def is_device_array(obj):
    return isinstance(obj, DeviceArray) or hasattr(obj, '__cuda_array_interface__')


class PinnedArray:
    """numpy array over page-locked host memory (for asynchronous, full-rate host<->device copies)."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(shape, dtype=np.int64)) * dtype.itemsize, 1)
        p = C.c_void_p()
        _lib.call('ofk_rt_host_alloc', C.byref(p), nbytes)
        self._ptr = p.value
        buf = (C.c_byte * nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)

    def __del__(self):
        try:
            if self._ptr and not _closing:
                self.array = None
                _lib.call('ofk_rt_host_free', self._ptr)
                self._ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """Returns (numpy array, keep-alive handle). The array is valid while the handle lives."""
    p = PinnedArray(shape, dtype)
    return p.array, p
