"""Argument validators of the public API. Types of exception and message texts are the reference's
(utils.py:25-88 of oflibnumpy) because they are part of the drop-in contract; the checks run on the host before any
device call."""
import numpy as np

DEFAULT_THRESHOLD = 1e-3  # utils.py:22


def get_valid_ref(ref):
    if ref is None:
        return 't'
    if not isinstance(ref, str):
        raise TypeError("Error setting flow reference: Input is not a string")
    if ref not in ('s', 't'):
        raise ValueError("Error setting flow reference: Input is not 's' or 't', but {}".format(ref))
    return ref


def get_valid_padding(padding, error_string=None):
    prefix = error_string or ''
    if not isinstance(padding, (list, tuple)):
        raise TypeError(prefix + "Padding needs to be a tuple or a list list of values [top, bot, left, right]")
    if len(padding) != 4:
        raise ValueError(prefix + "Padding list needs to be a list or tuple of length 4 [top, bot, left, right]")
    if any(not isinstance(p, int) for p in padding):
        raise ValueError(prefix + "Padding list [top, bot, left, right] items need to be integers")
    if any(p < 0 for p in padding):
        raise ValueError(prefix + "Padding list [top, bot, left, right] items need to be 0 or larger")
    return padding


def validate_shape(shape):
    if not isinstance(shape, (list, tuple)):
        raise TypeError("Error creating flow from matrix: Dims need to be a list or a tuple")
    if len(shape) != 2:
        raise ValueError("Error creating flow from matrix: Dims need to be a list or a tuple of length 2")
    if any((item <= 0 or not isinstance(item, int)) for item in shape):
        raise ValueError("Error creating flow from matrix: Dims need to be a list or a tuple of integers above zero")


def validate_flow_array(flow, error_string=None, finite_on_device=False):
    """Host-side shape/type checks of a flow ndarray; returns it as C-contiguous float32. The NaN / Inf test
    (utils.py:55-56) runs here unless `finite_on_device` is set and the array is float32 already: the caller then
    uploads first and tests on the device (upload_flow_array) -- np.isfinite costs more than the upload at megapixel
    sizes. Other dtypes are always tested before the float32 cast, as the reference does."""
    prefix = error_string or ''
    if not isinstance(flow, np.ndarray):
        raise TypeError(prefix + "Flow is not a numpy array")
    if flow.ndim != 3:
        raise ValueError(prefix + "Flow array is not 3-dimensional")
    if flow.shape[2] != 2:
        raise ValueError(prefix + "Flow array does not have 2 channels")
    if not (finite_on_device and flow.dtype == np.float32) and not np.isfinite(flow).all():
        raise ValueError(prefix + "Flow array contains NaN or Inf values")
    return np.ascontiguousarray(flow, dtype=np.float32)


def validate_transform_list(transform_list):
    """Checks of from_transforms (utils.py:363-388)."""
    pre = "Error creating flow from transforms: "
    if not isinstance(transform_list, list):
        raise TypeError(pre + "Transform_list needs to be a list")
    if not all(isinstance(item, list) for item in transform_list):
        raise TypeError(pre + "Transform_list needs to be a list of lists")
    if not all(len(item) > 1 for item in transform_list):
        raise ValueError(pre + "Invalid transforms passed")
    expected = {'translation': 2, 'rotation': 3, 'scaling': 3}
    for t in transform_list:
        if t[0] not in expected:
            raise ValueError(pre + "Transform '{}' not recognised".format(t[0]))
        if len(t) - 1 != expected[t[0]]:
            raise ValueError(pre + "Not enough transform values passed for '{}' - expected {}, got {}"
                             .format(t[0], expected[t[0]], len(t) - 1))
        if not all(isinstance(item, (float, int)) for item in t[1:]):
            raise ValueError(pre + "Transform values for '{}' need to be integers or floats".format(t[0]))
