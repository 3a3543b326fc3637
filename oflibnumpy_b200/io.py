"""Dataset loaders (host file I/O; formats of utils.py:426-490 of the reference). Not on the compute path: the arrays
they return are uploaded by the Flow constructor."""
import numpy as np


def read_kitti_raw(path):
    """The uint16 (H,W,3) BGR image of a KITTI flow PNG exactly as OpenCV decodes it (input of ofk_decode_kitti)."""
    import cv2
    inp = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if inp is None:
        raise ValueError("Error loading flow from KITTI data: Flow data could not be loaded")
    if inp.ndim != 3 or inp.shape[-1] != 3:
        raise ValueError("Error loading flow from KITTI data: Loaded flow data has the wrong shape")
    if inp.dtype != np.uint16:
        inp = inp.astype(np.uint16)        # 8-bit PNGs: the reference converts whatever it gets to float64
    return np.ascontiguousarray(inp)


def read_sintel_invalid_raw(path):
    """The uint8 (H,W) invalid-pixel image of Sintel as OpenCV decodes it (input of ofk_decode_sintel_mask)."""
    if not isinstance(path, str):
        raise TypeError("Error loading flow from Sintel data: Path needs to be a string")
    import cv2
    mask = cv2.imread(path, 0)
    if mask is None:
        raise ValueError("Error loading flow from Sintel data: Invalid mask could not be loaded from path")
    return np.ascontiguousarray(mask)


def load_kitti(path):
    """KITTI uint16 PNG: channels (u, v, valid) with u, v stored as (value - 2**15) / 64. Returns float64 (H,W,3)."""
    import cv2  # only needed for PNG decoding
    inp = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if inp is None:
        raise ValueError("Error loading flow from KITTI data: Flow data could not be loaded")
    if inp.ndim != 3 or inp.shape[-1] != 3:
        raise ValueError("Error loading flow from KITTI data: Loaded flow data has the wrong shape")
    out = inp[..., ::-1].astype('float64')  # OpenCV decodes as BGR
    out[..., :2] = (out[..., :2] - 2 ** 15) / 64
    return out


def load_sintel(path):
    """Sintel .flo: b'PIEH', int32 width, int32 height, then H*W*2 little-endian float32."""
    if not isinstance(path, str):
        raise TypeError("Error loading flow from Sintel data: Path needs to be a string")
    with open(path, 'rb') as fh:
        if fh.read(4) != b'PIEH':
            raise ValueError("Error loading flow from Sintel data: Path not a valid .flo file")
        w = int.from_bytes(fh.read(4), 'little')
        h = int.from_bytes(fh.read(4), 'little')
        flow = np.frombuffer(fh.read(), dtype='<f4').reshape(h, w, 2)
    return flow


def load_sintel_mask(path):
    """Sintel invalid-pixel PNG -> boolean mask that is True on valid pixels."""
    if not isinstance(path, str):
        raise TypeError("Error loading flow from Sintel data: Path needs to be a string")
    import cv2
    mask = cv2.imread(path, 0)
    if mask is None:
        raise ValueError("Error loading flow from Sintel data: Invalid mask could not be loaded from path")
    return ~(mask.astype('bool'))
