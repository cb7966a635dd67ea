"""ctypes wrapper of oracle/_build/liboracle.so (the C restatements).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "_build" / "liboracle.so"

KEYLINE_DTYPE = np.dtype([
    ("angle", np.float32), ("class_id", np.int32), ("octave", np.int32),
    ("ptx", np.float32), ("pty", np.float32), ("response", np.float32), ("size", np.float32),
    ("startPointX", np.float32), ("startPointY", np.float32), ("endPointX", np.float32),
    ("endPointY", np.float32), ("sPointInOctaveX", np.float32), ("sPointInOctaveY", np.float32),
    ("ePointInOctaveX", np.float32), ("ePointInOctaveY", np.float32), ("lineLength", np.float32),
    ("numOfPixels", np.int32)])
assert KEYLINE_DTYPE.itemsize == 68

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _LIB.exists():
            subprocess.run(["make", "oracle"], cwd=_DIR.parent, check=True, capture_output=True)
        _lib = C.CDLL(str(_LIB))
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gaussian_kernel_q8(ksize, sigma):
    taps = np.zeros(ksize, np.int32)
    lib().orc_gaussian_kernel_q8(int(ksize), C.c_double(sigma), _p(taps))
    return taps


def gaussian_blur_u8(img, ksize, sigma):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_gaussian_blur_u8(_p(img), img.shape[1], img.shape[0], int(ksize), C.c_double(sigma), _p(out))
    return out


def sobel3_i16(img):
    img = np.ascontiguousarray(img, np.uint8)
    dx = np.empty(img.shape, np.int16); dy = np.empty(img.shape, np.int16)
    lib().orc_sobel3_i16(_p(img), img.shape[1], img.shape[0], _p(dx), _p(dy))
    return dx, dy


def keylines_from_segments(segs, w, h, min_length):
    segs = np.ascontiguousarray(segs, np.float32).reshape(-1, 4)
    out = np.zeros(len(segs), KEYLINE_DTYPE)
    f = lib().orc_keylines_from_segments
    f.restype = C.c_int
    n = f(_p(segs), len(segs), int(w), int(h), C.c_double(min_length), _p(out))
    return out[:n].copy()


def lbd_compute(img, keylines, want_float=False):
    img = np.ascontiguousarray(img, np.uint8)
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    n = len(kl)
    desc = np.zeros((n, 32), np.uint8)
    fl = np.zeros((n, 72), np.float32) if want_float else None
    lib().orc_lbd_compute(_p(img), img.shape[1], img.shape[0], _p(kl), n, _p(desc),
                          _p(fl) if want_float else None)
    return (desc, fl) if want_float else desc
