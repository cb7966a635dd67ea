"""TEST INFRASTRUCTURE.  ctypes access to oracle/_ref/liblinedesc_ref.so: the reference's own vendored
3rdparty/line_descriptor (LSDDetectorC::detect, BinaryDescriptor::compute), compiled UNMODIFIED from /root/reference by
oracle/ref_build/Makefile against a stand-in for the OpenCV API whose primitives on this path are the cv2-pinned
restatements of oracle/*.c.  Used only to pin oracle/lbd.c (KeyLine stage + LBD) against the shipped reference code.
The library is built where /root/reference exists and travels to the GPU box; available() is False otherwise."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from oracle.clib import KEYLINE_DTYPE

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_ref" / "liblinedesc_ref.so"
_lib = None


def build():
    """Compiles the vendored sources if the reference tree is present (no-op otherwise)."""
    if Path("/root/reference/3rdparty/line_descriptor/src/binary_descriptor_custom.cpp").exists():
        subprocess.run(["make", "-C", str(_DIR / "ref_build")], check=True, capture_output=True)


def available():
    if not _SO.exists():
        try:
            build()
        except Exception:
            return False
    if not _SO.exists():
        return False
    try:
        lib()
    except OSError:
        return False
    return True


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_SO))
        _lib.ref_keylines.restype = C.c_int
        _lib.ref_lbd.restype = C.c_int
    return _lib


def keylines(img, scale_arg=1, num_octaves=1, refine=0, scale=1.2, sigma_scale=0.6, quant=2.0, ang_th=22.5, log_eps=1.0,
             density_th=0.6, n_bins=1024, min_length=0.0, cap=65536):
    """LSDDetectorC::detect(image, keylines, scale_arg, num_octaves, opts) of the vendored code."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(cap, KEYLINE_DTYPE)
    n = lib().ref_keylines(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], int(scale_arg), int(num_octaves), int(refine),
                           C.c_double(scale), C.c_double(sigma_scale), C.c_double(quant), C.c_double(ang_th), C.c_double(log_eps),
                           C.c_double(density_th), int(n_bins), C.c_double(min_length), out.ctypes.data_as(C.c_void_p), cap)
    if n < 0:
        raise RuntimeError(f"ref_keylines failed ({n})")
    return out[:n].copy()


def lbd(img, kls):
    """BinaryDescriptor::compute(image, keylines, descriptors) of the vendored code: uint8 [n, 32]."""
    img = np.ascontiguousarray(img, np.uint8)
    kl = np.ascontiguousarray(kls, KEYLINE_DTYPE)
    desc = np.zeros((len(kl), 32), np.uint8)
    n = lib().ref_lbd(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], kl.ctypes.data_as(C.c_void_p), len(kl),
                      desc.ctypes.data_as(C.c_void_p))
    if n != len(kl):
        raise RuntimeError(f"ref_lbd failed ({n})")
    return desc
