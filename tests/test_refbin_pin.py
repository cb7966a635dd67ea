"""CPU tests: oracle/lbd.c pinned against the reference's OWN vendored line_descriptor code, compiled unmodified from
/root/reference into oracle/_ref/liblinedesc_ref.so (oracle/ref_build/Makefile; OpenCV primitives supplied by the
cv2-pinned restatements).  Skipped where neither the reference tree nor the prebuilt library is present."""
import ctypes as C

import numpy as np
import pytest

import plslam_b200 as plf
from oracle import clib, refbin
from plslam_b200 import synth

pytestmark = pytest.mark.skipif(not refbin.available(), reason="oracle/_ref/liblinedesc_ref.so not built (no /root/reference)")


def images():
    L, R = synth.scene_pair()
    E = synth.scene_pair(w=752, h=480, seed=9)[0]
    S = next(iter(synth.stream(plf.KITTI_CAMERA, 1, world=synth.World(seed=7), seed=42)))[1]
    return [L, R, E, S]


def test_lbd_descriptor_equals_vendored_binary_descriptor():
    """BinaryDescriptor::compute (binary_descriptor_custom.cpp:524-687, :1026-1372) == orc_lbd_compute, bit for bit."""
    total = 0
    for img in images():
        h, w = img.shape
        kl = clib.keylines_from_segments(clib.lsd(img), w, h, float(np.float32(0.025)) * min(w, h))
        kl["class_id"] = np.arange(len(kl), dtype=np.int32)
        assert len(kl) > 150
        assert np.array_equal(refbin.lbd(img, kl), clib.lbd_compute(img, kl))
        total += len(kl)
    assert total > 1500


def test_keyline_stage_equals_vendored_lsd_detector():
    """LSDDetectorC::detect (LSDDetector_custom.cpp:218-324) == orc_keylines_from_segments on every field except
    `angle`, where the reference build resolves atan2(float, float) to glibc's atan2f while the restatement (and the
    CUDA kernel) round the f64 atan2 to f32: a 1-ulp difference on ~10 % of the lines (deviations)."""
    libm = C.CDLL("libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    for img in images():
        h, w = img.shape
        minlen = float(np.float32(0.025)) * min(w, h)
        ref = refbin.keylines(img, min_length=minlen)
        mine = clib.keylines_from_segments(clib.lsd(img), w, h, minlen)
        assert len(ref) == len(mine) > 150
        for f in ref.dtype.names:
            if f != "angle":
                assert np.array_equal(ref[f], mine[f]), f
        dy = ref["endPointY"] - ref["startPointY"]; dx = ref["endPointX"] - ref["startPointX"]
        assert np.array_equal(ref["angle"], np.array([libm.atan2f(float(y), float(x)) for y, x in zip(dy, dx)], np.float32))
        assert np.array_equal(mine["angle"], np.arctan2(dy.astype(np.float64), dx.astype(np.float64)).astype(np.float32))
        ulp = np.abs(ref["angle"].view(np.int32).astype(np.int64) - mine["angle"].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1 and 0 < (ulp > 0).mean() < 0.3
