"""GPU parity tests for LSD + KeyLines (+ LBD) through the C ABI: segments bit-exact and in the same order as
cv2 4.13 / the C restatement."""
import ctypes
from pathlib import Path

import numpy as np
import pytest

import plslam_b200 as plf
from oracle import clib, synth
from oracle import frontend as ofe
from oracle.cvref import lsd_cv2

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "lines_v1.npz"


def cam(w, h):
    return dict(plf.KITTI_CAMERA, width=w, height=h)


def test_glibc_sincosf_port_matches_libm(fe):
    """The device port of glibc sinf/cosf used in region growing is bit-identical to the host libm."""
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = ctypes.c_float; libm.cosf.argtypes = [ctypes.c_float]
    libm.sinf.restype = ctypes.c_float; libm.sinf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 2 * np.pi, 150000), rng.uniform(0, 1e-3, 5000), rng.uniform(0.7, 0.9, 20000),
                        [0.0, 1e-5, 2.0 ** -12, np.pi / 4, np.pi / 2, np.pi, 6.2831855]]).astype(np.float32)
    s, c = fe.debug_sincosf(x)
    rs = np.array([libm.sinf(float(v)) for v in x], np.float32)
    rc = np.array([libm.cosf(float(v)) for v in x], np.float32)
    assert np.array_equal(s.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), rc.view(np.uint32))


def test_lsd_golden_cv2_segments(built):
    g = np.load(GOLD)
    h, w = g["left"].shape
    with plf.Frontend(camera=cam(w, h)) as fe:
        segs = fe.lsd(g["left"])
        assert segs.shape == g["segs"].shape and np.array_equal(segs, g["segs"])   # cv2 output, same order


@pytest.mark.parametrize("w,h,seed", [(1242, 375, 1), (752, 480, 9), (640, 360, 33)])
def test_lsd_vs_cv2_live(built, w, h, seed):
    pytest.importorskip("cv2")
    L, R = synth.scene_pair(w=w, h=h, seed=seed)
    with plf.Frontend(camera=cam(w, h)) as fe:
        for img in (L, R):
            ref = lsd_cv2(img)
            segs = fe.lsd(img)
            assert len(ref) > 100 and segs.shape == ref.shape and np.array_equal(segs, ref)


def test_lsd_scale_variants(built):
    L, _ = synth.scene_pair(w=500, h=300, seed=12, n_rect=80, n_lines=40)
    for sc in (0.8, 1.0):
        with plf.Frontend(camera=cam(500, 300), lsd_scale=sc) as fe:
            ref = clib.lsd(L, scale=sc)
            segs = fe.lsd(L)
            assert segs.shape == ref.shape and np.array_equal(segs, ref)


def test_lsd_edge_cases(built):
    with plf.Frontend(camera=cam(150, 100)) as fe:
        assert len(fe.lsd(np.full((100, 150), 90, np.uint8))) == 0            # flat: no gradient anywhere
        rng = np.random.default_rng(1)
        noise = rng.integers(0, 256, (100, 150), dtype=np.uint8)              # dense gradients, tiny regions
        assert np.array_equal(fe.lsd(noise), clib.lsd(noise))
        blob = np.zeros((100, 150), np.uint8); blob[20:80, 30:120] = 200       # long straight edges to the limits
        assert np.array_equal(fe.lsd(blob), clib.lsd(blob))


@pytest.mark.parametrize("nfeat", [300, 100, 0])
def test_detect_lines_vs_oracle(built, nfeat):
    """LSDDetectorC KeyLine stage + top-K + LBD on the C1 scene (KITTI size)."""
    L, R = synth.scene_pair()
    lim = plf.default_limits(); lim.max_lines = 2048
    with plf.Frontend(limits=lim, lsd_nfeatures=nfeat) as fe:
        for img in (L, R):
            kl, desc = fe.detect_lines(img)
            okl, odesc = ofe.detect_lines(img, lsd_nfeatures=nfeat)
            assert len(kl) == len(okl) and (nfeat == 0 or len(kl) == nfeat)
            assert kl.tobytes() == okl.tobytes()
            assert np.array_equal(desc, odesc)


def test_low_texture_lines_dominant(built):
    """BASELINE config 5 shape: nearly flat albedo, many long segments (> 400 lines, lsd_nfeatures = 0)."""
    L, _ = synth.scene_pair(seed=77, n_rect=12, n_lines=520)
    lim = plf.default_limits(); lim.max_lines = 4096
    with plf.Frontend(limits=lim, lsd_nfeatures=0) as fe:
        kl, desc = fe.detect_lines(L)
        okl, odesc = ofe.detect_lines(L, lsd_nfeatures=0)
        assert len(okl) > 400 and kl.tobytes() == okl.tobytes() and np.array_equal(desc, odesc)
