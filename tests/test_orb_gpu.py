"""GPU parity tests for ORB through the C ABI: keypoints and descriptors bit-exact vs cv2 4.13 / the golden."""
from pathlib import Path

import numpy as np
import pytest

import plslam_b200 as plf
from oracle import clib, synth
from oracle.cvref import orb_cv2

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "orb_v1.npz"
FIELDS = ["x", "y", "size", "angle", "response", "octave"]


def assert_same(kp, desc, rk, rd):
    assert len(kp) == len(rk)
    for f in FIELDS:
        assert np.array_equal(kp[f], rk[f]), f
    assert np.array_equal(desc, rd)


@pytest.mark.parametrize("nf", [800, 300])
def test_orb_golden_bit_exact(built, nf):
    g = np.load(GOLD)
    h, w = g["left"].shape
    with plf.Frontend(camera=dict(plf.KITTI_CAMERA, width=w, height=h), orb_nfeatures=nf) as fe:
        for name in ("l", "r"):
            img = g["left"] if name == "l" else g["right"]
            kp, desc = fe.orb(img)
            assert_same(kp, desc, g[f"kp_{name}_{nf}"], g[f"desc_{name}_{nf}"])


@pytest.mark.parametrize("shape,seed,nf", [((375, 1242), 1, 1500), ((480, 752), 9, 1200), ((375, 1242), 2, 800)])
def test_orb_vs_cv2_live(built, shape, seed, nf):
    pytest.importorskip("cv2")
    h, w = shape
    L, R = synth.scene_pair(w=w, h=h, seed=seed, n_rect=260)
    with plf.Frontend(camera=dict(plf.KITTI_CAMERA, width=w, height=h), orb_nfeatures=nf) as fe:
        for img in (L, R):
            kp, desc = fe.orb(img)
            rk, rd = orb_cv2(img, nfeatures=nf)
            assert len(rk) > nf * 0.5
            assert_same(kp, desc, rk, rd)


def test_orb_edge_cases(built):
    h, w = 120, 160
    with plf.Frontend(camera=dict(plf.KITTI_CAMERA, width=w, height=h), orb_nfeatures=100, orb_nlevels=2) as fe:
        flat = np.full((h, w), 128, np.uint8)
        kp, desc = fe.orb(flat)                         # no corners at all
        assert len(kp) == 0 and desc.shape == (0, 32)
        rng = np.random.default_rng(0)
        noise = rng.integers(0, 256, (h, w), dtype=np.uint8)   # corners everywhere: ties at the retainBest boundary
        kp, desc = fe.orb(noise)
        ok, od = clib.orb(noise, 100, nlevels=2)
        assert_same(kp, desc, ok, od)
    # single level (config.yaml:118 / config_fast.yaml:62 use orb_nlevels 1)
    L, _ = synth.scene_pair(w=320, h=200, seed=5, n_rect=60)
    with plf.Frontend(camera=dict(plf.KITTI_CAMERA, width=320, height=200), orb_nfeatures=600, orb_nlevels=1) as fe:
        kp, desc = fe.orb(L)
        rk, rd = orb_cv2(L, nfeatures=600, nlevels=1)
        assert_same(kp, desc, rk, rd)


def test_orb_capacity_error(built):
    lim = plf.default_limits()
    lim.max_keypoints = 64
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (200, 300), dtype=np.uint8)
    with plf.Frontend(camera=dict(plf.KITTI_CAMERA, width=300, height=200), limits=lim, orb_nfeatures=500) as fe:
        with pytest.raises(plf.PlfError, match="capacity"):
            fe.orb(noise)
