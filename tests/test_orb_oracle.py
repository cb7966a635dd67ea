"""CPU tests: the C restatement of OpenCV's ORB pinned bit-exact against cv2 4.13 (live) and the golden."""
from pathlib import Path

import numpy as np
import pytest

from oracle import clib, synth
from oracle.cvref import orb_cv2

GOLD = Path(__file__).parent / "golden" / "orb_v1.npz"
FIELDS = ["x", "y", "size", "angle", "response", "octave"]


def same_kps(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in FIELDS)


@pytest.mark.parametrize("name,nf", [("l", 800), ("r", 800), ("l", 300), ("r", 300)])
def test_oracle_matches_golden(name, nf):
    g = np.load(GOLD)
    img = g["left"] if name == "l" else g["right"]
    kp, desc = clib.orb(img, nf)
    assert same_kps(kp, g[f"kp_{name}_{nf}"])
    assert np.array_equal(desc, g[f"desc_{name}_{nf}"])


def test_oracle_matches_cv2_live_full_size():
    pytest.importorskip("cv2")
    L, R = synth.scene_pair()
    for img in (L, R):
        kp, desc = clib.orb(img, 1500)
        rk, rd = orb_cv2(img, nfeatures=1500)
        assert same_kps(kp, rk) and np.array_equal(desc, rd)


def test_resize_exact_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (sw, sh, dw, dh) in [(1242, 375, 1035, 312), (1035, 312, 862, 260), (752, 480, 627, 400), (53, 37, 64, 44)]:
        img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        assert np.array_equal(clib.resize_linear_exact(img, dw, dh),
                              cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT))


def test_fast_atan2_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for _ in range(2000):
        y, x = (float(v) for v in rng.integers(-200000, 200000, 2))
        assert np.float32(clib.fast_atan2(y, x)) == np.float32(cv2.fastAtan2(y, x))
    assert clib.fast_atan2(0.0, 0.0) == cv2.fastAtan2(0.0, 0.0)
