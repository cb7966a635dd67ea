"""CPU tests: the C restatement of OpenCV's LSD (refine = 0) pinned bit-exact — same segments, same order —
against cv2 4.13 (live) and the committed golden segments (tests/golden/lines_v1.npz was produced by cv2)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import clib, synth
from oracle.cvref import lsd_cv2

GOLD = Path(__file__).parent / "golden" / "lines_v1.npz"


def test_lsd_oracle_matches_golden_cv2_segments():
    g = np.load(GOLD)
    for trig in (0, 1):
        segs = clib.lsd(g["left"], trig_mode=trig)
        assert segs.shape == g["segs"].shape and np.array_equal(segs, g["segs"])


@pytest.mark.parametrize("w,h,seed", [(1242, 375, 1), (752, 480, 9), (640, 360, 33)])
def test_lsd_oracle_matches_cv2_live(w, h, seed):
    pytest.importorskip("cv2")
    L, R = synth.scene_pair(w=w, h=h, seed=seed)
    for img in (L, R):
        ref = lsd_cv2(img)
        mine = clib.lsd(img)
        assert len(ref) > 100 and mine.shape == ref.shape and np.array_equal(mine, ref)


def test_lsd_scale_variants_vs_cv2():
    """scale 0.8 (OpenCV default; sigma = 0.6/0.8) and scale 1.0 (no resample) also match."""
    pytest.importorskip("cv2")
    L, _ = synth.scene_pair(w=500, h=300, seed=12, n_rect=80, n_lines=40)
    for sc in (0.8, 1.0):
        ref = lsd_cv2(L, scale=sc)
        mine = clib.lsd(L, scale=sc)
        assert mine.shape == ref.shape and np.array_equal(mine, ref)


def test_lsd_flat_image_has_no_segments():
    assert len(clib.lsd(np.full((100, 150), 90, np.uint8))) == 0
