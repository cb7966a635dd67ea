"""Generates tests/golden/orb_v1.npz: cv2 4.13 ORB (reference parameters, config_euroc.yaml:59-67) keypoints and
descriptors of a small seeded scene, canonically ordered (octave, y, x).
Run in the build container: python tests/golden/make_golden_orb.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import synth  # noqa: E402
from oracle.cvref import orb_cv2  # noqa: E402


def main():
    L, R = synth.scene_pair(w=480, h=300, seed=4, n_rect=90, n_lines=40, disparity=14)
    out = dict(left=L, right=R)
    for name, img in (("l", L), ("r", R)):
        for nf in (800, 300):
            kp, desc = orb_cv2(img, nfeatures=nf)
            out[f"kp_{name}_{nf}"] = kp
            out[f"desc_{name}_{nf}"] = desc
    np.savez_compressed(Path(__file__).parent / "orb_v1.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
