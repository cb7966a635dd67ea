"""Generates tests/golden/lines_v1.npz: a small seeded scene, its cv2 4.13 LSD segments (reference
parameters, config/config/config_euroc.yaml:68-77), cv2 blur/Sobel outputs, and the KeyLines + LBD descriptors
produced by the REFERENCE'S OWN vendored line_descriptor code compiled unmodified (oracle/_ref, LSDDetectorC::detect +
BinaryDescriptor::compute) - checked here to equal the oracle's restatement bit for bit before they are written.
Run in the build container (needs /root/reference): python tests/golden/make_golden_lines.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "pl-slam_b200"))
from oracle import clib, refbin, synth  # noqa: E402


def main():
    L, R = synth.scene_pair(w=400, h=240, seed=3, n_rect=50, n_lines=30, disparity=12)
    lsd = cv2.createLineSegmentDetector(0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024)
    segs = lsd.detect(L)[0].reshape(-1, 4).astype(np.float32)
    blur = cv2.GaussianBlur(L, (5, 5), 1)
    dx = cv2.Sobel(blur, cv2.CV_16S, 1, 0, ksize=3)
    dy = cv2.Sobel(blur, cv2.CV_16S, 0, 1, ksize=3)
    kl = clib.keylines_from_segments(segs, 400, 240, 0.025 * 240)
    desc, fl = clib.lbd_compute(L, kl, want_float=True)
    ref_kl = refbin.keylines(L, min_length=0.025 * 240)          # the reference's compiled LSDDetector_custom.cpp
    assert ref_kl.tobytes() == kl.tobytes(), "oracle KeyLines differ from the reference binary"
    assert np.array_equal(refbin.lbd(L, ref_kl), desc), "oracle LBD differs from the reference binary"
    np.savez_compressed(Path(__file__).parent / "lines_v1.npz", left=L, right=R, segs=segs, blur=blur,
                        dx=dx, dy=dy, keylines=kl, lbd=desc, lbd_float=fl)
    print("segments", len(segs), "keylines", len(kl))


if __name__ == "__main__":
    main()
