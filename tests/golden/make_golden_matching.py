"""Generates tests/golden/matching_v1.npz with cv2 4.13 BFMatcher (the reference's matcher library).
Run in the build container:  python tests/golden/make_golden_matching.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def knn_cv2(d1, d2):
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    mm = bf.knnMatch(d1, d2, k=2)
    out = np.full((len(d1), 4), -1, np.int32)
    for i, row in enumerate(mm):
        if len(row) > 0:
            out[i, 0] = row[0].trainIdx; out[i, 1] = int(row[0].distance)
        if len(row) > 1:
            out[i, 2] = row[1].trainIdx; out[i, 3] = int(row[1].distance)
    return out


def main():
    rng = np.random.default_rng(5)
    cases = {}
    # i.i.d. random 256-bit descriptors (SURVEY 8d micro-benchmark distribution)
    cases["rand"] = (rng.integers(0, 256, (700, 32), dtype=np.uint8),
                     rng.integers(0, 256, (650, 32), dtype=np.uint8))
    # tie-heavy: few distinct bit patterns, many equal distances
    base = rng.integers(0, 256, (12, 32), dtype=np.uint8)
    t1 = base[rng.integers(0, 12, 400)].copy(); t2 = base[rng.integers(0, 12, 380)].copy()
    t1[:, 0] ^= rng.integers(0, 4, 400).astype(np.uint8); t2[:, 1] ^= rng.integers(0, 4, 380).astype(np.uint8)
    cases["ties"] = (t1, t2)
    # correlated: d2 = noisy copy of a permutation of d1 (realistic stereo-like distances)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    perm = rng.permutation(500)
    noise = (rng.random((500, 256)) < 0.08)
    b = np.packbits(np.unpackbits(a[perm], axis=1) ^ noise, axis=1)
    cases["corr"] = (a, b)
    out = {}
    for k, (d1, d2) in cases.items():
        out[f"{k}_d1"] = d1; out[f"{k}_d2"] = d2
        out[f"{k}_knn12"] = knn_cv2(d1, d2); out[f"{k}_knn21"] = knn_cv2(d2, d1)
    np.savez_compressed(Path(__file__).parent / "matching_v1.npz", **out)
    print("wrote matching_v1.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
