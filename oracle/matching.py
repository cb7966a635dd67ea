"""Oracle: Hamming kNN(2) + NNR + mutual-consistency matcher (numpy; test infrastructure only).

Follows
  * cv::BFMatcher(NORM_HAMMING, crossCheck=false).knnMatch(k=2) — tie rule = (distance, trainIdx)
    lexicographic (verified against cv2 4.13 in tests/test_matching_oracle.py);
  * popcount distance: 3rdparty/line_descriptor/src/bitops_custom.hpp:83-96;
  * stvo-pl match()/matchNNR() as used at src/mapHandler.cpp:277,424,597,712,3223,3249
    (stvo-pl itself is not vendored => unpinned): accept row i when
    best.distance < second.distance * nnr evaluated in float32, then, if best_lr_matches,
    keep only mutual ratio-accepted pairs.
"""
import numpy as np


def hamming_matrix(d1: np.ndarray, d2: np.ndarray) -> np.ndarray:
    d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
    d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
    a = d1.view(np.uint64)[:, None, :]
    b = d2.view(np.uint64)[None, :, :]
    return np.bitwise_count(a ^ b).sum(axis=2).astype(np.int32)


def hamming_knn2(d1, d2):
    """Returns idx1, dist1, idx2, dist2 (int32[n1]); -1 where fewer than 1/2 train rows exist."""
    D = hamming_matrix(d1, d2)
    n1, n2 = D.shape
    idx1 = np.full(n1, -1, np.int32); dist1 = np.full(n1, -1, np.int32)
    idx2 = np.full(n1, -1, np.int32); dist2 = np.full(n1, -1, np.int32)
    if n1 == 0 or n2 == 0:
        return idx1, dist1, idx2, dist2
    key = D.astype(np.int64) * 65536 + np.arange(n2, dtype=np.int64)[None, :]
    if n2 >= 2:
        part = np.sort(key, axis=1)[:, :2]
        idx2[:] = (part[:, 1] % 65536).astype(np.int32)
        dist2[:] = (part[:, 1] // 65536).astype(np.int32)
    else:
        part = key
    idx1[:] = (part[:, 0] % 65536).astype(np.int32)
    dist1[:] = (part[:, 0] // 65536).astype(np.int32)
    return idx1, dist1, idx2, dist2


def match_nnr(d1, d2, nnr):
    idx1, dist1, idx2, dist2 = hamming_knn2(d1, d2)
    nnr = np.float32(nnr)
    ok = (idx1 >= 0) & (idx2 >= 0) & (dist1.astype(np.float32) < dist2.astype(np.float32) * nnr)
    return np.where(ok, idx1, -1).astype(np.int32)


def match(d1, d2, nnr, best_lr=True):
    """stvo-pl match(): returns (matches_12, count)."""
    m12 = match_nnr(d1, d2, nnr)
    if best_lr:
        m21 = match_nnr(d2, d1, nnr)
        for i in range(len(m12)):
            j = m12[i]
            if j >= 0 and m21[j] != i:
                m12[i] = -1
    return m12, int((m12 >= 0).sum())
