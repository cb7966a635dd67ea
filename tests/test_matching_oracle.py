"""CPU tests: the numpy matcher oracle against cv2.BFMatcher (the library the reference's matcher
calls) and against the committed golden vectors."""
from pathlib import Path

import numpy as np
import pytest

from oracle import matching as om

GOLD = Path(__file__).parent / "golden" / "matching_v1.npz"


@pytest.mark.parametrize("case", ["rand", "ties", "corr"])
def test_oracle_knn_matches_golden(case):
    g = np.load(GOLD)
    d1, d2 = g[f"{case}_d1"], g[f"{case}_d2"]
    for a, b, key in ((d1, d2, "knn12"), (d2, d1, "knn21")):
        i1, s1, i2, s2 = om.hamming_knn2(a, b)
        got = np.stack([i1, s1, i2, s2], 1)
        assert np.array_equal(got, g[f"{case}_{key}"])


def test_oracle_knn_live_cv2_ties():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (6, 32), dtype=np.uint8)
    d1 = base[rng.integers(0, 6, 300)]
    d2 = base[rng.integers(0, 6, 310)]
    mm = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(d1, d2, k=2)
    i1, s1, i2, s2 = om.hamming_knn2(d1, d2)
    for i, row in enumerate(mm):
        assert (row[0].trainIdx, int(row[0].distance), row[1].trainIdx, int(row[1].distance)) == \
            (i1[i], s1[i], i2[i], s2[i])


def test_match_semantics_small():
    # hand-built: 3 queries, 3 train; q0==t1, q1==t1 (competing), q2 far from everything
    z = np.zeros((3, 32), np.uint8)
    t = z.copy(); t[0, :4] = 0xFF; t[1, 4:6] = 0xFF; t[2, 8:20] = 0xFF
    q = z.copy(); q[0] = t[1]; q[1] = t[1]; q[1, 31] = 1; q[2, :] = 0x55
    m, c = om.match(q, t, 0.9, True)
    assert m[0] == 1 and m[1] == -1 and c == (m >= 0).sum()   # mutual check keeps only q0<->t1
    m2, _ = om.match(q, t, 0.9, False)
    assert m2[0] == 1 and m2[1] == 1                            # without it both claim t1


def test_match_edge_cases():
    e = np.zeros((0, 32), np.uint8)
    one = np.zeros((1, 32), np.uint8)
    five = np.arange(160, dtype=np.uint8).reshape(5, 32)
    assert om.match(e, five, 0.9)[1] == 0
    m, c = om.match(five, e, 0.9)
    assert c == 0 and (m == -1).all()
    m, c = om.match(five, one, 0.9)   # a single train row has no second neighbour -> no match
    assert c == 0 and (m == -1).all()
