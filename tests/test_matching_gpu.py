"""GPU parity tests for the Hamming kNN(2) / NNR / mutual matcher, through the C ABI."""
from pathlib import Path

import numpy as np
import pytest

from oracle import matching as om

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "matching_v1.npz"


@pytest.mark.parametrize("case", ["rand", "ties", "corr"])
def test_knn2_golden_bit_exact(fe, case):
    g = np.load(GOLD)
    d1, d2 = g[f"{case}_d1"], g[f"{case}_d2"]
    for a, b, key in ((d1, d2, "knn12"), (d2, d1, "knn21")):
        got = np.stack(fe.hamming_knn2(a, b), 1)
        assert np.array_equal(got, g[f"{case}_{key}"])   # cv2 BFMatcher output, bit-exact


@pytest.mark.parametrize("nnr", [0.9, 0.75, 1.0])
@pytest.mark.parametrize("best_lr", [True, False])
@pytest.mark.parametrize("case", ["rand", "ties", "corr"])
def test_match_vs_oracle(fe, case, nnr, best_lr):
    g = np.load(GOLD)
    d1, d2 = g[f"{case}_d1"], g[f"{case}_d2"]
    m, c = fe.match(d1, d2, nnr, best_lr)
    mo, co = om.match(d1, d2, nnr, best_lr)
    assert c == co and np.array_equal(m, mo)


def test_match_edge_cases(fe):
    e = np.zeros((0, 32), np.uint8)
    one = np.zeros((1, 32), np.uint8)
    five = np.arange(160, dtype=np.uint8).reshape(5, 32)
    m, c = fe.match(e, five, 0.9)
    assert c == 0 and len(m) == 0
    m, c = fe.match(five, e, 0.9)
    assert c == 0 and (m == -1).all()
    m, c = fe.match(five, one, 0.9)
    assert c == 0 and (m == -1).all()
    i1, s1, i2, s2 = fe.hamming_knn2(five, one)
    assert (i1 == 0).all() and (i2 == -1).all() and (s2 == -1).all()
    # ragged sizes around the tile / block boundaries
    rng = np.random.default_rng(3)
    for n1, n2 in [(1, 2), (63, 65), (64, 256), (65, 257), (255, 513), (1, 1000), (1000, 2)]:
        a = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        b = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        assert np.array_equal(np.stack(fe.hamming_knn2(a, b), 1), np.stack(om.hamming_knn2(a, b), 1))
        m, c = fe.match(a, b, 0.95, True)
        mo, co = om.match(a, b, 0.95, True)
        assert c == co and np.array_equal(m, mo)


def test_match_full_size_properties(fe):
    """KITTI-config size (1500 x 1500): oracle parity + size-independent properties."""
    rng = np.random.default_rng(42)
    a = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    noise = rng.random((1500, 256)) < 0.1
    perm = rng.permutation(1500)
    b = np.packbits(np.unpackbits(a[perm], axis=1) ^ noise, axis=1)
    m, c = fe.match(a, b, 0.9, True)
    mo, co = om.match(a, b, 0.9, True)
    assert c == co and np.array_equal(m, mo)
    # mutual matching is an involution: matching b->a gives the inverse map
    m21, c21 = fe.match(b, a, 0.9, True)
    assert c21 == c
    for i in np.nonzero(m >= 0)[0]:
        assert m21[m[i]] == i
    # planted correspondences are recovered
    inv = np.empty(1500, np.int64); inv[perm] = np.arange(1500)
    hit = m >= 0
    assert hit.sum() > 1400 and (m[hit] == inv[hit]).all()
    # self-matching: every row is its own nearest neighbour at distance 0
    i1, s1, _, _ = fe.hamming_knn2(a, a)
    assert (i1 == np.arange(1500)).all() and (s1 == 0).all()


def test_max_sizes(fe):
    """Largest supported train set (65535 rows, index packs into 16 bits)."""
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (65535, 32), dtype=np.uint8)
    b[65534] = a[7]   # plant an exact match at the last index
    i1, s1, i2, s2 = fe.hamming_knn2(a, b)
    assert i1[7] == 65534 and s1[7] == 0
    assert np.array_equal(np.stack([i1, s1, i2, s2], 1), np.stack(om.hamming_knn2(a, b), 1))
