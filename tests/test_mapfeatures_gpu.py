"""GPU parity: plf_median_descriptors (batched MapPoint/MapLine::updateAverageDescDir) vs oracle/mapfeatures.py."""
import numpy as np
import pytest

import plslam_b200 as plf
from oracle import mapfeatures as mf

pytestmark = pytest.mark.gpu


def _landmarks(rng, sizes, tie_heavy=False):
    descs, dirs = [], []
    for n in sizes:
        if tie_heavy:      # few distinct descriptors -> many equal distances and equal medians
            base = rng.integers(0, 256, (3, 32), dtype=np.uint8)
            d = base[rng.integers(0, 3, n)]
        else:              # noisy copies of one descriptor, as a tracked landmark's observations are
            proto = rng.integers(0, 256, 32, dtype=np.uint8)
            flips = (rng.random((n, 256)) < 0.08)
            d = proto[None, :] ^ np.packbits(flips, axis=1)
        descs.append(d.astype(np.uint8))
        dirs.append(rng.normal(0, 1, (n, 3)))
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return np.concatenate(descs), off, np.concatenate(dirs)


@pytest.mark.parametrize("tie_heavy", [False, True])
def test_median_descriptors_vs_oracle(fe, tie_heavy):
    rng = np.random.default_rng(7 + tie_heavy)
    sizes = [2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 64] + list(rng.integers(2, 40, 300))
    desc, off, dirs = _landmarks(rng, sizes, tie_heavy)
    idx, md = fe.median_descriptors(desc, off, dirs)
    for l in range(len(sizes)):
        ri, rd = mf.median_descriptor(desc[off[l]:off[l + 1]], dirs[off[l]:off[l + 1]])
        assert idx[l] == ri, (l, sizes[l])
        assert np.array_equal(md[l], rd), l      # same f64 sum order -> bit-identical mean direction


def test_median_descriptors_edge_cases(fe):
    rng = np.random.default_rng(3)
    idx, md = fe.median_descriptors(np.zeros((0, 32), np.uint8), np.array([0], np.int32))
    assert len(idx) == 0 and md is None
    d = rng.integers(0, 256, (4, 32), dtype=np.uint8)
    idx, md = fe.median_descriptors(d, np.array([0, 2, 4], np.int32))          # directions optional
    assert list(idx) == [0, 0] and md is None
    with pytest.raises(plf.PlfError, match="one observation"):
        fe.median_descriptors(d, np.array([0, 1, 4], np.int32))
    with pytest.raises(plf.PlfError, match="max 64"):
        fe.median_descriptors(rng.integers(0, 256, (65, 32), dtype=np.uint8), np.array([0, 65], np.int32))
