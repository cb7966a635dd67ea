"""CPU checks of the landmark descriptor-maintenance oracle (oracle/mapfeatures.py, restating src/mapFeatures.cpp:51-93)."""
import numpy as np

from oracle import mapfeatures as mf


def _desc(bits_list):
    out = np.zeros((len(bits_list), 32), np.uint8)
    for r, nbits in enumerate(bits_list):     # descriptor r has its first nbits bits set: d(r, s) = |nbits_r - nbits_s|
        for b in range(nbits):
            out[r, b // 8] |= 1 << (b % 8)
    return out


def test_two_observations_first_wins():
    # n = 2: position int(1 + 0.5) = 1 of the sorted row {0, d} is d for both rows -> the strict '<' keeps row 0
    idx, _ = mf.median_descriptor(_desc([0, 40]))
    assert idx == 0


def test_position_read_is_upper_median():
    # distances on a line: rows at 0, 10, 11, 12, 100 bits.  n = 5 -> position int(1 + 2) = 3 of each sorted row.
    d = _desc([0, 10, 11, 12, 100])
    # row sorted distances: r0 [0,10,11,12,100]->12; r1 [0,1,2,10,90]->10; r2 [0,1,1,11,89]->11; r3 [0,1,2,12,88]->12; r4 ->90
    idx, _ = mf.median_descriptor(d)
    assert idx == 1


def test_ties_keep_first_and_direction_mean():
    d = _desc([5, 5, 5, 5])                     # all identical: every median is 0 -> row 0
    dirs = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [1.0, 1.0, 1.0]])
    idx, md = mf.median_descriptor(d, dirs)
    assert idx == 0 and np.array_equal(md, np.array([0.5, 0.5, 0.5]))


def test_hamming_matches_numpy_unpackbits():
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
    assert mf.hamming(a, b) == int(np.unpackbits(a ^ b).sum())
