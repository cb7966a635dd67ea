"""TEST INFRASTRUCTURE (oracle) - CPU restatement of the landmark descriptor maintenance of the reference:
MapPoint::updateAverageDescDir / MapLine::updateAverageDescDir, src/mapFeatures.cpp:51-93 and :121-163 (identical
bodies).  Restated from the in-tree source; there are no reference tests or golden vectors for it ("parity unpinned":
the reference cannot be compiled here)."""
import numpy as np

_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming(a, b):
    """cv::norm(a, b, NORM_HAMMING) on two 32-byte rows (src/mapFeatures.cpp:63,133)."""
    return int(_POP[np.bitwise_xor(a, b)].sum())


def median_descriptor(desc, dirs=None):
    """desc: uint8 [n, 32] (n >= 2: a landmark's first observation sets med_desc directly, :25-38, and the position
    read at :76 is past the row for n = 1).  Returns (max_idx, med_obs_dir or None)."""
    n = len(desc)
    assert n >= 2
    conf = np.zeros((n, n), np.int32)                                   # :57-68
    for i in range(n):
        for j in range(i + 1, n):
            conf[i, j] = conf[j, i] = hamming(desc[i], desc[j])
    max_dist, max_idx = 99999, 0                                        # :71-72
    for i in range(n):
        row = sorted(int(v) for v in conf[i])                           # :75-78
        idx_median = row[int(1 + 0.5 * (n - 1))]                        # :79
        if idx_median < max_dist:                                       # :80 (strict: first row wins ties)
            max_dist, max_idx = idx_median, i
    med_dir = None
    if dirs is not None:                                                # :87-90 (accumulator started from zero: the
        acc = np.zeros(3)                                               #  reference's Vector3d is uninitialised)
        for i in range(n):
            acc = acc + np.asarray(dirs[i], np.float64)
        med_dir = acc / n
    return max_idx, med_dir
