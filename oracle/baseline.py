"""CPU baseline of the front-end hot path: the reference's own third-party arithmetic (OpenCV ORB, OpenCV LSD,
OpenCV BFMatcher — via python cv2 4.13, the only OpenCV in this image) + the C restatements of the vendored LBD /
KeyLine stage and of the in-tree Gauss-Newton.  TEST / BENCH INFRASTRUCTURE ONLY: timed by bench.py as
`cpu_baseline` / `--impl reference`, never part of the product path.

Timed interval = what app/plslam_dataset.cpp:126-129 brackets (extract L+R -> stereo match -> f2f track -> pose).
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from oracle import clib
from oracle import frontend as ofe
from oracle.cvref import lsd_cv2, orb_cv2


def match_cv2(d1, d2, nnr, best_lr=True):
    """stvo-pl match() with cv::BFMatcher doing the kNN (as the reference does)."""
    import cv2
    n1, n2 = len(d1), len(d2)
    m12 = np.full(n1, -1, np.int32)
    if n1 == 0 or n2 < 2:
        return m12, 0
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)

    def nnr_dir(a, b):
        out = np.full(len(a), -1, np.int32)
        if len(a) == 0 or len(b) < 2:
            return out
        for i, row in enumerate(bf.knnMatch(a, b, k=2)):
            if len(row) == 2 and np.float32(row[0].distance) < np.float32(row[1].distance) * np.float32(nnr):
                out[i] = row[0].trainIdx
        return out
    m12 = nnr_dir(d1, d2)
    if best_lr:
        m21 = nnr_dir(d2, d1)
        hit = np.nonzero(m12 >= 0)[0]
        bad = hit[m21[m12[hit]] != hit]
        m12[bad] = -1
    return m12, int((m12 >= 0).sum())


def make_fns(prm):
    orb_fn = lambda im: orb_cv2(im, nfeatures=prm["orb_nfeatures"], nlevels=prm["orb_nlevels"], fast_th=prm["orb_fast_th"])
    lines_fn = lambda im: ofe.detect_lines(im, prm["lsd_nfeatures"], prm["min_line_length"], segs=lsd_cv2(im))
    return orb_fn, lines_fn


_G = {}


def _extract_one(idx):
    cam, pairs, prm = _G["cam"], _G["pairs"], _G["prm"]
    orb_fn, lines_fn = make_fns(prm)
    L, R = pairs[idx]
    return ofe.extract_stereo(cam, L, R, prm, orb_fn, lines_fn, match_cv2)


def effective_cores():
    """Host cores this process can actually use: the affinity mask, capped by a cgroup CPU quota when there is one (a
    GPU slot of a shared box shows all of the node's CPUs in os.cpu_count() but is scheduled on a fraction of them)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    quota = None
    try:                                                    # cgroup v2
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(p)
    except Exception:
        try:                                                # cgroup v1
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and p > 0:
                quota = q / p
        except Exception:
            pass
    if quota is not None:
        n = max(1, min(n, int(quota + 0.5)))
    return n


class Workers:
    """A pool of worker processes (fork) that outlives one run(): forking and joining 128 processes costs seconds, which
    must not be charged to every timed step of the baseline."""

    def __init__(self, cam, pairs, prm=None, threads=None, cv_threads=1):
        import multiprocessing as mp

        import cv2
        self.prm = dict(ofe.DEFAULTS, **(prm or {}))
        self.threads = max(1, min(threads or effective_cores(), len(pairs)))
        cv2.setNumThreads(cv_threads)                   # one OpenCV thread per worker (inherited by the fork); measured on the
        _G.update(cam=cam, pairs=pairs, prm=self.prm)   # box, BLAS / OpenMP thread limits make no difference to this path
        self.cam, self.pairs = cam, pairs
        self.pool = mp.get_context("fork").Pool(self.threads) if self.threads > 1 else None

    def extract(self, n=None):
        idx = range(len(self.pairs) if n is None else n)
        if self.pool is None:
            return [_extract_one(i) for i in idx]
        return self.pool.map(_extract_one, idx, chunksize=1)

    def run(self, n=None):
        """One pass over the first n pairs as one sequence: extraction + stereo association spread over the workers,
        then frame-to-frame tracking and pose refinement sequentially, as the reference's loop does."""
        import time
        t0 = time.perf_counter()
        frames = self.extract(n)
        t1 = time.perf_counter()
        pairs = self.pairs if n is None else self.pairs[:n]
        res = ofe.run_sequence(self.cam, pairs, self.prm, match_fn=match_cv2, frames=frames)
        self.last_split = (t1 - t0, time.perf_counter() - t1)   # (parallel extraction, sequential tracking + pose) seconds
        return res

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def run(cam, pairs, prm=None, threads=None, cv_threads=1):
    """Processes the stereo pairs as one sequence.  Extraction + stereo association of the frames are independent
    and are spread over `threads` worker processes (fork; default = all host cores); frame-to-frame tracking and
    the pose refinement then run sequentially, as the reference's loop does.  Returns (results, cores used)."""
    with Workers(cam, pairs, prm, threads, cv_threads) as w:
        return w.run(), w.threads
