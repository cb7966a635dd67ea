"""GPU parity of the local bundle adjustment (plf_local_ba, csrc/lba.cu) against the CPU restatement of
MapHandler::levMarquardtOptimizationLBA (oracle/lba.c; src/mapHandler.cpp:1332-1989).  The oracle factors the dense
damped Hessian (LDL^T, like the reference's SimplicialLDLT); the device solves the same normal equations through the Schur
complement on the landmarks, so agreement is to rounding: 1e-6 relative on the update of every block."""
import numpy as np
import pytest

import plslam_b200 as plf
from oracle import clib
from plslam_b200 import synth

pytestmark = pytest.mark.gpu
CAM = plf.KITTI_CAMERA
TOL = 1e-6


def _check(p, got, ref):
    assert got["iters"] == ref["iters"]
    for k in ("kf_pose", "pt", "ls"):
        step = np.abs(ref[k] - p[k]).max() if ref[k].size else 0.0
        if ref[k].size:
            assert np.abs(got[k] - ref[k]).max() <= TOL * max(step, 1e-3), k
    assert np.array_equal(got["pt_moved"], ref["pt_moved"]) and np.array_equal(got["ls_moved"], ref["ls_moved"])
    assert abs(got["lambda_"] - ref["lambda_"]) <= 1e-9 * abs(ref["lambda_"])
    if np.isfinite(ref["err"]):
        assert abs(got["err"] - ref["err"]) <= 1e-9 * abs(ref["err"])


@pytest.mark.parametrize("quirks", [1, 0])
@pytest.mark.parametrize("seed,kw", [(1, {}), (2, dict(n_kf=8, n_fixed=3, n_pt=400, n_ls=120)), (3, dict(n_ls=0)),
                                     (4, dict(n_pt=0, n_ls=60)), (5, dict(n_kf=1, n_fixed=4)), (6, dict(n_fixed=0, n_kf=6))])
def test_local_ba_matches_the_oracle(fe, seed, kw, quirks):
    p = synth.lba_problem(CAM, seed=seed, **kw)
    ref = clib.local_ba(CAM, p, clib.lba_opts(ref_quirks=quirks))
    got = fe.local_ba(p, ref_quirks=quirks)
    assert ref["rc"] == 0
    _check(p, got, ref)


def test_local_ba_iteration_limits_and_landmark_only(fe):
    p = synth.lba_problem(CAM, seed=7, n_kf=3)
    for it in (1, 2, 4):
        _check(p, fe.local_ba(p, max_iters=it, ref_quirks=0), clib.local_ba(CAM, p, clib.lba_opts(max_iters=it, ref_quirks=0)))
    # every observing keyframe fixed: structure-only refinement (no reduced system)
    q = synth.lba_problem(CAM, seed=8, n_kf=0, n_fixed=5)
    _check(q, fe.local_ba(q, ref_quirks=0), clib.local_ba(CAM, q, clib.lba_opts(ref_quirks=0)))


def test_local_ba_rejects_bad_input(fe):
    p = synth.lba_problem(CAM, seed=9)
    empty = dict(p, pt_obs_lm=np.zeros(0, np.int32), pt_obs_kf=np.zeros(0, np.int32), pt_obs_xy=np.zeros((0, 2)),
                 ls_obs_lm=np.zeros(0, np.int32), ls_obs_kf=np.zeros(0, np.int32), ls_obs_le=np.zeros((0, 3)))
    with pytest.raises(plf.PlfError):
        fe.local_ba(empty)
    bad = dict(p, pt_obs_lm=p["pt_obs_lm"][::-1].copy())   # not grouped in ascending landmark order
    with pytest.raises(plf.PlfError):
        fe.local_ba(bad)
