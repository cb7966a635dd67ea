"""GPU parity test of plf_loop_closure_pose against oracle/loopclosure.py, the restatement of
MapHandler::isLoopClosure + computeRelativePoseRobustGN (src/mapHandler.cpp:3192-3300, :3566-3957): same matches, same
decision on every branch, same surviving correspondences; pose within 1e-6 relative on the se(3) log (the GN bar of
tests/test_gn_gpu.py); and the lookForCommonMatches refinement (:768-808) on plf_gn_pose."""
import numpy as np
import pytest

import plslam_b200 as plf
from oracle import clib
from oracle import loopclosure as olc
from test_loopclosure_oracle import CAM, PRM, keyframes

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-12)


def check(fe, kf0, kf1, prm, lc=None):
    r = olc.is_loop_closure(CAM, kf0, kf1, prm, lc)
    g = fe.loop_closure_pose(kf0, kf1, lc)
    assert g["estimated"] == r["estimated"] and g["accepted"] == r["accepted"]
    assert (g["common_pt"], g["common_ls"]) == (r["common_pt"], r["common_ls"])
    for k in ("inl_ratio_pt", "inl_ratio_ls"):
        assert g[k] == r[k] or (np.isnan(g[k]) and np.isnan(r[k]))
    if r["estimated"]:
        assert rel(g["x_inc"], r["x_inc"]) < 1e-6 and abs(g["err"] - r["err"]) <= 1e-6 * max(r["err"], 1e-12)
        assert abs(g["max_cov_eig"] - r["max_cov_eig"]) <= 1e-5 * r["max_cov_eig"]
        assert abs(g["t"] - r["t"]) < 1e-9 and abs(g["r"] - r["r"]) < 1e-7 and g["ratio_inliers"] == r["ratio_inliers"]
    if r["accepted"]:
        assert np.array_equal(g["pt_pairs"], r["pt_pairs"]) and np.array_equal(g["ls_pairs"], r["ls_pairs"])
        assert rel(g["pose_inc"], r["pose_inc"]) < 1e-6
    return r, g


def test_loop_closure_pose_matches_oracle(built):
    x_true = np.array([0.08, -0.03, 0.4, 0.01, -0.04, 0.006])
    with plf.Frontend(camera=CAM, max_iters=PRM["max_iters"], max_iters_ref=PRM["max_iters_ref"]) as fe:
        n0 = fe.launches
        kf0, kf1, _, _, _ = keyframes(seed=1, x_true=x_true)
        r, _ = check(fe, kf0, kf1, PRM)
        assert r["accepted"] and fe.launches - n0 >= 5                # matcher (2 directions x 2) + GN ran on the device
        for seed in (2, 3, 4, 5, 6, 7):
            kf0, kf1, _, _, _ = keyframes(seed=seed, outlier_frac=0.05 * (seed % 4))
            check(fe, kf0, kf1, PRM)
        # every rejection branch and the overridden inlier test
        kf0, kf1, _, _, _ = keyframes(seed=3, x_true=np.array([0.5, 0.2, 1.9, 0.02, 0.05, 0.0]))
        assert not check(fe, kf0, kf1, PRM)[0]["accepted"]
        assert check(fe, kf0, kf1, PRM, dict(lc_trs=3.0))[0]["accepted"]
        kf0, kf1, _, _, _ = keyframes(seed=4)
        for lc in (dict(lc_unc=1e-12), dict(lc_res=1e-9), dict(lc_rot=0.01), dict(lc_inl=2.0), dict(lc_inlier_ratio=101.0)):
            check(fe, kf0, kf1, PRM, lc)
        # unrelated descriptors: stops at the pre-condition; empty frame: 0 / 0 ratios
        rng = np.random.default_rng(9)
        bad = dict(kf1, pdesc=rng.integers(0, 256, kf1["pdesc"].shape, dtype=np.uint8))
        assert not check(fe, kf0, bad, PRM)[0]["estimated"]
        empty = dict(pdesc=np.zeros((0, 32), np.uint8), P=np.zeros((0, 3)), pl=np.zeros((0, 2)), ldesc=np.zeros((0, 32), np.uint8),
                     sP=np.zeros((0, 3)), eP=np.zeros((0, 3)), le=np.zeros((0, 3)))
        assert not check(fe, kf0, empty, PRM)[0]["estimated"]
    with plf.Frontend(camera=CAM, has_lines=0) as fe:                  # points only (:3288-3292)
        kf0, kf1, _, _, _ = keyframes(seed=5)
        r, g = check(fe, kf0, kf1, dict(PRM, has_lines=False))
        assert r["accepted"] and len(g["ls_pairs"]) == 0


def test_keyframe_refinement_lookforcommonmatches(built):
    """lookForCommonMatches (:768-808) hands matched_pt / matched_ls to StereoFrameHandler::optimizePose and accepts the
    refined DT when n_inliers > minFeatures and the inlier ratios reach kfInlierRatio: optimizePose == plf_gn_pose with the
    ctx thresholds (stvo-pl min_error / min_error_change)."""
    kf0, kf1, g, _, _ = keyframes(seed=8, shuffle=False)
    o = clib.gn_opts(1e-7, 5, 10, 1e-7, 1e-7)
    ref = clib.gn_pose(CAM, g["P"], g["obs"], g["sP"], g["eP"], g["le"], opts=o)
    with plf.Frontend(camera=CAM) as fe:
        got = fe.gn_pose(g["P"], g["obs"], g["sP"], g["eP"], g["le"])
    assert np.array_equal(got["inlier_pt"], ref["inlier_pt"]) and np.array_equal(got["inlier_ls"], ref["inlier_ls"])
    assert rel(got["x"], ref["x"]) < 1e-6
    ratio_pt = 100.0 * got["n_inliers"][0] / len(g["P"]); ratio_ls = 100.0 * got["n_inliers"][1] / len(g["sP"])
    assert sum(got["n_inliers"]) > 10 and ratio_pt >= 50.0 and ratio_ls >= 50.0        # kf_inlier_ratio-style acceptance
    DT = clib.inverse_se3(got["T"])                                                      # curr_frame->DT
    assert np.linalg.norm(clib.logmap_se3(DT) + g["x_true"]) < 0.05 or np.linalg.norm(got["x"] - g["x_true"]) < 0.05
