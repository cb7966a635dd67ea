"""CPU tests of oracle/loopclosure.py: the restatement of MapHandler::isLoopClosure +
computeRelativePoseRobustGN (src/mapHandler.cpp:3192-3300, :3566-3957) on planted problems - acceptance of a true
revisit, each rejection branch, the returned pose and correspondences."""
import numpy as np

from oracle import clib
from oracle import loopclosure as olc
from plslam_b200 import synth

CAM = dict(width=752, height=480, fx=458.654, fy=457.296, cx=367.215, cy=248.375, b=0.110077842)
PRM = dict(min_ratio_12_p=0.9, min_ratio_12_l=0.9, best_lr_matches=True, has_points=True, has_lines=True, homog_th=1e-7,
           max_iters=5, max_iters_ref=10)


def keyframes(seed=0, x_true=None, n_pts=300, n_lines=80, flip_bits=6, shuffle=True, outlier_frac=0.1):
    """kf0 holds 3-D features with random 256-bit descriptors; kf1 observes them under the planted increment with noisy
    copies of the descriptors (flip_bits random bit flips), in a shuffled order, plus unrelated distractors."""
    g = synth.gn_problem(CAM, n_pts=n_pts, n_lines=n_lines, seed=seed, x_true=x_true, outlier_frac=outlier_frac)
    rng = np.random.default_rng(seed + 100)

    def noisy(d):
        out = d.copy()
        for row in out:
            for b in rng.integers(0, 256, flip_bits):
                row[b >> 3] ^= np.uint8(1 << (b & 7))
        return out
    pd0 = rng.integers(0, 256, (n_pts, 32), dtype=np.uint8); ld0 = rng.integers(0, 256, (n_lines, 32), dtype=np.uint8)
    n_dp, n_dl = n_pts // 5, n_lines // 5                          # distractors in kf1
    pd1 = np.concatenate([noisy(pd0), rng.integers(0, 256, (n_dp, 32), dtype=np.uint8)])
    ld1 = np.concatenate([noisy(ld0), rng.integers(0, 256, (n_dl, 32), dtype=np.uint8)])
    pl1 = np.concatenate([g["obs"], rng.uniform(0, 400, (n_dp, 2))]); le1 = np.concatenate([g["le"], rng.normal(0, 1, (n_dl, 3))])
    pp = rng.permutation(len(pd1)) if shuffle else np.arange(len(pd1))
    lp = rng.permutation(len(ld1)) if shuffle else np.arange(len(ld1))
    kf0 = dict(pdesc=pd0, P=g["P"], pl=np.zeros((n_pts, 2)), ldesc=ld0, sP=g["sP"], eP=g["eP"], le=np.zeros((n_lines, 3)))
    kf1 = dict(pdesc=pd1[pp], P=np.zeros((len(pp), 3)), pl=pl1[pp], ldesc=ld1[lp], sP=np.zeros((len(lp), 3)),
               eP=np.zeros((len(lp), 3)), le=le1[lp])
    return kf0, kf1, g, np.argsort(pp), np.argsort(lp)


def test_true_revisit_is_accepted_and_pose_recovered():
    x_true = np.array([0.08, -0.03, 0.4, 0.01, -0.04, 0.006])
    kf0, kf1, g, inv_p, inv_l = keyframes(seed=1, x_true=x_true)
    r = olc.is_loop_closure(CAM, kf0, kf1, PRM)
    assert r["estimated"] and r["accepted"]
    assert r["common_pt"] == 300 and r["common_ls"] == 80            # every planted descriptor pair is found
    # x_inc = logmap(T_inc) with T_inc ~ the planted increment; pose_inc is its inverse
    assert np.linalg.norm(r["x_inc"] - x_true) < 0.02
    assert np.allclose(clib.expmap_se3(r["pose_inc"]) @ clib.expmap_se3(r["x_inc"]), np.eye(4), atol=1e-9)
    # correspondences: (i1, i2) with i2 the shuffled position of i1, gross outliers removed by the chi2 gate
    assert all(inv_p[i1] == i2 for i1, i2 in r["pt_pairs"]) and all(inv_l[i1] == i2 for i1, i2 in r["ls_pairs"])
    assert 230 <= len(r["pt_pairs"]) <= 275 and 55 <= len(r["ls_pairs"]) <= 74   # ~10 % planted outliers dropped
    assert r["max_cov_eig"] < 0.01 and r["err"] < 1.0


def test_unrelated_keyframes_stop_at_the_inlier_ratio_gate():
    kf0, kf1, _, _, _ = keyframes(seed=2)
    rng = np.random.default_rng(9)
    kf1["pdesc"] = rng.integers(0, 256, kf1["pdesc"].shape, dtype=np.uint8)
    r = olc.is_loop_closure(CAM, kf0, kf1, PRM)
    assert not r["estimated"] and not r["accepted"] and r["inl_ratio_pt"] < 30.0


def test_rejection_branches():
    big = np.array([0.5, 0.2, 1.9, 0.02, 0.05, 0.0])                  # |t| = 1.97 > lc_trs = 1.5
    kf0, kf1, _, _, _ = keyframes(seed=3, x_true=big)
    r = olc.is_loop_closure(CAM, kf0, kf1, PRM)
    assert r["estimated"] and not r["accepted"] and r["t"] > 1.5
    assert olc.is_loop_closure(CAM, kf0, kf1, PRM, dict(lc_trs=3.0))["accepted"]
    kf0, kf1, _, _, _ = keyframes(seed=4)
    assert not olc.is_loop_closure(CAM, kf0, kf1, PRM, dict(lc_unc=1e-12))["accepted"]      # covariance test
    assert not olc.is_loop_closure(CAM, kf0, kf1, PRM, dict(lc_res=1e-9))["accepted"]       # residual test
    assert not olc.is_loop_closure(CAM, kf0, kf1, PRM, dict(lc_rot=0.01))["accepted"]       # rotation test
    assert olc.is_loop_closure(CAM, kf0, kf1, PRM, dict(lc_inl=2.0))["accepted"]            # lc_inl is overridden (:3900)


def test_points_only_and_empty_frames():
    kf0, kf1, _, _, _ = keyframes(seed=5)
    r = olc.is_loop_closure(CAM, kf0, kf1, dict(PRM, has_lines=False))
    assert r["accepted"] and len(r["ls_pairs"]) == 0 and r["common_ls"] == 0
    empty = dict(pdesc=np.zeros((0, 32), np.uint8), P=np.zeros((0, 3)), pl=np.zeros((0, 2)), ldesc=np.zeros((0, 32), np.uint8),
                 sP=np.zeros((0, 3)), eP=np.zeros((0, 3)), le=np.zeros((0, 3)))
    r = olc.is_loop_closure(CAM, kf0, empty, PRM)
    assert not r["estimated"] and not r["accepted"]            # 0 / 0 ratios compare false, as in the reference
