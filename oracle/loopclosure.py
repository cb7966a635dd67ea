"""TEST INFRASTRUCTURE (oracle).  CPU restatement of the loop-closure relative pose of pl-slam:
MapHandler::isLoopClosure (src/mapHandler.cpp:3192-3300) + MapHandler::computeRelativePoseRobustGN (:3566-3957).

  matches      match(kf0.pdesc_l, kf1.pdesc_l, minRatio12P) :3223, match(ldesc_l, ..., minRatio12L) :3249
               (oracle/matching.py: stvo-pl match(), [UPSTREAM-RECALL])
  rows         P from kf0, pl_obs from kf1 :3230-3232; sP, eP from kf0, le_obs from kf1 :3256-3259
  pre-gate     inl_ratio = max(100 c / n0, 100 c / n1) > lcInlierRatio for the enabled feature types :3277-3299
  pose         two-stage robust GN from identity, chi2 gate between the stages, stop tests on DBL_EPSILON
               (oracle/gn.c, the restatement of :3566-3872)
  decision     e < lcRes, max eigenvalue of H^-1 < lcUnc, |t| < lcTrs, |w| (deg) < lcRot; the inlier-ratio test is
               computed and then overridden (`lc_inl = true;` :3900)   :3875-3906
  outputs      surviving (inlier) correspondences and pose_inc = logmap(inverse(expmap(x_inc))) :3909-3951
"""
import sys

import numpy as np

from oracle import clib
from oracle import matching as om

DEFAULT_LC = dict(lc_res=1.0, lc_unc=0.01, lc_inl=0.3, lc_trs=1.5, lc_rot=35.0, lc_inlier_ratio=30.0)  # src/slamConfig.cpp:73-83


def _std_max(a, b):
    return b if a < b else a      # std::max(a, b): NaN-propagation of the reference's expression


def _div(a, b):
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(a) / np.float64(b))


def is_loop_closure(cam, kf0, kf1, prm, lc=None):
    """kf0 / kf1: dicts with pdesc [n,32], P [n,3], pl [n,2], ldesc [m,32], sP, eP [m,3], le [m,3].
    prm: front-end parameters (min_ratio_12_p/l, best_lr_matches, has_points, has_lines, homog_th, max_iters,
    max_iters_ref).  Returns a dict mirroring plf_lc_result (+ pt_pairs / ls_pairs)."""
    lc = dict(DEFAULT_LC, **(lc or {}))
    has_p, has_l = prm.get("has_points", True), prm.get("has_lines", True)
    n_pt0, n_pt1, n_ls0, n_ls1 = len(kf0["pdesc"]), len(kf1["pdesc"]), len(kf0["ldesc"]), len(kf1["ldesc"])
    out = dict(accepted=False, estimated=False, common_pt=0, common_ls=0, pt_pairs=np.zeros((0, 2), np.int32),
               ls_pairs=np.zeros((0, 2), np.int32))
    mp = np.zeros(0, np.int32); ml = np.zeros(0, np.int32)
    if has_p and n_pt0 and n_pt1:
        mp, out["common_pt"] = om.match(kf0["pdesc"], kf1["pdesc"], prm["min_ratio_12_p"], prm["best_lr_matches"])
    if has_l and n_ls0 and n_ls1:
        ml, out["common_ls"] = om.match(kf0["ldesc"], kf1["ldesc"], prm["min_ratio_12_l"], prm["best_lr_matches"])
    ip = np.nonzero(mp >= 0)[0]; il = np.nonzero(ml >= 0)[0]
    P, obs = kf0["P"][ip].reshape(-1, 3), kf1["pl"][mp[ip]].reshape(-1, 2)
    sP, eP, le = kf0["sP"][il].reshape(-1, 3), kf0["eP"][il].reshape(-1, 3), kf1["le"][ml[il]].reshape(-1, 3)
    rp = _std_max(_div(100.0 * out["common_pt"], n_pt0), _div(100.0 * out["common_pt"], n_pt1))
    rl = _std_max(_div(100.0 * out["common_ls"], n_ls0), _div(100.0 * out["common_ls"], n_ls1))
    out.update(inl_ratio_pt=rp, inl_ratio_ls=rl)
    if has_p and has_l:
        cond = rp > lc["lc_inlier_ratio"] and rl > lc["lc_inlier_ratio"]
    elif has_p:
        cond = rp > lc["lc_inlier_ratio"]
    elif has_l:
        cond = rl > lc["lc_inlier_ratio"]
    else:
        cond = False
    if not cond:
        return out
    eps = sys.float_info.epsilon
    o = clib.gn_opts(prm["homog_th"], prm["max_iters"], prm["max_iters_ref"], eps, eps, eps)
    r = clib.gn_pose(cam, P, obs, sP, eP, le, opts=o)
    out["estimated"] = True
    x = np.asarray(r["x"], np.float64)
    cov = np.asarray(r["cov"], np.float64).reshape(6, 6)
    eig = np.linalg.eigvalsh(0.5 * (cov + cov.T))
    inl_p, inl_l = np.asarray(r["inlier_pt"], bool), np.asarray(r["inlier_ls"], bool)
    n = len(P) + len(sP)
    out.update(err=r["err"], max_cov_eig=float(eig[-1]), ratio_inliers=_div(int(inl_p.sum() + inl_l.sum()), n),
               t=float(np.linalg.norm(x[:3])), r=float(np.linalg.norm(x[3:]) * np.float32(180.0) / np.pi), x_inc=x)
    ok = (r["err"] < lc["lc_res"] and out["max_cov_eig"] < lc["lc_unc"] and out["t"] < lc["lc_trs"] and out["r"] < lc["lc_rot"])
    if not ok:
        return out
    out["accepted"] = True
    out["pt_pairs"] = np.stack([ip[inl_p], mp[ip][inl_p]], 1).astype(np.int32).reshape(-1, 2)
    out["ls_pairs"] = np.stack([il[inl_l], ml[il][inl_l]], 1).astype(np.int32).reshape(-1, 2)
    out["pose_inc"] = clib.logmap_se3(clib.inverse_se3(clib.expmap_se3(x)))
    return out
