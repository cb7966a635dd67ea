"""CPU checks of the local-bundle-adjustment restatement (oracle/lba.c; MapHandler::levMarquardtOptimizationLBA,
src/mapHandler.cpp:1332-1989)."""
import numpy as np

import plslam_b200 as plf
from oracle import clib
from plslam_b200 import synth

CAM = plf.KITTI_CAMERA


def _residuals(p, X, nkf):
    """the reference's scalar residuals (norm of the point reprojection error) at X"""
    kf, pt = X[:6 * nkf].reshape(-1, 6), X[6 * nkf:].reshape(-1, 3)
    r = []
    for lm, k, xy in zip(p["pt_obs_lm"], p["pt_obs_kf"], p["pt_obs_xy"]):
        T = clib.expmap_se3(kf[k]) if k >= 0 else p["fixed_T"][-1 - k]
        Ti = np.linalg.inv(T)
        Pc = Ti[:3, :3] @ pt[lm] + Ti[:3, 3]
        r.append(np.hypot(xy[0] - (CAM["cx"] + CAM["fx"] * Pc[0] / Pc[2]), xy[1] - (CAM["cy"] + CAM["fy"] * Pc[1] / Pc[2])))
    return np.array(r)


def test_first_step_is_the_damped_gauss_newton_step_on_the_pose_block():
    """The restated Jacobians / weights / update parametrisation reproduce a numerically differentiated Gauss-Newton step
    (Cauchy-weighted scalar residuals, H_ii += lambda H_ii, T <- T inverse(exp(dx))) on the keyframe poses."""
    p = synth.lba_problem(CAM, seed=1, pose_noise=0.003, lm_noise=0.001, px_noise=0.5, n_ls=0, n_pt=60)   # residuals away from 0: |e| is smooth
    nkf, npt = len(p["kf_pose"]), len(p["pt"])
    N = 6 * nkf + 3 * npt
    X0 = np.concatenate([p["kf_pose"].ravel(), p["pt"].ravel()])

    def apply(X, D):
        Y = X.copy()
        for i in range(nkf):
            Y[6 * i:6 * i + 6] = clib.logmap_se3(clib.expmap_se3(X[6 * i:6 * i + 6]) @ clib.inverse_se3(clib.expmap_se3(D[6 * i:6 * i + 6])))
        Y[6 * nkf:] += D[6 * nkf:]
        return Y
    r0 = _residuals(p, X0, nkf)
    J = np.zeros((len(r0), N))
    for j in range(N):
        D = np.zeros(N); D[j] = 1e-6
        J[:, j] = (_residuals(p, apply(X0, D), nkf) - r0) / 1e-6
    w = 1 / (1 + r0 ** 2)
    H = (J.T * w) @ J
    g = -(J.T * w) @ r0
    D = np.linalg.solve(H + np.diag(1e-5 * np.abs(np.diag(H)).max() * np.diag(H)), g)
    r = clib.local_ba(CAM, p, clib.lba_opts(ref_quirks=0, max_iters=1))
    want = apply(X0, D)[:6 * nkf].reshape(-1, 6) - p["kf_pose"]
    got = r["kf_pose"] - p["kf_pose"]
    assert np.abs(got - want).max() < 2e-2 * np.abs(want).max()
    assert _residuals(p, np.concatenate([r["kf_pose"].ravel(), r["pt"].ravel()]), nkf).mean() < r0.mean()


def test_reference_quirks_are_observable():
    p = synth.lba_problem(CAM, seed=2)
    a = clib.local_ba(CAM, p, clib.lba_opts(ref_quirks=1))
    b = clib.local_ba(CAM, p, clib.lba_opts(ref_quirks=0))
    assert a["rc"] == 0 and b["rc"] == 0 and np.isfinite(a["kf_pose"]).all() and np.isfinite(b["kf_pose"]).all()
    # (q1) err_prev = +inf after the first pass: the as-written mode always accepts the second step; the modes part ways
    assert a["iters"] != b["iters"] or not np.allclose(a["ls"], b["ls"])
    # a problem without observations is refused like the reference (:1324-1328 returns -1)
    empty = dict(p, pt_obs_lm=np.zeros(0, np.int32), pt_obs_kf=np.zeros(0, np.int32), pt_obs_xy=np.zeros((0, 2)),
                 ls_obs_lm=np.zeros(0, np.int32), ls_obs_kf=np.zeros(0, np.int32), ls_obs_le=np.zeros((0, 3)))
    assert clib.local_ba(CAM, empty)["rc"] == -1
