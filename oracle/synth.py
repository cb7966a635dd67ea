"""Seeded synthetic stereo imagery.  TEST / BENCH INFRASTRUCTURE ONLY.

scene_pair(): the C1 "plumbing" pair — filled random-gray rectangles + thin lines on gray 90, 3x3
sigma-0.8 blur, +-3 uniform noise, right image = left shifted by a constant disparity.
Stream rendering (planted SE(3) trajectory through a 3-D world of textured quads and segments) is in
world_* below.  Drawing uses cv2 (python OpenCV is part of the image on both the build container and
the GPU box); everything is deterministic given the seed.
"""
import cv2
import numpy as np


def scene_pair(w=1242, h=375, seed=1, n_rect=220, n_lines=120, disparity=20, noise=3):
    rng = np.random.default_rng(seed)
    W = w + disparity
    img = np.full((h, W), 90, np.uint8)
    for _ in range(n_rect):
        x0 = int(rng.integers(-40, W)); y0 = int(rng.integers(-40, h))
        rw = int(rng.integers(12, 160)); rh = int(rng.integers(10, 110))
        g = int(rng.integers(20, 236))
        cv2.rectangle(img, (x0, y0), (x0 + rw, y0 + rh), g, -1)
    for _ in range(n_lines):
        p0 = (int(rng.integers(0, W)), int(rng.integers(0, h)))
        ang = rng.uniform(0, np.pi); L = rng.uniform(30, 300)
        p1 = (int(p0[0] + L * np.cos(ang)), int(p0[1] + L * np.sin(ang)))
        cv2.line(img, p0, p1, int(rng.integers(0, 256)), int(rng.integers(1, 3)), cv2.LINE_8)
    img = cv2.GaussianBlur(img, (3, 3), 0.8)
    def noisy(a, r):
        n = r.integers(-noise, noise + 1, a.shape)
        return np.clip(a.astype(np.int16) + n, 0, 255).astype(np.uint8)
    left = noisy(img[:, disparity:disparity + w], rng)
    right = noisy(img[:, 0:w], rng)
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


# ---- synthetic correspondences for the pose optimiser (planted pose) ---------------
def project(cam, P):
    P = np.asarray(P, np.float64)
    return np.stack([cam["cx"] + cam["fx"] * P[:, 0] / P[:, 2], cam["cy"] + cam["fy"] * P[:, 1] / P[:, 2]], 1)


def gn_problem(cam, n_pts=300, n_lines=80, seed=0, x_true=None, px_noise=0.3, outlier_frac=0.1):
    """3-D points / segments in the previous camera frame + their observations in the current frame under a
    planted increment T_true (current <- previous), with pixel noise and gross outliers."""
    from oracle import clib
    rng = np.random.default_rng(seed)
    if x_true is None:
        x_true = np.array([0.05, -0.02, 0.9, 0.01, -0.03, 0.005])
    T = clib.expmap_se3(x_true)
    def rand_pts(n):
        z = rng.uniform(4, 40, n)
        u = rng.uniform(30, cam["width"] - 30, n); v = rng.uniform(20, cam["height"] - 20, n)
        return np.stack([(u - cam["cx"]) * z / cam["fx"], (v - cam["cy"]) * z / cam["fy"], z], 1)
    P = rand_pts(n_pts)
    Pc = P @ T[:3, :3].T + T[:3, 3]
    obs = project(cam, Pc) + rng.normal(0, px_noise, (n_pts, 2))
    n_out = int(outlier_frac * n_pts)
    obs[:n_out] += rng.uniform(-60, 60, (n_out, 2))
    sP = rand_pts(n_lines)
    eP = sP + rng.normal(0, 1.0, (n_lines, 3)) * np.array([1.5, 1.0, 0.5])
    eP[:, 2] = np.maximum(eP[:, 2], 2.0)
    sp = project(cam, sP @ T[:3, :3].T + T[:3, 3]) + rng.normal(0, px_noise, (n_lines, 2))
    ep = project(cam, eP @ T[:3, :3].T + T[:3, 3]) + rng.normal(0, px_noise, (n_lines, 2))
    n_outl = int(outlier_frac * n_lines)
    sp[:n_outl] += rng.uniform(-40, 40, (n_outl, 2))
    sph = np.concatenate([sp, np.ones((n_lines, 1))], 1); eph = np.concatenate([ep, np.ones((n_lines, 1))], 1)
    le = np.cross(sph, eph)
    le = le / np.sqrt(le[:, 0:1] ** 2 + le[:, 1:2] ** 2)      # le = (sp x ep) / ||(le0, le1)||  
    return dict(P=P, obs=obs, sP=sP, eP=eP, le=le, T_true=T, x_true=np.asarray(x_true, np.float64))
