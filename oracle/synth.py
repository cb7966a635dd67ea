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
