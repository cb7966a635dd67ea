"""CPU tests: the ORB / LSD restatements stay bit-exact against cv2 4.13 on the imagery the benchmark and the
pipeline tests actually use (rendered stereo streams: KITTI shape, EuRoC shape, lines-dominant corridor), not only on
the rectangle scenes of the golden files."""
import numpy as np
import pytest

import plslam_b200 as plf
from oracle import clib
from oracle.cvref import lsd_cv2, orb_cv2
from plslam_b200 import synth

FIELDS = ["x", "y", "size", "angle", "response", "octave"]


def corridor_world(n_segs=450, seed=3):
    world = synth.World(seed=21, length=90.0, n_quads=0, n_segs=0)
    rng = np.random.default_rng(seed)
    segs = []
    for _ in range(n_segs):
        x = rng.choice([-1, 1]) * rng.uniform(1.0, 12.0); y = rng.uniform(-4, 4)
        z0 = rng.uniform(1.0, 30.0); z1 = z0 + rng.uniform(15, 60)
        g = float(90 + rng.choice([-1, 1]) * rng.uniform(24, 32))
        segs.append((np.array([x, y, z0]), np.array([x + rng.normal(0, 0.05), y + rng.normal(0, 0.05), z1]), g, int(rng.integers(2, 4))))
    world.segs = segs
    return world


def frames(kind):
    if kind == "kitti":
        return [f[0] for f in synth.stream(plf.KITTI_CAMERA, 2, world=synth.World(seed=7), seed=42)]
    if kind == "euroc":
        world = synth.World(seed=8, length=40.0, n_quads=220, n_segs=120, half_width=5.0, half_height=3.0)
        return [f[1] for f in synth.stream(plf.EUROC_CAMERA, 2, world=world, seed=43, step=0.08, yaw_deg=0.8)]
    return [f[0] for f in synth.stream(plf.KITTI_CAMERA, 2, world=corridor_world(), seed=17, noise=2)]


@pytest.mark.parametrize("kind,nf", [("kitti", 1500), ("euroc", 1200), ("corridor", 150)])
def test_orb_oracle_vs_cv2_on_stream_frames(kind, nf):
    pytest.importorskip("cv2")
    for img in frames(kind):
        kp, desc = clib.orb(img, nf)
        rk, rd = orb_cv2(img, nfeatures=nf)
        assert len(kp) == len(rk) and all(np.array_equal(kp[f], rk[f]) for f in FIELDS)
        assert np.array_equal(desc, rd)


@pytest.mark.parametrize("kind", ["kitti", "euroc", "corridor"])
def test_lsd_oracle_vs_cv2_on_stream_frames(kind):
    pytest.importorskip("cv2")
    for img in frames(kind):
        ref = lsd_cv2(img)
        mine = clib.lsd(img)
        assert len(ref) > 50 and mine.shape == ref.shape and np.array_equal(mine, ref)
