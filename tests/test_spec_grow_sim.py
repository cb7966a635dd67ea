"""Round-2 groundwork: the speculative parallel region growing with in-order commit (oracle/tools/spec_grow_sim.py)
reproduces the sequential region growing exactly (CPU simulation on a small crop of the bench frame)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle" / "tools"))


def test_speculative_growing_equals_sequential():
    pytest.importorskip("cv2")
    import spec_grow_sim as sim
    import plslam_b200 as plf
    from plslam_b200 import synth
    L = next(iter(synth.stream(plf.KITTI_CAMERA, 1, world=synth.World(seed=7), seed=42)))[0]
    ang, order = sim.level_lines(np.ascontiguousarray(L[100:260, 300:560]))
    ang = ang.tolist()
    ref, seq_steps = sim.sequential(ang, order)
    assert len(ref) > 100 and seq_steps == len(order)          # every defined pixel ends up in exactly one region
    for M in (4, 16, 64):
        out, st = sim.speculative(ang, order, M)
        assert out == ref                                        # same regions, same pixels, same acceptance order
        assert st["ticks"] < seq_steps                           # and a shorter critical path
