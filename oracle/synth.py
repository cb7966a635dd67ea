"""The synthetic-input generator moved to the package (pl-slam_b200/plslam_b200/synth.py): it is data generation, not
part of the oracle.  Kept as an alias for the tests that import it from here."""
from plslam_b200.synth import *  # noqa: F401,F403
from plslam_b200.synth import World, expmap_se3, gn_problem, project, render_view, scene_pair, stream, trajectory  # noqa: F401
