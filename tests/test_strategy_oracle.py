"""CPU checks of the oracle's windowed matching strategy (oracle/frontend.py, matching_strategy != 0): the restated
control flow of stvo-pl's stereo matchGrid() association and of the in-tree analogue's window-then-match() tracking
(src/mapHandler.cpp:247-278, :379-425)."""
import numpy as np

import plslam_b200 as plf
from oracle import clib, synth
from oracle import frontend as ofe
from oracle import matching as om


def _stream():
    cam = dict(plf.KITTI_CAMERA, width=480, height=300, cx=240.0, cy=150.0, fx=400.0, fy=400.0)
    world = synth.World(seed=3, length=40.0, n_quads=120, n_segs=60, half_width=6.0, half_height=3.0)
    return cam, list(synth.stream(cam, 3, world=world, seed=5, step=0.1))


def test_windowed_strategy_tracks_the_planted_motion():
    cam, frames = _stream()
    prm = dict(ofe.DEFAULTS, orb_nfeatures=500, lsd_nfeatures=100, matching_strategy=3)
    out = ofe.run_sequence(cam, [(a, b) for a, b, _ in frames], prm)
    out0 = ofe.run_sequence(cam, [(a, b) for a, b, _ in frames], dict(prm, matching_strategy=0))
    assert [o["status"] for o in out] == [2, 0, 0]
    # the epipolar window finds the same or more stereo pairs than the ratio test over the whole image
    assert all(o["n_pt"] >= o0["n_pt"] for o, o0 in zip(out, out0))
    T_true = np.linalg.inv(frames[0][2]) @ frames[-1][2]
    assert np.linalg.norm(out[-1]["Tfw"][:3, 3] - T_true[:3, 3]) < 0.08 * np.linalg.norm(T_true[:3, 3])


def test_f2f_window_falls_back_to_match():
    cam, frames = _stream()
    prm = dict(ofe.DEFAULTS, orb_nfeatures=500, lsd_nfeatures=100)
    f0 = ofe.extract_stereo(cam, frames[0][0], frames[0][1], prm)
    f1 = ofe.extract_stereo(cam, frames[1][0], frames[1][1], prm)
    brute_p, _ = om.match(f0.pdesc, f1.pdesc, prm["min_ratio_12_p"], True)
    brute_l, _ = om.match(f0.ldesc, f1.ldesc, prm["min_ratio_12_l"], True)
    # impossible thresholds: the window result is always "too few" -> match() for both kinds
    mp, ml = ofe.track_matches(cam, f0, f1, dict(prm, matching_strategy=3, min_pt_matches=min(len(f0.pdesc), len(f1.pdesc)) - 1,
                                                 min_ls_matches=min(len(f0.ldesc), len(f1.ldesc)) - 1))
    assert np.array_equal(mp, brute_p) and np.array_equal(ml, brute_l)
    # thresholds of 0: the window result always stands, and every pair it reports lies inside the window
    mp, ml = ofe.track_matches(cam, f0, f1, dict(prm, matching_strategy=3, min_pt_matches=0, min_ls_matches=0))
    iw, ih = ofe.grid_scales(cam)
    q = ofe.projection(cam, f0.pt_P)
    for i, j in enumerate(mp):
        if j >= 0:
            assert abs(int(q[i, 0] * iw) - int(f1.pt_pl[j, 0] * iw)) <= 3 and abs(int(q[i, 1] * ih) - int(f1.pt_pl[j, 1] * ih)) <= 3
    assert (mp >= 0).sum() > 20
