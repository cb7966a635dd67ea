"""Oracle compositions of the per-image / per-pair front-end stages (numpy + the C restatements).
TEST / BENCH INFRASTRUCTURE ONLY.

detect_lines   stvo-pl StereoFrame::detectLineFeatures, LSD branch ([UPSTREAM-RECALL],
               unpinned): LSDDetectorC::detect (LSDDetector_custom.cpp:218-324) -> if more than lsd_nfeatures lines
               (and lsd_nfeatures != 0): sort by response descending (std::sort ties are implementation-defined in
               the reference; canonical here: detection order), keep lsd_nfeatures, class_id = rank ->
               BinaryDescriptor::compute (binary_descriptor_custom.cpp:524).
"""
import numpy as np

from oracle import clib


def detect_lines(img, lsd_nfeatures=300, min_line_length=0.025, lsd_kwargs=None, segs=None):
    h, w = img.shape
    if segs is None:
        segs = clib.lsd(img, **(lsd_kwargs or {}))
    kl = clib.keylines_from_segments(segs, w, h, float(np.float32(min_line_length)) * min(w, h))
    if lsd_nfeatures != 0 and len(kl) > lsd_nfeatures:
        order = np.lexsort((np.arange(len(kl)), -kl["response"].astype(np.float64)))
        kl = kl[order[:lsd_nfeatures]].copy()
        kl["class_id"] = np.arange(lsd_nfeatures, dtype=np.int32)
    desc = clib.lbd_compute(img, kl) if len(kl) else np.zeros((0, 32), np.uint8)
    return kl, desc
