"""cv2 (python OpenCV 4.13) as the oracle of record for the third-party arithmetic the reference calls:
ORB, LSD, BFMatcher.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference links OpenCV 3 C++ (CMakeLists.txt:7); python cv2 4.13 is the only OpenCV in this image.  Version
skew is accepted and stated in DESIGN.md."""
import numpy as np

KP_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("size", np.float32), ("angle", np.float32),
                     ("response", np.float32), ("octave", np.int32), ("class_id", np.int32)])


def orb_cv2(img, nfeatures=800, scale_factor=1.2, nlevels=4, edge_th=19, wta_k=2, patch=31, fast_th=20):
    """cv::ORB::create(...)->detectAndCompute as stvo-pl calls it, returned in canonical (octave, y, x) order."""
    import cv2
    orb = cv2.ORB_create(int(nfeatures), float(scale_factor), int(nlevels), int(edge_th), 0, int(wta_k),
                         cv2.ORB_FAST_SCORE, int(patch), int(fast_th))
    kps, desc = orb.detectAndCompute(img, None)
    out = np.zeros(len(kps), KP_DTYPE)
    for i, k in enumerate(kps):
        out[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    if len(kps) == 0:
        return out, np.zeros((0, 32), np.uint8)
    order = np.lexsort((out["x"], out["y"], out["octave"]))
    return out[order], np.ascontiguousarray(desc[order])


def lsd_cv2(img, refine=0, scale=1.2, sigma_scale=0.6, quant=2.0, ang_th=22.5, log_eps=1.0, density_th=0.6, n_bins=1024):
    import cv2
    lsd = cv2.createLineSegmentDetector(int(refine), float(scale), float(sigma_scale), float(quant), float(ang_th),
                                        float(log_eps), float(density_th), int(n_bins))
    segs = lsd.detect(img)[0]
    return np.zeros((0, 4), np.float32) if segs is None else segs.reshape(-1, 4).astype(np.float32)
