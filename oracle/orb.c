/*
 * oracle/orb.c — CPU restatement of OpenCV's ORB (detectAndCompute, FAST score) as the reference calls it.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * The reference reaches ORB through stvo-pl StereoFrame::detectPointFeatures ->
 * cv::ORB::create(nfeatures, scaleFactor, nlevels, edgeTh, 0, wtaK, scoreType, patchSize, fastTh)
 *   ->detectAndCompute(img, Mat(), kps, desc)      (parameters config/config/config_euroc.yaml:59-67;
 * descriptor rows consumed at src/mapHandler.cpp:86-88,302).  OpenCV is a third-party dependency that is
 * not vendored under /root/reference; its published algorithm (modules/features2d/src/orb.cpp, fast.cpp,
 * fast_score.cpp; imgproc resize INTER_LINEAR_EXACT and the float sepFilter2D Gaussian) is restated
 * here and PINNED bit-exact against python cv2 4.13 in tests/test_orb_oracle.py (keypoints, angles,
 * responses and descriptors on seeded images).  cv2 itself remains the oracle of record for ORB.
 *
 * Keypoint order: OpenCV's order inside a pyramid level is whatever std::nth_element + std::partition
 * leave behind (KeyPointsFilter::retainBest) — implementation-defined.  We define the canonical order
 * (octave, y, x) and compare against cv2 after sorting cv2's output the same way.
 *
 * Restricted to what the front-end uses: WTA_K = 2, FAST score (orb_score = 1), firstLevel = 0, no mask.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "orb_pattern.h"

#define ORB_MAX_LEVELS 8

static inline int cv_round_f(float v) { return (int)nearbyintf(v); }
static inline int cv_round_d(double v) { return (int)nearbyint(v); }

/* ---- cv::resize(INTER_LINEAR_EXACT), CV_8UC1 --------------------------------------------------------
 * Q8.8 coefficients (round half even of the fractional source coordinate), horizontal pass in Q8.8,
 * vertical in Q16.16, one rounding. scale_x/scale_y = source step per destination pixel
 * (1 / inv_scale; when dsize is given, inv_scale = dsize / ssize). */
static void linear_coeffs(int srcsize, int dstsize, double scale, int* ofs, int* c1, int* minofs, int* maxofs) {
  int mn = 0, mx = dstsize;
  for (int v = 0; v < dstsize; v++) {
    double fval = scale * ((double)v + 0.5) - 0.5;
    int ival = (int)floor(fval);
    ofs[v] = 0;
    c1[v] = 0;
    if (ival >= 0 && srcsize > 1) {
      if (ival < srcsize - 1) {
        ofs[v] = ival;
        c1[v] = cv_round_d((fval - ival) * 256.0);
      } else {
        ofs[v] = srcsize - 1;
        if (v < mx) mx = v;
      }
    } else if (v + 1 > mn) {
      mn = v + 1;
    }
  }
  *minofs = mn;
  *maxofs = mx;
}

void orc_resize_linear_exact(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, double inv_scale_x,
                             double inv_scale_y) {
  int* ox = (int*)malloc(sizeof(int) * dw * 2);
  int* cx = ox + dw;
  int* oy = (int*)malloc(sizeof(int) * dh * 2);
  int* cy = oy + dh;
  int minx, maxx, miny, maxy;
  linear_coeffs(sw, dw, 1.0 / inv_scale_x, ox, cx, &minx, &maxx);
  linear_coeffs(sh, dh, 1.0 / inv_scale_y, oy, cy, &miny, &maxy);
  uint16_t* H = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)sh * dw);
  for (int y = 0; y < sh; y++) {
    const uint8_t* s = src + (size_t)y * sw;
    uint16_t* h = H + (size_t)y * dw;
    for (int x = 0; x < dw; x++) {
      if (x < minx) h[x] = (uint16_t)(s[0] * 256);
      else if (x >= maxx) h[x] = (uint16_t)(s[sw - 1] * 256);
      else h[x] = (uint16_t)(s[ox[x]] * (256 - cx[x]) + s[ox[x] + 1] * cx[x]);
    }
  }
  for (int y = 0; y < dh; y++) {
    uint8_t* d = dst + (size_t)y * dw;
    const uint16_t *h0, *h1;
    uint32_t w0, w1;
    if (y < miny) { h0 = h1 = H; w0 = 256; w1 = 0; }
    else if (y >= maxy) { h0 = h1 = H + (size_t)(sh - 1) * dw; w0 = 256; w1 = 0; }
    else { h0 = H + (size_t)oy[y] * dw; h1 = h0 + dw; w0 = 256 - cy[y]; w1 = cy[y]; }
    for (int x = 0; x < dw; x++) {
      uint32_t v = (h0[x] * w0 + h1[x] * w1 + 32768u) >> 16;
      d[x] = (uint8_t)(v > 255 ? 255 : v);
    }
  }
  free(ox); free(oy); free(H);
}

/* ---- FAST-9/16 corner score (fast_score.cpp cornerScore<16>) ----------------------------------------
 * = (max over the 16 contiguous 9-arcs of the minimum |centre - arc pixel| on the brighter or darker
 * side) - 1; 0 when the pixel is not a corner for `threshold`. */
static const int fast_off[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                    {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

static int fast_score_px(const uint8_t* p, int stride, int threshold) {
  int d[32];
  const int v = p[0];
  for (int k = 0; k < 16; k++) d[k] = d[k + 16] = v - p[fast_off[k][1] * stride + fast_off[k][0]];
  int best = 0; /* best arc strength over both polarities */
  for (int s = 0; s < 16; s++) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; k++) {
      if (d[s + k] < mn) mn = d[s + k];
      if (d[s + k] > mx) mx = d[s + k];
    }
    if (mn > best) best = mn;   /* all 9 darker than centre by >= mn */
    if (-mx > best) best = -mx; /* all 9 brighter */
  }
  /* corner iff some arc has all |diff| > threshold  <=>  best > threshold; score = best - 1 */
  return best > threshold ? best - 1 : 0;
}

/* score map for one level: 0 outside [3,w-3) x [3,h-3) */
void orc_fast_score_map(const uint8_t* img, int w, int h, int threshold, uint8_t* score) {
  memset(score, 0, (size_t)w * h);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) score[(size_t)y * w + x] = (uint8_t)fast_score_px(img + (size_t)y * w + x, w, threshold);
}

/* cv::fastAtan2 (scalar path, mathfuncs_core): 7th-order polynomial, degrees. */
float orc_fast_atan2(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
  const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
  const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
  const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
  }
  return i;
}

/* The descriptor-stage blur: GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on a pyramid
 * ROI (orb.cpp detectAndCompute).  Because the ROI is a submatrix and BORDER_ISOLATED is not set, OpenCV does not take
 * its CV_8U fixed-point path here but sepFilter2D with the CV_32F Gaussian kernel: float row filter (sequential over
 * the 7 taps), float symmetric column filter, then saturate_cast<uchar> (round half even).  On FMA-capable x86 hosts
 * the AVX2 dispatch of filter.simd.hpp contracts every multiply-add; this restatement uses fmaf to match.  Pinned
 * pixel-exact against cv2.sepFilter2D and, end to end, against cv2 ORB descriptors (tests/test_orb_oracle.py). */
void orc_gaussian_kernel_f32(int ksize, double sigma, float* k) {
  double d[33], sum = 0;
  double scale2x = -0.5 / (sigma * sigma);
  for (int i = 0; i < ksize; i++) {
    double x = i - (ksize - 1) * 0.5;
    d[i] = exp(scale2x * x * x);
    sum += d[i];
  }
  sum = 1. / sum;
  for (int i = 0; i < ksize; i++) k[i] = (float)(d[i] * sum);
}

void orc_orb_blur7(const uint8_t* src, int w, int h, uint8_t* dst) {
  float k[7];
  orc_gaussian_kernel_f32(7, 2.0, k);
  float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const uint8_t* s = src + (size_t)y * w;
      float a = k[0] * (float)s[reflect101(x - 3, w)];
      for (int t = 1; t < 7; t++) a = fmaf(k[t], (float)s[reflect101(x + t - 3, w)], a);
      tmp[(size_t)y * w + x] = a;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float a = k[3] * tmp[(size_t)y * w + x];
      for (int t = 1; t <= 3; t++)
        a = fmaf(k[3 + t], tmp[(size_t)reflect101(y + t, h) * w + x] + tmp[(size_t)reflect101(y - t, h) * w + x], a);
      int v = (int)nearbyintf(a);
      dst[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
  free(tmp);
}

typedef struct { int x, y, score; } raw_kp;

/* Full detectAndCompute.  Outputs in canonical order (octave, y, x). Returns the number of keypoints
 * (<= cap) or -1 if cap is too small. */
int orc_orb_detect_and_compute(const uint8_t* image, int w, int h, int nfeatures, float scaleFactorF, int nlevels,
                               int edgeThreshold, int patchSize, int fastThreshold, orc_keypoint* kps,
                               uint8_t* desc, int cap) {
  if (nlevels > ORB_MAX_LEVELS) nlevels = ORB_MAX_LEVELS;
  const double scaleFactor = (double)scaleFactorF; /* ORB::create takes float, stores double */
  float layerScale[ORB_MAX_LEVELS];
  int lw[ORB_MAX_LEVELS], lh[ORB_MAX_LEVELS];
  uint8_t* lvl[ORB_MAX_LEVELS];
  for (int l = 0; l < nlevels; l++) {
    float scale = (float)pow(scaleFactor, (double)l);
    layerScale[l] = scale;
    float inv_scale = 1.0f / scale;
    lw[l] = cv_round_f(w * inv_scale);
    lh[l] = cv_round_f(h * inv_scale);
    lvl[l] = (uint8_t*)malloc((size_t)lw[l] * lh[l]);
    if (l == 0) memcpy(lvl[0], image, (size_t)w * h);
    else
      orc_resize_linear_exact(lvl[l - 1], lw[l - 1], lh[l - 1], lvl[l], lw[l], lh[l], (double)lw[l] / lw[l - 1],
                              (double)lh[l] / lh[l - 1]);
  }
  /* features per level (orb.cpp computeKeyPoints) */
  int nfeaturesPerLevel[ORB_MAX_LEVELS];
  {
    float factor = (float)(1.0 / scaleFactor);
    float ndesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
      nfeaturesPerLevel[l] = cv_round_f(ndesired);
      sum += nfeaturesPerLevel[l];
      ndesired *= factor;
    }
    nfeaturesPerLevel[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
  }
  /* umax (circular patch row extents) */
  const int halfPatchSize = patchSize / 2;
  int umax[64];
  {
    int v, v0, vmax = (int)floor(halfPatchSize * sqrt(2.f) / 2 + 1);
    int vmin = (int)ceil(halfPatchSize * sqrt(2.f) / 2);
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt((double)halfPatchSize * halfPatchSize - v * v));
    for (v = halfPatchSize, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }
  int total = 0;
  for (int l = 0; l < nlevels; l++) {
    const int W = lw[l], H = lh[l];
    uint8_t* score = (uint8_t*)malloc((size_t)W * H);
    orc_fast_score_map(lvl[l], W, H, fastThreshold, score);
    /* NMS (strictly greater than the 8 neighbours) + border filter (runByImageBorder) */
    int hist[256];
    memset(hist, 0, sizeof hist);
    raw_kp* cand = (raw_kp*)malloc(sizeof(raw_kp) * (size_t)W * H / 4 + 16);
    int nc = 0;
    for (int y = 3; y < H - 3; y++)
      for (int x = 3; x < W - 3; x++) {
        const uint8_t* s = score + (size_t)y * W + x;
        int sc = s[0];
        if (!sc) continue;
        if (!(sc > s[-1] && sc > s[1] && sc > s[-W - 1] && sc > s[-W] && sc > s[-W + 1] && sc > s[W - 1] && sc > s[W] &&
              sc > s[W + 1]))
          continue;
        if (!(x >= edgeThreshold && x < W - edgeThreshold && y >= edgeThreshold && y < H - edgeThreshold)) continue;
        cand[nc].x = x; cand[nc].y = y; cand[nc].score = sc; nc++;
        hist[sc]++;
      }
    /* retainBest(n): keep everything with response >= the n-th largest response */
    int n = nfeaturesPerLevel[l], thr = 0;
    if (n < nc) {
      if (n == 0) thr = 256; /* retainBest(0) clears */
      else {
        int acc = 0;
        for (int s = 255; s >= 0; s--) {
          acc += hist[s];
          if (acc >= n) { thr = s; break; }
        }
      }
    }
    for (int i = 0; i < nc; i++) {
      if (cand[i].score < thr) continue;
      if (total >= cap) { free(score); free(cand); for (int k = 0; k < nlevels; k++) free(lvl[k]); return -1; }
      orc_keypoint* kp = &kps[total++];
      /* IC angle on the unblurred level */
      const uint8_t* center = lvl[l] + (size_t)cand[i].y * W + cand[i].x;
      int m_01 = 0, m_10 = 0;
      for (int u = -halfPatchSize; u <= halfPatchSize; ++u) m_10 += u * center[u];
      for (int v = 1; v <= halfPatchSize; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
          int val_plus = center[u + v * W], val_minus = center[u - v * W];
          v_sum += (val_plus - val_minus);
          m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
      }
      kp->angle = orc_fast_atan2((float)m_01, (float)m_10);
      kp->octave = l;
      kp->response = (float)cand[i].score;
      kp->size = patchSize * layerScale[l];
      kp->x = (float)cand[i].x * layerScale[l];
      kp->y = (float)cand[i].y * layerScale[l];
      kp->lx = cand[i].x;
      kp->ly = cand[i].y;
    }
    free(score);
    free(cand);
  }
  /* descriptors on the 7x7 sigma-2 blurred levels */
  int done = 0;
  for (int l = 0; l < nlevels; l++) {
    const int W = lw[l], H = lh[l];
    uint8_t* blur = (uint8_t*)malloc((size_t)W * H);
    /* GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on a pyramid ROI: a submatrix
     * without BORDER_ISOLATED does not take OpenCV's bit-exact path but the generic 8U separable filter, whose
     * taps are cvRound(getGaussianKernel(7, 2, CV_32F) * 256) = 18 34 49 55 49 34 18 (sum 257). */
    orc_orb_blur7(lvl[l], W, H, blur);
    for (; done < total && kps[done].octave == l; done++) {
      const orc_keypoint* kp = &kps[done];
      float scale = 1.f / layerScale[l];
      float angle = kp->angle;
      angle *= (float)(3.14159265358979323846 / 180.f);
      float a = (float)cos(angle), b = (float)sin(angle);
      const uint8_t* center = blur + (size_t)cv_round_f(kp->y * scale) * W + cv_round_f(kp->x * scale);
      uint8_t* d = desc + (size_t)done * 32;
      const int* pat = orb_bit_pattern_31;
      for (int i = 0; i < 32; i++, pat += 32) {
        int val = 0;
        for (int j = 0; j < 8; j++) {
          const int* q = pat + 4 * j;
          float x0 = q[0] * a - q[1] * b, y0 = q[0] * b + q[1] * a;
          float x1 = q[2] * a - q[3] * b, y1 = q[2] * b + q[3] * a;
          int t0 = center[cv_round_f(y0) * W + cv_round_f(x0)];
          int t1 = center[cv_round_f(y1) * W + cv_round_f(x1)];
          val |= (t0 < t1) << j;
        }
        d[i] = (uint8_t)val;
      }
    }
    free(blur);
  }
  for (int k = 0; k < nlevels; k++) free(lvl[k]);
  return total;
}
