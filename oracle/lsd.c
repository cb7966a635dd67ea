/*
 * oracle/lsd.c — CPU restatement of OpenCV's LineSegmentDetector (LSD) as the reference calls it.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * The reference calls cv::createLineSegmentDetector(refine, scale, sigma_scale, quant, ang_th, log_eps,
 * density_th, n_bins)->detect(img, lines) from LSDDetectorC::detectImpl
 * (3rdparty/line_descriptor/src/LSDDetector_custom.cpp:246-264) with refine = 0 = LSD_REFINE_NONE
 * (config/config/config_euroc.yaml:69): Gaussian pre-blur + INTER_LINEAR_EXACT resample to `scale`,
 * level-line angles + gradient magnitude, 1024-bin pseudo-ordering of seeds, region growing, rectangle
 * fit.  With refine = 0 no NFA validation and no refinement runs.  OpenCV's lsd.cpp is a third-party
 * dependency that is not vendored under /root/reference; this file restates its published algorithm
 * (von Gioi et al., "LSD: a Line Segment Detector", IPOL 2012, as implemented in OpenCV imgproc) and is
 * PINNED against python cv2 4.13 in tests/test_lsd_oracle.py.
 *
 * Seed order.  OpenCV 4.x sorts the (bin, pixel) list with std::sort, which is not stable: the order of
 * pixels inside one magnitude bin is whatever libstdc++'s introsort produces.  OpenCV 3.x (the version
 * the reference pins, CMakeLists.txt:7) used per-bin lists in raster order.  Both are implemented:
 *   order_mode 0  raster order inside a bin (stable counting sort)  <- oracle of record for the CUDA path
 *   order_mode 1  emulation of libstdc++ std::sort on the raster-ordered list <- used only to pin this
 *                 restatement bit-for-bit against cv2 4.13
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define NOTDEF (-1024.0)
#define LSD_PI 3.1415926535897932384626433832795
#define M_3_2_PI ((3 * LSD_PI) / 2)
#define M_2__PI (2 * LSD_PI)
#define DEG_TO_RADS (LSD_PI / 180)

typedef struct { int x, y, norm; } norm_point;

/* ---- libstdc++ std::sort(first, last, comp) with comp(a,b) = a.norm > b.norm ---------------------- */
#define COMP(a, b) ((a).norm > (b).norm)
static void np_swap(norm_point* a, norm_point* b) { norm_point t = *a; *a = *b; *b = t; }

static void adjust_heap(norm_point* first, long hole, long len, norm_point value) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (COMP(first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  long parent = (hole - 1) / 2;
  while (hole > top && COMP(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void heap_sort(norm_point* first, norm_point* last) {
  long len = last - first;
  if (len >= 2)
    for (long parent = (len - 2) / 2;; parent--) {
      adjust_heap(first, parent, len, first[parent]);
      if (parent == 0) break;
    }
  while (last - first > 1) {
    --last;
    norm_point value = *last;
    *last = *first;
    adjust_heap(first, 0, last - first, value);
  }
}
static void introsort_loop(norm_point* first, norm_point* last, long depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { heap_sort(first, last); return; }
    --depth_limit;
    norm_point* mid = first + (last - first) / 2;
    norm_point *a = first + 1, *b = mid, *c = last - 1;
    if (COMP(*a, *b)) {
      if (COMP(*b, *c)) np_swap(first, b);
      else if (COMP(*a, *c)) np_swap(first, c);
      else np_swap(first, a);
    } else if (COMP(*a, *c)) np_swap(first, a);
    else if (COMP(*b, *c)) np_swap(first, c);
    else np_swap(first, b);
    norm_point *lo = first + 1, *hi = last;
    for (;;) {
      while (COMP(*lo, *first)) ++lo;
      --hi;
      while (COMP(*first, *hi)) --hi;
      if (!(lo < hi)) break;
      np_swap(lo, hi);
      ++lo;
    }
    introsort_loop(lo, last, depth_limit);
    last = lo;
  }
}
static void unguarded_linear_insert(norm_point* last) {
  norm_point val = *last;
  norm_point* next = last - 1;
  while (COMP(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void insertion_sort(norm_point* first, norm_point* last) {
  if (first == last) return;
  for (norm_point* i = first + 1; i != last; ++i) {
    if (COMP(*i, *first)) {
      norm_point val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(norm_point));
      *first = val;
    } else unguarded_linear_insert(i);
  }
}
static void libstdcxx_sort(norm_point* first, norm_point* last) {
  if (first == last) return;
  long n = last - first, lg = 0;
  while ((n >> (lg + 1)) > 0) lg++;
  introsort_loop(first, last, lg * 2);
  if (last - first > 16) {
    insertion_sort(first, first + 16);
    for (norm_point* i = first + 16; i != last; ++i) unguarded_linear_insert(i);
  } else insertion_sort(first, last);
}

/* ---- LSD ----------------------------------------------------------------------------------------------- */
typedef struct { int x, y; double modgrad; } region_point;

typedef struct {
  int w, h;
  double* angles;   /* level-line angle (radians, [0,2pi)) or NOTDEF */
  double* modgrad;
  float* cosa;      /* per-pixel cos/sin of float(angle) used by the region-angle update */
  float* sina;
  uint8_t* used;
} lsd_maps;

static inline int is_aligned(const lsd_maps* m, int x, int y, double theta, double prec) {
  if (x < 0 || y < 0 || x >= m->w || y >= m->h) return 0;
  const double a = m->angles[(size_t)y * m->w + x];
  if (a == NOTDEF) return 0;
  double n_theta = theta - a;
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > M_3_2_PI) {
    n_theta -= M_2__PI;
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

static inline double angle_diff(double a, double b) {
  double diff = a - b;
  while (diff <= -LSD_PI) diff += M_2__PI;
  while (diff > LSD_PI) diff -= M_2__PI;
  if (diff < 0.0) diff = -diff;
  return diff;
}

/* Computes the scaled image exactly as flsd() does.  Returns malloc'ed buffer; *ow,*oh = its size. */
uint8_t* orc_lsd_scaled_image(const uint8_t* img, int w, int h, double scale, double sigma_scale, int* ow, int* oh) {
  if (scale == 1.0) {
    uint8_t* out = (uint8_t*)malloc((size_t)w * h);
    memcpy(out, img, (size_t)w * h);
    *ow = w; *oh = h;
    return out;
  }
  const double sigma = (scale < 1) ? (sigma_scale / scale) : sigma_scale;
  const double sprec = 3;
  const unsigned int hk = (unsigned int)(ceil(sigma * sqrt(2 * sprec * log(10.0))));
  const int ksize = 1 + 2 * (int)hk;
  uint8_t* blur = (uint8_t*)malloc((size_t)w * h);
  orc_gaussian_blur_u8(img, w, h, ksize, sigma, blur);
  const int dw = (int)nearbyint(w * scale), dh = (int)nearbyint(h * scale);
  uint8_t* out = (uint8_t*)malloc((size_t)dw * dh);
  orc_resize_linear_exact(blur, w, h, out, dw, dh, scale, scale);
  free(blur);
  *ow = dw; *oh = dh;
  return out;
}

/* flsd with LSD_REFINE_NONE. segs: cap x 4 floats. Returns count or -1 on overflow.
 * trig_mode: how `cos(float(angle))` in region_grow resolves (0: float cosf, 1: double cos of the float) */
int orc_lsd_detect(const uint8_t* img, int w, int h, double scale, double sigma_scale, double quant, double ang_th,
                   int n_bins, int order_mode, int trig_mode, float* segs, int cap) {
  int W, H;
  uint8_t* simg = orc_lsd_scaled_image(img, w, h, scale, sigma_scale, &W, &H);
  const double prec = LSD_PI * ang_th / 180;
  const double p = ang_th / 180;
  const double rho = quant / sin(prec);
  const size_t N = (size_t)W * H;
  lsd_maps m;
  m.w = W; m.h = H;
  m.angles = (double*)malloc(sizeof(double) * N);
  m.modgrad = (double*)malloc(sizeof(double) * N);
  m.cosa = (float*)malloc(sizeof(float) * N);
  m.sina = (float*)malloc(sizeof(float) * N);
  m.used = (uint8_t*)calloc(N, 1);
  /* ll_angle */
  for (int x = 0; x < W; x++) m.angles[(size_t)(H - 1) * W + x] = NOTDEF;
  for (int y = 0; y < H; y++) m.angles[(size_t)y * W + W - 1] = NOTDEF;
  for (size_t i = 0; i < N; i++) m.modgrad[i] = 0; /* last row/col are never read as seeds */
  double max_grad = -1;
  for (int y = 0; y < H - 1; y++) {
    const uint8_t* r0 = simg + (size_t)y * W;
    const uint8_t* r1 = r0 + W;
    for (int x = 0; x < W - 1; x++) {
      int DA = r1[x + 1] - r0[x];
      int BC = r0[x + 1] - r1[x];
      int gx = DA + BC, gy = DA - BC;
      double norm = sqrt((gx * gx + gy * gy) / 4.0);
      m.modgrad[(size_t)y * W + x] = norm;
      if (norm <= rho) {
        m.angles[(size_t)y * W + x] = NOTDEF;
      } else {
        double a = orc_fast_atan2((float)gx, (float)(-gy)) * DEG_TO_RADS;
        m.angles[(size_t)y * W + x] = a;
        if (trig_mode == 0) {
          m.cosa[(size_t)y * W + x] = cosf((float)a);
          m.sina[(size_t)y * W + x] = sinf((float)a);
        }
        if (norm > max_grad) max_grad = norm;
      }
    }
  }
  /* pseudo-ordering */
  const double bin_coef = (max_grad > 0) ? (double)(n_bins - 1) / max_grad : 0;
  const size_t NP = (size_t)(W - 1) * (H - 1);
  norm_point* ordered = (norm_point*)malloc(sizeof(norm_point) * (NP + 1));
  {
    size_t k = 0;
    for (int y = 0; y < H - 1; y++)
      for (int x = 0; x < W - 1; x++) {
        ordered[k].x = x; ordered[k].y = y;
        ordered[k].norm = (int)(m.modgrad[(size_t)y * W + x] * bin_coef);
        k++;
      }
  }
  if (order_mode == 1) {
    libstdcxx_sort(ordered, ordered + NP);
  } else {
    /* stable: bins descending, raster order inside a bin */
    size_t* start = (size_t*)calloc((size_t)n_bins + 1, sizeof(size_t));
    for (size_t i = 0; i < NP; i++) start[n_bins - 1 - ordered[i].norm + 1]++;
    for (int b = 0; b < n_bins; b++) start[b + 1] += start[b];
    norm_point* tmp = (norm_point*)malloc(sizeof(norm_point) * (NP + 1));
    for (size_t i = 0; i < NP; i++) tmp[start[n_bins - 1 - ordered[i].norm]++] = ordered[i];
    free(ordered); free(start);
    ordered = tmp;
  }
  const double LOG_NT = 5 * (log10((double)W) + log10((double)H)) / 2 + log10(11.0);
  const size_t min_reg_size = (size_t)(-LOG_NT / log10(p));
  region_point* reg = (region_point*)malloc(sizeof(region_point) * N);
  int nseg = 0, overflow = 0;
#ifdef LSD_STATS
  long st_regions = 0, st_points = 0, st_small = 0, st_small_pts = 0, st_defined = 0;
  for (size_t i = 0; i < N; i++) st_defined += m.angles[i] != NOTDEF;
#endif
  for (size_t i = 0; i < NP && !overflow; i++) {
    const int sx = ordered[i].x, sy = ordered[i].y;
    const size_t si = (size_t)sy * W + sx;
    if (m.used[si] || m.angles[si] == NOTDEF) continue;
    /* region_grow */
    size_t nreg = 0;
    double reg_angle = m.angles[si];
    reg[nreg].x = sx; reg[nreg].y = sy; reg[nreg].modgrad = m.modgrad[si]; nreg++;
    float sumdx = (float)cos(reg_angle);
    float sumdy = (float)sin(reg_angle);
    m.used[si] = 1;
    for (size_t r = 0; r < nreg; r++) {
      const int rx = reg[r].x, ry = reg[r].y;
      const int xx_min = rx - 1 > 0 ? rx - 1 : 0, xx_max = rx + 1 < W - 1 ? rx + 1 : W - 1;
      const int yy_min = ry - 1 > 0 ? ry - 1 : 0, yy_max = ry + 1 < H - 1 ? ry + 1 : H - 1;
      for (int yy = yy_min; yy <= yy_max; ++yy)
        for (int xx = xx_min; xx <= xx_max; ++xx) {
          const size_t pi = (size_t)yy * W + xx;
          if (m.used[pi] != 1 && is_aligned(&m, xx, yy, reg_angle, prec)) {
            m.used[pi] = 1;
            reg[nreg].x = xx; reg[nreg].y = yy; reg[nreg].modgrad = m.modgrad[pi]; nreg++;
            if (trig_mode == 0) {
              sumdx += m.cosa[pi];
              sumdy += m.sina[pi];
            } else {
              const double af = (double)(float)m.angles[pi];
              sumdx += cos(af); /* float += double: promoted, then narrowed */
              sumdy += sin(af);
            }
            reg_angle = orc_fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
          }
        }
    }
#ifdef LSD_STATS
    st_regions++; st_points += nreg;
    if (nreg < min_reg_size) { st_small++; st_small_pts += nreg; }
#endif
    if (nreg < min_reg_size) continue;
    /* region2rect */
    double x = 0, y = 0, sum = 0;
    for (size_t k = 0; k < nreg; k++) {
      const double weight = reg[k].modgrad;
      x += (double)reg[k].x * weight;
      y += (double)reg[k].y * weight;
      sum += weight;
    }
    x /= sum;
    y /= sum;
    /* get_theta */
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (size_t k = 0; k < nreg; k++) {
      const double weight = reg[k].modgrad;
      const double dx = (double)reg[k].x - x, dy = (double)reg[k].y - y;
      Ixx += dy * dy * weight;
      Iyy += dx * dx * weight;
      Ixy -= dx * dy * weight;
    }
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)orc_fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                           : (double)orc_fast_atan2((float)Ixy, (float)(lambda - Iyy));
    theta *= DEG_TO_RADS;
    if (angle_diff(theta, reg_angle) > prec) theta += LSD_PI;
    const double dx = cos(theta), dy = sin(theta);
    double l_min = 0, l_max = 0;
    for (size_t k = 0; k < nreg; k++) {
      const double regdx = (double)reg[k].x - x, regdy = (double)reg[k].y - y;
      const double l = regdx * dx + regdy * dy;
      if (l > l_max) l_max = l;
      else if (l < l_min) l_min = l;
    }
    double x1 = x + l_min * dx, y1 = y + l_min * dy, x2 = x + l_max * dx, y2 = y + l_max * dy;
    x1 += 0.5; y1 += 0.5; x2 += 0.5; y2 += 0.5;
    if (scale != 1) { x1 /= scale; y1 /= scale; x2 /= scale; y2 /= scale; }
    if (nseg >= cap) { overflow = 1; break; }
    segs[4 * nseg] = (float)x1; segs[4 * nseg + 1] = (float)y1; segs[4 * nseg + 2] = (float)x2; segs[4 * nseg + 3] = (float)y2;
    nseg++;
  }
#ifdef LSD_STATS
  fprintf(stderr, "LSD_STATS %dx%d defined=%ld regions=%ld points=%ld small_regions=%ld small_points=%ld kept=%d\n", W, H,
          st_defined, st_regions, st_points, st_small, st_small_pts, nseg);
#endif
  free(reg); free(ordered); free(m.angles); free(m.modgrad); free(m.cosa); free(m.sina); free(m.used); free(simg);
  return overflow ? -1 : nseg;
}
