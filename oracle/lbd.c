/*
 * oracle/lbd.c — CPU restatement of the vendored LBD line descriptor and its prelude.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into the product library.
 *
 * Follows, function by function (paths relative to the reference tree):
 *   orc_gaussian_blur_u8   cv::GaussianBlur on CV_8U as called at
 *                          3rdparty/line_descriptor/src/binary_descriptor_custom.cpp:358 (5x5, sigma 1),
 *                          OpenCV >= 4 fixed-point path (Q8.8 kernel, one rounding at the end);
 *                          pinned bit-exact against cv2 4.13 in tests/test_lbd_oracle.py.
 *   orc_sobel3_i16         cv::Sobel(CV_16S, ksize 3) as called at binary_descriptor_custom.cpp:395-396.
 *   orc_lbd_weights        BinaryDescriptor ctor, binary_descriptor_custom.cpp:217-259 (integer-division
 *                          centres, unnormalised Gaussians).
 *   orc_lbd_compute        computeLBD, binary_descriptor_custom.cpp:1026-1372, and the binary packing
 *                          binaryConversion :401-412 over combinations[32][2] :74-107 as driven by
 *                          computeImpl :645-685.
 *   orc_keylines_from_segments   LSDDetectorC::detectImpl, LSDDetector_custom.cpp:267-308 (+ checkLineExtremes
 *                          :76-102): clamp, min-length filter, KeyLine fill, LineIterator pixel count.
 *
 * Floating-point conventions mirrored (compile with -ffp-contract=off):
 *   - float expressions are evaluated in float in source order, no FMA contraction;
 *   - cos/sin/round resolve to the double C functions (the translation unit only includes
 *     <cmath>; the float argument is promoted) and the result is narrowed on assignment; the KeyLine
 *     angle's atan2(float, float) resolves to the float overload, i.e. glibc's atan2f;
 *   - sqrt resolves to std::sqrt(float) (cv namespace has `using std::sqrt`), so 1/sqrt(x) is a
 *     float division.
 * The reference itself is built with -O3 -march=native (CMakeLists.txt:27), i.e. its FMA behaviour
 * depends on the build host.
 * PINNED: the vendored sources themselves are compiled, unmodified, into oracle/_ref/liblinedesc_ref.so
 * (oracle/ref_build/, -ffp-contract=off) and tests/test_refbin_pin.py checks orc_lbd_compute bit-for-bit
 * against BinaryDescriptor::compute, and orc_keylines_from_segments field-for-field against
 * LSDDetectorC::detect - every field, including `angle` (glibc atan2f; the CUDA kernel carries a bit-exact
 * port of it, pl-slam_b200/csrc/glibc_atan2f.cuh).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
  }
  return i;
}

/* OpenCV getGaussianKernel (double) followed by the fixed-point error-diffusion conversion used by
 * the CV_8U bit-exact GaussianBlur (Q8.8, centre tap absorbs the remainder so taps sum to 256). */
void orc_gaussian_kernel_q8(int ksize, double sigma, int* taps) {
  double k[33];
  double sum = 0;
  if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
  double scale2x = -0.5 / (sigma * sigma);
  for (int i = 0; i < ksize; i++) {
    double x = i - (ksize - 1) * 0.5;
    k[i] = exp(scale2x * x * x);
    sum += k[i];
  }
  sum = 1. / sum;
  for (int i = 0; i < ksize; i++) k[i] *= sum;
  double err = 0;
  int s = 0;
  for (int i = 0; i < ksize / 2; i++) {
    double adj = k[i] * 256.0 + err;
    int v0 = (int)nearbyint(adj);
    err = adj - v0;
    taps[i] = taps[ksize - 1 - i] = v0;
    s += v0;
  }
  taps[ksize / 2] = 256 - 2 * s;
}

/* Separable integer filter with Q8 taps and one final rounding ((acc + 2^15) >> 16, saturated):
 * the common core of OpenCV's CV_8U bit-exact GaussianBlur and of its generic 8U separable filter
 * (createSeparableLinearFilter with bits = 8). BORDER_REFLECT_101. */
void orc_blur_taps_u8(const uint8_t* src, int w, int h, int ksize, const int* taps, uint8_t* dst) {
  const int r = ksize / 2;
  uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t a = 0;
      for (int t = 0; t < ksize; t++) a += (uint32_t)taps[t] * src[(size_t)y * w + reflect101(x + t - r, w)];
      tmp[(size_t)y * w + x] = a;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t a = 0;
      for (int t = 0; t < ksize; t++) a += (uint32_t)taps[t] * tmp[(size_t)reflect101(y + t - r, h) * w + x];
      uint32_t v = (a + (1u << 15)) >> 16;
      dst[(size_t)y * w + x] = (uint8_t)(v > 255 ? 255 : v);
    }
  free(tmp);
}

void orc_gaussian_blur_u8(const uint8_t* src, int w, int h, int ksize, double sigma, uint8_t* dst) {
  int taps[33];
  orc_gaussian_kernel_q8(ksize, sigma, taps);
  orc_blur_taps_u8(src, w, h, ksize, taps, dst);
}

void orc_sobel3_i16(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy) {
  for (int y = 0; y < h; y++) {
    const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * w;
    const uint8_t* r1 = src + (size_t)y * w;
    const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * w;
    for (int x = 0; x < w; x++) {
      int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      int gx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
      int gy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
      dx[(size_t)y * w + x] = (int16_t)gx;
      dy[(size_t)y * w + x] = (int16_t)gy;
    }
  }
}

#define NUM_OF_BANDS 9
#define WIDTH_OF_BAND 7

/* binary_descriptor_custom.cpp:217-259 */
void orc_lbd_weights(double* gaussCoefL /*21*/, double* gaussCoefG /*63*/) {
  double u = (WIDTH_OF_BAND * 3 - 1) / 2; /* integer division: 10 */
  double sigma = (WIDTH_OF_BAND * 2 + 1) / 2; /* 7 */
  double invsigma2 = -1 / (2 * sigma * sigma);
  for (int i = 0; i < WIDTH_OF_BAND * 3; i++) {
    double dis = i - u;
    gaussCoefL[i] = exp(dis * dis * invsigma2);
  }
  u = (NUM_OF_BANDS * WIDTH_OF_BAND - 1) / 2; /* 31 */
  sigma = u;
  invsigma2 = -1 / (2 * sigma * sigma);
  for (int i = 0; i < NUM_OF_BANDS * WIDTH_OF_BAND; i++) {
    double dis = i - u;
    gaussCoefG[i] = exp(dis * dis * invsigma2);
  }
}

/* binary_descriptor_custom.cpp:74-107 */
static const int combinations[32][2] = {
    {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
    {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
    {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

/* One line: binary_descriptor_custom.cpp:1088-1341.  des = 72 floats. */
static void lbd_one(const int16_t* pdx, const int16_t* pdy, int width, int height, const orc_keyline* kl,
                    const double* gaussCoefL, const double* gaussCoefG, float* des) {
  const short heightOfLSP = WIDTH_OF_BAND * NUM_OF_BANDS;
  const short halfHeight = (heightOfLSP - 1) / 2;
  const short realWidth = (short)width;
  const short imageWidth = realWidth - 1, imageHeight = (short)(height - 1);
  float band[8][NUM_OF_BANDS]; /* pgdL, ngdL, pgdL2, ngdL2, pgdO, ngdO, pgdO2, ngdO2 */
  memset(band, 0, sizeof band);
  float* pgdLBandSum = band[0]; float* ngdLBandSum = band[1];
  float* pgdL2BandSum = band[2]; float* ngdL2BandSum = band[3];
  float* pgdOBandSum = band[4]; float* ngdOBandSum = band[5];
  float* pgdO2BandSum = band[6]; float* ngdO2BandSum = band[7];

  const short lengthOfLSP = (short)kl->numOfPixels;
  const short halfWidth = (lengthOfLSP - 1) / 2;
  const float lineMiddlePointX = (float)(0.5 * (kl->sPointInOctaveX + kl->ePointInOctaveX));
  const float lineMiddlePointY = (float)(0.5 * (kl->sPointInOctaveY + kl->ePointInOctaveY));
  float dL[2], dO[2];
  dL[0] = (float)cos((double)kl->angle);
  dL[1] = (float)sin((double)kl->angle);
  dO[0] = -dL[1];
  dO[1] = dL[0];
  float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + lineMiddlePointX;
  float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + lineMiddlePointY;
  for (short hID = 0; hID < heightOfLSP; hID++) {
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
      short tempCor = (short)round((double)sCorX);
      short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
      tempCor = (short)round((double)sCorY);
      short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
      short dx = pdx[yCor * realWidth + xCor];
      short dy = pdy[yCor * realWidth + xCor];
      float gDL = dx * dL[0] + dy * dL[1];
      float gDO = dx * dO[0] + dy * dO[1];
      if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
      if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
      sCorX += dL[0];
      sCorY += dL[1];
    }
    sCorX0 -= dL[1];
    sCorY0 += dL[0];
    float coef = (float)gaussCoefG[hID];
    pgdLRowSum = coef * pgdLRowSum;
    ngdLRowSum = coef * ngdLRowSum;
    float pgdL2RowSum = pgdLRowSum * pgdLRowSum;
    float ngdL2RowSum = ngdLRowSum * ngdLRowSum;
    pgdORowSum = coef * pgdORowSum;
    ngdORowSum = coef * ngdORowSum;
    float pgdO2RowSum = pgdORowSum * pgdORowSum;
    float ngdO2RowSum = ngdORowSum * ngdORowSum;
    short bandID = (short)(hID / WIDTH_OF_BAND);
    for (int pass = 0; pass < 3; pass++) {
      /* pass 0: own band (weights [7..13]); pass 1: band above ([14..20]); pass 2: band below ([0..6]) */
      int b = pass == 0 ? bandID : pass == 1 ? bandID - 1 : bandID + 1;
      if (b < 0 || b >= NUM_OF_BANDS) continue;
      int wi = hID % WIDTH_OF_BAND + (pass == 0 ? WIDTH_OF_BAND : pass == 1 ? 2 * WIDTH_OF_BAND : 0);
      coef = (float)gaussCoefL[wi];
      pgdLBandSum[b] += coef * pgdLRowSum;
      ngdLBandSum[b] += coef * ngdLRowSum;
      pgdL2BandSum[b] += coef * coef * pgdL2RowSum;
      ngdL2BandSum[b] += coef * coef * ngdL2RowSum;
      pgdOBandSum[b] += coef * pgdORowSum;
      ngdOBandSum[b] += coef * ngdORowSum;
      pgdO2BandSum[b] += coef * coef * pgdO2RowSum;
      ngdO2BandSum[b] += coef * coef * ngdO2RowSum;
    }
  }
  const float invN2 = (float)(1.0 / (WIDTH_OF_BAND * 2.0));
  const float invN3 = (float)(1.0 / (WIDTH_OF_BAND * 3.0));
  for (int b = 0; b < NUM_OF_BANDS; b++) {
    float invN = (b == 0 || b == NUM_OF_BANDS - 1) ? invN2 : invN3;
    int d = b * 8;
    float temp = pgdLBandSum[b] * invN;
    des[d] = temp;
    des[d + 4] = sqrtf(pgdL2BandSum[b] * invN - temp * temp);
    temp = ngdLBandSum[b] * invN;
    des[d + 1] = temp;
    des[d + 5] = sqrtf(ngdL2BandSum[b] * invN - temp * temp);
    temp = pgdOBandSum[b] * invN;
    des[d + 2] = temp;
    des[d + 6] = sqrtf(pgdO2BandSum[b] * invN - temp * temp);
    temp = ngdOBandSum[b] * invN;
    des[d + 3] = temp;
    des[d + 7] = sqrtf(ngdO2BandSum[b] * invN - temp * temp);
  }
  float tempM = 0, tempS = 0;
  for (int b = 0; b < NUM_OF_BANDS; b++) {
    const float* v = des + 8 * b;
    tempM += v[0] * v[0]; tempM += v[1] * v[1]; tempM += v[2] * v[2]; tempM += v[3] * v[3];
    tempS += v[4] * v[4]; tempS += v[5] * v[5]; tempS += v[6] * v[6]; tempS += v[7] * v[7];
  }
  tempM = 1 / sqrtf(tempM);
  tempS = 1 / sqrtf(tempS);
  for (int b = 0; b < NUM_OF_BANDS; b++) {
    float* v = des + 8 * b;
    v[0] = v[0] * tempM; v[1] = v[1] * tempM; v[2] = v[2] * tempM; v[3] = v[3] * tempM;
    v[4] = v[4] * tempS; v[5] = v[5] * tempS; v[6] = v[6] * tempS; v[7] = v[7] * tempS;
  }
  for (int i = 0; i < 72; i++)
    if ((double)des[i] > 0.4) des[i] = (float)0.4;
  float temp = 0;
  for (int i = 0; i < 72; i++) temp += des[i] * des[i];
  temp = 1 / sqrtf(temp);
  for (int i = 0; i < 72; i++) des[i] = des[i] * temp;
}

/* computeImpl + computeLBD on one image, one octave.  desc_bin: n x 32 bytes; desc_float (optional): n x 72 */
void orc_lbd_compute(const uint8_t* img, int w, int h, const orc_keyline* kls, int n, uint8_t* desc_bin,
                     float* desc_float) {
  uint8_t* blur = (uint8_t*)malloc((size_t)w * h);
  int16_t* dx = (int16_t*)malloc(sizeof(int16_t) * (size_t)w * h);
  int16_t* dy = (int16_t*)malloc(sizeof(int16_t) * (size_t)w * h);
  orc_gaussian_blur_u8(img, w, h, 5, 1.0, blur);
  orc_sobel3_i16(blur, w, h, dx, dy);
  double gL[21], gG[63];
  orc_lbd_weights(gL, gG);
  for (int l = 0; l < n; l++) {
    float des[72];
    lbd_one(dx, dy, w, h, &kls[l], gL, gG, des);
    if (desc_float) memcpy(desc_float + (size_t)l * 72, des, sizeof des);
    for (int c = 0; c < 32; c++) {
      const float* f1 = des + 8 * combinations[c][0];
      const float* f2 = des + 8 * combinations[c][1];
      uint8_t r = 0;
      for (int i = 0; i < 8; i++)
        if (f1[i] > f2[i]) r += (uint8_t)(1 << i);
      desc_bin[(size_t)l * 32 + c] = r;
    }
  }
  free(blur); free(dx); free(dy);
}

/* cvRound: round half to even (SSE cvtss2si under the default rounding mode) */
static inline int cv_round_f(float v) { return (int)nearbyintf(v); }

/* LSDDetector_custom.cpp:267-308 with octave 0 only (numOctaves = 1, octaveScale = pow(scale,0) = 1).
 * segs: m x 4 floats (x1,y1,x2,y2) as returned by cv::LineSegmentDetector::detect.
 * Returns the number of KeyLines written (class_id = running counter of accepted lines). */
int orc_keylines_from_segments(const float* segs, int m, int w, int h, double min_length, orc_keyline* out) {
  int class_counter = -1, n = 0;
  for (int k = 0; k < m; k++) {
    float e[4] = {segs[4 * k], segs[4 * k + 1], segs[4 * k + 2], segs[4 * k + 3]};
    /* checkLineExtremes :76-102 */
    if (e[0] < 0) e[0] = 0;
    if (e[0] >= w) e[0] = (float)w - 1.0f;
    if (e[2] < 0) e[2] = 0;
    if (e[2] >= w) e[2] = (float)w - 1.0f;
    if (e[1] < 0) e[1] = 0;
    if (e[1] >= h) e[1] = (float)h - 1.0f;
    if (e[3] < 0) e[3] = 0;
    if (e[3] >= h) e[3] = (float)h - 1.0f;
    double length = (float)sqrt(pow((double)(e[0] - e[2]), 2) + pow((double)(e[1] - e[3]), 2));
    if (!(length > min_length)) continue;
    orc_keyline kl;
    const float octaveScale = 1.0f;
    kl.startPointX = e[0] * octaveScale; kl.startPointY = e[1] * octaveScale;
    kl.endPointX = e[2] * octaveScale; kl.endPointY = e[3] * octaveScale;
    kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1];
    kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
    kl.lineLength = (float)length;
    /* cv::LineIterator(img, Point2f, Point2f).count, 8-connected: Point2f -> Point via cvRound,
     * count = max(|dx|,|dy|) + 1 (both endpoints are inside the image after the clamp) */
    int x1 = cv_round_f(e[0]), y1 = cv_round_f(e[1]), x2 = cv_round_f(e[2]), y2 = cv_round_f(e[3]);
    int adx = abs(x2 - x1), ady = abs(y2 - y1);
    kl.numOfPixels = (adx > ady ? adx : ady) + 1;
    /* `atan2(float, float)` at :286 resolves to the float overload = glibc atan2f in the reference's C++ build
     * (established by compiling the vendored source, oracle/ref_build; tests/test_refbin_pin.py) */
    kl.angle = atan2f(kl.endPointY - kl.startPointY, kl.endPointX - kl.startPointX);
    kl.class_id = ++class_counter;
    kl.octave = 0;
    kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
    kl.response = kl.lineLength / (float)(w > h ? w : h);
    kl.ptx = (kl.endPointX + kl.startPointX) / 2;
    kl.pty = (kl.endPointY + kl.startPointY) / 2;
    out[n++] = kl;
  }
  return n;
}
