// Port of glibc's single-precision atanf / atan2f (sysdeps/ieee754/flt-32/s_atanf.c, e_atan2f.c: the fdlibm float
// algorithm, as shipped in this image's glibc 2.39), written so that the same source compiles for the device (nvcc,
// --fmad=false) and for the host (gcc -ffp-contract=off).
//
// Why: compiling the reference's vendored LSDDetector_custom.cpp (oracle/ref_build) showed that the KeyLine angle
// `atan2(endY - startY, endX - startX)` on float operands (:286) resolves to atan2f in the C++ build, which is 1 ulp
// away from the narrowed f64 atan2 on ~10 % of the lines (what round 1 computed).  CUDA's
// atan2f is not glibc's, so bit-exactness needs this port.  It is verified bit-for-bit against libm on the host
// (tests/test_glibc_atan2f_port.py; 60 M inputs offline) and is what k_keylines (lsd.cu) calls; the oracle's
// orc_keylines_from_segments calls the host libm atan2f, and tests/test_refbin_pin.py requires every KeyLine field of
// the oracle to equal the reference's own compiled LSDDetector_custom.cpp.
#pragma once
#ifdef __CUDACC__
#define PLF_LIBM_FN __device__ __forceinline__
#define PLF_F2I(x) __float_as_int(x)
#define PLF_I2F(i) __int_as_float(i)
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define PLF_LIBM_FN static inline
static inline int32_t plf_f2i_(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
static inline float plf_i2f_(int32_t i) { float x; memcpy(&x, &i, 4); return x; }
#define PLF_F2I(x) plf_f2i_(x)
#define PLF_I2F(i) plf_i2f_(i)
#endif

PLF_LIBM_FN float glibc_atanf(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                        9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                        4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
  const float one = 1.0f, huge = 1.0e30f;
  float w, s1, s2, z;
  int id;
  const int hx = PLF_F2I(x), ix = hx & 0x7fffffff;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return atanhi[3] + atanlo[3];
    return -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) {
      if (huge + x > one) return x;
    }
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {                                                   // |x| < 1.1875
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - one) / (2.0f + x); }  // 7/16 <= |x| < 11/16
      else { id = 1; x = (x - one) / (x + one); }                          // 11/16 <= |x| < 19/16
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (one + 1.5f * x); }  // |x| < 2.4375
      else { id = 3; x = -1.0f / x; }
    }
  }
  z = x * x;
  w = z * z;
  s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

PLF_LIBM_FN float glibc_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  float z;
  const int hx = PLF_F2I(x), ix = hx & 0x7fffffff, hy = PLF_F2I(y), iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;  // NaN
  if (hx == 0x3f800000) return glibc_atanf(y);            // x = 1
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);      // 2 * sign(x) + sign(y)
  if (iy == 0) {                                          // y = 0
    if (m < 2) return y;
    return m == 2 ? pi + tiny : -pi - tiny;
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;  // x = 0
  if (ix == 0x7f800000) {                                          // x = inf
    if (iy == 0x7f800000) {
      if (m == 0) return pi_o_4 + tiny;
      if (m == 1) return -pi_o_4 - tiny;
      return m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny;
    }
    if (m == 0) return 0.0f;
    if (m == 1) return -0.0f;
    return m == 2 ? pi + tiny : -pi - tiny;
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;  // y = inf
  const int k = (iy - ix) >> 23;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;                 // |y/x| > 2^60
  else if (hx < 0 && k < -60) z = 0.0f;                  // |y|/x < -2^60
  else z = glibc_atanf(fabsf(y / x));
  if (m == 0) return z;
  if (m == 1) return PLF_I2F(PLF_F2I(z) ^ (int)0x80000000);
  if (m == 2) return pi - (z - pi_lo);
  return (z - pi_lo) - pi;
}
