// Device port of glibc's single-precision sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h, glibc >= 2.28;
// verified against the libm.so.6 of this image, glibc 2.39, in tests/test_lsd_gpu.py).
//
// Why: OpenCV's LSD updates the region angle with `cos(float(angle))` / `sin(float(angle))`, which resolve to the
// host libm's cosf/sinf.  CUDA's cosf/sinf are not bit-identical to glibc's, so the region-growing kernel carries
// this port: double-precision range reduction by pi/2 and the degree-8/7 minimax polynomials with glibc's published
// coefficients, rounded once to float.  Multiply-adds are fused the way the x86-64 FMA build of glibc fuses them;
// since everything is evaluated in double and rounded once, fused vs unfused changes the float result with
// probability ~1e-9.  Valid for |y| < 120 (the LSD angles lie in [0, 2*pi)).
#pragma once

__device__ __forceinline__ float glibc_sinf_poly(double x, double x2, bool neg_tab, int n) {
  // table[0]: c0..c4 = 1, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16
  //           s1..s3 = -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13
  // table[1] = same with the cosine coefficients negated
  const double sg = neg_tab ? -1.0 : 1.0;
  if ((n & 1) == 0) {
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double x3 = x * x2;
    const double t1 = __fma_rn(x2, s3, s2);
    const double x7 = x3 * x2;
    const double s = __fma_rn(x3, s1, x);
    return (float)__fma_rn(x7, t1, s);
  } else {
    const double c0 = sg * 1.0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5,
                 c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = x2 * x2;
    const double t2 = __fma_rn(x2, c4, c3);
    const double t1 = __fma_rn(x2, c1, c0);
    const double x6 = x4 * x2;
    const double c = __fma_rn(x4, c2, t1);
    return (float)__fma_rn(x6, t2, c);
  }
}

// want_cos = false: sinf(y); true: cosf(y)
__device__ __forceinline__ float glibc_sincosf1(float y, bool want_cos) {
  double x = (double)y;
  const float ay = fabsf(y);
  if (ay < 0x1.921fb6p-1f) {  // |y| < pi/4 (abstop12 comparison; boundary cases coincide for our inputs)
    if (ay < 0x1p-12f) return want_cos ? 1.0f : y;
    return glibc_sinf_poly(x, x * x, false, want_cos ? 1 : 0);
  }
  // reduce_fast: n = round(x * 2/pi) via the 2^24-scaled truncation trick
  const double r = x * 0x1.45F306DC9C883p+23;
  const int n = ((int)r + 0x800000) >> 24;
  x = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
  const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign[] = {1,-1,-1,1}
  return glibc_sinf_poly(x * sign, x * x, (n & 2) != 0, want_cos ? (n ^ 1) : n);
}

__device__ __forceinline__ float glibc_sinf(float y) { return glibc_sincosf1(y, false); }
__device__ __forceinline__ float glibc_cosf(float y) { return glibc_sincosf1(y, true); }
