// Local bundle adjustment on the GPU.
//
// Replaces MapHandler::levMarquardtOptimizationLBA (src/mapHandler.cpp:1332-1989), the numerical core of
// MapHandler::localBundleAdjustment (:1220-1330): Levenberg-Marquardt over the local keyframes' poses (x_kf_w, se(3)
// vectors), the local 3-D points and the local 3-D line segments (two end points), with Cauchy-weighted scalar residuals
// (the norm of the point reprojection error / of the two end-point distances to the observed line).  The reference
// rebuilds a DENSE N x N Hessian every iteration and factors its sparse view with SimplicialLDLT; here the same normal
// equations are solved through the Schur complement on the landmarks - every observation couples ONE keyframe with ONE
// landmark, so the landmark block is block-diagonal (3x3 / 6x6) and each coupling block is rank one:
//   k_lba_rows      one thread per observation: residual, weight, the two Jacobian rows (:1384-1400 points, :1441-1503
//                   lines); accumulates the keyframe blocks A (6x6) / g_c and the error
//   k_lba_lm_accum  one thread per landmark: D = sum w Jx Jx^T, g_l (its observations are contiguous, in the reference's
//                   order), max |diagonal| for the initial lambda (:1543-1550)
//   k_lba_lm_schur  one thread per landmark: D' = D + lambda diag(D), D'^-1, and the reduced-system contributions
//                   S[a][b] -= w_b (y_a . Jx_b) Jt_a Jt_b^T, g_red[a] -= (y_a . g_l) Jt_a with y_a = w_a D'^-1 Jx_a
//   k_lba_solve     one CTA: S += A', dense LDL^T (no pivoting, as SimplicialLDLT) and the two triangular solves
//   k_lba_backsub   one thread per landmark: dx_l = D'^-1 (g_l - sum_a w_a Jx_a (Jt_a . dx_c[a]))
//   k_lba_apply     T <- T * inverse_se3(expmap_se3(dx)) / X += dx (:1558-1570), |dx|^2
// The host keeps the reference's control flow (first pass on the map's values, lambda schedule :1785-1803, stop tests
// :1775,:1805).  opts.ref_quirks selects between the reference as written and its evident intent for the four oddities
// documented in oracle/lba.c (division by the zero counters :1541, both line end points read from one block of X at
// stride 3 :1678-1679, the map's pose for line rows :1681, the literal clamp); oracle = oracle/lba.c.
#include "plf_internal.h"
#include "plf_se3.cuh"

struct LbaDev {
  int nkf, npt, nls, N, npo, nlo, nobs, nlm;
  double* X; double* X0; double* T0inv;       // N, N, nkf x 16 (inverse of the map pose of the local keyframes)
  const double* fixedTinv;                     // n_fixed x 16
  const int* obs_lm; const int* obs_kf;        // nobs (points first, then lines)
  const double* obs_z;                         // nobs x 3 (points: x, y, -; lines: the observed line equation)
  const int* lm_start;                         // nlm + 1 (points then lines): first observation of each landmark
  double* Jt; double* Jx; double* Wn;          // nobs x 6, nobs x 6, nobs x 2 {w, r w}
  double* D; double* gl; double* Dinv;         // nlm x 36, nlm x 6, nlm x 36
  double* A; double* gc;                       // nkf x 36, nkf x 6
  double* S; double* gred; double* dxc;        // n x n, n, n   (n = 6 nkf)
  double* DX;                                  // N
  double* scal;                                // [0] err, [1] |dx|^2, [2] Hmax (as ordered bits)
};
struct LbaCam { double fx, fy, cx, cy; };

__device__ __forceinline__ void lba_jac_pose(double gz2, double fxdx, double fydy, double gx, double gy, double gz, double* J) {
  J[0] = +gz2 * fxdx * gz;
  J[1] = +gz2 * fydy * gz;
  J[2] = -gz2 * (fxdx * gx + fydy * gy);
  J[3] = -gz2 * (fxdx * gx * gy + fydy * gy * gy + fydy * gz * gz);
  J[4] = +gz2 * (fxdx * gx * gx + fxdx * gz * gz + fydy * gx * gy);
  J[5] = +gz2 * (fydy * gx * gz - fxdx * gy * gz);
}
__device__ __forceinline__ void lba_rowvec_R(const double* v, const double* T, double* out) {
  for (int j = 0; j < 3; ++j) out[j] = v[0] * T[j] + v[1] * T[4 + j] + v[2] * T[8 + j];
}
__device__ __forceinline__ void lba_xform(const double* T, const double* P, double* Q) {
  for (int i = 0; i < 3; ++i) Q[i] = T[4 * i] * P[0] + T[4 * i + 1] * P[1] + T[4 * i + 2] * P[2] + T[4 * i + 3];
}

// mode 0: first pass on the map's values (X0 / T0inv); mode 1: LM iteration on X; quirks: see the header
__global__ void __launch_bounds__(128) k_lba_rows(LbaDev d, LbaCam c, int mode, int quirks, double homog_th) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  double e2w = 0.0;
  if (k < d.nobs) {
    const bool is_line = k >= d.npo;
    const int lm = d.obs_lm[k], kf = d.obs_kf[k];
    const double* X = mode == 0 ? d.X0 : d.X;
    double Ti[16];
    const bool map_pose = mode == 0 || (is_line && quirks);   // (q3)
    if (kf >= 0) {
      if (map_pose) {
        for (int i = 0; i < 16; ++i) Ti[i] = d.T0inv[16 * kf + i];
      } else {
        double T[16];
        d_expmap(X + 6 * kf, T);
        d_inverse_se3(T, Ti);
      }
    } else {
      for (int i = 0; i < 16; ++i) Ti[i] = d.fixedTinv[16 * (-1 - kf) + i];
    }
    double JT[6] = {0, 0, 0, 0, 0, 0}, JX[6] = {0, 0, 0, 0, 0, 0}, nrm;
    const double* z = d.obs_z + 3 * k;
    if (!is_line) {
      double Xwi[3];
      lba_xform(Ti, X + 6 * d.nkf + 3 * lm, Xwi);
      const double px = c.cx + c.fx * Xwi[0] / Xwi[2], py = c.cy + c.fy * Xwi[1] / Xwi[2];
      const double dx = z[0] - px, dy = z[1] - py;
      nrm = sqrt(dx * dx + dy * dy);
      const double gx = Xwi[0], gy = Xwi[1], gz = Xwi[2];
      const double gz2 = 1.0 / fmax(homog_th, gz * gz);
      const double fxdx = c.fx * dx, fydy = c.fy * dy;
      lba_jac_pose(gz2, fxdx, fydy, gx, gy, gz, JT);
      const double den = fmax(homog_th, nrm);
      for (int i = 0; i < 6; ++i) JT[i] = JT[i] / den;
      double J3[3] = {+gz2 * fxdx * gz, +gz2 * fydy * gz, -gz2 * (fxdx * gx + fydy * gy)};
      lba_rowvec_R(J3, Ti, JX);
      for (int i = 0; i < 3; ++i) JX[i] = JX[i] / den;
    } else {
      const double hom = (mode == 1 && quirks) ? 0.0000001 : homog_th;   // (q4)
      const double *P, *Q;
      if (mode == 1 && quirks) {
        P = Q = X + 6 * d.nkf + 3 * d.npt + 3 * lm;                      // (q2)
      } else {
        P = X + 6 * d.nkf + 3 * d.npt + 6 * lm;
        Q = P + 3;
      }
      double Pwi[3], Qwi[3];
      lba_xform(Ti, P, Pwi);
      lba_xform(Ti, Q, Qwi);
      const double ppx = c.cx + c.fx * Pwi[0] / Pwi[2], ppy = c.cy + c.fy * Pwi[1] / Pwi[2];
      const double qpx = c.cx + c.fx * Qwi[0] / Qwi[2], qpy = c.cy + c.fy * Qwi[1] / Qwi[2];
      const double e0 = z[0] * ppx + z[1] * ppy + z[2], e1 = z[0] * qpx + z[1] * qpy + z[2];
      nrm = sqrt(e0 * e0 + e1 * e1);
      const double fxlx = c.fx * e0, fyly = c.fy * e1;   // as the reference writes it (:1470-1471)
      const double den = fmax(hom, nrm);
      double JP[6], JQ[6], J3[3], JPw[3], JQw[3];
      {
        const double gx = Pwi[0], gy = Pwi[1], gz = Pwi[2], gz2 = 1.0 / fmax(hom, gz * gz);
        lba_jac_pose(gz2, fxlx, fyly, gx, gy, gz, JP);
        J3[0] = +gz2 * fxlx * gz; J3[1] = +gz2 * fyly * gz; J3[2] = -gz2 * (fxlx * gx + fyly * gy);
        lba_rowvec_R(J3, Ti, JPw);
        for (int i = 0; i < 3; ++i) JPw[i] = JPw[i] * e0 / den;
      }
      {
        const double gx = Qwi[0], gy = Qwi[1], gz = Qwi[2], gz2 = 1.0 / fmax(hom, gz * gz);
        lba_jac_pose(gz2, fxlx, fyly, gx, gy, gz, JQ);
        J3[0] = +gz2 * fxlx * gz; J3[1] = +gz2 * fyly * gz; J3[2] = -gz2 * (fxlx * gx + fyly * gy);
        lba_rowvec_R(J3, Ti, JQw);
        for (int i = 0; i < 3; ++i) JQw[i] = JQw[i] * e1 / den;
      }
      for (int i = 0; i < 6; ++i) JT[i] = (JP[i] * e0 + JQ[i] * e1) / den;
      for (int i = 0; i < 3; ++i) { JX[i] = JPw[i]; JX[3 + i] = JQw[i]; }
    }
    const double w = 1.0 / (1.0 + nrm * nrm);   // robustWeightCauchy
    for (int i = 0; i < 6; ++i) { d.Jt[6 * (size_t)k + i] = JT[i]; d.Jx[6 * (size_t)k + i] = JX[i]; }
    d.Wn[2 * (size_t)k] = w;
    d.Wn[2 * (size_t)k + 1] = nrm * w;
    e2w = nrm * nrm * w;
    if (kf >= 0) {
      for (int i = 0; i < 6; ++i) {
        atomicAdd(&d.gc[6 * kf + i], JT[i] * nrm * w);
        for (int j = 0; j < 6; ++j) atomicAdd(&d.A[36 * kf + 6 * i + j], JT[i] * JT[j] * w);
      }
    }
  }
  // error: warp reduce, one atomic per warp
  for (int off = 16; off > 0; off >>= 1) e2w += __shfl_xor_sync(0xFFFFFFFFu, e2w, off);
  if ((threadIdx.x & 31) == 0 && e2w != 0.0) atomicAdd(&d.scal[0], e2w);
}

__device__ __forceinline__ void lba_atomic_max_pos(double* addr, double v) {   // v >= 0: the bit patterns are ordered
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void __launch_bounds__(128) k_lba_lm_accum(LbaDev d) {
  const int l = blockIdx.x * 128 + threadIdx.x;
  if (l >= d.nlm) return;
  const int dim = l < d.npt ? 3 : 6;
  double D[36], g[6];
  for (int i = 0; i < 36; ++i) D[i] = 0.0;
  for (int i = 0; i < 6; ++i) g[i] = 0.0;
  for (int k = d.lm_start[l]; k < d.lm_start[l + 1]; ++k) {
    const double* jx = d.Jx + 6 * (size_t)k;
    const double w = d.Wn[2 * (size_t)k], rn = d.Wn[2 * (size_t)k + 1];
    for (int i = 0; i < dim; ++i) {
      g[i] += jx[i] * rn;
      for (int j = 0; j < dim; ++j) D[6 * i + j] += jx[i] * jx[j] * w;
    }
  }
  double mx = 0.0;
  for (int i = 0; i < dim; ++i) mx = fmax(mx, fabs(D[7 * i]));
  for (int i = 0; i < 36; ++i) d.D[36 * (size_t)l + i] = D[i];
  for (int i = 0; i < 6; ++i) d.gl[6 * (size_t)l + i] = g[i];
  lba_atomic_max_pos(&d.scal[2], mx);
}

__global__ void k_lba_kf_diagmax(LbaDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * d.nkf) lba_atomic_max_pos(&d.scal[2], fabs(d.A[36 * (i / 6) + 7 * (i % 6)]));
}

// in-place inverse of a symmetric positive definite dim x dim matrix (row stride 6) by Gauss-Jordan without pivoting
__device__ void lba_inv_spd(double* M, int dim, double* Minv) {
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) Minv[6 * i + j] = i == j ? 1.0 : 0.0;
  for (int p = 0; p < dim; ++p) {
    const double ip = 1.0 / M[7 * p];
    for (int j = 0; j < dim; ++j) { M[6 * p + j] *= ip; Minv[6 * p + j] *= ip; }
    for (int i = 0; i < dim; ++i) {
      if (i == p) continue;
      const double f = M[6 * i + p];
      for (int j = 0; j < dim; ++j) { M[6 * i + j] -= f * M[6 * p + j]; Minv[6 * i + j] -= f * Minv[6 * p + j]; }
    }
  }
}

__global__ void __launch_bounds__(128) k_lba_lm_schur(LbaDev d, double lambda) {
  const int l = blockIdx.x * 128 + threadIdx.x;
  if (l >= d.nlm) return;
  const int dim = l < d.npt ? 3 : 6, n = 6 * d.nkf;
  double D[36], Di[36], g[6];
  for (int i = 0; i < 36; ++i) D[i] = d.D[36 * (size_t)l + i];
  for (int i = 0; i < 6; ++i) g[i] = d.gl[6 * (size_t)l + i];
  for (int i = 0; i < dim; ++i) D[7 * i] += lambda * D[7 * i];
  lba_inv_spd(D, dim, Di);
  for (int i = 0; i < 36; ++i) d.Dinv[36 * (size_t)l + i] = Di[i];
  const int s = d.lm_start[l], e = d.lm_start[l + 1];
  for (int a = s; a < e; ++a) {
    const int ka = d.obs_kf[a];
    if (ka < 0) continue;
    const double* jxa = d.Jx + 6 * (size_t)a;
    const double* jta = d.Jt + 6 * (size_t)a;
    const double wa = d.Wn[2 * (size_t)a];
    double y[6];
    for (int i = 0; i < dim; ++i) {
      double v = 0.0;
      for (int j = 0; j < dim; ++j) v += Di[6 * i + j] * jxa[j];
      y[i] = wa * v;
    }
    double yg = 0.0;
    for (int i = 0; i < dim; ++i) yg += y[i] * g[i];
    for (int i = 0; i < 6; ++i) atomicAdd(&d.gred[6 * ka + i], -yg * jta[i]);
    for (int b = s; b < e; ++b) {
      const int kb = d.obs_kf[b];
      if (kb < 0) continue;
      const double* jxb = d.Jx + 6 * (size_t)b;
      const double* jtb = d.Jt + 6 * (size_t)b;
      double cab = 0.0;
      for (int i = 0; i < dim; ++i) cab += y[i] * jxb[i];
      cab *= d.Wn[2 * (size_t)b];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) atomicAdd(&d.S[(size_t)(6 * ka + i) * n + 6 * kb + j], -cab * jta[i] * jtb[j]);
    }
  }
}

// One CTA: S += A' (block diagonal, damped), g_red += g_c, dense LDL^T without pivoting, solve S dx_c = g_red.
__global__ void __launch_bounds__(256) k_lba_solve(LbaDev d, double lambda) {
  const int n = 6 * d.nkf, tid = threadIdx.x;
  double* S = d.S;
  for (int i = tid; i < 36 * d.nkf; i += 256) {
    const int kf = i / 36, r = (i % 36) / 6, cc = i % 6;
    double v = d.A[i];
    if (r == cc) v += lambda * v;
    S[(size_t)(6 * kf + r) * n + 6 * kf + cc] += v;
  }
  for (int i = tid; i < n; i += 256) d.gred[i] += d.gc[i];
  __syncthreads();
  // right-looking LDL^T on the lower triangle: after column j, S[i][j] = L_ij, S[j][j] = D_j
  for (int j = 0; j < n; ++j) {
    const double dj = S[(size_t)j * n + j];
    __syncthreads();
    for (int i = j + 1 + tid; i < n; i += 256) S[(size_t)i * n + j] /= dj;
    __syncthreads();
    // trailing update: S[i][k] -= L_ij D_j L_kj for j < k <= i
    const int m = n - j - 1;
    for (int t = tid; t < m * m; t += 256) {
      const int i = j + 1 + t / m, k = j + 1 + t % m;
      if (k <= i) S[(size_t)i * n + k] -= S[(size_t)i * n + j] * dj * S[(size_t)k * n + j];
    }
    __syncthreads();
  }
  if (tid == 0) {   // n <= 384: the two triangular solves are a few thousand operations
    double* x = d.dxc;
    for (int i = 0; i < n; ++i) {
      double v = d.gred[i];
      for (int k = 0; k < i; ++k) v -= S[(size_t)i * n + k] * x[k];
      x[i] = v;
    }
    for (int i = 0; i < n; ++i) x[i] /= S[(size_t)i * n + i];
    for (int i = n - 1; i >= 0; --i) {
      double v = x[i];
      for (int k = i + 1; k < n; ++k) v -= S[(size_t)k * n + i] * x[k];
      x[i] = v;
    }
    for (int i = 0; i < n; ++i) d.DX[i] = x[i];
  }
}

__global__ void __launch_bounds__(128) k_lba_backsub(LbaDev d) {
  const int l = blockIdx.x * 128 + threadIdx.x;
  if (l >= d.nlm) return;
  const int dim = l < d.npt ? 3 : 6;
  double r[6];
  for (int i = 0; i < 6; ++i) r[i] = d.gl[6 * (size_t)l + i];
  for (int a = d.lm_start[l]; a < d.lm_start[l + 1]; ++a) {
    const int ka = d.obs_kf[a];
    if (ka < 0) continue;
    const double* jt = d.Jt + 6 * (size_t)a;
    const double* jx = d.Jx + 6 * (size_t)a;
    double s = 0.0;
    for (int i = 0; i < 6; ++i) s += jt[i] * d.dxc[6 * ka + i];
    s *= d.Wn[2 * (size_t)a];
    for (int i = 0; i < dim; ++i) r[i] -= jx[i] * s;
  }
  const double* Di = d.Dinv + 36 * (size_t)l;
  const int off = l < d.npt ? 6 * d.nkf + 3 * l : 6 * d.nkf + 3 * d.npt + 6 * (l - d.npt);
  for (int i = 0; i < dim; ++i) {
    double v = 0.0;
    for (int j = 0; j < dim; ++j) v += Di[6 * i + j] * r[j];
    d.DX[off + i] = v;
  }
}

// |DX|^2 always; the step itself only when `apply`
__global__ void __launch_bounds__(128) k_lba_apply(LbaDev d, int apply) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  double n2 = 0.0;
  if (i < d.nkf) {
    for (int c = 0; c < 6; ++c) n2 += d.DX[6 * i + c] * d.DX[6 * i + c];
    if (apply) {
      double T[16];
      d_expmap(d.X + 6 * i, T);
      d_update_pose(T, d.DX + 6 * i);   // T <- T * inverse_se3(expmap_se3(dx))
      double x[6];
      d_logmap(T, x);
      for (int c = 0; c < 6; ++c) d.X[6 * i + c] = x[c];
    }
  } else if (i < d.nkf + (d.N - 6 * d.nkf)) {
    const int j = 6 * d.nkf + (i - d.nkf);
    n2 = d.DX[j] * d.DX[j];
    if (apply) d.X[j] += d.DX[j];
  }
  for (int off = 16; off > 0; off >>= 1) n2 += __shfl_xor_sync(0xFFFFFFFFu, n2, off);
  if ((threadIdx.x & 31) == 0 && n2 != 0.0) atomicAdd(&d.scal[1], n2);
}

__global__ void k_lba_prep(LbaDev d, const double* fixedT, double* fixedTinv, int n_fixed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d.nkf) {
    double T[16], Ti[16];
    d_expmap(d.X0 + 6 * i, T);
    d_inverse_se3(T, Ti);
    for (int k = 0; k < 16; ++k) d.T0inv[16 * i + k] = Ti[k];
  }
  if (i < n_fixed) {
    double Ti[16];
    d_inverse_se3(fixedT + 16 * i, Ti);
    for (int k = 0; k < 16; ++k) fixedTinv[16 * i + k] = Ti[k];
  }
}

__global__ void k_lba_moved(LbaDev d, uint8_t* pt_moved, uint8_t* ls_moved) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.nlm) return;
  const int dim = l < d.npt ? 3 : 6;
  const int off = l < d.npt ? 6 * d.nkf + 3 * l : 6 * d.nkf + 3 * d.npt + 6 * (l - d.npt);
  double n2 = 0.0;
  for (int c = 0; c < dim; ++c) { const double v = d.X[off + c] - d.X0[off + c]; n2 += v * v; }
  const uint8_t m = sqrt(n2) > 0.01;   // :1829, :1843
  if (l < d.npt) pt_moved[l] = m; else ls_moved[l - d.npt] = m;
}

static size_t lba_al(size_t x) { return (x + 255) & ~size_t(255); }

extern "C" plf_status plf_local_ba(plf_ctx* ctx, const plf_lba_opts* opts, const plf_lba_problem* p, plf_lba_result* out) {
  if (!ctx || !opts || !p || p->n_kf < 0 || p->n_pt < 0 || p->n_ls < 0 || p->n_pt_obs < 0 || p->n_ls_obs < 0 || p->n_fixed < 0)
    return plf_fail(ctx, PLF_ERR_INVALID, "plf_local_ba: bad arguments");
  const int nkf = p->n_kf, npt = p->n_pt, nls = p->n_ls, npo = p->n_pt_obs, nlo = p->n_ls_obs, nobs = npo + nlo, nlm = npt + nls;
  const int N = 6 * nkf + 3 * npt + 6 * nls, n = 6 * nkf;
  if (nobs == 0 || N == 0) return plf_fail(ctx, PLF_ERR_INVALID, "plf_local_ba: empty problem (the reference returns -1, :1324-1328)");
  if (nkf > 64) return plf_fail(ctx, PLF_ERR_CAPACITY, "plf_local_ba: %d local keyframes (limit 64)", nkf);
  if ((nkf && !p->kf_pose) || (npt && !p->pt) || (nls && !p->ls) || (npo && (!p->pt_obs_lm || !p->pt_obs_kf || !p->pt_obs_xy)) ||
      (nlo && (!p->ls_obs_lm || !p->ls_obs_kf || !p->ls_obs_le)) || (p->n_fixed && !p->fixed_T))
    return plf_fail(ctx, PLF_ERR_INVALID, "plf_local_ba: null array");
  // observations are grouped per landmark in non-decreasing landmark order (how the reference builds its lists, :1249-1277)
  std::vector<int> lm_start(nlm + 1, 0), h_lm(nobs), h_kf(nobs);
  std::vector<double> h_z((size_t)nobs * 3, 0.0);
  for (int k = 0; k < nobs; ++k) {
    const bool is_line = k >= npo;
    const int lm = is_line ? p->ls_obs_lm[k - npo] : p->pt_obs_lm[k], kf = is_line ? p->ls_obs_kf[k - npo] : p->pt_obs_kf[k];
    const int lim = is_line ? nls : npt;
    if (lm < 0 || lm >= lim || kf >= nkf || kf < -p->n_fixed)
      return plf_fail(ctx, PLF_ERR_INVALID, "plf_local_ba: observation %d refers to landmark %d / keyframe %d out of range", k, lm, kf);
    const int prev = k == 0 || k == npo ? -1 : (is_line ? p->ls_obs_lm[k - npo - 1] : p->pt_obs_lm[k - 1]);
    if (lm < prev) return plf_fail(ctx, PLF_ERR_INVALID, "plf_local_ba: observations must be grouped by landmark in ascending order");
    h_lm[k] = lm; h_kf[k] = kf;
    lm_start[(is_line ? npt : 0) + lm + 1]++;
    if (is_line) for (int c = 0; c < 3; ++c) h_z[3 * (size_t)k + c] = p->ls_obs_le[3 * (size_t)(k - npo) + c];
    else { h_z[3 * (size_t)k] = p->pt_obs_xy[2 * (size_t)k]; h_z[3 * (size_t)k + 1] = p->pt_obs_xy[2 * (size_t)k + 1]; }
  }
  for (int l = 0; l < nlm; ++l) lm_start[l + 1] += lm_start[l];
  PLF_CUDA(ctx, cudaSetDevice(ctx->device));
  // one scratch block
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += lba_al(bytes); return o; };
  const size_t oX = take((size_t)N * 8), oX0 = take((size_t)N * 8), oT0 = take((size_t)std::max(nkf, 1) * 128),
               oFT = take((size_t)std::max(p->n_fixed, 1) * 128), oFTi = take((size_t)std::max(p->n_fixed, 1) * 128),
               oLm = take((size_t)nobs * 4), oKf = take((size_t)nobs * 4), oZ = take((size_t)nobs * 24), oLs = take((size_t)(nlm + 1) * 4),
               oJt = take((size_t)nobs * 48), oJx = take((size_t)nobs * 48), oWn = take((size_t)nobs * 16),
               oD = take((size_t)std::max(nlm, 1) * 288), oGl = take((size_t)std::max(nlm, 1) * 48), oDi = take((size_t)std::max(nlm, 1) * 288),
               oA = take((size_t)std::max(nkf, 1) * 288), oGc = take((size_t)std::max(nkf, 1) * 48),
               oS = take((size_t)std::max(n * n, 1) * 8), oGr = take((size_t)std::max(n, 1) * 8), oDc = take((size_t)std::max(n, 1) * 8),
               oDX = take((size_t)N * 8), oSc = take(64), oPm = take((size_t)std::max(npt, 1)), oLmv = take((size_t)std::max(nls, 1));
  uint8_t* base = (uint8_t*)plf_scratch(ctx, 10, off);
  if (!base) return PLF_ERR_CUDA;
  LbaDev d;
  d.nkf = nkf; d.npt = npt; d.nls = nls; d.N = N; d.npo = npo; d.nlo = nlo; d.nobs = nobs; d.nlm = nlm;
  d.X = (double*)(base + oX); d.X0 = (double*)(base + oX0); d.T0inv = (double*)(base + oT0);
  double* dFT = (double*)(base + oFT); double* dFTi = (double*)(base + oFTi);
  d.fixedTinv = dFTi;
  int* dLm = (int*)(base + oLm); int* dKf = (int*)(base + oKf); double* dZ = (double*)(base + oZ); int* dLs = (int*)(base + oLs);
  d.obs_lm = dLm; d.obs_kf = dKf; d.obs_z = dZ; d.lm_start = dLs;
  d.Jt = (double*)(base + oJt); d.Jx = (double*)(base + oJx); d.Wn = (double*)(base + oWn);
  d.D = (double*)(base + oD); d.gl = (double*)(base + oGl); d.Dinv = (double*)(base + oDi);
  d.A = (double*)(base + oA); d.gc = (double*)(base + oGc);
  d.S = (double*)(base + oS); d.gred = (double*)(base + oGr); d.dxc = (double*)(base + oDc);
  d.DX = (double*)(base + oDX); d.scal = (double*)(base + oSc);
  uint8_t* dPm = base + oPm; uint8_t* dLmv = base + oLmv;
  cudaStream_t cs = ctx->stream;
  ctx->cur = cs;
  std::vector<double> hX(N);
  for (int i = 0; i < 6 * nkf; ++i) hX[i] = p->kf_pose[i];
  for (int i = 0; i < 3 * npt; ++i) hX[6 * nkf + i] = p->pt[i];
  for (int i = 0; i < 6 * nls; ++i) hX[6 * nkf + 3 * npt + i] = p->ls[i];
  PLF_CUDA(ctx, cudaMemcpyAsync(d.X, hX.data(), (size_t)N * 8, cudaMemcpyHostToDevice, cs));
  PLF_CUDA(ctx, cudaMemcpyAsync(d.X0, hX.data(), (size_t)N * 8, cudaMemcpyHostToDevice, cs));
  if (p->n_fixed) PLF_CUDA(ctx, cudaMemcpyAsync(dFT, p->fixed_T, (size_t)p->n_fixed * 128, cudaMemcpyHostToDevice, cs));
  PLF_CUDA(ctx, cudaMemcpyAsync(dLm, h_lm.data(), (size_t)nobs * 4, cudaMemcpyHostToDevice, cs));
  PLF_CUDA(ctx, cudaMemcpyAsync(dKf, h_kf.data(), (size_t)nobs * 4, cudaMemcpyHostToDevice, cs));
  PLF_CUDA(ctx, cudaMemcpyAsync(dZ, h_z.data(), (size_t)nobs * 24, cudaMemcpyHostToDevice, cs));
  PLF_CUDA(ctx, cudaMemcpyAsync(dLs, lm_start.data(), (size_t)(nlm + 1) * 4, cudaMemcpyHostToDevice, cs));
  k_lba_prep<<<(std::max(nkf, p->n_fixed) + 63) / 64 + 1, 64, 0, cs>>>(d, dFT, dFTi, p->n_fixed);
  PLF_LAUNCH_CHECK(ctx);
  const LbaCam cam = {ctx->cam.fx, ctx->cam.fy, ctx->cam.cx, ctx->cam.cy};
  const int q = opts->ref_quirks ? 1 : 0;
  double lambda = opts->lambda;
  double h_scal[3];
  // builds the normal equations of `mode` and leaves err / Hmax in scal
  auto build = [&](int mode) -> plf_status {
    PLF_CUDA(ctx, cudaMemsetAsync(d.A, 0, (size_t)std::max(nkf, 1) * 288, cs));
    PLF_CUDA(ctx, cudaMemsetAsync(d.gc, 0, (size_t)std::max(nkf, 1) * 48, cs));
    PLF_CUDA(ctx, cudaMemsetAsync(d.S, 0, (size_t)std::max(n * n, 1) * 8, cs));
    PLF_CUDA(ctx, cudaMemsetAsync(d.gred, 0, (size_t)std::max(n, 1) * 8, cs));
    PLF_CUDA(ctx, cudaMemsetAsync(d.scal, 0, 64, cs));
    k_lba_rows<<<(nobs + 127) / 128, 128, 0, cs>>>(d, cam, mode, q, opts->homog_th);
    PLF_LAUNCH_CHECK(ctx);
    k_lba_lm_accum<<<(nlm + 127) / 128, 128, 0, cs>>>(d);
    PLF_LAUNCH_CHECK(ctx);
    if (mode == 0 && nkf) {
      k_lba_kf_diagmax<<<(6 * nkf + 63) / 64, 64, 0, cs>>>(d);
      PLF_LAUNCH_CHECK(ctx);
    }
    PLF_CUDA(ctx, cudaMemcpyAsync(h_scal, d.scal, 24, cudaMemcpyDeviceToHost, cs));
    PLF_CUDA(ctx, cudaStreamSynchronize(cs));
    return PLF_OK;
  };
  // solves the damped system and (optionally) applies the step; leaves |DX| in *dx_norm
  auto solve = [&](int apply, double* dx_norm) -> plf_status {
    k_lba_lm_schur<<<(nlm + 127) / 128, 128, 0, cs>>>(d, lambda);
    PLF_LAUNCH_CHECK(ctx);
    if (nkf) {
      k_lba_solve<<<1, 256, 0, cs>>>(d, lambda);
      PLF_LAUNCH_CHECK(ctx);
    }
    k_lba_backsub<<<(nlm + 127) / 128, 128, 0, cs>>>(d);
    PLF_LAUNCH_CHECK(ctx);
    k_lba_apply<<<(nkf + (N - 6 * nkf) + 127) / 128, 128, 0, cs>>>(d, apply);
    PLF_LAUNCH_CHECK(ctx);
    PLF_CUDA(ctx, cudaMemcpyAsync(h_scal, d.scal, 24, cudaMemcpyDeviceToHost, cs));
    PLF_CUDA(ctx, cudaStreamSynchronize(cs));
    *dx_norm = sqrt(h_scal[1]);
    return PLF_OK;
  };
  plf_status st;
  // ---- first pass on the map's values (:1352-1541), lambda *= Hmax (:1543-1550), first step applied unconditionally
  if ((st = build(0))) return st;
  double err = h_scal[0];
  volatile int never_incremented = 0;   // Npt_obs + Nls_obs of the reference
  if (q) err /= (double)never_incremented;   // (q1) :1541
  else err /= (double)nobs;
  double Hmax;
  { unsigned long long bits; memcpy(&bits, &h_scal[2], 8); memcpy(&Hmax, &bits, 8); }
  lambda *= Hmax;
  double dxn = 0.0;
  if ((st = solve(1, &dxn))) return st;
  double err_prev = err;
  int iters;
  for (iters = 1; iters < opts->max_iters; ++iters) {
    if ((st = build(1))) return st;
    err = h_scal[0] / (double)(npt + nls);                                                       // :1773
    if (fabs(err - err_prev) < opts->min_error_change || err < opts->min_error) break;           // :1775
    const bool grew = err > err_prev;
    if ((st = solve(grew ? 0 : 1, &dxn))) return st;                                             // :1777-1803
    if (grew) lambda /= opts->lambda_k; else lambda *= opts->lambda_k;
    if (dxn < opts->min_error_change) break;                                                     // :1805
    err_prev = err;
  }
  k_lba_moved<<<(nlm + 127) / 128, 128, 0, cs>>>(d, dPm, dLmv);
  PLF_LAUNCH_CHECK(ctx);
  PLF_CUDA(ctx, cudaMemcpyAsync(hX.data(), d.X, (size_t)N * 8, cudaMemcpyDeviceToHost, cs));
  if (p->pt_moved && npt) PLF_CUDA(ctx, cudaMemcpyAsync(p->pt_moved, dPm, (size_t)npt, cudaMemcpyDeviceToHost, cs));
  if (p->ls_moved && nls) PLF_CUDA(ctx, cudaMemcpyAsync(p->ls_moved, dLmv, (size_t)nls, cudaMemcpyDeviceToHost, cs));
  PLF_CUDA(ctx, cudaStreamSynchronize(cs));
  for (int i = 0; i < 6 * nkf; ++i) p->kf_pose[i] = hX[i];
  for (int i = 0; i < 3 * npt; ++i) p->pt[i] = hX[6 * nkf + i];
  for (int i = 0; i < 6 * nls; ++i) p->ls[i] = hX[6 * nkf + 3 * npt + i];
  if (out) { out->iters = iters; out->err = err; out->lambda = lambda; }
  return PLF_OK;
}
