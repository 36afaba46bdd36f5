// icp_math.cuh — the float32 algebra of the ICP loop, written once for host and device.
// Every operation is individually rounded (Go/amd64 semantics): device code uses the
// _rn intrinsics (never contracted into FMA), host code relies on -ffp-contract=off.
#pragma once

#include <math.h>

#include "common.cuh"

namespace pcg {
namespace im {

#ifdef __CUDACC__
#define PCG_HD __host__ __device__ inline
#else
#define PCG_HD inline
#endif

PCG_HD float mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
PCG_HD float add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
PCG_HD float sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
PCG_HD float div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

struct M4 {
  float m[16];  // column-major, index = col*4 + row (mat/mat4.go:8-10)
};

// mat/mat4.go:16-28
PCG_HD M4 m4mul(const M4& m, const M4& a) {
  M4 out;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float sum = 0.f;
      for (int k = 0; k < 4; k++) sum = add(sum, mul(m.m[4 * k + i], a.m[4 * j + k]));
      out.m[4 * j + i] = sum;
    }
  return out;
}
// mat/mat4.go:30-36
PCG_HD M4 m4factor(const M4& m, float f) {
  M4 out;
  for (int i = 0; i < 16; i++) out.m[i] = mul(m.m[i], f);
  return out;
}
// mat/mat4.go:38-44
PCG_HD M4 m4add(const M4& m, const M4& a) {
  M4 out;
  for (int i = 0; i < 16; i++) out.m[i] = add(m.m[i], a.m[i]);
  return out;
}
// mat/transform.go:7-14
PCG_HD M4 m4translate(float x, float y, float z) {
  M4 o;
  for (int i = 0; i < 16; i++) o.m[i] = 0.f;
  o.m[0] = o.m[5] = o.m[10] = o.m[15] = 1.f;
  o.m[12] = x;
  o.m[13] = y;
  o.m[14] = z;
  return o;
}
// mat/mat4.go:130-137
PCG_HD void m4transform(const float* m, float x, float y, float z, float* ox, float* oy, float* oz) {
  float den = add(add(add(mul(m[3], x), mul(m[7], y)), mul(m[11], z)), m[15]);
  float w = div(1.0f, den);
  *ox = mul(add(add(add(mul(m[0], x), mul(m[4], y)), mul(m[8], z)), m[12]), w);
  *oy = mul(add(add(add(mul(m[1], x), mul(m[5], y)), mul(m[9], z)), m[13]), w);
  *oz = mul(add(add(add(mul(m[2], x), mul(m[6], y)), mul(m[10], z)), m[14]), w);
}

// pc/registration/icp/rodrigues.go:11-33
PCG_HD M4 rodrigues(float vx, float vy, float vz) {
  float nsq = add(add(mul(vx, vx), mul(vy, vy)), mul(vz, vz));
  float ang = (float)sqrt((double)nsq);  // mat/vec3.go:22-24
  M4 r;
  for (int i = 0; i < 16; i++) r.m[i] = 0.f;
  r.m[1] = vz;
  r.m[2] = -vy;
  r.m[4] = -vz;
  r.m[6] = vx;
  r.m[8] = vy;
  r.m[9] = -vx;
  M4 id = m4translate(0.f, 0.f, 0.f);
  float f0, f1;
  if (ang < 0.1f) {
    f0 = 1.f;
    f1 = 0.5f;
  } else {
    f0 = div((float)sin((double)ang), ang);
    f1 = div((float)(1.0 - cos((double)ang)), mul(ang, ang));
  }
  return m4add(m4add(id, m4factor(r, f0)), m4factor(m4mul(r, r), f1));
}

struct Sums {  // the nine float32 accumulators of Evaluate (evaluator.go:130-144)
  float value, sum_weight, g[6], rms;
};

struct Eval {  // Value, Gradient, DistRMS (evaluator.go:25-30)
  float value, g[6], dist_rms;
};

// evaluator.go:156-186 : normalisation, DistRMS, rotation limit
PCG_HD Eval evaluate_tail(const Sums& s) {
  Eval e;
  float f = 1.f;
  if (s.sum_weight > 1.f) f = div(1.f, s.sum_weight);
  e.value = mul(s.value, f);
  float two_f = mul(2.f, f);
  for (int i = 0; i < 6; i++) e.g[i] = mul(s.g[i], two_f);
  e.dist_rms = (float)sqrt((double)mul(s.rms, f));
  float rot_limit = 1.f;
  float dist = (float)sqrt((double)e.value);
  for (int i = 3; i < 6; i++) {
    float d = mul(e.g[i], e.dist_rms);
    if (d < 0.f) d = -d;
    if (dist < d) {
      float l = div(dist, d);
      if (rot_limit > l) rot_limit = l;
    }
  }
  for (int i = 3; i < 6; i++) e.g[i] = mul(e.g[i], rot_limit);
  return e;
}

struct UpdaterCfg {  // GradientDescentUpdaterFactory after defaults (updater.go:24-37)
  float weight[6], threshold[6];
  int max_iteration;
  int kind;  // PCG_UPDATER_GRADIENT_DESCENT (the reference's) | PCG_UPDATER_GAUSS_NEWTON
};

inline UpdaterCfg make_updater(const pcg_icp_params& p) {
  UpdaterCfg u;
  bool wz = true, tz = true;
  for (int k = 0; k < 6; k++) {
    wz = wz && p.weight[k] == 0.f;
    tz = tz && p.threshold[k] == 0.f;
  }
  for (int k = 0; k < 6; k++) {
    u.weight[k] = wz ? 0.3f : p.weight[k];
    u.threshold[k] = tz ? 0.01f : p.threshold[k];
  }
  u.max_iteration = p.max_iteration == 0 ? 20 : p.max_iteration;
  u.kind = p.updater;
  return u;
}

// updater.go:44-71. Returns converged; *iter is the updater's i.
PCG_HD bool updater_update(const UpdaterCfg& u, int* iter, M4* trans, const Eval& ev) {
  bool flat = true;
  for (int j = 0; j < 6; j++) {
    float g = ev.g[j];
    if (g < -u.threshold[j] || u.threshold[j] < g) {
      flat = false;
      break;
    }
  }
  if (flat) return true;
  float factor_iter = -sub(1.f, div((float)(*iter), (float)u.max_iteration));
  float delta[6];
  for (int k = 0; k < 6; k++) delta[k] = mul(mul(factor_iter, u.weight[k]), ev.g[k]);
  M4 dt = m4translate(delta[0], delta[1], delta[2]);
  M4 dr = rodrigues(delta[3], delta[4], delta[5]);
  *trans = m4mul(dt, m4mul(dr, *trans));
  (*iter)++;
  return *iter >= u.max_iteration;
}

// ---- normal equations ---------------------------------------------------------
// The reference declares Evaluated.Hessian (mat.Mat6, evaluator.go:25-30) but never writes it.
// Here it is the Gauss-Newton Hessian of Value = f * sum |pt - pb|^2 for the increment
// (dt, dw) applied as Translate(dt) * Rodrigues(dw) on the left (updater.go:65-68): the residual of
// a pair moves by J = [I | -[pt]x], so H = 2f * sum J^T J, which only needs the pair count, sum pt
// and the six second moments of pt (nine float64 sums).  Gradient = 2f * sum J^T r is the
// reference's (before its rotation limit).  Layout: 6x6, symmetric, index = col*6 + row.
struct HSums {
  double n;        // pairs
  double p[3];     // sum x, y, z of the transformed target points
  double pp[6];    // sum xx, xy, xz, yy, yz, zz
};

PCG_HD void normal_matrix(const HSums& h, double A[36]) {  // sum J^T J
  for (int i = 0; i < 36; i++) A[i] = 0.0;
  const double Sx = h.p[0], Sy = h.p[1], Sz = h.p[2];
  const double xx = h.pp[0], xy = h.pp[1], xz = h.pp[2], yy = h.pp[3], yz = h.pp[4], zz = h.pp[5];
  for (int i = 0; i < 3; i++) A[i * 6 + i] = h.n;
  // rows t, cols w: -[S]x ; rows w, cols t: [S]x
  const double tw[3][3] = {{0.0, Sz, -Sy}, {-Sz, 0.0, Sx}, {Sy, -Sx, 0.0}};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      A[(3 + c) * 6 + r] = tw[r][c];
      A[r * 6 + (3 + c)] = tw[r][c];  // symmetric counterpart: element (3+c, r)
    }
  const double ww[3][3] = {{yy + zz, -xy, -xz}, {-xy, xx + zz, -yz}, {-xz, -yz, xx + yy}};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) A[(3 + c) * 6 + (3 + r)] = ww[r][c];
}

// Evaluated.Hessian = 2f * sum J^T J with the f of evaluator.go:156-159
PCG_HD void hessian_from_sums(const HSums& h, float sum_weight, float out36[36]) {
  double A[36];
  normal_matrix(h, A);
  float f = 1.f;
  if (sum_weight > 1.f) f = div(1.f, sum_weight);
  const double s = 2.0 * (double)f;
  for (int i = 0; i < 36; i++) out36[i] = (float)(A[i] * s);
}

// Solves (A + lambda I) d = -b by Cholesky in float64; lambda starts at a relative 1e-12 and grows
// until the factorisation succeeds (A is positive semi-definite by construction).
PCG_HD bool solve_normal_equations(const double A_in[36], const double b[6], double d[6]) {
  double tr = 0.0;
  for (int i = 0; i < 6; i++) tr += A_in[i * 6 + i];
  if (!(tr > 0.0) || !(tr < 1.0e300)) return false;
  double lambda = 1.0e-12 * tr / 6.0;
  for (int attempt = 0; attempt < 12; attempt++, lambda *= 100.0) {
    double L[36];
    bool ok = true;
    for (int i = 0; i < 36; i++) L[i] = 0.0;
    for (int j = 0; j < 6 && ok; j++) {
      double djj = A_in[j * 6 + j] + lambda;
      for (int k = 0; k < j; k++) djj -= L[k * 6 + j] * L[k * 6 + j];  // L[col*6+row]
      if (!(djj > 0.0)) {
        ok = false;
        break;
      }
      const double ljj = sqrt(djj);
      L[j * 6 + j] = ljj;
      for (int i = j + 1; i < 6; i++) {
        double v = A_in[j * 6 + i];
        for (int k = 0; k < j; k++) v -= L[k * 6 + i] * L[k * 6 + j];
        L[j * 6 + i] = v / ljj;
      }
    }
    if (!ok) continue;
    double y[6];
    for (int i = 0; i < 6; i++) {  // L y = -b
      double v = -b[i];
      for (int k = 0; k < i; k++) v -= L[k * 6 + i] * y[k];
      y[i] = v / L[i * 6 + i];
    }
    for (int i = 5; i >= 0; i--) {  // L^T d = y
      double v = y[i];
      for (int k = i + 1; k < 6; k++) v -= L[i * 6 + k] * d[k];
      d[i] = v / L[i * 6 + i];
    }
    return true;
  }
  return false;
}

// Gauss-Newton updater: same convergence test and the same composition of the increment as
// gradientDescentUpdater.Update (updater.go:45-54,65-70); the step is the solution of the normal
// equations instead of a damped gradient.  g_raw = sum J^T r (the reference's sums before 2f and
// before the rotation limit).  Returns converged.
PCG_HD bool updater_update_gn(const UpdaterCfg& u, int* iter, M4* trans, const Eval& ev, const Sums& sums,
                              const HSums& h) {
  bool flat = true;
  for (int j = 0; j < 6; j++) {
    float g = ev.g[j];
    if (g < -u.threshold[j] || u.threshold[j] < g) {
      flat = false;
      break;
    }
  }
  if (flat) return true;
  double A[36], b[6], d[6];
  normal_matrix(h, A);
  for (int k = 0; k < 6; k++) b[k] = (double)sums.g[k];
  if (!solve_normal_equations(A, b, d)) return true;  // degenerate geometry: stop where we are
  M4 dt = m4translate((float)d[0], (float)d[1], (float)d[2]);
  M4 dr = rodrigues((float)d[3], (float)d[4], (float)d[5]);
  *trans = m4mul(dt, m4mul(dr, *trans));
  (*iter)++;
  return *iter >= u.max_iteration;
}

}  // namespace im
}  // namespace pcg
