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

}  // namespace im
}  // namespace pcg
