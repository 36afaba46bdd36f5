// vg_common.cuh — grid arithmetic of filter.VoxelGrid shared by the VoxelGrid translation units:
// Filter's parameters (voxelgrid.go:45-63,137-138) and the (chunk id, voxel key) of a point
// (voxelgrid.go:76-79,88,149-151), in Go's float32 semantics.
#pragma once

#include "common.cuh"
#include "icp_math.cuh"

namespace pcg {

// ---- MinMaxVec3 (pc/minmax.go:9-26) ------------------------------------------------
// Go keeps the FIRST occurrence of the extreme value (strict comparisons, -0 == +0), so
// candidates are packed as (order-preserving value bits, index) and reduced with 64-bit
// min / max; the winning index is then dereferenced so the sign of a zero survives.
__device__ __forceinline__ uint32_t ordered_bits(float f) {
  if (f == 0.0f) f = 0.0f;  // -0 -> +0: they compare equal in Go
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ---- grid parameters, computed on the host in Go's float32 semantics ----------------
// (this translation unit's host code is compiled with -ffp-contract=off)
struct VgParams {
  float vmin[3];
  float leaf[3];
  float chunk_size[3];
  int32_t chunked;
  int64_t nx, ny, n_chunks;
  int64_t xs, ys;
  int64_t n_voxels;
  int32_t key_bits;
  int32_t small;  // every count fits 30 bits: in-range points take the 32-bit arithmetic path
  int64_t zs;
  float rleaf[3], rchunk[3];  // 1/leaf, 1/chunk_size: quotient estimates (the exact division decides near integers)
};

PCG_HD bool go_int_hd(float f, long long* out) {
  // Go's int(float32): truncation toward zero; out of int64 range / NaN is implementation-specific
  if (!(fabsf(f) < 9.0e18f)) return false;
#ifdef __CUDA_ARCH__
  *out = __float2ll_rz(f);
#else
  *out = (long long)f;
#endif
  return true;
}

PCG_HD int bits_for(long long count) {  // bits to represent values in [0, count)
  int b = 0;
  while (b < 63 && (1ll << b) < count) b++;
  return b;
}

// Filter's grid arithmetic (voxelgrid.go:45-63,137-138) in Go's float32 semantics; shared by the
// host path and the fused kernel (im:: ops are the _rn intrinsics on the device and plain
// operators, compiled with -ffp-contract=off, on the host).
PCG_HD pcg_status vg_make_params(const float vmin[3], const float vmax[3], const float leaf[3], const long long chunk[3],
                                 VgParams* Pout, int* total_bits_out) {
  VgParams P;
  for (int k = 0; k < 3; k++) {
    P.vmin[k] = vmin[k];
    P.leaf[k] = leaf[k];
    P.chunk_size[k] = 0.f;
  }
  float size_for_grid[3];
  long long nz = 1;
  P.nx = P.ny = 1;
  P.n_chunks = 1;
  P.chunked = (chunk[0] * chunk[1] * chunk[2] != 0) ? 1 : 0;  // voxelgrid.go:45
  if (!P.chunked) {
    for (int k = 0; k < 3; k++) size_for_grid[k] = vmax[k];  // sic: voxelgrid.go:46 passes vMax as size
  } else {
    float size[3];
    for (int k = 0; k < 3; k++) {
      size[k] = im::sub(vmax[k], vmin[k]);                 // :49
      float cs = im::mul(leaf[k], (float)chunk[k]);        // :50-54
      float lim = im::add(size[k], leaf[k]);               // :58
      if (cs > lim) cs = lim;
      P.chunk_size[k] = cs;
      size_for_grid[k] = cs;
    }
    long long c[3];
    for (int k = 0; k < 3; k++) {
      if (!go_int_hd(im::div(size[k], P.chunk_size[k]), &c[k])) return PCG_E_REF_UNDEFINED;  // :62
      c[k] += 1;
    }
    P.nx = c[0];
    P.ny = c[1];
    nz = c[2];
    if (c[0] <= 0 || c[1] <= 0 || c[2] <= 0 || c[0] > (1ll << 40) / c[1] || c[0] * c[1] > (1ll << 40) / c[2])
      return PCG_E_TOO_LARGE;
    P.n_chunks = c[0] * c[1] * nz;
  }
  long long s[3];
  for (int k = 0; k < 3; k++) {
    if (!go_int_hd(im::div(size_for_grid[k], leaf[k]), &s[k])) return PCG_E_REF_UNDEFINED;  // :137
  }
  P.xs = s[0];
  P.ys = s[1];
  // nVoxels = (xs+1)*(ys+1)*(zs+1)  :138 ; a non-positive product leaves the dense array empty,
  // so the first indexed write panics.
  const double nv = (double)(s[0] + 1) * (double)(s[1] + 1) * (double)(s[2] + 1);
  if (nv >= 9.0e18 || nv <= -9.0e18) return PCG_E_TOO_LARGE;
  P.n_voxels = (s[0] + 1) * (s[1] + 1) * (s[2] + 1);
  if (P.n_voxels <= 0) return PCG_E_REF_WOULD_PANIC;
  P.key_bits = bits_for(P.n_voxels);
  P.zs = s[2];
  for (int k = 0; k < 3; k++) {
    P.rleaf[k] = im::div(1.f, leaf[k]);
    P.rchunk[k] = P.chunked ? im::div(1.f, P.chunk_size[k]) : 0.f;
  }
  P.small = (P.n_voxels < (1ll << 30) && P.n_chunks < (1ll << 30) && s[0] >= 0 && s[1] >= 0 && s[2] >= 0) ? 1 : 0;
  int total_bits = P.key_bits + bits_for(P.n_chunks);
  if (total_bits > 64) return PCG_E_TOO_LARGE;
  if (total_bits == 0) total_bits = 1;
  *Pout = P;
  *total_bits_out = total_bits;
  return PCG_OK;
}

inline const char* vg_status_message(pcg_status s) {
  switch (s) {
    case PCG_E_REF_UNDEFINED: return "voxel / chunk grid size is not finite or out of int64 range";
    case PCG_E_TOO_LARGE: return "voxel grid or chunk table too large";
    case PCG_E_REF_WOULD_PANIC: return "reference would index an empty voxel array";
    default: return "voxel grid parameter error";
  }
}

enum { kFlagPanic = 1, kFlagUndefined = 2 };

__device__ __forceinline__ bool go_int_dev(float f, long long* out) {
  if (!(fabsf(f) < 9.0e18f)) return false;
  *out = __float2ll_rz(f);
  return true;
}

// voxelgrid.go:69-75,109-110 : vcMin = vMin + cid2xyz(cid) (*) chunkSize
__device__ __forceinline__ void chunk_min_xyz(const VgParams& P, long long x, long long y, long long z, float out[3]) {
  out[0] = __fadd_rn(P.vmin[0], __fmul_rn((float)x, P.chunk_size[0]));
  out[1] = __fadd_rn(P.vmin[1], __fmul_rn((float)y, P.chunk_size[1]));
  out[2] = __fadd_rn(P.vmin[2], __fmul_rn((float)z, P.chunk_size[2]));
}
__device__ __forceinline__ void chunk_min(const VgParams& P, long long cid, float out[3]) {
  if (!P.chunked) {
    out[0] = P.vmin[0];
    out[1] = P.vmin[1];
    out[2] = P.vmin[2];
    return;
  }
  if (P.n_chunks <= 0x7fffffffll) {  // 32-bit division is several times cheaper
    const uint32_t c = (uint32_t)cid, nx = (uint32_t)P.nx, ny = (uint32_t)P.ny;
    const uint32_t t = c / nx;
    chunk_min_xyz(P, c - t * nx, t % ny, t / ny, out);
  } else {
    const long long t = cid / P.nx;
    chunk_min_xyz(P, cid % P.nx, t % P.ny, t / P.ny, out);
  }
}

// voxelgrid.go:76-79,88 (vec2cid) and :149-151 (voxel key) for one point:
// (chunk id << key_bits) | (x + xs*(y + ys*z)).
// General path: 64-bit arithmetic, every out-of-range case of the reference (aliasing, panics) handled.
static __device__ __noinline__ unsigned long long voxel_key_general(const VgParams& P, const float3 pt, int* bad) {
  long long cid = 0;
  float vc[3] = {P.vmin[0], P.vmin[1], P.vmin[2]};
  if (P.chunked) {
    long long cx, cy, cz;
    bool ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.x, P.vmin[0]), P.chunk_size[0]), &cx);
    ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.y, P.vmin[1]), P.chunk_size[1]), &cy) && ok;
    ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.z, P.vmin[2]), P.chunk_size[2]), &cz) && ok;
    if (!ok) {
      *bad |= kFlagUndefined;
      cx = cy = cz = 0;
    }
    cid = ((cz * P.ny) + cy) * P.nx + cx;
    if (cid < 0 || cid >= P.n_chunks) {
      *bad |= kFlagPanic;
      cid = 0;
      cx = cy = cz = 0;
    }
    if (cx >= 0 && cx < P.nx && cy >= 0 && cy < P.ny && cz >= 0)
      chunk_min_xyz(P, cx, cy, cz, vc);  // cid2xyz(cid) == (cx, cy, cz) when every coordinate is in range
    else
      chunk_min(P, cid, vc);             // out-of-range coordinates alias into another chunk (voxelgrid.go:69-79)
  }
  long long x, y, z;
  bool ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.x, vc[0]), P.leaf[0]), &x);
  ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.y, vc[1]), P.leaf[1]), &y) && ok;
  ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.z, vc[2]), P.leaf[2]), &z) && ok;
  if (!ok) {
    *bad |= kFlagUndefined;
    x = y = z = 0;
  }
  long long key = x + P.xs * (y + P.ys * z);
  if (key < 0 || key >= P.n_voxels) {
    *bad |= kFlagPanic;
    key = 0;
  }
  return ((unsigned long long)cid << P.key_bits) | (unsigned long long)key;
}

// Truncated quotients int(float32(a / d)) of three coordinates at once, branch-free on the common path.
// q = a * (1/d) is within a few ulp of the correctly rounded quotient, so int(q) can differ from the
// reference's int(a / d) only when an integer lies between the two, i.e. when q is (relatively) within 2^-19
// of an integer.  Those lanes - and only those - evaluate the IEEE division.  Returns false when a quotient does
// not fit 31 bits (NaN included): the general path then takes over.
__device__ __forceinline__ bool div3_to_int32(const float a[3], const float d[3], const float r[3], int out[3]) {
  float q[3];
  bool ok = true, near = false;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    q[k] = __fmul_rn(a[k], r[k]);
    const float qa = fabsf(q[k]);
    ok = ok && qa < 1073741824.0f;
    out[k] = __float2int_rz(q[k]);
    const float fr = fabsf(__fsub_rn(q[k], (float)out[k]));  // exact: |q| < 2^30 leaves no rounding here
    const float e = __fmul_rn(qa, 1.9073486e-6f);
    near = near || fr <= e || fr >= __fsub_rn(1.0f, e);
  }
  if (!ok) return false;
  if (near) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float ex = __fdiv_rn(a[k], d[k]);
      ok = ok && fabsf(ex) < 1073741824.0f;
      out[k] = __float2int_rz(ex);
    }
  }
  return ok;
}

// The launch constants of the key arithmetic, read from shared memory once per thread.
struct KeyConsts {
  float vmin[3], leaf[3], rleaf[3], chunk_size[3], rchunk[3];
  uint32_t nx, ny, n_chunks, xs, ys, zs, n_voxels;
  int key_bits, chunked, small;
  __device__ __forceinline__ explicit KeyConsts(const VgParams& P) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      vmin[k] = P.vmin[k];
      leaf[k] = P.leaf[k];
      rleaf[k] = P.rleaf[k];
      chunk_size[k] = P.chunk_size[k];
      rchunk[k] = P.rchunk[k];
    }
    nx = (uint32_t)P.nx;
    ny = (uint32_t)P.ny;
    n_chunks = (uint32_t)P.n_chunks;
    xs = (uint32_t)P.xs;
    ys = (uint32_t)P.ys;
    zs = (uint32_t)P.zs;
    n_voxels = (uint32_t)P.n_voxels;
    key_bits = P.key_bits;
    chunked = P.chunked;
    small = P.small;
  }
};

// voxelgrid.go:76-79,88 (vec2cid) and :149-151 (voxel key) for one point:
// (chunk id << key_bits) | (x + xs*(y + ys*z)).  Points whose chunk and voxel coordinates are in range - all of
// them on any input the reference accepts - are done in 32-bit arithmetic (same float operations, same
// truncation).  Returns false when the point needs the general path.
__device__ __forceinline__ bool voxel_key_fast(const KeyConsts& C, const float3 pt, unsigned long long* out) {
  if (!C.small) return false;
  uint32_t cid = 0;
  float vc[3] = {C.vmin[0], C.vmin[1], C.vmin[2]};
  const float p[3] = {pt.x, pt.y, pt.z};
  bool fast = true;
  if (C.chunked) {
    float a[3];
    int c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) a[k] = __fsub_rn(p[k], C.vmin[k]);
    fast = div3_to_int32(a, C.chunk_size, C.rchunk, c);
    fast = fast && (uint32_t)c[0] < C.nx && (uint32_t)c[1] < C.ny && c[2] >= 0;
    const unsigned long long c64 = ((unsigned long long)(uint32_t)c[2] * C.ny + (uint32_t)c[1]) * C.nx + (uint32_t)c[0];
    fast = fast && c64 < (unsigned long long)C.n_chunks;
    cid = (uint32_t)c64;
#pragma unroll
    for (int k = 0; k < 3; k++) vc[k] = __fadd_rn(C.vmin[k], __fmul_rn((float)c[k], C.chunk_size[k]));  // voxelgrid.go:109-110
  }
  float a[3];
  int x[3];
#pragma unroll
  for (int k = 0; k < 3; k++) a[k] = __fsub_rn(p[k], vc[k]);
  fast = div3_to_int32(a, C.leaf, C.rleaf, x) && fast;
  // coordinates inside [0, size+1] keep x + xs*(y + ys*z) below 2^32 when n_voxels < 2^30
  fast = fast && (uint32_t)x[0] <= C.xs + 1u && (uint32_t)x[1] <= C.ys + 1u && (uint32_t)x[2] <= C.zs + 1u;
  const uint32_t key = (uint32_t)x[0] + C.xs * ((uint32_t)x[1] + C.ys * (uint32_t)x[2]);
  fast = fast && key < C.n_voxels;
  *out = ((unsigned long long)cid << C.key_bits) | (unsigned long long)key;
  return fast;
}
__device__ __forceinline__ unsigned long long voxel_key_of(const VgParams& P, const KeyConsts& C, const float3 pt,
                                                           int* bad) {
  unsigned long long k;
  if (voxel_key_fast(C, pt, &k)) return k;
  return voxel_key_general(P, pt, bad);
}

// Small pinned staging area per host thread for device->host results (pageable
// destinations would add a staging copy to every readback).
struct PinnedScratch {
  unsigned char* p = nullptr;
  PinnedScratch() { cudaMallocHost((void**)&p, 256); }
};
inline unsigned char* pinned_scratch() {
  static thread_local PinnedScratch s;  // lives as long as the thread; 256 bytes
  if (!s.p) throw StatusError{PCG_E_CUDA, "cudaMallocHost failed"};
  return s.p;
}
inline unsigned char* vg_pinned_state() { return pinned_scratch() + 128; }

inline void vg_throw_on_flags(int h_flags) {
  if (h_flags & kFlagUndefined)
    throw StatusError{PCG_E_REF_UNDEFINED, "a voxel coordinate is not finite / out of int64 range"};
  if (h_flags & kFlagPanic)
    throw StatusError{PCG_E_REF_WOULD_PANIC,
                      "reference would panic: voxel or chunk index out of range (voxelgrid.go:46,89,151)"};
}

}  // namespace pcg
