// common.cuh — shared device/host helpers for libpcgol_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/pcgol_b200.h"

namespace pcg {

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;

struct CudaError {
  cudaError_t e;
  const char* what;
  const char* file;
  int line;
};

#define PCG_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) throw ::pcg::CudaError{e__, #expr, __FILE__, __LINE__}; \
  } while (0)

// Optional per-kernel timing (pcg_profile_enable): CUDA events on the launching stream
// around every launch, aggregated by kernel name. Off by default; costs one relaxed load.
extern std::atomic<int> g_profile;
void prof_begin(const char* name, cudaStream_t s);
void prof_end(cudaStream_t s);

// Kernel launch with launch counting + immediate launch-error check.
#define PCG_LAUNCH_NAMED(name, kernel, grid, block, smem, stream, ...)        \
  do {                                                                        \
    const bool prof__ = ::pcg::g_profile.load(std::memory_order_relaxed) != 0; \
    if (prof__) ::pcg::prof_begin((name), (stream));                          \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);               \
    if (prof__) ::pcg::prof_end((stream));                                    \
    ::pcg::g_launches.fetch_add(1, std::memory_order_relaxed);                \
    PCG_CUDA(cudaGetLastError());                                             \
  } while (0)
#define PCG_LAUNCH(kernel, grid, block, smem, stream, ...) \
  PCG_LAUNCH_NAMED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

struct StatusError {
  pcg_status s;
  std::string msg;
};

// ---- stream-ordered scratch memory -----------------------------------------
// cudaMallocAsync from the device's default pool (release threshold raised at
// first use so that repeated calls do not return memory to the OS).
void ensure_pool(int device);

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() {}
  DevBuf(size_t count, cudaStream_t stream) { alloc(count, stream); }
  void alloc(size_t count, cudaStream_t stream) {
    release();
    n = count;
    s = stream;
    if (count) PCG_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), stream));
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    release();
    p = o.p;
    n = o.n;
    s = o.s;
    o.p = nullptr;
    return *this;
  }
  size_t bytes() const { return n * sizeof(T); }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    PCG_CUDA(cudaSetDevice(dev));
    ensure_pool(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Pinned staging buffer for small device->host results.
struct PinnedBuf {
  void* p = nullptr;
  explicit PinnedBuf(size_t bytes) { PCG_CUDA(cudaMallocHost(&p, bytes)); }
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
};

// ---- interleaved record accessor (pc.PointCloud layout) ----------------------
// pc/pointcloud.go:130-163: x,y,z float32 inside `stride`-byte records. The
// aligned case is the reference's float32Iterator fast path (iterator.go:96-149),
// the unaligned one its binaryFloat32Iterator (iterator.go:64-94).
struct CloudView {
  const uint8_t* data;
  int64_t n;
  int64_t stride;
  int32_t off[3];
  int32_t aligned;  // stride, offsets and base all multiples of 4
  int32_t packed;   // aligned && off == {o, o+4, o+8}
};

inline CloudView make_view(const void* data, int64_t n, int64_t stride, const int64_t off[3]) {
  CloudView v;
  v.data = (const uint8_t*)data;
  v.n = n;
  v.stride = stride;
  static const int64_t kDefault[3] = {0, 4, 8};
  if (!off) off = kDefault;
  for (int k = 0; k < 3; k++) v.off[k] = (int32_t)off[k];
  v.aligned = ((stride & 3) == 0 && (off[0] & 3) == 0 && (off[1] & 3) == 0 && (off[2] & 3) == 0 &&
               (((uintptr_t)data) & 3) == 0)
                  ? 1
                  : 0;
  v.packed = (v.aligned && off[1] == off[0] + 4 && off[2] == off[0] + 8) ? 1 : 0;
  return v;
}

inline void check_view_args(const void* data, int64_t n, int64_t stride, const int64_t off[3]) {
  if (n < 0) throw StatusError{PCG_E_INVALID_ARG, "negative point count"};
  if (n > 0 && !data) throw StatusError{PCG_E_INVALID_ARG, "null cloud pointer"};
  if (stride < 12) throw StatusError{PCG_E_INVALID_ARG, "record stride must be >= 12 bytes"};
  if (n >= ((int64_t)1 << 31)) throw StatusError{PCG_E_TOO_LARGE, "more than 2^31-1 points in one cloud"};
  if (off) {
    for (int k = 0; k < 3; k++)
      if (off[k] < 0 || off[k] + 4 > stride) throw StatusError{PCG_E_INVALID_ARG, "xyz offset outside the record"};
  }
}

#ifdef __CUDACC__
// binaryFloat32Iterator path (unaligned records): out of line so that the aligned fast path is
// three plain loads and not an if-converted mix of both.
static __device__ __noinline__ float load_f32_bytes(const uint8_t* p) {
  uint32_t b = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  return __uint_as_float(b);
}
__device__ __forceinline__ float load_f32_any(const uint8_t* p, int aligned) {
  if (aligned) return __ldg((const float*)p);
  return load_f32_bytes(p);
}
__device__ __forceinline__ void store_f32_any(uint8_t* p, float f, int aligned) {
  if (aligned) {
    *(float*)p = f;
    return;
  }
  uint32_t b = __float_as_uint(f);
  p[0] = (uint8_t)b;
  p[1] = (uint8_t)(b >> 8);
  p[2] = (uint8_t)(b >> 16);
  p[3] = (uint8_t)(b >> 24);
}
__device__ __forceinline__ float3 load_xyz(const CloudView& v, int64_t i) {
  const uint8_t* r = v.data + i * v.stride;
  if (v.aligned) {
    return make_float3(__ldg((const float*)(r + v.off[0])), __ldg((const float*)(r + v.off[1])),
                       __ldg((const float*)(r + v.off[2])));
  }
  return make_float3(load_f32_bytes(r + v.off[0]), load_f32_bytes(r + v.off[1]), load_f32_bytes(r + v.off[2]));
}

// mat/vec3.go:18-20,38-40 : ((dx*dx + dy*dy) + dz*dz), d = a - b, every op rounded, no FMA
__device__ __forceinline__ float dist_sq_ref(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ uint64_t shfl_down_u64(uint64_t v, int d) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_down_sync(0xffffffffu, lo, d);
  hi = __shfl_down_sync(0xffffffffu, hi, d);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int d) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, d);
  hi = __shfl_xor_sync(0xffffffffu, hi, d);
  return ((uint64_t)hi << 32) | lo;
}
#endif

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

}  // namespace pcg
