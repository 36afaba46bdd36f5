// voxelgrid.cu — filter.VoxelGrid (pc/filter/voxelgrid/voxelgrid.go:35-187) on sm_100a.
//
// The reference fills a dense voxel array per chunk and scans it in index order.
// Here memory is proportional to the points: every point gets the 64-bit key
// (chunk id << key_bits | voxel key) the reference would have used as
// (chunk loop position, dense-array index); a stable radix sort of (key, point index)
// puts voxels in the reference's output order and the members of each voxel in the
// reference's accumulation order; one pass over the sorted list then sums every
// voxel sequentially (bit-exact float32 centroid), copies the first member's whole
// record and overwrites x,y,z when the voxel has more than one member.
#include <cooperative_groups.h>

#include "common.cuh"
#include "icp_math.cuh"
#include "radix_sort.cuh"

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

__global__ void __launch_bounds__(256) minmax_kernel(CloudView v, unsigned long long* __restrict__ out6) {
  __shared__ unsigned long long s_red[8][6];
  unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0ull, 0ull, 0ull};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += stride) {
    float3 p = load_xyz(v, i);
    float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (c[k] != c[k]) continue;  // NaN never wins a comparison in the reference
      unsigned long long o = (unsigned long long)ordered_bits(c[k]) << 32;
      unsigned long long a = o | (uint32_t)i, b = o | (0xffffffffu - (uint32_t)i);
      mn[k] = a < mn[k] ? a : mn[k];
      mx[k] = b > mx[k] ? b : mx[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      unsigned long long a = shfl_xor_u64(mn[k], d), b = shfl_xor_u64(mx[k], d);
      mn[k] = a < mn[k] ? a : mn[k];
      mx[k] = b > mx[k] ? b : mx[k];
    }
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      s_red[warp][k] = mn[k];
      s_red[warp][3 + k] = mx[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {  // one global atomic per CTA and component
    const int k = threadIdx.x;
    unsigned long long r = s_red[0][k];
    for (int w = 1; w < 8; w++) {
      unsigned long long o = s_red[w][k];
      r = k < 3 ? (o < r ? o : r) : (o > r ? o : r);
    }
    if (k < 3) {
      if (r != ~0ull) atomicMin(&out6[k], r);
    } else {
      if (r != 0ull) atomicMax(&out6[k], r);
    }
  }
}

__global__ void minmax_finalize_kernel(CloudView v, const unsigned long long* __restrict__ in6,
                                       float* __restrict__ out6) {
  int k = threadIdx.x;
  if (k >= 6) return;
  int c = k % 3;
  float3 p0 = load_xyz(v, 0);
  float first = c == 0 ? p0.x : (c == 1 ? p0.y : p0.z);
  float r;
  if (first != first) {
    r = first;  // min/max start at point 0; a NaN there is never replaced (minmax.go:13,17-22)
  } else {
    unsigned long long w = in6[k];
    uint32_t idx = k < 3 ? (uint32_t)w : 0xffffffffu - (uint32_t)w;
    float3 p = load_xyz(v, idx);
    r = c == 0 ? p.x : (c == 1 ? p.y : p.z);
  }
  out6[k] = r;
}

// Small pinned staging area per host thread for device->host results (pageable
// destinations would add a staging copy to every readback).
struct PinnedScratch {
  unsigned char* p = nullptr;
  PinnedScratch() { cudaMallocHost((void**)&p, 256); }
};
static unsigned char* pinned_scratch() {
  static thread_local PinnedScratch s;  // lives as long as the thread; 256 bytes
  if (!s.p) throw StatusError{PCG_E_CUDA, "cudaMallocHost failed"};
  return s.p;
}

void minmax_device(const CloudView& v, float mn[3], float mx[3], cudaStream_t stream) {
  DevBuf<unsigned long long> acc(6, stream);
  DevBuf<float> res(6, stream);
  unsigned long long* init = (unsigned long long*)(pinned_scratch() + 64);
  for (int k = 0; k < 3; k++) {
    init[k] = ~0ull;
    init[3 + k] = 0ull;
  }
  PCG_CUDA(cudaMemcpyAsync(acc.p, init, 6 * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
  int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(v.n, 256));
  PCG_LAUNCH(minmax_kernel, blocks, 256, 0, stream, v, acc.p);
  PCG_LAUNCH(minmax_finalize_kernel, 1, 32, 0, stream, v, acc.p, res.p);
  float* h = (float*)pinned_scratch();
  PCG_CUDA(cudaMemcpyAsync(h, res.p, 6 * sizeof(float), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  for (int k = 0; k < 3; k++) {
    mn[k] = h[k];
    mx[k] = h[3 + k];
  }
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
  int total_bits = P.key_bits + bits_for(P.n_chunks);
  if (total_bits > 64) return PCG_E_TOO_LARGE;
  if (total_bits == 0) total_bits = 1;
  *Pout = P;
  *total_bits_out = total_bits;
  return PCG_OK;
}

static const char* vg_status_message(pcg_status s) {
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
__device__ __forceinline__ unsigned long long voxel_key_of(const VgParams& P, const float3 pt, int* bad) {
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

// Keys for the multi-kernel path; the digit histograms of the sort are accumulated here so the
// keys are not read a second time.
template <typename K>
__global__ void __launch_bounds__(256)
    voxel_key_kernel(CloudView v, VgParams P, K* __restrict__ keys, int* __restrict__ flags,
                     uint32_t* __restrict__ hist, int passes) {
  __shared__ uint32_t s_hist[rsort::kMaxPasses * rsort::kRadix];
  rsort::hist_zero(s_hist, passes);
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (v.n + stride - 1) / stride;
  int bad = 0;
  for (int64_t r = 0; r < rounds; r++) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < v.n;
    K out_key = 0;
    if (valid) {
      out_key = (K)voxel_key_of(P, load_xyz(v, i), &bad);
      keys[i] = out_key;
    }
    rsort::hist_add_key(s_hist, out_key, valid, 0, passes);
  }
  if (bad) atomicOr(flags, bad);
  __syncthreads();
  rsort::hist_flush(s_hist, hist, passes);
}

// ---- segmented centroid + record gather ----------------------------------------------
// One pass over the sorted (key, index) list.  Phase A: every position gathers its point
// and stores p = pt - vcMin in shared memory (all threads busy, latencies overlap).
// Phase B: segment heads are counted per tile and chained between CTAs with a decoupled
// look-back so that every head knows its output slot; the head thread then adds its voxel's
// members in list order out of shared memory (voxelgrid.go:148-158) — the float32 sum is the
// reference's — and writes the first member's record with the centroid (voxelgrid.go:173-184).
constexpr int kSegThreads = 256;
constexpr int kSegItems = 4;
constexpr int kSegTile = kSegThreads * kSegItems;

template <typename K>
__global__ void __launch_bounds__(kSegThreads)
    voxel_reduce_kernel(CloudView v, VgParams P, const K* __restrict__ keys, const uint32_t* __restrict__ vals,
                        uint8_t* __restrict__ out, uint32_t* __restrict__ tile_counter,
                        unsigned long long* __restrict__ status, long long* __restrict__ n_out) {
  __shared__ float s_p[3][kSegTile];
  __shared__ K s_key[kSegTile];
  __shared__ uint32_t s_scan[rsort::kWarps];
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_prefix;
  const uint32_t tid = threadIdx.x;
  const uint32_t n = (uint32_t)v.n;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_base = tile * kSegTile;
  const uint32_t tile_count = min((uint32_t)kSegTile, n - tile_base);

  // Phase A (striped: coalesced key/index loads, independent gathers)
#pragma unroll
  for (int j = 0; j < kSegItems; j++) {
    const uint32_t l = j * kSegThreads + tid;
    if (l < tile_count) {
      const K key = keys[tile_base + l];
      const float3 pt = load_xyz(v, vals[tile_base + l]);
      float vc[3];
      chunk_min(P, (long long)((unsigned long long)key >> P.key_bits), vc);
      s_key[l] = key;
      s_p[0][l] = __fsub_rn(pt.x, vc[0]);
      s_p[1][l] = __fsub_rn(pt.y, vc[1]);
      s_p[2][l] = __fsub_rn(pt.z, vc[2]);
    }
  }
  __syncthreads();

  // Phase B (blocked: thread t owns positions 4t .. 4t+3)
  const uint32_t l0 = tid * kSegItems;
  K prev = 0;
  if (l0 > 0 && l0 - 1 < tile_count)
    prev = s_key[l0 - 1];
  else if (l0 == 0 && tile_base > 0)
    prev = keys[tile_base - 1];
  uint32_t heads = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < kSegItems; j++) {
    const uint32_t l = l0 + j;
    if (l < tile_count) {
      const K k = s_key[l];
      const bool h = (tile_base + l == 0) || k != prev;
      heads |= (h ? 1u : 0u) << j;
      cnt += h ? 1u : 0u;
      prev = k;
    }
  }
  uint32_t total = 0;
  const uint32_t excl = rsort::block_excl_scan_256(cnt, s_scan, &total);
  if (tid == 0) {
    volatile unsigned long long* st = status;
    unsigned long long prefix = 0;
    if (tile == 0) {
      st[0] = (2ull << 62) | (unsigned long long)total;
    } else {
      st[tile] = (1ull << 62) | (unsigned long long)total;
      int64_t pv = (int64_t)tile - 1;
      for (;;) {
        unsigned long long w = st[pv];
        unsigned long long state = w >> 62;
        if (state == 0) continue;
        prefix += w & ((1ull << 62) - 1);
        if (state == 2) break;
        pv--;
      }
      st[tile] = (2ull << 62) | (prefix + total);
    }
    s_prefix = prefix;
    if ((uint64_t)tile_base + kSegTile >= n) *n_out = (long long)(prefix + total);
  }
  __syncthreads();
  uint64_t rank = s_prefix + excl;
  const int out_aligned = v.aligned && ((((uintptr_t)out) & 3) == 0);

#pragma unroll
  for (int j = 0; j < kSegItems; j++) {
    if (!((heads >> j) & 1u)) continue;
    const uint32_t l = l0 + j;
    const K key = s_key[l];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t num = 0;
    uint32_t ll = l;
    do {  // members inside this tile: shared memory
      sx = __fadd_rn(sx, s_p[0][ll]);
      sy = __fadd_rn(sy, s_p[1][ll]);
      sz = __fadd_rn(sz, s_p[2][ll]);
      num++;
      ll++;
    } while (ll < tile_count && s_key[ll] == key);
    float vc[3];
    chunk_min(P, (long long)((unsigned long long)key >> P.key_bits), vc);
    if (ll == tile_count) {  // the voxel continues in the next tile(s): finish from global memory
      uint32_t g = tile_base + tile_count;
      while (g < n && keys[g] == key) {
        const float3 pt = load_xyz(v, vals[g]);
        sx = __fadd_rn(sx, __fsub_rn(pt.x, vc[0]));
        sy = __fadd_rn(sy, __fsub_rn(pt.y, vc[1]));
        sz = __fadd_rn(sz, __fsub_rn(pt.z, vc[2]));
        num++;
        g++;
      }
    }
    const uint32_t first = vals[tile_base + l];
    uint8_t* dst = out + rank * (uint64_t)v.stride;
    const uint8_t* src = v.data + (uint64_t)first * (uint64_t)v.stride;
    if (out_aligned) {
      const uint32_t* s4 = (const uint32_t*)src;
      uint32_t* d4 = (uint32_t*)dst;
      const int words = (int)(v.stride >> 2);
      for (int b = 0; b < words; b++) d4[b] = __ldg(s4 + b);
    } else {
      for (int64_t b = 0; b < v.stride; b++) dst[b] = src[b];
    }
    if (num > 1) {
      const float inv = __fdiv_rn(1.0f, (float)num);  // 1.0 / float32(n)   voxelgrid.go:179
      store_f32_any(dst + v.off[0], __fadd_rn(__fmul_rn(sx, inv), vc[0]), out_aligned);
      store_f32_any(dst + v.off[1], __fadd_rn(__fmul_rn(sy, inv), vc[1]), out_aligned);
      store_f32_any(dst + v.off[2], __fadd_rn(__fmul_rn(sz, inv), vc[2]), out_aligned);
    }
    rank++;
  }
}

template <typename K>
static void run_sorted_reduce(const CloudView& v, const VgParams& P, int total_bits, uint8_t* d_out,
                              long long* d_n_out, int* d_flags, cudaStream_t stream) {
  const uint32_t n = (uint32_t)v.n;
  DevBuf<K> keys0(n, stream), keys1(n, stream);
  DevBuf<uint32_t> vals0(n, stream), vals1(n, stream);
  rsort::Sorter<K> sorter;
  sorter.prepare(n, 0, total_bits, stream);
  const int kblocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(n, 256));
  PCG_LAUNCH((voxel_key_kernel<K>), kblocks, 256, 0, stream, v, P, keys0.p, d_flags, sorter.hist(), sorter.passes);
  K* kk[2] = {keys0.p, keys1.p};
  uint32_t* vbuf[2] = {vals0.p, vals1.p};
  int res = 0;
  sorter.run(kk, vbuf, /*identity_vals=*/true, /*keep_keys=*/true, stream, &res);
  const int tiles = div_up(n, kSegTile);
  DevBuf<unsigned long long> status((size_t)tiles + 1, stream);
  PCG_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), stream));
  uint32_t* counter = (uint32_t*)(status.p + tiles);
  PCG_LAUNCH((voxel_reduce_kernel<K>), tiles, kSegThreads, 0, stream, v, P, kk[res], vbuf[res], d_out, counter,
             status.p, d_n_out);
}

// ======================================================================================
// Fused path: the whole Filter as ONE cooperative kernel, one CTA per tile, tiles <= SMs.
// For clouds up to ~1.2M points every phase below is shorter than a kernel launch, so the
// multi-kernel pipeline above spends most of its time in launch gaps, workspace memsets and
// the host round trip for min/max.  Here the phases are separated by grid-wide barriers:
//   0  MinMaxVec3 of the tile -> 6 atomics                                   | grid.sync
//      every CTA derives the grid parameters itself (vg_make_params, device float32 ops)
//   1  voxel keys of the tile, kept in registers
//   2  per 8-bit digit: rank in shared memory (warp match-any), publish the tile's digit
//      counts                                                                 | grid.sync
//      offsets = sum of the counts of preceding tiles (independent loads, no look-back
//      chain), scatter through shared memory                                  | grid.sync
//   3  stage p = pt - vcMin of the tile's sorted slice in shared memory, count voxel heads,
//      publish                                                                | grid.sync
//      heads add their members in list order and write record + centroid.
// The host launches once and reads back 16 bytes.
namespace cg = cooperative_groups;

namespace fused {
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

struct Work {
  unsigned long long* acc;  // [6]: ~min / max packed (value bits, index), reduced with atomicMax; zero-initialised
  uint32_t* counts;         // [tiles][256]
  uint32_t* head_counts;    // [tiles]
  void* keys[2];            // n keys each (uint32 or uint64 depending on the bits needed)
  uint32_t* vals[2];
  long long* result;        // [0] = records written, [1] = status | flags << 8 ; zero-initialised
};

struct Args {
  float leaf[3];
  long long chunk[3];
};

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp /*[kWarps]*/, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) {
    uint32_t c = s_warp[w];
    if (w < (int)warp) base += c;
    tot += c;
  }
  __syncthreads();
  if (total) *total = tot;
  return base + incl - v;
}

struct Smem {
  uint32_t warp_hist[kWarps][rsort::kRadix];
  uint32_t digit_start[rsort::kRadix];
  uint32_t global_base[rsort::kRadix];
  uint32_t scan[kWarps];
  unsigned long long red[kWarps][6];
  VgParams P;
  int status;
  int total_bits;
  uint32_t prefix;
  uint32_t part[2][2][rsort::kRadix];  // [half][total|prefix][digit]
  float mm[6];
};

template <typename K, int IPT>
__device__ __forceinline__ void run(const CloudView& v, const Work& w, uint8_t* __restrict__ out, Smem& sm,
                                    unsigned char* dyn, cg::grid_group& grid) {
  constexpr int kTile = kThreads * IPT;
  const VgParams& P = sm.P;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = (uint32_t)v.n;
  const uint32_t tile = blockIdx.x, tiles = gridDim.x;
  const uint32_t tile_base = tile * (uint32_t)kTile;
  const uint32_t tile_count = min((uint32_t)kTile, n - tile_base);
  const uint32_t warp_base = tile_base + warp * (32u * IPT);
  K* kbuf[2] = {(K*)w.keys[0], (K*)w.keys[1]};

  // ---- phase 1: keys (warp-blocked, lane-striped positions: index order == (item, lane) order)
  K keys[IPT];
  uint32_t vals[IPT];
  int bad = 0;
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    const uint32_t pos = warp_base + i * 32 + lane;
    vals[i] = pos;
    keys[i] = pos < n ? (K)voxel_key_of(P, load_xyz(v, pos), &bad) : (K)0;
  }
  if (bad) atomicOr((unsigned long long*)&w.result[1], (unsigned long long)bad << 8);

  // ---- phase 2: stable LSD radix sort, one grid-wide exchange per digit
  K* s_keys = reinterpret_cast<K*>(dyn);
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(dyn + (size_t)kTile * sizeof(K));
  int cur = 0;
  for (int shift = 0; shift < sm.total_bits; shift += rsort::kRadixBits) {
    for (int i = tid; i < kWarps * rsort::kRadix; i += kThreads) (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t offs[IPT];
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const bool valid = (warp_base + i * 32 + lane) < n;
      const uint32_t d = valid ? rsort::digit_of(keys[i], shift) : (uint32_t)rsort::kRadix;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t pre = 0;
      if (valid && (int)lane == leader) {
        pre = sm.warp_hist[warp][d];
        sm.warp_hist[warp][d] = pre + __popc(peers);
      }
      pre = __shfl_sync(0xffffffffu, pre, leader);
      offs[i] = pre + rank;
      __syncwarp();
    }
    __syncthreads();
    uint32_t count = 0;
    if (tid < rsort::kRadix) {
#pragma unroll
      for (int wv = 0; wv < kWarps; wv++) {
        uint32_t c = sm.warp_hist[wv][tid];
        sm.warp_hist[wv][tid] = count;
        count += c;
      }
      w.counts[tile * rsort::kRadix + tid] = count;
    }
    grid.sync();
    // offsets of this tile = counts of the preceding tiles (independent loads: no look-back chain);
    // both halves of the CTA walk half of the tiles each
    uint32_t total = 0, prefix = 0;
    {
      const uint32_t d = tid & (rsort::kRadix - 1), half = tid >> 8;
      const uint32_t t0 = half ? tiles / 2 : 0, t1 = half ? tiles : tiles / 2;
      uint32_t tot = 0, pre = 0;
#pragma unroll 8
      for (uint32_t t = t0; t < t1; t++) {
        const uint32_t c = __ldcg(&w.counts[t * rsort::kRadix + d]);
        tot += c;
        pre += t < tile ? c : 0u;
      }
      sm.part[half][0][d] = tot;
      sm.part[half][1][d] = pre;
      __syncthreads();
      if (tid < rsort::kRadix) {
        total = sm.part[0][0][tid] + sm.part[1][0][tid];
        prefix = sm.part[0][1][tid] + sm.part[1][1][tid];
      }
    }
    const uint32_t digit_base = block_excl_scan(total, sm.scan, nullptr);
    const uint32_t dstart = block_excl_scan(count, sm.scan, nullptr);
    if (tid < rsort::kRadix) {
      sm.global_base[tid] = digit_base + prefix;
      sm.digit_start[tid] = dstart;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      if ((warp_base + i * 32 + lane) < n) {
        const uint32_t d = rsort::digit_of(keys[i], shift);
        const uint32_t pos = sm.digit_start[d] + sm.warp_hist[warp][d] + offs[i];
        s_keys[pos] = keys[i];
        s_vals[pos] = vals[i];
      }
    }
    __syncthreads();
    K* ko = kbuf[cur ^ 1];
    uint32_t* vo = w.vals[cur ^ 1];
    for (uint32_t s = tid; s < tile_count; s += kThreads) {
      const K k = s_keys[s];
      const uint32_t d = rsort::digit_of(k, shift);
      const uint32_t dst = sm.global_base[d] + (s - sm.digit_start[d]);
      ko[dst] = k;
      vo[dst] = s_vals[s];
    }
    grid.sync();
    cur ^= 1;
    if (shift + rsort::kRadixBits < sm.total_bits) {
#pragma unroll
      for (int i = 0; i < IPT; i++) {
        const uint32_t pos = warp_base + i * 32 + lane;
        if (pos < n) {
          keys[i] = __ldcg(&kbuf[cur][pos]);
          vals[i] = __ldcg(&w.vals[cur][pos]);
        }
      }
    }
  }
  const K* __restrict__ skeys = kbuf[cur];
  const uint32_t* __restrict__ svals = w.vals[cur];

  // ---- phase 3: segmented centroid (same arithmetic as voxel_reduce_kernel)
  K* s_key = reinterpret_cast<K*>(dyn);
  float* s_p = reinterpret_cast<float*>(dyn + (size_t)kTile * sizeof(K));  // [3][kTile]
  long long vc_cid = -1;  // sorted positions change chunk rarely: keep vcMin of the last chunk id
  float vc[3] = {0.f, 0.f, 0.f};
  for (uint32_t l = tid; l < tile_count; l += kThreads) {
    const K key = __ldcg(&skeys[tile_base + l]);
    const float3 pt = load_xyz(v, __ldcg(&svals[tile_base + l]));
    const long long cid = (long long)((unsigned long long)key >> P.key_bits);
    if (cid != vc_cid) {
      chunk_min(P, cid, vc);
      vc_cid = cid;
    }
    s_key[l] = key;
    s_p[l] = __fsub_rn(pt.x, vc[0]);
    s_p[kTile + l] = __fsub_rn(pt.y, vc[1]);
    s_p[2 * kTile + l] = __fsub_rn(pt.z, vc[2]);
  }
  __syncthreads();
  const uint32_t l0 = tid * IPT;
  K prev = 0;
  if (l0 > 0 && l0 - 1 < tile_count)
    prev = s_key[l0 - 1];
  else if (l0 == 0 && tile_base > 0)
    prev = __ldcg(&skeys[tile_base - 1]);
  uint32_t heads = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < IPT; j++) {
    const uint32_t l = l0 + j;
    if (l < tile_count) {
      const K k = s_key[l];
      const bool h = (tile_base + l == 0) || k != prev;
      heads |= (h ? 1u : 0u) << j;
      cnt += h ? 1u : 0u;
      prev = k;
    }
  }
  uint32_t total_heads = 0;
  const uint32_t excl = block_excl_scan(cnt, sm.scan, &total_heads);
  if (tid == 0) w.head_counts[tile] = total_heads;
  grid.sync();
  if (warp == 0) {
    uint32_t part = 0;
    for (uint32_t t = lane; t < tile; t += 32) part += __ldcg(&w.head_counts[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) {
      sm.prefix = part;
      if (tile == tiles - 1) w.result[0] = (long long)(part + total_heads);
    }
  }
  __syncthreads();
  uint64_t rank = (uint64_t)sm.prefix + excl;
  const int out_aligned = v.aligned && ((((uintptr_t)out) & 3) == 0);
#pragma unroll
  for (int j = 0; j < IPT; j++) {
    if (!((heads >> j) & 1u)) continue;
    const uint32_t l = l0 + j;
    const K key = s_key[l];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t num = 0;
    uint32_t ll = l;
    do {
      sx = __fadd_rn(sx, s_p[ll]);
      sy = __fadd_rn(sy, s_p[kTile + ll]);
      sz = __fadd_rn(sz, s_p[2 * kTile + ll]);
      num++;
      ll++;
    } while (ll < tile_count && s_key[ll] == key);
    {
      const long long cid = (long long)((unsigned long long)key >> P.key_bits);
      if (cid != vc_cid) {
        chunk_min(P, cid, vc);
        vc_cid = cid;
      }
    }
    if (ll == tile_count) {  // the voxel continues in the next tile(s)
      uint32_t g = tile_base + tile_count;
      while (g < n && __ldcg(&skeys[g]) == key) {
        const float3 pt = load_xyz(v, __ldcg(&svals[g]));
        sx = __fadd_rn(sx, __fsub_rn(pt.x, vc[0]));
        sy = __fadd_rn(sy, __fsub_rn(pt.y, vc[1]));
        sz = __fadd_rn(sz, __fsub_rn(pt.z, vc[2]));
        num++;
        g++;
      }
    }
    const uint32_t first = __ldcg(&svals[tile_base + l]);
    uint8_t* dst = out + rank * (uint64_t)v.stride;
    const uint8_t* src = v.data + (uint64_t)first * (uint64_t)v.stride;
    if (out_aligned) {
      const uint32_t* s4 = (const uint32_t*)src;
      uint32_t* d4 = (uint32_t*)dst;
      const int words = (int)(v.stride >> 2);
      for (int b = 0; b < words; b++) d4[b] = __ldg(s4 + b);
    } else {
      for (int64_t b = 0; b < v.stride; b++) dst[b] = src[b];
    }
    if (num > 1) {
      const float inv = __fdiv_rn(1.0f, (float)num);  // 1.0 / float32(n)   voxelgrid.go:179
      store_f32_any(dst + v.off[0], __fadd_rn(__fmul_rn(sx, inv), vc[0]), out_aligned);
      store_f32_any(dst + v.off[1], __fadd_rn(__fmul_rn(sy, inv), vc[1]), out_aligned);
      store_f32_any(dst + v.off[2], __fadd_rn(__fmul_rn(sz, inv), vc[2]), out_aligned);
    }
    rank++;
  }
}

template <int IPT>
__global__ void __launch_bounds__(kThreads, 1) voxelgrid_fused_kernel(CloudView v, Args args, Work w, uint8_t* out) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ Smem sm;
  constexpr int kTile = kThreads * IPT;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = (uint32_t)v.n;
  const uint32_t tile_base = blockIdx.x * (uint32_t)kTile;

  // ---- phase 0: MinMaxVec3 (pc/minmax.go:9-26), first occurrence wins (see minmax_kernel)
  unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0ull, 0ull, 0ull};
  for (int i = 0; i < IPT; i++) {
    const uint32_t pos = tile_base + i * kThreads + tid;
    if (pos < n) {
      const float3 p = load_xyz(v, pos);
      const float c[3] = {p.x, p.y, p.z};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (c[k] != c[k]) continue;
        const unsigned long long o = (unsigned long long)ordered_bits(c[k]) << 32;
        const unsigned long long a = o | pos, b = o | (0xffffffffu - pos);
        mn[k] = a < mn[k] ? a : mn[k];
        mx[k] = b > mx[k] ? b : mx[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long a = shfl_xor_u64(mn[k], d), b = shfl_xor_u64(mx[k], d);
      mn[k] = a < mn[k] ? a : mn[k];
      mx[k] = b > mx[k] ? b : mx[k];
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      sm.red[warp][k] = ~mn[k];  // min as max of the complement: one zero-initialised accumulator type
      sm.red[warp][3 + k] = mx[k];
    }
  }
  __syncthreads();
  if (tid < 6) {
    unsigned long long r = sm.red[0][tid];
    for (int wv = 1; wv < kWarps; wv++) r = sm.red[wv][tid] > r ? sm.red[wv][tid] : r;
    if (r != 0ull) atomicMax(&w.acc[tid], r);
  }
  grid.sync();
  if (tid < 6) {
    const int k = tid, c = k % 3;
    const float3 p0 = load_xyz(v, 0);
    const float first = c == 0 ? p0.x : (c == 1 ? p0.y : p0.z);
    float r = first;  // a NaN at point 0 is never replaced (minmax.go:13,17-22)
    if (first == first) {
      unsigned long long a = __ldcg(&w.acc[k]);
      if (k < 3) a = ~a;
      const uint32_t idx = k < 3 ? (uint32_t)a : 0xffffffffu - (uint32_t)a;
      const float3 p = load_xyz(v, idx);
      r = c == 0 ? p.x : (c == 1 ? p.y : p.z);
    }
    sm.mm[k] = r;
  }
  __syncthreads();
  if (tid == 0) {
    float mm[6];
    for (int k = 0; k < 6; k++) mm[k] = sm.mm[k];
    int tb = 0;
    VgParams P;
    const pcg_status st = vg_make_params(mm, mm + 3, args.leaf, args.chunk, &P, &tb);
    sm.status = st;
    if (st == PCG_OK) {
      sm.P = P;
      sm.total_bits = tb;
    } else if (blockIdx.x == 0) {
      w.result[1] = (long long)st;
    }
  }
  __syncthreads();
  if (sm.status != PCG_OK) return;  // every CTA computed the same status: uniform exit
  if (sm.total_bits <= 32)
    run<uint32_t, IPT>(v, w, out, sm, dyn, grid);
  else
    run<unsigned long long, IPT>(v, w, out, sm, dyn, grid);
}

template <int IPT>
constexpr size_t dyn_smem_bytes() {
  return (size_t)kThreads * IPT * 20;  // max(sort staging 12 B, reduce staging 8 + 12 B) per position
}

}  // namespace fused

// Largest cloud the fused kernel takes: one tile per SM.
static int64_t fused_capacity(int ipt) { return (int64_t)kNumSMs * fused::kThreads * ipt; }

template <int IPT>
static void launch_fused(const CloudView& v, const fused::Args& args, const fused::Work& w, uint8_t* d_out,
                         cudaStream_t stream) {
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  PCG_CUDA(cudaGetDevice(&dev));
  const size_t smem = fused::dyn_smem_bytes<IPT>();
  if (!(configured.load(std::memory_order_relaxed) & (1ull << dev))) {
    PCG_CUDA(cudaFuncSetAttribute(fused::voxelgrid_fused_kernel<IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  const int tiles = div_up(v.n, (int64_t)fused::kThreads * IPT);
  CloudView vv = v;
  fused::Args aa = args;
  fused::Work ww = w;
  uint8_t* oo = d_out;
  void* params[] = {&vv, &aa, &ww, &oo};
  const bool prof = g_profile.load(std::memory_order_relaxed) != 0;
  if (prof) prof_begin("voxelgrid_fused_kernel<IPT>", stream);
  PCG_CUDA(cudaLaunchCooperativeKernel((const void*)fused::voxelgrid_fused_kernel<IPT>, dim3(tiles), dim3(fused::kThreads),
                                       params, smem, stream));
  if (prof) prof_end(stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

static pcg_status voxelgrid_filter_fused(const CloudView& v, const float leaf[3], const int64_t chunk[3],
                                         uint8_t* d_out, int64_t* n_out, cudaStream_t stream) {
  const uint32_t n = (uint32_t)v.n;
  const int ipt = v.n <= fused_capacity(4) ? 4 : 16;
  const int tiles = div_up(v.n, (int64_t)fused::kThreads * ipt);
  // one allocation: [acc 6 x u64 | result 2 x i64 | counts | head_counts | keys0 | keys1 | vals0 | vals1]
  const size_t head_bytes = 64 + 16 + 48;
  const size_t counts_bytes = ((size_t)tiles * rsort::kRadix + tiles) * sizeof(uint32_t);
  const size_t keys_bytes = ((size_t)n * 8 + 255) & ~(size_t)255;
  const size_t vals_bytes = ((size_t)n * 4 + 255) & ~(size_t)255;
  const size_t counts_pad = (counts_bytes + 255) & ~(size_t)255;
  DevBuf<uint8_t> ws(128 + counts_pad + 2 * keys_bytes + 2 * vals_bytes, stream);
  (void)head_bytes;
  PCG_CUDA(cudaMemsetAsync(ws.p, 0, 128, stream));
  fused::Work w;
  w.acc = (unsigned long long*)ws.p;
  w.result = (long long*)(ws.p + 64);
  w.counts = (uint32_t*)(ws.p + 128);
  w.head_counts = w.counts + (size_t)tiles * rsort::kRadix;
  uint8_t* p = ws.p + 128 + counts_pad;
  w.keys[0] = p;
  w.keys[1] = p + keys_bytes;
  w.vals[0] = (uint32_t*)(p + 2 * keys_bytes);
  w.vals[1] = (uint32_t*)(p + 2 * keys_bytes + vals_bytes);
  fused::Args args;
  for (int k = 0; k < 3; k++) {
    args.leaf[k] = leaf[k];
    args.chunk[k] = (long long)chunk[k];
  }
  if (ipt == 4)
    launch_fused<4>(v, args, w, d_out, stream);
  else
    launch_fused<16>(v, args, w, d_out, stream);
  long long* h = (long long*)pinned_scratch();
  PCG_CUDA(cudaMemcpyAsync(h, w.result, 2 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  const pcg_status st = (pcg_status)(h[1] & 0xff);
  const int flags = (int)((h[1] >> 8) & 0xff);
  if (st != PCG_OK) throw StatusError{st, vg_status_message(st)};
  if (flags & kFlagUndefined)
    throw StatusError{PCG_E_REF_UNDEFINED, "a voxel coordinate is not finite / out of int64 range"};
  if (flags & kFlagPanic)
    throw StatusError{PCG_E_REF_WOULD_PANIC,
                      "reference would panic: voxel or chunk index out of range (voxelgrid.go:46,89,151)"};
  *n_out = h[0];
  return PCG_OK;
}

// Filter (voxelgrid.go:35-134).  `v` and d_out are device pointers.  Synchronises `stream`.
pcg_status voxelgrid_filter_device(const CloudView& v, const float leaf[3], const int64_t chunk[3], uint8_t* d_out,
                                   int64_t* n_out, cudaStream_t stream) {
  *n_out = 0;
  if (v.n == 0) throw StatusError{PCG_E_NO_POINT, "no point"};
  static const bool no_fused = getenv("PCG_VG_NO_FUSED") != nullptr;  // comparison runs only
  if (!no_fused && v.n <= fused_capacity(16)) {
    try {
      return voxelgrid_filter_fused(v, leaf, chunk, d_out, n_out, stream);
    } catch (const CudaError& e) {
      // a device that cannot co-schedule one CTA per tile (fewer SMs than a B200): multi-kernel path
      if (e.e != cudaErrorCooperativeLaunchTooLarge && e.e != cudaErrorNotSupported) throw;
      cudaGetLastError();
    }
  }
  float vmin[3], vmax[3];
  minmax_device(v, vmin, vmax, stream);

  VgParams P;
  int total_bits = 0;
  const long long chunk_ll[3] = {(long long)chunk[0], (long long)chunk[1], (long long)chunk[2]};
  const pcg_status prc = vg_make_params(vmin, vmax, leaf, chunk_ll, &P, &total_bits);
  if (prc != PCG_OK) throw StatusError{prc, vg_status_message(prc)};

  DevBuf<long long> d_n(1, stream);
  DevBuf<int> d_flags(1, stream);
  PCG_CUDA(cudaMemsetAsync(d_flags.p, 0, sizeof(int), stream));
  PCG_CUDA(cudaMemsetAsync(d_n.p, 0, sizeof(long long), stream));
  if (total_bits <= 32)
    run_sorted_reduce<uint32_t>(v, P, total_bits, d_out, d_n.p, d_flags.p, stream);
  else
    run_sorted_reduce<unsigned long long>(v, P, total_bits, d_out, d_n.p, d_flags.p, stream);
  long long* ph_n = (long long*)pinned_scratch();
  int* ph_flags = (int*)(pinned_scratch() + 16);
  PCG_CUDA(cudaMemcpyAsync(ph_n, d_n.p, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaMemcpyAsync(ph_flags, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  const long long h_n = *ph_n;
  const int h_flags = *ph_flags;
  if (h_flags & kFlagUndefined)
    throw StatusError{PCG_E_REF_UNDEFINED, "a voxel coordinate is not finite / out of int64 range"};
  if (h_flags & kFlagPanic)
    throw StatusError{PCG_E_REF_WOULD_PANIC,
                      "reference would panic: voxel or chunk index out of range (voxelgrid.go:46,89,151)"};
  *n_out = h_n;
  return PCG_OK;
}

}  // namespace pcg
