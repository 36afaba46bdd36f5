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
#include "common.cuh"
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

static inline bool go_int_host(float f, int64_t* out) {
  if (!(std::fabs(f) < 9.0e18f)) return false;
  *out = (int64_t)f;
  return true;
}

static int bits_for(int64_t count) {  // bits to represent values in [0, count)
  int b = 0;
  while (b < 63 && ((int64_t)1 << b) < count) b++;
  return b;
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

// voxelgrid.go:76-79,88 (vec2cid) and :149-151 (voxel key); the digit histograms of the
// sort are accumulated here so the keys are not read a second time.
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
      float3 pt = load_xyz(v, i);
      long long cid = 0;
      float vc[3] = {P.vmin[0], P.vmin[1], P.vmin[2]};
      if (P.chunked) {
        long long cx, cy, cz;
        bool ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.x, P.vmin[0]), P.chunk_size[0]), &cx);
        ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.y, P.vmin[1]), P.chunk_size[1]), &cy) && ok;
        ok = go_int_dev(__fdiv_rn(__fsub_rn(pt.z, P.vmin[2]), P.chunk_size[2]), &cz) && ok;
        if (!ok) {
          bad |= kFlagUndefined;
          cx = cy = cz = 0;
        }
        cid = ((cz * P.ny) + cy) * P.nx + cx;
        if (cid < 0 || cid >= P.n_chunks) {
          bad |= kFlagPanic;
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
        bad |= kFlagUndefined;
        x = y = z = 0;
      }
      long long key = x + P.xs * (y + P.ys * z);
      if (key < 0 || key >= P.n_voxels) {
        bad |= kFlagPanic;
        key = 0;
      }
      out_key = (K)(((unsigned long long)cid << P.key_bits) | (unsigned long long)key);
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

// Filter (voxelgrid.go:35-134).  `v` and d_out are device pointers.  Synchronises `stream`.
pcg_status voxelgrid_filter_device(const CloudView& v, const float leaf[3], const int64_t chunk[3], uint8_t* d_out,
                                   int64_t* n_out, cudaStream_t stream) {
  *n_out = 0;
  if (v.n == 0) throw StatusError{PCG_E_NO_POINT, "no point"};
  float vmin[3], vmax[3];
  minmax_device(v, vmin, vmax, stream);

  VgParams P;
  std::memset(&P, 0, sizeof(P));
  for (int k = 0; k < 3; k++) {
    P.vmin[k] = vmin[k];
    P.leaf[k] = leaf[k];
  }
  float size_for_grid[3];
  int64_t nz = 1;
  P.nx = P.ny = 1;
  P.n_chunks = 1;
  P.chunked = (chunk[0] * chunk[1] * chunk[2] != 0) ? 1 : 0;  // voxelgrid.go:45
  if (!P.chunked) {
    for (int k = 0; k < 3; k++) size_for_grid[k] = vmax[k];  // sic: voxelgrid.go:46 passes vMax as size
  } else {
    float size[3];
    for (int k = 0; k < 3; k++) {
      size[k] = vmax[k] - vmin[k];                      // :49
      float cs = leaf[k] * (float)chunk[k];             // :50-54
      float lim = size[k] + leaf[k];                    // :58
      if (cs > lim) cs = lim;
      P.chunk_size[k] = cs;
      size_for_grid[k] = cs;
    }
    int64_t c[3];
    for (int k = 0; k < 3; k++) {
      float q = size[k] / P.chunk_size[k];              // :62
      if (!go_int_host(q, &c[k])) throw StatusError{PCG_E_REF_UNDEFINED, "chunk grid size is not finite"};
      c[k] += 1;
    }
    P.nx = c[0];
    P.ny = c[1];
    nz = c[2];
    if (c[0] <= 0 || c[1] <= 0 || c[2] <= 0 || c[0] > (1ll << 40) / c[1] || c[0] * c[1] > (1ll << 40) / c[2])
      throw StatusError{PCG_E_TOO_LARGE, "chunk table too large"};
    P.n_chunks = c[0] * c[1] * nz;
  }
  int64_t s[3];
  for (int k = 0; k < 3; k++) {
    float q = size_for_grid[k] / leaf[k];               // :137
    if (!go_int_host(q, &s[k])) throw StatusError{PCG_E_REF_UNDEFINED, "voxel grid size is not finite"};
  }
  P.xs = s[0];
  P.ys = s[1];
  // nVoxels = (xs+1)*(ys+1)*(zs+1)  :138 ; a non-positive product leaves the dense array empty,
  // so the first indexed write panics.
  long double nv = (long double)(s[0] + 1) * (long double)(s[1] + 1) * (long double)(s[2] + 1);
  if (nv >= 9.0e18L || nv <= -9.0e18L) throw StatusError{PCG_E_TOO_LARGE, "voxel grid too large"};
  P.n_voxels = (s[0] + 1) * (s[1] + 1) * (s[2] + 1);
  if (P.n_voxels <= 0) throw StatusError{PCG_E_REF_WOULD_PANIC, "reference would index an empty voxel array"};
  P.key_bits = bits_for(P.n_voxels);
  int total_bits = P.key_bits + bits_for(P.n_chunks);
  if (total_bits > 64) throw StatusError{PCG_E_TOO_LARGE, "chunk id and voxel key do not fit 64 bits"};
  if (total_bits == 0) total_bits = 1;

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
