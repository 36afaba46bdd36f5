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

#include "bvh.cuh"
#include "icp_math.cuh"
#include "radix_sort.cuh"
#include "vg_common.cuh"
#include "vg_packed.cuh"

namespace pcg {

std::atomic<int> g_vg_path{0};

// ---- MinMaxVec3 (pc/minmax.go:9-26) ----------------------------------------------------------------------------------
// Go keeps the FIRST occurrence of the extreme value (strict comparisons).  Equal floats have equal bits except for
// the two zeros, so which occurrence wins only shows in the sign of a zero result.  The scan therefore keeps plain
// float minima / maxima (FMNMX ignores NaN operands like the reference's comparisons do) plus, per axis, the first
// index at which a zero occurs and that zero's sign - a handful of instructions per coordinate instead of a 64-bit
// (value, index) compare-and-select.  A CTA publishes six words (ordered value bits << 32 | low), low = (global index
// << 1 | sign) of the first zero when the extreme IS zero, else 0; 64-bit atomicMin / atomicMax combine them - over
// CTAs, and over ranks for a sharded cloud (minmax_shard_words_kernel).
struct MinMaxAcc {
  float mn[3], mx[3];
  uint32_t zc[3];  // (global index << 1 | sign bit) of the first zero seen, per axis
  __device__ __forceinline__ MinMaxAcc() {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      mn[k] = __int_as_float(0x7f800000);
      mx[k] = __int_as_float(0xff800000);
      zc[k] = 0xffffffffu;
    }
  }
  __device__ __forceinline__ void take(float c, int k, uint32_t g2) {  // g2 = global index << 1
    mn[k] = fminf(mn[k], c);
    mx[k] = fmaxf(mx[k], c);
    if (c == 0.0f) zc[k] = min(zc[k], g2 | (__float_as_uint(c) >> 31));
  }
  // CTA-wide reduction and the six atomics.  s_f / s_z: [warps][6] and [warps][3] scratch.
  // complement_min: the minima are combined as atomicMax of the complemented word (zero-initialised accumulators).
  __device__ __forceinline__ void publish(unsigned long long* __restrict__ out6, float (*s_f)[6], uint32_t (*s_z)[3],
                                          int warps, bool complement_min = false) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
        mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
        zc[k] = min(zc[k], __shfl_xor_sync(0xffffffffu, zc[k], d));
      }
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        s_f[warp][k] = mn[k];
        s_f[warp][3 + k] = mx[k];
        s_z[warp][k] = zc[k];
      }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      const int k = threadIdx.x, c = k % 3;
      float r = s_f[0][k];
      uint32_t z = s_z[0][c];
      for (int w = 1; w < warps; w++) {
        r = k < 3 ? fminf(r, s_f[w][k]) : fmaxf(r, s_f[w][k]);
        z = min(z, s_z[w][c]);
      }
      if (z != 0xffffffffu || r != 0.0f) {  // (a zero extreme always comes with the index of a zero)
        uint32_t low = 0;
        if (r == 0.0f) low = k < 3 ? z : ((0x7fffffffu - (z >> 1)) << 1) | (z & 1u);
        const unsigned long long w = ((unsigned long long)ordered_bits(r) << 32) | low;
        if (k >= 3)
          atomicMax(&out6[k], w);
        else if (complement_min)
          atomicMax(&out6[k], ~w);
        else
          atomicMin(&out6[k], w);
      }
    }
  }
};

__global__ void __launch_bounds__(256)
    minmax_kernel(CloudView v, uint32_t index_base, unsigned long long* __restrict__ out6) {
  __shared__ float s_f[8][6];
  __shared__ uint32_t s_z[8][3];
  MinMaxAcc acc;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += stride) {
    const float3 p = load_xyz(v, i);
    const uint32_t g2 = (index_base + (uint32_t)i) << 1;
    acc.take(p.x, 0, g2);
    acc.take(p.y, 1, g2);
    acc.take(p.z, 2, g2);
  }
  acc.publish(out6, s_f, s_z, 8);
}

// ---- MinMaxVec3 as a bulk-async pipeline (sm_100a data movement) -----------------------------------------------
// The scan above issues three 4-byte loads per point and is latency-bound.  Here the records travel global -> shared
// memory as 1-D bulk copies (cp.async.bulk, the TMA engine without a tensor map; UBLKCP in SASS) into a ring of
// kMmStages tiles, each completion counted in bytes on an mbarrier; one thread keeps the ring full while all 256
// threads reduce the tile that has landed.  Any record layout with 4-byte aligned x/y/z works: a tile is a contiguous
// run of whole records.
constexpr int kMmThreads = 256;
constexpr int kMmTilePoints = 1024;
constexpr int kMmStages = 4;
constexpr int kMmMaxStride = 32;  // bytes per record the shared-memory ring is sized for

__global__ void __launch_bounds__(kMmThreads)
    minmax_bulk_kernel(CloudView v, uint32_t index_base, unsigned long long* __restrict__ out6) {
  extern __shared__ __align__(128) unsigned char mm_dyn[];
  __shared__ __align__(8) uint64_t full[kMmStages];
  __shared__ float s_f[kMmThreads / 32][6];
  __shared__ uint32_t s_z[kMmThreads / 32][3];
  const uint32_t tid = threadIdx.x;
  const uint32_t stride = (uint32_t)v.stride;
  const uint32_t tile_bytes = kMmTilePoints * stride;
  const uint32_t n = (uint32_t)v.n;
  const uint32_t full_tiles = n / kMmTilePoints;
  MinMaxAcc acc;
  if (tid == 0) {
    for (int s = 0; s < kMmStages; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; the ring is primed with the first kMmStages of them
  const uint32_t my_tiles = full_tiles > blockIdx.x ? (full_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (uint32_t j = 0; j < my_tiles && j < (uint32_t)kMmStages; j++) {
      const uint64_t t = (uint64_t)blockIdx.x + (uint64_t)j * gridDim.x;
      mbar_expect_tx(&full[j], tile_bytes);
      bulk_g2s(mm_dyn + (size_t)j * tile_bytes, v.data + t * tile_bytes, tile_bytes, &full[j]);
    }
  }
  for (uint32_t j = 0; j < my_tiles; j++) {
    const uint32_t s = j % kMmStages, parity = (j / kMmStages) & 1u;
    mbar_wait(&full[s], parity);
    const unsigned char* tile = mm_dyn + (size_t)s * tile_bytes;
    const uint32_t first = index_base + (blockIdx.x + j * gridDim.x) * kMmTilePoints;
#pragma unroll
    for (int q = 0; q < kMmTilePoints / kMmThreads; q++) {
      const uint32_t p = q * kMmThreads + tid;
      const unsigned char* r = tile + (size_t)p * stride;
      const uint32_t g2 = (first + p) << 1;
      acc.take(*reinterpret_cast<const float*>(r + v.off[0]), 0, g2);
      acc.take(*reinterpret_cast<const float*>(r + v.off[1]), 1, g2);
      acc.take(*reinterpret_cast<const float*>(r + v.off[2]), 2, g2);
    }
    __syncthreads();  // everybody has read the slot: it can be refilled
    if (tid == 0 && j + kMmStages < my_tiles) {
      const uint64_t t = (uint64_t)blockIdx.x + (uint64_t)(j + kMmStages) * gridDim.x;
      mbar_expect_tx(&full[s], tile_bytes);
      bulk_g2s(mm_dyn + (size_t)s * tile_bytes, v.data + t * tile_bytes, tile_bytes, &full[s]);
    }
  }
  // the points after the last whole tile: plain loads, first CTA
  if (blockIdx.x == 0) {
    for (uint32_t i = full_tiles * kMmTilePoints + tid; i < n; i += kMmThreads) {
      const float3 p = load_xyz(v, i);
      const uint32_t g2 = (index_base + i) << 1;
      acc.take(p.x, 0, g2);
      acc.take(p.y, 1, g2);
      acc.take(p.z, 2, g2);
    }
  }
  acc.publish(out6, s_f, s_z, kMmThreads / 32);
}

// Launches the bulk-async scan when the layout allows it (4-byte aligned fields, records of up to 32 bytes, 16-byte
// aligned base), else the plain one.  acc: [3] minima initialised to ~0, [3] maxima initialised to 0.
static void launch_minmax(const CloudView& v, uint32_t index_base, unsigned long long* acc, cudaStream_t stream) {
  if (v.n <= 0) return;
  const bool bulk = v.aligned && v.stride <= kMmMaxStride && (((uintptr_t)v.data) & 15) == 0 && v.n >= 64 * kMmTilePoints;
  if (bulk) {
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    PCG_CUDA(cudaGetDevice(&dev));
    const size_t smem = (size_t)kMmStages * kMmTilePoints * (size_t)v.stride;
    if (!(configured.load(std::memory_order_relaxed) & (1ull << dev))) {
      PCG_CUDA(cudaFuncSetAttribute(minmax_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kMmStages * kMmTilePoints * kMmMaxStride));
      configured.fetch_or(1ull << dev, std::memory_order_relaxed);
    }
    const int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, v.n / kMmTilePoints);
    PCG_LAUNCH(minmax_bulk_kernel, blocks, kMmThreads, smem, stream, v, index_base, acc);
  } else {
    const int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(v.n, 256));
    PCG_LAUNCH(minmax_kernel, blocks, 256, 0, stream, v, index_base, acc);
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
    const unsigned long long w = in6[k];
    const uint32_t o = (uint32_t)(w >> 32);
    uint32_t bits = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    if (bits == 0u && (w & 1ull)) bits = 0x80000000u;  // the first zero was a -0
    r = __uint_as_float(bits);
  }
  out6[k] = r;
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
  launch_minmax(v, 0u, acc.p, stream);
  PCG_LAUNCH(minmax_finalize_kernel, 1, 32, 0, stream, v, acc.p, res.p);
  float* h = (float*)pinned_scratch();
  PCG_CUDA(cudaMemcpyAsync(h, res.p, 6 * sizeof(float), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  for (int k = 0; k < 3; k++) {
    mn[k] = h[k];
    mx[k] = h[3 + k];
  }
}


// Keys for the multi-kernel path; the digit histograms of the sort are accumulated here so the
// keys are not read a second time.
template <typename K>
__global__ void __launch_bounds__(256)
    voxel_key_kernel(CloudView v, VgParams P, K* __restrict__ keys, float4* __restrict__ xyz4,
                     int* __restrict__ flags, uint32_t* __restrict__ hist, int passes) {
  __shared__ uint32_t s_hist[rsort::kMaxPasses * rsort::kRadix];
  rsort::hist_zero(s_hist, passes);
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (v.n + stride - 1) / stride;
  int bad = 0;
  const KeyConsts C(P);
  for (int64_t r = 0; r < rounds; r++) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < v.n;
    K out_key = 0;
    if (valid) {
      const float3 pt = load_xyz(v, i);
      xyz4[i] = make_float4(pt.x, pt.y, pt.z, 0.f);  // aligned copy: the sorted gather is one 16-byte load per point
      out_key = (K)voxel_key_of(P, C, pt, &bad);
      keys[i] = out_key;
    }
    rsort::hist_add_key(s_hist, out_key, valid, 0, passes);
  }
  if (bad) atomicOr(flags, bad);
  __syncthreads();
  rsort::hist_flush(s_hist, hist, passes);
}

// MinMaxVec3 of one slice of a cloud that is split over several GPUs: six 64-bit words that ONE signed MIN all-reduce
// (what NCCL offers) combines across ranks.  Word k = (ordered value bits << 32 | global index << 1 | sign of a zero)
// for the minima; the maxima use the complemented index and are stored bitwise-negated, so MIN reduces them too; the
// top bit is flipped for the signed order.  The first occurrence of the extreme value wins, across ranks as well; NaN
// coordinates take no part (minmax.go:17-22) - except at global point 0, where the reference keeps the NaN for ever
// (minmax.go:13): the slice that starts at index 0 then publishes a marker no real value can produce.
__global__ void minmax_shard_init_kernel(unsigned long long* __restrict__ acc) {
  const int k = threadIdx.x;
  if (k < 6) acc[k] = k < 3 ? ~0ull : 0ull;
}
__global__ void minmax_shard_words_kernel(CloudView v, int first_slice, const unsigned long long* __restrict__ acc,
                                          long long* __restrict__ out6) {
  const int k = threadIdx.x;
  if (k >= 6) return;
  unsigned long long w = acc[k];
  if (first_slice && v.n > 0) {
    const float3 p0 = load_xyz(v, 0);
    const int c = k % 3;
    const float first = c == 0 ? p0.x : (c == 1 ? p0.y : p0.z);
    if (first != first) w = k < 3 ? 0ull : ~0ull;  // beats every candidate of every rank
  }
  if (k >= 3) w = ~w;
  out6[k] = (long long)(w ^ 0x8000000000000000ull);
}
void minmax_packed_device(const CloudView& v, uint32_t index_base, long long* d_out6, cudaStream_t stream) {
  DevBuf<unsigned long long> acc(6, stream);
  PCG_LAUNCH(minmax_shard_init_kernel, 1, 32, 0, stream, acc.p);
  launch_minmax(v, index_base, acc.p, stream);
  PCG_LAUNCH(minmax_shard_words_kernel, 1, 32, 0, stream, v, index_base == 0 ? 1 : 0, acc.p, d_out6);
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
                        const float4* __restrict__ xyz4, uint8_t* __restrict__ out, uint32_t* __restrict__ tile_counter,
                        unsigned long long* __restrict__ status, long long* __restrict__ n_out) {
  __shared__ float s_pt[3][kSegTile];  // raw points of the tile's sorted slice
  __shared__ K s_key[kSegTile];
  __shared__ uint16_t s_src[kSegTile];  // position of the tile's r-th voxel head
  __shared__ uint32_t s_scan[rsort::kWarps];
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_prefix;
  __shared__ struct {
    unsigned long long key, rank;
    float sx, sy, sz, fx, fy, fz, vc[3];
    uint32_t num, l;
    int pending;
  } s_cont;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = (uint32_t)v.n;
  if (tid == 0) {
    s_tile = atomicAdd(tile_counter, 1u);
    s_cont.pending = 0;
  }
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_base = tile * kSegTile;
  const uint32_t tile_count = min((uint32_t)kSegTile, n - tile_base);

  // Phase A (striped: coalesced key/index loads, independent 16-byte gathers)
  {
    K kk[kSegItems];
    uint32_t vv[kSegItems];
#pragma unroll
    for (int j = 0; j < kSegItems; j++) {
      const uint32_t l = j * kSegThreads + tid;
      if (l < tile_count) {
        kk[j] = keys[tile_base + l];
        vv[j] = vals[tile_base + l];
      }
    }
#pragma unroll
    for (int j = 0; j < kSegItems; j++) {
      const uint32_t l = j * kSegThreads + tid;
      if (l < tile_count) {
        const float4 pt = __ldg(&xyz4[vv[j]]);
        s_key[l] = kk[j];
        s_pt[0][l] = pt.x;
        s_pt[1][l] = pt.y;
        s_pt[2][l] = pt.z;
      }
    }
  }
  __syncthreads();

  // Phase B (blocked: thread t looks at positions 4t .. 4t+3): voxel heads, compacted
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
  {
    uint32_t r = excl;
#pragma unroll
    for (int j = 0; j < kSegItems; j++)
      if ((heads >> j) & 1u) s_src[r++] = (uint16_t)(l0 + j);
  }
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
  const int out_aligned = v.aligned && ((((uintptr_t)out) & 3) == 0);

  // voxelgrid.go:173-184: the first member's record, x/y/z replaced by the centroid when there are several members
  auto emit = [&](uint32_t l, uint32_t num, float sx, float sy, float sz, float fx, float fy, float fz,
                  const float* vcm, uint64_t rk) {
    const uint32_t first = vals[tile_base + l];
    uint8_t* dst = out + rk * (uint64_t)v.stride;
    const uint8_t* src = v.data + (uint64_t)first * (uint64_t)v.stride;
    if (out_aligned) {
      const uint32_t* s4 = (const uint32_t*)src;
      uint32_t* d4 = (uint32_t*)dst;
      const int words = (int)(v.stride >> 2);
      if (v.packed && words == 3) {  // xyz-only records: the point itself
        d4[0] = __float_as_uint(fx);
        d4[1] = __float_as_uint(fy);
        d4[2] = __float_as_uint(fz);
      } else {
        for (int b = 0; b < words; b++) d4[b] = __ldg(s4 + b);
      }
    } else {
      for (int64_t b = 0; b < v.stride; b++) dst[b] = src[b];
    }
    if (num > 1) {
      const float inv = __fdiv_rn(1.0f, (float)num);  // 1.0 / float32(n)   voxelgrid.go:179
      store_f32_any(dst + v.off[0], __fadd_rn(__fmul_rn(sx, inv), vcm[0]), out_aligned);
      store_f32_any(dst + v.off[1], __fadd_rn(__fmul_rn(sy, inv), vcm[1]), out_aligned);
      store_f32_any(dst + v.off[2], __fadd_rn(__fmul_rn(sz, inv), vcm[2]), out_aligned);
    }
  };

  // One voxel per thread per round (see the fused kernel): the members of a voxel are added in list order out of
  // shared memory (voxelgrid.go:148-158) - the float32 sum is the reference's.
  long long vc_cid = -1;
  float vc[3] = {0.f, 0.f, 0.f};
  for (uint32_t r = tid; r < total; r += kSegThreads) {
    const uint32_t l = s_src[r];
    const uint64_t rank = s_prefix + r;
    const K key = s_key[l];
    {
      const long long cid = (long long)((unsigned long long)key >> P.key_bits);
      if (cid != vc_cid) {
        chunk_min(P, cid, vc);
        vc_cid = cid;
      }
    }
    const float fx = s_pt[0][l], fy = s_pt[1][l], fz = s_pt[2][l];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t num = 0;
    uint32_t ll = l;
    do {
      sx = __fadd_rn(sx, __fsub_rn(s_pt[0][ll], vc[0]));
      sy = __fadd_rn(sy, __fsub_rn(s_pt[1][ll], vc[1]));
      sz = __fadd_rn(sz, __fsub_rn(s_pt[2][ll], vc[2]));
      num++;
      ll++;
    } while (ll < tile_count && s_key[ll] == key);
    if (ll == tile_count && tile_base + tile_count < n) {  // may run on into the next tile(s): finished by warp 0
      s_cont.key = (unsigned long long)key;
      s_cont.rank = rank;
      s_cont.sx = sx;
      s_cont.sy = sy;
      s_cont.sz = sz;
      s_cont.fx = fx;
      s_cont.fy = fy;
      s_cont.fz = fz;
      s_cont.vc[0] = vc[0];
      s_cont.vc[1] = vc[1];
      s_cont.vc[2] = vc[2];
      s_cont.num = num;
      s_cont.l = l;
      s_cont.pending = 1;
    } else {
      emit(l, num, sx, sy, sz, fx, fy, fz, vc, rank);
    }
  }
  __syncthreads();
  if (warp == 0 && s_cont.pending) {
    const K key = (K)s_cont.key;
    float sx = s_cont.sx, sy = s_cont.sy, sz = s_cont.sz;
    const float c0 = s_cont.vc[0], c1 = s_cont.vc[1], c2 = s_cont.vc[2];
    uint32_t num = s_cont.num;
    for (uint32_t g = tile_base + tile_count;; g += 32) {
      const uint32_t idx = g + lane;
      const bool match = idx < n && keys[idx] == key;
      const uint32_t m = __ballot_sync(0xffffffffu, match);
      const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;  // members are consecutive: the leading matches
      float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((int)lane < run) pt = __ldg(&xyz4[vals[idx]]);
      for (int q = 0; q < run; q++) {  // the additions stay in list order (every lane carries the same sums)
        sx = __fadd_rn(sx, __fsub_rn(__shfl_sync(0xffffffffu, pt.x, q), c0));
        sy = __fadd_rn(sy, __fsub_rn(__shfl_sync(0xffffffffu, pt.y, q), c1));
        sz = __fadd_rn(sz, __fsub_rn(__shfl_sync(0xffffffffu, pt.z, q), c2));
      }
      num += (uint32_t)run;
      if (run < 32) break;
    }
    if (lane == 0) emit(s_cont.l, num, sx, sy, sz, s_cont.fx, s_cont.fy, s_cont.fz, s_cont.vc, s_cont.rank);
  }
}

template <typename K>
static void run_sorted_reduce(const CloudView& v, const VgParams& P, int total_bits, uint8_t* d_out,
                              long long* d_n_out, int* d_flags, cudaStream_t stream) {
  const uint32_t n = (uint32_t)v.n;
  DevBuf<K> keys0(n, stream), keys1(n, stream);
  DevBuf<uint32_t> vals0(n, stream), vals1(n, stream);
  DevBuf<float4> xyz4(n, stream);
  rsort::Sorter<K> sorter;
  sorter.prepare(n, 0, total_bits, stream);
  const int kblocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(n, 256));
  PCG_LAUNCH((voxel_key_kernel<K>), kblocks, 256, 0, stream, v, P, keys0.p, xyz4.p, d_flags, sorter.hist(),
             sorter.passes);
  K* kk[2] = {keys0.p, keys1.p};
  uint32_t* vbuf[2] = {vals0.p, vals1.p};
  int res = 0;
  sorter.run(kk, vbuf, /*identity_vals=*/true, /*keep_keys=*/true, stream, &res);
  const int tiles = div_up(n, kSegTile);
  DevBuf<unsigned long long> status((size_t)tiles + 1, stream);
  PCG_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), stream));
  uint32_t* counter = (uint32_t*)(status.p + tiles);
  PCG_LAUNCH((voxel_reduce_kernel<K>), tiles, kSegThreads, 0, stream, v, P, kk[res], vbuf[res], xyz4.p, d_out,
             counter, status.p, d_n_out);
}

// The multi-kernel pipeline for a whole cloud: packed 64-bit words (vg_packed.cuh) whenever key + index bits fit 64,
// else (key, index) pairs.  Test hook g_vg_path == 2 forces the pairs.
static void run_pipeline(const CloudView& v, const VgParams& P, int total_bits, uint8_t* d_out, long long* d_n_out,
                         int* d_flags, cudaStream_t stream) {
  if (g_vg_path.load(std::memory_order_relaxed) != 2 && vgp::fits(v.n, total_bits))
    vgp::run(v, P, total_bits, d_out, d_n_out, d_flags, stream);
  else if (total_bits <= 32)
    run_sorted_reduce<uint32_t>(v, P, total_bits, d_out, d_n_out, d_flags, stream);
  else
    run_sorted_reduce<unsigned long long>(v, P, total_bits, d_out, d_n_out, d_flags, stream);
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

#ifdef PCG_VG_TIMING
__device__ unsigned long long g_vg_stamps[64];
__device__ unsigned long long g_vg_max[64];  // the same stamps, latest CTA
__device__ unsigned long long g_vg_tile[3][160];  // per tile: reduce start, reduce end, voxel heads
__device__ int g_vg_nstamps;
#define PCG_VG_STAMP()                                                          \
  do {                                                                          \
    if (threadIdx.x == 0) {                                                     \
      unsigned long long t__;                                                   \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                   \
      atomicMax(&g_vg_max[sm.stamp_i], t__);                                  \
      if (blockIdx.x == 0 && g_vg_nstamps < 64) g_vg_stamps[g_vg_nstamps++] = t__; \
      sm.stamp_i++;                                                           \
    }                                                                           \
  } while (0)
#else
#define PCG_VG_STAMP() \
  do {                 \
  } while (0)
#endif

namespace fused {
#ifndef PCG_VG_THREADS
#define PCG_VG_THREADS 512
#endif
constexpr int kThreads = PCG_VG_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kParts = kThreads / (rsort::kRadix / 4);  // groups of 64 threads (4 digits each) that share the walk over the tiles
constexpr int kIptSmall = 2048 / kThreads;         // tile of 2048 points (clouds up to 148 * 2048)
constexpr int kIptLarge = 8192 / kThreads;         // tile of 8192 points

template <int IPT>
__host__ __device__ constexpr size_t dyn_smem_bytes() {
  // max(sort staging 12 B, reduce staging (8 + 12) B with one pad slot per IPT positions) per position,
  // + 2 B per position for the compacted list of voxel heads
  return (size_t)(kThreads * IPT + kThreads) * 20 + (size_t)kThreads * IPT * 2;
}

struct Work {
  unsigned long long* acc;  // [6]: ~min / max words of MinMaxAcc::publish, reduced with atomicMax; zero-initialised
  uint32_t* counts;         // [tiles][256]
  uint32_t* head_counts;    // [tiles]
  void* keys[2];            // n keys each (uint32 or uint64 depending on the bits needed)
  uint32_t* vals[2];
  long long* result;        // PINNED HOST memory (mapped): [0] = records written, [1] = status | flags << 8; the
                            // kernel stores there directly, so the host needs no copy after the launch
  unsigned long long* flags;  // device: kFlag* bits raised while the keys are computed; zero-initialised
  float4* xyz4;             // n points as aligned {x, y, z, -}: the sorted gather of phase 3 is one 16-byte load per
                            // point instead of three 4-byte loads scattered over two sectors
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
  float red_f[kWarps][6];
  uint32_t red_z[kWarps][3];
  VgParams P;
  int status;
  int total_bits;
  uint32_t prefix;
  __align__(16) uint32_t part[kParts][2][rsort::kRadix];  // [part][total|prefix][digit]
  float mm[6];
  __align__(8) uint64_t tile_bar;  // bulk copy of the tile's records into shared memory (phase 0)
  int tile_in_smem;
  // the tile's last voxel when it runs on into the next tile(s): finished by warp 0 with parallel loads
  struct {
    unsigned long long key, rank;
    float sx, sy, sz, fx, fy, fz, vc[3];
    uint32_t num, l;
    int pending;
  } cont;
#ifdef PCG_VG_TIMING
  int stamp_i;  // thread 0's running stamp index
#endif
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
  {
    // The key arithmetic is a few hundred instructions: kept in a ROLLED loop (its code is fetched once and
    // reused IPT times - fully unrolled it was instruction-fetch bound), results parked in shared memory.
    K* s_park = reinterpret_cast<K*>(dyn);
    bool general = false;
    const KeyConsts C(P);
#pragma unroll 2
    for (int i = 0; i < IPT; i++) {
      const uint32_t pos = warp_base + i * 32 + lane;
      unsigned long long k = ~0ull;
      if (pos < n) {
        float3 pt;
        if (sm.tile_in_smem) {
          const unsigned char* r = dyn + dyn_smem_bytes<IPT>() - (size_t)kTile * (size_t)v.stride +
                                   (size_t)(pos - tile_base) * (size_t)v.stride;
          pt = make_float3(*reinterpret_cast<const float*>(r + v.off[0]), *reinterpret_cast<const float*>(r + v.off[1]),
                           *reinterpret_cast<const float*>(r + v.off[2]));
        } else {
          pt = load_xyz(v, pos);
        }
        w.xyz4[pos] = make_float4(pt.x, pt.y, pt.z, 0.f);
        if (!voxel_key_fast(C, pt, &k)) {
          general = true;
          k = ~0ull;
        }
      }
      s_park[warp * (32 * IPT) + i * 32 + lane] = (K)k;
    }
    if (__any_sync(0xffffffffu, general)) {  // stragglers (out-of-range / non-finite coordinates): general path
#pragma unroll 1
      for (int i = 0; i < IPT; i++) {
        const uint32_t pos = warp_base + i * 32 + lane;
        if (pos < n && s_park[warp * (32 * IPT) + i * 32 + lane] == (K)~0ull) {
          unsigned long long k;
          if (!voxel_key_fast(C, load_xyz(v, pos), &k)) k = voxel_key_general(P, load_xyz(v, pos), &bad);
          s_park[warp * (32 * IPT) + i * 32 + lane] = (K)k;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const uint32_t pos = warp_base + i * 32 + lane;
      vals[i] = pos;
      keys[i] = pos < n ? s_park[warp * (32 * IPT) + i * 32 + lane] : (K)0;
    }
    __syncthreads();  // the parking area is the sort's staging buffer
  }
  if (bad) atomicOr(w.flags, (unsigned long long)bad);
  PCG_VG_STAMP();  // keys done

  // ---- phase 2: stable LSD radix sort, one grid-wide exchange per digit
  K* s_keys = reinterpret_cast<K*>(dyn);
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(dyn + (size_t)kTile * sizeof(K));
  int cur = 0;
  for (int shift = 0; shift < sm.total_bits; shift += rsort::kRadixBits) {
    for (int i = tid; i < kWarps * rsort::kRadix; i += kThreads) (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t offs[IPT];
    uint32_t peers_of[IPT];
#pragma unroll
    for (int i = 0; i < IPT; i++) {  // the match instructions are independent: issue them back to back
      const bool valid = (warp_base + i * 32 + lane) < n;
      const uint32_t d = valid ? rsort::digit_of(keys[i], shift) : (uint32_t)rsort::kRadix;
      peers_of[i] = __match_any_sync(0xffffffffu, d);
    }
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const bool valid = (warp_base + i * 32 + lane) < n;
      const uint32_t d = valid ? rsort::digit_of(keys[i], shift) : (uint32_t)rsort::kRadix;
      const uint32_t peers = peers_of[i];
      const int leader = __ffs(peers) - 1;
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t pre = 0;
      // atomics of one warp on one address execute in program order: the returned counts need no barrier between
      // the rounds, and the rounds overlap instead of waiting for a load-add-store each
      if (valid && (int)lane == leader) pre = atomicAdd(&sm.warp_hist[warp][d], (uint32_t)__popc(peers));
      pre = __shfl_sync(0xffffffffu, pre, leader);
      offs[i] = pre + rank;
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
    PCG_VG_STAMP();  // ranked
    grid.sync();
    PCG_VG_STAMP();  // sync A
    // offsets of this tile = counts of the preceding tiles (independent loads: no look-back chain);
    // every group of 256 threads walks its share of the tiles
    uint32_t total = 0, prefix = 0;
    {
      // thread = (4 consecutive digits, one share of the tiles): 16-byte loads, all of them independent
      const uint32_t d4 = (tid & 63u) * 4u, part = tid >> 6;
      const uint32_t t0 = part * tiles / kParts, t1 = (part + 1) * tiles / kParts;
      uint4 tot = make_uint4(0, 0, 0, 0), pre = make_uint4(0, 0, 0, 0);
      // every CTA reads the same table at the same moment: start at a CTA-specific tile so that the
      // requests spread over the L2 slices (integer sums: the order is free)
      const uint32_t len = t1 - t0;
      uint32_t t = len ? t0 + tile % len : t0;
#pragma unroll 4
      for (uint32_t k = 0; k < len; k++) {
        const uint4 c = __ldcg(reinterpret_cast<const uint4*>(&w.counts[t * rsort::kRadix + d4]));
        tot.x += c.x;
        tot.y += c.y;
        tot.z += c.z;
        tot.w += c.w;
        if (t < tile) {
          pre.x += c.x;
          pre.y += c.y;
          pre.z += c.z;
          pre.w += c.w;
        }
        t = t + 1 == t1 ? t0 : t + 1;
      }
      *reinterpret_cast<uint4*>(&sm.part[part][0][d4]) = tot;
      *reinterpret_cast<uint4*>(&sm.part[part][1][d4]) = pre;
      __syncthreads();
      if (tid < rsort::kRadix) {
#pragma unroll
        for (int k = 0; k < kParts; k++) {
          total += sm.part[k][0][tid];
          prefix += sm.part[k][1][tid];
        }
      }
    }
    PCG_VG_STAMP();  // walked
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
    PCG_VG_STAMP();  // scattered
    grid.sync();
    PCG_VG_STAMP();  // sync B
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
  PCG_VG_STAMP();  // sort done (incl. reload)
  const K* __restrict__ skeys = kbuf[cur];
  const uint32_t* __restrict__ svals = w.vals[cur];

  // ---- phase 3: segmented centroid (same arithmetic as voxel_reduce_kernel)
  // Staging: key + raw point of every sorted position (independent gathers, all in flight at once).
  // Thread t later walks positions t*IPT .. t*IPT+IPT-1: one pad word per IPT positions keeps the lanes of a warp
  // on different banks (stride IPT would put them all on one or two).
  constexpr int kPadTile = kTile + kTile / IPT;
  auto pad = [](uint32_t l) { return l + l / (uint32_t)IPT; };
  K* s_key = reinterpret_cast<K*>(dyn);
  float* s_pt = reinterpret_cast<float*>(dyn + (size_t)kPadTile * sizeof(K));  // [3][kPadTile]
  {
    K kk[IPT];
    uint32_t vv[IPT];
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const uint32_t l = i * kThreads + tid;
      if (l < tile_count) {
        kk[i] = __ldcg(&skeys[tile_base + l]);
        vv[i] = __ldcg(&svals[tile_base + l]);
      }
    }
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const uint32_t l = i * kThreads + tid;
      if (l < tile_count) {
        const float4 pt = __ldcg(&w.xyz4[vv[i]]);
        const uint32_t pl = pad(l);
        s_key[pl] = kk[i];
        s_pt[pl] = pt.x;
        s_pt[kPadTile + pl] = pt.y;
        s_pt[2 * kPadTile + pl] = pt.z;
      }
    }
  }
  __syncthreads();
  const uint32_t l0 = tid * IPT;
  K prev = 0;
  if (l0 > 0 && l0 - 1 < tile_count)
    prev = s_key[pad(l0 - 1)];
  else if (l0 == 0 && tile_base > 0)
    prev = __ldcg(&skeys[tile_base - 1]);
  uint32_t heads = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < IPT; j++) {
    const uint32_t l = l0 + j;
    if (l < tile_count) {
      const K k = s_key[pad(l)];
      const bool h = (tile_base + l == 0) || k != prev;
      heads |= (h ? 1u : 0u) << j;
      cnt += h ? 1u : 0u;
      prev = k;
    }
  }
  uint32_t total_heads = 0;
  const uint32_t excl = block_excl_scan(cnt, sm.scan, &total_heads);
  if (tid == 0) w.head_counts[tile] = total_heads;
  PCG_VG_STAMP();  // staged + heads
  grid.sync();
  PCG_VG_STAMP();  // sync C
#ifdef PCG_VG_TIMING
  if (tid == 0 && tile < 160) {
    unsigned long long t__;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));
    g_vg_tile[0][tile] = t__;
    g_vg_tile[2][tile] = total_heads;
  }
#endif
  if (warp == 0) {
    uint32_t part = 0;
    for (uint32_t t = lane; t < tile; t += 32) part += __ldcg(&w.head_counts[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) {
      sm.prefix = part;
      sm.cont.pending = 0;
      if (tile == tiles - 1) {  // the flags were raised before the sort's grid-wide barriers
        w.result[1] = (long long)(__ldcg(w.flags) << 8);
        w.result[0] = (long long)(part + total_heads);
      }
    }
  }
  __syncthreads();
  const int out_aligned = v.aligned && ((((uintptr_t)out) & 3) == 0);
  // records that are exactly x,y,z (pc.Vec3Slice / xyz-only PCD): the first member's record is its point, already
  // in shared memory - no gather from the input
  const bool xyz_only = out_aligned && v.packed && v.stride == 12;
  long long vc_cid = -1;  // sorted positions change chunk rarely: keep vcMin of the last chunk id
  float vc[3] = {0.f, 0.f, 0.f};
  // voxelgrid.go:173-184: the first member's record, x/y/z replaced by the centroid when there are several members
  auto emit = [&](uint32_t l, uint32_t num, float sx, float sy, float sz, float fx, float fy, float fz,
                  const float* vcm, uint64_t rk) {
    float ox = fx, oy = fy, oz = fz;  // num == 1: the original bytes (voxelgrid.go:176-178)
    if (num > 1) {
      const float inv = __fdiv_rn(1.0f, (float)num);  // 1.0 / float32(n)   voxelgrid.go:179
      ox = __fadd_rn(__fmul_rn(sx, inv), vcm[0]);
      oy = __fadd_rn(__fmul_rn(sy, inv), vcm[1]);
      oz = __fadd_rn(__fmul_rn(sz, inv), vcm[2]);
    }
    uint8_t* dst = out + rk * (uint64_t)v.stride;
    if (xyz_only) {
      float* d3 = reinterpret_cast<float*>(dst);
      d3[0] = ox;
      d3[1] = oy;
      d3[2] = oz;
    } else {
      const uint32_t first = __ldcg(&svals[tile_base + l]);
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
        store_f32_any(dst + v.off[0], ox, out_aligned);
        store_f32_any(dst + v.off[1], oy, out_aligned);
        store_f32_any(dst + v.off[2], oz, out_aligned);
      }
    }
  };
  // Voxels differ a lot in size (a few hundred members next to the sensor, one far away) and the members of one
  // voxel must be added one after the other.  Walking "my 16 positions" would make a warp wait, 16 times over,
  // for its longest voxel; instead the heads are compacted (s_src[r] = position of the tile's r-th voxel) and
  // voxel r goes to thread r mod kThreads: a tile of few large voxels is one short round, and neighbouring
  // lanes write neighbouring records.
  uint16_t* s_src = reinterpret_cast<uint16_t*>(dyn + (size_t)kPadTile * 20);
  {
    uint32_t r = excl;
#pragma unroll
    for (int j = 0; j < IPT; j++)
      if ((heads >> j) & 1u) s_src[r++] = (uint16_t)(l0 + j);
  }
  __syncthreads();
  for (uint32_t r = tid; r < total_heads; r += kThreads) {
    const uint32_t l = s_src[r];
    const uint64_t rank = (uint64_t)sm.prefix + r;
    const K key = s_key[pad(l)];
    {
      const long long cid = (long long)((unsigned long long)key >> P.key_bits);
      if (cid != vc_cid) {
        chunk_min(P, cid, vc);
        vc_cid = cid;
      }
    }
    const float fx = s_pt[pad(l)], fy = s_pt[kPadTile + pad(l)], fz = s_pt[2 * kPadTile + pad(l)];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t num = 0;
    uint32_t ll = l;
    do {  // p = pt - vcMin, sum += p in list order (voxelgrid.go:148-158)
      const uint32_t pl = pad(ll);
      sx = __fadd_rn(sx, __fsub_rn(s_pt[pl], vc[0]));
      sy = __fadd_rn(sy, __fsub_rn(s_pt[kPadTile + pl], vc[1]));
      sz = __fadd_rn(sz, __fsub_rn(s_pt[2 * kPadTile + pl], vc[2]));
      num++;
      ll++;
    } while (ll < tile_count && s_key[pad(ll)] == key);
    if (ll == tile_count && tile_base + tile_count < n) {
      // the tile's last voxel may run on into the next tile(s): one thread walking those members would chain
      // dependent global loads (key -> index -> point) - warp 0 finishes it below with the loads side by side
      sm.cont.key = (unsigned long long)key;
      sm.cont.rank = rank;
      sm.cont.sx = sx;
      sm.cont.sy = sy;
      sm.cont.sz = sz;
      sm.cont.fx = fx;
      sm.cont.fy = fy;
      sm.cont.fz = fz;
      sm.cont.vc[0] = vc[0];
      sm.cont.vc[1] = vc[1];
      sm.cont.vc[2] = vc[2];
      sm.cont.num = num;
      sm.cont.l = l;
      sm.cont.pending = 1;
    } else {
      emit(l, num, sx, sy, sz, fx, fy, fz, vc, rank);
    }
  }
  __syncthreads();
  if (warp == 0 && sm.cont.pending) {
    const K key = (K)sm.cont.key;
    float sx = sm.cont.sx, sy = sm.cont.sy, sz = sm.cont.sz;
    const float c0 = sm.cont.vc[0], c1 = sm.cont.vc[1], c2 = sm.cont.vc[2];
    uint32_t num = sm.cont.num;
    for (uint32_t g = tile_base + tile_count;; g += 32) {
      const uint32_t idx = g + lane;
      const bool match = idx < n && __ldcg(&skeys[idx]) == key;
      const uint32_t m = __ballot_sync(0xffffffffu, match);
      const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;  // members are consecutive: the leading matches
      float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((int)lane < run) pt = __ldcg(&w.xyz4[__ldcg(&svals[idx])]);
      for (int q = 0; q < run; q++) {  // the additions stay in list order (every lane carries the same sums)
        sx = __fadd_rn(sx, __fsub_rn(__shfl_sync(0xffffffffu, pt.x, q), c0));
        sy = __fadd_rn(sy, __fsub_rn(__shfl_sync(0xffffffffu, pt.y, q), c1));
        sz = __fadd_rn(sz, __fsub_rn(__shfl_sync(0xffffffffu, pt.z, q), c2));
      }
      num += (uint32_t)run;
      if (run < 32) break;
    }
    if (lane == 0) emit(sm.cont.l, num, sx, sy, sz, sm.cont.fx, sm.cont.fy, sm.cont.fz, sm.cont.vc, sm.cont.rank);
  }
  PCG_VG_STAMP();  // reduced
#ifdef PCG_VG_TIMING
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t__;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));
    atomicMax(&g_vg_stamps[63], t__);
    if (tile < 160) g_vg_tile[1][tile] = t__;
  }
#endif
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

#ifdef PCG_VG_TIMING
  if (threadIdx.x == 0) sm.stamp_i = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    g_vg_nstamps = 0;
    g_vg_stamps[63] = 0;
  }
#endif
  PCG_VG_STAMP();  // start
  // The tile's records arrive in shared memory as ONE bulk copy (cp.async.bulk, completion counted on an mbarrier):
  // min/max and the key arithmetic then read them from there instead of issuing three 4-byte global loads per point
  // twice.  The tile sits at the END of the dynamic buffer (the key parking area of phase 1 uses its start).  Needs
  // whole tiles of 4-byte aligned records that fit beside the parking area; otherwise the plain loads.
  const uint32_t tile_bytes = (uint32_t)kTile * (uint32_t)v.stride;
  unsigned char* s_tile = dyn + dyn_smem_bytes<IPT>() - tile_bytes;
  if (tid == 0) {
    const bool ok = v.aligned && (((uintptr_t)v.data) & 15) == 0 && tile_base + (uint32_t)kTile <= n &&
                    (size_t)tile_bytes + (size_t)kTile * 8 <= dyn_smem_bytes<IPT>();
    sm.tile_in_smem = ok ? 1 : 0;
    if (ok) {
      mbar_init(&sm.tile_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(&sm.tile_bar, tile_bytes);
      bulk_g2s(s_tile, v.data + (size_t)tile_base * (size_t)v.stride, tile_bytes, &sm.tile_bar);
    }
  }
  __syncthreads();
  const bool tile_in_smem = sm.tile_in_smem != 0;
  if (tile_in_smem) mbar_wait(&sm.tile_bar, 0);
  auto tile_xyz = [&](uint32_t local) {
    const unsigned char* r = s_tile + (size_t)local * (size_t)v.stride;
    return make_float3(*reinterpret_cast<const float*>(r + v.off[0]), *reinterpret_cast<const float*>(r + v.off[1]),
                       *reinterpret_cast<const float*>(r + v.off[2]));
  };
  // ---- phase 0: MinMaxVec3 (pc/minmax.go:9-26), first occurrence wins (see minmax_kernel)
  {
    MinMaxAcc acc;
#pragma unroll 4
    for (int i = 0; i < IPT; i++) {
      const uint32_t pos = tile_base + i * kThreads + tid;
      if (pos < n) {
        const float3 p = tile_in_smem ? tile_xyz(i * kThreads + tid) : load_xyz(v, pos);
        acc.take(p.x, 0, pos << 1);
        acc.take(p.y, 1, pos << 1);
        acc.take(p.z, 2, pos << 1);
      }
    }
    acc.publish(w.acc, sm.red_f, sm.red_z, kWarps, /*complement_min=*/true);
  }
  PCG_VG_STAMP();  // minmax local
  grid.sync();
  PCG_VG_STAMP();  // sync 0
  if (tid < 6) {
    const int k = tid, c = k % 3;
    const float3 p0 = load_xyz(v, 0);
    const float first = c == 0 ? p0.x : (c == 1 ? p0.y : p0.z);
    float r = first;  // a NaN at point 0 is never replaced (minmax.go:13,17-22)
    if (first == first) {
      unsigned long long a = __ldcg(&w.acc[k]);
      if (k < 3) a = ~a;
      const uint32_t o = (uint32_t)(a >> 32);
      uint32_t bits = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
      if (bits == 0u && (a & 1ull)) bits = 0x80000000u;  // the first zero was a -0
      r = __uint_as_float(bits);
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
  PCG_VG_STAMP();  // params
  if (sm.status != PCG_OK) return;  // every CTA computed the same status: uniform exit
  if (sm.total_bits <= 32)
    run<uint32_t, IPT>(v, w, out, sm, dyn, grid);
  else
    run<unsigned long long, IPT>(v, w, out, sm, dyn, grid);
}

}  // namespace fused

#ifdef PCG_VG_TIMING
}  // namespace pcg
extern "C" int pcg_debug_vg_stamps(unsigned long long* out) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, pcg::g_vg_nstamps, sizeof(int));
  cudaMemcpyFromSymbol(out, pcg::g_vg_stamps, sizeof(unsigned long long) * 64);
  cudaMemcpyFromSymbol(out + 64, pcg::g_vg_max, sizeof(unsigned long long) * 64);
  cudaMemcpyFromSymbol(out + 128, pcg::g_vg_tile, sizeof(unsigned long long) * 480);
  unsigned long long z[64] = {0};
  cudaMemcpyToSymbol(pcg::g_vg_max, z, sizeof(z));
  return n;
}
namespace pcg {
#endif

// Largest cloud the fused kernel takes: one tile per SM.
static int64_t fused_capacity(int ipt) { return (int64_t)kNumSMs * fused::kThreads * ipt; }
constexpr int kFusedIptSmall = fused::kIptSmall, kFusedIptLarge = fused::kIptLarge;

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
  const int ipt = v.n <= fused_capacity(kFusedIptSmall) ? kFusedIptSmall : kFusedIptLarge;
  const int tiles = div_up(v.n, (int64_t)fused::kThreads * ipt);
  // one allocation: [acc 6 x u64 | flags | counts | head_counts | keys0 | keys1 | vals0 | vals1 | xyz4]
  const size_t head_bytes = 64 + 16 + 48;
  const size_t counts_bytes = ((size_t)tiles * rsort::kRadix + tiles) * sizeof(uint32_t);
  const size_t keys_bytes = ((size_t)n * 8 + 255) & ~(size_t)255;
  const size_t vals_bytes = ((size_t)n * 4 + 255) & ~(size_t)255;
  const size_t counts_pad = (counts_bytes + 255) & ~(size_t)255;
  DevBuf<uint8_t> ws(128 + counts_pad + 2 * keys_bytes + 2 * vals_bytes + (size_t)n * sizeof(float4), stream);
  (void)head_bytes;
  PCG_CUDA(cudaMemsetAsync(ws.p, 0, 128, stream));
  fused::Work w;
  w.acc = (unsigned long long*)ws.p;
  w.flags = (unsigned long long*)(ws.p + 64);
  long long* h = (long long*)pinned_scratch();  // cudaMallocHost memory: addressable from the device (UVA)
  h[0] = 0;
  h[1] = 0;
  w.result = h;
  w.counts = (uint32_t*)(ws.p + 128);
  w.head_counts = w.counts + (size_t)tiles * rsort::kRadix;
  uint8_t* p = ws.p + 128 + counts_pad;
  w.keys[0] = p;
  w.keys[1] = p + keys_bytes;
  w.vals[0] = (uint32_t*)(p + 2 * keys_bytes);
  w.vals[1] = (uint32_t*)(p + 2 * keys_bytes + vals_bytes);
  w.xyz4 = (float4*)(p + 2 * keys_bytes + 2 * vals_bytes);  // 256-byte aligned like the blocks before it
  fused::Args args;
  for (int k = 0; k < 3; k++) {
    args.leaf[k] = leaf[k];
    args.chunk[k] = (long long)chunk[k];
  }
  if (ipt == kFusedIptSmall)
    launch_fused<kFusedIptSmall>(v, args, w, d_out, stream);
  else
    launch_fused<kFusedIptLarge>(v, args, w, d_out, stream);
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
  // Clouds of up to 1.2M points: the whole Filter in one cooperative kernel; larger ones: the multi-kernel LSD
  // pipeline below.  g_vg_path is a test hook (pcg_debug_set_vg_path): 1 forces the multi-kernel pipeline so that the
  // parity tests cover it at every size.
  const int path = g_vg_path.load(std::memory_order_relaxed);
  if (path == 0 && v.n <= fused_capacity(kFusedIptLarge)) {
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
  run_pipeline(v, P, total_bits, d_out, d_n.p, d_flags.p, stream);
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


// ======================================================================================
// One large VoxelGrid over several GPUs: chunks are independent units in the reference
// (voxelgrid.go:102-116), so rank r filters the chunks [cid_lo, cid_hi) of the replicated cloud and the ranks'
// outputs, concatenated in rank order, are the reference's output.  MinMaxVec3, the grid and the chunk table are
// those of the WHOLE cloud; the points of the range are compacted in index order (the stable sort then keeps the
// reference's accumulation order), sorted and reduced by the kernels above.
__global__ void __launch_bounds__(256)
    chunk_hist_kernel(CloudView v, VgParams P, int shift, unsigned long long n_ids, int64_t sample_step,
                      unsigned int* __restrict__ hist, int* __restrict__ flags) {
  // a sampled histogram takes one run of 32 consecutive points out of every 32*sample_step (coalesced, and a warp
  // still sees neighbouring points)
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = (t >> 5) * 32 * sample_step + (t & 31);
  const bool live = i < v.n;
  int bad = 0;
  const KeyConsts C(P);
  unsigned long long cid = ~0ull;
  if (live) cid = voxel_key_of(P, C, load_xyz(v, i), &bad) >> shift;
  if (bad) atomicOr(flags, bad);
  // scans are spatially coherent: a warp's 32 points fall into a few chunks, so one lane per distinct chunk adds
  const unsigned peers = __match_any_sync(0xffffffffu, cid);
  if (live && cid < n_ids && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[cid], (unsigned)__popc(peers));
}

template <typename K>
__global__ void __launch_bounds__(256)
    range_key_kernel(CloudView v, VgParams P, int shift, unsigned long long cid_lo, unsigned long long cid_hi,
                     K* __restrict__ keys, float4* __restrict__ xyz4, uint32_t* __restrict__ block_counts,
                     int* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool in = false;
  if (i < v.n) {
    int bad = 0;
    const KeyConsts C(P);
    const float3 pt = load_xyz(v, i);
    const unsigned long long key = voxel_key_of(P, C, pt, &bad);
    const unsigned long long cid = key >> shift;
    in = cid >= cid_lo && cid < cid_hi;
    // an input the reference would panic on fails on every rank, whichever range the offending point falls in
    if (bad) atomicOr(flags, bad);
    keys[i] = (K)key;
    if (in) xyz4[i] = make_float4(pt.x, pt.y, pt.z, 0.f);  // the gather only reads the range's own points
  }
  const int cnt = __syncthreads_count(in);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)cnt;
}

// Ordered compaction of the range's points: block offsets from the scan of block_counts, rank inside the block by
// ballot.  Index order is kept, so equal keys stay in the reference's accumulation order through the stable sort.
template <typename K>
__global__ void __launch_bounds__(256)
    range_compact_kernel(const K* __restrict__ keys, const long long* __restrict__ block_offs, uint32_t n, int shift,
                         unsigned long long cid_lo, unsigned long long cid_hi, K* __restrict__ keys_out,
                         uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t wsum[8];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool in = false;
  K key = 0;
  if (i < n) {
    key = keys[i];
    const unsigned long long cid = (unsigned long long)key >> shift;
    in = cid >= cid_lo && cid < cid_hi;
  }
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if (lane == 0) wsum[warp] = (uint32_t)__popc(b);
  __syncthreads();
  if (!in) return;
  uint32_t base = 0;
  for (int w = 0; w < warp; w++) base += wsum[w];
  const long long o = block_offs[blockIdx.x] + base + __popc(b & ((1u << lane) - 1u));
  keys_out[o] = key;
  vals_out[o] = i;
}

static void vg_params_device(const CloudView& v, const float leaf[3], const int64_t chunk[3], VgParams* P,
                             int* total_bits, cudaStream_t stream, const float* mm6 = nullptr) {
  float vmin[3], vmax[3];
  if (mm6) {  // the bounds of the WHOLE cloud, reduced over the ranks by the caller (this rank holds a slice)
    for (int k = 0; k < 3; k++) {
      vmin[k] = mm6[k];
      vmax[k] = mm6[3 + k];
    }
  } else {
    if (v.n == 0) throw StatusError{PCG_E_NO_POINT, "no point"};
    minmax_device(v, vmin, vmax, stream);
  }
  const long long chunk_ll[3] = {(long long)chunk[0], (long long)chunk[1], (long long)chunk[2]};
  const pcg_status prc = vg_make_params(vmin, vmax, leaf, chunk_ll, P, total_bits);
  if (prc != PCG_OK) throw StatusError{prc, vg_status_message(prc)};
}


// The ids a sharded Filter is cut by: chunk ids when the filter is chunked (the loop of voxelgrid.go:102-116); for an
// un-chunked filter (one chunk) the top <= 12 bits of the voxel key - voxels are independent and emitted in ascending
// key order (voxelgrid.go:172-184), so contiguous key ranges concatenate to the reference's output just the same.
static int range_shift(const VgParams& P) { return P.n_chunks > 1 ? P.key_bits : std::max(0, P.key_bits - 12); }
static int64_t range_ids(const VgParams& P) {
  return P.n_chunks > 1 ? (int64_t)P.n_chunks : (int64_t)(((unsigned long long)(P.n_voxels - 1)) >> range_shift(P)) + 1;
}

// Points per id, for balancing the ranges over the ranks.
int64_t voxelgrid_chunk_histogram_device(const CloudView& v, const float leaf[3], const int64_t chunk[3],
                                         int64_t sample_step, int64_t* hist_out, int64_t cap, cudaStream_t stream,
                                         const float* mm6) {
  if (sample_step < 1) throw StatusError{PCG_E_INVALID_ARG, "sample_step < 1"};
  VgParams P;
  int total_bits = 0;
  vg_params_device(v, leaf, chunk, &P, &total_bits, stream, mm6);
  if (P.n_chunks > ((int64_t)1 << 26)) throw StatusError{PCG_E_TOO_LARGE, "more than 2^26 chunks"};
  const int64_t ids = range_ids(P);
  const int shift = range_shift(P);
  if (!hist_out || cap < ids) return ids;  // size query
  DevBuf<unsigned int> hist((size_t)ids, stream);
  DevBuf<int> d_flags(1, stream);
  PCG_CUDA(cudaMemsetAsync(hist.p, 0, hist.bytes(), stream));
  PCG_CUDA(cudaMemsetAsync(d_flags.p, 0, sizeof(int), stream));
  const int64_t runs = div_up(div_up(v.n, (int64_t)32), sample_step);
  if (runs > 0)
    PCG_LAUNCH(chunk_hist_kernel, (unsigned)div_up(runs * 32, (int64_t)256), 256, 0, stream, v, P, shift,
               (unsigned long long)ids, sample_step, hist.p, d_flags.p);
  std::vector<unsigned int> h((size_t)ids);
  int h_flags = 0;
  PCG_CUDA(cudaMemcpyAsync(h.data(), hist.p, hist.bytes(), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaMemcpyAsync(&h_flags, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  vg_throw_on_flags(h_flags);
  for (int64_t c = 0; c < ids; c++) hist_out[c] = (int64_t)h[(size_t)c];
  return ids;
}

template <typename K>
static void run_range_reduce(const CloudView& v, const VgParams& P, int total_bits, unsigned long long cid_lo,
                             unsigned long long cid_hi, uint8_t* d_out, long long* d_n_out, int* d_flags,
                             cudaStream_t stream) {
  const uint32_t n = (uint32_t)v.n;
  DevBuf<K> keys_all(n, stream);
  DevBuf<float4> xyz4(n, stream);
  const uint32_t blocks = (uint32_t)div_up(n, 256);
  DevBuf<uint32_t> block_counts(blocks, stream);
  DevBuf<long long> offs((size_t)blocks + 1, stream);
  const int shift = range_shift(P);
  PCG_LAUNCH((range_key_kernel<K>), blocks, 256, 0, stream, v, P, shift, cid_lo, cid_hi, keys_all.p, xyz4.p, block_counts.p,
             d_flags);
  scan_counts(block_counts.p, offs.p, blocks, stream);
  long long m = 0;
  PCG_CUDA(cudaMemcpyAsync(&m, offs.p + blocks, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  if (m == 0) return;  // *d_n_out stays 0
  const uint32_t nr = (uint32_t)m;
  DevBuf<K> keys0(nr, stream), keys1(nr, stream);
  DevBuf<uint32_t> vals0(nr, stream), vals1(nr, stream);
  PCG_LAUNCH((range_compact_kernel<K>), blocks, 256, 0, stream, keys_all.p, offs.p, n, shift, cid_lo, cid_hi,
             keys0.p, vals0.p);
  rsort::Sorter<K> sorter;
  sorter.prepare(nr, 0, total_bits, stream);
  sorter.histogram(keys0.p, stream);
  K* kk[2] = {keys0.p, keys1.p};
  uint32_t* vbuf[2] = {vals0.p, vals1.p};
  int res = 0;
  sorter.run(kk, vbuf, /*identity_vals=*/false, /*keep_keys=*/true, stream, &res);
  const int tiles = div_up(nr, kSegTile);
  DevBuf<unsigned long long> status((size_t)tiles + 1, stream);
  PCG_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), stream));
  uint32_t* counter = (uint32_t*)(status.p + tiles);
  CloudView vr = v;
  vr.n = nr;  // length of the sorted list; records and points are still addressed by their original index
  PCG_LAUNCH((voxel_reduce_kernel<K>), tiles, kSegThreads, 0, stream, vr, P, kk[res], vbuf[res], xyz4.p, d_out,
             counter, status.p, d_n_out);
}

// Point-sharded Filter, exchange plan: every rank holds a slice of the cloud; the chunk ids [cuts[r], cuts[r+1]) belong
// to rank r.  Writes the stable order of this slice's points by owner (d_perm[n]: first the points of rank 0's chunks
// in index order, then rank 1's, ...) and counts[r] = points owned by rank r - what an all-to-all of the records needs.
struct OwnerCuts {
  unsigned long long cut[9];  // cut[r] = first id of rank r, r = 1 .. world-1 (entries past world are never reached)
  int world;
};
__global__ void __launch_bounds__(256)
    owner_key_kernel(CloudView v, VgParams P, int shift, OwnerCuts oc, uint32_t* __restrict__ owner,
                     uint32_t* __restrict__ hist, int* __restrict__ flags) {
  __shared__ uint32_t s_hist[rsort::kRadix];
  rsort::hist_zero(s_hist, 1);
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (v.n + stride - 1) / stride;
  int bad = 0;
  const KeyConsts C(P);
  for (int64_t r = 0; r < rounds; r++) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < v.n;
    uint32_t o = 0;
    if (valid) {
      const unsigned long long id = voxel_key_of(P, C, load_xyz(v, i), &bad) >> shift;
      for (int q = 1; q < oc.world; q++) o += id >= oc.cut[q] ? 1u : 0u;
      owner[i] = o;
    }
    rsort::hist_add_key(s_hist, o, valid, 0, 1);
  }
  if (bad) atomicOr(flags, bad);
  __syncthreads();
  rsort::hist_flush(s_hist, hist, 1);
}

// records (whole, `stride` bytes each) in the order of perm: 4-byte words when everything is aligned, else bytes
__global__ void __launch_bounds__(256)
    gather_records_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n, uint32_t stride,
                          int words, uint8_t* __restrict__ dst) {
  if (words > 0) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)n * (uint32_t)words;
    if (t >= total) return;
    const uint32_t r = (uint32_t)(t / (uint32_t)words), w = (uint32_t)(t % (uint32_t)words);
    reinterpret_cast<uint32_t*>(dst)[t] = __ldg(reinterpret_cast<const uint32_t*>(src + (uint64_t)perm[r] * stride) + w);
  } else {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n * stride) return;
    const uint32_t r = (uint32_t)(t / stride), b = (uint32_t)(t % stride);
    dst[t] = src[(uint64_t)perm[r] * stride + b];
  }
}

void voxelgrid_owner_order_device(const CloudView& v, const float leaf[3], const int64_t chunk[3], const float* mm6,
                                  const int64_t* cuts, int world, uint32_t* d_perm, int64_t* counts, uint8_t* d_send,
                                  cudaStream_t stream) {
  if (world < 1 || world > 8) throw StatusError{PCG_E_INVALID_ARG, "1 to 8 ranks"};
  for (int r = 0; r < world; r++) counts[r] = 0;
  VgParams P;
  int total_bits = 0;
  vg_params_device(v, leaf, chunk, &P, &total_bits, stream, mm6);
  if (v.n == 0) return;
  OwnerCuts oc;
  std::memset(&oc, 0, sizeof(oc));
  oc.world = world;
  for (int r = 1; r < world; r++) {
    if (cuts[r] < cuts[r - 1]) throw StatusError{PCG_E_INVALID_ARG, "chunk cuts are not monotone"};
    oc.cut[r] = (unsigned long long)cuts[r];
  }
  const uint32_t n = (uint32_t)v.n;
  DevBuf<uint32_t> k0(n, stream), k1(n, stream), v0(n, stream);
  DevBuf<int> d_flags(1, stream);
  PCG_CUDA(cudaMemsetAsync(d_flags.p, 0, sizeof(int), stream));
  rsort::Sorter<uint32_t> sorter;
  sorter.prepare(n, 0, 8, stream);
  const int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(n, 256));
  PCG_LAUNCH(owner_key_kernel, blocks, 256, 0, stream, v, P, range_shift(P), oc, k0.p, sorter.hist(), d_flags.p);
  uint32_t* kk[2] = {k0.p, k1.p};
  uint32_t* vv[2] = {v0.p, d_perm};  // one pass: the payload ends on side 1
  int res = 0;
  sorter.run(kk, vv, /*identity_vals=*/true, /*keep_keys=*/false, stream, &res);
  if (d_send) {
    const bool aligned = (v.stride & 3) == 0 && (((uintptr_t)v.data | (uintptr_t)d_send) & 3) == 0;
    const int words = aligned ? (int)(v.stride >> 2) : 0;
    const uint64_t work = aligned ? (uint64_t)n * (uint64_t)words : (uint64_t)n * (uint64_t)v.stride;
    PCG_LAUNCH(gather_records_kernel, (unsigned)((work + 255) / 256), 256, 0, stream, v.data, d_perm, n, (uint32_t)v.stride,
               words, d_send);
  }
  uint32_t h[8];
  int h_flags = 0;
  PCG_CUDA(cudaMemcpyAsync(h, sorter.hist(), sizeof(h), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaMemcpyAsync(&h_flags, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  vg_throw_on_flags(h_flags);
  for (int r = 0; r < world; r++) counts[r] = (int64_t)h[r];
}

// Filter restricted to the chunks [cid_lo, cid_hi). Synchronises `stream`.
pcg_status voxelgrid_filter_chunks_device(const CloudView& v, const float leaf[3], const int64_t chunk[3],
                                          int64_t cid_lo, int64_t cid_hi, uint8_t* d_out, int64_t* n_out,
                                          cudaStream_t stream, const float* mm6) {
  *n_out = 0;
  VgParams P;
  int total_bits = 0;
  vg_params_device(v, leaf, chunk, &P, &total_bits, stream, mm6);
  if (v.n == 0) return PCG_OK;  // a rank that owns no point of the cloud
  if (cid_lo < 0 || cid_hi < cid_lo) throw StatusError{PCG_E_INVALID_ARG, "bad chunk range"};
  DevBuf<long long> d_n(1, stream);
  DevBuf<int> d_flags(1, stream);
  PCG_CUDA(cudaMemsetAsync(d_flags.p, 0, sizeof(int), stream));
  PCG_CUDA(cudaMemsetAsync(d_n.p, 0, sizeof(long long), stream));
  if (mm6) {
    // point-sharded Filter: the caller sent this rank exactly the points of its chunks (owner_order + all-to-all), so
    // there is nothing to select - the plain pipeline under the global bounds
    run_pipeline(v, P, total_bits, d_out, d_n.p, d_flags.p, stream);
  } else if (total_bits <= 32)
    run_range_reduce<uint32_t>(v, P, total_bits, (unsigned long long)cid_lo, (unsigned long long)cid_hi, d_out, d_n.p,
                               d_flags.p, stream);
  else
    run_range_reduce<unsigned long long>(v, P, total_bits, (unsigned long long)cid_lo, (unsigned long long)cid_hi,
                                         d_out, d_n.p, d_flags.p, stream);
  long long* ph_n = (long long*)pinned_scratch();
  int* ph_flags = (int*)(pinned_scratch() + 16);
  PCG_CUDA(cudaMemcpyAsync(ph_n, d_n.p, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaMemcpyAsync(ph_flags, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  vg_throw_on_flags(*ph_flags);
  *n_out = *ph_n;
  return PCG_OK;
}

}  // namespace pcg
