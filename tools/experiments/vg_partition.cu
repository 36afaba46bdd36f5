// vg_partition.cu — filter.VoxelGrid (pc/filter/voxelgrid/voxelgrid.go:35-187) as a sample-partition pipeline.
//
// The LSD pipeline of voxelgrid.cu moves every (key, index) pair through HBM once per 8-bit digit (five times for
// a 50M-point map) and then gathers the points in sorted order.  Here every point crosses HBM twice:
//
//   minmax      MinMaxVec3 of the cloud; the last CTA derives the grid parameters on the device (no host round trip)
//   sample      32 keys per bucket, one from a random position of every stratum of the cloud (a fixed stride
//               aliases with the 64 beams of a scan), sorted; every 32nd sample is a splitter
//   partition   one pass over the cloud: key -> bucket (binary search over the splitters in shared memory) ->
//               the point's record {x, y, z, index} and its key relative to the bucket go to the bucket's region.
//               A warp ranks its 256 consecutive points per bucket in index order and reserves one run per
//               bucket with one global atomic, so a bucket is a sequence of runs, each ascending by index
//   bucket      one CTA per bucket, everything in shared memory: the runs are put in index order (they arrive in
//               the order the warps reserved them), a stable LSD radix sort on the bits the bucket's keys differ in,
//               voxel heads, one thread per voxel adds the members in list order = the reference's accumulation
//               order (voxelgrid.go:148-158), records written at the slot the voxel has in the output (voxel
//               counts chained between CTAs by a decoupled look-back)
//
// Buckets are cut by key, so a voxel never straddles two of them, and they are emitted in key order: the output
// bytes are those of the LSD pipeline (and of the reference).  A bucket region holds kCap records; the splitters
// aim at kTarget.  If a bucket overflows (one voxel with thousands of points, an adversarial order) or the cloud
// needs more buckets than the splitter table holds, the caller falls back to the LSD pipeline.
#include "radix_sort.cuh"
#include "vg_common.cuh"

namespace pcg {

namespace vgp {

constexpr int kCap = 4096;            // records per bucket region
constexpr int kTarget = 1792;         // expected records per bucket (2.29x headroom: 7 sigma with 32 samples per bucket)
constexpr int kSamplesPerBucket = 32;
constexpr int kMaxBuckets = 28000;    // splitter table in shared memory: 4 B each
constexpr int kChunkShift = 8;        // a warp ranks 256 consecutive points: run id = index >> 8

struct State {  // device memory, zero-initialised per call
  unsigned long long acc[6];  // ~min / max packed (ordered value bits, index), reduced with atomicMax
  unsigned int ticket_minmax;
  unsigned int ticket_bucket;
  unsigned int flags;     // kFlag* bits raised while keys are computed
  unsigned int overflow;  // a bucket region or a 32-bit relative key overflowed: the result is not valid
  int status;             // pcg_status of vg_make_params
  int total_bits;
  int shift;              // splitters compare key >> shift (keys wider than 32 bits)
  int pad_;
  long long n_out;
  VgParams P;
};

// ---- minmax + parameters ------------------------------------------------------------------------------------
struct MinMaxAcc {
  unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0ull, 0ull, 0ull};
  __device__ __forceinline__ void add(float c, int k, uint32_t pos) {
    if (c != c) return;  // NaN never wins a comparison in the reference
    const unsigned long long o = (unsigned long long)ordered_bits(c) << 32;
    const unsigned long long a = o | pos, b = o | (0xffffffffu - pos);
    mn[k] = a < mn[k] ? a : mn[k];
    mx[k] = b > mx[k] ? b : mx[k];
  }
};

__global__ void __launch_bounds__(256)
    minmax_params_kernel(CloudView v, float3 leaf, longlong3 chunk, State* __restrict__ st) {
  __shared__ unsigned long long s_red[8][6];
  __shared__ bool s_last;
  __shared__ float s_mm[6];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = (uint32_t)v.n;
  MinMaxAcc a;
  if (v.packed && v.stride == 12 && v.off[0] == 0 && ((uintptr_t)v.data & 15) == 0) {
    // xyz-only records: 4 points = 3 aligned 16-byte loads
    const float4* __restrict__ q = reinterpret_cast<const float4*>(v.data);
    const uint32_t groups = n / 4;
    for (uint32_t g = blockIdx.x * blockDim.x + tid; g < groups; g += gridDim.x * blockDim.x) {
      const float4 f0 = __ldg(q + 3 * (size_t)g), f1 = __ldg(q + 3 * (size_t)g + 1), f2 = __ldg(q + 3 * (size_t)g + 2);
      const uint32_t p = 4 * g;
      a.add(f0.x, 0, p);
      a.add(f0.y, 1, p);
      a.add(f0.z, 2, p);
      a.add(f0.w, 0, p + 1);
      a.add(f1.x, 1, p + 1);
      a.add(f1.y, 2, p + 1);
      a.add(f1.z, 0, p + 2);
      a.add(f1.w, 1, p + 2);
      a.add(f2.x, 2, p + 2);
      a.add(f2.y, 0, p + 3);
      a.add(f2.z, 1, p + 3);
      a.add(f2.w, 2, p + 3);
    }
    if (blockIdx.x == 0 && tid < (n & 3u)) {
      const uint32_t p = groups * 4 + tid;
      const float3 pt = load_xyz(v, p);
      a.add(pt.x, 0, p);
      a.add(pt.y, 1, p);
      a.add(pt.z, 2, p);
    }
  } else {
    for (uint32_t p = blockIdx.x * blockDim.x + tid; p < n; p += gridDim.x * blockDim.x) {
      const float3 pt = load_xyz(v, p);
      a.add(pt.x, 0, p);
      a.add(pt.y, 1, p);
      a.add(pt.z, 2, p);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long x = shfl_xor_u64(a.mn[k], d), y = shfl_xor_u64(a.mx[k], d);
      a.mn[k] = x < a.mn[k] ? x : a.mn[k];
      a.mx[k] = y > a.mx[k] ? y : a.mx[k];
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      s_red[warp][k] = ~a.mn[k];  // min as max of the complement: zero-initialised accumulators
      s_red[warp][3 + k] = a.mx[k];
    }
  }
  __syncthreads();
  if (tid < 6) {
    unsigned long long r = s_red[0][tid];
    for (int w = 1; w < 8; w++) r = s_red[w][tid] > r ? s_red[w][tid] : r;
    if (r != 0ull) atomicMax(&st->acc[tid], r);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&st->ticket_minmax, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  // the last CTA: MinMaxVec3's result (pc/minmax.go:9-26) and Filter's grid (voxelgrid.go:45-63,137-138)
  __threadfence();
  if (tid < 6) {
    const int k = tid, c = k % 3;
    const float3 p0 = load_xyz(v, 0);
    const float first = c == 0 ? p0.x : (c == 1 ? p0.y : p0.z);
    float r = first;  // a NaN at point 0 is never replaced (minmax.go:13,17-22)
    if (first == first) {
      unsigned long long w = __ldcg(&st->acc[k]);
      if (k < 3) w = ~w;
      const uint32_t idx = k < 3 ? (uint32_t)w : 0xffffffffu - (uint32_t)w;
      const float3 p = load_xyz(v, idx);
      r = c == 0 ? p.x : (c == 1 ? p.y : p.z);
    }
    s_mm[k] = r;
  }
  __syncthreads();
  if (tid == 0) {
    float mm[6];
    for (int k = 0; k < 6; k++) mm[k] = s_mm[k];
    const float lf[3] = {leaf.x, leaf.y, leaf.z};
    const long long ch[3] = {chunk.x, chunk.y, chunk.z};
    int tb = 0;
    VgParams P;
    const pcg_status rc = vg_make_params(mm, mm + 3, lf, ch, &P, &tb);
    st->status = (int)rc;
    if (rc == PCG_OK) {
      st->P = P;
      st->total_bits = tb;
      st->shift = tb > 32 ? tb - 32 : 0;
    }
  }
}

// ---- samples ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(256)
    sample_kernel(CloudView v, const State* __restrict__ st, uint32_t n_samples, uint32_t* __restrict__ skeys,
                  uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_hist[4 * rsort::kRadix];
  if (st->status != PCG_OK) return;
  rsort::hist_zero(s_hist, 4);
  __syncthreads();
  const VgParams& P = st->P;
  const KeyConsts C(P);
  const int shift = st->shift;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n_samples;
  uint32_t sk = 0;
  if (valid) {
    const uint64_t lo = (uint64_t)i * (uint64_t)v.n / n_samples, hi = (uint64_t)(i + 1) * (uint64_t)v.n / n_samples;
    const uint64_t pos = lo + (hi > lo ? mix32(i) % (uint32_t)(hi - lo) : 0u);
    int bad = 0;
    sk = (uint32_t)(voxel_key_of(P, C, load_xyz(v, (int64_t)pos), &bad) >> shift);
    skeys[i] = sk;
  }
  rsort::hist_add_key(s_hist, sk, valid, 0, 4);
  __syncthreads();
  rsort::hist_flush(s_hist, hist, 4);
}

__global__ void __launch_bounds__(256)
    splitters_kernel(const uint32_t* __restrict__ sorted, uint32_t n_samples, uint32_t n_buckets, uint32_t padded,
                     uint32_t* __restrict__ spl) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= padded) return;
  spl[j] = j + 1 < n_buckets ? sorted[(uint64_t)(j + 1) * n_samples / n_buckets] : 0xffffffffu;
}

// ---- partition -------------------------------------------------------------------------------------------------
constexpr int kPartThreads = 256;
constexpr int kPartWarps = kPartThreads / 32;
constexpr int kPartIpt = 8;       // points per thread per chunk: a warp owns (1 << kChunkShift) consecutive points
constexpr int kHashSlots = 512;   // per warp: at most 256 distinct buckets per chunk
static_assert(32 * kPartIpt == (1 << kChunkShift), "a warp ranks one chunk");

struct PartSmem {
  uint32_t tab[kPartWarps][kHashSlots];          // bucket -> count, then -> first slot of the warp's run
  uint32_t rel[kPartIpt][kPartThreads];          // parked per point: key relative to its bucket
  uint32_t where[kPartIpt][kPartThreads];        // bucket << 17 | hash slot << 8 | rank in the warp's run
};
static_assert(kMaxBuckets < (1 << 15), "bucket id is packed into 15 bits");

__global__ void __launch_bounds__(kPartThreads, 4)
    partition_kernel(CloudView v, State* __restrict__ st, const uint32_t* __restrict__ spl, uint32_t n_buckets,
                     uint32_t* __restrict__ cursor, float4* __restrict__ rec, uint32_t* __restrict__ key32) {
  __shared__ PartSmem sm;
  if (st->status != PCG_OK) return;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_spl = n_buckets - 1;
  for (uint32_t i = tid; i < kPartWarps * kHashSlots; i += kPartThreads) (&sm.tab[0][0])[i] = 0;
  __syncthreads();
  uint32_t top = 1;  // largest power of two <= n_spl (0 splitters: the search below does nothing)
  while (top * 2 <= n_spl) top *= 2;
  if (n_spl == 0) top = 0;
  const KeyConsts C(st->P);
  const int shift = st->shift;
  uint32_t* tab = sm.tab[warp];
  const uint32_t n = (uint32_t)v.n;
  const uint32_t chunks = (n + (1u << kChunkShift) - 1) >> kChunkShift;
  const uint32_t lt = (1u << lane) - 1u;
  int bad = 0;
  bool over = false;
  for (uint32_t chunk = blockIdx.x * kPartWarps + warp; chunk < chunks; chunk += gridDim.x * kPartWarps) {
    const uint32_t base = chunk << kChunkShift;
    // the key arithmetic is a few hundred instructions: a rolled loop (its code is fetched once), the next point's
    // coordinates loaded one iteration ahead, results parked in shared memory
    float3 nxt = base + lane < n ? load_xyz(v, base + lane) : make_float3(0.f, 0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < kPartIpt; i++) {
      const uint32_t pos = base + i * 32 + lane;
      const bool valid = pos < n;
      const float3 pt = nxt;
      if (i + 1 < kPartIpt && pos + 32 < n) nxt = load_xyz(v, pos + 32);
      uint32_t b = 0xffffffffu, rel = 0;
      if (valid) {
        unsigned long long key;
        if (!voxel_key_fast(C, pt, &key)) key = voxel_key_general(st->P, pt, &bad);
        const uint32_t sk = (uint32_t)(key >> shift);
        // bucket = number of splitters <= key >> shift (the table is read-only and small: it lives in L1)
        // (branch-free: the table is padded with 0xffffffff up to 2 * top entries)
        uint32_t lo = 0;
        for (uint32_t step = top; step; step >>= 1) lo += __ldg(&spl[lo + step - 1]) <= sk ? step : 0u;
        b = min(lo, n_spl);
        const unsigned long long kmin = b ? (unsigned long long)__ldg(&spl[b - 1]) << shift : 0ull;
        const unsigned long long r = key - kmin;
        over = over || (r >> 32) != 0 || (uint32_t)r == 0xffffffffu;  // (0xffffffff marks an empty hash slot in bucket_kernel)
        rel = (uint32_t)r;
      }
      // rank among the warp's points of the same bucket, in index order (item-major, lane-minor)
      const uint32_t peers = __match_any_sync(0xffffffffu, b);
      const int leader = __ffs(peers) - 1;
      uint32_t h = 0, prev = 0;
      if (valid && (int)lane == leader) {
        const uint32_t tag = (b + 1) << 16;
        h = (b * 0x9e3779b1u) >> 23;  // 9 bits
        for (;;) {
          const uint32_t old = atomicCAS(&tab[h], 0u, tag);
          if (old == 0u || (old & 0xffff0000u) == tag) break;
          h = (h + 1) & (kHashSlots - 1);
        }
        prev = atomicAdd(&tab[h], (uint32_t)__popc(peers)) & 0xffffu;
      }
      h = __shfl_sync(0xffffffffu, h, leader);
      prev = __shfl_sync(0xffffffffu, prev, leader);
      sm.rel[i][tid] = rel;
      sm.where[i][tid] = (b << 17) | (h << 8) | (prev + __popc(peers & lt));
    }
    __syncwarp();
    // one reservation per bucket the chunk touches; the atomics of a lane are independent of each other
    {
      uint32_t e[kHashSlots / 32], at[kHashSlots / 32];
#pragma unroll
      for (int j = 0; j < kHashSlots / 32; j++) e[j] = tab[j * 32 + lane];
#pragma unroll
      for (int j = 0; j < kHashSlots / 32; j++) {
        at[j] = 0;
        if (e[j]) at[j] = atomicAdd(&cursor[(e[j] >> 16) - 1], e[j] & 0xffffu);
      }
#pragma unroll
      for (int j = 0; j < kHashSlots / 32; j++) {
        if (e[j]) {
          over = over || at[j] + (e[j] & 0xffffu) > (uint32_t)kCap;
          tab[j * 32 + lane] = at[j];
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kPartIpt; i++) {
      const uint32_t pos = base + i * 32 + lane;
      if (pos < n) {
        const uint32_t w = sm.where[i][tid];
        const uint32_t at = tab[(w >> 8) & (kHashSlots - 1)] + (w & 0xffu);
        if (at < (uint32_t)kCap) {
          const float3 pt = load_xyz(v, pos);  // L1
          const size_t dst = (size_t)(w >> 17) * kCap + at;
          rec[dst] = make_float4(pt.x, pt.y, pt.z, __uint_as_float(pos));
          key32[dst] = sm.rel[i][tid];
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kHashSlots / 32; j++) tab[j * 32 + lane] = 0;
    __syncwarp();
  }
  if (bad) atomicOr(&st->flags, (unsigned int)bad);
  if (over) atomicOr(&st->overflow, 1u);
}

// ---- bucket: group by voxel, order, reduce -----------------------------------------------------------------------
// One CTA per bucket, everything in shared memory, no radix passes:
//   1  every record's key goes into a hash table (open addressing): slot -> member count; the first record of a key
//      appends it to the list of distinct keys.  The number of distinct keys = voxels of the bucket is published
//      for the look-back at once, so no later bucket ever waits for this one's work
//   2  the distinct keys (a few hundred) are sorted (bitonic): voxel u = u-th smallest key; exclusive scan of the
//      member counts = first position of every voxel
//   3  records are scattered into their voxel's range (atomic fill: arbitrary order inside the voxel), then every
//      record counts the members of its voxel with a smaller point index: that is its place in the reference's
//      accumulation order (voxelgrid.go:148-158 adds the points in iteration order)
//   4  points gathered in final order; one thread per voxel adds its members sequentially; record + centroid written
//      at the voxel's slot of the output (bucket offsets by decoupled look-back)
// A voxel with more than kMaxMembers points makes step 3 quadratic: the bucket raises the overflow flag and the
// caller runs the LSD pipeline.
constexpr int kBktThreads = 256;
constexpr int kBktWarps = kBktThreads / 32;
constexpr int kBktIpt = kCap / kBktThreads;  // 16 rows of 256 records
constexpr int kHashMax = 2 * kCap;
constexpr uint32_t kMaxMembers = 255;

struct BucketSmem {
  union {
    struct {
      uint32_t hkey[kHashMax];      // slot -> key (0xffffffff = empty), later slot -> voxel
      uint32_t hcnt[kHashMax / 2];  // slot -> member count, two 16-bit counters per word; later the fill cursors
    } h;
    float xyz[3][kCap];             // points in final order (steps 4)
  } u;
  uint32_t dkey[kCap];     // distinct keys, sorted in step 2
  uint16_t dslot[kCap];    // their hash slots; from step 3 on: final position -> record
  uint16_t vstart[kCap + 2];  // first position of voxel u (vstart[U] = m)
  uint32_t g_idx[kCap];    // grouped position -> point index
  union {
    struct {
      uint16_t g_l[kCap];  // grouped position -> record
      uint16_t g_u[kCap];  // grouped position -> voxel
    } g;
    struct {               // scratch of the distinct-key sort (step 2)
      uint16_t hist[kBktWarps][256];
      uint32_t digit_start[256];
    } r;
  } w;
  uint32_t scan[kBktWarps];
  uint32_t bucket, distinct, heavy, maxkey;
  unsigned long long look[kBktWarps][2];
};
static_assert(sizeof(BucketSmem) <= 113 * 1024, "two buckets per SM");

__device__ __forceinline__ uint32_t bkt_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
  {
    const uint32_t c = lane < (uint32_t)kBktWarps ? s_warp[lane] : 0u;
    base = c & (lane < warp ? 0xffffffffu : 0u);
    tot = c;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      base += __shfl_xor_sync(0xffffffffu, base, d);
      tot += __shfl_xor_sync(0xffffffffu, tot, d);
    }
  }
  __syncthreads();
  if (total) *total = tot;
  return base + incl - v;
}

__device__ __forceinline__ uint32_t cnt_get(const uint32_t* hcnt, uint32_t s) { return (hcnt[s >> 1] >> (16 * (s & 1))) & 0xffffu; }
__device__ __forceinline__ uint32_t cnt_inc(uint32_t* hcnt, uint32_t s) {  // returns the counter's previous value
  return (atomicAdd(&hcnt[s >> 1], 1u << (16 * (s & 1))) >> (16 * (s & 1))) & 0xffffu;
}

// Stable LSD radix sort (8-bit digits) of the `count` (key, payload) pairs in sm.dkey / sm.dslot, in place, on key
// bits [0, nbits).  All threads call; ends with a barrier.
__device__ __forceinline__ void bkt_sort_distinct(BucketSmem& sm, uint32_t count, int nbits) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t kPerWarp = 32u * kBktIpt;
  const uint32_t warp_base = warp * kPerWarp;
  if (nbits <= 0 || count <= 1) return;
  uint32_t keys[kBktIpt], pp[kBktIpt];
#pragma unroll
  for (int i = 0; i < kBktIpt; i++) {
    const uint32_t p = warp_base + i * 32 + lane;
    keys[i] = p < count ? sm.dkey[p] : 0u;
    pp[i] = p < count ? sm.dslot[p] : 0u;
  }
  const uint32_t used_warps = (count + kPerWarp - 1) / kPerWarp;  // warps that hold any pair
  for (int shift = 0; shift < nbits; shift += 8) {
    for (uint32_t i = tid; i < used_warps * 128u; i += kBktThreads) reinterpret_cast<uint32_t*>(&sm.w.r.hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t offs[kBktIpt];
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      offs[i] = 0;
      if (warp_base + i * 32 < count) {  // warp-uniform
        const bool valid = warp_base + i * 32 + lane < count;
        const uint32_t d = valid ? ((keys[i] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && (int)lane == leader) {
          pre = sm.w.r.hist[warp][d];
          sm.w.r.hist[warp][d] = (uint16_t)(pre + __popc(peers));
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        offs[i] = pre + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
      }
    }
    __syncthreads();
    uint32_t cnt = 0;
    {  // thread = digit: exclusive offsets of the warps' counts
      for (uint32_t w = 0; w < used_warps; w++) {
        const uint32_t c = sm.w.r.hist[w][tid];
        sm.w.r.hist[w][tid] = (uint16_t)cnt;
        cnt += c;
      }
    }
    const uint32_t dstart = bkt_excl_scan(cnt, sm.scan, nullptr);
    sm.w.r.digit_start[tid] = dstart;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      if (warp_base + i * 32 < count && warp_base + i * 32 + lane < count) {
        const uint32_t d = (keys[i] >> shift) & 255u;
        const uint32_t at = sm.w.r.digit_start[d] + sm.w.r.hist[warp][d] + offs[i];
        sm.dkey[at] = keys[i];
        sm.dslot[at] = (uint16_t)pp[i];
      }
    }
    __syncthreads();
    if (shift + 8 < nbits) {
#pragma unroll
      for (int i = 0; i < kBktIpt; i++) {
        const uint32_t p = warp_base + i * 32 + lane;
        if (warp_base + i * 32 < count && p < count) {
          keys[i] = sm.dkey[p];
          pp[i] = sm.dslot[p];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kBktThreads, 2)
    bucket_kernel(CloudView v, State* __restrict__ st, const uint32_t* __restrict__ spl, uint32_t n_buckets,
                  const uint32_t* __restrict__ cursor, const float4* __restrict__ rec,
                  const uint32_t* __restrict__ key32, unsigned long long* __restrict__ status,
                  uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char dyn[];
  BucketSmem& sm = *reinterpret_cast<BucketSmem*>(dyn);
  if (st->status != PCG_OK || st->overflow) return;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    sm.bucket = atomicAdd(&st->ticket_bucket, 1u);  // dynamic id: every predecessor is running or done
    sm.distinct = 0;
    sm.heavy = 0;
    sm.maxkey = 0;
  }
  __syncthreads();
  const uint32_t bucket = sm.bucket;
  const uint32_t m = min(__ldcg(&cursor[bucket]), (uint32_t)kCap);
  const float4* __restrict__ brec = rec + (size_t)bucket * kCap;
  const uint32_t* __restrict__ bkey = key32 + (size_t)bucket * kCap;
  volatile unsigned long long* stv = status;
  constexpr uint32_t kEmpty = 0xffffffffu;  // never a key: partition_kernel raises the overflow flag for it

  uint32_t U = 0;
  if (m > 0) {
    // ---- 1. hash the keys
    uint32_t hslots = 256;  // power of two >= 2 m
    while (hslots < 2 * m) hslots *= 2;
    for (uint32_t i = tid; i < hslots; i += kBktThreads) sm.u.h.hkey[i] = kEmpty;
    for (uint32_t i = tid; i < hslots / 2; i += kBktThreads) sm.u.h.hcnt[i] = 0;
    __syncthreads();
    const int hshift = 32 - (31 - __clz(hslots));
    uint32_t slot2[kBktIpt / 2];  // the records' hash slots, two per register
    uint32_t idx[kBktIpt];        // their point indices
    uint32_t mx = 0;
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      if ((i & 1) == 0) slot2[i / 2] = 0;
      idx[i] = 0;
      const uint32_t l = i * kBktThreads + tid;
      if (i * kBktThreads < m && l < m) {
        const uint32_t k = __ldcs(bkey + l);
        idx[i] = __float_as_uint(__ldg(&brec[l].w));
        mx = max(mx, k);
        uint32_t h = (k * 0x9e3779b1u) >> hshift;
        for (;;) {
          const uint32_t old = atomicCAS(&sm.u.h.hkey[h], kEmpty, k);
          if (old == kEmpty) {
            const uint32_t at = atomicAdd(&sm.distinct, 1u);
            sm.dkey[at] = k;
            sm.dslot[at] = (uint16_t)h;
            break;
          }
          if (old == k) break;
          h = (h + 1) & (hslots - 1);
        }
        cnt_inc(sm.u.h.hcnt, h);
        slot2[i / 2] |= h << (16 * (i & 1));
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane == 0 && mx) atomicMax(&sm.maxkey, mx);
    __syncthreads();
    U = sm.distinct;
    if (tid == 0) stv[bucket] = ((bucket == 0 ? 2ull : 1ull) << 62) | (unsigned long long)U;
    // ---- 2. sort the distinct keys (with their slots)
    bkt_sort_distinct(sm, U, sm.maxkey ? 32 - __clz(sm.maxkey) : 0);
    // ---- first position of every voxel; slot -> voxel
    {
      const uint32_t u0 = tid * kBktIpt;
      uint32_t len[kBktIpt], sum = 0, big = 0;
#pragma unroll
      for (int j = 0; j < kBktIpt; j++) {
        len[j] = u0 + j < U ? cnt_get(sm.u.h.hcnt, sm.dslot[u0 + j]) : 0u;
        big = max(big, len[j]);
        sum += len[j];
      }
      if (big > kMaxMembers) sm.heavy = 1;
      uint32_t e = bkt_excl_scan(sum, sm.scan, nullptr);  // (its barriers also order the count reads before the resets)
#pragma unroll
      for (int j = 0; j < kBktIpt; j++) {
        if (u0 + j < U) {
          const uint32_t s = sm.dslot[u0 + j];
          sm.vstart[u0 + j] = (uint16_t)e;
          sm.u.h.hkey[s] = u0 + j;
        }
        e += len[j];
      }
      if (tid == 0) sm.vstart[U] = (uint16_t)m;
      for (uint32_t i = tid; i < hslots / 2; i += kBktThreads) sm.u.h.hcnt[i] = 0;
    }
    __syncthreads();
    if (sm.heavy) {  // uniform
      if (tid == 0) atomicOr(&st->overflow, 2u);
    }
    // ---- 3. records into their voxel's range, then into the reference's accumulation order
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      const uint32_t l = i * kBktThreads + tid;
      if (i * kBktThreads < m && l < m) {
        const uint32_t s = (slot2[i / 2] >> (16 * (i & 1))) & 0xffffu;
        const uint32_t u = sm.u.h.hkey[s];
        const uint32_t p = (uint32_t)sm.vstart[u] + cnt_inc(sm.u.h.hcnt, s);
        sm.w.g.g_l[p] = (uint16_t)l;
        sm.w.g.g_u[p] = (uint16_t)u;
        sm.g_idx[p] = idx[i];
      }
    }
    __syncthreads();
    if (!sm.heavy) {
#pragma unroll 1
      for (int i = 0; i < kBktIpt; i++) {
        const uint32_t p = i * kBktThreads + tid;
        if (p < m) {
          const uint32_t l = sm.w.g.g_l[p], u = sm.w.g.g_u[p];
          const uint32_t a = sm.vstart[u], e = sm.vstart[u + 1];
          const uint32_t mine = sm.g_idx[p];
          uint32_t r = 0;
          for (uint32_t j = a; j < e; j++) r += sm.g_idx[j] < mine ? 1u : 0u;
          sm.dslot[a + r] = (uint16_t)l;
        }
      }
    }
    __syncthreads();
    // ---- 4. points in final order (the hash table is dead)
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      const uint32_t p = i * kBktThreads + tid;
      if (i * kBktThreads < m && p < m) {
        const float4 r4 = __ldg(&brec[sm.dslot[p]]);
        sm.u.xyz[0][p] = r4.x;
        sm.u.xyz[1][p] = r4.y;
        sm.u.xyz[2][p] = r4.z;
      }
    }
  } else if (tid == 0) {
    stv[bucket] = ((bucket == 0 ? 2ull : 1ull) << 62);
  }
  // ---- output slot of the bucket's first voxel: look-back over the preceding buckets, one status word per thread
  // (every running bucket has published its count already; only finished ones hold inclusive sums)
  unsigned long long prefix = 0;
  if (bucket > 0) {
    int64_t hi = (int64_t)bucket - 1;  // window [hi - 255, hi], thread t reads hi - t
    for (;;) {
      const int64_t idx = hi - (int64_t)tid;
      unsigned long long w = 2ull << 62;  // before the first bucket: an inclusive 0
      if (idx >= 0) {
        do {
          w = stv[idx];
        } while ((w >> 62) == 0);
      }
      const uint32_t incl = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
      const int stop = incl ? __ffs(incl) - 1 : 31;  // lanes are ordered nearest first
      unsigned long long val = (int)lane <= stop ? (w & ((1ull << 62) - 1)) : 0ull;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) val += shfl_xor_u64(val, d);
      if (lane == 0) {
        sm.look[warp][0] = val;
        sm.look[warp][1] = incl ? 1ull : 0ull;
      }
      __syncthreads();
      bool found = false;
      for (int w2 = 0; w2 < kBktWarps && !found; w2++) {  // warps are ordered nearest first
        prefix += sm.look[w2][0];
        found = sm.look[w2][1] != 0ull;
      }
      __syncthreads();
      if (found) break;
      hi -= kBktThreads;
    }
    if (tid == 0) stv[bucket] = (2ull << 62) | (prefix + U);
  }
  if (tid == 0 && bucket == n_buckets - 1) st->n_out = (long long)(prefix + U);
  __syncthreads();
  if (m == 0 || sm.heavy) return;
  // ---- one voxel per thread per round: members added in list order (voxelgrid.go:148-158), record of the first
  // member with the centroid (voxelgrid.go:173-184)
  const VgParams& P = st->P;
  const int out_aligned = v.aligned && ((((uintptr_t)out) & 3) == 0);
  const bool xyz_only = out_aligned && v.packed && v.stride == 12;
  const unsigned long long kmin = bucket ? (unsigned long long)__ldg(&spl[bucket - 1]) << st->shift : 0ull;
  const int key_bits = P.key_bits;
  long long vc_cid = -1;
  float vc[3] = {0.f, 0.f, 0.f};
  for (uint32_t r = tid; r < U; r += kBktThreads) {
    const uint32_t l = sm.vstart[r], end = sm.vstart[r + 1];
    const unsigned long long key = kmin + sm.dkey[r];
    const long long cid = (long long)(key >> key_bits);
    if (cid != vc_cid) {
      chunk_min(P, cid, vc);
      vc_cid = cid;
    }
    const float fx = sm.u.xyz[0][l], fy = sm.u.xyz[1][l], fz = sm.u.xyz[2][l];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (uint32_t ll = l; ll < end; ll++) {  // p = pt - vcMin, sum += p
      sx = __fadd_rn(sx, __fsub_rn(sm.u.xyz[0][ll], vc[0]));
      sy = __fadd_rn(sy, __fsub_rn(sm.u.xyz[1][ll], vc[1]));
      sz = __fadd_rn(sz, __fsub_rn(sm.u.xyz[2][ll], vc[2]));
    }
    const uint32_t num = end - l;
    float ox = fx, oy = fy, oz = fz;  // num == 1: the original bytes (voxelgrid.go:176-178)
    if (num > 1) {
      const float inv = __fdiv_rn(1.0f, (float)num);  // 1.0 / float32(n)   voxelgrid.go:179
      ox = __fadd_rn(__fmul_rn(sx, inv), vc[0]);
      oy = __fadd_rn(__fmul_rn(sy, inv), vc[1]);
      oz = __fadd_rn(__fmul_rn(sz, inv), vc[2]);
    }
    uint8_t* dst = out + (prefix + r) * (unsigned long long)v.stride;
    if (xyz_only) {
      float* d3 = reinterpret_cast<float*>(dst);
      d3[0] = ox;
      d3[1] = oy;
      d3[2] = oz;
    } else {
      const uint32_t first = __float_as_uint(__ldg(&brec[sm.dslot[l]].w));
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
  }
}

}  // namespace vgp

// Whether the partition pipeline takes a cloud of n points (else: the LSD pipeline of voxelgrid.cu).
bool vgp_eligible(int64_t n) { return n > 0 && (n + vgp::kTarget - 1) / vgp::kTarget <= vgp::kMaxBuckets; }

// Filter by the partition pipeline.  `v` and d_out are device pointers.  Synchronises `stream`.
// Returns false (and leaves *n_out alone) when a bucket overflowed: the caller runs the LSD pipeline instead.
bool voxelgrid_filter_partition(const CloudView& v, const float leaf[3], const int64_t chunk[3], uint8_t* d_out,
                                int64_t* n_out, cudaStream_t stream) {
  using namespace vgp;
  const uint32_t n = (uint32_t)v.n;
  const uint32_t n_buckets = n <= (uint32_t)kCap ? 1u : (uint32_t)((n + kTarget - 1) / kTarget);
  const uint32_t n_samples = n_buckets > 1 ? n_buckets * kSamplesPerBucket : 0u;

  // one allocation for the small state: [State | cursor[n_buckets] | look-back status[n_buckets]]
  const size_t state_bytes = (sizeof(State) + 255) & ~(size_t)255;
  const size_t cursor_bytes = ((size_t)n_buckets * 4 + 255) & ~(size_t)255;
  const size_t status_bytes = (size_t)n_buckets * 8;
  DevBuf<uint8_t> small(state_bytes + cursor_bytes + status_bytes, stream);
  PCG_CUDA(cudaMemsetAsync(small.p, 0, small.bytes(), stream));
  State* st = reinterpret_cast<State*>(small.p);
  uint32_t* cursor = reinterpret_cast<uint32_t*>(small.p + state_bytes);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(small.p + state_bytes + cursor_bytes);
  uint32_t spl_padded = 2;  // 2 * (largest power of two <= n_buckets - 1), at least 2
  while (spl_padded < 2 * n_buckets) spl_padded *= 2;
  DevBuf<uint32_t> spl(spl_padded, stream);
  DevBuf<float4> rec((size_t)n_buckets * kCap, stream);
  DevBuf<uint32_t> key32((size_t)n_buckets * kCap, stream);

  const int mm_blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 8, div_up(n, 256 * 4));
  PCG_LAUNCH(minmax_params_kernel, std::max(mm_blocks, 1), 256, 0, stream, v, make_float3(leaf[0], leaf[1], leaf[2]),
             make_longlong3((long long)chunk[0], (long long)chunk[1], (long long)chunk[2]), st);
  if (n_samples) {
    DevBuf<uint32_t> sk0(n_samples, stream), sk1(n_samples, stream), sv0(n_samples, stream), sv1(n_samples, stream);
    rsort::Sorter<uint32_t> sorter;
    sorter.prepare(n_samples, 0, 32, stream);
    PCG_LAUNCH(sample_kernel, div_up(n_samples, 256), 256, 0, stream, v, st, n_samples, sk0.p, sorter.hist());
    uint32_t* kk[2] = {sk0.p, sk1.p};
    uint32_t* vv[2] = {sv0.p, sv1.p};
    int res = 0;
    sorter.run(kk, vv, /*identity_vals=*/true, /*keep_keys=*/true, stream, &res);
    PCG_LAUNCH(splitters_kernel, div_up(spl_padded, 256), 256, 0, stream, kk[res], n_samples, n_buckets, spl_padded,
               spl.p);
  }
  {
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    PCG_CUDA(cudaGetDevice(&dev));
    if (!(configured.load(std::memory_order_relaxed) & (1ull << dev))) {
      PCG_CUDA(cudaFuncSetAttribute(bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketSmem)));
      configured.fetch_or(1ull << dev, std::memory_order_relaxed);
    }
  }
  const uint32_t chunks = (n + (1u << kChunkShift) - 1) >> kChunkShift;
  const int part_blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(chunks, kPartWarps));
  PCG_LAUNCH(partition_kernel, part_blocks, kPartThreads, 0, stream, v, st, spl.p, n_buckets, cursor, rec.p,
             key32.p);
  PCG_LAUNCH(bucket_kernel, n_buckets, kBktThreads, sizeof(BucketSmem), stream, v, st, spl.p, n_buckets, cursor,
             rec.p, key32.p, status, d_out);
  State* h = reinterpret_cast<State*>(vg_pinned_state());
  PCG_CUDA(cudaMemcpyAsync(h, st, offsetof(State, P), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  const pcg_status rc = (pcg_status)h->status;
  if (rc != PCG_OK) throw StatusError{rc, vg_status_message(rc)};
  vg_throw_on_flags((int)h->flags);
  if (h->overflow) return false;
  *n_out = h->n_out;
  return true;
}

}  // namespace pcg
