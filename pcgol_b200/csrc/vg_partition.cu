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
    splitters_kernel(const uint32_t* __restrict__ sorted, uint32_t n_samples, uint32_t n_buckets,
                     uint32_t* __restrict__ spl) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j + 1 < n_buckets) spl[j] = sorted[(uint64_t)(j + 1) * n_samples / n_buckets];
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
        uint32_t lo = 0;
        for (uint32_t step = top; step; step >>= 1) {
          const uint32_t t = lo + step;
          if (t <= n_spl && __ldg(&spl[t - 1]) <= sk) lo = t;
        }
        b = lo;
        const unsigned long long kmin = b ? (unsigned long long)__ldg(&spl[b - 1]) << shift : 0ull;
        const unsigned long long r = key - kmin;
        over = over || (r >> 32) != 0;
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

// ---- bucket: order, sort, reduce -------------------------------------------------------------------------------
constexpr int kBktThreads = 512;
constexpr int kBktWarps = kBktThreads / 32;
constexpr int kBktIpt = kCap / kBktThreads;  // 8

struct BucketSmem {
  uint32_t key[kCap];               // staging: keys
  uint16_t pos[kCap];               // staging: payload (position of the record in the bucket region / run number)
  uint16_t hist[kBktWarps][256];    // digit counts per warp; later the positions of the voxel heads
  union {
    uint32_t hset[2 * kCap];   // distinct keys of the bucket (open addressing)
    struct {
      uint32_t chunk[kCap];    // run r: index >> kChunkShift of its records
      uint16_t start[kCap];    // first record
      uint16_t rank[kCap];     // rank of the run by chunk
      uint16_t nstart[kCap];   // length by rank, then (scanned) first position in index order
      uint16_t rec_run[kCap];  // run of record l
    } run;
    float xyz[3][kCap];        // points in sorted order
  } u;
  uint32_t scan[kBktWarps];
  uint32_t digit_start[256];
  uint32_t bucket, maxkey, minchunk, maxchunk, distinct;
  unsigned long long look[kBktWarps][2];
  unsigned long long prefix;
};

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

// Stable LSD radix sort (8-bit digits) of the `count` (key, payload) pairs in sm.key / sm.pos, in place, on key bits
// [0, nbits).  Position p = warp * 256 + i * 32 + lane is the order the ranks follow.  All threads call; ends with
// a barrier.
__device__ __forceinline__ void bkt_radix_sort(BucketSmem& sm, uint32_t count, int nbits) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t warp_base = warp * (32u * kBktIpt);
  if (nbits <= 0 || count <= 1) return;
  uint32_t keys[kBktIpt], pp[kBktIpt];
#pragma unroll
  for (int i = 0; i < kBktIpt; i++) {
    const uint32_t p = warp_base + i * 32 + lane;
    keys[i] = p < count ? sm.key[p] : 0u;
    pp[i] = p < count ? sm.pos[p] : 0u;
  }
  const uint32_t used_warps = (count + 32u * kBktIpt - 1) / (32u * kBktIpt);  // warps that hold any pair
  for (int shift = 0; shift < nbits; shift += 8) {
    for (uint32_t i = tid; i < used_warps * 128u; i += kBktThreads) reinterpret_cast<uint32_t*>(&sm.hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t offs[kBktIpt];
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      if (warp_base + i * 32 < count) {  // warp-uniform
        const bool valid = warp_base + i * 32 + lane < count;
        const uint32_t d = valid ? ((keys[i] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && (int)lane == leader) {
          pre = sm.hist[warp][d];
          sm.hist[warp][d] = (uint16_t)(pre + __popc(peers));
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        offs[i] = pre + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
      }
    }
    __syncthreads();
    uint32_t cnt = 0;
    if (tid < 256) {
      for (uint32_t w = 0; w < used_warps; w++) {
        const uint32_t c = sm.hist[w][tid];
        sm.hist[w][tid] = (uint16_t)cnt;
        cnt += c;
      }
    }
    const uint32_t dstart = bkt_excl_scan(cnt, sm.scan, nullptr);
    if (tid < 256) sm.digit_start[tid] = dstart;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      if (warp_base + i * 32 + lane < count) {
        const uint32_t d = (keys[i] >> shift) & 255u;
        const uint32_t at = sm.digit_start[d] + sm.hist[warp][d] + offs[i];
        sm.key[at] = keys[i];
        sm.pos[at] = (uint16_t)pp[i];
      }
    }
    __syncthreads();
    if (shift + 8 < nbits) {
#pragma unroll
      for (int i = 0; i < kBktIpt; i++) {
        const uint32_t p = warp_base + i * 32 + lane;
        if (p < count) {
          keys[i] = sm.key[p];
          pp[i] = sm.pos[p];
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
    sm.maxkey = 0;
    sm.minchunk = 0xffffffffu;
    sm.maxchunk = 0;
    sm.distinct = 0;
  }
  __syncthreads();
  const uint32_t bucket = sm.bucket;
  const uint32_t m = min(__ldcg(&cursor[bucket]), (uint32_t)kCap);
  const float4* __restrict__ brec = rec + (size_t)bucket * kCap;
  const uint32_t* __restrict__ bkey = key32 + (size_t)bucket * kCap;
  volatile unsigned long long* stv = status;
  const uint32_t l0 = tid * kBktIpt;
  constexpr uint32_t kEmpty = 0xffffffffu;

  uint32_t total_heads = 0;
  uint32_t keys[kBktIpt];
  if (m > 0) {
    // ---- 1. load (striped); chunk ids to shared memory.  The number of distinct keys (= voxels of the bucket) is
    // counted with a hash set and published at once: no later bucket ever waits for this one's sort.
    uint32_t hslots = 256;  // power of two >= 2 m
    while (hslots < 2 * m) hslots *= 2;
    for (uint32_t i = tid; i < hslots; i += kBktThreads) sm.u.hset[i] = kEmpty;
    uint32_t mx = 0, cmin = 0xffffffffu, cmax = 0;
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      const uint32_t l = i * kBktThreads + tid;
      keys[i] = 0;
      if (l < m) {
        keys[i] = __ldcs(bkey + l);
        const uint32_t c = __float_as_uint(__ldg(&brec[l].w)) >> kChunkShift;
        sm.key[l] = c;
        mx = max(mx, keys[i]);
        cmin = min(cmin, c);
        cmax = max(cmax, c);
      }
    }
    __syncthreads();
    uint32_t fresh = 0;
    const int hshift = 32 - (31 - __clz(hslots));
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      const uint32_t l = i * kBktThreads + tid;
      if (l < m) {
        const uint32_t k = keys[i] == kEmpty ? kEmpty - 1 : keys[i];  // the marker itself shares a slot with its
        uint32_t h = (k * 0x9e3779b1u) >> hshift;                      // neighbour (corrected below)
        for (;;) {
          const uint32_t old = atomicCAS(&sm.u.hset[h], kEmpty, k);
          if (old == kEmpty) {
            fresh++;
            break;
          }
          if (old == k) break;
          h = (h + 1) & (hslots - 1);
        }
      }
    }
    // keys 0xffffffff and 0xfffffffe were counted as one: both present -> one more voxel
    bool has_e = false, has_e1 = false;
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      const uint32_t l = i * kBktThreads + tid;
      has_e = has_e || (l < m && keys[i] == kEmpty);
      has_e1 = has_e1 || (l < m && keys[i] == kEmpty - 1);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, d));
      cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
      fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
    }
    if (lane == 0) {
      if (mx) atomicMax(&sm.maxkey, mx);
      atomicMin(&sm.minchunk, cmin);
      atomicMax(&sm.maxchunk, cmax);
      if (fresh) atomicAdd(&sm.distinct, fresh);
    }
    const int both = __syncthreads_or(has_e ? 1 : 0) && __syncthreads_or(has_e1 ? 1 : 0);
    total_heads = sm.distinct + (both ? 1u : 0u);
  }
  if (tid == 0) stv[bucket] = ((bucket == 0 ? 2ull : 1ull) << 62) | (unsigned long long)total_heads;
  if (m > 0) {
    // ---- 2. runs (blocked): a run = consecutive records of one chunk (one warp's reservation)
    uint32_t headmask = 0, cnt = 0;
    {
      uint32_t prev = l0 > 0 && l0 - 1 < m ? sm.key[l0 - 1] : 0xffffffffu;
#pragma unroll
      for (int j = 0; j < kBktIpt; j++) {
        const uint32_t l = l0 + j;
        if (l < m) {
          const uint32_t c = sm.key[l];
          const bool h = l == 0 || c != prev;
          headmask |= (h ? 1u : 0u) << j;
          cnt += h ? 1u : 0u;
          prev = c;
        }
      }
    }
    uint32_t runs = 0;
    const uint32_t rexcl = bkt_excl_scan(cnt, sm.scan, &runs);  // (its barriers also retire the hash set)
    uint32_t my_chunk[kBktIpt];
    {
      uint32_t r = rexcl;  // runs before this thread's first head; a record's run = heads up to and including it - 1
#pragma unroll
      for (int j = 0; j < kBktIpt; j++) {
        const uint32_t l = l0 + j;
        my_chunk[j] = 0;
        if (l < m) {
          if ((headmask >> j) & 1u) {
            my_chunk[j] = sm.key[l];
            sm.u.run.chunk[r] = my_chunk[j];
            sm.u.run.start[r] = (uint16_t)l;
            r++;
          }
          sm.u.run.rec_run[l] = (uint16_t)(r - 1);
        }
      }
    }
    __syncthreads();
    const bool ordered = runs <= 1;
    if (!ordered) {
      // ---- 3. rank of every run by chunk id (unique per run): radix sort of (chunk - min chunk, run)
      const uint32_t minchunk = sm.minchunk;
      const int cbits = 32 - __clz(sm.maxchunk - minchunk);  // >= 1: two runs have different chunks
      {
        uint32_t r = rexcl;
#pragma unroll
        for (int j = 0; j < kBktIpt; j++) {
          if ((headmask >> j) & 1u) {
            sm.key[r] = my_chunk[j] - minchunk;
            sm.pos[r] = (uint16_t)r;
            r++;
          }
        }
      }
      __syncthreads();
      bkt_radix_sort(sm, runs, cbits);
      // ---- 4. first position of every run in index order: exclusive scan of the lengths by rank
      uint32_t len[kBktIpt], sum = 0;
#pragma unroll
      for (int j = 0; j < kBktIpt; j++) {
        len[j] = 0;
        if (l0 + j < runs) {
          const uint32_t r = sm.pos[l0 + j];
          sm.u.run.rank[r] = (uint16_t)(l0 + j);
          len[j] = (r + 1 < runs ? (uint32_t)sm.u.run.start[r + 1] : m) - sm.u.run.start[r];
        }
        sum += len[j];
      }
      uint32_t e = bkt_excl_scan(sum, sm.scan, nullptr);
#pragma unroll
      for (int j = 0; j < kBktIpt; j++) {
        if (l0 + j < runs) sm.u.run.nstart[l0 + j] = (uint16_t)e;
        e += len[j];
      }
      __syncthreads();
    }
    // ---- 5. records in index order -> staging (key, position in the region)
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      const uint32_t l = i * kBktThreads + tid;
      if (l < m) {
        uint32_t np = l;
        if (!ordered) {
          const uint32_t r = sm.u.run.rec_run[l];
          np = (uint32_t)sm.u.run.nstart[sm.u.run.rank[r]] + (l - sm.u.run.start[r]);
        }
        sm.key[np] = keys[i];
        sm.pos[np] = (uint16_t)l;
      }
    }
    __syncthreads();
    // ---- 6. stable LSD radix sort on the bits the keys of this bucket use
    const uint32_t maxkey = sm.maxkey;
    bkt_radix_sort(sm, m, maxkey ? 32 - __clz(maxkey) : 0);
    // ---- 7. points in sorted order (the run table is dead)
#pragma unroll
    for (int i = 0; i < kBktIpt; i++) {
      const uint32_t l = i * kBktThreads + tid;
      if (l < m) {
        const float4 r4 = __ldg(&brec[sm.pos[l]]);
        sm.u.xyz[0][l] = r4.x;
        sm.u.xyz[1][l] = r4.y;
        sm.u.xyz[2][l] = r4.z;
      }
    }
    // ---- 8. voxel heads (blocked), compacted
    headmask = 0;
    cnt = 0;
    {
      uint32_t prev = l0 > 0 && l0 - 1 < m ? sm.key[l0 - 1] : 0u;
#pragma unroll
      for (int j = 0; j < kBktIpt; j++) {
        const uint32_t l = l0 + j;
        if (l < m) {
          const uint32_t k = sm.key[l];
          const bool h = l == 0 || k != prev;
          headmask |= (h ? 1u : 0u) << j;
          cnt += h ? 1u : 0u;
          prev = k;
        }
      }
    }
    const uint32_t hexcl = bkt_excl_scan(cnt, sm.scan, nullptr);  // also orders the xyz stores before the walk
    {
      uint16_t* s_src = &sm.hist[0][0];
      uint32_t r = hexcl;
#pragma unroll
      for (int j = 0; j < kBktIpt; j++)
        if ((headmask >> j) & 1u) s_src[r++] = (uint16_t)(l0 + j);
    }
  }
  // ---- 9. output slot of the bucket's first voxel: look-back over the preceding buckets, one status word per
  // thread (every running bucket has published its count already; only finished ones hold inclusive sums)
  unsigned long long prefix = 0;
  if (bucket > 0) {
    int64_t hi = (int64_t)bucket - 1;  // window [hi - 511, hi], thread t reads hi - t
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
    if (tid == 0) stv[bucket] = (2ull << 62) | (prefix + total_heads);
  }
  if (tid == 0 && bucket == n_buckets - 1) st->n_out = (long long)(prefix + total_heads);
  __syncthreads();
  if (m == 0) return;
  // ---- 10. one voxel per thread per round: members added in list order (voxelgrid.go:148-158), record of the
  // first member with the centroid (voxelgrid.go:173-184)
  const VgParams& P = st->P;
  const int out_aligned = v.aligned && ((((uintptr_t)out) & 3) == 0);
  const bool xyz_only = out_aligned && v.packed && v.stride == 12;
  const unsigned long long kmin = bucket ? (unsigned long long)__ldg(&spl[bucket - 1]) << st->shift : 0ull;
  const uint16_t* s_src = &sm.hist[0][0];
  const int key_bits = P.key_bits;
  long long vc_cid = -1;
  float vc[3] = {0.f, 0.f, 0.f};
  for (uint32_t r = tid; r < total_heads; r += kBktThreads) {
    const uint32_t l = s_src[r];
    const uint32_t end = r + 1 < total_heads ? s_src[r + 1] : m;
    const unsigned long long key = kmin + sm.key[l];
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
      const uint32_t first = __float_as_uint(__ldg(&brec[sm.pos[l]].w));
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
  DevBuf<uint32_t> spl(std::max<uint32_t>(n_buckets, 1u), stream);
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
    PCG_LAUNCH(splitters_kernel, div_up(n_buckets, 256), 256, 0, stream, kk[res], n_samples, n_buckets, spl.p);
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
