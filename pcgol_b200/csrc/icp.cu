// icp.cu — point-to-point ICP by gradient descent (pc/registration/icp) on sm_100a.
//
// One iteration of PointToPointICPGradient.Fit (icp.go:48-65):
//   icp_terms_kernel   per target point: Mat4.Transform with the accumulated transform
//                      (icp.go:62-64), exact nearest neighbour in the base index
//                      (correspondence.go:22-37) and the nine per-pair terms of
//                      Evaluate (evaluator.go:130-144), fused — the pair list never exists.
//   FAST mode          every CTA folds its terms into float64 partials; the CTA that finishes
//                      last (ticket counter) adds the partials in a fixed order (deterministic,
//                      GPU-count independent up to float64 rounding) and runs the tail — one
//                      kernel per iteration.
//   STRICT mode        the terms are stored at the target's index and the exact replay (see "fast exact
//                      replay" below: binade-parallel integer sums, element-wise only around binade
//                      crossings) reproduces the reference's sequential float32 accumulation:
//                      bit-identical trajectory.  icp_replay_kernel is the plain sequential chain.
//   tail               evaluator.go:156-186 + gradientDescentUpdater.Update (updater.go:44-71)
//                      run on the device, so the loop never returns to the host; once `done`
//                      is set the remaining launches fall through.
#include <algorithm>
#include <mutex>
#include <thread>
#include <exception>
#include <cstdlib>
#include <vector>

#include "bulk_async.cuh"
#include "bvh.cuh"
#include "icp_math.cuh"

namespace pcg {

struct IcpState {
  im::M4 trans;
  im::UpdaterCfg cfg;
  int32_t min_pairs;
  int32_t iter;           // gradientDescentUpdater.i
  int32_t num_iteration;  // Stat.NumIteration
  int32_t done;
  int32_t status;
  int32_t evaluate_only;
  im::Eval ev;
  long long n_pairs;
  unsigned int pair_counter;
  unsigned int ticket;   // CTAs that have published their part of the current reduction
  float sums[12];        // the nine accumulators of the current Evaluate (strict replay, one CTA each)
  // normal equations (never read unless want_hessian / Gauss-Newton)
  int32_t want_hessian;
  int32_t weight_fn;      // pcg_weight_fn (EvaluateWeightFn family, evaluator.go:19-23)
  float weight_param;
  float pad_;
  double hsum[9];        // sum x,y,z,xx,xy,xz,yy,yz,zz of the matched, transformed target points
  float hess[36];        // Evaluated.Hessian of the last Evaluate
};

// Block size of the fused correspondence kernel.  Measured on B200 (100k-point iteration): the walk itself likes
// small blocks (strict mode, no block reduction: 38.6 us at 32 threads, 49 at 128), the float64 block sums + the
// last block's fold over all partials like few, large blocks (fast mode: 41.6 us at 256 threads, 46 at 128, 73 at 32).
constexpr int kTermThreadsStrict = 32;
constexpr int kTermThreadsReduce = 256;
inline int term_threads(int mode, bool hess) {
  return ((mode & ~PCG_ICP_WITH_HESSIAN) == PCG_ICP_STRICT && !hess) ? kTermThreadsStrict : kTermThreadsReduce;
}
constexpr int kTerms = 9;  // Value, SumW, G0..G5, R
constexpr int kHTerms = 9; // sum p (3) + second moments of p (6): the Gauss-Newton Hessian

// One large ICP over several GPUs driven by ONE process: exchange of the iteration's sums over NVLink peer memory,
// done by the last CTA of the terms kernel (see icp_last_block).  n_dev <= 1: no exchange.
constexpr int kMaxPeers = 8;
constexpr int kXchg = 10;  // Value, SumW, G0..5, R, nPairs
struct PeerCtx {
  int n_dev, rank;
  unsigned int seq_base;           // exchanges completed by earlier Fits on these buffers (flags only grow)
  double* slots[kMaxPeers];        // device d's buffer [2][kMaxPeers][kXchg], addressable from this device
  unsigned int* flags[kMaxPeers];  // device d's flags [kMaxPeers]: last exchange the source has published
};

__device__ __forceinline__ void icp_last_block(IcpState* __restrict__ st, const double* __restrict__ partials,
                                               const double* __restrict__ hpartials, int nblocks, bool do_main,
                                               bool do_hess, float* s_sum, double* s_tot, int* s_last,
                                               const PeerCtx& pc);

// Block-wide float64 sum of K per-thread values -> out[blockIdx.x * K + k] (fixed order).
template <int K, int THREADS>
__device__ __forceinline__ void block_sum_f64(const float* t, double (*s_red)[K], double* __restrict__ out) {
  const int tid = threadIdx.x;
  double d[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    d[k] = (double)t[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d[k] += __shfl_down_sync(0xffffffffu, d[k], o);
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) s_red[tid >> 5][k] = d[k];
  }
  __syncthreads();
  if (tid < K) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) s += s_red[w][tid];
    out[(int64_t)blockIdx.x * K + tid] = s;
  }
  __syncthreads();
}

// APPROX: KDTree.MinDistSq > 0 on the base search (kdtree.go:19-22).  HESS: also accumulate the
// nine moments of the normal equations (icp_math.cuh).
template <int MODE, bool APPROX, bool HESS, int THREADS>
__global__ void __launch_bounds__(THREADS)
    icp_terms_kernel(IndexView base, CloudView tgt, const uint32_t* __restrict__ perm, uint32_t* __restrict__ warm,
                     float max_dist_sq, float min_dist_sq, IcpState* __restrict__ st, float* __restrict__ terms,
                     int64_t n_pad, double* __restrict__ partials, double* __restrict__ hpartials, int finalize,
                     PeerCtx pc) {
  if (st->done) return;
  __shared__ float s_m[16];
  __shared__ int s_first;
  __shared__ int s_last;
  __shared__ float s_sum[kTerms];
  __shared__ double s_tot[kXchg];
  __shared__ double s_red[THREADS / 32][kTerms];
  const int tid = threadIdx.x;
  if (tid < 16) s_m[tid] = st->trans.m[tid];
  if (tid == 0) s_first = st->num_iteration == 0;
  const int weight_fn = st->weight_fn;
  const float weight_param = st->weight_param;
  __syncthreads();
  const int64_t slot = (int64_t)blockIdx.x * THREADS + tid;
  float t[kTerms];
  float ht[kHTerms];
#pragma unroll
  for (int k = 0; k < kTerms; k++) t[k] = 0.f;
#pragma unroll
  for (int k = 0; k < kHTerms; k++) ht[k] = 0.f;
  int matched = 0;
  if (slot < tgt.n) {
    // Morton-ordered visit (coherent warps); the terms still land at the target's own index,
    // so the sequential replay below adds them in target order
    const int64_t i = perm ? (int64_t)perm[slot] : slot;
    float3 p = load_xyz(tgt, i);
    float x0 = p.x, y0 = p.y, z0 = p.z;
    // icp.go:27-30: the first Evaluate sees the raw target; later ones the ORIGINAL target
    // moved by the accumulated transform (icp.go:62-64)
    if (!s_first) im::m4transform(s_m, p.x, p.y, p.z, &x0, &y0, &z0);
    uint64_t best = nn_init(max_dist_sq);
    const uint64_t init = best;
    uint32_t pos = 0;
    // warm start: the point this target matched in the previous iteration (the transform moves little per
    // iteration) is a real candidate, so starting from its distance leaves the (DistSq, ID) arg-min unchanged
    // and prunes almost all backtracking
    const uint32_t w0 = warm ? warm[slot] : 0xffffffffu;
    if (w0 != 0xffffffffu) {
      const float4 c = load_point(base.pts, w0);
      const float d = dist_sq_ref(c.x, c.y, c.z, x0, y0, z0);
      const uint64_t packed = ((uint64_t)__float_as_uint(d) << 32) | (uint64_t)__float_as_uint(c.w);
      if (packed < best) {
        best = packed;
        pos = w0;
      }
    }
    if (!(APPROX && best != init && __uint_as_float((uint32_t)(best >> 32)) < min_dist_sq))
      nn_traverse4<APPROX>(base, x0, y0, z0, best, pos, min_dist_sq);
    if (warm && best != init) warm[slot] = pos;
    if (best != init) {  // correspondence.go:27-29
      matched = 1;
      const float4 pb = load_point(base.pts, pos);
      const float x1 = pb.x, y1 = pb.y, z1 = pb.z;
      const float dsq = __uint_as_float((uint32_t)(best >> 32));
      // evaluator.go:130-144; with w = 1 (DefaultEvaluateWeightFn) w*v == v exactly
      t[0] = dsq;
      t[1] = 1.f;
      t[2] = im::sub(x0, x1);
      t[3] = im::sub(y0, y1);
      t[4] = im::sub(z0, z1);
      t[5] = im::sub(im::mul(z0, y1), im::mul(y0, z1));
      t[6] = im::sub(im::mul(x0, z1), im::mul(z0, x1));
      t[7] = im::sub(im::mul(y0, x1), im::mul(x0, y1));
      t[8] = im::add(im::add(im::mul(x0, x0), im::mul(y0, y0)), im::mul(z0, z0));
      if (weight_fn != PCG_WEIGHT_CONSTANT) {  // the parametric EvaluateWeightFn family (pcg_weight_fn)
        float w;
        if (weight_fn == PCG_WEIGHT_TRUNCATED)
          w = dsq < weight_param ? 1.f : 0.f;
        else
          w = dsq <= weight_param ? 1.f : __fsqrt_rn(im::div(weight_param, dsq));
        t[1] = w;
#pragma unroll
        for (int k = 0; k < kTerms; k++)
          if (k != 1) t[k] = im::mul(w, t[k]);
      }
      if (HESS) {
        ht[0] = x0;
        ht[1] = y0;
        ht[2] = z0;
        ht[3] = im::mul(x0, x0);
        ht[4] = im::mul(x0, y0);
        ht[5] = im::mul(x0, z0);
        ht[6] = im::mul(y0, y0);
        ht[7] = im::mul(y0, z0);
        ht[8] = im::mul(z0, z0);
      }
    }
    if (MODE == PCG_ICP_STRICT) {
      // unmatched targets contribute +0: x + (+0) == x for every partial sum the reference can hold
#pragma unroll
      for (int k = 0; k < kTerms; k++) terms[(int64_t)k * n_pad + i] = t[k];
    }
  }
  const int block_pairs = __syncthreads_count(matched);
  if (tid == 0 && block_pairs) atomicAdd(&st->pair_counter, (unsigned int)block_pairs);
  if (HESS) block_sum_f64<kHTerms, THREADS>(ht, s_red, hpartials);
  if (MODE == PCG_ICP_FAST) block_sum_f64<kTerms, THREADS>(t, s_red, partials);
  if (finalize && (MODE == PCG_ICP_FAST || HESS))
    icp_last_block(st, partials, hpartials, (int)gridDim.x, MODE == PCG_ICP_FAST, HESS, s_sum, s_tot, &s_last, pc);
}

constexpr int kFinishThreads = 32 * kTerms;
constexpr int kChunk = 512;  // floats per staged chunk

__device__ __forceinline__ float4 load_terms4(const float* __restrict__ s, int64_t idx, int64_t n) {
  if (idx + 3 < n) return *reinterpret_cast<const float4*>(s + idx);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx < n) v.x = s[idx];
  if (idx + 1 < n) v.y = s[idx + 1];
  if (idx + 2 < n) v.z = s[idx + 2];
  return v;
}

// Tail of Evaluate + Update from the nine sums (one thread).
__device__ __forceinline__ void icp_finalize(IcpState* __restrict__ st, const float* sum9) {
  const long long n_pairs = (long long)st->pair_counter;
  st->pair_counter = 0;
  st->n_pairs = n_pairs;
  st->num_iteration++;                        // icp.go:50
  if (n_pairs < (long long)st->min_pairs) {   // evaluator.go:97-106
    st->status = PCG_E_NOT_ENOUGH_PAIRS;
    st->done = 1;
    return;
  }
  im::Sums sums;
  sums.value = sum9[0];
  sums.sum_weight = sum9[1];
  for (int k = 0; k < 6; k++) sums.g[k] = sum9[2 + k];
  sums.rms = sum9[8];
  const im::Eval ev = im::evaluate_tail(sums);
  st->ev = ev;                                // icp.go:54
  im::HSums hs = {};
  if (st->want_hessian) {
    hs.n = (double)n_pairs;
    for (int k = 0; k < 3; k++) hs.p[k] = st->hsum[k];
    for (int k = 0; k < 6; k++) hs.pp[k] = st->hsum[3 + k];
    im::hessian_from_sums(hs, sums.sum_weight, st->hess);
  }
  if (st->evaluate_only) {
    st->done = 1;
    return;
  }
  im::M4 trans = st->trans;
  int iter = st->iter;
  const bool converged = st->cfg.kind == PCG_UPDATER_GAUSS_NEWTON
                             ? im::updater_update_gn(st->cfg, &iter, &trans, ev, sums, hs)
                             : im::updater_update(st->cfg, &iter, &trans, ev);  // icp.go:57
  st->trans = trans;
  st->iter = iter;
  if (converged) st->done = 1;
}

// Called by every CTA of icp_terms_kernel after it wrote its partials; the last one to arrive
// reduces all of them (fixed order: lane-strided columns, then a shuffle tree).  do_main: the nine
// Evaluate sums (FAST mode) followed by the tail; do_hess: the nine moments of the normal equations.
__device__ __forceinline__ void icp_last_block(IcpState* __restrict__ st, const double* __restrict__ partials,
                                               const double* __restrict__ hpartials, int nblocks, bool do_main,
                                               bool do_hess, float* s_sum, double* s_tot, int* s_last,
                                               const PeerCtx& pc) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __threadfence();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&st->ticket, 1u);
    *s_last = (t == (unsigned int)nblocks - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  if (do_hess) {
    for (int k = warp; k < kHTerms; k += (int)(blockDim.x >> 5)) {
      double s = 0.0;
      for (int b = lane; b < nblocks; b += 32) s += __ldcg(&hpartials[(int64_t)b * kHTerms + k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) st->hsum[k] = s;
    }
  }
  if (do_main) {
    for (int k = warp; k < kTerms; k += (int)(blockDim.x >> 5)) {
      double s = 0.0;
      for (int b = lane; b < nblocks; b += 32) s += __ldcg(&partials[(int64_t)b * kTerms + k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) s_tot[k] = s;
    }
  }
  __syncthreads();
  if (do_main && pc.n_dev > 1) {
    // ---- exchange over NVLink peer memory: this device's ten sums are STORED into slot [parity][rank] of every
    // device's buffer (plain stores to peer-mapped addresses), fenced, then its flag is raised on every device;
    // it then waits until every device's flag shows this exchange and adds the slots in device order - the same
    // order everywhere, so all devices apply the identical update without a broadcast.  Slots are double-buffered by
    // iteration parity: nobody can be more than one exchange ahead of the slowest device.
    const unsigned int it = (unsigned int)st->num_iteration;
    const unsigned int seq = pc.seq_base + it + 1u;
    const int parity = (int)(it & 1u);
    if (tid == 0) s_tot[9] = (double)st->pair_counter;
    __syncthreads();
    for (int idx = tid; idx < pc.n_dev * kXchg; idx += (int)blockDim.x) {
      const int d = idx / kXchg, k = idx % kXchg;
      volatile double* dst = pc.slots[d] + ((size_t)parity * kMaxPeers + pc.rank) * kXchg + k;
      *dst = s_tot[k];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pc.n_dev) {
      volatile unsigned int* f = pc.flags[tid] + pc.rank;
      *f = seq;
      volatile unsigned int* mine = pc.flags[pc.rank] + tid;
      while ((int)(*mine - seq) < 0) {  // flags only grow (wrap-safe comparison)
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid < kXchg) {
      double s = 0.0;
      const volatile double* src = pc.slots[pc.rank] + (size_t)parity * kMaxPeers * kXchg;
      for (int d = 0; d < pc.n_dev; d++) s += src[(size_t)d * kXchg + tid];
      s_tot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) st->pair_counter = (unsigned int)(long long)s_tot[9];  // icp_finalize takes the pair count from here
  }
  if (do_main && tid < kTerms) s_sum[tid] = (float)s_tot[tid];
  __syncthreads();
  if (tid == 0) {
    st->ticket = 0;
    if (do_main) icp_finalize(st, s_sum);
  }
}

// STRICT mode: CTA k (one warp) owns accumulator k.  All lanes stream the next chunk from
// global memory while lane 0 adds the current one in order out of shared memory: the float32
// sum is the reference's (evaluator.go:130-144).  Nine CTAs land on nine SMs, so each
// dependent FADD chain has an issue port to itself; the last CTA to finish runs the tail.
__global__ void __launch_bounds__(32)
    icp_replay_kernel(IcpState* __restrict__ st, const float* __restrict__ terms, int64_t n, int64_t n_pad) {
  if (st->done) return;
  __shared__ __align__(16) float s_buf[2][kChunk];
  const int lane = threadIdx.x, k = blockIdx.x;
  const float* __restrict__ src = terms + (int64_t)k * n_pad;
  const int64_t nchunks = (n + kChunk - 1) / kChunk;
  float acc = 0.f;
  float4 r[kChunk / 128];
#pragma unroll
  for (int j = 0; j < kChunk / 128; j++) r[j] = load_terms4(src, (int64_t)j * 128 + lane * 4, n);
#pragma unroll
  for (int j = 0; j < kChunk / 128; j++) *reinterpret_cast<float4*>(&s_buf[0][j * 128 + lane * 4]) = r[j];
  __syncwarp();
  for (int64_t c = 0; c < nchunks; c++) {
    const bool more = c + 1 < nchunks;
    if (more) {
#pragma unroll
      for (int j = 0; j < kChunk / 128; j++) r[j] = load_terms4(src, (c + 1) * kChunk + (int64_t)j * 128 + lane * 4, n);
    }
    if (lane == 0) {
      const float4* b4 = reinterpret_cast<const float4*>(&s_buf[c & 1][0]);
#pragma unroll 16
      for (int j = 0; j < kChunk / 4; j++) {
        const float4 v = b4[j];
        acc = __fadd_rn(acc, v.x);
        acc = __fadd_rn(acc, v.y);
        acc = __fadd_rn(acc, v.z);
        acc = __fadd_rn(acc, v.w);
      }
    }
    __syncwarp();
    if (more) {
#pragma unroll
      for (int j = 0; j < kChunk / 128; j++)
        *reinterpret_cast<float4*>(&s_buf[(c + 1) & 1][j * 128 + lane * 4]) = r[j];
    }
    __syncwarp();
  }
  if (lane == 0) {
    st->sums[k] = acc;
    __threadfence();
    const unsigned int t = atomicAdd(&st->ticket, 1u);
    if (t == (unsigned int)kTerms - 1u) {
      __threadfence();
      float sum9[kTerms];
      for (int j = 0; j < kTerms; j++) sum9[j] = __ldcg(&st->sums[j]);
      st->ticket = 0;
      icp_finalize(st, sum9);
    }
  }
}

// ---- STRICT mode, fast exact replay -------------------------------------------------------------
// The reference adds the terms one by one in float32.  While the accumulator s stays inside one
// binade [2^e, 2^(e+1)) (either sign), s is an integer multiple M of ulp = 2^(e-23) and every step
// is fl(s + x) = (M + RN(x / ulp)) * ulp: the rounding only sees x - except for an exact tie
// (x / ulp = q + 1/2), where round-to-even looks at the parity of M + q and always leaves an even
// result.  So inside a binade the sequential float32 sum is an INTEGER sum driven by a two-state
// automaton (the parity of the running integer).  Elements are maps parity -> (increment, parity),
// their composition is associative, hence the sum is parallel:
//   1. a float64 prefix sum guesses the accumulator (only its binade matters) at every chunk start;
//   2. every chunk is folded in parallel (one warp, ordered tree composition) into, for both start
//      parities: the total increment and the min / max of the running prefix, in units of ulp;
//   3. one warp walks the chunks in order with the TRUE accumulator: if it lies in the guessed
//      binade and every prefix keeps it strictly inside (one unit of margin at both ends, because
//      the pre-rounded terms are off by at most half a unit), the chunk is ONE exact integer add;
//      otherwise the chunk is replayed element by element like the reference (from shared memory).
// The result is the reference's float32 sum, bit for bit, for any input; LiDAR residuals take the
// element-wise path only around the few binade crossings of each accumulator.
#ifndef PCG_REPLAY_CHUNK
#define PCG_REPLAY_CHUNK 256
#endif
constexpr int kReplayChunk = PCG_REPLAY_CHUNK;        // elements folded into one summary
constexpr int kReplayPerLane = kReplayChunk / 32;     // consecutive elements per lane when a warp folds a chunk
static_assert(kReplayChunk % 128 == 0 || kReplayChunk == 64 || kReplayChunk == 32, "chunk = whole float4 rows per warp");
constexpr int kReplayThreads = 512;
constexpr int kReplayBatch = 256;  // chunk summaries staged in shared memory per round of the walk
constexpr int kReplayStage = 128;  // chunks whose elements can wait in shared memory per round (128 KB)

// Summary of one chunk as the walk consumes it, in the float domain.  p = parity of the accumulator's integer
// mantissa M (acc = M * ulp, 2^23 <= |M| < 2^24).  The chunk is one exact add iff acc lies in [lo_pos[p], hi_pos[p]]
// (positive accumulator) or in [lo_neg[p], hi_neg[p]] (negative): the interval is the set of accumulators of the
// guessed binade that every running prefix keeps strictly inside it.  Then acc += t[p], t[p] = (integer total) *
// ulp: both operands and the result are multiples of ulp below 2^24 * ulp, so the float addition is exact.
// Unusable summaries hold empty intervals (lo = +inf, hi = -inf).
struct __align__(16) ReplayChunk {
  float t[2];
  float lo_pos[2], hi_pos[2], lo_neg[2], hi_neg[2];
  float pad_[2];
};

// Parity automaton of a run of elements: for start parity p, the integer increment, the parity
// afterwards and the extremes of the running increment.
struct ParityMap {
  long long off[2], mn[2], mx[2];
  int np;  // bit p = parity after the run when started with parity p
};
__device__ __forceinline__ ParityMap compose(const ParityMap& a, const ParityMap& b) {  // a first, then b
  ParityMap r;
  r.np = 0;
#pragma unroll
  for (int p = 0; p < 2; p++) {
    const int mid = (a.np >> p) & 1;
    r.off[p] = a.off[p] + b.off[mid];
    r.mn[p] = min(a.mn[p], a.off[p] + b.mn[mid]);
    r.mx[p] = max(a.mx[p], a.off[p] + b.mx[mid]);
    r.np |= ((b.np >> mid) & 1) << p;
  }
  return r;
}
__device__ __forceinline__ long long shfl_down_ll(long long v, int d) {
  return (long long)shfl_down_u64((uint64_t)v, d);
}

__device__ __forceinline__ int float_exponent(float f) { return (int)((__float_as_uint(f) >> 23) & 0xff) - 127; }

constexpr int kReplayChunkRecords = 1;

// Phase A (whole GPU): float64 sum of every chunk of every stream; one warp per chunk.
__global__ void __launch_bounds__(256)
    icp_replay_sums_kernel(const IcpState* __restrict__ st, const float* __restrict__ terms, int64_t n, int64_t n_pad,
                           int64_t nchunks, int streams, double* __restrict__ chunk_sums) {
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nchunks * streams) return;
  const int64_t k = w / nchunks, c = w - k * nchunks;
  const float* __restrict__ x = terms + k * n_pad;
  const int64_t base = c * kReplayChunk + lane * kReplayPerLane;
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < kReplayPerLane; j++) s += (base + j < n) ? (double)x[base + j] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) chunk_sums[w] = s;
}

// Phase B (whole GPU): parity-automaton summary of every chunk in units of the guessed binade's ulp.
__global__ void __launch_bounds__(256)
    icp_replay_summaries_kernel(const IcpState* __restrict__ st, const float* __restrict__ terms, int64_t n,
                                int64_t n_pad, int64_t nchunks, int streams, const double* __restrict__ chunk_sums,
                                ReplayChunk* __restrict__ chunks) {
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nchunks * streams) return;
  const int64_t k = w / nchunks, c = w - k * nchunks;
  const float* __restrict__ x = terms + k * n_pad;
  // guess of the accumulator at the chunk start: float64 sum of the preceding chunks (only its binade matters)
  double g = 0.0;
  for (int64_t j = lane; j < c; j += 32) g += __ldg(&chunk_sums[k * nchunks + j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
  const float guess = (float)g;
  const int e = float_exponent(guess);
  bool regular = guess != 0.f && e > -100 && e < 128;  // not zero / (near) denormal / inf / nan
  const double scale = regular ? __longlong_as_double((long long)(1023 + 23 - e) << 52) : 0.0;  // 2^(23-e), exact
  const int64_t base = c * kReplayChunk + lane * kReplayPerLane;
  ParityMap m;
  m.off[0] = m.off[1] = 0;
  m.mn[0] = m.mn[1] = 0x7fffffffffffffffll;
  m.mx[0] = m.mx[1] = -0x7fffffffffffffffll;
  m.np = 2;  // identity: parity p stays p
#pragma unroll
  for (int j = 0; j < kReplayPerLane; j++) {
    const float xv = (base + j < n) ? x[base + j] : 0.f;
    const double y = (double)xv * scale;  // exact: power-of-two scaling inside double's range
    double q = floor(y);
    const bool sane = fabs(y) < 1.0e12;    // else: far beyond the accumulator, inf or nan -> not this binade
    if (!sane) {
      regular = false;
      q = 0.0;
    }
    const double frac = sane ? y - q : 0.0;
    const long long qi = (long long)q;
    ParityMap el;
    if (frac == 0.5) {
      // tie: round to even -> the result M + off is even whatever the start parity
      el.off[0] = qi + (qi & 1);        // M even: M + qi even iff qi even
      el.off[1] = qi + ((qi + 1) & 1);  // M odd:  M + qi even iff qi odd
      el.np = 0;
    } else {
      const long long r = qi + (frac > 0.5 ? 1 : 0);
      el.off[0] = el.off[1] = r;
      el.np = (r & 1) ? 1 : 2;  // odd increment flips the parity (bit p = p ^ 1), even keeps it
    }
    el.mn[0] = el.mx[0] = el.off[0];
    el.mn[1] = el.mx[1] = el.off[1];
    m = compose(m, el);
  }
  // ordered tree composition across the warp: lane l <- (lanes l .. l+o-1) then (lanes l+o .. l+2o-1)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    ParityMap right;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      right.off[p] = shfl_down_ll(m.off[p], o);
      right.mn[p] = shfl_down_ll(m.mn[p], o);
      right.mx[p] = shfl_down_ll(m.mx[p], o);
    }
    right.np = __shfl_down_sync(0xffffffffu, m.np, o);
    if ((lane & (2 * o - 1)) == 0) m = compose(m, right);
  }
  const bool all_regular = __all_sync(0xffffffffu, regular);
  if (lane == 0) {
    ReplayChunk rc;
    const long long kLo = 1ll << 23, kHi = 1ll << 24;
    const float inf = __int_as_float(0x7f800000);
    const bool usable = all_regular && e <= 100;  // 2^24 * ulp must stay finite
    const float ulp = usable ? __uint_as_float((uint32_t)(e - 23 + 127) << 23) : 0.f;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      // every running value M + prefix must stay strictly inside the binade, with one unit of margin
      // (the pre-rounded terms are off by at most half a unit): positive M in [2^23+1, 2^24-1] - prefix, ...
      long long lo_pos = max(kLo + 1 - m.mn[p], kLo), hi_pos = min(kHi - 1 - m.mx[p], kHi - 1);
      long long lo_neg = max(-kHi + 1 - m.mn[p], -(kHi - 1)), hi_neg = min(-kLo - 1 - m.mx[p], -kLo);
      const bool tot_ok = m.off[p] > -kHi && m.off[p] < kHi;
      const bool pos_ok = usable && tot_ok && lo_pos <= hi_pos, neg_ok = usable && tot_ok && lo_neg <= hi_neg;
      // |integers| < 2^24 times a power of two: exact
      rc.t[p] = (usable && tot_ok) ? __fmul_rn((float)m.off[p], ulp) : 0.f;
      rc.lo_pos[p] = pos_ok ? __fmul_rn((float)lo_pos, ulp) : inf;
      rc.hi_pos[p] = pos_ok ? __fmul_rn((float)hi_pos, ulp) : -inf;
      rc.lo_neg[p] = neg_ok ? __fmul_rn((float)lo_neg, ulp) : inf;
      rc.hi_neg[p] = neg_ok ? __fmul_rn((float)hi_neg, ulp) : -inf;
    }
    // Will the walk have to replay this chunk element by element?  The float64 guess failing the interval test (for
    // either parity) predicts it (exactly, on every stream tried): the walk kernel then has the chunk's elements
    // waiting in shared memory.  A wrong prediction only costs the on-demand load.
    auto inside = [&](int p) {
      return (guess >= rc.lo_pos[p] && guess <= rc.hi_pos[p]) || (guess >= rc.lo_neg[p] && guess <= rc.hi_neg[p]);
    };
    rc.pad_[0] = (inside(0) && inside(1)) ? 0.f : 1.f;
    rc.pad_[1] = 0.f;
    chunks[w] = rc;
  }
}

// Phase C (one CTA per stream): in-order walk with the true accumulator.  The chunk summaries are
// staged in shared memory by the whole CTA; warp 0 walks (every lane carries the same accumulator).
__global__ void __launch_bounds__(kReplayThreads)
    icp_replay_walk_kernel(IcpState* __restrict__ st, const float* __restrict__ terms, int64_t n, int64_t n_pad,
                           int64_t nchunks, const ReplayChunk* __restrict__ chunks_all) {
  if (st->done) return;
  extern __shared__ __align__(128) float s_stage[];  // [kReplayStage][kReplayChunk]: elements of the chunks to replay
  __shared__ __align__(16) float s_x[kReplayChunk];
  __shared__ ReplayChunk s_chunks[kReplayBatch];
  __shared__ short s_slot[kReplayBatch];  // staging slot of the batch's i-th chunk, -1: not staged
  __shared__ int s_wcount[kReplayBatch / 32];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.x;
  const float* __restrict__ x = terms + (int64_t)k * n_pad;
  const ReplayChunk* __restrict__ chunks = chunks_all + (int64_t)k * nchunks;
  const bool can_stage = ((reinterpret_cast<uintptr_t>(x)) & 15) == 0;  // bulk copies need 16-byte aligned rows
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float acc = 0.f;
  unsigned int n_fast = 0, n_slow = 0;
  uint32_t stage_phase = 0;
  const long long t_walk0 = clock64();
  for (int64_t c0 = 0; c0 < nchunks; c0 += kReplayBatch) {
    const int batch = (int)min((int64_t)kReplayBatch, nchunks - c0);
    __syncthreads();
    {
      const int4* src = reinterpret_cast<const int4*>(chunks + c0);
      int4* dst = reinterpret_cast<int4*>(s_chunks);
      const int words = batch * (int)(sizeof(ReplayChunk) / sizeof(int4));
      for (int i = tid; i < words; i += kReplayThreads) dst[i] = src[i];
    }
    __syncthreads();
    // The chunks the walk is predicted to replay arrive as bulk-async copies (one 1 KB cp.async.bulk each, all in
    // flight together, completion counted on an mbarrier) while the walk is still in its exact-add steps.
    bool staged_any = false;
    {
      bool want = false;
      if (tid < batch)
        want = can_stage && s_chunks[tid].pad_[0] != 0.f && (c0 + tid + 1) * (int64_t)kReplayChunk <= n;
      const unsigned m = __ballot_sync(0xffffffffu, want);
      if (tid < kReplayBatch && lane == 0) s_wcount[warp] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kReplayBatch / 32; w++) {
        const int c = s_wcount[w];
        before += w < warp ? c : 0;
        total += c;
      }
      const int slot = before + __popc(m & ((1u << lane) - 1u));
      const bool take = want && slot < kReplayStage;
      if (tid < kReplayBatch) s_slot[tid] = take ? (short)slot : (short)-1;
      const int nstaged = min(total, kReplayStage);
      staged_any = nstaged > 0;
      if (staged_any) {
        if (tid == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slots were last read by ordinary loads
          mbar_expect_tx(&s_bar, (uint32_t)nstaged * kReplayChunk * (uint32_t)sizeof(float));
        }
        __syncthreads();  // the expectation is posted before any copy can complete the phase
        if (take)
          bulk_g2s(s_stage + (size_t)slot * kReplayChunk, x + (c0 + tid) * (int64_t)kReplayChunk,
                   kReplayChunk * (uint32_t)sizeof(float), &s_bar);
      }
      __syncthreads();  // s_slot is complete
    }
    if (warp == 0) {
      if (staged_any) {
        mbar_wait(&s_bar, stage_phase & 1u);
        stage_phase++;
      }
      // the reference's own loop over one chunk (from shared memory)
      auto replay = [&](float a, int64_t chunk) {
        const int64_t base = chunk * kReplayChunk;
        const int slot = s_slot[(int)(chunk - c0)];
        if (slot < 0) {  // not predicted (or no room): load now
#pragma unroll
          for (int j = 0; j < kReplayChunk / 32; j++) {
            const int64_t i = base + j * 32 + lane;
            s_x[j * 32 + lane] = i < n ? x[i] : 0.f;
          }
          __syncwarp();
        }
        const float4* b4 = reinterpret_cast<const float4*>(slot < 0 ? s_x : s_stage + (size_t)slot * kReplayChunk);
#pragma unroll 8
        for (int j = 0; j < kReplayChunk / 4; j++) {
          const float4 v = b4[j];
          a = __fadd_rn(a, v.x);
          a = __fadd_rn(a, v.y);
          a = __fadd_rn(a, v.z);
          a = __fadd_rn(a, v.w);
        }
        __syncwarp();
        return a;
      };
      // 32 chunks per round, one per lane.  The dependency between chunks is three instructions - parity of the
      // accumulator, select the chunk's total for that parity, add - so that chain runs first (the totals come from
      // shared memory, independent of the accumulator), every lane keeping the accumulator its own chunk starts from;
      // the interval tests (ten comparisons per chunk) then run side by side, one chunk per lane.  The chain stops in
      // front of every chunk the summaries kernel predicted to need a replay, so nothing is computed twice; a chunk
      // that fails its test unpredicted is replayed from the accumulator it really starts with, and the chain resumes
      // behind it.
      for (int g0 = 0; g0 < batch; g0 += 32) {
        const int cnt = min(32, batch - g0);
        float t0 = 0.f, t1 = 0.f, lp0 = 0.f, lp1 = 0.f, hp0 = 0.f, hp1 = 0.f, ln0 = 0.f, ln1 = 0.f, hn0 = 0.f, hn1 = 0.f;
        bool predicted = false;
        if (lane < cnt) {
          const float4* q = reinterpret_cast<const float4*>(&s_chunks[g0 + lane]);
          const float4 qa = q[0], qb = q[1], qc = q[2];  // {t0,t1,lp0,lp1} {hp0,hp1,ln0,ln1} {hn0,hn1,flag,-}
          t0 = qa.x, t1 = qa.y, lp0 = qa.z, lp1 = qa.w;
          hp0 = qb.x, hp1 = qb.y, ln0 = qb.z, ln1 = qb.w;
          hn0 = qc.x, hn1 = qc.y;
          predicted = qc.z != 0.f;
        }
        auto passes = [&](float a) {  // this lane's chunk: one exact add from accumulator a?
          const bool p = (__float_as_uint(a) & 1u) != 0;
          const float lp = p ? lp1 : lp0, hp = p ? hp1 : hp0, ln = p ? ln1 : ln0, hn = p ? hn1 : hn0;
          return (a >= lp && a <= hp) || (a >= ln && a <= hn);
        };
        const unsigned predmask = __ballot_sync(0xffffffffu, predicted);
        int start = 0;
        while (start < cnt) {
          const unsigned ahead = predmask & (0xffffffffu << start);
          const int stop = ahead ? __ffs(ahead) - 1 : cnt;  // first predicted replay at or behind start
          float mine = 0.f;  // the accumulator this lane's chunk starts from
          float a = acc;
#pragma unroll 4
          for (int j = start; j < stop; j++) {
            const float2 t = *reinterpret_cast<const float2*>(&s_chunks[g0 + j]);
            if ((int)lane == j) mine = a;
            a = __fadd_rn(a, (__float_as_uint(a) & 1u) ? t.y : t.x);
          }
          const unsigned bad = __ballot_sync(0xffffffffu, (int)lane >= start && (int)lane < stop && !passes(mine));
          if (bad != 0u) {  // not predicted: replay the first failing chunk from its true accumulator
            const int f = __ffs(bad) - 1;
            n_fast += (unsigned)(f - start);
            acc = replay(__shfl_sync(0xffffffffu, mine, f), c0 + g0 + f);
            n_slow++;
            start = f + 1;
            continue;
          }
          acc = a;
          n_fast += (unsigned)(stop - start);
          if (stop < cnt) {  // the predicted chunk, from the true accumulator (it may pass after all)
            const bool ok = (__ballot_sync(0xffffffffu, passes(acc)) >> stop) & 1u;
            if (ok) {
              const float s0 = __shfl_sync(0xffffffffu, t0, stop), s1 = __shfl_sync(0xffffffffu, t1, stop);
              acc = __fadd_rn(acc, (__float_as_uint(acc) & 1u) ? s1 : s0);
              n_fast++;
            } else {
              acc = replay(acc, c0 + g0 + stop);
              n_slow++;
            }
          }
          start = stop + 1;
        }
      }
    }
  }
  if (tid == 0) {
    st->sums[k] = acc;
    if (gridDim.x == 1) {  // test hook launch: expose the walk statistics
      st->sums[9] = (float)n_fast;
      st->sums[10] = (float)n_slow;
      st->sums[11] = (float)(clock64() - t_walk0);
    }
    __threadfence();
    const unsigned int t = atomicAdd(&st->ticket, 1u);
    if (t == (unsigned int)kTerms - 1u) {
      __threadfence();
      float sum9[kTerms];
      for (int j = 0; j < kTerms; j++) sum9[j] = __ldcg(&st->sums[j]);
      st->ticket = 0;
      icp_finalize(st, sum9);
    }
  }
}


// Launches the three phases for `streams` accumulators (9 in a Fit, 1 from the test hook).
static void launch_exact_replay(IcpState* st, const float* terms, int64_t n, int64_t n_pad, int streams,
                                ReplayChunk* chunks, double* sums, cudaStream_t stream) {
  const int64_t nchunks = std::max<int64_t>(1, (n + kReplayChunk - 1) / kReplayChunk);
  const int blocks = div_up(nchunks * streams, 256 / 32);
  PCG_LAUNCH(icp_replay_sums_kernel, blocks, 256, 0, stream, st, terms, n, n_pad, nchunks, streams, sums);
  PCG_LAUNCH(icp_replay_summaries_kernel, blocks, 256, 0, stream, st, terms, n, n_pad, nchunks, streams, sums, chunks);
  constexpr size_t kStageBytes = (size_t)kReplayStage * kReplayChunk * sizeof(float);
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  PCG_CUDA(cudaGetDevice(&dev));
  if (!(configured.load(std::memory_order_relaxed) & (1ull << dev))) {
    PCG_CUDA(cudaFuncSetAttribute(icp_replay_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStageBytes));
    configured.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  PCG_LAUNCH(icp_replay_walk_kernel, streams, kReplayThreads, kStageBytes, stream, st, terms, n, n_pad, nchunks, chunks);
}

// Sharded ICP: fold the per-CTA float64 partials into 16 doubles for the all-reduce.
__global__ void __launch_bounds__(kFinishThreads)
    icp_partial_reduce_kernel(IcpState* __restrict__ st, const double* __restrict__ partials, int nblocks,
                              double* __restrict__ out16) {
  if (st->done) return;  // device-resident sharded loop: nothing was accumulated
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(int64_t)b * kTerms + warp];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out16[warp] = s;
  if (tid == 0) {
    out16[9] = (double)st->pair_counter;
    st->pair_counter = 0;  // the next iteration of a resident loop starts from zero
    for (int k = 10; k < 16; k++) out16[k] = 0.0;
  }
}

// Device-resident sharded loop: tail of Evaluate + Update from the all-reduced sums, identical on every rank.
__global__ void icp_shard_finish_kernel(IcpState* __restrict__ st, const double* __restrict__ reduced16) {
  if (threadIdx.x != 0 || st->done) return;
  float sum9[kTerms];
  for (int k = 0; k < kTerms; k++) sum9[k] = (float)reduced16[k];
  st->pair_counter = (unsigned int)(long long)reduced16[9];  // icp_finalize takes the pair count from here
  icp_finalize(st, sum9);
}

struct IcpWork {
  DevBuf<IcpState> st;
  DevBuf<float> terms;
  DevBuf<double> partials;
  DevBuf<double> hpartials;  // [nblocks][9] when the normal equations are accumulated
  DevBuf<uint32_t> perm;
  DevBuf<uint32_t> warm;              // per visit slot: position in the index of the previous iteration's match
  DevBuf<ReplayChunk> replay_chunks;  // strict replay: [9][chunks]
  DevBuf<double> replay_sums;         // strict replay: [9][chunks]
  int nblocks = 0;
  int64_t n_pad = 0;
};

static IcpState make_state(const pcg_icp_params& prm, bool evaluate_only) {
  IcpState h;
  std::memset(&h, 0, sizeof(h));
  h.trans = im::m4translate(0.f, 0.f, 0.f);  // icp.go:47
  h.cfg = im::make_updater(prm);
  h.min_pairs = prm.min_pairs == 0 ? 6 : prm.min_pairs;  // evaluator.go:92-95
  h.evaluate_only = evaluate_only ? 1 : 0;
  h.want_hessian = ((prm.mode & PCG_ICP_WITH_HESSIAN) || prm.updater == PCG_UPDATER_GAUSS_NEWTON) ? 1 : 0;
  h.weight_fn = prm.weight_fn;
  h.weight_param = prm.weight_param;
  return h;
}

static void check_icp_params(const pcg_icp_params& prm) {
  const int mode = prm.mode & ~PCG_ICP_WITH_HESSIAN;
  if (mode != PCG_ICP_STRICT && mode != PCG_ICP_FAST) throw StatusError{PCG_E_INVALID_ARG, "unknown ICP mode"};
  if (prm.updater != PCG_UPDATER_GRADIENT_DESCENT && prm.updater != PCG_UPDATER_GAUSS_NEWTON)
    throw StatusError{PCG_E_INVALID_ARG, "unknown ICP updater"};
  if (!(prm.min_dist_sq >= 0.f)) throw StatusError{PCG_E_INVALID_ARG, "MinDistSq must be >= 0"};
  if (prm.weight_fn < PCG_WEIGHT_CONSTANT || prm.weight_fn > PCG_WEIGHT_HUBER)
    throw StatusError{PCG_E_INVALID_ARG, "unknown weight function"};
  if (prm.weight_fn != PCG_WEIGHT_CONSTANT && !(prm.weight_param > 0.f))
    throw StatusError{PCG_E_INVALID_ARG, "the weight function's parameter (a squared distance) must be > 0"};
  if (prm.weight_fn != PCG_WEIGHT_CONSTANT && ((prm.mode & PCG_ICP_WITH_HESSIAN) || prm.updater == PCG_UPDATER_GAUSS_NEWTON))
    throw StatusError{PCG_E_INVALID_ARG, "the Hessian extension needs the constant weight function"};
}

// One launch of the fused correspondence + terms kernel for the (mode, MinDistSq, Hessian) combination.
template <int MODE>
static void launch_terms(const Index& base, const CloudView& tgt, const uint32_t* perm, uint32_t* warm, float mdsq,
                         float min_dist_sq, bool hess, IcpState* st, float* terms, int64_t n_pad, double* partials, double* hpartials,
                         int nblocks, int finalize, cudaStream_t stream, const PeerCtx* peer = nullptr) {
  PeerCtx pc;
  std::memset(&pc, 0, sizeof(pc));
  if (peer) pc = *peer;
  const char* name = MODE == PCG_ICP_STRICT ? "(icp_terms_kernel<PCG_ICP_STRICT>)" : "(icp_terms_kernel<PCG_ICP_FAST>)";
  const bool approx = min_dist_sq > 0.f;
  min_dist_sq = fminf(min_dist_sq, mdsq);  // only a real hit can end a search early: see nearest_device
  // nblocks was sized with term_threads(MODE, hess): strict without the Hessian moments walks in 32-thread blocks
#define PCG_TERMS(A, H, T)                                                                                         \
  PCG_LAUNCH_NAMED(name, (icp_terms_kernel<MODE, A, H, T>), nblocks, T, 0, stream, base.view(), tgt, perm, warm,    \
                   mdsq, min_dist_sq, st, terms, n_pad, partials, hpartials, finalize, pc)
  constexpr int kT = MODE == PCG_ICP_STRICT ? kTermThreadsStrict : kTermThreadsReduce;
  if (approx && hess)
    PCG_TERMS(true, true, kTermThreadsReduce);
  else if (approx)
    PCG_TERMS(true, false, kT);
  else if (hess)
    PCG_TERMS(false, true, kTermThreadsReduce);
  else
    PCG_TERMS(false, false, kT);
#undef PCG_TERMS
}

// Enqueues a whole Fit (or a single Evaluate) on `stream` without synchronising.
// Returns false if max_iteration exceeds what is enqueued at once (caller then polls).
constexpr int kMaxEnqueuedIterations = 64;

static void icp_enqueue_iterations(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, IcpWork& w,
                                   int iterations, cudaStream_t stream) {
  const float mdsq = prm.max_dist * prm.max_dist;  // kdtree.go:91
  const int mode = prm.mode & ~PCG_ICP_WITH_HESSIAN;
  const bool hess = (prm.mode & PCG_ICP_WITH_HESSIAN) || prm.updater == PCG_UPDATER_GAUSS_NEWTON;
  for (int it = 0; it < iterations; it++) {
    if (mode == PCG_ICP_STRICT) {
      launch_terms<PCG_ICP_STRICT>(base, tgt, w.perm.p, w.warm.p, mdsq, prm.min_dist_sq, hess, w.st.p, w.terms.p, w.n_pad,
                                   w.partials.p, w.hpartials.p, w.nblocks, 1, stream);
      launch_exact_replay(w.st.p, w.terms.p, tgt.n, w.n_pad, kTerms, w.replay_chunks.p, w.replay_sums.p, stream);
    } else {
      launch_terms<PCG_ICP_FAST>(base, tgt, w.perm.p, w.warm.p, mdsq, prm.min_dist_sq, hess, w.st.p, w.terms.p, w.n_pad,
                                 w.partials.p, w.hpartials.p, w.nblocks, 1, stream);
    }
  }
}

static void icp_prepare(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, bool evaluate_only,
                        IcpWork& w, cudaStream_t stream) {
  const bool hess = (prm.mode & PCG_ICP_WITH_HESSIAN) || prm.updater == PCG_UPDATER_GAUSS_NEWTON;
  base.wait(stream);  // the index may have been built on another stream
  w.nblocks = std::max(1, div_up(tgt.n, term_threads(prm.mode, hess)));
  if (tgt.n >= kMinQueriesToReorder && base.n > 0) {
    // the target moves by a small rigid transform per iteration: the order of the raw target stays coherent
    w.perm.alloc((size_t)tgt.n, stream);
    query_order_device(base, tgt, w.perm.p, stream);
  }
  w.n_pad = (tgt.n + 3) & ~(int64_t)3;
  w.st.alloc(1, stream);
  if (!evaluate_only && tgt.n > 0) {
    w.warm.alloc((size_t)tgt.n, stream);
    PCG_CUDA(cudaMemsetAsync(w.warm.p, 0xff, w.warm.bytes(), stream));
  }
  if ((prm.mode & PCG_ICP_WITH_HESSIAN) || prm.updater == PCG_UPDATER_GAUSS_NEWTON)
    w.hpartials.alloc((size_t)w.nblocks * kHTerms, stream);
  if ((prm.mode & ~PCG_ICP_WITH_HESSIAN) == PCG_ICP_STRICT) {
    w.terms.alloc((size_t)std::max<int64_t>(4, w.n_pad) * kTerms, stream);
    const size_t nchunks = (size_t)std::max<int64_t>(1, (tgt.n + kReplayChunk - 1) / kReplayChunk);
    w.replay_chunks.alloc(nchunks * kTerms * kReplayChunkRecords, stream);
    w.replay_sums.alloc(nchunks * kTerms, stream);
  } else
    w.partials.alloc((size_t)w.nblocks * kTerms, stream);
  IcpState h = make_state(prm, evaluate_only);
  PCG_CUDA(cudaMemcpyAsync(w.st.p, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
}

static void state_to_outputs(const IcpState& h, float trans[16], pcg_icp_stat* stat) {
  if (trans) std::memcpy(trans, h.trans.m, sizeof(float) * 16);
  if (stat) {
    std::memset(stat, 0, sizeof(*stat));
    stat->evaluated.value = h.ev.value;
    for (int k = 0; k < 6; k++) stat->evaluated.gradient[k] = h.ev.g[k];
    stat->evaluated.dist_rms = h.ev.dist_rms;
    if (h.want_hessian) std::memcpy(stat->evaluated.hessian, h.hess, sizeof(h.hess));
    stat->num_iteration = h.num_iteration;
    stat->n_pairs = h.n_pairs;
  }
}

// PointToPointICPGradient.Fit (icp.go:23-67) / Evaluate only. Synchronises `stream`.
pcg_status icp_fit_device(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, bool evaluate_only,
                          float trans[16], pcg_icp_stat* stat, cudaStream_t stream) {
  check_icp_params(prm);
  IcpWork w;
  icp_prepare(base, tgt, prm, evaluate_only, w, stream);
  const int total = evaluate_only ? 1 : im::make_updater(prm).max_iteration;
  IcpState h;
  int enq = 0;
  for (;;) {
    // every Update that does not converge increments i, so `total` Evaluate calls always suffice
    const int batch = std::min(kMaxEnqueuedIterations, std::max(1, total - enq));
    icp_enqueue_iterations(base, tgt, prm, w, batch, stream);
    enq += batch;
    PCG_CUDA(cudaMemcpyAsync(&h, w.st.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PCG_CUDA(cudaStreamSynchronize(stream));
    if (h.done) break;
  }
  state_to_outputs(h, trans, stat);
  return (pcg_status)h.status;
}

constexpr int kFarmMaxStreams = 32;
struct FarmResources {
  int device = -1, n_streams = 0;
  cudaStream_t streams[kFarmMaxStreams];
  cudaEvent_t start = nullptr, done[kFarmMaxStreams];
  IcpState* results = nullptr;  // pinned
  size_t capacity = 0;
};
static FarmResources& farm_resources(int device, int count) {
  constexpr int kMaxDevices = 64;
  if (device < 0 || device >= kMaxDevices) throw StatusError{PCG_E_INVALID_ARG, "device ordinal out of range"};
  static thread_local FarmResources per_device[kMaxDevices];  // never destroyed: the runtime may be gone at thread exit
  FarmResources& r = per_device[device];
  if (r.device != device) {
    r.n_streams = std::min(kFarmMaxStreams, 8);  // pairs in flight
    for (int s = 0; s < r.n_streams; s++) {
      PCG_CUDA(cudaStreamCreateWithFlags(&r.streams[s], cudaStreamNonBlocking));
      PCG_CUDA(cudaEventCreateWithFlags(&r.done[s], cudaEventDisableTiming));
    }
    PCG_CUDA(cudaEventCreateWithFlags(&r.start, cudaEventDisableTiming));
    r.device = device;
  }
  if ((size_t)count > r.capacity) {
    if (r.results) cudaFreeHost(r.results);
    r.results = nullptr;
    r.capacity = 0;
    const size_t cap = std::max<size_t>(64, (size_t)count * 2);
    PCG_CUDA(cudaMallocHost((void**)&r.results, cap * sizeof(IcpState)));
    r.capacity = cap;
  }
  return r;
}

// Scan-pair farm (BASELINE config 4): independent pairs round-robined over a few
// streams so that index builds and fits of different pairs overlap.
void icp_fit_pairs_device(int32_t count, const void* const* d_base, const int64_t* n_base,
                          const void* const* d_target, const int64_t* n_target, int64_t stride,
                          const int64_t xyz_off[3], const pcg_icp_params& prm, int device, float* trans_out,
                          pcg_icp_stat* stat_out, pcg_status* status_out, cudaStream_t stream) {
  if (count <= 0) return;
  check_icp_params(prm);
  const int total = im::make_updater(prm).max_iteration;
  if (total > kMaxEnqueuedIterations)
    throw StatusError{PCG_E_INVALID_ARG, "pcg_icp_fit_pairs_dev supports MaxIteration <= 64"};
  // Streams, events and the pinned result buffer live as long as the calling thread: creating them per call costs
  // milliseconds (cudaMallocHost synchronises the device) against ~20 ms of work for 64 pairs.
  FarmResources& res = farm_resources(device, count);
  const int ns = std::min<int>(res.n_streams, count);
  cudaStream_t* streams = res.streams;
  cudaEvent_t start = res.start;
  cudaEvent_t* done = res.done;
  PCG_CUDA(cudaEventRecord(start, stream));
  for (int s = 0; s < ns; s++) PCG_CUDA(cudaStreamWaitEvent(streams[s], start, 0));
  IcpState* h = res.results;
  {
    std::vector<IcpWork> works((size_t)count);
    std::vector<Index*> indices((size_t)count, nullptr);
    try {
      for (int i = 0; i < count; i++) {
        cudaStream_t s = streams[i % ns];
        check_view_args(d_base[i], n_base[i], stride, xyz_off);
        check_view_args(d_target[i], n_target[i], stride, xyz_off);
        CloudView bv = make_view(d_base[i], n_base[i], stride, xyz_off);
        CloudView tv = make_view(d_target[i], n_target[i], stride, xyz_off);
        indices[i] = index_build_device(bv, device, s);
        icp_prepare(*indices[i], tv, prm, false, works[i], s);
        icp_enqueue_iterations(*indices[i], tv, prm, works[i], total, s);
        PCG_CUDA(cudaMemcpyAsync(&h[i], works[i].st.p, sizeof(IcpState), cudaMemcpyDeviceToHost, s));
      }
      for (int s = 0; s < ns; s++) {
        PCG_CUDA(cudaEventRecord(done[s], streams[s]));
        PCG_CUDA(cudaStreamWaitEvent(stream, done[s], 0));
      }
      for (int s = 0; s < ns; s++) PCG_CUDA(cudaStreamSynchronize(streams[s]));
      PCG_CUDA(cudaStreamSynchronize(stream));
    } catch (...) {
      for (int s = 0; s < ns; s++) cudaStreamSynchronize(streams[s]);
      for (auto* ix : indices) index_free_async(ix, stream);
      works.clear();
      throw;
    }
    for (auto* ix : indices) index_free_async(ix, stream);  // every pair's stream has been synchronised
  }
  for (int i = 0; i < count; i++) {
    state_to_outputs(h[i], trans_out ? trans_out + 16 * (size_t)i : nullptr, stat_out ? &stat_out[i] : nullptr);
    if (status_out) status_out[i] = (pcg_status)h[i].status;
  }
}

// One shard's contribution to a single large ICP (see pcg_icp_partial_dev).
void icp_partial_device(const Index& base, const CloudView& tgt, float max_dist, const float trans[16], bool first,
                        const uint32_t* d_order, double* d_partial16, cudaStream_t stream) {
  base.wait(stream);
  IcpWork w;
  pcg_icp_params prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.mode = PCG_ICP_FAST;
  w.nblocks = std::max(1, div_up(tgt.n, term_threads(PCG_ICP_FAST, false)));
  w.n_pad = (tgt.n + 3) & ~(int64_t)3;
  w.st.alloc(1, stream);
  w.partials.alloc((size_t)w.nblocks * kTerms, stream);
  IcpState h = make_state(prm, true);
  std::memcpy(h.trans.m, trans, sizeof(float) * 16);
  h.num_iteration = first ? 0 : 1;  // only "is this the first Evaluate" matters to the terms kernel
  PCG_CUDA(cudaMemcpyAsync(w.st.p, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
  const float mdsq = max_dist * max_dist;
  launch_terms<PCG_ICP_FAST>(base, tgt, d_order, nullptr, mdsq, 0.f, false, w.st.p, w.terms.p, w.n_pad, w.partials.p, nullptr,
                             w.nblocks, 0, stream);
  PCG_LAUNCH(icp_partial_reduce_kernel, 1, kFinishThreads, 0, stream, w.st.p, w.partials.p, w.nblocks, d_partial16);
}

// ---- one large ICP, target sharded over GPUs, loop resident on the device -------------------------------
// Per iteration and rank: terms kernel on the rank's slice (transform read from the device state) -> 16 float64
// -> all-reduce by the caller (NCCL, stream-ordered) -> finish kernel.  No host round trip inside the loop; every
// rank applies the identical update, so the transforms stay bit-identical without a broadcast.
struct IcpShard {
  const Index* base = nullptr;
  CloudView tgt;
  pcg_icp_params prm;
  IcpWork w;
};

IcpShard* icp_shard_new(const Index& base, const CloudView& tgt, const pcg_icp_params& prm_in, cudaStream_t stream) {
  pcg_icp_params prm = prm_in;
  check_icp_params(prm);
  if (prm.updater != PCG_UPDATER_GRADIENT_DESCENT || (prm.mode & PCG_ICP_WITH_HESSIAN))
    throw StatusError{PCG_E_INVALID_ARG, "the sharded loop exchanges the nine Evaluate sums only (gradient-descent updater)"};
  prm.mode = PCG_ICP_FAST;  // a sequential float32 sum has one order: it cannot be sharded
  IcpShard* sh = new IcpShard();
  try {
    sh->base = &base;
    sh->tgt = tgt;
    sh->prm = prm;
    icp_prepare(base, tgt, prm, false, sh->w, stream);
  } catch (...) {
    delete sh;
    throw;
  }
  return sh;
}

void icp_shard_free(IcpShard* sh) { delete sh; }

void icp_shard_partial(IcpShard& sh, double* d_partial16, cudaStream_t stream) {
  const float mdsq = sh.prm.max_dist * sh.prm.max_dist;
  launch_terms<PCG_ICP_FAST>(*sh.base, sh.tgt, sh.w.perm.p, sh.w.warm.p, mdsq, sh.prm.min_dist_sq, false, sh.w.st.p, sh.w.terms.p,
                             sh.w.n_pad, sh.w.partials.p, nullptr, sh.w.nblocks, 0, stream);
  PCG_LAUNCH(icp_partial_reduce_kernel, 1, kFinishThreads, 0, stream, sh.w.st.p, sh.w.partials.p, sh.w.nblocks,
             d_partial16);
}

void icp_shard_finish(IcpShard& sh, const double* d_reduced16, cudaStream_t stream) {
  PCG_LAUNCH(icp_shard_finish_kernel, 1, 32, 0, stream, sh.w.st.p, d_reduced16);
}

pcg_status icp_shard_result(IcpShard& sh, float trans[16], pcg_icp_stat* stat, int32_t* done, cudaStream_t stream) {
  IcpState h;
  PCG_CUDA(cudaMemcpyAsync(&h, sh.w.st.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  state_to_outputs(h, trans, stat);
  if (done) *done = h.done;
  return (pcg_status)h.status;
}

// ---- one large ICP over several GPUs of one process: NVLink peer exchange inside the iteration kernel ---------------
// The same sharding as above (target split, base index replicated) for a caller that owns all devices from ONE process
// (the Go shim: icp.Fit has no notion of ranks).  Every device runs the single-GPU fast loop on its slice - one
// icp_terms_kernel per iteration - and the last CTA of that kernel exchanges the ten sums with the other devices over
// peer-mapped memory before it runs the tail of Evaluate + Update (icp_last_block): no library collective, no extra
// kernel and no host round trip per iteration.  The host enqueues MaxIteration kernels per device (finished ones fall
// through) and joins once.  Exchange buffers, flags and streams are created once per device and reused.
struct PeerResources {
  cudaStream_t stream = nullptr;
  double* slots = nullptr;        // cudaMalloc (not the stream-ordered pool): peer access follows the device
  unsigned int* flags = nullptr;
};
static std::mutex g_peer_mu;           // one multi-GPU Fit at a time: the exchange buffers are per device
static PeerResources g_peer[64];
static unsigned int g_peer_seq = 0;    // exchanges ever enqueued (the flags only grow)
static uint64_t g_peer_enabled = 0;    // bit (a * 8 + b) for small ordinals: peer access a -> b already enabled

static PeerResources& peer_resources(int device) {
  if (device < 0 || device >= 64) throw StatusError{PCG_E_INVALID_ARG, "device ordinal out of range"};
  PeerResources& r = g_peer[device];
  if (!r.stream) {
    PCG_CUDA(cudaSetDevice(device));
    ensure_pool(device);
    PCG_CUDA(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
    PCG_CUDA(cudaMalloc((void**)&r.slots, sizeof(double) * 2 * kMaxPeers * kXchg));
    PCG_CUDA(cudaMalloc((void**)&r.flags, sizeof(unsigned int) * kMaxPeers));
    PCG_CUDA(cudaMemset(r.slots, 0, sizeof(double) * 2 * kMaxPeers * kXchg));
    PCG_CUDA(cudaMemset(r.flags, 0, sizeof(unsigned int) * kMaxPeers));
    PCG_CUDA(cudaDeviceSynchronize());
  }
  return r;
}

// PointToPointICPGradient.Fit with the target sharded over the devices of bases[] (one replica of the base index
// per device), one host thread, no library collective.  d_targets[r] (n_targets[r] records) lives on the device of
// bases[r].  Fast mode (float64 partial sums); the transform is the one the single-GPU fast Fit produces up to the
// rounding of the float64 sums.
pcg_status icp_fit_multi_device(int n_dev, const Index* const* bases, const void* const* d_targets,
                                const int64_t* n_targets, int64_t stride, const int64_t xyz_off[3],
                                const pcg_icp_params& prm_in, float trans[16], pcg_icp_stat* stat) {
  if (n_dev < 1 || n_dev > kMaxPeers) throw StatusError{PCG_E_INVALID_ARG, "1 to 8 devices"};
  pcg_icp_params prm = prm_in;
  check_icp_params(prm);
  if (prm.updater != PCG_UPDATER_GRADIENT_DESCENT || (prm.mode & PCG_ICP_WITH_HESSIAN) || prm.min_dist_sq > 0.f)
    throw StatusError{PCG_E_INVALID_ARG, "the multi-GPU loop is the exact-NN gradient-descent Fit (nine Evaluate sums)"};
  prm.mode = PCG_ICP_FAST;  // a sequential float32 sum has one order: it cannot be sharded
  for (int r = 0; r < n_dev; r++) {
    for (int q = 0; q < r; q++)
      if (bases[q]->device == bases[r]->device) throw StatusError{PCG_E_INVALID_ARG, "one index per device"};
    if (bases[r]->n != bases[0]->n) throw StatusError{PCG_E_INVALID_ARG, "the base replicas differ"};
    check_view_args(d_targets[r], n_targets[r], stride, xyz_off);
  }
  std::lock_guard<std::mutex> lk(g_peer_mu);
  int prev = -1;
  cudaGetDevice(&prev);
  struct Dev {
    IcpWork w;
    CloudView tgt;
  };
  std::vector<Dev> devs((size_t)n_dev);
  auto restore = [&]() {
    for (int r = 0; r < n_dev; r++) {  // the stream-ordered workspace must not outlive its stream's work
      cudaSetDevice(bases[r]->device);
      if (g_peer[bases[r]->device].stream) cudaStreamSynchronize(g_peer[bases[r]->device].stream);
      devs[(size_t)r].w = IcpWork();
    }
    if (prev >= 0) cudaSetDevice(prev);
  };
  try {
    // peer access between every pair (NVLink / NVSwitch), enabled once
    for (int r = 0; r < n_dev && n_dev > 1; r++) {
      const int a = bases[r]->device;
      for (int q = 0; q < n_dev; q++) {
        const int b = bases[q]->device;
        if (q == r) continue;
        const bool small = a < 8 && b < 8;
        if (small && (g_peer_enabled >> (a * 8 + b)) & 1ull) continue;
        PCG_CUDA(cudaSetDevice(a));
        int can = 0;
        PCG_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
        if (!can) throw StatusError{PCG_E_INVALID_ARG, "devices without peer access"};
        const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PCG_CUDA(e);
        cudaGetLastError();
        if (small) g_peer_enabled |= 1ull << (a * 8 + b);
      }
    }
    PeerCtx pc;
    std::memset(&pc, 0, sizeof(pc));
    pc.n_dev = n_dev;
    pc.seq_base = g_peer_seq;
    for (int r = 0; r < n_dev; r++) {
      PeerResources& pr = peer_resources(bases[r]->device);
      pc.slots[r] = pr.slots;
      pc.flags[r] = pr.flags;
    }
    const int max_it = im::make_updater(prm).max_iteration;
    g_peer_seq += (unsigned int)max_it;
    const float mdsq = prm.max_dist * prm.max_dist;  // kdtree.go:91
    // One host thread per device prepares its slice (visit order, workspace) and enqueues all iterations: launching
    // 8 x 20 kernels from a single thread costs more than the kernels run (a device spins on its peers' flags, so
    // every device must get its kernels early).
    std::vector<std::exception_ptr> errors((size_t)n_dev);
    auto work = [&](int r) {
      try {
        Dev& d = devs[(size_t)r];
        PCG_CUDA(cudaSetDevice(bases[r]->device));
        cudaStream_t stream = g_peer[bases[r]->device].stream;
        d.tgt = make_view(d_targets[r], n_targets[r], stride, xyz_off);
        icp_prepare(*bases[r], d.tgt, prm, false, d.w, stream);
        PeerCtx mine = pc;
        mine.rank = r;
        for (int it = 0; it < max_it; it++)
          launch_terms<PCG_ICP_FAST>(*bases[r], d.tgt, d.w.perm.p, d.w.warm.p, mdsq, 0.f, false, d.w.st.p, d.w.terms.p,
                                     d.w.n_pad, d.w.partials.p, d.w.hpartials.p, d.w.nblocks, 1, stream, &mine);
      } catch (...) {
        errors[(size_t)r] = std::current_exception();
      }
    };
    {
      std::vector<std::thread> threads;
      for (int r = 1; r < n_dev; r++) threads.emplace_back(work, r);
      work(0);
      for (auto& t : threads) t.join();
    }
    for (int r = 0; r < n_dev; r++) {
      if (!errors[(size_t)r]) continue;
      // A device that failed to enqueue its iterations would leave the others spinning on its flag for ever: raise
      // that flag on every device from the host (the streams are non-blocking, a plain copy runs beside them); the
      // sums they then read are meaningless, and so is the result - the call fails.
      const unsigned int last = pc.seq_base + (unsigned int)max_it;
      for (int q = 0; q < n_dev; q++) {
        cudaSetDevice(bases[q]->device);
        cudaMemcpy(pc.flags[q] + r, &last, sizeof(last), cudaMemcpyHostToDevice);
      }
    }
    for (int r = 0; r < n_dev; r++)
      if (errors[(size_t)r]) std::rethrow_exception(errors[(size_t)r]);
    IcpState h;
    PCG_CUDA(cudaSetDevice(bases[0]->device));
    PCG_CUDA(cudaMemcpyAsync(&h, devs[0].w.st.p, sizeof(h), cudaMemcpyDeviceToHost, g_peer[bases[0]->device].stream));
    for (int r = 0; r < n_dev; r++) {
      PCG_CUDA(cudaSetDevice(bases[r]->device));
      PCG_CUDA(cudaStreamSynchronize(g_peer[bases[r]->device].stream));
    }
    state_to_outputs(h, trans, stat);
    restore();
    return (pcg_status)h.status;
  } catch (...) {
    restore();
    throw;
  }
}

// Tail of Evaluate + Update on the host from all-reduced sums (see pcg_icp_finish).
pcg_status icp_finish_host(const double partial16[16], const pcg_icp_params& prm, int32_t* iter, float trans[16],
                           pcg_evaluated* ev_out, int32_t* converged) {
  const long long n_pairs = (long long)partial16[9];
  const int min_pairs = prm.min_pairs == 0 ? 6 : prm.min_pairs;
  if (n_pairs < min_pairs) return PCG_E_NOT_ENOUGH_PAIRS;
  im::Sums s;
  s.value = (float)partial16[0];
  s.sum_weight = (float)partial16[1];
  for (int k = 0; k < 6; k++) s.g[k] = (float)partial16[2 + k];
  s.rms = (float)partial16[8];
  im::Eval ev = im::evaluate_tail(s);
  if (ev_out) {
    std::memset(ev_out, 0, sizeof(*ev_out));
    ev_out->value = ev.value;
    for (int k = 0; k < 6; k++) ev_out->gradient[k] = ev.g[k];
    ev_out->dist_rms = ev.dist_rms;
  }
  im::UpdaterCfg cfg = im::make_updater(prm);
  im::M4 t;
  std::memcpy(t.m, trans, sizeof(t.m));
  int i = *iter;
  bool c = im::updater_update(cfg, &i, &t, ev);
  std::memcpy(trans, t.m, sizeof(t.m));
  *iter = i;
  *converged = c ? 1 : 0;
  return PCG_OK;
}

// Test hook: the reference-order (sequential) float32 sum of a device array through either replay kernel.
float debug_sequential_sum(const float* d_x, int64_t n, bool exact_path, cudaStream_t stream, float* stats3) {
  const int64_t n_pad = (n + 3) & ~(int64_t)3;
  DevBuf<IcpState> st(1, stream);
  PCG_CUDA(cudaMemsetAsync(st.p, 0, sizeof(IcpState), stream));
  const size_t nchunks = (size_t)std::max<int64_t>(1, (n + kReplayChunk - 1) / kReplayChunk);
  DevBuf<ReplayChunk> chunks(nchunks * kReplayChunkRecords, stream);
  DevBuf<double> sums(nchunks, stream);
  if (exact_path)
    launch_exact_replay(st.p, d_x, n, n_pad, 1, chunks.p, sums.p, stream);
  else
    PCG_LAUNCH(icp_replay_kernel, 1, 32, 0, stream, st.p, d_x, n, n_pad);
  IcpState h;
  PCG_CUDA(cudaMemcpyAsync(&h, st.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  if (stats3) {
    stats3[0] = h.sums[9];
    stats3[1] = h.sums[10];
    stats3[2] = h.sums[11];
  }
  return h.sums[0];
}

// NearestPointCorresponder.Pairs building block: nearest neighbour of every target point.
template <bool APPROX>
__global__ void __launch_bounds__(128)
    icp_pairs_kernel(IndexView base, CloudView tgt, float max_dist_sq, float min_dist_sq, int32_t* __restrict__ ids,
                     float* __restrict__ dsq) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tgt.n) return;
  float3 p = load_xyz(tgt, i);
  uint64_t best = nn_init(max_dist_sq);
  const uint64_t init = best;
  uint32_t pos = 0;
  nn_traverse4<APPROX>(base, p.x, p.y, p.z, best, pos, min_dist_sq);
  const bool hit = best != init;
  ids[i] = hit ? (int32_t)(uint32_t)best : -1;
  dsq[i] = hit ? __uint_as_float((uint32_t)(best >> 32)) : max_dist_sq;
}

void icp_pairs_device(const Index& base, const CloudView& tgt, float max_dist, float min_dist_sq, int32_t* d_ids,
                      float* d_dsq, cudaStream_t stream) {
  base.wait(stream);
  if (tgt.n == 0) return;
  if (min_dist_sq > 0.f)  // threshold capped at maxDist^2: see nearest_device
    PCG_LAUNCH(icp_pairs_kernel<true>, div_up(tgt.n, 128), 128, 0, stream, base.view(), tgt, max_dist * max_dist,
               fminf(min_dist_sq, max_dist * max_dist), d_ids, d_dsq);
  else
    PCG_LAUNCH(icp_pairs_kernel<false>, div_up(tgt.n, 128), 128, 0, stream, base.view(), tgt, max_dist * max_dist, 0.f,
               d_ids, d_dsq);
}

}  // namespace pcg
