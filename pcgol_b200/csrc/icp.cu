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
//   STRICT mode        the terms are stored at the target's index and icp_replay_kernel (nine
//                      single-warp CTAs, one accumulator each, on nine SMs) replays the
//                      reference's sequential float32 accumulation: bit-identical trajectory.
//   tail               evaluator.go:156-186 + gradientDescentUpdater.Update (updater.go:44-71)
//                      run on the device, so the loop never returns to the host; once `done`
//                      is set the remaining launches fall through.
#include <algorithm>
#include <vector>

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
};

constexpr int kTermThreads = 128;
constexpr int kTerms = 9;  // Value, SumW, G0..G5, R

__device__ __forceinline__ void icp_fast_last_block(IcpState* __restrict__ st, const double* __restrict__ partials,
                                                    int nblocks, float* s_sum, int* s_last);

template <int MODE>
__global__ void __launch_bounds__(kTermThreads)
    icp_terms_kernel(IndexView base, CloudView tgt, const uint32_t* __restrict__ perm, float max_dist_sq,
                     IcpState* __restrict__ st, float* __restrict__ terms, int64_t n_pad,
                     double* __restrict__ partials, int finalize) {
  if (st->done) return;
  __shared__ float s_m[16];
  __shared__ int s_first;
  __shared__ int s_last;
  __shared__ float s_sum[kTerms];
  __shared__ double s_red[kTermThreads / 32][kTerms];
  const int tid = threadIdx.x;
  if (tid < 16) s_m[tid] = st->trans.m[tid];
  if (tid == 0) s_first = st->num_iteration == 0;
  __syncthreads();
  const int64_t slot = (int64_t)blockIdx.x * kTermThreads + tid;
  float t[kTerms];
#pragma unroll
  for (int k = 0; k < kTerms; k++) t[k] = 0.f;
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
    PCG_NN_TRAVERSE(base, x0, y0, z0, best, pos);
    if (best != init) {  // correspondence.go:27-29
      matched = 1;
      const float4 pb = __ldg(base.pts + pos);
      const float x1 = pb.x, y1 = pb.y, z1 = pb.z;
      const float dsq = __uint_as_float((uint32_t)(best >> 32));
      // evaluator.go:130-144 with w = 1 (DefaultEvaluateWeightFn): w*v == v exactly
      t[0] = dsq;
      t[1] = 1.f;
      t[2] = im::sub(x0, x1);
      t[3] = im::sub(y0, y1);
      t[4] = im::sub(z0, z1);
      t[5] = im::sub(im::mul(z0, y1), im::mul(y0, z1));
      t[6] = im::sub(im::mul(x0, z1), im::mul(z0, x1));
      t[7] = im::sub(im::mul(y0, x1), im::mul(x0, y1));
      t[8] = im::add(im::add(im::mul(x0, x0), im::mul(y0, y0)), im::mul(z0, z0));
    }
    if (MODE == PCG_ICP_STRICT) {
      // unmatched targets contribute +0: x + (+0) == x for every partial sum the reference can hold
#pragma unroll
      for (int k = 0; k < kTerms; k++) terms[(int64_t)k * n_pad + i] = t[k];
    }
  }
  const int block_pairs = __syncthreads_count(matched);
  if (tid == 0 && block_pairs) atomicAdd(&st->pair_counter, (unsigned int)block_pairs);
  if (MODE == PCG_ICP_FAST) {
    double d[kTerms];
#pragma unroll
    for (int k = 0; k < kTerms; k++) {
      d[k] = (double)t[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d[k] += __shfl_down_sync(0xffffffffu, d[k], o);
    }
    if ((tid & 31) == 0) {
#pragma unroll
      for (int k = 0; k < kTerms; k++) s_red[tid >> 5][k] = d[k];
    }
    __syncthreads();
    if (tid < kTerms) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kTermThreads / 32; w++) s += s_red[w][tid];
      partials[(int64_t)blockIdx.x * kTerms + tid] = s;
    }
    if (finalize) icp_fast_last_block(st, partials, (int)gridDim.x, s_sum, &s_last);
  }
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
  if (st->evaluate_only) {
    st->done = 1;
    return;
  }
  im::M4 trans = st->trans;
  int iter = st->iter;
  const bool converged = im::updater_update(st->cfg, &iter, &trans, ev);  // icp.go:57
  st->trans = trans;
  st->iter = iter;
  if (converged) st->done = 1;
}

// FAST mode: called by every CTA of icp_terms_kernel after it wrote its partials; the last one
// to arrive reduces all of them (fixed order: lane-strided columns, then a shuffle tree).
__device__ __forceinline__ void icp_fast_last_block(IcpState* __restrict__ st, const double* __restrict__ partials,
                                                    int nblocks, float* s_sum, int* s_last) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __threadfence();
  if (tid == 0) {
    const unsigned int t = atomicAdd(&st->ticket, 1u);
    *s_last = (t == (unsigned int)nblocks - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  for (int k = warp; k < kTerms; k += kTermThreads / 32) {
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += __ldcg(&partials[(int64_t)b * kTerms + k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) s_sum[k] = (float)s;
  }
  __syncthreads();
  if (tid == 0) {
    st->ticket = 0;
    icp_finalize(st, s_sum);
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

// Sharded ICP: fold the per-CTA float64 partials into 16 doubles for the all-reduce.
__global__ void __launch_bounds__(kFinishThreads)
    icp_partial_reduce_kernel(IcpState* __restrict__ st, const double* __restrict__ partials, int nblocks,
                              double* __restrict__ out16) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(int64_t)b * kTerms + warp];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out16[warp] = s;
  if (tid == 0) {
    out16[9] = (double)st->pair_counter;
    for (int k = 10; k < 16; k++) out16[k] = 0.0;
  }
}

struct IcpWork {
  DevBuf<IcpState> st;
  DevBuf<float> terms;
  DevBuf<double> partials;
  DevBuf<uint32_t> perm;
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
  return h;
}

// Enqueues a whole Fit (or a single Evaluate) on `stream` without synchronising.
// Returns false if max_iteration exceeds what is enqueued at once (caller then polls).
constexpr int kMaxEnqueuedIterations = 64;

static void icp_enqueue_iterations(const Index& base, const CloudView& tgt, float max_dist, int mode, IcpWork& w,
                                   int iterations, cudaStream_t stream) {
  const float mdsq = max_dist * max_dist;  // kdtree.go:91
  for (int it = 0; it < iterations; it++) {
    if (mode == PCG_ICP_STRICT) {
      PCG_LAUNCH((icp_terms_kernel<PCG_ICP_STRICT>), w.nblocks, kTermThreads, 0, stream, base.view(), tgt, w.perm.p, mdsq,
                 w.st.p, w.terms.p, w.n_pad, w.partials.p, 0);
      PCG_LAUNCH(icp_replay_kernel, kTerms, 32, 0, stream, w.st.p, w.terms.p, tgt.n, w.n_pad);
    } else {
      PCG_LAUNCH((icp_terms_kernel<PCG_ICP_FAST>), w.nblocks, kTermThreads, 0, stream, base.view(), tgt, w.perm.p, mdsq,
                 w.st.p, w.terms.p, w.n_pad, w.partials.p, 1);
    }
  }
}

static void icp_prepare(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, bool evaluate_only,
                        IcpWork& w, cudaStream_t stream) {
  w.nblocks = std::max(1, div_up(tgt.n, kTermThreads));
  if (tgt.n >= kMinQueriesToReorder && base.n > 0) {
    // the target moves by a small rigid transform per iteration: the order of the raw target stays coherent
    w.perm.alloc((size_t)tgt.n, stream);
    query_order_device(base, tgt, w.perm.p, stream);
  }
  w.n_pad = (tgt.n + 3) & ~(int64_t)3;
  w.st.alloc(1, stream);
  if (prm.mode == PCG_ICP_STRICT)
    w.terms.alloc((size_t)std::max<int64_t>(4, w.n_pad) * kTerms, stream);
  else
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
    stat->num_iteration = h.num_iteration;
    stat->n_pairs = h.n_pairs;
  }
}

// PointToPointICPGradient.Fit (icp.go:23-67) / Evaluate only. Synchronises `stream`.
pcg_status icp_fit_device(const Index& base, const CloudView& tgt, const pcg_icp_params& prm, bool evaluate_only,
                          float trans[16], pcg_icp_stat* stat, cudaStream_t stream) {
  if (prm.mode != PCG_ICP_STRICT && prm.mode != PCG_ICP_FAST)
    throw StatusError{PCG_E_INVALID_ARG, "unknown ICP mode"};
  IcpWork w;
  icp_prepare(base, tgt, prm, evaluate_only, w, stream);
  const int total = evaluate_only ? 1 : im::make_updater(prm).max_iteration;
  IcpState h;
  int enq = 0;
  for (;;) {
    // every Update that does not converge increments i, so `total` Evaluate calls always suffice
    const int batch = std::min(kMaxEnqueuedIterations, std::max(1, total - enq));
    icp_enqueue_iterations(base, tgt, prm.max_dist, prm.mode, w, batch, stream);
    enq += batch;
    PCG_CUDA(cudaMemcpyAsync(&h, w.st.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PCG_CUDA(cudaStreamSynchronize(stream));
    if (h.done) break;
  }
  state_to_outputs(h, trans, stat);
  return (pcg_status)h.status;
}

// Scan-pair farm (BASELINE config 4): independent pairs round-robined over a few
// streams so that index builds and fits of different pairs overlap.
void icp_fit_pairs_device(int32_t count, const void* const* d_base, const int64_t* n_base,
                          const void* const* d_target, const int64_t* n_target, int64_t stride,
                          const int64_t xyz_off[3], const pcg_icp_params& prm, int device, float* trans_out,
                          pcg_icp_stat* stat_out, pcg_status* status_out, cudaStream_t stream) {
  if (count <= 0) return;
  if (prm.mode != PCG_ICP_STRICT && prm.mode != PCG_ICP_FAST)
    throw StatusError{PCG_E_INVALID_ARG, "unknown ICP mode"};
  const int total = im::make_updater(prm).max_iteration;
  if (total > kMaxEnqueuedIterations)
    throw StatusError{PCG_E_INVALID_ARG, "pcg_icp_fit_pairs_dev supports MaxIteration <= 64"};
  constexpr int kStreams = 8;
  const int ns = std::min<int>(kStreams, count);
  cudaStream_t streams[kStreams];
  cudaEvent_t start, done[kStreams];
  for (int s = 0; s < ns; s++) PCG_CUDA(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking));
  PCG_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
  for (int s = 0; s < ns; s++) PCG_CUDA(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
  PCG_CUDA(cudaEventRecord(start, stream));
  for (int s = 0; s < ns; s++) PCG_CUDA(cudaStreamWaitEvent(streams[s], start, 0));
  PinnedBuf results((size_t)count * sizeof(IcpState));
  IcpState* h = (IcpState*)results.p;
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
        icp_enqueue_iterations(*indices[i], tv, prm.max_dist, prm.mode, works[i], total, s);
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
      for (auto* ix : indices) index_free(ix);
      works.clear();
      for (int s = 0; s < ns; s++) cudaStreamDestroy(streams[s]);
      throw;
    }
    for (auto* ix : indices) index_free(ix);
  }
  for (int i = 0; i < count; i++) {
    state_to_outputs(h[i], trans_out ? trans_out + 16 * (size_t)i : nullptr, stat_out ? &stat_out[i] : nullptr);
    if (status_out) status_out[i] = (pcg_status)h[i].status;
  }
  cudaEventDestroy(start);
  for (int s = 0; s < ns; s++) {
    cudaEventDestroy(done[s]);
    cudaStreamDestroy(streams[s]);
  }
}

// One shard's contribution to a single large ICP (see pcg_icp_partial_dev).
void icp_partial_device(const Index& base, const CloudView& tgt, float max_dist, const float trans[16], bool first,
                        const uint32_t* d_order, double* d_partial16, cudaStream_t stream) {
  IcpWork w;
  pcg_icp_params prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.mode = PCG_ICP_FAST;
  w.nblocks = std::max(1, div_up(tgt.n, kTermThreads));
  w.n_pad = (tgt.n + 3) & ~(int64_t)3;
  w.st.alloc(1, stream);
  w.partials.alloc((size_t)w.nblocks * kTerms, stream);
  IcpState h = make_state(prm, true);
  std::memcpy(h.trans.m, trans, sizeof(float) * 16);
  h.num_iteration = first ? 0 : 1;  // only "is this the first Evaluate" matters to the terms kernel
  PCG_CUDA(cudaMemcpyAsync(w.st.p, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
  const float mdsq = max_dist * max_dist;
  PCG_LAUNCH((icp_terms_kernel<PCG_ICP_FAST>), w.nblocks, kTermThreads, 0, stream, base.view(), tgt, d_order, mdsq,
             w.st.p, w.terms.p, w.n_pad, w.partials.p, 0);
  PCG_LAUNCH(icp_partial_reduce_kernel, 1, kFinishThreads, 0, stream, w.st.p, w.partials.p, w.nblocks, d_partial16);
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

// NearestPointCorresponder.Pairs building block: nearest neighbour of every target point.
__global__ void __launch_bounds__(128)
    icp_pairs_kernel(IndexView base, CloudView tgt, float max_dist_sq, int32_t* __restrict__ ids,
                     float* __restrict__ dsq) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tgt.n) return;
  float3 p = load_xyz(tgt, i);
  uint64_t best = nn_init(max_dist_sq);
  const uint64_t init = best;
  uint32_t pos = 0;
  PCG_NN_TRAVERSE(base, p.x, p.y, p.z, best, pos);
  const bool hit = best != init;
  ids[i] = hit ? (int32_t)(uint32_t)best : -1;
  dsq[i] = hit ? __uint_as_float((uint32_t)(best >> 32)) : max_dist_sq;
}

void icp_pairs_device(const Index& base, const CloudView& tgt, float max_dist, int32_t* d_ids, float* d_dsq,
                      cudaStream_t stream) {
  if (tgt.n == 0) return;
  PCG_LAUNCH(icp_pairs_kernel, div_up(tgt.n, 128), 128, 0, stream, base.view(), tgt, max_dist * max_dist, d_ids,
             d_dsq);
}

}  // namespace pcg
