// radix_sort.cuh — stable LSD radix sort of (key, uint32 payload) pairs, hand-written
// for sm_100a.  One upfront kernel histograms every digit position; each 8-bit pass is
// then a single "onesweep" kernel: tiles are ranked in shared memory with warp
// match-any multi-split, the per-tile digit counts are chained between CTAs by
// decoupled look-back (one packed 64-bit status word per tile and digit, so no fence
// is needed between flag and value), and keys are staged in shared memory so that the
// scatter writes coalesced runs.  Stability (equal keys keep input order) is what makes
// the VoxelGrid centroid sums reproduce the reference's accumulation order.
#pragma once

#include <cstdlib>

#include "common.cuh"

namespace pcg {
namespace rsort {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kThreads = 256;  // one thread per digit for the look-back
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPasses = 8;

constexpr uint64_t kValueMask = (1ull << 54) - 1;
constexpr int kEpochShift = 54;
constexpr int kStateShift = 62;
constexpr uint64_t kStateAggregate = 1ull;
constexpr uint64_t kStateInclusive = 2ull;

template <typename K>
__device__ __forceinline__ uint32_t digit_of(K k, int shift) {
  return (uint32_t)(k >> shift) & (kRadix - 1);
}

// Warp-aggregated accumulation of one key into a CTA-shared histogram [passes][256]:
// lanes holding the same digit elect one leader that adds their count (clustered keys
// would otherwise serialise on one shared-memory address).  All 32 lanes must call.
template <typename K>
__device__ __forceinline__ void hist_add_key(uint32_t* __restrict__ sh, K key, bool valid, int begin_bit,
                                             int passes) {
  const uint32_t lane = threadIdx.x & 31;
  for (int p = 0; p < passes; p++) {
    const uint32_t d = valid ? digit_of(key, begin_bit + p * kRadixBits) : 0xffffffffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (valid && (int)lane == __ffs(peers) - 1) atomicAdd(&sh[p * kRadix + d], (uint32_t)__popc(peers));
  }
}
__device__ __forceinline__ void hist_zero(uint32_t* sh, int passes) {
  for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) sh[i] = 0;
}
__device__ __forceinline__ void hist_flush(const uint32_t* sh, uint32_t* __restrict__ hist, int passes) {
  for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
    uint32_t c = sh[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// All digit histograms in one read of the keys (the per-position histogram of a
// multiset does not depend on its order, so it is valid for every later pass).
// Producers of keys can fold this into their own kernel with hist_add_key instead.
template <typename K>
__global__ void __launch_bounds__(256) histogram_kernel(const K* __restrict__ keys, uint32_t n, int begin_bit,
                                                        int passes, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kMaxPasses * kRadix];
  hist_zero(sh, passes);
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (n + stride - 1) / stride;
  for (uint32_t r = 0; r < rounds; r++) {
    uint32_t i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    K k = valid ? keys[i] : (K)0;
    hist_add_key(sh, k, valid, begin_bit, passes);
  }
  __syncthreads();
  hist_flush(sh, hist, passes);
}

// Exclusive scan of one value per thread across a 256-thread block.
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp /*[kWarps]*/, uint32_t* total) {
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

template <typename K, int IPT>
struct TileSmem {
  static constexpr int kTile = kThreads * IPT;
  static constexpr size_t kBytes = (size_t)kTile * (sizeof(K) + sizeof(uint32_t));
};

// Up to kMaxBatch independent sorts of the same length advance together: blockIdx.y selects the sort.  (Three
// axis sorts of the KD build in one launch per digit keep the machine full where one 1M-key pass does not.)
constexpr int kMaxBatch = 3;
template <typename K>
struct PassArgs {
  const K* keys_in[kMaxBatch];
  K* keys_out[kMaxBatch];
  const uint32_t* vals_in[kMaxBatch];
  uint32_t* vals_out[kMaxBatch];
  const uint32_t* hist[kMaxBatch];
  uint32_t* tile_counter[kMaxBatch];
  unsigned long long* status[kMaxBatch];
};

template <typename K, int IPT>
__global__ void __launch_bounds__(kThreads)
    onesweep_kernel(PassArgs<K> A, uint32_t n, int shift, uint32_t epoch) {
  const K* __restrict__ keys_in = A.keys_in[blockIdx.y];
  K* __restrict__ keys_out = A.keys_out[blockIdx.y];
  const uint32_t* __restrict__ vals_in = A.vals_in[blockIdx.y];
  uint32_t* __restrict__ vals_out = A.vals_out[blockIdx.y];
  const uint32_t* __restrict__ hist = A.hist[blockIdx.y];
  uint32_t* __restrict__ tile_counter = A.tile_counter[blockIdx.y];
  unsigned long long* __restrict__ status = A.status[blockIdx.y];
  constexpr int kTile = kThreads * IPT;
  __shared__ uint32_t s_warp_hist[kWarps][kRadix];
  __shared__ uint32_t s_digit_start[kRadix];
  __shared__ uint32_t s_global_base[kRadix];
  __shared__ uint32_t s_scan[kWarps];
  __shared__ uint32_t s_tile;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  K* s_keys = reinterpret_cast<K*>(s_dyn);
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_dyn + (size_t)kTile * sizeof(K));

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);  // dynamic tile id: predecessors are always running
#pragma unroll
  for (int w = 0; w < kWarps; w++) s_warp_hist[w][tid] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_base = tile * (uint32_t)kTile;
  const uint32_t warp_base = tile_base + warp * (32u * IPT);

  K keys[IPT];
  uint32_t vals[IPT];
  uint32_t offs[IPT];
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    uint32_t idx = warp_base + i * 32 + lane;
    bool valid = idx < n;
    keys[i] = valid ? keys_in[idx] : (K)~(K)0;
    vals[i] = vals_in ? (valid ? vals_in[idx] : 0u) : idx;
  }
  // Rank inside the warp: items are visited in index order (i major, lane minor).
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    uint32_t idx = warp_base + i * 32 + lane;
    bool valid = idx < n;
    uint32_t d = valid ? digit_of(keys[i], shift) : (uint32_t)kRadix;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pre = 0;
    // atomics of one warp on one address execute in program order: no barrier between the rounds
    if (valid && (int)lane == leader) pre = atomicAdd(&s_warp_hist[warp][d], (uint32_t)__popc(peers));
    pre = __shfl_sync(0xffffffffu, pre, leader);
    offs[i] = pre + rank;
  }
  __syncthreads();

  // Thread t owns digit t: scan the warp histograms (warp order == index order).
  uint32_t count = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) {
    uint32_t c = s_warp_hist[w][tid];
    s_warp_hist[w][tid] = count;
    count += c;
  }

  // Decoupled look-back over predecessor tiles for this digit.
  {
    volatile unsigned long long* st = status;
    const uint64_t tag = (uint64_t)epoch << kEpochShift;
    uint64_t excl = 0;
    const size_t mine = (size_t)tile * kRadix + tid;
    if (tile == 0) {
      st[mine] = (kStateInclusive << kStateShift) | tag | (uint64_t)count;
    } else {
      st[mine] = (kStateAggregate << kStateShift) | tag | (uint64_t)count;
      int64_t prev = (int64_t)tile - 1;
      for (;;) {
        uint64_t w = st[(size_t)prev * kRadix + tid];
        uint64_t state = w >> kStateShift;
        if (state == 0 || ((w >> kEpochShift) & 0xffu) != epoch) continue;  // not published yet for this pass
        excl += w & kValueMask;
        if (state == kStateInclusive) break;
        prev--;
      }
      st[mine] = (kStateInclusive << kStateShift) | tag | (excl + count);
    }
    uint32_t digit_base = block_excl_scan_256(hist[tid], s_scan, nullptr);
    s_global_base[tid] = digit_base + (uint32_t)excl;
  }
  s_digit_start[tid] = block_excl_scan_256(count, s_scan, nullptr);
  __syncthreads();

  // Stage the tile in digit order, then write coalesced runs.
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    uint32_t idx = warp_base + i * 32 + lane;
    if (idx < n) {
      uint32_t d = digit_of(keys[i], shift);
      uint32_t pos = s_digit_start[d] + s_warp_hist[warp][d] + offs[i];
      s_keys[pos] = keys[i];
      s_vals[pos] = vals[i];
    }
  }
  __syncthreads();
  const uint32_t tile_count = min((uint32_t)kTile, n - tile_base);
  for (uint32_t s = tid; s < tile_count; s += kThreads) {
    K k = s_keys[s];
    uint32_t d = digit_of(k, shift);
    uint32_t dst = s_global_base[d] + (s - s_digit_start[d]);
    if (keys_out) keys_out[dst] = k;
    vals_out[dst] = s_vals[s];
  }
}

inline int num_passes(int begin_bit, int end_bit) {
  int bits = end_bit - begin_bit;
  return bits <= 0 ? 0 : (bits + kRadixBits - 1) / kRadixBits;
}

template <typename K, int IPT>
inline void launch_pass_batch(const PassArgs<K>& args, int batch, uint32_t n, int shift, uint32_t epoch,
                              cudaStream_t stream) {
  constexpr size_t smem = TileSmem<K, IPT>::kBytes;
  static std::atomic<uint64_t> configured{0};  // bit per device; the attribute is per device
  int dev = 0;
  PCG_CUDA(cudaGetDevice(&dev));
  if (!(configured.load(std::memory_order_relaxed) & (1ull << dev))) {
    PCG_CUDA(cudaFuncSetAttribute(onesweep_kernel<K, IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  uint32_t tiles = (n + TileSmem<K, IPT>::kTile - 1) / TileSmem<K, IPT>::kTile;
  PCG_LAUNCH((onesweep_kernel<K, IPT>), dim3(tiles, (unsigned)batch), kThreads, smem, stream, args, n, shift, epoch);
}

template <typename K, int IPT>
inline void launch_pass(const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, uint32_t n, int shift,
                        uint32_t epoch, const uint32_t* hist, uint32_t* counter, unsigned long long* status,
                        cudaStream_t stream) {
  PassArgs<K> a;
  std::memset(&a, 0, sizeof(a));
  a.keys_in[0] = kin;
  a.keys_out[0] = kout;
  a.vals_in[0] = vin;
  a.vals_out[0] = vout;
  a.hist[0] = hist;
  a.tile_counter[0] = counter;
  a.status[0] = status;
  launch_pass_batch<K, IPT>(a, 1, n, shift, epoch, stream);
}

inline int pick_ipt(uint32_t n) {
  // largest tile that still gives every SM about four CTAs
  for (int ipt : {16, 8}) {
    if ((uint64_t)n >= (uint64_t)kNumSMs * 4 * kThreads * ipt) return ipt;
  }
  return 4;
}

// Stable LSD sort of n (key, payload) pairs by key bits [begin_bit, end_bit).
//   prepare()    allocates and clears the workspace (histograms, tile counters, look-back status)
//   hist()       device histogram [passes][256]; either call histogram() or let the kernel that
//                produces the keys accumulate into it (hist_add_key / hist_flush)
//   run()        one onesweep kernel per 8-bit digit; keys[0]/vals[0] hold the input, the buffers
//                ping-pong and *result (0 or 1) tells which side holds the output.
//                identity_vals: payload = input position (vals[0] is not read);
//                keep_keys=false skips the key write of the last pass.
template <typename K>
struct Sorter {
  uint32_t n = 0;
  int begin_bit = 0, passes = 0, ipt = 16;
  uint32_t tiles = 0;
  size_t hist_words = 0;
  DevBuf<uint32_t> head;
  DevBuf<unsigned long long> status;

  void prepare(uint32_t n_, int begin_bit_, int end_bit_, cudaStream_t stream) {
    n = n_;
    begin_bit = begin_bit_;
    passes = num_passes(begin_bit_, end_bit_);
    if (passes == 0) passes = 1;  // degenerate range: one pass over (all-equal) digits keeps the order
    if (passes > kMaxPasses) throw StatusError{PCG_E_INVALID_ARG, "radix sort: more than 64 key bits"};
    if (n == 0) return;
    ipt = pick_ipt(n);
    const uint32_t tile = (uint32_t)kThreads * ipt;
    tiles = (n + tile - 1) / tile;
    hist_words = (size_t)passes * kRadix;
    head.alloc(hist_words + 64, stream);
    status.alloc((size_t)tiles * kRadix, stream);
    PCG_CUDA(cudaMemsetAsync(head.p, 0, head.bytes(), stream));
    PCG_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), stream));
  }
  uint32_t* hist() { return head.p; }

  void histogram(const K* keys, cudaStream_t stream) {
    if (n == 0) return;
    int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(n, 256));
    PCG_LAUNCH((histogram_kernel<K>), blocks, 256, 0, stream, keys, n, begin_bit, passes, head.p);
  }

  void run(K* keys[2], uint32_t* vals[2], bool identity_vals, bool keep_keys, cudaStream_t stream, int* result) {
    *result = 0;
    if (n == 0) return;
    int cur = 0;
    for (int p = 0; p < passes; p++) {
      const int shift = begin_bit + p * kRadixBits;
      const bool last = p == passes - 1;
      const K* kin = keys[cur];
      K* kout = (last && !keep_keys) ? nullptr : keys[cur ^ 1];
      const uint32_t* vin = (p == 0 && identity_vals) ? nullptr : vals[cur];
      uint32_t* vout = vals[cur ^ 1];
      const uint32_t* h = head.p + (size_t)p * kRadix;
      uint32_t* counter = head.p + hist_words + p;
      const uint32_t epoch = (uint32_t)(p + 1);
      if (ipt == 4)
        launch_pass<K, 4>(kin, kout, vin, vout, n, shift, epoch, h, counter, status.p, stream);
      else if (ipt == 8)
        launch_pass<K, 8>(kin, kout, vin, vout, n, shift, epoch, h, counter, status.p, stream);
      else
        launch_pass<K, 16>(kin, kout, vin, vout, n, shift, epoch, h, counter, status.p, stream);
      cur ^= 1;
    }
    *result = cur;
  }
};

// `batch` prepared sorters of identical length and bit range advance pass by pass in shared launches.
// keys[b][2] / vals[b][2] as in Sorter::run; every sort ends on the same side (*result).
template <typename K>
void run_batched(Sorter<K>* sorters, int batch, K* (*keys)[2], uint32_t* (*vals)[2], bool identity_vals,
                 bool keep_keys, cudaStream_t stream, int* result) {
  *result = 0;
  if (batch <= 0 || sorters[0].n == 0) return;
  if (batch > kMaxBatch) throw StatusError{PCG_E_INVALID_ARG, "radix sort: batch too large"};
  const Sorter<K>& s0 = sorters[0];
  int cur = 0;
  for (int p = 0; p < s0.passes; p++) {
    const bool last = p == s0.passes - 1;
    PassArgs<K> a;
    std::memset(&a, 0, sizeof(a));
    for (int b = 0; b < batch; b++) {
      a.keys_in[b] = keys[b][cur];
      a.keys_out[b] = (last && !keep_keys) ? nullptr : keys[b][cur ^ 1];
      a.vals_in[b] = (p == 0 && identity_vals) ? nullptr : vals[b][cur];
      a.vals_out[b] = vals[b][cur ^ 1];
      a.hist[b] = sorters[b].head.p + (size_t)p * kRadix;
      a.tile_counter[b] = sorters[b].head.p + sorters[b].hist_words + p;
      a.status[b] = sorters[b].status.p;
    }
    const int shift = s0.begin_bit + p * kRadixBits;
    const uint32_t epoch = (uint32_t)(p + 1);
    // tile size for the work of the whole batch (the workspace was sized for the smaller tiles of one sort)
    const int ipt = std::max(s0.ipt, pick_ipt((uint32_t)std::min<uint64_t>((uint64_t)s0.n * (uint64_t)batch, 0xffffffffull)));
    if (ipt == 4)
      launch_pass_batch<K, 4>(a, batch, s0.n, shift, epoch, stream);
    else if (ipt == 8)
      launch_pass_batch<K, 8>(a, batch, s0.n, shift, epoch, stream);
    else
      launch_pass_batch<K, 16>(a, batch, s0.n, shift, epoch, stream);
    cur ^= 1;
  }
  *result = cur;
}

template <typename K>
void sort_pairs(K* keys[2], uint32_t* vals[2], uint32_t n, int begin_bit, int end_bit, bool identity_vals,
                bool keep_keys, cudaStream_t stream, int* result) {
  *result = 0;
  if (n == 0) return;
  Sorter<K> s;
  s.prepare(n, begin_bit, end_bit, stream);
  s.histogram(keys[0], stream);
  s.run(keys, vals, identity_vals, keep_keys, stream, result);
}

}  // namespace rsort
}  // namespace pcg
