// vg_packed.cuh — the large-cloud VoxelGrid pipeline on packed 64-bit words (sm_100a).
//
// Every point becomes ONE word  (chunk id << key_bits | voxel key) << idx_bits | point index : the stable LSD sort of
// (key, index) pairs of voxelgrid.cu turns into a keys-only sort of 8-byte words over the key bits alone - the index
// rides in the low bits, and a stable sort on the key bits leaves equal keys in index order, which is the reference's
// accumulation order (voxelgrid.go:148-158).  With 9-bit digits a 36-bit key (the 50M-point map) takes 4 passes.
//
// A pass is NOT a decoupled look-back chain: measured on the 50M-point map, every tile of a onesweep pass walks back
// over ~40 predecessor tiles (one L2 round trip each, 4 KB of status words per step) because several hundred tiles
// are in flight at once - the chain, not HBM, set the pace (0.46 ms per pass for 1.2 GB).  Here the offsets come from
// a table instead: the input is cut into super-tiles of S x 4096 words, a histogram kernel counts the pass's digit
// per super-tile, a small scan turns the table into the first output slot of every (super-tile, digit), and the
// scatter kernel's CTAs are independent of one another.  24 bytes per word and pass cross HBM instead of 16, none of
// them waits for another CTA.
//
//   key_kernel      point -> word; counts the first pass's digit per super-tile in the same read
//   hist_kernel     digit counts per super-tile for the later passes (one streaming read)
//   base_kernel     table -> first output slot per (super-tile, digit)
//   scatter_kernel  one CTA per super-tile; its tiles arrive in shared memory as bulk-async copies (cp.async.bulk +
//                   mbarrier, UBLKCP / SYNCS in SASS), the next one in flight while the current one is ranked (warp
//                   match-any) and written out in runs through the buffer it arrived in
//   head_count_kernel + scan   voxels per reduce tile -> output slot of every tile's first voxel
//   reduce_kernel   segmented centroid + record emit over the sorted words; the points are gathered by index in
//                   sorted order from an aligned float4 copy - one 16-byte load per point, and the sweep keeps the
//                   copy L2-resident (gathering inside the last scatter pass instead was measured: no locality,
//                   80 bytes of DRAM traffic per point)
//
// Used when key bits + index bits fit 64 (always, for clouds the reference can hold); the (key, index) pipeline of
// voxelgrid.cu stays for wider keys and for the chunk-range variant.
#pragma once

#include "bulk_async.cuh"
#include "radix_sort.cuh"
#include "vg_common.cuh"

namespace pcg {
namespace vgp {

typedef unsigned long long u64;

constexpr int kBits = 9;
constexpr int kRadix = 1 << kBits;  // 512 digits: two per thread
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kIpt = 16;
constexpr int kTile = kThreads * kIpt;  // 4096 words = 32 KB
constexpr int kTileBytes = kTile * 8;
constexpr int kMaxPasses = 7;           // 63 key bits
constexpr int kRowBlock = 64;           // super-tiles per block of the base table
#ifndef PCG_VGP_CTAS
#define PCG_VGP_CTAS 3
#endif
constexpr int kCtasPerSm = PCG_VGP_CTAS;
#ifndef PCG_VGP_SUPER
#define PCG_VGP_SUPER 4
#endif
constexpr int kSuperTiles = PCG_VGP_SUPER;  // tiles per super-tile for large clouds
constexpr int kBufs = 2;  // tile buffers per scatter CTA: the next tile is in flight while one is ranked (a single
                          // buffer at 4-5 CTAs per SM was measured 17-38 % slower)

inline int passes_for(int total_bits) { return total_bits <= 0 ? 1 : (total_bits + kBits - 1) / kBits; }

__device__ __forceinline__ uint32_t digit_of(u64 w, int shift) { return (uint32_t)(w >> shift) & (kRadix - 1); }

// One lane per distinct digit of the warp adds (neighbouring points share most digits).  All 32 lanes must call.
__device__ __forceinline__ void hist_add(uint32_t* s_hist, uint32_t d, bool valid) {
  const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
  if (valid && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[d], (uint32_t)__popc(peers));
}

// ---- words + the first pass's counts -----------------------------------------------------------------------------
// One CTA per tile of 4096 points; H0[super-tile][digit] += the tile's counts (zeroed by the host).
__global__ void __launch_bounds__(kThreads)
    key_kernel(CloudView v, VgParams P, int idx_bits, uint32_t tiles_per_super, u64* __restrict__ words,
               float4* __restrict__ xyz4, uint32_t* __restrict__ H0, uint32_t* __restrict__ T0, int* __restrict__ flags) {
  __shared__ uint32_t s_hist[kRadix];
  s_hist[threadIdx.x] = 0;
  s_hist[threadIdx.x + kThreads] = 0;
  __syncthreads();
  const uint32_t n = (uint32_t)v.n;
  const uint32_t base = blockIdx.x * (uint32_t)kTile;
  int bad = 0;
  const KeyConsts C(P);
#pragma unroll 2
  for (int j = 0; j < kIpt; j++) {
    const uint32_t i = base + j * kThreads + threadIdx.x;
    const bool valid = i < n;
    u64 key = 0;
    if (valid) {
      const float3 pt = load_xyz(v, i);
      // aligned copy: the gather in sorted order is then ONE 16-byte load per point.  Three 4-byte loads from 12-byte
      // records touch 96 cache lines per warp and window - measured, the L1 tag rate then bounds the reduce
      xyz4[i] = make_float4(pt.x, pt.y, pt.z, 0.f);
      key = voxel_key_of(P, C, pt, &bad);
      words[i] = (key << idx_bits) | (u64)i;
    }
    hist_add(s_hist, (uint32_t)key & (kRadix - 1), valid);
  }
  if (bad) atomicOr(flags, bad);
  __syncthreads();
  const uint32_t super = blockIdx.x / tiles_per_super;
  uint32_t* row = H0 + (size_t)super * kRadix;
  uint32_t* trow = T0 + (size_t)(super / kRowBlock) * kRadix;  // column sums over the row block (see base_kernel)
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const uint32_t c = s_hist[threadIdx.x + q * kThreads];
    if (c) {
      atomicAdd(&row[threadIdx.x + q * kThreads], c);
      atomicAdd(&trow[threadIdx.x + q * kThreads], c);
    }
  }
}

// ---- digit counts per super-tile (passes after the first) ------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    hist_kernel(const u64* __restrict__ in, uint32_t n, int shift, uint32_t tiles_per_super, uint32_t* __restrict__ H,
                uint32_t* __restrict__ T) {
  __shared__ uint32_t s_hist[kRadix];
  s_hist[threadIdx.x] = 0;
  s_hist[threadIdx.x + kThreads] = 0;
  __syncthreads();
  const uint64_t begin = (uint64_t)blockIdx.x * tiles_per_super * kTile;
  const uint64_t end = min((uint64_t)n, begin + (uint64_t)tiles_per_super * kTile);
  for (uint64_t t = begin; t < end; t += 8 * kThreads) {  // eight independent loads per thread in flight
    u64 w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint64_t i = t + j * kThreads + threadIdx.x;
      w[j] = i < end ? __ldcs(in + i) : 0ull;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) hist_add(s_hist, digit_of(w[j], shift), t + j * kThreads + threadIdx.x < end);
  }
  __syncthreads();
  uint32_t* row = H + (size_t)blockIdx.x * kRadix;
  uint32_t* trow = T + (size_t)(blockIdx.x / kRowBlock) * kRadix;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const uint32_t c = s_hist[threadIdx.x + q * kThreads];
    row[threadIdx.x + q * kThreads] = c;
    if (c) atomicAdd(&trow[threadIdx.x + q * kThreads], c);
  }
}

// ---- table -> first output slot of every (super-tile, digit) ----------------------------------------------------------
// B[s][d] = (slots of all smaller digits) + (counts of digit d in the super-tiles before s).  T[rb][d] = the counts of
// digit d over the 64 super-tiles of row block rb, accumulated by the kernels that count.
// 1024 threads: two per digit, each half takes every other row block of T and 32 of the block's 64 rows.
__global__ void __launch_bounds__(2 * kRadix)
    base_kernel(const uint32_t* __restrict__ H, const uint32_t* __restrict__ T, uint32_t supers, uint32_t rowblocks,
                uint32_t* __restrict__ B) {
  __shared__ uint32_t s_warp[kRadix / 32];
  __shared__ uint32_t s_tot[kRadix], s_pre[kRadix], s_half[kRadix];
  const uint32_t d = threadIdx.x & (kRadix - 1), h = threadIdx.x >> kBits, lane = d & 31, warp = d >> 5;
  uint32_t tot = 0, pre = 0;
#pragma unroll 8
  for (uint32_t rb = h; rb < rowblocks; rb += 2) {
    const uint32_t t = T[(size_t)rb * kRadix + d];
    tot += t;
    pre += rb < blockIdx.x ? t : 0u;
  }
  // this half's 32 rows of the block, all loads in flight together
  constexpr int kHalfRows = kRowBlock / 2;
  const uint32_t s0 = blockIdx.x * kRowBlock + h * kHalfRows, s1 = min(supers, s0 + kHalfRows);
  uint32_t hv[kHalfRows];
  uint32_t mine = 0;
#pragma unroll
  for (int q = 0; q < kHalfRows; q++) {
    hv[q] = s0 + q < s1 ? H[(size_t)(s0 + q) * kRadix + d] : 0u;
    mine += hv[q];
  }
  if (h == 1) {
    s_tot[d] = tot;
    s_pre[d] = pre;
  } else {
    s_half[d] = mine;
  }
  __syncthreads();
  if (h == 0) {
    tot += s_tot[d];
    pre += s_pre[d];
  }
  uint32_t incl = tot;  // exclusive scan of the digit totals (first half's 512 threads)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  if (h == 0 && lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (h == 0) {
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kRadix / 32; w++) wbase += w < (int)warp ? s_warp[w] : 0u;
    s_tot[d] = wbase + incl - tot + pre;  // first slot of digit d in this row block
  }
  __syncthreads();
  uint32_t run = s_tot[d] + (h == 1 ? s_half[d] : 0u);
#pragma unroll
  for (int q = 0; q < kHalfRows; q++) {
    if (s0 + q < s1) B[(size_t)(s0 + q) * kRadix + d] = run;
    run += hv[q];
  }
}

// ---- one stable 9-bit pass over a super-tile ---------------------------------------------------------------------------
struct ScatterArgs {
  const u64* in;
  u64* out;
  const uint32_t* B;  // [supers][512] first output slot of the super-tile's words with that digit
  uint32_t n;
  uint32_t tiles_per_super;
  int shift;
};

__global__ void __launch_bounds__(kThreads, kCtasPerSm) scatter_kernel(ScatterArgs A) {
  extern __shared__ __align__(128) unsigned char vgp_dyn[];  // two tile buffers
  __shared__ __align__(16) uint16_t s_warp_hist[kWarps][kRadix];
  __shared__ uint32_t s_delta[kRadix];  // first output slot of the digit's run minus its position in the staged tile
  __shared__ uint32_t s_scan[kWarps];
  __shared__ __align__(8) uint64_t s_bar[2];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = A.n;
  const uint32_t tiles = (n + kTile - 1) / kTile;
  const uint32_t first_tile = blockIdx.x * A.tiles_per_super;
  const uint32_t end_tile = min(tiles, first_tile + A.tiles_per_super);
  const int shift = A.shift;

  auto fetch = [&](uint32_t t, int b) {  // thread 0: start the copy of tile t into buffer b
    const uint32_t cnt = min((uint32_t)kTile, n - t * (uint32_t)kTile);
    const uint32_t bytes = (cnt * 8u + 15u) & ~15u;  // the buffers are padded: a 16-byte granule may overhang n
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was last written by ordinary stores
    mbar_expect_tx(&s_bar[b], bytes);
    bulk_g2s(vgp_dyn + (size_t)b * kTileBytes, A.in + (size_t)t * kTile, bytes, &s_bar[b]);
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fetch(first_tile, 0);
  }
  // running first slot of this thread's two digits (2t, 2t+1)
  uint32_t base0 = A.B[(size_t)blockIdx.x * kRadix + 2 * tid], base1 = A.B[(size_t)blockIdx.x * kRadix + 2 * tid + 1];
  __syncthreads();

  uint32_t* hist32 = reinterpret_cast<uint32_t*>(&s_warp_hist[0][0]);  // [kWarps][256] pairs of 16-bit counters
  for (uint32_t tile = first_tile, j = 0; tile < end_tile; tile++, j++) {
    const int b = (int)(j & 1u);
    // the other buffer is free (its staged words were written out before the barrier that ended the last iteration)
    if (tid == 0 && tile + 1 < end_tile) fetch(tile + 1, b ^ 1);
    const uint32_t tile_base = tile * (uint32_t)kTile;
    const uint32_t tile_count = min((uint32_t)kTile, n - tile_base);
    u64* buf = reinterpret_cast<u64*>(vgp_dyn + (size_t)b * kTileBytes);
#pragma unroll
    for (int q = 0; q < kWarps * kRadix / 2 / kThreads; q++) hist32[q * kThreads + tid] = 0;
    mbar_wait(&s_bar[b], (j >> 1) & 1u);
    u64 keys[kIpt];
    const uint32_t wl = warp * (32u * kIpt) + lane;
#pragma unroll
    for (int i = 0; i < kIpt; i++) keys[i] = buf[wl + i * 32];
    __syncthreads();  // counters are zero; everyone holds its words: the buffer becomes the staging area

    // rank inside the warp: words are visited in index order (i major, lane minor)
    uint16_t offs[kIpt];
#pragma unroll
    for (int h = 0; h < kIpt; h += 8) {
      uint32_t peers_of[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {  // the match instructions are independent: issued back to back
        const bool valid = wl + (h + i) * 32 < tile_count;
        peers_of[i] = __match_any_sync(0xffffffffu, valid ? digit_of(keys[h + i], shift) : (uint32_t)kRadix);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const bool valid = wl + (h + i) * 32 < tile_count;
        const uint32_t d = digit_of(keys[h + i], shift);
        const uint32_t peers = peers_of[i];
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        // one atomic on the pair of 16-bit counters (no carry: a tile holds 4096 words); atomics of one warp on one
        // address execute in program order, so the returned counts need no barrier between the rounds
        if (valid && (int)lane == leader)
          pre = (atomicAdd(&hist32[warp * (kRadix / 2) + (d >> 1)], (uint32_t)__popc(peers) << ((d & 1u) * 16u)) >>
                 ((d & 1u) * 16u)) & 0xffffu;
        pre = __shfl_sync(0xffffffffu, pre, leader);
        offs[h + i] = (uint16_t)(pre + __popc(peers & ((1u << lane) - 1u)));
      }
    }
    __syncthreads();

    // thread t owns digits 2t and 2t+1 (one 32-bit word of every warp's counters; no carry: counts <= 4096)
    uint32_t cnt2 = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) cnt2 += hist32[w * (kRadix / 2) + tid];
    const uint32_t c0 = cnt2 & 0xffffu, c1 = cnt2 >> 16;
    const uint32_t ds = rsort::block_excl_scan_256(c0 + c1, s_scan, nullptr);  // position of the digit's run in the tile
    {
      uint32_t run = ds | ((ds + c0) << 16);
#pragma unroll
      for (int w = 0; w < kWarps; w++) {
        const uint32_t c = hist32[w * (kRadix / 2) + tid];
        hist32[w * (kRadix / 2) + tid] = run;
        run += c;
      }
      s_delta[2 * tid] = base0 - ds;
      s_delta[2 * tid + 1] = base1 - (ds + c0);
      base0 += c0;
      base1 += c1;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kIpt; i++) {
      if (wl + i * 32 < tile_count) {
        const uint32_t d = digit_of(keys[i], shift);
        buf[(uint32_t)s_warp_hist[warp][d] + offs[i]] = keys[i];
      }
    }
    __syncthreads();
    for (uint32_t s = tid; s < tile_count; s += kThreads) {
      const u64 k = buf[s];
      A.out[s_delta[digit_of(k, shift)] + s] = k;
    }
    __syncthreads();  // the staged words are read: the buffer can take the tile after next
  }
}

// ---- voxel heads per reduce tile -------------------------------------------------------------------------------------
// counts[t] = first words of a voxel among the sorted positions [1024 t, 1024 (t+1)); one warp per tile, a streaming
// read.  Their exclusive scan is the output slot of every tile's first voxel.
constexpr int kRedThreads = 256;
#ifndef PCG_VGP_RED_ITEMS
#define PCG_VGP_RED_ITEMS 4
#endif
constexpr int kRedItems = PCG_VGP_RED_ITEMS;
constexpr int kRedTile = kRedThreads * kRedItems;

__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_up1_u64(u64 v) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_up_sync(0xffffffffu, lo, 1);
  hi = __shfl_up_sync(0xffffffffu, hi, 1);
  return ((u64)hi << 32) | lo;
}

__global__ void __launch_bounds__(256)
    head_count_kernel(const u64* __restrict__ words, uint32_t n, int idx_bits, uint32_t* __restrict__ counts) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t seg_begin = tile * (uint32_t)kRedTile;
  if (seg_begin >= n) return;
  const uint32_t seg_end = min(n, seg_begin + (uint32_t)kRedTile);
  u64 prev_last = seg_begin > 0 ? words[seg_begin - 1] >> idx_bits : 0ull;
  uint32_t cnt = 0;
  for (int w0 = 0; w0 < kRedTile / 32; w0 += 8) {
    u64 k[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const uint32_t i = seg_begin + (w0 + q) * 32 + lane;
      k[q] = i < seg_end ? words[i] >> idx_bits : ~0ull;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const uint32_t i = seg_begin + (w0 + q) * 32 + lane;
      u64 pv = shfl_up1_u64(k[q]);
      if (lane == 0) pv = prev_last;
      cnt += __popc(__ballot_sync(0xffffffffu, i < seg_end && (i == 0 || k[q] != pv)));
      prev_last = shfl_u64(k[q], 31);
    }
  }
  if (lane == 0) counts[tile] = cnt;
}

// ---- segmented centroid + record emit over the sorted words ------------------------------------------------------------
#ifndef PCG_VGP_RED_CTAS
#define PCG_VGP_RED_CTAS 6
#endif
__global__ void __launch_bounds__(kRedThreads, PCG_VGP_RED_CTAS)
    reduce_kernel(CloudView v, VgParams P, const u64* __restrict__ words, const float4* __restrict__ xyz4, uint32_t n,
                  int idx_bits, uint8_t* __restrict__ out, const long long* __restrict__ first_slot,
                  long long* __restrict__ n_out) {
  __shared__ float4 s_pt[kRedTile];  // raw points of the tile's sorted slice (one 16-byte access per member)
  __shared__ u64 s_key[kRedTile];
  __shared__ uint16_t s_src[kRedTile];  // position of the tile's r-th voxel head
  __shared__ u64 s_after_key[32];       // the 32 positions after the tile
  __shared__ float4 s_after_pt[32];
  __shared__ uint32_t s_scan[kRedThreads / 32];
  __shared__ struct {
    u64 key, rank;
    float sx, sy, sz, fx, fy, fz, vc[3];
    uint32_t num, l;
    int pending;
  } s_cont;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 idx_mask = (1ull << idx_bits) - 1;
  if (tid == 0) s_cont.pending = 0;
  const uint32_t tile = blockIdx.x;
  const uint32_t tile_base = tile * kRedTile;
  // output slot of the tile's first voxel: heads of all preceding tiles (head_count_kernel + scan).  A decoupled
  // look-back inside this kernel was measured to BE the kernel's time: with ~740 tiles in flight the nearest
  // inclusive prefix lies ~600 tiles back, 19 L2 round trips of 32 predecessors each, per tile.
  const u64 base_rank = (u64)first_slot[tile];
  if (tid == 0 && (uint64_t)tile_base + kRedTile >= n) *n_out = first_slot[gridDim.x];
  const uint32_t tile_count = min((uint32_t)kRedTile, n - tile_base);

  // Phase A (striped: coalesced word loads, independent gathers from the input records; the sweep in sorted order
  // keeps the gathers of neighbouring voxels in the L2)
  // Loaded alongside (so that no dependent global load is left for later): the word before the tile (is the tile's
  // first word a head?) and, by the last warp, the 32 words + points after the tile - the tile's last voxel usually
  // ends among them.
  u64 word_before = 0;
  {
    u64 ww[kRedItems];
#pragma unroll
    for (int j = 0; j < kRedItems; j++) {
      const uint32_t l = j * kRedThreads + tid;
      ww[j] = l < tile_count ? words[tile_base + l] : 0ull;
    }
    if (tid == 0 && tile_base > 0) word_before = words[tile_base - 1];
    u64 wn = 0;
    const uint32_t after = tile_base + tile_count + lane;
    if (warp == kRedThreads / 32 - 1 && after < n) wn = words[after];
#pragma unroll
    for (int j = 0; j < kRedItems; j++) {
      const uint32_t l = j * kRedThreads + tid;
      if (l < tile_count) {
        const float4 pt = __ldg(xyz4 + (ww[j] & idx_mask));
        s_key[l] = ww[j] >> idx_bits;
        s_pt[l] = pt;
      }
    }
    if (warp == kRedThreads / 32 - 1) {
      s_after_key[lane] = after < n ? wn >> idx_bits : ~0ull;
      s_after_pt[lane] = after < n ? __ldg(xyz4 + (wn & idx_mask)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  // Phase B (blocked: thread t looks at positions 4t .. 4t+3): voxel heads, compacted
  const uint32_t l0 = tid * kRedItems;
  u64 prev = 0;
  if (l0 > 0 && l0 - 1 < tile_count)
    prev = s_key[l0 - 1];
  else if (l0 == 0 && tile_base > 0)
    prev = word_before >> idx_bits;
  uint32_t heads = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < kRedItems; j++) {
    const uint32_t l = l0 + j;
    if (l < tile_count) {
      const u64 k = s_key[l];
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
    for (int j = 0; j < kRedItems; j++)
      if ((heads >> j) & 1u) s_src[r++] = (uint16_t)(l0 + j);
  }
  const int out_aligned = v.aligned && ((((uintptr_t)out) & 3) == 0);
  const bool xyz_only = out_aligned && v.packed && v.stride == 12;

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
    if (xyz_only) {  // the record is the point itself
      float* d3 = reinterpret_cast<float*>(dst);
      d3[0] = ox;
      d3[1] = oy;
      d3[2] = oz;
      return;
    }
    const uint32_t first = (uint32_t)(words[tile_base + l] & idx_mask);
    const uint8_t* src = v.data + (uint64_t)first * (uint64_t)v.stride;
    if (out_aligned) {
      const uint32_t* s4 = (const uint32_t*)src;
      uint32_t* d4 = (uint32_t*)dst;
      const int nw = (int)(v.stride >> 2);
      for (int b = 0; b < nw; b++) d4[b] = __ldg(s4 + b);
    } else {
      for (int64_t b = 0; b < v.stride; b++) dst[b] = src[b];
    }
    if (num > 1) {
      store_f32_any(dst + v.off[0], ox, out_aligned);
      store_f32_any(dst + v.off[1], oy, out_aligned);
      store_f32_any(dst + v.off[2], oz, out_aligned);
    }
  };

  // One voxel per thread per round: the members of a voxel are added in list order out of shared memory
  // (voxelgrid.go:148-158) - the float32 sum is the reference's.  add(r) leaves voxel r's sums in the locals below and
  // returns false when the voxel may run on into the next tile (warp 0 finishes it at the end).
  long long vc_cid = -1;
  float vc[3] = {0.f, 0.f, 0.f};
  uint32_t a_l = 0, a_num = 0;
  float a_sx = 0.f, a_sy = 0.f, a_sz = 0.f, a_fx = 0.f, a_fy = 0.f, a_fz = 0.f;
  auto add = [&](uint32_t r) -> bool {
    const uint32_t l = s_src[r];
    const u64 key = s_key[l];
    {
      const long long cid = (long long)(key >> P.key_bits);
      if (cid != vc_cid) {
        chunk_min(P, cid, vc);
        vc_cid = cid;
      }
    }
    // the voxel's members are the positions up to the next head: the trip count is known, no key is compared
    const uint32_t l_end = r + 1 < total ? (uint32_t)s_src[r + 1] : tile_count;
    a_l = l;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    {
      const float4 p = s_pt[l];
      a_fx = p.x;
      a_fy = p.y;
      a_fz = p.z;
    }
    for (uint32_t ll = l; ll < l_end; ll++) {
      const float4 p = s_pt[ll];
      sx = __fadd_rn(sx, __fsub_rn(p.x, vc[0]));
      sy = __fadd_rn(sy, __fsub_rn(p.y, vc[1]));
      sz = __fadd_rn(sz, __fsub_rn(p.z, vc[2]));
    }
    const uint32_t num = l_end - l;
    a_sx = sx;
    a_sy = sy;
    a_sz = sz;
    a_num = num;
    if (l_end == tile_count && tile_base + tile_count < n) {
      s_cont.key = key;
      s_cont.rank = r;  // relative to the tile's first voxel
      s_cont.sx = sx;
      s_cont.sy = sy;
      s_cont.sz = sz;
      s_cont.fx = a_fx;
      s_cont.fy = a_fy;
      s_cont.fz = a_fz;
      s_cont.vc[0] = vc[0];
      s_cont.vc[1] = vc[1];
      s_cont.vc[2] = vc[2];
      s_cont.num = num;
      s_cont.l = l;
      s_cont.pending = 1;
      return false;
    }
    return true;
  };
  __syncthreads();  // s_src is complete
  for (uint32_t r = tid; r < total; r += kRedThreads)
    if (add(r)) emit(a_l, a_num, a_sx, a_sy, a_sz, a_fx, a_fy, a_fz, vc, base_rank + r);
  __syncthreads();
  if (warp == 0 && s_cont.pending) {
    const u64 key = s_cont.key;
    float sx = s_cont.sx, sy = s_cont.sy, sz = s_cont.sz;
    const float c0 = s_cont.vc[0], c1 = s_cont.vc[1], c2 = s_cont.vc[2];
    uint32_t num = s_cont.num;
    for (uint32_t g = tile_base + tile_count;; g += 32) {
      const uint32_t idx = g + lane;
      const bool staged = g == tile_base + tile_count;  // the first window was loaded with the tile
      u64 w = 0;
      bool match = false;
      if (staged) {
        match = s_after_key[lane] == key;
      } else if (idx < n) {
        w = words[idx];
        match = (w >> idx_bits) == key;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, match);
      const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;  // members are consecutive: the leading matches
      float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((int)lane < run) pt = staged ? s_after_pt[lane] : __ldg(xyz4 + (w & idx_mask));
      for (int q = 0; q < run; q++) {  // the additions stay in list order (every lane carries the same sums)
        sx = __fadd_rn(sx, __fsub_rn(__shfl_sync(0xffffffffu, pt.x, q), c0));
        sy = __fadd_rn(sy, __fsub_rn(__shfl_sync(0xffffffffu, pt.y, q), c1));
        sz = __fadd_rn(sz, __fsub_rn(__shfl_sync(0xffffffffu, pt.z, q), c2));
      }
      num += (uint32_t)run;
      if (run < 32) break;
    }
    if (lane == 0)
      emit(s_cont.l, num, sx, sy, sz, s_cont.fx, s_cont.fy, s_cont.fz, s_cont.vc, base_rank + s_cont.rank);
  }
}

inline bool fits(int64_t n, int total_bits) { return bits_for(n) + total_bits <= 64 && total_bits <= kMaxPasses * kBits; }

// keys -> sort -> reduce for the whole cloud; *d_n_out = records written.  Stream-ordered, no host synchronisation.
inline void run(const CloudView& v, const VgParams& P, int total_bits, uint8_t* d_out, long long* d_n_out, int* d_flags,
                cudaStream_t stream) {
  const uint32_t n = (uint32_t)v.n;
  const int idx_bits = std::max(1, bits_for((long long)n));
  const int passes = passes_for(total_bits);
  const uint32_t tiles = (n + kTile - 1) / kTile;
  // super-tiles: as large as leaves the scatter kernel at least six full waves of CTAs, at most 4096 of them
  uint32_t S = std::min<uint32_t>((uint32_t)kSuperTiles, std::max<uint32_t>(1u, tiles / (kNumSMs * kCtasPerSm * 6)));
  while ((tiles + S - 1) / S > 4096u) S *= 2;
  const uint32_t supers = (tiles + S - 1) / S;
  const uint32_t rowblocks = (supers + kRowBlock - 1) / kRowBlock;
  const uint32_t rtiles = (n + kRedTile - 1) / kRedTile;
  // ONE allocation per call (a pool that served many small requests before hands out eight large ones slowly):
  // [T: passes x rowblocks x 512 | H: supers x 512 | B: supers x 512 | head counts | first slots | words 0 | words 1 |
  //  float4 points]; T and H (first pass) are accumulated, hence zeroed
  const size_t table = (size_t)supers * kRadix, ttable = (size_t)rowblocks * kRadix;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t tab_bytes = up(((size_t)passes * ttable + 2 * table) * sizeof(uint32_t));
  const size_t hc_bytes = up((size_t)rtiles * sizeof(uint32_t));
  const size_t fs_bytes = up(((size_t)rtiles + 1) * sizeof(long long));
  const size_t w_bytes = up(((size_t)n + 2) * sizeof(u64));
  DevBuf<uint8_t> ws(tab_bytes + hc_bytes + fs_bytes + 2 * w_bytes + (size_t)n * sizeof(float4), stream);
  uint32_t* T = reinterpret_cast<uint32_t*>(ws.p);
  uint32_t* H = T + (size_t)passes * ttable;
  uint32_t* B = H + table;
  uint32_t* head_counts = reinterpret_cast<uint32_t*>(ws.p + tab_bytes);
  long long* first_slot = reinterpret_cast<long long*>(ws.p + tab_bytes + hc_bytes);
  u64* w0 = reinterpret_cast<u64*>(ws.p + tab_bytes + hc_bytes + fs_bytes);
  u64* w1 = reinterpret_cast<u64*>(ws.p + tab_bytes + hc_bytes + fs_bytes + w_bytes);
  float4* xyz4 = reinterpret_cast<float4*>(ws.p + tab_bytes + hc_bytes + fs_bytes + 2 * w_bytes);
  PCG_CUDA(cudaMemsetAsync(T, 0, ((size_t)passes * ttable + table) * sizeof(uint32_t), stream));
  PCG_LAUNCH_NAMED("vgp::key_kernel", key_kernel, tiles, kThreads, 0, stream, v, P, idx_bits, S, w0, xyz4, H, T, d_flags);

  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  PCG_CUDA(cudaGetDevice(&dev));
  if (!(configured.load(std::memory_order_relaxed) & (1ull << dev))) {
    PCG_CUDA(cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBufs * kTileBytes));
    configured.fetch_or(1ull << dev, std::memory_order_relaxed);
  }
  u64* buf[2] = {w0, w1};
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    const int shift = idx_bits + p * kBits;
    uint32_t* Tp = T + (size_t)p * ttable;
    if (p > 0)
      PCG_LAUNCH_NAMED("vgp::hist_kernel", hist_kernel, supers, kThreads, 0, stream, buf[cur], n, shift, S, H, Tp);
    PCG_LAUNCH_NAMED("vgp::base_kernel", base_kernel, rowblocks, 2 * kRadix, 0, stream, H, Tp, supers, rowblocks, B);
    ScatterArgs a;
    a.in = buf[cur];
    a.out = buf[cur ^ 1];
    a.B = B;
    a.n = n;
    a.tiles_per_super = S;
    a.shift = shift;
    PCG_LAUNCH_NAMED("vgp::scatter_kernel", scatter_kernel, supers, kThreads, kBufs * kTileBytes, stream, a);
    cur ^= 1;
  }
  PCG_LAUNCH_NAMED("vgp::head_count_kernel", head_count_kernel, (rtiles + 7) / 8, 256, 0, stream, buf[cur], n, idx_bits,
                   head_counts);
  scan_counts(head_counts, first_slot, rtiles, stream);
  PCG_LAUNCH_NAMED("vgp::reduce_kernel", reduce_kernel, rtiles, kRedThreads, 0, stream, v, P, buf[cur], xyz4, n,
                   idx_bits, d_out, first_slot, d_n_out);
}

}  // namespace vgp
}  // namespace pcg
