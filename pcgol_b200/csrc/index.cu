// index.cu — GPU-built spatial index behind storage.Search (replaces kdtree.New /
// KDTree.Nearest / KDTree.Range, pc/storage/kdtree/kdtree.go:33-222) on sm_100a.
//
// Build: bounding box -> 48-bit Morton keys (16 bits per axis, cubic cells) -> stable
// radix sort of (key, index) -> gather float4 {x,y,z,index} in Morton order -> leaf
// boxes over 8 consecutive points and an implicit, heap-indexed binary tree of boxes
// above them (see bvh.cuh).  No pointers, no recursion, two kernels for the tree.
#include "bvh.cuh"
#include "radix_sort.cuh"

namespace pcg {


__device__ __forceinline__ uint32_t ord_bits(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_float(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

__global__ void bbox_init_kernel(uint32_t* out6) {
  if (threadIdx.x < 6) out6[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
}

// out6: ordered-bit min x,y,z then max x,y,z over finite coordinates
__global__ void __launch_bounds__(256) bbox_kernel(CloudView v, uint32_t* __restrict__ out6) {
  __shared__ uint32_t s_red[8][6];
  uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += stride) {
    float3 p = load_xyz(v, i);
    float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (!isfinite(c[k])) continue;
      uint32_t o = ord_bits(c[k]);
      mn[k] = min(mn[k], o);
      mx[k] = max(mx[k], o);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn[k] = min(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
      mx[k] = max(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      s_red[threadIdx.x >> 5][k] = mn[k];
      s_red[threadIdx.x >> 5][3 + k] = mx[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    uint32_t r = s_red[0][k];
    for (int w = 1; w < 8; w++) r = k < 3 ? min(r, s_red[w][k]) : max(r, s_red[w][k]);
    if (k < 3)
      atomicMin(&out6[k], r);
    else
      atomicMax(&out6[k], r);
  }
}

__device__ __forceinline__ unsigned long long spread16(uint32_t x) {  // 16 bits -> every third bit
  unsigned long long v = x & 0xffffu;
  v = (v | (v << 16)) & 0x0000ff0000ffull;
  v = (v | (v << 8)) & 0x00f00f00f00full;
  v = (v | (v << 4)) & 0x0c30c30c30c3ull;
  v = (v | (v << 2)) & 0x249249249249ull;
  return v;
}

// 3-D Hilbert index (Skilling's transpose algorithm) of BITS-bit coordinates.  Unlike the Morton
// curve the Hilbert curve never jumps: runs of consecutive points - the leaves and every implicit
// node above them - stay compact, which is what makes the packed tree prune well
// (measured: 12 -> see DESIGN.md leaf scans per query with Morton order).
template <int BITS>
__device__ __forceinline__ unsigned long long hilbert_key(uint32_t x0, uint32_t x1, uint32_t x2) {
  uint32_t X[3] = {x0, x1, x2};
  const uint32_t M = 1u << (BITS - 1);
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    const uint32_t Pm = Q - 1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (X[i] & Q) {
        X[0] ^= Pm;
      } else {
        const uint32_t t = (X[0] ^ X[i]) & Pm;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  X[1] ^= X[0];
  X[2] ^= X[1];
  uint32_t t = 0;
  for (uint32_t Q = M; Q > 1; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t;
  X[1] ^= t;
  X[2] ^= t;
  return (spread16(X[0]) << 2) | (spread16(X[1]) << 1) | spread16(X[2]);
}

__global__ void __launch_bounds__(256)
    gather_points_kernel(CloudView v, const uint32_t* __restrict__ order, float4* __restrict__ pts, uint32_t padded) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= padded) return;
  float4 o;
  if (i < (uint32_t)v.n) {
    uint32_t src = order[i];
    float3 p = load_xyz(v, src);
    o = make_float4(p.x, p.y, p.z, __uint_as_float(src));
  } else {
    const float inf = __int_as_float(0x7f800000);
    o = make_float4(inf, inf, inf, __uint_as_float(0xffffffffu));
  }
  store_point(pts, i, o);  // coordinate-major inside the leaf's line (bvh.cuh)
}

// Leaf boxes and the 8 levels above them, one CTA per 256 leaves.
__global__ void __launch_bounds__(256)
    leaf_boxes_kernel(const float4* __restrict__ pts, float4* __restrict__ boxes, uint32_t leaves, uint32_t P) {
  __shared__ float s_lo[256][3];
  __shared__ float s_hi[256][3];
  const uint32_t t = threadIdx.x;
  const uint32_t leaf = blockIdx.x * 256 + t;
  const float inf = __int_as_float(0x7f800000);
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  if (leaf < leaves) {
#pragma unroll
    for (int j = 0; j < kLeaf; j++) {
      const float4 p = load_point(pts, leaf * kLeaf + j);
      if (__float_as_uint(p.w) == 0xffffffffu) continue;  // padding
      // fminf/fmaxf drop NaN operands; +-inf coordinates stay out of the boxes as well
      if (isfinite(p.x)) { lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x); }
      if (isfinite(p.y)) { lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y); }
      if (isfinite(p.z)) { lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z); }
    }
  }
  if (leaf < P) store_box(boxes, P + leaf, lo, hi);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    s_lo[t][k] = lo[k];
    s_hi[t][k] = hi[k];
  }
  __syncthreads();
  // heights 1..8 inside the CTA: node ids ((P + blockIdx*256) >> h) + j
  uint32_t width = 256;
  uint32_t first = P + blockIdx.x * 256;
  for (int h = 1; h <= 8; h++) {
    width >>= 1;
    first >>= 1;
    if (first == 0) break;  // above the root
    const uint32_t valid = min(width, P >> h);  // a tree smaller than one CTA has fewer nodes per level
    float l[3], u[3];
    if (t < valid) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        l[k] = fminf(s_lo[2 * t][k], s_lo[2 * t + 1][k]);
        u[k] = fmaxf(s_hi[2 * t][k], s_hi[2 * t + 1][k]);
      }
    }
    __syncthreads();
    if (t < valid) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        s_lo[t][k] = l[k];
        s_hi[t][k] = u[k];
      }
      store_box(boxes, first + t, l, u);
    }
    __syncthreads();
  }
}

// Levels above height 8 (P/256 nodes and fewer), one CTA walking up level by level.
__global__ void __launch_bounds__(1024) top_boxes_kernel(float4* __restrict__ boxes, uint32_t P) {
  // level with `count` nodes starting at node id `count` (heap indexing), children already written
  for (uint32_t count = P >> 9; count >= 1; count >>= 1) {
    for (uint32_t j = threadIdx.x; j < count; j += blockDim.x) {
      const uint32_t node = count + j;
      // plain loads: the children were written by this kernel (or the one before) - not through the read-only path
      const float* c0 = reinterpret_cast<const float*>(boxes) + (size_t)((2 * node) >> 2) * 32 + ((2 * node) & 3u);
      float lo[3], hi[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        lo[k] = fminf(c0[4 * k], c0[4 * k + 1]);
        hi[k] = fmaxf(c0[12 + 4 * k], c0[12 + 4 * k + 1]);
      }
      store_box(boxes, node, lo, hi);
    }
    __syncthreads();
  }
}


// ---- KD order ---------------------------------------------------------------------------------
// The implicit tree (bvh.cuh) only fixes WHICH slots a node covers: node at height h owns 2^h
// consecutive leaves.  Which points sit in those slots is free, and it decides how well the boxes
// prune.  A space-filling-curve order is one radix sort away but its runs are only as compact as
// the curve; the order below is the one a KD-tree build produces: every node splits its points
// along the widest axis of their bounding box, the first 2^(h-1) leaves' worth of points (by rank
// along that axis) going to the left child.  Measured on the 1M-point LiDAR scan with 10M queries
// (tools/nn_model.cpp, same counters as the kernel): 27.8 node steps + 3.5 leaf scans per query
// against 59.9 + 7.4 for the Hilbert order.
//
// Build = the classic presorted-list KD construction, level-synchronous and pointer-free:
//   three stable radix sorts give the point ids ordered by x, by y and by z;
//   per level, every segment (fixed, aligned slot range) picks its axis from the ends of its three
//   lists (that IS its bounding box), marks each of its points left/right by its rank in the list
//   of that axis, and stably partitions the two other lists by that mark (segmented prefix sums).
// All three lists stay sorted inside every segment, so the next level needs nothing else.
constexpr int kKdTileItems = 8;
constexpr int kKdTile = 256 * kKdTileItems;

__device__ __forceinline__ float coord_of(const CloudView& v, uint32_t id, int axis) {
  const float3 p = load_xyz(v, id);
  return axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
}

// ordered-bit keys of x, y, z (a total order that also places NaN / inf somewhere), with the digit
// histograms of the three sorts accumulated on the way
__global__ void __launch_bounds__(256)
    kd_keys_kernel(CloudView v, uint32_t* __restrict__ k0, uint32_t* __restrict__ k1, uint32_t* __restrict__ k2,
                   uint32_t* __restrict__ h0, uint32_t* __restrict__ h1, uint32_t* __restrict__ h2, int passes) {
  __shared__ uint32_t s_hist[3][4 * rsort::kRadix];
  for (int c = 0; c < 3; c++) rsort::hist_zero(s_hist[c], passes);
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (v.n + stride - 1) / stride;
  for (int64_t r = 0; r < rounds; r++) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < v.n;
    uint32_t a = 0, b = 0, c = 0;
    if (valid) {
      const float3 p = load_xyz(v, i);
      a = ord_bits(p.x);
      b = ord_bits(p.y);
      c = ord_bits(p.z);
      k0[i] = a;
      k1[i] = b;
      k2[i] = c;
    }
    rsort::hist_add_key(s_hist[0], a, valid, 0, passes);
    rsort::hist_add_key(s_hist[1], b, valid, 0, passes);
    rsort::hist_add_key(s_hist[2], c, valid, 0, passes);
  }
  __syncthreads();
  rsort::hist_flush(s_hist[0], h0, passes);
  rsort::hist_flush(s_hist[1], h1, passes);
  rsort::hist_flush(s_hist[2], h2, passes);
}

struct KdLists {
  const uint32_t* in[3];
  uint32_t* out[3];
};

// One level of the build while a segment still spans several tiles (one segment per CTA).  Segment s covers
// positions [s*S, min((s+1)*S, n)) and splits at m = s*S + S/2 when m lies before its end:
//   axis   = widest extent, read off the ends of the segment's three sorted lists (that IS its bounding box);
//   pivot  = the element at position m of the list of that axis; a point goes left iff it ranks before the pivot
//            in that list, i.e. iff (key, id) < (pivot key, pivot id) - the lists were sorted stably by key from
//            the identity, so ties are in id order - which every thread decides from the point's own coordinate;
//   the two other lists are partitioned stably; the left-counts of the segment's earlier tiles arrive through a
//   decoupled look-back (tiles take their ids from a counter, so predecessors are always running; one packed word
//   per tile and list: state | level | count), restarted at every segment's first tile.
constexpr unsigned long long kKdAggregate = 1ull << 62, kKdInclusive = 2ull << 62;
__global__ void __launch_bounds__(256)
    kd_level_kernel(CloudView v, KdLists L, uint32_t n, int log2S, uint32_t tiles, uint32_t* __restrict__ tile_counter,
                    unsigned long long* __restrict__ status) {
  __shared__ uint32_t s_scan[rsort::kWarps];
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_tile;
  __shared__ int s_axis;
  __shared__ uint32_t s_pivot_key, s_pivot_id;
  const uint32_t tid = threadIdx.x;
  const uint32_t S = 1u << log2S;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_base = tile * kKdTile;  // < n: there are ceil(n / kKdTile) tiles
  const uint32_t s = tile_base >> log2S, b = s << log2S, m = b + (S >> 1);
  const uint32_t e = min(b + S, n);
  if (tid == 0) {
    int axis = 3;
    if (m < e) {
      float ext[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float lo = coord_of(v, L.in[c][b], c), hi = coord_of(v, L.in[c][e - 1], c);
        float d = hi - lo;
        if (!(d >= 0.f)) d = 0.f;  // NaN ends
        ext[c] = d;
      }
      axis = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
      const uint32_t pid = L.in[axis][m];
      s_pivot_id = pid;
      s_pivot_key = ord_bits(coord_of(v, pid, axis));
    }
    s_axis = axis;
  }
  __syncthreads();
  const int axis = s_axis;
  const uint32_t pkey = s_pivot_key, pid = s_pivot_id;
  const uint32_t base = tile_base + tid * kKdTileItems;
#pragma unroll 1
  for (int c = 0; c < 3; c++) {
    uint32_t id[kKdTileItems];
    uint32_t left_mask = 0, cnt = 0;
    const bool part = axis != 3 && axis != c;
#pragma unroll
    for (int j = 0; j < kKdTileItems; j++) {
      id[j] = 0;
      if (base + j < n) id[j] = L.in[c][base + j];
    }
    if (part) {
#pragma unroll
      for (int j = 0; j < kKdTileItems; j++) {
        if (base + j < n) {
          const uint32_t key = ord_bits(coord_of(v, id[j], axis));
          if (key < pkey || (key == pkey && id[j] < pid)) {
            left_mask |= 1u << j;
            cnt++;
          }
        }
      }
    }
    uint32_t total = 0;
    const uint32_t excl = rsort::block_excl_scan_256(cnt, s_scan, &total);
    if (tid < 32) {  // warp 0 looks back 32 predecessor tiles per round
      uint32_t prefix = 0;
      if (part) {
        volatile unsigned long long* st = status + (size_t)c * tiles;
        const unsigned long long tag = (unsigned long long)(uint32_t)log2S << 48;
        const uint32_t first_tile = b / kKdTile;
        if (tile == first_tile) {
          if (tid == 0) st[tile] = kKdInclusive | tag | total;
        } else {
          if (tid == 0) st[tile] = kKdAggregate | tag | total;
          long long look = (long long)tile - 1;
          for (;;) {
            const long long idx = look - (long long)tid;
            const bool before = idx < (long long)first_tile;  // ahead of the segment: contributes nothing, ends the walk
            unsigned long long w = 0;
            if (!before) w = st[idx];
            const bool ready = before || ((w >> 62) != 0 && ((w >> 48) & 0xffu) == (unsigned long long)(uint32_t)log2S);
            if (!__all_sync(0xffffffffu, ready)) continue;  // someone has not published at this level yet
            const uint32_t stop = __ballot_sync(0xffffffffu, before || (w >> 62) == 2);
            const int last = stop ? __ffs(stop) - 1 : 31;  // lanes 0..last contribute
            uint32_t val = (!before && (int)tid <= last) ? (uint32_t)w : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            prefix += val;
            if (stop) break;
            look -= 32;
          }
          if (tid == 0) st[tile] = kKdInclusive | tag | (unsigned long long)(prefix + total);
        }
      }
      if (tid == 0) s_prefix = prefix;
    }
    __syncthreads();
    uint32_t lb = s_prefix + excl;  // lefts of this segment that precede this thread's items
#pragma unroll
    for (int j = 0; j < kKdTileItems; j++) {
      const uint32_t pos = base + j;
      if (pos < n) {
        uint32_t dst = pos;
        if (part) {
          const bool is_left = (left_mask >> j) & 1u;
          dst = is_left ? b + lb : m + ((pos - b) - lb);
          lb += is_left ? 1u : 0u;
        }
        L.out[c][dst] = id[j];
      }
    }
    __syncthreads();  // s_prefix is reused by the next list
  }
}

// Levels whose segments fit one tile: the whole tail of the build in shared memory, one CTA per
// kKdLocal slots.  Points get tile-local ids (their position in the tile's x-list), so the three
// lists are 16-bit and the left/right marks live in shared memory; per level: axis per segment,
// marks, two stable partitions.  Only the x-list is needed afterwards: it is the final order.
constexpr int kKdLocalLog2 = 11;
constexpr int kKdLocal = 1 << kKdLocalLog2;
static_assert(kKdLocal == kKdTile, "the local kernel takes over where segments stop spanning tiles");

__global__ void __launch_bounds__(256)
    kd_local_kernel(CloudView v, const uint32_t* __restrict__ in0, const uint32_t* __restrict__ in1,
                    const uint32_t* __restrict__ in2, uint32_t n, int log2S_first, int log2L,
                    uint32_t* __restrict__ loc, uint32_t* __restrict__ order_out) {
  __shared__ uint16_t s_list[2][3][kKdLocal];
  __shared__ uint32_t s_gid[kKdLocal];
  __shared__ uint8_t s_side[kKdLocal];
  __shared__ uint8_t s_axis[kKdLocal / kKdTileItems];  // segments are never smaller than one thread's items
  __shared__ uint32_t s_scan[rsort::kWarps];
  __shared__ uint32_t s_excl[256];
  const uint32_t tid = threadIdx.x;
  const uint32_t tile_base = blockIdx.x * kKdLocal;
  const uint32_t cnt = min((uint32_t)kKdLocal, n - tile_base);
  for (uint32_t i = tid; i < cnt; i += 256) {
    const uint32_t g = in0[tile_base + i];
    s_gid[i] = g;
    loc[g] = i;
    s_list[0][0][i] = (uint16_t)i;
  }
  __syncthreads();  // loc[] of this tile's points is complete (the three lists hold the same point set)
  for (uint32_t i = tid; i < cnt; i += 256) {
    s_list[0][1][i] = (uint16_t)loc[in1[tile_base + i]];
    s_list[0][2][i] = (uint16_t)loc[in2[tile_base + i]];
  }
  __syncthreads();
  int cur = 0;
  const uint32_t base = tid * kKdTileItems;
  for (int log2S = log2S_first; log2S > log2L; log2S--) {
    const uint32_t S = 1u << log2S;
    if (tid < ((uint32_t)kKdLocal >> log2S)) {
      const uint32_t b = tid << log2S;
      int axis = 3;
      if (b < cnt) {
        const uint32_t e = min(b + S, cnt), m = b + (S >> 1);
        if (m < e) {
          float ext[3];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const float lo = coord_of(v, s_gid[s_list[cur][c][b]], c), hi = coord_of(v, s_gid[s_list[cur][c][e - 1]], c);
            float d = hi - lo;
            if (!(d >= 0.f)) d = 0.f;
            ext[c] = d;
          }
          axis = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
        }
      }
      s_axis[tid] = (uint8_t)axis;
    }
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += 256) {
      const uint32_t seg = i >> log2S;
      const int a = s_axis[seg];
      if (a != 3) s_side[s_list[cur][a][i]] = i >= (seg << log2S) + (S >> 1) ? 1 : 0;
    }
    __syncthreads();
    const uint32_t seg = base >> log2S, b = seg << log2S, m = b + (S >> 1);
    const int axis = base < cnt ? s_axis[seg] : 3;
#pragma unroll 1
    for (int c = 0; c < 3; c++) {
      const bool part = axis != 3 && axis != c;
      uint16_t id[kKdTileItems];
      uint32_t left_mask = 0, nl = 0;
#pragma unroll
      for (int j = 0; j < kKdTileItems; j++) {
        id[j] = 0;
        if (base + j < cnt) {
          id[j] = s_list[cur][c][base + j];
          if (part && !s_side[id[j]]) {
            left_mask |= 1u << j;
            nl++;
          }
        }
      }
      const uint32_t excl = rsort::block_excl_scan_256(nl, s_scan, nullptr);
      s_excl[tid] = excl;
      __syncthreads();
      uint32_t lb = excl - s_excl[b / kKdTileItems];
#pragma unroll
      for (int j = 0; j < kKdTileItems; j++) {
        const uint32_t pos = base + j;
        if (pos < cnt) {
          uint32_t dst = pos;
          if (part) {
            const bool is_left = (left_mask >> j) & 1u;
            dst = is_left ? b + lb : m + ((pos - b) - lb);
            lb += is_left ? 1u : 0u;
          }
          s_list[cur ^ 1][c][dst] = id[j];
        }
      }
      __syncthreads();
    }
    cur ^= 1;
  }
  for (uint32_t i = tid; i < cnt; i += 256) order_out[tile_base + i] = s_gid[s_list[cur][0][i]];
}

// Writes the KD order into lists[.]; returns the buffer that holds it (a permutation of 0..n-1).
static const uint32_t* kd_order_device(const CloudView& v, uint32_t n, uint32_t P, DevBuf<uint32_t> (&lists)[6],
                                       cudaStream_t stream) {
  DevBuf<uint32_t> keys[3], keys_alt[3];
  rsort::Sorter<uint32_t> sorter[3];
  for (int c = 0; c < 3; c++) {
    keys[c].alloc(n, stream);
    keys_alt[c].alloc(n, stream);
    sorter[c].prepare(n, 0, 32, stream);
  }
  for (int c = 0; c < 6; c++) lists[c].alloc(n, stream);
  const int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(n, 256));
  PCG_LAUNCH(kd_keys_kernel, blocks, 256, 0, stream, v, keys[0].p, keys[1].p, keys[2].p, sorter[0].hist(),
             sorter[1].hist(), sorter[2].hist(), sorter[0].passes);
  uint32_t* cur[3];
  uint32_t* alt[3];
  {
    // the three axis sorts share every launch (blockIdx.y = axis)
    uint32_t* kk[3][2];
    uint32_t* vv[3][2];
    for (int c = 0; c < 3; c++) {
      kk[c][0] = keys[c].p;
      kk[c][1] = keys_alt[c].p;
      vv[c][0] = lists[c].p;
      vv[c][1] = lists[3 + c].p;
    }
    int res = 0;
    rsort::run_batched<uint32_t>(sorter, 3, kk, vv, /*identity_vals=*/true, /*keep_keys=*/false, stream, &res);
    for (int c = 0; c < 3; c++) {
      cur[c] = vv[c][res];
      alt[c] = vv[c][res ^ 1];
    }
  }
  const uint32_t M = P * (uint32_t)kLeaf;  // slots of the complete tree
  if (P <= 1) return cur[0];
  int log2M = 0;
  while ((1u << log2M) < M) log2M++;
  int log2L = 0;
  while ((1 << log2L) < kLeaf) log2L++;
  const uint32_t tiles = (uint32_t)div_up(n, kKdTile);
  DevBuf<unsigned long long> status((size_t)3 * tiles + 32, stream);  // look-back words + one tile counter per level
  PCG_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), stream));
  uint32_t* counters = reinterpret_cast<uint32_t*>(status.p + (size_t)3 * tiles);
  int log2S = log2M;
  for (; log2S > log2L && log2S > kKdLocalLog2; log2S--) {  // segments that span several tiles
    KdLists L;
    for (int c = 0; c < 3; c++) {
      L.in[c] = cur[c];
      L.out[c] = alt[c];
    }
    PCG_LAUNCH(kd_level_kernel, tiles, 256, 0, stream, v, L, n, log2S, tiles, counters + log2S, status.p);
    for (int c = 0; c < 3; c++) std::swap(cur[c], alt[c]);
  }
  if (log2S > log2L) {  // the rest (children of the last level are single leaves) in shared memory
    DevBuf<uint32_t> loc(n, stream);  // point id -> tile-local id
    PCG_LAUNCH(kd_local_kernel, tiles, 256, 0, stream, v, cur[0], cur[1], cur[2], n, log2S, log2L, loc.p, alt[0]);
    return alt[0];
  }
  return cur[0];
}

Index* index_build_device(const CloudView& v, int device, cudaStream_t stream) {
  Index* ix = new Index();
  ix->device = device;
  ix->n = v.n;
  const uint32_t n = (uint32_t)v.n;
  ix->leaves = (n + kLeaf - 1) / kLeaf;
  uint32_t P = 1;
  while (P < ix->leaves) P <<= 1;
  ix->P = P;
  if (n == 0) return ix;
  try {
    const size_t padded = (size_t)ix->leaves * kLeaf;
    const size_t pts_bytes = padded * sizeof(float4);
    const size_t box_bytes = (size_t)4 * std::max<uint32_t>(P, 2u) * sizeof(float4);  // whole lines of four nodes (bvh.cuh)
    PCG_CUDA(cudaMallocAsync((void**)&ix->pts, pts_bytes, stream));
    PCG_CUDA(cudaMallocAsync((void**)&ix->boxes, box_bytes, stream));
    ix->bytes = (int64_t)(pts_bytes + box_bytes);

    PCG_CUDA(cudaMallocAsync((void**)&ix->bbox, 8 * sizeof(uint32_t), stream));
    PCG_LAUNCH(bbox_init_kernel, 1, 32, 0, stream, ix->bbox);
    int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(n, 256));
    PCG_LAUNCH(bbox_kernel, blocks, 256, 0, stream, v, ix->bbox);

    {
      DevBuf<uint32_t> lists[6];
      const uint32_t* order = kd_order_device(v, n, P, lists, stream);
      PCG_LAUNCH(gather_points_kernel, div_up(padded, 256), 256, 0, stream, v, order, ix->pts, (uint32_t)padded);
    }
    PCG_LAUNCH(leaf_boxes_kernel, div_up(P, 256), 256, 0, stream, ix->pts, ix->boxes, ix->leaves, P);
    if (P >= 512) PCG_LAUNCH(top_boxes_kernel, 1, 1024, 0, stream, ix->boxes, P);
    PCG_CUDA(cudaEventCreateWithFlags(&ix->ready, cudaEventDisableTiming));
    PCG_CUDA(cudaEventRecord(ix->ready, stream));
  } catch (...) {
    index_free(ix);
    throw;
  }
  return ix;
}

void index_free(Index* ix);
// A copy of a built index on another device (device-to-device over NVLink when the devices are peers, staged
// through the host by the runtime otherwise): the replica every GPU needs for sharded queries and the sharded ICP.
// Much cheaper than a second build, and bit-identical to the source by construction.
Index* index_replicate(const Index& src, int device) {
  Index* ix = new Index();
  ix->device = device;
  ix->n = src.n;
  ix->leaves = src.leaves;
  ix->P = src.P;
  if (src.n == 0) return ix;
  int prev = -1;
  cudaGetDevice(&prev);
  try {
    PCG_CUDA(cudaSetDevice(src.device));
    PCG_CUDA(cudaDeviceSynchronize());  // the source may still be building on some stream
    PCG_CUDA(cudaSetDevice(device));
    const size_t pts_bytes = (size_t)src.leaves * kLeaf * sizeof(float4);
    const size_t box_bytes = (size_t)4 * std::max<uint32_t>(src.P, 2u) * sizeof(float4);
    PCG_CUDA(cudaMalloc((void**)&ix->pts, pts_bytes));
    PCG_CUDA(cudaMalloc((void**)&ix->boxes, box_bytes));
    PCG_CUDA(cudaMalloc((void**)&ix->bbox, 8 * sizeof(uint32_t)));
    PCG_CUDA(cudaMemcpyPeer(ix->pts, device, src.pts, src.device, pts_bytes));
    PCG_CUDA(cudaMemcpyPeer(ix->boxes, device, src.boxes, src.device, box_bytes));
    PCG_CUDA(cudaMemcpyPeer(ix->bbox, device, src.bbox, src.device, 8 * sizeof(uint32_t)));
    ix->bytes = (int64_t)(pts_bytes + box_bytes);
    if (src.inv) {  // DeletePoint was used: the tombstone table travels too
      PCG_CUDA(cudaMalloc((void**)&ix->inv, (size_t)src.n * sizeof(uint32_t)));
      PCG_CUDA(cudaMemcpyPeer(ix->inv, device, src.inv, src.device, (size_t)src.n * sizeof(uint32_t)));
      ix->bytes += src.n * (int64_t)sizeof(uint32_t);
    }
    PCG_CUDA(cudaDeviceSynchronize());
    PCG_CUDA(cudaEventCreateWithFlags(&ix->ready, cudaEventDisableTiming));
    PCG_CUDA(cudaEventRecord(ix->ready, cudaStreamPerThread));
  } catch (...) {
    if (prev >= 0) cudaSetDevice(prev);
    index_free(ix);
    throw;
  }
  if (prev >= 0) cudaSetDevice(prev);
  return ix;
}

void index_free(Index* ix) {
  if (!ix) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ix->device);
  if (ix->pts) cudaFree(ix->pts);
  if (ix->boxes) cudaFree(ix->boxes);
  if (ix->bbox) cudaFree(ix->bbox);
  if (ix->inv) cudaFree(ix->inv);
  if (ix->ready) cudaEventDestroy(ix->ready);
  if (prev >= 0) cudaSetDevice(prev);
  delete ix;
}

// ---- KDTree.DeletePoint (kdtree.go:322-332) as tombstones ------------------------------------
__global__ void __launch_bounds__(256)
    inverse_slots_kernel(const float4* __restrict__ pts, uint32_t padded, uint32_t* __restrict__ inv) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= padded) return;
  const uint32_t id = __float_as_uint(load_point(pts, i).w);
  if (id != 0xffffffffu) inv[id] = i;
}

__global__ void __launch_bounds__(256)
    delete_points_kernel(float4* __restrict__ pts, const uint32_t* __restrict__ inv, const int64_t* __restrict__ ids,
                         int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = inv[ids[i]];
  const float inf = __int_as_float(0x7f800000);
  // x,y,z only: the id in .w stays (deleting twice is a no-op, like the reference's second walk)
  float* p = reinterpret_cast<float*>(pts) + (size_t)(slot >> 3) * 32 + (slot & 7u);  // x, y, z of the slot (bvh.cuh)
  p[0] = inf;
  p[8] = inf;
  p[16] = inf;
}

void index_delete_points_device(Index& ix, const int64_t* d_ids, int64_t n, cudaStream_t stream) {
  if (n == 0 || ix.n == 0) return;
  std::lock_guard<std::mutex> lk(ix.mu);
  ix.wait(stream);
  if (!ix.inv) {
    PCG_CUDA(cudaMallocAsync((void**)&ix.inv, (size_t)ix.n * sizeof(uint32_t), stream));
    ix.bytes += ix.n * (int64_t)sizeof(uint32_t);
    const uint32_t padded = ix.leaves * kLeaf;
    PCG_LAUNCH(inverse_slots_kernel, div_up(padded, 256), 256, 0, stream, ix.pts, padded, ix.inv);
  }
  PCG_LAUNCH(delete_points_kernel, div_up(n, 256), 256, 0, stream, ix.pts, ix.inv, d_ids, n);
  if (ix.ready) PCG_CUDA(cudaEventRecord(ix.ready, stream));  // later queries on other streams see the tombstones
}

// For owners that know the stream the index was last used on (the scan-pair farm): cudaFree synchronises
// the whole device, a stream-ordered free does not.
void index_free_async(Index* ix, cudaStream_t stream) {
  if (!ix) return;
  if (ix->pts) cudaFreeAsync(ix->pts, stream);
  if (ix->boxes) cudaFreeAsync(ix->boxes, stream);
  if (ix->bbox) cudaFreeAsync(ix->bbox, stream);
  if (ix->inv) cudaFreeAsync(ix->inv, stream);
  if (ix->ready) cudaEventDestroy(ix->ready);
  delete ix;
}

#ifdef PCG_NN_STATS
}  // namespace pcg
extern "C" void pcg_debug_nn_stats(unsigned long long out[4]) {
  using namespace pcg;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_nn_stats, sizeof(unsigned long long) * 4);
  unsigned long long z[4] = {0, 0, 0, 0};
  cudaMemcpyToSymbol(g_nn_stats, z, sizeof(z));
}
namespace pcg {
#endif

// ---- query ordering -----------------------------------------------------------------------
#ifndef PCG_QBITS
#define PCG_QBITS 10
#endif
constexpr int kQueryBitsPerAxis = PCG_QBITS;  // 10 -> 30-bit Morton key, four radix passes

__global__ void __launch_bounds__(256)
    query_key_kernel(CloudView q, const uint32_t* __restrict__ bbox6, uint32_t* __restrict__ keys,
                     uint32_t* __restrict__ hist, int passes) {
  __shared__ uint32_t s_hist[rsort::kMaxPasses * rsort::kRadix];
  rsort::hist_zero(s_hist, passes);
  __syncthreads();
  float lo[3], ext = 0.f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    lo[k] = ord_to_float(bbox6[k]);
    ext = fmaxf(ext, ord_to_float(bbox6[3 + k]) - lo[k]);
  }
  const float cells = (float)(1 << kQueryBitsPerAxis);
  const float scale = (ext > 0.f && isfinite(ext)) ? cells / ext : 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (q.n + stride - 1) / stride;
  for (int64_t r = 0; r < rounds; r++) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < q.n;
    uint32_t key = 0;
    if (valid) {
      float3 p = load_xyz(q, i);
      float c[3] = {p.x, p.y, p.z};
      uint32_t u[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        float t = (c[k] - lo[k]) * scale;  // queries outside the box clamp to its faces; NaN -> 0
        u[k] = (uint32_t)fminf(fmaxf(t, 0.f), cells - 1.f);
      }
      key = (uint32_t)hilbert_key<kQueryBitsPerAxis>(u[0], u[1], u[2]);
      keys[i] = key;
    }
    rsort::hist_add_key(s_hist, key, valid, 0, passes);
  }
  __syncthreads();
  rsort::hist_flush(s_hist, hist, passes);
}

void query_order_device(const Index& ix, const CloudView& q, uint32_t* d_perm, cudaStream_t stream) {
  const uint32_t n = (uint32_t)q.n;
  if (n == 0) return;
  DevBuf<uint32_t> keys0(n, stream), keys1(n, stream), vals1(n, stream);
  rsort::Sorter<uint32_t> sorter;
  sorter.prepare(n, 0, 3 * kQueryBitsPerAxis, stream);
  const int blocks = (int)std::min<int64_t>((int64_t)kNumSMs * 4, div_up(n, 256));
  PCG_LAUNCH(query_key_kernel, blocks, 256, 0, stream, q, ix.bbox, keys0.p, sorter.hist(), sorter.passes);
  // the payload ends on side (passes & 1): make d_perm that side so no copy is needed
  uint32_t* kk[2] = {keys0.p, keys1.p};
  uint32_t* vv[2] = {vals1.p, d_perm};
  if ((sorter.passes & 1) == 0) std::swap(vv[0], vv[1]);
  int res = 0;
  sorter.run(kk, vv, /*identity_vals=*/true, /*keep_keys=*/false, stream, &res);
  if (vv[res] != d_perm)
    PCG_CUDA(cudaMemcpyAsync(d_perm, vv[res], (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
}

// ---- KDTree.Nearest, batched (kdtree.go:83-92) -----------------------------------------
// A warp walks `per_thread` consecutive runs of 32 queries of the (Hilbert-ordered) visit list; lane l takes query l
// of every run, so the lanes of a warp always hold neighbouring queries (their walks touch the same nodes).  Queries
// 32 positions apart are still neighbours in space, so the winner of a lane's previous query is a real point close to
// its next one: its distance is a valid upper bound to start the next walk with (it is one of the candidates the
// search would see anyway, so the (DistSq, ID) arg-min is unchanged) and prunes most of the backtracking.
constexpr int kNnThreads = 128;
template <bool APPROX, bool PACKET>
__global__ void __launch_bounds__(kNnThreads)
    nearest_kernel(IndexView ix, CloudView q, const uint32_t* __restrict__ perm, int per_thread, float max_range_sq,
                   float min_dist_sq, int32_t* __restrict__ ids, float* __restrict__ dist_sq,
                   pcg_neighbor* __restrict__ aos) {
  __shared__ unsigned long long s_stack[PACKET ? kNnThreads / 32 : 1][PACKET ? kMaxStack + 8 : 1];
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t first = (t >> 5) * 32 * per_thread + (t & 31);
  const uint64_t init = nn_init(max_range_sq);
  uint32_t warm = 0xffffffffu;
  for (int k = 0; k < per_thread; k++) {
    const int64_t slot = first + (int64_t)k * 32;
    if (PACKET) {
      // the warp walks together: lanes past the end of the list stay in the loop as passengers
      if (first - (t & 31) + (int64_t)k * 32 >= q.n) return;  // warp-uniform
    } else if (slot >= q.n) {
      return;
    }
    const bool live = slot < q.n;
    const int64_t i = live ? (perm ? (int64_t)perm[slot] : slot) : 0;
    const float3 p = live ? load_xyz(q, i) : make_float3(0.f, 0.f, 0.f);
    uint64_t best = init;
    uint32_t pos = 0;
    if (live && warm != 0xffffffffu) {
      const float4 c = load_point(ix.pts, warm);
      const float d = dist_sq_ref(c.x, c.y, c.z, p.x, p.y, p.z);
      const uint64_t packed = ((uint64_t)__float_as_uint(d) << 32) | (uint64_t)__float_as_uint(c.w);
      if (packed < best) {
        best = packed;
        pos = warm;
      }
    }
    if (PACKET) {
      nn_traverse_packet<APPROX>(ix, p.x, p.y, p.z, live, best, pos, s_stack[threadIdx.x >> 5], min_dist_sq);
      __syncwarp();
    } else if (!(APPROX && best != init && __uint_as_float((uint32_t)(best >> 32)) < min_dist_sq)) {
      nn_traverse4<APPROX>(ix, p.x, p.y, p.z, best, pos, min_dist_sq);
    }
    if (!live) continue;
    const bool hit = best != init;
    if (hit) warm = pos;
    const int32_t id = hit ? (int32_t)(uint32_t)best : -1;
    const float d = hit ? __uint_as_float((uint32_t)(best >> 32)) : max_range_sq;
    if (ids) {
      ids[i] = id;
      dist_sq[i] = d;
    }
    if (aos) {
      pcg_neighbor nb;
      nb.id = (int64_t)id;
      nb.dist_sq = d;
      nb.pad_ = 0;
      aos[i] = nb;
    }
  }
}

void nearest_device(const Index& ix, const CloudView& q, float max_range, float min_dist_sq, int32_t* d_ids,
                    float* d_dist_sq, pcg_neighbor* d_aos, cudaStream_t stream) {
  if (q.n == 0) return;
  ix.wait(stream);
  const float mrsq = max_range * max_range;  // kdtree.go:91
  DevBuf<uint32_t> perm;
  if (q.n >= kMinQueriesToReorder && ix.n > 0) {
    perm.alloc((size_t)q.n, stream);
    query_order_device(ix, q, perm.p, stream);
  }
  // queries per thread: as many as still leave every SM a few thousand threads
  const int per_thread = !perm.p ? 1 : (q.n >= (int64_t)kNumSMs * 2048 * 8 ? 8 : (q.n >= (int64_t)kNumSMs * 2048 * 2 ? 4 : 2));
  // blocks cover whole warps of `per_thread` runs of 32 queries
  const int64_t warps = div_up(q.n, 32 * (int64_t)per_thread);
  const int blocks = div_up(warps * 32, kNnThreads);
  // dense batches (several ordered queries per indexed point): neighbouring queries need the same nodes, the warp
  // walks the tree as a packet; sparse ones: one walk per lane
  const bool packet = perm.p != nullptr && q.n >= 2 * ix.n;
  if (min_dist_sq > 0.f) {  // KDTree.MinDistSq > 0: approximate search (kdtree.go:19-22)
    // a miss keeps DistSq == maxRange^2: capping the threshold there means only a real hit can end the search
    // early (the reference's early miss for maxRange^2 < MinDistSq, kdtree.go:100-106, is not reproduced)
    if (packet)
      PCG_LAUNCH((nearest_kernel<true, true>), blocks, kNnThreads, 0, stream, ix.view(), q, perm.p, per_thread, mrsq,
                 fminf(min_dist_sq, mrsq), d_ids, d_dist_sq, d_aos);
    else
      PCG_LAUNCH((nearest_kernel<true, false>), blocks, kNnThreads, 0, stream, ix.view(), q, perm.p, per_thread, mrsq,
                 fminf(min_dist_sq, mrsq), d_ids, d_dist_sq, d_aos);
    return;
  }
  if (packet)
    PCG_LAUNCH((nearest_kernel<false, true>), blocks, kNnThreads, 0, stream, ix.view(), q, perm.p, per_thread, mrsq, 0.f,
               d_ids, d_dist_sq, d_aos);
  else
    PCG_LAUNCH((nearest_kernel<false, false>), blocks, kNnThreads, 0, stream, ix.view(), q, perm.p, per_thread, mrsq, 0.f,
               d_ids, d_dist_sq, d_aos);
}

// ---- KDTree.Range, batched (kdtree.go:148-197) -----------------------------------------
__global__ void __launch_bounds__(128)
    range_count_kernel(IndexView ix, CloudView q, float max_range_sq, uint32_t* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.n) return;
  float3 p = load_xyz(q, i);
  uint32_t c = 0;
  range_traverse(ix, p.x, p.y, p.z, max_range_sq, [&](uint32_t, float) { c++; });
  counts[i] = c;
}

__global__ void __launch_bounds__(128)
    range_fill_kernel(IndexView ix, CloudView q, float max_range_sq, const long long* __restrict__ offsets,
                      unsigned long long* __restrict__ packed, int* __restrict__ mismatch) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.n) return;
  float3 p = load_xyz(q, i);
  // the offsets are the caller's (count pass): never write past the list's end, and report a list whose length
  // differs from what this query finds (other queries / range than in the count pass, stale offsets)
  unsigned long long* dst = packed + offsets[i];
  unsigned long long* const end = packed + offsets[i + 1];
  bool over = false;
  range_traverse(ix, p.x, p.y, p.z, max_range_sq, [&](uint32_t id, float d) {
    if (dst < end)
      *dst++ = ((unsigned long long)__float_as_uint(d) << 32) | id;
    else
      over = true;
  });
  if (over || dst != end) atomicOr(mismatch, 1);
}

// exclusive scan uint32 counts -> int64 offsets (n+1 entries), decoupled look-back
constexpr int kScanItems = 8;
constexpr int kScanTile = 256 * kScanItems;
__global__ void __launch_bounds__(256)
    scan_counts_kernel(const uint32_t* __restrict__ counts, long long* __restrict__ offsets, uint32_t n,
                       uint32_t* __restrict__ tile_counter, unsigned long long* __restrict__ status) {
  __shared__ uint32_t s_scan[rsort::kWarps];
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_prefix;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * kScanTile + tid * kScanItems;
  uint32_t c[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    c[j] = (base + j < n) ? counts[base + j] : 0u;
    sum += c[j];
  }
  uint32_t total = 0;
  uint32_t excl = rsort::block_excl_scan_256(sum, s_scan, &total);
  if (tid == 0) {
    volatile unsigned long long* st = status;
    unsigned long long prefix = 0;
    if (tile == 0) {
      st[0] = (2ull << 62) | (unsigned long long)total;
    } else {
      st[tile] = (1ull << 62) | (unsigned long long)total;
      int64_t prev = (int64_t)tile - 1;
      for (;;) {
        unsigned long long w = st[prev];
        unsigned long long state = w >> 62;
        if (state == 0) continue;
        prefix += w & ((1ull << 62) - 1);
        if (state == 2) break;
        prev--;
      }
      st[tile] = (2ull << 62) | (prefix + total);
    }
    s_prefix = prefix;
    if ((uint64_t)(tile + 1) * kScanTile >= n) offsets[n] = (long long)(prefix + total);
  }
  __syncthreads();
  unsigned long long run = s_prefix + excl;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    if (base + j < n) offsets[base + j] = (long long)run;
    run += c[j];
  }
}

// Per-query sort of packed (DistSq bits << 32 | ID): ascending == (DistSq, ID) order.
// Lists of up to kSortSmem entries are bitonic-sorted in shared memory, longer ones in
// place in global memory by the same CTA.
constexpr int kSortThreads = 128;
constexpr int kSortSmem = 2048;

__device__ __forceinline__ void bitonic_sort(unsigned long long* a, uint32_t len_pow2, uint32_t tid,
                                             uint32_t nthreads) {
  for (uint32_t k = 2; k <= len_pow2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < len_pow2; i += nthreads) {
        uint32_t l = i ^ j;
        if (l > i) {
          unsigned long long x = a[i], y = a[l];
          bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kSortThreads)
    range_sort_kernel(const long long* __restrict__ offsets, unsigned long long* __restrict__ packed,
                      unsigned long long* __restrict__ scratch, const long long* __restrict__ scratch_off) {
  __shared__ unsigned long long s[kSortSmem];
  const long long b = offsets[blockIdx.x], e = offsets[blockIdx.x + 1];
  const uint32_t len = (uint32_t)(e - b);
  if (len <= 1) return;
  uint32_t p2 = 1;
  while (p2 < len) p2 <<= 1;
  unsigned long long* a;
  if (p2 <= kSortSmem) {
    a = s;
  } else {
    a = scratch + scratch_off[blockIdx.x];
  }
  for (uint32_t i = threadIdx.x; i < p2; i += kSortThreads) a[i] = i < len ? packed[b + i] : ~0ull;
  __syncthreads();
  bitonic_sort(a, p2, threadIdx.x, kSortThreads);
  for (uint32_t i = threadIdx.x; i < len; i += kSortThreads) packed[b + i] = a[i];
}

// scratch sizes for lists longer than kSortSmem (rounded up to a power of two), else 0
__global__ void __launch_bounds__(256)
    range_scratch_kernel(const uint32_t* __restrict__ counts, uint32_t* __restrict__ need, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t len = counts[i], p2 = 1;
  while (p2 < len) p2 <<= 1;
  need[i] = p2 > (uint32_t)kSortSmem ? p2 : 0u;
}

__global__ void __launch_bounds__(256)
    range_unpack_kernel(const unsigned long long* __restrict__ packed, pcg_neighbor* __restrict__ out, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  unsigned long long w = packed[i];
  pcg_neighbor nb;
  nb.id = (int64_t)(uint32_t)w;
  nb.dist_sq = __uint_as_float((uint32_t)(w >> 32));
  nb.pad_ = 0;
  out[i] = nb;
}

void scan_counts(const uint32_t* counts, long long* offsets, uint32_t n, cudaStream_t stream) {
  const int tiles = std::max(1, div_up(n, kScanTile));
  DevBuf<unsigned long long> status((size_t)tiles + 1, stream);
  PCG_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), stream));
  PCG_LAUNCH(scan_counts_kernel, tiles, 256, 0, stream, counts, offsets, n, (uint32_t*)(status.p + tiles), status.p);
}

// Returns device CSR: offsets (nq+1) and neighbours (total), sorted per query. Synchronises.
// Count pass: offsets[nq+1] (exclusive scan of the per-query neighbour counts). Synchronises.
void range_count_device(const Index& ix, const CloudView& q, float max_range, DevBuf<long long>& offsets,
                        int64_t* total_out, cudaStream_t stream) {
  const uint32_t nq = (uint32_t)q.n;
  ix.wait(stream);
  const float mrsq = max_range * max_range;  // kdtree.go:157
  offsets.alloc((size_t)nq + 1, stream);
  *total_out = 0;
  if (nq == 0) {
    PCG_CUDA(cudaMemsetAsync(offsets.p, 0, sizeof(long long), stream));
    PCG_CUDA(cudaStreamSynchronize(stream));
    return;
  }
  DevBuf<uint32_t> counts(nq, stream);
  PCG_LAUNCH(range_count_kernel, div_up(nq, 128), 128, 0, stream, ix.view(), q, mrsq, counts.p);
  scan_counts(counts.p, offsets.p, nq, stream);
  long long total = 0;
  PCG_CUDA(cudaMemcpyAsync(&total, offsets.p + nq, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  *total_out = total;
}

__global__ void __launch_bounds__(256)
    range_scratch_from_offsets_kernel(const long long* __restrict__ offsets, uint32_t* __restrict__ need, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t len = (uint32_t)(offsets[i + 1] - offsets[i]), p2 = 1;
  while (p2 < len) p2 <<= 1;
  need[i] = p2 > (uint32_t)kSortSmem ? p2 : 0u;
}

// Fill pass for known offsets: neighbours of every query, each list sorted by (DistSq, ID). Synchronises.
void range_fill_device(const Index& ix, const CloudView& q, float max_range, const long long* d_offsets,
                       int64_t total, pcg_neighbor* d_out, cudaStream_t stream) {
  const uint32_t nq = (uint32_t)q.n;
  if (nq == 0 || total == 0) return;
  ix.wait(stream);
  const float mrsq = max_range * max_range;
  DevBuf<uint32_t> need(nq, stream);
  DevBuf<long long> scratch_off((size_t)nq + 1, stream);
  PCG_LAUNCH(range_scratch_from_offsets_kernel, div_up(nq, 256), 256, 0, stream, d_offsets, need.p, nq);
  scan_counts(need.p, scratch_off.p, nq, stream);
  long long scratch_total = 0;
  PCG_CUDA(cudaMemcpyAsync(&scratch_total, scratch_off.p + nq, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  DevBuf<unsigned long long> packed((size_t)total, stream);
  DevBuf<unsigned long long> scratch((size_t)std::max<long long>(1, scratch_total), stream);
  DevBuf<int> mismatch(1, stream);
  PCG_CUDA(cudaMemsetAsync(mismatch.p, 0, sizeof(int), stream));
  PCG_LAUNCH(range_fill_kernel, div_up(nq, 128), 128, 0, stream, ix.view(), q, mrsq, d_offsets, packed.p, mismatch.p);
  int h_mismatch = 0;
  PCG_CUDA(cudaMemcpyAsync(&h_mismatch, mismatch.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PCG_CUDA(cudaStreamSynchronize(stream));
  if (h_mismatch)
    throw StatusError{PCG_E_INVALID_ARG, "range_fill: the offsets do not match these queries / this range (count pass differs)"};
  PCG_LAUNCH(range_sort_kernel, nq, kSortThreads, 0, stream, d_offsets, packed.p, scratch.p, scratch_off.p);
  PCG_LAUNCH(range_unpack_kernel, div_up(total, 256), 256, 0, stream, packed.p, d_out, (long long)total);
  PCG_CUDA(cudaStreamSynchronize(stream));
}

// Returns device CSR: offsets (nq+1) and neighbours (total), sorted per query. Synchronises.
void range_device(const Index& ix, const CloudView& q, float max_range, DevBuf<long long>& offsets,
                  DevBuf<pcg_neighbor>& out, int64_t* total_out, cudaStream_t stream) {
  range_count_device(ix, q, max_range, offsets, total_out, stream);
  if (*total_out == 0) return;
  out.alloc((size_t)*total_out, stream);
  range_fill_device(ix, q, max_range, offsets.p, *total_out, out.p, stream);
}

}  // namespace pcg
