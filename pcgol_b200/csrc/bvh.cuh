// bvh.cuh — device view of the spatial index and the exact nearest-neighbour / range
// traversal shared by the storage.Search kernels (index.cu) and the fused ICP
// iteration (icp.cu).
//
// Layout in HBM (all built on the GPU, see index.cu):
//   pts   float4[leaves*kLeaf]  points in Morton order: x, y, z, original index (bits);
//                               the tail is padded with (+inf,+inf,+inf, 0xffffffff)
//   boxes float4[2*2*P]         implicit complete binary tree over the leaves, heap
//                               indexed: node k has children 2k and 2k+1, root = 1,
//                               leaf l is node P + l (P = leaves rounded up to 2^m).
//                               Box of node k = {boxes[2k] = lo.xyz, boxes[2k+1] = hi.xyz};
//                               both child boxes of a node are one aligned 64-byte line.
//
// Exactness.  The reference distance is ((dx*dx + dy*dy) + dz*dz) in float32 with
// d = point - query and every operation rounded (mat/vec3.go:18-20,38-40).  The box
// distance below applies the same operations to the per-axis gap between query and
// box; float subtraction, multiplication and addition are monotone, so it never
// exceeds the reference distance of any point inside the box and pruning with it is
// exact.  Candidates are compared as (DistSq bits << 32 | index): non-negative floats
// order like their bit patterns, so one 64-bit min is the (DistSq, ID) lexicographic
// arg-min — the reference's own brute-force oracle (kdtree_test.go:955-968).
#pragma once

#include <mutex>

#include "common.cuh"

namespace pcg {

#ifndef PCG_LEAF
#define PCG_LEAF 8
#endif
constexpr int kLeaf = PCG_LEAF;  // points per leaf (8 * 16 B = one 128-byte line)
constexpr int kMaxStack = 40;   // > log2(2^31 / kLeaf); the 4-ary walk pushes up to 3 per two levels

struct IndexView {
  const float4* pts;
  const float4* boxes;
  uint32_t P;       // leaves rounded up to a power of two (>= 1)
  uint32_t n;       // points
};

#ifdef __CUDACC__
// Packed float32 pairs (sm_100a: FADD2 / FMUL2, one instruction for two IEEE round-to-nearest operations - the same
// per-element results as the scalar instructions).  x and y travel as a pair (they sit in adjacent registers after a
// 16-byte load), z stays scalar.
#ifndef PCG_NO_F32X2
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// ((dx*dx + dy*dy) + dz*dz) with d = a - b: bit-identical to dist_sq_ref (mat/vec3.go:18-20,38-40)
__device__ __forceinline__ float dist_sq_pair(float ax, float ay, float az, unsigned long long bxy, float bz) {
  const unsigned long long d = f2_sub(f2_pack(ax, ay), bxy);
  float sx, sy;
  f2_unpack(f2_mul(d, d), sx, sy);
  const float dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(sx, sy), __fmul_rn(dz, dz));
}
#endif

__device__ __forceinline__ float box_dist_sq(const float4 lo, const float4 hi, float qx, float qy, float qz) {
  // per axis: the query minus its clamp into [lo, hi] (exact selection), ONE rounded subtraction - the same rounded
  // operation the reference applies to (point - query), and |q - clamp(q)| <= |q - p| for every p in the box, so by
  // monotonicity of the rounding the bound never exceeds the reference distance of a point inside.  An empty box
  // (lo = +inf, hi = -inf) clamps to -inf: distance +inf.
#ifndef PCG_NO_F32X2
  const unsigned long long e =
      f2_sub(f2_pack(qx, qy), f2_pack(fminf(fmaxf(qx, lo.x), hi.x), fminf(fmaxf(qy, lo.y), hi.y)));
  float sx, sy;
  f2_unpack(f2_mul(e, e), sx, sy);
  const float ez = __fsub_rn(qz, fminf(fmaxf(qz, lo.z), hi.z));
  return __fadd_rn(__fadd_rn(sx, sy), __fmul_rn(ez, ez));
#else
  const float ex = __fsub_rn(qx, fminf(fmaxf(qx, lo.x), hi.x));
  const float ey = __fsub_rn(qy, fminf(fmaxf(qy, lo.y), hi.y));
  const float ez = __fsub_rn(qz, fminf(fmaxf(qz, lo.z), hi.z));
  return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
#endif
}

__device__ __forceinline__ uint64_t nn_init(float max_range_sq) {
  // strict "<" against maxRange^2 with ID -1: nothing packs below (bits, 0) at equal distance
  return (uint64_t)__float_as_uint(max_range_sq) << 32;
}

// Exact nearest neighbour: nn_traverse4 below.  `best` must be nn_init(maxRange^2) or a tighter bound that belongs
// to a real point (a warm start); on return it is unchanged for a miss, else (DistSq bits << 32 | original index)
// and best_pos is the winner's position in ix.pts.
// 4-ary view of the same heap: a step looks at the four grandchildren 4k..4k+3 of node k (their
// boxes are one aligned 128-byte line), so the chain of dependent loads per descent is half as
// long.  Children are ordered by their box distance with the child slot packed into the two
// low mantissa bits; clearing those bits again only lowers the bound, so pruning stays exact.
#ifdef PCG_NN_STATS
__device__ unsigned long long g_nn_stats[4];  // node steps, leaf scans, pushes, queries (tuning builds only)
#define PCG_STAT(i) atomicAdd(&g_nn_stats[i], 1ull)
#else
#define PCG_STAT(i) \
  do {              \
  } while (0)
#endif

// APPROX mirrors KDTree.MinDistSq (kdtree.go:19-22,104,120,140): the search stops at the first
// candidate with DistSq < min_dist_sq.  The answer is then either the exact nearest neighbour or a
// real point closer than sqrt(MinDistSq) - the contract every MinDistSq answer of the reference
// satisfies (which point is traversal-dependent there too).
template <bool APPROX = false>
__device__ __forceinline__ void nn_traverse4(const IndexView& ix, float qx, float qy, float qz, uint64_t& best,
                                             uint32_t& best_pos, float min_dist_sq = 0.f) {
  if (ix.n == 0) return;
  PCG_STAT(3);
  if (qx != qx || qy != qy || qz != qz) return;
  // one 64-bit word per deferred child: (box-distance bits << 32 | node id) -> one local-memory access per
  // push and per pop
  unsigned long long stack[kMaxStack + 8];
  int sp = 0;
  float bestd = __uint_as_float((uint32_t)(best >> 32));
#ifndef PCG_NO_F32X2
  const unsigned long long qxy = f2_pack(qx, qy);
#endif
  {
    const float d = box_dist_sq(__ldg(ix.boxes + 2), __ldg(ix.boxes + 3), qx, qy, qz);
    if (!(d <= bestd)) return;
  }
  const uint32_t P = ix.P;
  uint32_t node = 1;
  // leaves sit at depth log2(P); with an odd depth the first step is binary so that 4-ary steps land on them
  if (P > 1 && (__ffs(P) - 1) & 1) {
    const float4* cb = ix.boxes + 4;
    const float d0 = box_dist_sq(__ldg(cb), __ldg(cb + 1), qx, qy, qz);
    const float d1 = box_dist_sq(__ldg(cb + 2), __ldg(cb + 3), qx, qy, qz);
    const bool first0 = d0 <= d1;
    const float dn = first0 ? d0 : d1, df = first0 ? d1 : d0;
    if (!(dn <= bestd)) return;
    if (df <= bestd) stack[sp++] = ((unsigned long long)__float_as_uint(df) << 32) | (first0 ? 3u : 2u);
    node = first0 ? 2u : 3u;
  }
  for (;;) {
    while (node < P) {
      const float4* cb = ix.boxes + 8 * (size_t)node;  // boxes of nodes 4*node .. 4*node+3
      PCG_STAT(0);
      uint32_t key[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float d = box_dist_sq(__ldg(cb + 2 * j), __ldg(cb + 2 * j + 1), qx, qy, qz);
        key[j] = (__float_as_uint(d) & ~3u) | (uint32_t)j;  // d >= 0 (or +inf / NaN pattern for empty boxes)
      }
      // sorting network for 4 keys (ascending)
#define PCG_CSWAP(a, b)                 \
  {                                     \
    const uint32_t lo_ = min(a, b);     \
    b = max(a, b);                      \
    a = lo_;                            \
  }
      PCG_CSWAP(key[0], key[1]);
      PCG_CSWAP(key[2], key[3]);
      PCG_CSWAP(key[0], key[2]);
      PCG_CSWAP(key[1], key[3]);
      PCG_CSWAP(key[1], key[2]);
#undef PCG_CSWAP
      const float dn = __uint_as_float(key[0] & ~3u);
      if (!(dn <= bestd)) {
        node = 0;
        break;
      }
#pragma unroll
      for (int j = 3; j >= 1; j--) {  // farthest first: the nearest of them is popped first
        const float dj = __uint_as_float(key[j] & ~3u);
        if (dj <= bestd) {
          PCG_STAT(2);
          stack[sp++] = ((unsigned long long)(key[j] & ~3u) << 32) | (4 * node + (key[j] & 3u));
        }
      }
      node = 4 * node + (key[0] & 3u);
    }
    if (node) {
      PCG_STAT(1);
      const uint32_t base = (node - P) * kLeaf;
      const float4* lp = ix.pts + base;
#pragma unroll
      for (int j = 0; j < kLeaf; j++) {
        const float4 p = __ldg(lp + j);
#ifndef PCG_NO_F32X2
        const float d = dist_sq_pair(p.x, p.y, p.z, qxy, qz);
#else
        const float d = dist_sq_ref(p.x, p.y, p.z, qx, qy, qz);
#endif
        const uint64_t packed = ((uint64_t)__float_as_uint(d) << 32) | (uint64_t)__float_as_uint(p.w);
        if (packed < best) {
          best = packed;
          best_pos = base + j;
        }
      }
      bestd = __uint_as_float((uint32_t)(best >> 32));
      if (APPROX && bestd < min_dist_sq) return;
    }
    node = 0;
    while (sp > 0) {
      const unsigned long long e = stack[--sp];
      if (__uint_as_float((uint32_t)(e >> 32)) <= bestd) {
        node = (uint32_t)e;
        break;
      }
    }
    if (!node) break;
  }
}

// Warp-packet form of nn_traverse4 for DENSE query batches: the 32 lanes of a warp hold 32 neighbouring queries (the
// visit list is Hilbert-ordered) and walk the tree TOGETHER - a node is entered when ANY lane still needs it, every lane
// then tests the node's four children (or the leaf's eight points) against its own query.  Control flow is uniform, so
// all 32 lanes stay active where the per-lane walk diverges (about 11 of 32 active on config 3); the price is that a
// lane also steps through nodes only its neighbours need, which pays off exactly when the lanes' search balls overlap
// (many more queries than points).  Exactness is that of the per-lane walk: every lane compares its own exact box
// and point distances; the shared decisions only ever ADD nodes (a child is entered if any lane's box distance is
// within that lane's bound; a deferred child is dropped only if its smallest box distance over the warp exceeds
// every lane's bound).  `stack` = kMaxStack + 8 words of shared memory owned by the warp.
template <bool APPROX = false>
__device__ __forceinline__ void nn_traverse_packet(const IndexView& ix, float qx, float qy, float qz, bool active,
                                                   uint64_t& best, uint32_t& best_pos, unsigned long long* stack,
                                                   float min_dist_sq = 0.f) {
  constexpr unsigned kFull = 0xffffffffu;
  const float inf = __int_as_float(0x7f800000);
  active = active && ix.n != 0 && qx == qx && qy == qy && qz == qz;
  float bestd = __uint_as_float((uint32_t)(best >> 32));
  if (APPROX && active && bestd < min_dist_sq) active = false;
  const uint32_t lane = threadIdx.x & 31;
  int sp = 0;
#ifndef PCG_NO_F32X2
  const unsigned long long qxy = f2_pack(qx, qy);
#endif
  {
    const float d = active ? box_dist_sq(__ldg(ix.boxes + 2), __ldg(ix.boxes + 3), qx, qy, qz) : inf;
    if (__ballot_sync(kFull, active && d <= bestd) == 0) return;
  }
  const uint32_t P = ix.P;
  uint32_t node = 1;
  if (P > 1 && (__ffs(P) - 1) & 1) {  // odd depth: one binary step first (see nn_traverse4)
    const float4* cb = ix.boxes + 4;
    const float d0 = active ? box_dist_sq(__ldg(cb), __ldg(cb + 1), qx, qy, qz) : inf;
    const float d1 = active ? box_dist_sq(__ldg(cb + 2), __ldg(cb + 3), qx, qy, qz) : inf;
    const uint32_t m0 = __reduce_min_sync(kFull, __float_as_uint(d0)), m1 = __reduce_min_sync(kFull, __float_as_uint(d1));
    const bool w0 = __ballot_sync(kFull, d0 <= bestd) != 0, w1 = __ballot_sync(kFull, d1 <= bestd) != 0;
    if (!w0 && !w1) return;
    const bool first0 = w0 && (!w1 || m0 <= m1);
    if (w0 && w1 && lane == 0) stack[0] = ((unsigned long long)(first0 ? m1 : m0) << 32) | (first0 ? 3u : 2u);
    if (w0 && w1) sp = 1;
    node = first0 ? 2u : 3u;
  }
  for (;;) {
    while (node < P) {
      const float4* cb = ix.boxes + 8 * (size_t)node;  // boxes of nodes 4*node .. 4*node+3
      uint32_t key[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float d = active ? box_dist_sq(__ldg(cb + 2 * j), __ldg(cb + 2 * j + 1), qx, qy, qz) : inf;
        const uint32_t m = __reduce_min_sync(kFull, __float_as_uint(d));  // d >= 0: the bits order like the values
        const bool want = __ballot_sync(kFull, d <= bestd) != 0;
        key[j] = want ? ((m & ~3u) | (uint32_t)j) : 0xffffffffu;
      }
#define PCG_CSWAP(a, b)                 \
  {                                     \
    const uint32_t lo_ = min(a, b);     \
    b = max(a, b);                      \
    a = lo_;                            \
  }
      PCG_CSWAP(key[0], key[1]);
      PCG_CSWAP(key[2], key[3]);
      PCG_CSWAP(key[0], key[2]);
      PCG_CSWAP(key[1], key[3]);
      PCG_CSWAP(key[1], key[2]);
#undef PCG_CSWAP
      if (key[0] == 0xffffffffu) {
        node = 0;
        break;
      }
#pragma unroll
      for (int j = 3; j >= 1; j--) {  // farthest first: the nearest of them is popped first
        if (key[j] != 0xffffffffu) {
          if (lane == 0) stack[sp] = ((unsigned long long)(key[j] & ~3u) << 32) | (4 * node + (key[j] & 3u));
          sp++;
        }
      }
      node = 4 * node + (key[0] & 3u);
    }
    if (node) {
      const uint32_t base = (node - P) * kLeaf;
      const float4* lp = ix.pts + base;
      if (active) {
#pragma unroll
        for (int j = 0; j < kLeaf; j++) {
          const float4 p = __ldg(lp + j);
#ifndef PCG_NO_F32X2
          const float d = dist_sq_pair(p.x, p.y, p.z, qxy, qz);
#else
          const float d = dist_sq_ref(p.x, p.y, p.z, qx, qy, qz);
#endif
          const uint64_t packed = ((uint64_t)__float_as_uint(d) << 32) | (uint64_t)__float_as_uint(p.w);
          if (packed < best) {
            best = packed;
            best_pos = base + j;
          }
        }
        bestd = __uint_as_float((uint32_t)(best >> 32));
        if (APPROX && bestd < min_dist_sq) active = false;
      }
    }
    __syncwarp();  // lane 0's pushes are visible to the warp
    // the largest bound any lane still holds: a deferred child whose smallest box distance exceeds it is dead
    const uint32_t wmax = __reduce_max_sync(kFull, active ? __float_as_uint(bestd) : 0u);
    if (__ballot_sync(kFull, active) == 0) return;
    node = 0;
    while (sp > 0) {
      const unsigned long long e = stack[--sp];
      if ((uint32_t)(e >> 32) <= wmax) {
        node = (uint32_t)e;
        break;
      }
    }
    if (!node) break;
  }
}

// Visits every point with DistSq < max_range_sq (strict, kdtree.go:167,179) and calls
// f(original index, DistSq).
template <typename F>
__device__ __forceinline__ void range_traverse(const IndexView& ix, float qx, float qy, float qz, float max_range_sq,
                                               F f) {
  if (ix.n == 0) return;
  if (qx != qx || qy != qy || qz != qz) return;
  uint32_t stack_node[kMaxStack];
  int sp = 0;
  {
    float d = box_dist_sq(ix.boxes[2], ix.boxes[3], qx, qy, qz);
    if (!(d < max_range_sq)) return;
  }
  const uint32_t P = ix.P;
  uint32_t node = 1;
  for (;;) {
    while (node < P) {
      const float4* cb = ix.boxes + 4 * (size_t)node;
      const float4 l0 = __ldg(cb), h0 = __ldg(cb + 1), l1 = __ldg(cb + 2), h1 = __ldg(cb + 3);
      const bool in0 = box_dist_sq(l0, h0, qx, qy, qz) < max_range_sq;
      const bool in1 = box_dist_sq(l1, h1, qx, qy, qz) < max_range_sq;
      if (in0 && in1) {
        stack_node[sp++] = 2 * node + 1;
        node = 2 * node;
      } else if (in0) {
        node = 2 * node;
      } else if (in1) {
        node = 2 * node + 1;
      } else {
        node = 0;
        break;
      }
    }
    if (node) {
      const float4* lp = ix.pts + (size_t)(node - P) * kLeaf;
#pragma unroll
      for (int j = 0; j < kLeaf; j++) {
        const float4 p = __ldg(lp + j);
        const float d = dist_sq_ref(p.x, p.y, p.z, qx, qy, qz);
        if (d < max_range_sq) f(__float_as_uint(p.w), d);
      }
    }
    if (sp == 0) break;
    node = stack_node[--sp];
  }
}
#endif

// Host-side handle behind the opaque pcg_index.
struct Index {
  int device = 0;
  int64_t n = 0;
  uint32_t leaves = 0;
  uint32_t P = 1;
  float4* pts = nullptr;
  float4* boxes = nullptr;
  uint32_t* bbox = nullptr;  // order-preserving bits of min x,y,z / max x,y,z over finite coordinates
  // KDTree.DeletePoint (kdtree.go:322-332): tombstones.  inv maps an original id to its slot in pts
  // (built on the first delete); a deleted slot keeps its id but its coordinates become +inf, so no
  // search can ever accept it (the boxes stay valid: they only get conservative).
  uint32_t* inv = nullptr;
  std::mutex mu;  // serialises DeletePoint calls (queries racing a delete are undefined, as in the reference)
  int64_t bytes = 0;
  // recorded on the stream the index was built (or last modified) on: every consumer makes its own stream wait for
  // it, so a build enqueued on one stream is ordered before queries enqueued on another
  cudaEvent_t ready = nullptr;
  void wait(cudaStream_t s) const {
    if (ready) cudaStreamWaitEvent(s, ready, 0);
  }
  IndexView view() const { return IndexView{pts, boxes, P, (uint32_t)n}; }
};

Index* index_build_device(const CloudView& v, int device, cudaStream_t stream);
// d_ids: n point ids, already validated against [0, ix.n). Enqueues on `stream`.
void index_delete_points_device(Index& ix, const int64_t* d_ids, int64_t n, cudaStream_t stream);
void index_free(Index* ix);
Index* index_replicate(const Index& src, int device);
void index_free_async(Index* ix, cudaStream_t stream);  // caller is on ix->device; `stream` ordered after the last use
// Queries visited in Morton order make the threads of a warp walk the same part of the tree.
// Writes a permutation (sorted position -> query index) into d_perm[q.n].
void query_order_device(const Index& ix, const CloudView& q, uint32_t* d_perm, cudaStream_t stream);
constexpr int64_t kMinQueriesToReorder = 1 << 14;

// Batched KDTree.Range building blocks (index.cu); both synchronise `stream`.
void range_count_device(const Index& ix, const CloudView& q, float max_range, DevBuf<long long>& offsets,
                        int64_t* total_out, cudaStream_t stream);
void range_fill_device(const Index& ix, const CloudView& q, float max_range, const long long* d_offsets,
                       int64_t total, pcg_neighbor* d_out, cudaStream_t stream);
// exclusive scan of uint32 counts into int64 offsets[n+1] (decoupled look-back)
void scan_counts(const uint32_t* counts, long long* offsets, uint32_t n, cudaStream_t stream);

}  // namespace pcg
