// bvh.cuh — device view of the spatial index and the exact nearest-neighbour / range
// traversal shared by the storage.Search kernels (index.cu) and the fused ICP
// iteration (icp.cu).
//
// Layout in HBM (all built on the GPU, see index.cu):
//   pts   float4[leaves*kLeaf]  points in KD order, one 128-byte line per leaf of 8 points, coordinate-major INSIDE
//                               the leaf: x[8] y[8] z[8] id[8] (original index bits); the tail is padded with
//                               (+inf,+inf,+inf, 0xffffffff).  load_point() returns slot pos as {x, y, z, id}.
//   boxes float4[2*2*P]         implicit complete binary tree over the leaves, heap indexed: node k has children 2k
//                               and 2k+1, root = 1, leaf l is node P + l (P = leaves rounded up to 2^m).  The boxes
//                               of the four nodes 4g .. 4g+3 (the children of a 4-ary step, two sibling pairs of the
//                               binary tree) are one 128-byte line, coordinate-major: lo.x[4] lo.y[4] lo.z[4]
//                               hi.x[4] hi.y[4] hi.z[4] (+ 8 floats of padding).  load_box() / store_box() address
//                               a single node.
// Coordinate-major lines put the same coordinate of two neighbouring boxes / points in adjacent registers after a
// 16-byte load, which is what the packed float32 pair instructions of sm_100a (FADD2 / FMUL2) take: a box or point
// test is 8 packed instructions per PAIR instead of 8 scalar ones per item.
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
// per-element results as the scalar instructions).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// ((x*x + y*y) + z*z) for both halves.  The squares are packed; the sums are scalar on purpose: ptxas contracts a packed
// multiply feeding a packed add into FFMA2 (even with explicit .rn and -fmad=false), and the reference never fuses.
__device__ __forceinline__ void sum_squares_x2(f32x2 x, f32x2 y, f32x2 z, float& a, float& b) {
  float xa, xb, ya, yb, za, zb;
  f2_unpack(f2_mul(x, x), xa, xb);
  f2_unpack(f2_mul(y, y), ya, yb);
  f2_unpack(f2_mul(z, z), za, zb);
  a = __fadd_rn(__fadd_rn(xa, ya), za);
  b = __fadd_rn(__fadd_rn(xb, yb), zb);
}
// The query as three pairs {q, q}.
struct Query2 {
  f32x2 x, y, z;
  __device__ __forceinline__ Query2(float qx, float qy, float qz)
      : x(f2_pack(qx, qx)), y(f2_pack(qy, qy)), z(f2_pack(qz, qz)) {}
};
// ((dx*dx + dy*dy) + dz*dz), d = point - query, for TWO points: bit-identical to dist_sq_ref (mat/vec3.go:18-20,38-40)
__device__ __forceinline__ void dist_sq_x2(float xa, float xb, float ya, float yb, float za, float zb, const Query2& q,
                                           float& da, float& db) {
  const f32x2 dx = f2_sub(f2_pack(xa, xb), q.x), dy = f2_sub(f2_pack(ya, yb), q.y), dz = f2_sub(f2_pack(za, zb), q.z);
  sum_squares_x2(dx, dy, dz, da, db);
}

__device__ __forceinline__ float box_dist_sq(const float lo[3], const float hi[3], float qx, float qy, float qz) {
  // per axis: the query minus its clamp into [lo, hi] (exact selection), ONE rounded subtraction - the same rounded
  // operation the reference applies to (point - query), and |q - clamp(q)| <= |q - p| for every p in the box, so by
  // monotonicity of the rounding the bound never exceeds the reference distance of a point inside.  An empty box
  // (lo = +inf, hi = -inf) clamps to -inf: distance +inf.
  const float ex = __fsub_rn(qx, fminf(fmaxf(qx, lo[0]), hi[0]));
  const float ey = __fsub_rn(qy, fminf(fmaxf(qy, lo[1]), hi[1]));
  const float ez = __fsub_rn(qz, fminf(fmaxf(qz, lo[2]), hi[2]));
  return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}
// The same bound for two boxes at once (coordinates given per box a / b).
__device__ __forceinline__ void box_dist_sq_x2(float lxa, float lxb, float lya, float lyb, float lza, float lzb, float hxa,
                                               float hxb, float hya, float hyb, float hza, float hzb, float qx, float qy,
                                               float qz, const Query2& q, float& da, float& db) {
  const f32x2 ex = f2_sub(q.x, f2_pack(fminf(fmaxf(qx, lxa), hxa), fminf(fmaxf(qx, lxb), hxb)));
  const f32x2 ey = f2_sub(q.y, f2_pack(fminf(fmaxf(qy, lya), hya), fminf(fmaxf(qy, lyb), hyb)));
  const f32x2 ez = f2_sub(q.z, f2_pack(fminf(fmaxf(qz, lza), hza), fminf(fmaxf(qz, lzb), hzb)));
  sum_squares_x2(ex, ey, ez, da, db);
}

// ---- addressing of the coordinate-major lines ---------------------------------------------------------------------
static_assert(kLeaf == 8, "a leaf is one 128-byte line: x[8] y[8] z[8] id[8]");
__device__ __forceinline__ float4 load_point(const float4* pts, uint32_t pos) {
  const float* f = reinterpret_cast<const float*>(pts) + (size_t)(pos >> 3) * 32 + (pos & 7u);
  return make_float4(__ldg(f), __ldg(f + 8), __ldg(f + 16), __ldg(f + 24));
}
__device__ __forceinline__ void store_point(float4* pts, uint32_t pos, float4 v) {
  float* f = reinterpret_cast<float*>(pts) + (size_t)(pos >> 3) * 32 + (pos & 7u);
  f[0] = v.x;
  f[8] = v.y;
  f[16] = v.z;
  f[24] = v.w;
}
__device__ __forceinline__ void load_box(const float4* boxes, uint32_t node, float lo[3], float hi[3]) {
  const float* f = reinterpret_cast<const float*>(boxes) + (size_t)(node >> 2) * 32 + (node & 3u);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    lo[k] = __ldg(f + 4 * k);
    hi[k] = __ldg(f + 12 + 4 * k);
  }
}
__device__ __forceinline__ void store_box(float4* boxes, uint32_t node, const float lo[3], const float hi[3]) {
  float* f = reinterpret_cast<float*>(boxes) + (size_t)(node >> 2) * 32 + (node & 3u);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    f[4 * k] = lo[k];
    f[12 + 4 * k] = hi[k];
  }
}
// The four boxes of nodes 4g .. 4g+3 (six 16-byte loads) and their bounds for one query.
struct BoxLine {
  float4 lx, ly, lz, hx, hy, hz;
  __device__ __forceinline__ BoxLine(const float4* boxes, uint32_t group) {
    const float4* g = boxes + 8 * (size_t)group;
    lx = __ldg(g), ly = __ldg(g + 1), lz = __ldg(g + 2), hx = __ldg(g + 3), hy = __ldg(g + 4), hz = __ldg(g + 5);
  }
  // nodes 4g and 4g+1 / 4g+2 and 4g+3
  __device__ __forceinline__ void dist_low(float qx, float qy, float qz, const Query2& q, float& d0, float& d1) const {
    box_dist_sq_x2(lx.x, lx.y, ly.x, ly.y, lz.x, lz.y, hx.x, hx.y, hy.x, hy.y, hz.x, hz.y, qx, qy, qz, q, d0, d1);
  }
  __device__ __forceinline__ void dist_high(float qx, float qy, float qz, const Query2& q, float& d2, float& d3) const {
    box_dist_sq_x2(lx.z, lx.w, ly.z, ly.w, lz.z, lz.w, hx.z, hx.w, hy.z, hy.w, hz.z, hz.w, qx, qy, qz, q, d2, d3);
  }
};
// The eight points of a leaf (eight 16-byte loads): squared distances to one query and the ids.
struct LeafLine {
  float4 xa, xb, ya, yb, za, zb, ia, ib;
  __device__ __forceinline__ explicit LeafLine(const float4* lp) {
    xa = __ldg(lp), xb = __ldg(lp + 1), ya = __ldg(lp + 2), yb = __ldg(lp + 3);
    za = __ldg(lp + 4), zb = __ldg(lp + 5), ia = __ldg(lp + 6), ib = __ldg(lp + 7);
  }
  __device__ __forceinline__ void dist8(const Query2& q, float d[8]) const {
    dist_sq_x2(xa.x, xa.y, ya.x, ya.y, za.x, za.y, q, d[0], d[1]);
    dist_sq_x2(xa.z, xa.w, ya.z, ya.w, za.z, za.w, q, d[2], d[3]);
    dist_sq_x2(xb.x, xb.y, yb.x, yb.y, zb.x, zb.y, q, d[4], d[5]);
    dist_sq_x2(xb.z, xb.w, yb.z, yb.w, zb.z, zb.w, q, d[6], d[7]);
  }
  __device__ __forceinline__ void ids(uint32_t id[8]) const {
    id[0] = __float_as_uint(ia.x), id[1] = __float_as_uint(ia.y), id[2] = __float_as_uint(ia.z), id[3] = __float_as_uint(ia.w);
    id[4] = __float_as_uint(ib.x), id[5] = __float_as_uint(ib.y), id[6] = __float_as_uint(ib.z), id[7] = __float_as_uint(ib.w);
  }
};

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
  const Query2 q2(qx, qy, qz);
  const uint32_t P = ix.P;
  uint32_t node = 1;
  {
    // nodes 0..3 share a line: the root's box and, for an odd depth, the first (binary) step so that the 4-ary
    // steps land on the leaves
    const BoxLine top(ix.boxes, 0);
    float dx, d1, d2, d3;
    top.dist_low(qx, qy, qz, q2, dx, d1);
    if (!(d1 <= bestd)) return;
    if (P > 1 && (__ffs(P) - 1) & 1) {
      top.dist_high(qx, qy, qz, q2, d2, d3);
      const bool first0 = d2 <= d3;
      const float dn = first0 ? d2 : d3, df = first0 ? d3 : d2;
      if (!(dn <= bestd)) return;
      if (df <= bestd) stack[sp++] = ((unsigned long long)__float_as_uint(df) << 32) | (first0 ? 3u : 2u);
      node = first0 ? 2u : 3u;
    }
  }
  for (;;) {
    while (node < P) {
      const BoxLine cb(ix.boxes, node);  // boxes of nodes 4*node .. 4*node+3
      PCG_STAT(0);
      float d[4];
      cb.dist_low(qx, qy, qz, q2, d[0], d[1]);
      cb.dist_high(qx, qy, qz, q2, d[2], d[3]);
      uint32_t key[4];
#pragma unroll
      for (int j = 0; j < 4; j++)
        key[j] = (__float_as_uint(d[j]) & ~3u) | (uint32_t)j;  // d >= 0 (or +inf / NaN pattern for empty boxes)
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
      const LeafLine leaf(ix.pts + base);
      float d[8];
      uint32_t id[8];
      leaf.dist8(q2, d);
      leaf.ids(id);
#pragma unroll
      for (int j = 0; j < kLeaf; j++) {
        const uint64_t packed = ((uint64_t)__float_as_uint(d[j]) << 32) | (uint64_t)id[j];
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
  const Query2 q2(qx, qy, qz);
  const uint32_t P = ix.P;
  uint32_t node = 1;
  {
    const BoxLine top(ix.boxes, 0);  // nodes 0..3: the root and, for an odd depth, the first (binary) step
    float dx, dr, d0, d1;
    top.dist_low(qx, qy, qz, q2, dx, dr);
    if (!active) dr = inf;
    if (__ballot_sync(kFull, active && dr <= bestd) == 0) return;
    if (P > 1 && (__ffs(P) - 1) & 1) {
      top.dist_high(qx, qy, qz, q2, d0, d1);
      if (!active) d0 = d1 = inf;
      const uint32_t m0 = __reduce_min_sync(kFull, __float_as_uint(d0)), m1 = __reduce_min_sync(kFull, __float_as_uint(d1));
      const bool w0 = __ballot_sync(kFull, d0 <= bestd) != 0, w1 = __ballot_sync(kFull, d1 <= bestd) != 0;
      if (!w0 && !w1) return;
      const bool first0 = w0 && (!w1 || m0 <= m1);
      if (w0 && w1 && lane == 0) stack[0] = ((unsigned long long)(first0 ? m1 : m0) << 32) | (first0 ? 3u : 2u);
      if (w0 && w1) sp = 1;
      node = first0 ? 2u : 3u;
    }
  }
  for (;;) {
    while (node < P) {
      const BoxLine cb(ix.boxes, node);  // boxes of nodes 4*node .. 4*node+3
      float d[4];
      cb.dist_low(qx, qy, qz, q2, d[0], d[1]);
      cb.dist_high(qx, qy, qz, q2, d[2], d[3]);
      uint32_t key[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float dj = active ? d[j] : inf;
        const uint32_t m = __reduce_min_sync(kFull, __float_as_uint(dj));  // d >= 0: the bits order like the values
        const bool want = __ballot_sync(kFull, dj <= bestd) != 0;
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
      const LeafLine leaf(ix.pts + base);
      if (active) {
        float d[8];
        uint32_t id[8];
        leaf.dist8(q2, d);
        leaf.ids(id);
#pragma unroll
        for (int j = 0; j < kLeaf; j++) {
          const uint64_t packed = ((uint64_t)__float_as_uint(d[j]) << 32) | (uint64_t)id[j];
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
  const Query2 q2(qx, qy, qz);
  {
    float lo[3], hi[3];
    load_box(ix.boxes, 1, lo, hi);
    const float d = box_dist_sq(lo, hi, qx, qy, qz);
    if (!(d < max_range_sq)) return;
  }
  const uint32_t P = ix.P;
  uint32_t node = 1;
  for (;;) {
    while (node < P) {
      // the children 2*node, 2*node+1 are one half of the line of nodes 4g .. 4g+3
      const float* g = reinterpret_cast<const float*>(ix.boxes) + (size_t)(node >> 1) * 32 + ((2 * node) & 3u);
      float d0, d1;
      {
        const float2 lx = __ldg(reinterpret_cast<const float2*>(g)), ly = __ldg(reinterpret_cast<const float2*>(g + 4)),
                     lz = __ldg(reinterpret_cast<const float2*>(g + 8)), hx = __ldg(reinterpret_cast<const float2*>(g + 12)),
                     hy = __ldg(reinterpret_cast<const float2*>(g + 16)), hz = __ldg(reinterpret_cast<const float2*>(g + 20));
        box_dist_sq_x2(lx.x, lx.y, ly.x, ly.y, lz.x, lz.y, hx.x, hx.y, hy.x, hy.y, hz.x, hz.y, qx, qy, qz, q2, d0, d1);
      }
      const bool in0 = d0 < max_range_sq;
      const bool in1 = d1 < max_range_sq;
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
      const LeafLine leaf(ix.pts + (size_t)(node - P) * kLeaf);
      float d[8];
      uint32_t id[8];
      leaf.dist8(q2, d);
      leaf.ids(id);
#pragma unroll
      for (int j = 0; j < kLeaf; j++)
        if (d[j] < max_range_sq) f(id[j], d[j]);
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
