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

#include "common.cuh"

namespace pcg {

constexpr int kLeaf = 8;        // points per leaf: 8 * 16 B = one 128-byte line
constexpr int kMaxStack = 40;   // > log2(2^31 / kLeaf)

struct IndexView {
  const float4* pts;
  const float4* boxes;
  uint32_t P;       // leaves rounded up to a power of two (>= 1)
  uint32_t n;       // points
};

#ifdef __CUDACC__
__device__ __forceinline__ float box_dist_sq(const float4 lo, const float4 hi, float qx, float qy, float qz) {
  float ex = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.0f);
  float ey = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.0f);
  float ez = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.0f);
  return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}

__device__ __forceinline__ uint64_t nn_init(float max_range_sq) {
  // strict "<" against maxRange^2 with ID -1: nothing packs below (bits, 0) at equal distance
  return (uint64_t)__float_as_uint(max_range_sq) << 32;
}

// Exact nearest neighbour. `best` must be nn_init(maxRange^2) (or a tighter bound);
// on return it is unchanged for a miss, else (DistSq bits << 32 | original index) and
// best_pos is the winner's position in ix.pts.
__device__ __forceinline__ void nn_traverse(const IndexView& ix, float qx, float qy, float qz, uint64_t& best,
                                            uint32_t& best_pos) {
  if (ix.n == 0) return;
  if (qx != qx || qy != qy || qz != qz) return;  // NaN never compares below the bound in the reference
  uint32_t stack_node[kMaxStack];
  float stack_d[kMaxStack];
  int sp = 0;
  float bestd = __uint_as_float((uint32_t)(best >> 32));
  {
    float d = box_dist_sq(ix.boxes[2], ix.boxes[3], qx, qy, qz);
    if (!(d <= bestd)) return;
  }
  const uint32_t P = ix.P;
  uint32_t node = 1;
  for (;;) {
    while (node < P) {
      const float4* cb = ix.boxes + 4 * (size_t)node;
      const float4 l0 = __ldg(cb), h0 = __ldg(cb + 1), l1 = __ldg(cb + 2), h1 = __ldg(cb + 3);
      const float d0 = box_dist_sq(l0, h0, qx, qy, qz);
      const float d1 = box_dist_sq(l1, h1, qx, qy, qz);
      const bool first0 = d0 <= d1;
      const float dn = first0 ? d0 : d1, df = first0 ? d1 : d0;
      const uint32_t near_node = 2 * node + (first0 ? 0u : 1u), far_node = 2 * node + (first0 ? 1u : 0u);
      if (!(dn <= bestd)) {
        node = 0;
        break;
      }
      if (df <= bestd) {
        stack_node[sp] = far_node;
        stack_d[sp] = df;
        sp++;
      }
      node = near_node;
    }
    if (node) {
      const uint32_t base = (node - P) * kLeaf;
      const float4* lp = ix.pts + base;
#pragma unroll
      for (int j = 0; j < kLeaf; j++) {
        const float4 p = __ldg(lp + j);
        const float d = dist_sq_ref(p.x, p.y, p.z, qx, qy, qz);
        const uint64_t packed = ((uint64_t)__float_as_uint(d) << 32) | (uint64_t)__float_as_uint(p.w);
        if (packed < best) {
          best = packed;
          best_pos = base + j;
        }
      }
      bestd = __uint_as_float((uint32_t)(best >> 32));
    }
    node = 0;
    while (sp > 0) {
      --sp;
      if (stack_d[sp] <= bestd) {
        node = stack_node[sp];
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
  int64_t bytes = 0;
  IndexView view() const { return IndexView{pts, boxes, P, (uint32_t)n}; }
};

Index* index_build_device(const CloudView& v, int device, cudaStream_t stream);
void index_free(Index* ix);

}  // namespace pcg
