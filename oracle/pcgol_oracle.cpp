// pcgol_oracle.cpp — CPU restatement of the seqsense/pcgol hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (pcgol_b200/, include/) may
// call, link or import this file.  It is used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// legs, and only as the checker or as the timed CPU baseline.
//
// The reference is Go; there is no Go toolchain in this image, so the reference
// itself cannot be executed here.  This file restates its algorithms in C++
// with float32 arithmetic in the reference's operation order.  It must be
// compiled with  -O2 -ffp-contract=off -fno-fast-math  (see oracle/Makefile):
// Go on amd64 (GOAMD64=v1) rounds after every float32 operation and never
// fuses multiply-add.
//
// Parity pinning: every golden vector the reference's own tests hold for this
// path is replayed against this file by tests/test_oracle_golden.py
// (kdtree_test.go, voxelgrid_test.go, correspondence_test.go,
// evaluator_test.go, icp_test.go, rodrigues_test.go, transform_test.go).
// Behaviour at scales beyond those vectors is pinned only by the reference's
// own property test (KD-tree == brute force, kdtree_test.go:794-834,887-924).
//
// Each function cites the reference file:line it follows (paths relative to
// the reference root).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------
// mat.Vec3  (mat/vec3.go)
// ----------------------------------------------------------------------------
struct Vec3 {
  float v[3];
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};

// mat/vec3.go:38-40
inline Vec3 vsub(const Vec3& a, const Vec3& b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
// mat/vec3.go:42-44
inline Vec3 vadd(const Vec3& a, const Vec3& b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
// mat/vec3.go:30-32
inline Vec3 vmul(const Vec3& a, float s) { return {{a[0] * s, a[1] * s, a[2] * s}}; }
// mat/vec3.go:34-36
inline Vec3 velmul(const Vec3& a, const Vec3& b) { return {{a[0] * b[0], a[1] * b[1], a[2] * b[2]}}; }
// mat/vec3.go:18-20   v0*v0 + v1*v1 + v2*v2, left-associated, each op rounded
inline float vnormsq(const Vec3& a) {
  float t0 = a[0] * a[0];
  float t1 = a[1] * a[1];
  float t2 = a[2] * a[2];
  float s = t0 + t1;
  return s + t2;
}
// mat/vec3.go:22-24
inline float vnorm(const Vec3& a) { return (float)std::sqrt((double)vnormsq(a)); }

// ----------------------------------------------------------------------------
// mat.Mat4  (mat/mat4.go) — column-major, index = col*4+row
// ----------------------------------------------------------------------------
struct Mat4 {
  float m[16];
};

// mat/mat4.go:16-28
Mat4 m4mul(const Mat4& m, const Mat4& a) {
  Mat4 out;
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) {
      float sum = 0;
      for (int k = 0; k < 4; k++) {
        float prod = m.m[4 * k + i] * a.m[4 * j + k];
        sum = sum + prod;
      }
      out.m[4 * j + i] = sum;
    }
  }
  return out;
}
// mat/mat4.go:30-36
Mat4 m4factor(const Mat4& m, float f) {
  Mat4 out;
  for (int i = 0; i < 16; i++) out.m[i] = m.m[i] * f;
  return out;
}
// mat/mat4.go:38-44
Mat4 m4add(const Mat4& m, const Mat4& a) {
  Mat4 out;
  for (int i = 0; i < 16; i++) out.m[i] = m.m[i] + a.m[i];
  return out;
}
// mat/mat4.go:130-137
inline Vec3 m4transform(const Mat4& M, const Vec3& a) {
  const float* m = M.m;
  float den = m[4 * 0 + 3] * a[0];
  den = den + m[4 * 1 + 3] * a[1];
  den = den + m[4 * 2 + 3] * a[2];
  den = den + m[4 * 3 + 3];
  float w = 1.0f / den;
  Vec3 out;
  for (int r = 0; r < 3; r++) {
    float s = m[4 * 0 + r] * a[0];
    s = s + m[4 * 1 + r] * a[1];
    s = s + m[4 * 2 + r] * a[2];
    s = s + m[4 * 3 + r];
    out[r] = s * w;
  }
  return out;
}
// mat/transform.go:7-14
Mat4 m4translate(float x, float y, float z) {
  return Mat4{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, x, y, z, 1}};
}
// mat/transform.go:25-35
Mat4 m4rotate(float x, float y, float z, float ang) {
  float s = (float)std::sin((double)ang);
  float c = (float)std::cos((double)ang);
  float omc = 1 - c;
  return Mat4{{c + x * x * omc, x * y * omc + z * s, x * z * omc - y * s, 0,
               y * x * omc - z * s, c + y * y * omc, y * z * omc + x * s, 0,
               z * x * omc + y * s, z * y * omc - x * s, c + z * z * omc, 0,
               0, 0, 0, 1}};
}

// pc/registration/icp/rodrigues.go:11-33
Mat4 rodrigues_to_rotation(const Vec3& v) {
  float ang = vnorm(v);
  Mat4 r{{0, v[2], -v[1], 0, -v[2], 0, v[0], 0, v[1], -v[0], 0, 0, 0, 0, 0, 0}};
  Mat4 i{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
  float f0, f1;
  if (ang < 0.1f) {
    f0 = 1;
    f1 = 0.5f;
  } else {
    f0 = (float)std::sin((double)ang) / ang;
    f1 = (float)(1 - std::cos((double)ang)) / (ang * ang);
  }
  return m4add(m4add(i, m4factor(r, f0)), m4factor(m4mul(r, r), f1));
}

// ----------------------------------------------------------------------------
// storage.Search  (pc/storage/search.go:8-17)
// ----------------------------------------------------------------------------
struct Neighbor {
  int64_t id;
  float dist_sq;
};

struct Search {
  std::vector<float> xyz;  // Vec3RandomAccessor flattened (n*3)
  int64_t n = 0;
  virtual ~Search() {}
  inline Vec3 at(int64_t i) const { return {{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}}; }
  // KDTree.DeletePoint (kdtree.go:322-332) / naiveSearch.deletePoint (kdtree_test.go:1003-1005).
  // Returns false for an id outside [0, Len()-1] (the reference returns an error and changes nothing).
  virtual bool delete_point(int64_t id) = 0;
  virtual Neighbor nearest(const Vec3& p, float max_range) const = 0;
  virtual void range(const Vec3& p, float max_range, std::vector<Neighbor>& out) const = 0;
};

// canonical order used by the reference's own test (kdtree_test.go:926-941)
inline bool neighbor_id_less(const Neighbor& a, const Neighbor& b) {
  if (a.dist_sq == b.dist_sq) return a.id < b.id;
  return a.dist_sq < b.dist_sq;
}

// Brute force: pc/storage/kdtree/kdtree_test.go:943-985 (naiveSearch)
struct Naive : Search {
  std::vector<uint8_t> deleted;  // kdtree_test.go:945-948 (deletedPoints)
  bool delete_point(int64_t id) override {
    if (id < 0 || id > n - 1) return false;
    if (deleted.empty()) deleted.assign((size_t)n, 0);
    deleted[(size_t)id] = 1;
    return true;
  }
  inline bool gone(int64_t i) const { return !deleted.empty() && deleted[(size_t)i]; }
  // kdtree_test.go:987-1001 (findMinimum): first live point with the smallest coordinate
  int64_t find_minimum(int dim) const {
    int64_t id = -1;
    float mn = 0.f;
    for (int64_t i = 0; i < n; i++) {
      if (gone(i)) continue;
      float v = xyz[3 * i + dim];
      if (id == -1 || v < mn) {
        mn = v;
        id = i;
      }
    }
    return id;
  }
  Neighbor nearest(const Vec3& p, float max_range) const override {
    float dsq = max_range * max_range;
    int64_t id = -1;
    for (int64_t i = 0; i < n; i++) {
      if (gone(i)) continue;
      float d1 = vnormsq(vsub(at(i), p));
      if (d1 < dsq) {
        id = i;
        dsq = d1;
      }
    }
    return {id, dsq};
  }
  void range(const Vec3& p, float max_range, std::vector<Neighbor>& out) const override {
    float th = max_range * max_range;
    out.clear();
    for (int64_t i = 0; i < n; i++) {
      if (gone(i)) continue;
      float d = vnormsq(vsub(at(i), p));
      if (d < th) out.push_back({i, d});
    }
    // reference: sort.Sort by DistSq only (unstable); canonicalised as (DistSq, ID)
    std::sort(out.begin(), out.end(), neighbor_id_less);
  }
};

// KD-tree: pc/storage/kdtree/kdtree.go
struct KDNode {
  int32_t child[2];  // -1 == nil
  int64_t id;
  int32_t dim;
};

struct KDTree : Search {
  std::vector<KDNode> nodes;
  int32_t root = -1;
  float min_dist_sq = 0;  // kdtree.go:20-22
  int max_depth = 0;

  // kdtree.go:348-370 (newNode): sort.Sort(indice by coord[depth%3]); median = indice[len/2]
  int32_t new_node(int64_t* ind, int64_t len, int depth) {
    int dim = depth % 3;
    const float* x = xyz.data();
    std::sort(ind, ind + len, [x, dim](int64_t a, int64_t b) { return x[3 * a + dim] < x[3 * b + dim]; });
    int64_t mid = len / 2;
    int64_t med = ind[mid];
    int32_t left = -1, right = -1;
    if (mid > 0) left = new_node(ind, mid, depth + 1);
    if (mid + 1 < len) right = new_node(ind + mid + 1, len - mid - 1, depth + 1);
    nodes.push_back({{left, right}, med, dim});
    return (int32_t)nodes.size() - 1;
  }
  // kdtree.go:385-395
  int depth_of(int32_t nd, int depth) const {
    if (nd < 0) return depth;
    return std::max(depth_of(nodes[nd].child[0], depth + 1), depth_of(nodes[nd].child[1], depth + 1));
  }
  // kdtree.go:33-56
  void build() {
    std::vector<int64_t> ids(n);
    for (int64_t i = 0; i < n; i++) ids[i] = i;
    nodes.reserve(n);
    if (n > 0) root = new_node(ids.data(), n, 0);
    max_depth = depth_of(root, 0);
  }

  // kdtree.go:224-264 (findMinimumImpl): id of the point with the smallest coordinate `dim` below nd
  // (-1 for nil).  dim > 2 is an error in the reference (-2 here).
  int64_t find_minimum(int32_t nd, int dim) const {
    if (dim > 2) return -2;
    if (nd < 0) return -1;
    const KDNode& nn = nodes[nd];
    auto min_node = [&](int d, int64_t a, int64_t b, int64_t c) {
      int64_t mn = a;
      if (b != -1 && xyz[3 * b + d] < xyz[3 * mn + d]) mn = b;
      if (c != -1 && xyz[3 * c + d] < xyz[3 * mn + d]) mn = c;
      return mn;
    };
    if (nn.dim == dim) {
      if (nn.child[0] < 0) return nn.id;
      return find_minimum(nn.child[0], dim);
    }
    const int64_t m0 = find_minimum(nn.child[0], dim);
    const int64_t m1 = find_minimum(nn.child[1], dim);
    return min_node(dim, nn.id, m0, m1);
  }

  // kdtree.go:266-320 (deleteNodeImpl): returns the node that replaces nd (-1 == nil)
  int32_t delete_node(int32_t nd, int64_t pid) {
    if (nd < 0) return -1;
    if (pid == nodes[nd].id) {
      if (nodes[nd].child[1] >= 0) {
        const int64_t mn = find_minimum(nodes[nd].child[1], nodes[nd].dim);
        const int32_t child = delete_node(nodes[nd].child[1], mn);
        nodes[nd].id = mn;
        nodes[nd].child[1] = child;
      } else if (nodes[nd].child[0] >= 0) {
        const int64_t mn = find_minimum(nodes[nd].child[0], nodes[nd].dim);
        const int32_t child = delete_node(nodes[nd].child[0], mn);
        nodes[nd].id = mn;
        nodes[nd].child[0] = -1;
        nodes[nd].child[1] = child;
      } else {
        return -1;
      }
      return nd;
    }
    const float pivot = xyz[3 * nodes[nd].id + nodes[nd].dim];
    const float v = xyz[3 * pid + nodes[nd].dim];
    if (v <= pivot) nodes[nd].child[0] = delete_node(nodes[nd].child[0], pid);
    if (v >= pivot) nodes[nd].child[1] = delete_node(nodes[nd].child[1], pid);
    return nd;
  }

  // kdtree.go:322-332 (DeletePoint)
  bool delete_point(int64_t pid) override {
    if (pid < 0 || pid > n - 1) return false;
    root = delete_node(root, pid);
    return true;
  }

  // kdtree.go:199-222 (searchLeafNode); the node stack is the tail of `st`
  void search_leaf(std::vector<int32_t>& st, const Vec3& p) const {
    for (;;) {
      const KDNode& parent = nodes[st.back()];
      int32_t c0 = parent.child[0], c1 = parent.child[1];
      if (c0 < 0 && c1 < 0) return;
      if (c0 < 0) {
        st.push_back(c1);
        continue;
      }
      if (c1 < 0) {
        st.push_back(c0);
        continue;
      }
      float pivot_val = xyz[3 * parent.id + parent.dim], val = p[parent.dim];
      if (pivot_val > val)
        st.push_back(c0);
      else
        st.push_back(c1);
    }
  }

  // kdtree.go:94-146 (nearestImpl); this stack occupies st[b..]
  Neighbor nearest_impl(std::vector<int32_t>& st, size_t b, const Vec3& p, float max_range_sq) const {
    size_t i = st.size() - 1;
    Neighbor n1{nodes[st[i]].id, vnormsq(vsub(at(nodes[st[i]].id), p))};
    if (n1.dist_sq > max_range_sq) {
      n1.id = -1;
      n1.dist_sq = max_range_sq;
    }
    if (n1.dist_sq < min_dist_sq) return n1;
    for (size_t jj = i; jj-- > b;) {
      const KDNode& nj = nodes[st[jj]];
      Vec3 pivot = at(nj.id);
      float from_pivot = p[nj.dim] - pivot[nj.dim];
      float from_pivot_sq = from_pivot * from_pivot;
      if (from_pivot_sq > n1.dist_sq) continue;
      float dsq_pivot = vnormsq(vsub(pivot, p));
      if (dsq_pivot < n1.dist_sq) {
        n1.id = nj.id;
        n1.dist_sq = dsq_pivot;
        if (n1.dist_sq < min_dist_sq) break;
      }
      int32_t next = (nj.child[0] == st[jj + 1]) ? nj.child[1] : nj.child[0];
      if (next < 0) continue;
      size_t nb = st.size();
      st.push_back(next);
      search_leaf(st, p);
      Neighbor n2 = nearest_impl(st, nb, p, n1.dist_sq);
      st.resize(nb);
      if (n2.id >= 0) {
        n1 = n2;
        if (n1.dist_sq < min_dist_sq) break;
      }
    }
    return n1;
  }

  // kdtree.go:83-92
  Neighbor nearest(const Vec3& p, float max_range) const override {
    if (root < 0) return {-1, max_range * max_range};
    static thread_local std::vector<int32_t> st;  // stands in for the sync.Pool of node stacks
    st.clear();
    st.push_back(root);
    search_leaf(st, p);
    return nearest_impl(st, 0, p, max_range * max_range);
  }

  // kdtree.go:163-197 (rangeImpl)
  void range_impl(std::vector<int32_t>& st, size_t b, const Vec3& p, float max_range_sq,
                  std::vector<Neighbor>& out) const {
    size_t i = st.size() - 1;
    int64_t id = nodes[st[i]].id;
    float dsq = vnormsq(vsub(at(id), p));
    if (dsq < max_range_sq) out.push_back({id, dsq});
    for (size_t jj = i; jj-- > b;) {
      const KDNode& nj = nodes[st[jj]];
      Vec3 pivot = at(nj.id);
      float from_pivot = p[nj.dim] - pivot[nj.dim];
      float from_pivot_sq = from_pivot * from_pivot;
      if (from_pivot_sq > max_range_sq) continue;
      float dsq_pivot = vnormsq(vsub(pivot, p));
      if (dsq_pivot < max_range_sq) out.push_back({nj.id, dsq_pivot});
      int32_t next = (nj.child[0] == st[jj + 1]) ? nj.child[1] : nj.child[0];
      if (next < 0) continue;
      size_t nb = st.size();
      st.push_back(next);
      search_leaf(st, p);
      range_impl(st, nb, p, max_range_sq, out);
      st.resize(nb);
    }
  }

  // kdtree.go:148-161
  void range(const Vec3& p, float max_range, std::vector<Neighbor>& out) const override {
    out.clear();
    if (root < 0) return;
    static thread_local std::vector<int32_t> st;
    st.clear();
    st.push_back(root);
    search_leaf(st, p);
    range_impl(st, 0, p, max_range * max_range, out);
    // reference sorts by DistSq only (kdtree.go:159,425-427; unstable, ties unordered);
    // canonical order per the reference's test is (DistSq, ID)  kdtree_test.go:926-941
    std::sort(out.begin(), out.end(), neighbor_id_less);
  }
};

// Go's int(float32) is a truncating conversion to int64 (CVTTSS2SQ on amd64).
// Out-of-range / NaN is implementation-specific in Go; flagged instead of emulated.
inline int64_t go_int(float f, bool* undefined) {
  if (!(std::fabs(f) < 9.0e18f)) {
    *undefined = true;
    return 0;
  }
  return (int64_t)f;
}

inline float load_f32(const uint8_t* p) {
  float f;
  std::memcpy(&f, p, 4);
  return f;
}
inline void store_f32(uint8_t* p, float f) { std::memcpy(p, &f, 4); }

// Interleaved-record accessor: pc/pointcloud.go:130-163 (Vec3Iterator) +
// pc/iterator.go:132-137 (float32Iterator.Vec3At) / :171-173 (naiveVec3Iterator.Vec3At)
struct Cloud {
  const uint8_t* data;
  int64_t n;
  int64_t stride;
  int64_t off[3];
  inline Vec3 at(int64_t i) const {
    const uint8_t* r = data + i * stride;
    return {{load_f32(r + off[0]), load_f32(r + off[1]), load_f32(r + off[2])}};
  }
};

// pc/filter/voxelgrid/voxelgrid.go:17-21
struct Voxel {
  Vec3 sum;
  int64_t num;
  int64_t index;
};
static_assert(sizeof(Voxel) == 32, "voxel is 32 bytes like the Go struct");

enum {
  ORC_OK = 0,
  ORC_E_NO_POINT = 1,         // pc/minmax.go:10-12  errors.New("no point")
  ORC_E_REF_WOULD_PANIC = 2,  // index out of range / negative makeslice in the reference
  ORC_E_REF_UNDEFINED = 3,    // float->int conversion out of range (implementation-specific in Go)
  ORC_E_TOO_LARGE = 4,        // dense voxel array beyond the cap given to the oracle
  ORC_E_NOT_ENOUGH_PAIRS = 5  // pc/registration/icp/evaluator.go:15-17
};

// pc/minmax.go:9-26
void minmax_vec3(const Cloud& c, Vec3* mn, Vec3* mx) {
  *mn = c.at(0);
  *mx = c.at(0);
  for (int64_t i = 1; i < c.n; i++) {
    Vec3 v = c.at(i);
    for (int k = 0; k < 3; k++) {
      if (v[k] < (*mn)[k]) (*mn)[k] = v[k];
      if (v[k] > (*mx)[k]) (*mx)[k] = v[k];
    }
  }
}

struct VoxelFilter {
  Vec3 leaf;
  std::vector<Voxel> voxels;  // f.voxels, reused across chunks (voxelgrid.go:14,139-145)
  int64_t cap_voxels;

  // pc/filter/voxelgrid/voxelgrid.go:136-187 (filterChunk).
  // `ids` == nullptr: iterate the whole cloud in order (RawIndex = i);
  // else iterate the index list (RawIndex = ids[j], pc/indice.go:17-23).
  int filter_chunk(const Vec3& v_min, const Vec3& size, const Cloud& c, const int64_t* ids, int64_t n_ids,
                   std::vector<uint8_t>& out, int64_t* n_out) {
    bool undef = false;
    int64_t xs = go_int(size[0] / leaf[0], &undef), ys = go_int(size[1] / leaf[1], &undef),
            zs = go_int(size[2] / leaf[2], &undef);
    if (undef) return ORC_E_REF_UNDEFINED;
    int64_t n_voxels = (xs + 1) * (ys + 1) * (zs + 1);
    if ((int64_t)voxels.size() < n_voxels) {
      if (n_voxels > cap_voxels) return ORC_E_TOO_LARGE;
      voxels.assign((size_t)n_voxels, Voxel{{{0, 0, 0}}, 0, 0});
    } else {
      // (a negative n_voxels never reaches make(): len(f.voxels) < n is false, the slice stays as is)
      std::fill(voxels.begin(), voxels.end(), Voxel{{{0, 0, 0}}, 0, 0});
    }
    int64_t n = 0;
    int64_t count = ids ? n_ids : c.n;
    for (int64_t j = 0; j < count; j++) {
      int64_t raw = ids ? ids[j] : j;
      Vec3 p = vsub(c.at(raw), v_min);
      int64_t x = go_int(p[0] / leaf[0], &undef), y = go_int(p[1] / leaf[1], &undef),
              z = go_int(p[2] / leaf[2], &undef);
      if (undef) return ORC_E_REF_UNDEFINED;
      int64_t key = x + xs * (y + ys * z);
      if (key < 0 || key >= (int64_t)voxels.size()) return ORC_E_REF_WOULD_PANIC;
      Voxel& v = voxels[(size_t)key];
      if (v.num == 0) {
        v.index = raw;
        n++;
      }
      v.num++;
      v.sum = vadd(v.sum, p);
    }
    size_t base = out.size();
    out.resize(base + (size_t)(c.stride * n));
    uint8_t* dst = out.data() + base;
    for (size_t i = 0; i < voxels.size(); i++) {
      const Voxel& v = voxels[i];
      if (v.num > 0) {
        std::memcpy(dst, c.data + v.index * c.stride, (size_t)c.stride);
        if (v.num > 1) {
          Vec3 cen = vadd(vmul(v.sum, 1.0f / (float)v.num), v_min);
          store_f32(dst + c.off[0], cen[0]);
          store_f32(dst + c.off[1], cen[1]);
          store_f32(dst + c.off[2], cen[2]);
        }
        dst += c.stride;
      }
    }
    *n_out += n;
    return ORC_OK;
  }

  // pc/filter/voxelgrid/voxelgrid.go:35-134 (Filter)
  int filter(const Cloud& c, const int64_t chunk[3], std::vector<uint8_t>& out, int64_t* n_out) {
    *n_out = 0;
    out.clear();
    if (c.n == 0) return ORC_E_NO_POINT;
    Vec3 v_min, v_max;
    minmax_vec3(c, &v_min, &v_max);
    if (chunk[0] * chunk[1] * chunk[2] == 0) {
      return filter_chunk(v_min, v_max /* sic: voxelgrid.go:46 */, c, nullptr, 0, out, n_out);
    }
    Vec3 size = vsub(v_max, v_min);
    Vec3 chunk_size{{leaf[0] * (float)chunk[0], leaf[1] * (float)chunk[1], leaf[2] * (float)chunk[2]}};
    for (int i = 0; i < 3; i++) {
      if (chunk_size[i] > size[i] + leaf[i]) chunk_size[i] = size[i] + leaf[i];
    }
    bool undef = false;
    int64_t nx = go_int(size[0] / chunk_size[0], &undef) + 1, ny = go_int(size[1] / chunk_size[1], &undef) + 1,
            nz = go_int(size[2] / chunk_size[2], &undef) + 1;
    if (undef) return ORC_E_REF_UNDEFINED;
    int64_t n_chunks = nx * ny * nz;
    if (n_chunks < 0 || n_chunks > (int64_t)1 << 32) return ORC_E_TOO_LARGE;
    std::vector<std::vector<int64_t>> indices((size_t)n_chunks);
    // voxelgrid.go:76-79,87-99  (count pass + fill pass collapsed: same order, same membership)
    for (int64_t i = 0; i < c.n; i++) {
      Vec3 p = vsub(c.at(i), v_min);
      int64_t x = go_int(p[0] / chunk_size[0], &undef), y = go_int(p[1] / chunk_size[1], &undef),
              z = go_int(p[2] / chunk_size[2], &undef);
      if (undef) return ORC_E_REF_UNDEFINED;
      int64_t cid = ((z * ny) + y) * nx + x;
      if (cid < 0 || cid >= n_chunks) return ORC_E_REF_WOULD_PANIC;
      indices[(size_t)cid].push_back(i);
    }
    // voxelgrid.go:102-116
    for (int64_t cid = 0; cid < n_chunks; cid++) {
      const auto& ind = indices[(size_t)cid];
      if (ind.empty()) continue;
      int64_t t = cid;
      int64_t x = t % nx;
      t = t / nx;
      int64_t y = t % ny;
      int64_t z = t / ny;
      Vec3 cp{{(float)x, (float)y, (float)z}};
      Vec3 vc_min = vadd(v_min, velmul(cp, chunk_size));
      int rc = filter_chunk(vc_min, chunk_size, c, ind.data(), (int64_t)ind.size(), out, n_out);
      if (rc != ORC_OK) return rc;
    }
    return ORC_OK;
  }
};

// Same result as VoxelFilter::filter without the dense voxel array (memory ∝ points), so that
// the large configurations can be checked on a small host.  Validated against the literal
// version above by tests/test_oracle_golden.py; the literal version is the specification.
int voxel_filter_sparse(const Cloud& c, const Vec3& leaf, const int64_t chunk[3], std::vector<uint8_t>& out,
                        int64_t* n_out) {
  *n_out = 0;
  out.clear();
  if (c.n == 0) return ORC_E_NO_POINT;
  Vec3 v_min, v_max;
  minmax_vec3(c, &v_min, &v_max);
  bool chunked = chunk[0] * chunk[1] * chunk[2] != 0;
  bool undef = false;
  Vec3 chunk_size{{0, 0, 0}};
  int64_t nx = 1, ny = 1, nz = 1;
  Vec3 size_for_xs = v_max;
  if (chunked) {
    Vec3 size = vsub(v_max, v_min);
    chunk_size = {{leaf[0] * (float)chunk[0], leaf[1] * (float)chunk[1], leaf[2] * (float)chunk[2]}};
    for (int i = 0; i < 3; i++)
      if (chunk_size[i] > size[i] + leaf[i]) chunk_size[i] = size[i] + leaf[i];
    nx = go_int(size[0] / chunk_size[0], &undef) + 1;
    ny = go_int(size[1] / chunk_size[1], &undef) + 1;
    nz = go_int(size[2] / chunk_size[2], &undef) + 1;
    size_for_xs = chunk_size;
  }
  int64_t xs = go_int(size_for_xs[0] / leaf[0], &undef), ys = go_int(size_for_xs[1] / leaf[1], &undef),
          zs = go_int(size_for_xs[2] / leaf[2], &undef);
  if (undef) return ORC_E_REF_UNDEFINED;
  int64_t n_voxels = (xs + 1) * (ys + 1) * (zs + 1);
  if (n_voxels < 0) n_voxels = 0;  // make() is skipped when len(f.voxels)=0 >= n; every key is then out of range
  int64_t n_chunks = nx * ny * nz;
  struct Item {
    int64_t cid, key, idx;
  };
  std::vector<Item> items((size_t)c.n);
  for (int64_t i = 0; i < c.n; i++) {
    Vec3 pt = c.at(i);
    int64_t cid = 0;
    Vec3 vc_min = v_min;
    if (chunked) {
      Vec3 p = vsub(pt, v_min);
      int64_t x = go_int(p[0] / chunk_size[0], &undef), y = go_int(p[1] / chunk_size[1], &undef),
              z = go_int(p[2] / chunk_size[2], &undef);
      if (undef) return ORC_E_REF_UNDEFINED;
      cid = ((z * ny) + y) * nx + x;
      if (cid < 0 || cid >= n_chunks) return ORC_E_REF_WOULD_PANIC;
      int64_t t = cid;
      int64_t cx = t % nx;
      t /= nx;
      int64_t cy = t % ny;
      int64_t cz = t / ny;
      Vec3 cp{{(float)cx, (float)cy, (float)cz}};
      vc_min = vadd(v_min, velmul(cp, chunk_size));
    }
    Vec3 p = vsub(pt, vc_min);
    int64_t x = go_int(p[0] / leaf[0], &undef), y = go_int(p[1] / leaf[1], &undef), z = go_int(p[2] / leaf[2], &undef);
    if (undef) return ORC_E_REF_UNDEFINED;
    int64_t key = x + xs * (y + ys * z);
    if (key < 0 || key >= n_voxels) return ORC_E_REF_WOULD_PANIC;
    items[(size_t)i] = {cid, key, i};
  }
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
    if (a.cid != b.cid) return a.cid < b.cid;
    return a.key < b.key;
  });
  size_t i = 0;
  while (i < items.size()) {
    size_t j = i;
    int64_t cid = items[i].cid;
    Vec3 vc_min = v_min;
    if (chunked) {
      int64_t t = cid;
      int64_t cx = t % nx;
      t /= nx;
      int64_t cy = t % ny;
      int64_t cz = t / ny;
      Vec3 cp{{(float)cx, (float)cy, (float)cz}};
      vc_min = vadd(v_min, velmul(cp, chunk_size));
    }
    Vec3 sum{{0, 0, 0}};
    while (j < items.size() && items[j].cid == cid && items[j].key == items[i].key) {
      sum = vadd(sum, vsub(c.at(items[j].idx), vc_min));
      j++;
    }
    int64_t num = (int64_t)(j - i);
    size_t base = out.size();
    out.resize(base + (size_t)c.stride);
    uint8_t* dst = out.data() + base;
    std::memcpy(dst, c.data + items[i].idx * c.stride, (size_t)c.stride);
    if (num > 1) {
      Vec3 cen = vadd(vmul(sum, 1.0f / (float)num), vc_min);
      store_f32(dst + c.off[0], cen[0]);
      store_f32(dst + c.off[1], cen[1]);
      store_f32(dst + c.off[2], cen[2]);
    }
    (*n_out)++;
    i = j;
  }
  return ORC_OK;
}

// ----------------------------------------------------------------------------
// pc/registration/icp
// ----------------------------------------------------------------------------
struct Pair {
  int64_t base_id, target_id;
  float dsq;
};

// correspondence.go:22-37
void icp_pairs(const Search& base, const float* target, int64_t n, float max_dist, std::vector<Pair>& out) {
  out.clear();
  out.reserve((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    Vec3 t{{target[3 * i], target[3 * i + 1], target[3 * i + 2]}};
    Neighbor nn = base.nearest(t, max_dist);
    if (nn.id < 0) continue;
    out.push_back({nn.id, i, nn.dist_sq});
  }
}

struct Evaluated {  // evaluator.go:25-30 (Hessian is never written by the reference)
  float value;
  float gradient[6];
  float dist_rms;
};

// EvaluateWeightFn (evaluator.go:19-23,72): a Go closure in the reference; here the parametric family the C ABI
// offers (pcg_weight_fn), every operation a rounded float32 one.  0 = DefaultEvaluateWeightFn (w = 1).
enum { ORC_WEIGHT_CONSTANT = 0, ORC_WEIGHT_TRUNCATED = 1, ORC_WEIGHT_HUBER = 2 };
static float eval_weight(int fn, float param, float dsq) {
  switch (fn) {
    case ORC_WEIGHT_TRUNCATED: return dsq < param ? 1.0f : 0.0f;
    case ORC_WEIGHT_HUBER: {
      if (dsq <= param) return 1.0f;
      float q = param / dsq;  // float32 division, then the correctly rounded float32 square root
      return std::sqrt(q);
    }
    default: return 1.0f;
  }
}

// evaluator.go:91-189; weight_fn = 0 is the default weight function (w = 1, evaluator.go:21-23)
template <typename Acc>
int icp_evaluate_t(const Search& base, const float* target, int64_t n, float max_dist, int min_pairs,
                   Evaluated* out, int64_t* n_pairs, int weight_fn = 0, float weight_param = 0.f) {
  if (min_pairs == 0) min_pairs = 6;
  static thread_local std::vector<Pair> pairs;
  icp_pairs(base, target, n, max_dist, pairs);
  if (n_pairs) *n_pairs = (int64_t)pairs.size();
  if ((int64_t)pairs.size() < (int64_t)min_pairs) return ORC_E_NOT_ENOUGH_PAIRS;
  Acc value = 0, sum_weight = 0, rms = 0;
  Acc g[6] = {0, 0, 0, 0, 0, 0};
  for (const Pair& pr : pairs) {
    Vec3 pb = base.at(pr.base_id);
    Vec3 pt{{target[3 * pr.target_id], target[3 * pr.target_id + 1], target[3 * pr.target_id + 2]}};
    float w = eval_weight(weight_fn, weight_param, pr.dsq);  // evaluator.go:130
    value += w * pr.dsq;
    sum_weight += w;
    float x0 = pt[0], y0 = pt[1], z0 = pt[2];
    float x1 = pb[0], y1 = pb[1], z1 = pb[2];
    g[0] += w * (x0 - x1);
    g[1] += w * (y0 - y1);
    g[2] += w * (z0 - z1);
    {
      float a = z0 * y1, b = y0 * z1;
      g[3] += w * (a - b);
    }
    {
      float a = x0 * z1, b = z0 * x1;
      g[4] += w * (a - b);
    }
    {
      float a = y0 * x1, b = x0 * y1;
      g[5] += w * (a - b);
    }
    rms += w * vnormsq(pt);
  }
  // From here on the reference works on float32 fields (evaluator.go:156-186).
  float valuef = (float)value, sum_weightf = (float)sum_weight, rmsf = (float)rms;
  float gf[6];
  for (int i = 0; i < 6; i++) gf[i] = (float)g[i];
  float f = 1;
  if (sum_weightf > 1) f = 1 / sum_weightf;
  valuef *= f;
  float two_f = 2 * f;
  for (int i = 0; i < 6; i++) gf[i] *= two_f;
  float dist_rms = (float)std::sqrt((double)(rmsf * f));
  float rot_limit = 1;
  float dist = (float)std::sqrt((double)valuef);
  for (int i = 3; i < 6; i++) {
    float d = gf[i] * dist_rms;
    if (d < 0) d = -d;
    if (dist < d) {
      float l = dist / d;
      if (rot_limit > l) rot_limit = l;
    }
  }
  for (int i = 3; i < 6; i++) gf[i] *= rot_limit;
  out->value = valuef;
  for (int i = 0; i < 6; i++) out->gradient[i] = gf[i];
  out->dist_rms = dist_rms;
  return ORC_OK;
}

struct IcpParams {
  float max_dist;      // NearestPointCorresponder.MaxDist  correspondence.go:18-20
  int32_t min_pairs;   // PointToPointEvaluator.MinPairs, 0 -> 6   evaluator.go:92-95
  float weight[6];     // GradientDescentUpdaterFactory.Weight, all-zero -> 0.3   updater.go:15,25-27
  float threshold[6];  // .Threshold, all-zero -> 0.01   updater.go:16,28-30
  int32_t max_iteration;  // .MaxIteration, 0 -> 20   updater.go:31-33
  int32_t f64_accumulate;  // oracle-only: accumulate the 9 sums in float64 (error budgeting)
  int32_t weight_fn;       // EvaluateWeightFn family (evaluator.go:19-23): 0 constant, 1 truncated, 2 Huber
  float weight_param;
};

struct IcpStat {  // stat.go:3-6
  Evaluated ev;
  int32_t num_iteration;
};

struct Updater {  // updater.go:39-42
  float weight[6], threshold[6];
  int max_iteration;
  int i = 0;
  // updater.go:24-37
  explicit Updater(const IcpParams& p) {
    bool wz = true, tz = true;
    for (int k = 0; k < 6; k++) {
      wz = wz && p.weight[k] == 0;
      tz = tz && p.threshold[k] == 0;
    }
    for (int k = 0; k < 6; k++) {
      weight[k] = wz ? 0.3f : p.weight[k];
      threshold[k] = tz ? 0.01f : p.threshold[k];
    }
    max_iteration = p.max_iteration == 0 ? 20 : p.max_iteration;
  }
  // updater.go:44-71
  bool update(Mat4* trans, const Evaluated& ev) {
    bool flat = true;
    for (int j = 0; j < 6; j++) {
      float g = ev.gradient[j];
      if (g < -threshold[j] || threshold[j] < g) {
        flat = false;
        break;
      }
    }
    if (flat) return true;
    float factor_iter = -(1 - ((float)i / (float)max_iteration));
    float delta[6];
    for (int k = 0; k < 6; k++) {
      float fw = factor_iter * weight[k];
      delta[k] = fw * ev.gradient[k];
    }
    Mat4 delta_trans = m4translate(delta[0], delta[1], delta[2]);
    Mat4 delta_rot = rodrigues_to_rotation(Vec3{{delta[3], delta[4], delta[5]}});
    *trans = m4mul(delta_trans, m4mul(delta_rot, *trans));
    i++;
    return i >= max_iteration;
  }
};

// icp.go:23-67
int icp_fit(const Search& base, const float* target, int64_t n, const IcpParams& prm, Mat4* trans_out,
            IcpStat* stat) {
  std::vector<float> tt(target, target + 3 * n);
  Updater up(prm);
  std::memset(stat, 0, sizeof(*stat));
  Mat4 trans = m4translate(0, 0, 0);
  for (;;) {
    Evaluated ev;
    int rc = prm.f64_accumulate
                 ? icp_evaluate_t<double>(base, tt.data(), n, prm.max_dist, prm.min_pairs, &ev, nullptr, prm.weight_fn,
                                          prm.weight_param)
                 : icp_evaluate_t<float>(base, tt.data(), n, prm.max_dist, prm.min_pairs, &ev, nullptr, prm.weight_fn,
                                         prm.weight_param);
    stat->num_iteration++;
    if (rc != ORC_OK) {
      *trans_out = trans;
      return rc;
    }
    stat->ev = ev;
    bool converged = up.update(&trans, ev);
    if (converged) break;
    for (int64_t i = 0; i < n; i++) {
      Vec3 t = m4transform(trans, Vec3{{target[3 * i], target[3 * i + 1], target[3 * i + 2]}});
      tt[3 * i] = t[0];
      tt[3 * i + 1] = t[1];
      tt[3 * i + 2] = t[2];
    }
  }
  *trans_out = trans;
  return ORC_OK;
}

// ----------------------------------------------------------------------------
// Extension check (NOT a restatement: the reference declares Evaluated.Hessian and
// HasHessian(), evaluator.go:25-36,76, but never fills / consumes them).  Literal
// definition of what the product's PCG_ICP_WITH_HESSIAN / PCG_UPDATER_GAUSS_NEWTON compute:
// per pair the 3x6 Jacobian J = [I | -[pt]x] of the residual r = pt - pb under the increment
// Translate(dt) * Rodrigues(dw) applied on the left (updater.go:65-68); A = sum J^T J,
// b = sum J^T r in float64, pair by pair.  Evaluated.Hessian = 2f * A (f of evaluator.go:156-159).
// ----------------------------------------------------------------------------
int icp_normal_equations(const Search& base, const float* target, int64_t n, float max_dist, int min_pairs,
                         double A[36], double b[6], int64_t* n_pairs) {
  if (min_pairs == 0) min_pairs = 6;
  std::vector<Pair> pairs;
  icp_pairs(base, target, n, max_dist, pairs);
  if (n_pairs) *n_pairs = (int64_t)pairs.size();
  for (int i = 0; i < 36; i++) A[i] = 0;
  for (int i = 0; i < 6; i++) b[i] = 0;
  if ((int64_t)pairs.size() < (int64_t)min_pairs) return ORC_E_NOT_ENOUGH_PAIRS;
  for (const Pair& pr : pairs) {
    Vec3 pb = base.at(pr.base_id);
    const double x = target[3 * pr.target_id], y = target[3 * pr.target_id + 1], z = target[3 * pr.target_id + 2];
    const double r[3] = {x - (double)pb[0], y - (double)pb[1], z - (double)pb[2]};
    // J[row][col]: d r / d (tx,ty,tz,wx,wy,wz);  d(w x p) = -[p]x dw
    const double J[3][6] = {{1, 0, 0, 0, z, -y}, {0, 1, 0, -z, 0, x}, {0, 0, 1, y, -x, 0}};
    for (int c = 0; c < 6; c++) {
      for (int rr = 0; rr < 6; rr++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += J[k][rr] * J[k][c];
        A[c * 6 + rr] += s;  // index = col*6 + row
      }
      double s = 0;
      for (int k = 0; k < 3; k++) s += J[k][c] * r[k];
      b[c] += s;
    }
  }
  return ORC_OK;
}

// Solves A d = -b (6x6) by Gaussian elimination with partial pivoting, float64.
bool solve6(const double A_in[36], const double b[6], double d[6]) {
  double M[6][7];
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) M[r][c] = A_in[c * 6 + r];
    M[r][6] = -b[r];
  }
  for (int c = 0; c < 6; c++) {
    int piv = c;
    for (int r = c + 1; r < 6; r++)
      if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
    if (!(std::fabs(M[piv][c]) > 0)) return false;
    if (piv != c)
      for (int k = 0; k < 7; k++) std::swap(M[piv][k], M[c][k]);
    for (int r = c + 1; r < 6; r++) {
      const double f = M[r][c] / M[c][c];
      for (int k = c; k < 7; k++) M[r][k] -= f * M[c][k];
    }
  }
  for (int r = 5; r >= 0; r--) {
    double v = M[r][6];
    for (int k = r + 1; k < 6; k++) v -= M[r][k] * d[k];
    d[r] = v / M[r][r];
  }
  return true;
}

// Fit with the Gauss-Newton step in place of the damped gradient: same loop (icp.go:23-67),
// same convergence test and increment composition (updater.go:45-54,65-70).
int icp_fit_gn(const Search& base, const float* target, int64_t n, const IcpParams& prm, Mat4* trans_out,
               IcpStat* stat) {
  std::vector<float> tt(target, target + 3 * n);
  Updater up(prm);
  std::memset(stat, 0, sizeof(*stat));
  Mat4 trans = m4translate(0, 0, 0);
  for (;;) {
    Evaluated ev;
    int rc = prm.f64_accumulate
                 ? icp_evaluate_t<double>(base, tt.data(), n, prm.max_dist, prm.min_pairs, &ev, nullptr, prm.weight_fn,
                                          prm.weight_param)
                 : icp_evaluate_t<float>(base, tt.data(), n, prm.max_dist, prm.min_pairs, &ev, nullptr, prm.weight_fn,
                                         prm.weight_param);
    stat->num_iteration++;
    if (rc != ORC_OK) {
      *trans_out = trans;
      return rc;
    }
    stat->ev = ev;
    bool flat = true;
    for (int j = 0; j < 6; j++) {
      float g = ev.gradient[j];
      if (g < -up.threshold[j] || up.threshold[j] < g) {
        flat = false;
        break;
      }
    }
    if (flat) break;
    double A[36], b[6], d[6];
    icp_normal_equations(base, tt.data(), n, prm.max_dist, prm.min_pairs, A, b, nullptr);
    if (!solve6(A, b, d)) break;
    Mat4 delta_trans = m4translate((float)d[0], (float)d[1], (float)d[2]);
    Mat4 delta_rot = rodrigues_to_rotation(Vec3{{(float)d[3], (float)d[4], (float)d[5]}});
    trans = m4mul(delta_trans, m4mul(delta_rot, trans));
    up.i++;
    if (up.i >= up.max_iteration) break;
    for (int64_t i = 0; i < n; i++) {
      Vec3 t = m4transform(trans, Vec3{{target[3 * i], target[3 * i + 1], target[3 * i + 2]}});
      tt[3 * i] = t[0];
      tt[3 * i + 1] = t[1];
      tt[3 * i + 2] = t[2];
    }
  }
  *trans_out = trans;
  return ORC_OK;
}

// ----------------------------------------------------------------------------
// pc/segmentation/regiongrowing/regiongrowing.go:23-56 (RegionGrowing.Segment)
// Range lists arrive in the canonical (DistSq, ID) order (the reference sorts by DistSq only and
// leaves ties unordered, so the reference's own test sorts the result before comparing,
// regiongrowing_test.go:186; with the canonical order the BFS order below is reproducible).
// ----------------------------------------------------------------------------
void region_growing_segment(const Search& search, const uint32_t* label, const Vec3& p, float max_range,
                            std::vector<int64_t>& indice) {
  indice.clear();
  std::vector<Neighbor> nb;
  search.range(p, max_range, nb);
  if (nb.empty()) return;
  const uint32_t target = label[nb[0].id];
  std::vector<int64_t> next;
  std::vector<uint8_t> to_visit((size_t)search.n, 0);
  for (const Neighbor& n : nb) {
    next.push_back(n.id);
    to_visit[(size_t)n.id] = 1;
  }
  for (size_t head = 0; head < next.size(); head++) {
    const int64_t id = next[head];
    if (label[id] != target) continue;
    indice.push_back(id);
    search.range(search.at(id), max_range, nb);
    for (const Neighbor& n : nb) {
      if (!to_visit[(size_t)n.id]) {
        next.push_back(n.id);
        to_visit[(size_t)n.id] = 1;
      }
    }
  }
}

template <typename F>
void parallel_for(int64_t n, int threads, F f) {
  if (threads <= 1 || n < 2) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  int64_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    int64_t lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] { f(lo, hi); });
  }
  for (auto& t : th) t.join();
}

}  // namespace

// ----------------------------------------------------------------------------
// C entry points (ctypes) — see oracle/oracle.py
// ----------------------------------------------------------------------------
extern "C" {

// kind: 0 = KD-tree (kdtree.New), 1 = brute force (naiveSearch)
void* orc_search_new(const float* xyz, int64_t n, int32_t kind) {
  Search* s;
  if (kind == 0)
    s = new KDTree();
  else
    s = new Naive();
  s->xyz.assign(xyz, xyz + 3 * n);
  s->n = n;
  if (kind == 0) static_cast<KDTree*>(s)->build();
  return s;
}
void orc_search_free(void* h) { delete static_cast<Search*>(h); }

// KDTree.MinDistSq (kdtree.go:20-22); ignored for brute force
void orc_search_set_min_dist_sq(void* h, float v) {
  if (auto* k = dynamic_cast<KDTree*>(static_cast<Search*>(h))) k->min_dist_sq = v;
}

// DeletePoint: 0 = ok, 1 = "does not correspond to any point in the tree" (kdtree.go:323-325)
int32_t orc_search_delete_point(void* h, int64_t id) { return static_cast<Search*>(h)->delete_point(id) ? 0 : 1; }
// findMinimumImpl from the root (kdtree_test.go:388-411) / naiveSearch.findMinimum; -2 = error (dim > 2)
int64_t orc_search_find_minimum(void* h, int32_t dim) {
  Search* s = static_cast<Search*>(h);
  if (auto* k = dynamic_cast<KDTree*>(s)) return k->find_minimum(k->root, dim);
  if (auto* nv = dynamic_cast<Naive*>(s)) return dim > 2 ? -2 : nv->find_minimum(dim);
  return -2;
}

int64_t orc_kdtree_num_nodes(void* h) {
  auto* k = dynamic_cast<KDTree*>(static_cast<Search*>(h));
  return k ? (int64_t)k->nodes.size() : -1;
}
int32_t orc_kdtree_max_depth(void* h) {
  auto* k = dynamic_cast<KDTree*>(static_cast<Search*>(h));
  return k ? k->max_depth : -1;
}
// Dump the tree (node arrays indexed by internal node number; returns root)
int32_t orc_kdtree_dump(void* h, int64_t* id, int32_t* dim, int32_t* left, int32_t* right) {
  auto* k = dynamic_cast<KDTree*>(static_cast<Search*>(h));
  if (!k) return -1;
  for (size_t i = 0; i < k->nodes.size(); i++) {
    id[i] = k->nodes[i].id;
    dim[i] = k->nodes[i].dim;
    left[i] = k->nodes[i].child[0];
    right[i] = k->nodes[i].child[1];
  }
  return k->root;
}
// searchLeafNode from the root (kdtree_test.go:250-279): returns the id at the top of the stack
int64_t orc_kdtree_search_leaf(void* h, const float* p) {
  auto* k = dynamic_cast<KDTree*>(static_cast<Search*>(h));
  if (!k || k->root < 0) return -1;
  std::vector<int32_t> st{k->root};
  k->search_leaf(st, Vec3{{p[0], p[1], p[2]}});
  return k->nodes[st.back()].id;
}

void orc_search_nearest(void* h, const float* q, int64_t nq, float max_range, int64_t* ids, float* dist_sq,
                        int32_t threads) {
  const Search* s = static_cast<Search*>(h);
  parallel_for(nq, threads, [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++) {
      Neighbor nb = s->nearest(Vec3{{q[3 * i], q[3 * i + 1], q[3 * i + 2]}}, max_range);
      ids[i] = nb.id;
      dist_sq[i] = nb.dist_sq;
    }
  });
}

// Range, CSR output. Pass ids == nullptr to obtain counts only (offsets[nq] = total).
// Lists are in canonical (DistSq, ID) order.
void orc_search_range(void* h, const float* q, int64_t nq, float max_range, int64_t* offsets, int64_t* ids,
                      float* dist_sq) {
  const Search* s = static_cast<Search*>(h);
  std::vector<Neighbor> tmp;
  int64_t total = 0;
  for (int64_t i = 0; i < nq; i++) {
    s->range(Vec3{{q[3 * i], q[3 * i + 1], q[3 * i + 2]}}, max_range, tmp);
    offsets[i] = total;
    if (ids) {
      for (size_t j = 0; j < tmp.size(); j++) {
        ids[total + (int64_t)j] = tmp[j].id;
        dist_sq[total + (int64_t)j] = tmp[j].dist_sq;
      }
    }
    total += (int64_t)tmp.size();
  }
  offsets[nq] = total;
}

// filter.VoxelGrid. mode 0 = literal dense array (the specification), 1 = sparse equivalent.
// `out` must hold n*stride bytes. cap_voxels bounds the dense array (entries of 32 B).
int32_t orc_voxelgrid_filter(const uint8_t* data, int64_t n, int64_t stride, const int64_t* off, const float* leaf,
                             const int64_t* chunk, int32_t mode, int64_t cap_voxels, uint8_t* out, int64_t* n_out) {
  Cloud c{data, n, stride, {off[0], off[1], off[2]}};
  std::vector<uint8_t> buf;
  int rc;
  if (mode == 0) {
    VoxelFilter f;
    f.leaf = {{leaf[0], leaf[1], leaf[2]}};
    f.cap_voxels = cap_voxels;
    rc = f.filter(c, chunk, buf, n_out);
  } else {
    rc = voxel_filter_sparse(c, Vec3{{leaf[0], leaf[1], leaf[2]}}, chunk, buf, n_out);
  }
  if (rc == ORC_OK && !buf.empty()) std::memcpy(out, buf.data(), buf.size());
  return rc;
}

void orc_minmax(const uint8_t* data, int64_t n, int64_t stride, const int64_t* off, float* mn, float* mx) {
  Cloud c{data, n, stride, {off[0], off[1], off[2]}};
  Vec3 a, b;
  minmax_vec3(c, &a, &b);
  for (int k = 0; k < 3; k++) {
    mn[k] = a[k];
    mx[k] = b[k];
  }
}

// NearestPointCorresponder.Pairs; arrays sized n; returns the number of pairs
int64_t orc_icp_pairs(void* base, const float* target, int64_t n, float max_dist, int64_t* base_id,
                      int64_t* target_id, float* dsq) {
  std::vector<Pair> pairs;
  icp_pairs(*static_cast<Search*>(base), target, n, max_dist, pairs);
  for (size_t i = 0; i < pairs.size(); i++) {
    base_id[i] = pairs[i].base_id;
    target_id[i] = pairs[i].target_id;
    dsq[i] = pairs[i].dsq;
  }
  return (int64_t)pairs.size();
}

// PointToPointEvaluator.Evaluate; out8 = {Value, Gradient[6], DistRMS}
int32_t orc_icp_evaluate_w(void* base, const float* target, int64_t n, float max_dist, int32_t min_pairs,
                           int32_t f64_accumulate, int32_t weight_fn, float weight_param, float* out8,
                           int64_t* n_pairs) {
  Evaluated ev{};
  int rc = f64_accumulate ? icp_evaluate_t<double>(*static_cast<Search*>(base), target, n, max_dist, min_pairs, &ev,
                                                   n_pairs, weight_fn, weight_param)
                          : icp_evaluate_t<float>(*static_cast<Search*>(base), target, n, max_dist, min_pairs, &ev,
                                                  n_pairs, weight_fn, weight_param);
  if (rc == ORC_OK) {
    out8[0] = ev.value;
    for (int i = 0; i < 6; i++) out8[1 + i] = ev.gradient[i];
    out8[7] = ev.dist_rms;
  }
  return rc;
}

int32_t orc_icp_evaluate(void* base, const float* target, int64_t n, float max_dist, int32_t min_pairs,
                         int32_t f64_accumulate, float* out8, int64_t* n_pairs) {
  Evaluated ev{};
  int rc = f64_accumulate
               ? icp_evaluate_t<double>(*static_cast<Search*>(base), target, n, max_dist, min_pairs, &ev, n_pairs)
               : icp_evaluate_t<float>(*static_cast<Search*>(base), target, n, max_dist, min_pairs, &ev, n_pairs);
  if (rc == ORC_OK) {
    out8[0] = ev.value;
    for (int i = 0; i < 6; i++) out8[1 + i] = ev.gradient[i];
    out8[7] = ev.dist_rms;
  }
  return rc;
}

// gradientDescentUpdater.Update applied once from updater state i; returns converged flag
int32_t orc_icp_update(const IcpParams* prm, int32_t i, float* trans16, const float* ev8) {
  Updater up(*prm);
  up.i = i;
  Evaluated ev;
  ev.value = ev8[0];
  for (int k = 0; k < 6; k++) ev.gradient[k] = ev8[1 + k];
  ev.dist_rms = ev8[7];
  Mat4 t;
  std::memcpy(t.m, trans16, sizeof(t.m));
  bool c = up.update(&t, ev);
  std::memcpy(trans16, t.m, sizeof(t.m));
  return c ? 1 : 0;
}

// PointToPointICPGradient.Fit; stat9 = {Value, Gradient[6], DistRMS, NumIteration}
int32_t orc_icp_fit(void* base, const float* target, int64_t n, const IcpParams* prm, float* trans16,
                    float* stat_ev8, int32_t* num_iteration) {
  Mat4 t;
  IcpStat st;
  int rc = icp_fit(*static_cast<Search*>(base), target, n, *prm, &t, &st);
  std::memcpy(trans16, t.m, sizeof(t.m));
  stat_ev8[0] = st.ev.value;
  for (int i = 0; i < 6; i++) stat_ev8[1 + i] = st.ev.gradient[i];
  stat_ev8[7] = st.ev.dist_rms;
  *num_iteration = st.num_iteration;
  return rc;
}

// Extension check: normal equations of one Evaluate. hess36 = 2f * sum J^T J (float32, index col*6+row),
// b6 = sum J^T r (float64).
int32_t orc_icp_normal_equations(void* base, const float* target, int64_t n, float max_dist, int32_t min_pairs,
                                 float* hess36, double* b6, int64_t* n_pairs) {
  double A[36];
  int64_t np = 0;
  int rc = icp_normal_equations(*static_cast<Search*>(base), target, n, max_dist, min_pairs, A, b6, &np);
  if (n_pairs) *n_pairs = np;
  float sw = (float)np, f = 1;
  if (sw > 1) f = 1 / sw;
  for (int i = 0; i < 36; i++) hess36[i] = (float)(A[i] * 2.0 * (double)f);
  return rc;
}

// Extension check: Fit with the Gauss-Newton updater.
int32_t orc_icp_fit_gn(void* base, const float* target, int64_t n, const IcpParams* prm, float* trans16,
                       float* stat_ev8, int32_t* num_iteration) {
  Mat4 t;
  IcpStat st;
  int rc = icp_fit_gn(*static_cast<Search*>(base), target, n, *prm, &t, &st);
  std::memcpy(trans16, t.m, sizeof(t.m));
  stat_ev8[0] = st.ev.value;
  for (int i = 0; i < 6; i++) stat_ev8[1 + i] = st.ev.gradient[i];
  stat_ev8[7] = st.ev.dist_rms;
  *num_iteration = st.num_iteration;
  return rc;
}

// RegionGrowing.Segment; out must hold n entries; returns the number of indices (BFS order)
int64_t orc_region_growing_segment(void* search, const uint32_t* label, const float* p, float max_range,
                                   int64_t* out) {
  std::vector<int64_t> ind;
  region_growing_segment(*static_cast<Search*>(search), label, Vec3{{p[0], p[1], p[2]}}, max_range, ind);
  std::memcpy(out, ind.data(), ind.size() * sizeof(int64_t));
  return (int64_t)ind.size();
}

// mat helpers for the golden tests
void orc_mat4_mul(const float* a, const float* b, float* out) {
  Mat4 A, B;
  std::memcpy(A.m, a, 64);
  std::memcpy(B.m, b, 64);
  Mat4 C = m4mul(A, B);
  std::memcpy(out, C.m, 64);
}
void orc_mat4_transform(const float* m, const float* xyz, int64_t n, float* out) {
  Mat4 M;
  std::memcpy(M.m, m, 64);
  for (int64_t i = 0; i < n; i++) {
    Vec3 t = m4transform(M, Vec3{{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}});
    out[3 * i] = t[0];
    out[3 * i + 1] = t[1];
    out[3 * i + 2] = t[2];
  }
}
void orc_translate(float x, float y, float z, float* out) {
  Mat4 M = m4translate(x, y, z);
  std::memcpy(out, M.m, 64);
}
void orc_rotate(float x, float y, float z, float ang, float* out) {
  Mat4 M = m4rotate(x, y, z, ang);
  std::memcpy(out, M.m, 64);
}
void orc_rodrigues(const float* v, float* out) {
  Mat4 M = rodrigues_to_rotation(Vec3{{v[0], v[1], v[2]}});
  std::memcpy(out, M.m, 64);
}
float orc_norm_sq(const float* v) { return vnormsq(Vec3{{v[0], v[1], v[2]}}); }

}  // extern "C"
