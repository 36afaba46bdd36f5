// pcgol_b200.hpp — header-only C++ host mirror of the pcgol interfaces on the hot path,
// over the C ABI of pcgol_b200.h.  The reference is Go and this image has no Go toolchain,
// so the host side that a Go program would reach through cgo (go/pcgolgpu, source only) is
// mirrored here in C++ with the same names, argument meaning and error behaviour:
//
//   pcgol::pc::Vec3Slice, pcgol::pc::PointCloud          pc/vec3slice.go:8, pc/pointcloud.go:72-78
//   pcgol::storage::Neighbor, pcgol::storage::Search     pc/storage/search.go:8-17
//   pcgol::storage::Index  (replaces kdtree.New)         pc/storage/kdtree/kdtree.go:33,83,148
//   pcgol::filter::VoxelGrid                             pc/filter/voxelgrid/voxelgrid.go:23-35
//   pcgol::icp::{NearestPointCorresponder, PointToPointEvaluator,
//                GradientDescentUpdaterFactory, PointToPointICPGradient, Stat, Evaluated}
//                                                        pc/registration/icp/*.go
//   pcgol::storage::Index::{MinDistSq, With, DeletePoint} kdtree.go:19-22,58-65,322-332
//   pcgol::segmentation::RegionGrowing                   pc/segmentation/regiongrowing/regiongrowing.go:11-56
//   pcgol::pc::{Unmarshal, Marshal, DeviceCloud}         pc/io.go:32-45,232-285 (+ a cloud resident in HBM)
//   pcgol::icp::GaussNewtonUpdaterFactory                not in the reference: consumes Evaluated.Hessian
//
// Go returns (value, error); here errors are exceptions carrying the C status:
//   pcgol::Error                    any failure (status(), what())
//   pcgol::icp::ErrNotEnoughPairs   evaluator.go:15-17 (carries the partial trans/stat like icp.go:51-53)
//   pcgol::pc::ErrNoPoint           minmax.go:10-12
#ifndef PCGOL_B200_HPP_
#define PCGOL_B200_HPP_

#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pcgol_b200.h"

namespace pcgol {

class Error : public std::runtime_error {
 public:
  Error(pcg_status s, const std::string& msg) : std::runtime_error(msg), status_(s) {}
  pcg_status status() const { return status_; }

 private:
  pcg_status status_;
};

inline void check(pcg_status s) {
  if (s != PCG_OK) {
    const char* m = pcg_last_error();
    throw Error(s, (m && *m) ? m : pcg_status_string(s));
  }
}

namespace mat {
using Vec3 = std::array<float, 3>;   // mat/vec3.go:8
using Vec6 = std::array<float, 6>;   // mat/vec6.go:3
using Mat4 = std::array<float, 16>;  // mat/mat4.go:8-10, column-major: index = col*4 + row
using Mat6 = std::array<float, 36>;  // mat/mat6.go:3
}  // namespace mat

namespace pc {

struct ErrNoPoint : Error {
  ErrNoPoint() : Error(PCG_E_NO_POINT, "no point") {}
};

using Vec3Slice = std::vector<mat::Vec3>;  // pc/vec3slice.go:8

// pc.PointCloud reduced to what the hot path reads: the interleaved record buffer and where
// x, y, z live in a record (pc/pointcloud.go:64-70,130-163).
struct PointCloud {
  std::vector<uint8_t> Data;
  int64_t Points = 0;
  int64_t Stride = 12;
  std::array<int64_t, 3> XYZOffset{{0, 4, 8}};
  int64_t Width = 0, Height = 1;
};

// A flat view of anything Vec3RandomAccessor-like (pc/randomaccess.go:7-12).
struct View {
  const void* data;
  int64_t n, stride;
  std::array<int64_t, 3> off;
};
inline View view(const Vec3Slice& v) { return View{v.data(), (int64_t)v.size(), 12, {{0, 4, 8}}}; }
inline View view(const PointCloud& p) { return View{p.Data.data(), p.Points, p.Stride, p.XYZOffset}; }

}  // namespace pc

namespace storage {

struct Neighbor {  // pc/storage/search.go:8-11
  int64_t ID;
  float DistSq;
};

// storage.Search (pc/storage/search.go:13-17)
class Search {
 public:
  virtual ~Search() {}
  virtual int64_t Len() const = 0;
  virtual Neighbor Nearest(const mat::Vec3& p, float maxRange) const = 0;
  virtual std::vector<Neighbor> Range(const mat::Vec3& p, float maxRange) const = 0;
};

// GPU index; `Index(ra)` is the drop-in for kdtree.New(ra).
class Index : public Search {
 public:
  explicit Index(const pc::View& v, int device = 0) {
    check(pcg_index_build(v.data, v.n, v.stride, v.off.data(), device, &h_));
  }
  explicit Index(const pc::Vec3Slice& ra, int device = 0) : Index(pc::view(ra), device) {}
  explicit Index(pcg_index* adopted) : h_(adopted) {}  // e.g. from pc::DeviceCloud::BuildIndex
  ~Index() override { pcg_index_free(h_); }

  // KDTree.MinDistSq (kdtree.go:19-22): > 0 turns Nearest / NearestBatch / the ICP correspondences into the
  // approximate search (exact answer, or a real point closer than sqrt(MinDistSq)).
  float MinDistSq = 0.f;
  // KDTree.DeletePoint (kdtree.go:322-332); an id outside [0, Len()-1] throws (the reference returns an error).
  void DeletePoint(int64_t pID) { check(pcg_index_delete_points(h_, &pID, 1)); }
  void DeletePoints(const std::vector<int64_t>& ids) { check(pcg_index_delete_points(h_, ids.data(), (int64_t)ids.size())); }
  Index(const Index&) = delete;
  Index& operator=(const Index&) = delete;

  int64_t Len() const override { return pcg_index_len(h_); }

  Neighbor Nearest(const mat::Vec3& p, float maxRange) const override {
    pcg_neighbor nb;
    check(pcg_index_nearest_approx(h_, p.data(), 1, 12, nullptr, maxRange, MinDistSq, &nb));
    return Neighbor{nb.id, nb.dist_sq};
  }
  std::vector<Neighbor> Range(const mat::Vec3& p, float maxRange) const override {
    std::vector<int64_t> off;
    return RangeBatch(pc::View{p.data(), 1, 12, {{0, 4, 8}}}, maxRange, &off);
  }
  // The batched calls the hot path uses.
  std::vector<Neighbor> NearestBatch(const pc::View& q, float maxRange) const {
    std::vector<pcg_neighbor> raw((size_t)q.n);
    check(pcg_index_nearest_approx(h_, q.data, q.n, q.stride, q.off.data(), maxRange, MinDistSq, raw.data()));
    std::vector<Neighbor> out((size_t)q.n);
    for (size_t i = 0; i < raw.size(); i++) out[i] = Neighbor{raw[i].id, raw[i].dist_sq};
    return out;
  }
  std::vector<Neighbor> RangeBatch(const pc::View& q, float maxRange, std::vector<int64_t>* offsets) const {
    offsets->assign((size_t)q.n + 1, 0);  // two-call protocol: count, then fill into caller-owned memory
    check(pcg_index_range_count(h_, q.data, q.n, q.stride, q.off.data(), maxRange, offsets->data()));
    const int64_t total = offsets->back();
    std::vector<pcg_neighbor> raw((size_t)total);
    check(pcg_index_range_fill(h_, q.data, q.n, q.stride, q.off.data(), maxRange, offsets->data(), raw.data()));
    std::vector<Neighbor> out((size_t)total);
    for (int64_t i = 0; i < total; i++) out[(size_t)i] = Neighbor{raw[(size_t)i].id, raw[(size_t)i].dist_sq};
    return out;
  }
  pcg_index* handle() const { return h_; }

 private:
  pcg_index* h_ = nullptr;
};

}  // namespace storage

namespace filter {

// filter.Filter (pc/filter/filter.go:7-9) + voxelgrid.New / WithChunkSize
class VoxelGrid {
 public:
  explicit VoxelGrid(const mat::Vec3& leafSize, std::array<int64_t, 3> chunkSize = {{0, 0, 0}}, int device = 0)
      : leaf_(leafSize), chunk_(chunkSize), device_(device) {}

  pc::PointCloud Filter(const pc::PointCloud& pp) const {
    pc::PointCloud out = pp;
    out.Data.assign((size_t)(pp.Points * pp.Stride), 0);
    int64_t n = 0;
    pcg_status s = pcg_voxelgrid_filter(pp.Data.data(), pp.Points, pp.Stride, pp.XYZOffset.data(), leaf_.data(),
                                        chunk_.data(), device_, out.Data.data(), &n);
    if (s == PCG_E_NO_POINT) throw pc::ErrNoPoint();
    check(s);
    out.Data.resize((size_t)(n * pp.Stride));
    out.Points = out.Width = n;  // voxelgrid.go:119-128
    out.Height = 1;
    return out;
  }

  // One large Filter over several GPUs (chunks are independent units, voxelgrid.go:102-116): points per chunk id of a
  // device-resident cloud, to cut [0, n_chunks) into one range per rank ...
  std::vector<int64_t> ChunkHistogramDev(const void* d_data, int64_t n, int64_t stride,
                                         const std::array<int64_t, 3>& xyzOff, int64_t sampleStep = 1,
                                         void* stream = nullptr) const {
    int64_t chunks = 0;
    check(pcg_voxelgrid_chunk_histogram_dev(d_data, n, stride, xyzOff.data(), leaf_.data(), chunk_.data(), device_,
                                            sampleStep, nullptr, 0, &chunks, stream));
    std::vector<int64_t> hist((size_t)chunks, 0);
    check(pcg_voxelgrid_chunk_histogram_dev(d_data, n, stride, xyzOff.data(), leaf_.data(), chunk_.data(), device_,
                                            sampleStep, hist.data(), chunks, &chunks, stream));
    return hist;
  }
  // ... and the Filter restricted to chunk ids [cidLo, cidHi): the ranks' outputs in rank order are Filter's output.
  int64_t FilterChunksDev(const void* d_data, int64_t n, int64_t stride, const std::array<int64_t, 3>& xyzOff,
                          int64_t cidLo, int64_t cidHi, void* d_out, void* stream = nullptr) const {
    int64_t m = 0;
    pcg_status s = pcg_voxelgrid_filter_chunks_dev(d_data, n, stride, xyzOff.data(), leaf_.data(), chunk_.data(), cidLo,
                                                   cidHi, device_, d_out, &m, stream);
    if (s == PCG_E_NO_POINT) throw pc::ErrNoPoint();
    check(s);
    return m;
  }

 private:
  mat::Vec3 leaf_;
  std::array<int64_t, 3> chunk_;
  int device_;
};

}  // namespace filter

namespace icp {

struct Evaluated {  // evaluator.go:25-30
  float Value = 0;
  mat::Vec6 Gradient{};
  mat::Mat6 Hessian{};
  float DistRMS = 0;
};
struct Stat {  // stat.go:3-6
  Evaluated Ev;
  int NumIteration = 0;
};

struct ErrNotEnoughPairs : Error {
  mat::Mat4 Trans{};
  Stat St;
  ErrNotEnoughPairs() : Error(PCG_E_NOT_ENOUGH_PAIRS, "not enough correspondence pairs") {}
};

struct PointToPointCorrespondence {  // correspondence.go:8-12
  int64_t BaseID, TargetID;
  float SquaredDistance;
};

struct NearestPointCorresponder {  // correspondence.go:18-37
  float MaxDist = 0;
  std::vector<PointToPointCorrespondence> Pairs(const storage::Index& base, const pc::View& target) const {
    std::vector<int64_t> b((size_t)target.n + 1), t((size_t)target.n + 1);
    std::vector<float> d((size_t)target.n + 1);
    int64_t m = 0;
    check(pcg_icp_pairs_approx(base.handle(), target.data, target.n, target.stride, target.off.data(), MaxDist,
                               base.MinDistSq, b.data(), t.data(), d.data(), &m));
    std::vector<PointToPointCorrespondence> out((size_t)m);
    for (int64_t i = 0; i < m; i++) out[(size_t)i] = {b[(size_t)i], t[(size_t)i], d[(size_t)i]};
    return out;
  }
};

inline Evaluated from_c(const pcg_evaluated& e) {
  Evaluated o;
  o.Value = e.value;
  std::memcpy(o.Gradient.data(), e.gradient, sizeof(e.gradient));
  std::memcpy(o.Hessian.data(), e.hessian, sizeof(e.hessian));
  o.DistRMS = e.dist_rms;
  return o;
}

struct PointToPointEvaluator {  // evaluator.go:69-189
  NearestPointCorresponder Corresponder;
  int MinPairs = 0;
  int Mode = PCG_ICP_STRICT;  // | PCG_ICP_WITH_HESSIAN to fill Evaluated.Hessian
  // PointToPointEvaluator.WeightFn (evaluator.go:19-23,72) is a closure in the reference; here one of pcg_weight_fn
  int WeightFn = PCG_WEIGHT_CONSTANT;
  float WeightParam = 0.f;
  bool HasGradient() const { return true; }
  bool HasHessian() const { return (Mode & PCG_ICP_WITH_HESSIAN) != 0; }  // false like evaluator.go:76 by default
  Evaluated Evaluate(const storage::Index& base, const pc::View& target) const {
    pcg_evaluated ev;
    int64_t np = 0;
    pcg_icp_params p;
    std::memset(&p, 0, sizeof(p));
    p.max_dist = Corresponder.MaxDist;
    p.min_pairs = MinPairs;
    p.mode = Mode;
    p.min_dist_sq = base.MinDistSq;
    p.weight_fn = WeightFn;
    p.weight_param = WeightParam;
    pcg_status s = pcg_icp_evaluate_params(base.handle(), target.data, target.n, target.stride, target.off.data(), &p,
                                           &ev, &np);
    if (s == PCG_E_NOT_ENOUGH_PAIRS) throw ErrNotEnoughPairs();
    check(s);
    return from_c(ev);
  }
};

struct GradientDescentUpdaterFactory {  // updater.go:18-37 ; zero == reference default
  mat::Vec6 Weight{};
  mat::Vec6 Threshold{};
  int MaxIteration = 0;
  int Kind = PCG_UPDATER_GRADIENT_DESCENT;
};
// Not in the reference: solves the normal equations accumulated with Evaluated.Hessian (the hook declared
// by evaluator.go:25-36) instead of a damped gradient step; same convergence test and MaxIteration cap.
struct GaussNewtonUpdaterFactory : GradientDescentUpdaterFactory {
  GaussNewtonUpdaterFactory() { Kind = PCG_UPDATER_GAUSS_NEWTON; }
};

struct PointToPointICPGradient {  // icp.go:18-67
  PointToPointEvaluator Evaluator;
  GradientDescentUpdaterFactory UpdaterFactory;

  std::pair<mat::Mat4, Stat> Fit(const storage::Index& base, const pc::View& target) const {
    pcg_icp_params p;
    std::memset(&p, 0, sizeof(p));
    p.max_dist = Evaluator.Corresponder.MaxDist;
    p.min_pairs = Evaluator.MinPairs;
    std::memcpy(p.weight, UpdaterFactory.Weight.data(), sizeof(p.weight));
    std::memcpy(p.threshold, UpdaterFactory.Threshold.data(), sizeof(p.threshold));
    p.max_iteration = UpdaterFactory.MaxIteration;
    p.mode = Evaluator.Mode;
    p.updater = UpdaterFactory.Kind;
    p.min_dist_sq = base.MinDistSq;
    p.weight_fn = Evaluator.WeightFn;
    p.weight_param = Evaluator.WeightParam;
    mat::Mat4 trans{};
    pcg_icp_stat st;
    pcg_status s = pcg_icp_fit(base.handle(), target.data, target.n, target.stride, target.off.data(), &p,
                               trans.data(), &st);
    Stat out;
    out.Ev = from_c(st.evaluated);
    out.NumIteration = st.num_iteration;
    if (s == PCG_E_NOT_ENOUGH_PAIRS) {
      ErrNotEnoughPairs e;
      e.Trans = trans;
      e.St = out;
      throw e;
    }
    check(s);
    return {trans, out};
  }
};

}  // namespace icp

namespace segmentation {

// regiongrowing.New(search, propertyIter) + Segment (regiongrowing.go:11-56): `labelOffset` is the byte offset
// of the uint32 property inside the records of `pp` (pc.Uint32Iterator(name)).
class RegionGrowing {
 public:
  RegionGrowing(const storage::Index& search, const pc::PointCloud& pp, int64_t labelOffset) : n_(pp.Points) {
    check(pcg_region_growing_new(search.handle(), pp.Data.data(), pp.Points, pp.Stride, pp.XYZOffset.data(),
                                 labelOffset, &h_));
  }
  ~RegionGrowing() { pcg_region_growing_free(h_); }
  RegionGrowing(const RegionGrowing&) = delete;
  RegionGrowing& operator=(const RegionGrowing&) = delete;
  std::vector<int64_t> Segment(const mat::Vec3& p, float maxRange) const {
    std::vector<int64_t> out((size_t)n_ + 1);
    int64_t m = 0;
    check(pcg_region_growing_segment(h_, p.data(), maxRange, out.data(), n_, &m));
    out.resize((size_t)m);
    return out;
  }

 private:
  pcg_region_growing* h_ = nullptr;
  int64_t n_;
};

}  // namespace segmentation

namespace pc {

// A pc.PointCloud whose Data stays in HBM: Unmarshal -> VoxelGrid -> BuildIndex -> ICP without host round trips.
class DeviceCloud {
 public:
  explicit DeviceCloud(pcg_cloud* h) : h_(h) {}
  ~DeviceCloud() { pcg_cloud_free(h_); }
  DeviceCloud(const DeviceCloud&) = delete;
  DeviceCloud& operator=(const DeviceCloud&) = delete;
  DeviceCloud(DeviceCloud&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  pcg_cloud_header Header() const {
    pcg_cloud_header h;
    check(pcg_cloud_get_header(h_, &h));
    return h;
  }
  std::vector<uint8_t> Download() const {
    std::vector<uint8_t> d((size_t)Header().data_bytes);
    check(pcg_cloud_download(h_, d.data(), (int64_t)d.size()));
    return d;
  }
  DeviceCloud VoxelGrid(const mat::Vec3& leaf, std::array<int64_t, 3> chunk = {{0, 0, 0}}) const {
    pcg_cloud* out = nullptr;
    pcg_status s = pcg_cloud_voxelgrid_filter(h_, leaf.data(), chunk.data(), &out);
    if (s == PCG_E_NO_POINT) throw ErrNoPoint();
    check(s);
    return DeviceCloud(out);
  }
  pcg_index* BuildIndex() const {  // wrap with storage::Index(handle)
    pcg_index* idx = nullptr;
    check(pcg_cloud_index_build(h_, &idx));
    return idx;
  }
  pcg_cloud* handle() const { return h_; }

 private:
  pcg_cloud* h_;
};

// pc.Unmarshal (io.go:32-45): errors keep the reference's classes through Error::status()
// (PCG_E_PCD_SYNTAX ~ strconv.ErrSyntax, PCG_E_PCD_EOF ~ io.EOF, PCG_E_PCD_CORRUPT ~ lzf.ErrDataCorruption).
inline DeviceCloud Unmarshal(const std::vector<uint8_t>& pcd, int device = 0) {
  pcg_cloud* c = nullptr;
  check(pcg_pcd_unmarshal(pcd.data(), (int64_t)pcd.size(), device, &c));
  return DeviceCloud(c);
}
// pc.Marshal (io.go:232-285): "DATA binary".
inline std::vector<uint8_t> Marshal(const DeviceCloud& c) {
  int64_t len = 0;
  pcg_pcd_marshal(c.handle(), nullptr, 0, &len);
  std::vector<uint8_t> out((size_t)len);
  check(pcg_pcd_marshal(c.handle(), out.data(), len, &len));
  return out;
}

}  // namespace pc
}  // namespace pcgol

#endif  // PCGOL_B200_HPP_
