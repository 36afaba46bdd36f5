// reference_tests.cpp — the reference's own table tests for the hot path, replayed through the
// C++ host mirror (include/pcgol_b200.hpp) on the GPU.  Each test names the Go test it follows.
// Build: g++ -O2 -ffp-contract=off -Iinclude tests/cpp/reference_tests.cpp -Lpcgol_b200 -lpcgol_b200
#include <cmath>
#include <cstdio>
#include <map>
#include <string>

#include "pcgol_b200.hpp"

using namespace pcgol;
using mat::Mat4;
using mat::Vec3;

static int failures = 0;
#define EXPECT(cond, ...)                         \
  do {                                            \
    if (!(cond)) {                                \
      failures++;                                 \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);                   \
      std::printf("\n");                          \
    }                                             \
  } while (0)

// mat helpers used by the Go tests to build inputs (mat/transform.go, mat/mat4.go)
static Mat4 Translate(float x, float y, float z) { return Mat4{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, x, y, z, 1}}; }
static Mat4 Rotate(float x, float y, float z, float ang) {
  float s = (float)std::sin((double)ang), c = (float)std::cos((double)ang);
  float o = 1 - c;
  return Mat4{{c + x * x * o, x * y * o + z * s, x * z * o - y * s, 0, y * x * o - z * s, c + y * y * o,
               y * z * o + x * s, 0, z * x * o + y * s, z * y * o - x * s, c + z * z * o, 0, 0, 0, 0, 1}};
}
static Mat4 Mul(const Mat4& m, const Mat4& a) {
  Mat4 out{};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float sum = 0;
      for (int k = 0; k < 4; k++) sum += m[4 * k + i] * a[4 * j + k];
      out[4 * j + i] = sum;
    }
  return out;
}
static Vec3 Transform(const Mat4& m, const Vec3& a) {
  float w = 1 / (m[3] * a[0] + m[7] * a[1] + m[11] * a[2] + m[15]);
  return Vec3{{(m[0] * a[0] + m[4] * a[1] + m[8] * a[2] + m[12]) * w, (m[1] * a[0] + m[5] * a[1] + m[9] * a[2] + m[13]) * w,
               (m[2] * a[0] + m[6] * a[1] + m[10] * a[2] + m[14]) * w}};
}
static float NormSq(const Vec3& v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
static Vec3 Sub(const Vec3& a, const Vec3& b) { return Vec3{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
static Vec3 Add(const Vec3& a, const Vec3& b) { return Vec3{{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }

// pc/storage/kdtree/kdtree_test.go:119-246 (TestKDtree / Nearest)
static void TestKDtree_Nearest() {
  pc::Vec3Slice pts{{{4, 1, 0}}, {{2, 2, 1}}, {{5, 0, 0}}, {{3, 0, 0}}, {{0, 1, 0}}, {{1, 0, 0}}, {{6, 2, 1}}};
  storage::Index kdt(pts);
  struct Case {
    Vec3 p;
    int64_t nodeID;
    float distSq, maxRange;
  } cases[] = {
      {{{5, 0, 0}}, 2, 0, 1.0f},          {{{5, 0, 0.1f}}, 2, 0.1f * 0.1f, 1.0f},  {{{4.9f, 0, 0}}, 2, 0.1f * 0.1f, 1.0f},
      {{{3, 0, 0}}, 3, 0, 1.0f},          {{{3, 0, 0.1f}}, 3, 0.1f * 0.1f, 1.0f},  {{{2.1f, 1.9f, 1}}, 1, 2 * 0.1f * 0.1f, 1.0f},
      {{{2.1f, 2.1f, 1}}, 1, 2 * 0.1f * 0.1f, 1.0f}, {{{3.9f, 1, 0}}, 0, 0.1f * 0.1f, 1.0f},
      {{{4.1f, 1, 0}}, 0, 0.1f * 0.1f, 1.0f}, {{{4.2f, 1, 0}}, -1, 0.1f * 0.1f, 0.1f},
  };
  const float eps = 0.00001f;
  for (auto& tt : cases) {
    storage::Neighbor nb = kdt.Nearest(tt.p, tt.maxRange);
    EXPECT(nb.ID == tt.nodeID, "Expected id: %lld, got: %lld", (long long)tt.nodeID, (long long)nb.ID);
    EXPECT(!(nb.DistSq < tt.distSq - eps || tt.distSq + eps < nb.DistSq), "Expected distance^2: %0.4f, got: %0.4f",
           tt.distSq, nb.DistSq);
  }
}

// kdtree_test.go:281-386 (TestKDtree / Range)
static void TestKDtree_Range() {
  pc::Vec3Slice pts{{{0.0f, 0.2f, 0.0f}}, {{3.0f, 0, 0}}, {{0.2f, 0, 0}}, {{0, 1.0f, 0}},
                    {{0, 0, 5.0f}},       {{0.5f, 0, 0}}, {{0, 0, 0.4f}}};
  storage::Index kdt(pts);
  struct Case {
    Vec3 p;
    float maxRange;
    std::vector<storage::Neighbor> neighbors;
  } cases[] = {
      {{{10, 10, 10}}, 1, {}},
      {{{0, 0.2f, 0}}, 0.05f, {{0, 0.0f}}},
      {{{0, 0.2f, 0}}, 0.3f, {{0, 0.0f}, {2, 0.08f}}},
      {{{0, 0.2f, 0}}, 0.45f, {{0, 0.0f}, {2, 0.08f}, {6, 0.2f}}},
      {{{0, 0.2f, 0}}, 0.6f, {{0, 0.0f}, {2, 0.08f}, {6, 0.2f}, {5, 0.29f}}},
  };
  const float eps = 0.00001f;
  for (auto& tt : cases) {
    auto got = kdt.Range(tt.p, tt.maxRange);
    EXPECT(got.size() == tt.neighbors.size(), "Expected number of neighbors: %zu, got %zu", tt.neighbors.size(),
           got.size());
    for (size_t i = 0; i < got.size() && i < tt.neighbors.size(); i++) {
      EXPECT(got[i].ID == tt.neighbors[i].ID, "Expected ID: %lld, got %lld", (long long)tt.neighbors[i].ID,
             (long long)got[i].ID);
      EXPECT(!(got[i].DistSq < tt.neighbors[i].DistSq - eps || tt.neighbors[i].DistSq + eps < got[i].DistSq),
             "Expected distance^2: %0.4f, got: %0.4f", tt.neighbors[i].DistSq, got[i].DistSq);
    }
  }
}

// pc/filter/voxelgrid/voxelgrid_test.go:12-110 (TestVoxelGrid)
static void TestVoxelGrid() {
  struct Rec {
    float x, y, z;
    uint32_t label;
  };
  const Rec in[6] = {{0.625f, 1.875f, 0.125f, 1}, {1.250f, 1.250f, 1.250f, 2}, {0.650f, 1.875f, 0.150f, 3},
                     {1.250f, 0.000f, 1.250f, 4}, {1.250f, 1.275f, 1.250f, 5}, {0.000f, 3.000f, 0.000f, 6}};
  struct Case {
    const char* name;
    std::array<int64_t, 3> chunk;
    Vec3 expected[4];
    uint32_t labels[4];
  } cases[] = {
      {"Default", {{0, 0, 0}}, {{{0, 3, 0}}, {{0.6375f, 1.875f, 0.1375f}}, {{1.25f, 0, 1.25f}}, {{1.25f, 1.2625f, 1.25f}}}, {6, 1, 4, 2}},
      {"WithChunkSize881", {{8, 8, 1}}, {{{0, 3, 0}}, {{0.6375f, 1.875f, 0.1375f}}, {{1.25f, 0, 1.25f}}, {{1.25f, 1.2625f, 1.25f}}}, {6, 1, 4, 2}},
      {"WithChunkSize333", {{3, 3, 3}}, {{{0.6375f, 1.875f, 0.1375f}}, {{0, 3, 0}}, {{1.25f, 0, 1.25f}}, {{1.25f, 1.2625f, 1.25f}}}, {1, 6, 4, 2}},
  };
  for (auto& tt : cases) {
    pc::PointCloud pp;
    pp.Stride = 16;
    pp.Points = pp.Width = 6;
    pp.Data.resize(sizeof(in));
    std::memcpy(pp.Data.data(), in, sizeof(in));
    filter::VoxelGrid vg(Vec3{{0.125f, 0.125f, 0.125f}}, tt.chunk);
    pc::PointCloud out = vg.Filter(pp);
    EXPECT(out.Points == 4, "%s: Wrong number of points, expected: 4, got: %lld", tt.name, (long long)out.Points);
    const Rec* r = reinterpret_cast<const Rec*>(out.Data.data());
    for (int i = 0; i < 4 && i < out.Points; i++) {
      EXPECT(r[i].x == tt.expected[i][0] && r[i].y == tt.expected[i][1] && r[i].z == tt.expected[i][2],
             "%s: Expected point: {%g %g %g}, got: {%g %g %g}", tt.name, tt.expected[i][0], tt.expected[i][1],
             tt.expected[i][2], r[i].x, r[i].y, r[i].z);
      EXPECT(r[i].label == tt.labels[i], "%s: Expected label: %x, got: %x", tt.name, tt.labels[i], r[i].label);
    }
  }
  bool threw = false;
  try {
    filter::VoxelGrid(Vec3{{0.1f, 0.1f, 0.1f}}).Filter(pc::PointCloud());
  } catch (const pc::ErrNoPoint&) {
    threw = true;  // pc/minmax.go:10-12
  }
  EXPECT(threw, "empty cloud must fail with \"no point\"");
}

// pc/registration/icp/correspondence_test.go:12-37
static void TestNearestPointCorresponder() {
  pc::Vec3Slice base{{{4, 1, 0}}, {{1, 1, 0}}, {{8, 1, 1}}, {{-5, 0, 1}}, {{0, 1, 0}}};
  storage::Index kdt(base);
  icp::NearestPointCorresponder corr{3};
  pc::Vec3Slice targets{{{8, 1, 1}}, {{-8, 1, 1}}, {{2, 1, 0}}};
  auto pairs = corr.Pairs(kdt, pc::view(targets));
  EXPECT(pairs.size() == 2, "Expected 2 pairs, got %zu", pairs.size());
  if (pairs.size() == 2) {
    EXPECT(pairs[0].BaseID == 2 && pairs[0].TargetID == 0 && pairs[0].SquaredDistance == 0, "pair 0 differs");
    EXPECT(pairs[1].BaseID == 1 && pairs[1].TargetID == 2 && pairs[1].SquaredDistance == 1, "pair 1 differs");
  }
}

// pc/registration/icp/evaluator_test.go:11-77 (Value equality part)
static void TestPointToPointEvaluator() {
  pc::Vec3Slice base{{{0, 0, 0}}, {{1, 1, 0}}, {{2, 2, 0}}, {{3, 1, 1}}, {{4, 0, 0}}};
  Vec3 delta{{0.25f, 0.125f, -0.125f}};
  pc::Vec3Slice target{Add(base[2], delta), Add(base[3], delta), Add(base[4], delta)};
  storage::Index kdt(base);
  icp::PointToPointEvaluator e{icp::NearestPointCorresponder{2}, 3};
  icp::Evaluated ev = e.Evaluate(kdt, pc::view(target));
  EXPECT(ev.Value == NormSq(delta), "Expected evaluated value: %f, got: %f", NormSq(delta), ev.Value);
  EXPECT(e.HasGradient() && !e.HasHessian(), "HasGradient/HasHessian");
  bool threw = false;
  try {
    icp::PointToPointEvaluator{icp::NearestPointCorresponder{2}, 0}.Evaluate(kdt, pc::view(target));
  } catch (const icp::ErrNotEnoughPairs&) {
    threw = true;  // MinPairs 0 -> 6
  }
  EXPECT(threw, "3 pairs < default MinPairs 6 must fail with ErrNotEnoughPairs");
}

// pc/registration/icp/icp_test.go:13-98 (TestPointToPointICPGradient; exact search instead of MinDistSq=0.01)
static void TestPointToPointICPGradient() {
  for (float zoff : {0.0f, 5.0f}) {
    pc::Vec3Slice base{{{-2.1f, 0, zoff}}, {{-1, 1, zoff}}, {{0, 2, zoff}}, {{1, 1, 1 + zoff}}, {{2, 0, zoff}}};
    std::map<std::string, Mat4> deltas{
        {"Trans(0,0,0)", Translate(0, 0, 0)},
        {"Trans(0.25,0.125,-0.125)", Translate(0.25f, 0.125f, -0.125f)},
        {"Trans(0.5,0.5,1)", Translate(0.5f, 0.5f, 1.0f)},
        {"Trans(-0.5,-0.5,0)", Translate(-0.5f, -0.5f, 0.0f)},
        {"Rot(1,0,0,0.2)", Rotate(1, 0, 0, 0.2f)},
        {"Rot(1,0,0,-0.2)", Rotate(1, 0, 0, -0.2f)},
        {"Rot(1,0,0,0.1)Trans(0.2,0,0)", Mul(Rotate(1, 0, 0, 0.1f), Translate(0.2f, 0, 0))},
        {"Rot(1,0,0,0.1)Trans(-0.2,0,0)", Mul(Rotate(1, 0, 0, 0.1f), Translate(-0.2f, 0, 0))},
        {"Trans(0.2,0,0)Rot(1,0,0,0.1)", Mul(Translate(0.2f, 0, 0), Rotate(1, 0, 0, 0.1f))},
        {"Trans(-0.2,0,0)Rot(1,0,0,0.1)", Mul(Translate(-0.2f, 0, 0), Rotate(1, 0, 0, 0.1f))},
        {"Rot(0,1,0,0.1)Trans(0.2,0,0)", Mul(Rotate(0, 1, 0, 0.1f), Translate(0.2f, 0, 0))},
        {"Rot(0,1,0,0.1)Trans(-0.2,0,0)", Mul(Rotate(0, 1, 0, 0.1f), Translate(-0.2f, 0, 0))},
        {"Trans(0.2,0,0)Rot(0,1,0,0.1)", Mul(Translate(0.2f, 0, 0), Rotate(0, 1, 0, 0.1f))},
        {"Trans(-0.2,0,0)Rot(0,1,0,0.1)", Mul(Translate(-0.2f, 0, 0), Rotate(0, 1, 0, 0.1f))},
    };
    storage::Index kdt(base);
    for (auto& kv : deltas) {
      const int indices[5] = {3, 1, 4, 0, 2};
      pc::Vec3Slice target(5);
      for (int i = 0; i < 5; i++) target[i] = Transform(kv.second, base[indices[i]]);
      icp::PointToPointICPGradient ppicp;
      ppicp.Evaluator = icp::PointToPointEvaluator{icp::NearestPointCorresponder{2}, 3};
      auto fit = ppicp.Fit(kdt, pc::view(target));
      float residual = 0;
      for (int i = 0; i < 5; i++) residual += NormSq(Sub(Transform(fit.first, target[i]), base[indices[i]]));
      residual /= 5.0f;
      EXPECT(0.05f >= residual, "%s (z+%g): residual %f", kv.first.c_str(), zoff, residual);
      EXPECT(fit.second.NumIteration >= 1 && fit.second.NumIteration <= 20, "NumIteration %d", fit.second.NumIteration);
    }
  }
}


// pc/storage/kdtree/kdtree_test.go:413-751 (DeletePoint): searches after deletion behave like naiveSearch.deletePoint
static void TestKDtree_DeletePoint() {
  pc::Vec3Slice pts{{{4, 1, 0}}, {{2, 2, 1}}, {{5, 0, 0}}, {{3, 0, 0}}, {{0, 1, 0}}, {{1, 0, 0}}, {{6, 2, 1}}};
  storage::Index kdt(pts);
  kdt.DeletePoint(5);
  EXPECT(kdt.Nearest(pts[5], 0.001f).ID < 0, "point 5 was not deleted");
  EXPECT(kdt.Nearest(pts[5], 1.5f).ID == 4, "nearest to deleted point 5 must be 4");
  kdt.DeletePoint(5);  // "TwiceTheSamePoint": no error
  EXPECT(kdt.Len() == 7, "Len() is the accessor's");
  bool threw = false;
  try {
    kdt.DeletePoint(123);  // "InvalidPointID"
  } catch (const Error& e) {
    threw = e.status() == PCG_E_INVALID_ARG && std::string(e.what()).find("123 does not correspond") != std::string::npos;
  }
  EXPECT(threw, "expected an error when deleting a point that is not in the tree");
  // kdtree_test.go:731-751: all points on a line
  pc::Vec3Slice line{{{4, 0, 0}}, {{1, 0, 0}}, {{2, 0, 0}}, {{3, 0, 0}}};
  storage::Index k2(line);
  for (int i = 0; i < 4; i++) {
    k2.DeletePoint(i);
    EXPECT(k2.Nearest(line[i], 0.001f).ID < 0, "%d was not deleted", i);
  }
}

// pc/registration/icp/icp_test.go:13-98 exactly as the reference runs it: kdtree.New(ra) with MinDistSq = 0.01
static void TestPointToPointICPGradient_MinDistSq() {
  pc::Vec3Slice base{{{-2.1f, 0, 0}}, {{-1, 1, 0}}, {{0, 2, 0}}, {{1, 1, 1}}, {{2, 0, 0}}};
  storage::Index kdt(base);
  kdt.MinDistSq = 0.01f;
  const int indices[5] = {3, 1, 4, 0, 2};
  for (const Mat4& delta : {Translate(0.25f, 0.125f, -0.125f), Rotate(1, 0, 0, 0.2f), Mul(Translate(0.2f, 0, 0), Rotate(0, 1, 0, 0.1f))}) {
    pc::Vec3Slice target(5);
    for (int i = 0; i < 5; i++) target[i] = Transform(delta, base[indices[i]]);
    for (int gn = 0; gn < 2; gn++) {  // the reference's updater, then the Gauss-Newton one
      icp::PointToPointICPGradient ppicp;
      ppicp.Evaluator = icp::PointToPointEvaluator{icp::NearestPointCorresponder{2}, 3};
      if (gn) ppicp.UpdaterFactory = icp::GaussNewtonUpdaterFactory();
      auto fit = ppicp.Fit(kdt, pc::view(target));
      float residual = 0;
      for (int i = 0; i < 5; i++) residual += NormSq(Sub(Transform(fit.first, target[i]), base[indices[i]]));
      EXPECT(0.05f >= residual / 5.0f, "MinDistSq fit (gn=%d): residual %f", gn, residual / 5.0f);
    }
  }
  icp::PointToPointEvaluator e{icp::NearestPointCorresponder{2}, 3, PCG_ICP_STRICT | PCG_ICP_WITH_HESSIAN};
  icp::Evaluated ev = e.Evaluate(kdt, pc::view(base));
  EXPECT(e.HasHessian() && ev.Hessian[0] == 2.0f && ev.Hessian[7] == 2.0f, "H_tt = 2f*N*I = 2 (got %f)", ev.Hessian[0]);
}

// pc/segmentation/regiongrowing/regiongrowing_test.go (two labelled clusters joined by a chain of another label)
static void TestRegionGrowingSegment() {
  pc::PointCloud pp;
  pp.Stride = 16;
  const int n = 30;
  pp.Points = pp.Width = n;
  pp.Data.assign((size_t)n * 16, 0);
  for (int i = 0; i < n; i++) {
    const float xyz[3] = {0.1f * i, 0, 0};
    const uint32_t label = (i >= 10 && i < 20) ? 7u : 1u;  // 1 1 1 ... 7 7 7 ... 1 1 1
    std::memcpy(&pp.Data[(size_t)i * 16], xyz, 12);
    std::memcpy(&pp.Data[(size_t)i * 16 + 12], &label, 4);
  }
  storage::Index kdt(pc::view(pp));
  segmentation::RegionGrowing rg(kdt, pp, 12);
  auto a = rg.Segment(Vec3{{0.0f, 0, 0}}, 0.15f);
  EXPECT(a.size() == 10, "first label-1 run: expected 10 points, got %zu", a.size());
  for (int64_t id : a) EXPECT(id < 10, "index %lld leaked across the label-7 run", (long long)id);
  auto b = rg.Segment(Vec3{{1.5f, 0, 0}}, 0.15f);
  EXPECT(b.size() == 10 && b[0] == 15, "label-7 run: expected 10 points starting at the seed's nearest, got %zu", b.size());
  EXPECT(rg.Segment(Vec3{{10, 10, 10}}, 0.15f).empty(), "no neighbour -> empty result");
}

// pc/io_test.go:16-255 ("Binary" vector written by Marshal: Unmarshal -> Marshal reproduces it; error classes)
static void TestUnmarshalMarshal() {
  const char* head =
      "VERSION 0.7\nFIELDS x y z label\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT 1\n"
      "VIEWPOINT 0.0000 0.0000 0.0000 1.0000 0.0000 0.0000 0.0000\nPOINTS 2\nDATA binary\n";
  std::vector<uint8_t> pcd(head, head + std::strlen(head));
  const float pts[2][3] = {{0.352f, -0.151f, -0.106f}, {-0.473f, 0.292f, -0.731f}};
  for (int i = 0; i < 2; i++) {
    const uint32_t label = (uint32_t)i + 3;
    pcd.insert(pcd.end(), (const uint8_t*)pts[i], (const uint8_t*)pts[i] + 12);
    pcd.insert(pcd.end(), (const uint8_t*)&label, (const uint8_t*)&label + 4);
  }
  pc::DeviceCloud dc = pc::Unmarshal(pcd);
  pcg_cloud_header h = dc.Header();
  EXPECT(h.points == 2 && h.n_fields == 4 && std::string(h.fields[3]) == "label" && h.width == 2, "header");
  EXPECT(pc::Marshal(dc) == pcd, "Marshal(Unmarshal(x)) != x");
  auto status_of = [](const char* text) {
    try {
      pc::Unmarshal(std::vector<uint8_t>(text, text + std::strlen(text)));
    } catch (const Error& e) {
      return (int)e.status();
    }
    return 0;
  };
  EXPECT(status_of("VERSION X") == PCG_E_PCD_SYNTAX, "ErrorVersion -> strconv.ErrSyntax");
  EXPECT(status_of("WIDTH X") == PCG_E_PCD_SYNTAX, "ErrorWidth -> strconv.ErrSyntax");
  EXPECT(status_of("VERSION 0.7\nFIELDS x\nSIZE 4\nTYPE F\nCOUNT 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary_compressed\n") ==
             PCG_E_PCD_EOF, "ErrorBinaryCompressedNCompressedEOF -> io.EOF");
  // resident pipeline: VoxelGrid -> index -> Nearest
  pc::DeviceCloud small = dc.VoxelGrid(Vec3{{10, 10, 10}}, {{4, 4, 4}});
  EXPECT(small.Header().points == 1 && small.Header().width == 1 && small.Header().height == 1, "one voxel");
  storage::Index idx(dc.BuildIndex());
  EXPECT(idx.Nearest(Vec3{{0.35f, -0.15f, -0.1f}}, 1.0f).ID == 0, "index over the resident cloud");
}

int main() {
  if (pcg_device_count() < 1) {
    std::printf("no CUDA device: %s\n", pcg_last_error());
    return 2;
  }
  TestKDtree_Nearest();
  TestKDtree_Range();
  TestVoxelGrid();
  TestNearestPointCorresponder();
  TestPointToPointEvaluator();
  TestPointToPointICPGradient();
  TestKDtree_DeletePoint();
  TestPointToPointICPGradient_MinDistSq();
  TestRegionGrowingSegment();
  TestUnmarshalMarshal();
  std::printf(failures ? "FAILED (%d)\n" : "ok (%d failures)\n", failures);
  return failures ? 1 : 0;
}
