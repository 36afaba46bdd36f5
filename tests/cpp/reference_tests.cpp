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
  std::printf(failures ? "FAILED (%d)\n" : "ok (%d failures)\n", failures);
  return failures ? 1 : 0;
}
