// Inputs: tools/nn_model_dump.py writes /tmp/nnm/tgt.bin (target cloud) and /tmp/nnm/q.bin (query sample).
// Build: g++ -O2 -o /tmp/nnm/model tools/nn_model.cpp ; run: /tmp/nnm/model [queries]   (PERFECT=1 adds the perfect-bound floor)
// CPU model of the GPU index: counts node steps / leaf scans per query for different point orders.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>
using namespace std;
struct P3 { float x, y, z; };
static vector<P3> readf(const char* p) {
  FILE* f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  vector<P3> v(n / 12); if (fread(v.data(), 12, v.size(), f) != v.size()) abort(); fclose(f); return v;
}
static uint64_t spread16(uint32_t x) { uint64_t v = x & 0xffff; v = (v | (v << 16)) & 0x0000ff0000ffull; v = (v | (v << 8)) & 0x00f00f00f00full; v = (v | (v << 4)) & 0x0c30c30c30c3ull; v = (v | (v << 2)) & 0x249249249249ull; return v; }
static uint64_t hilbert(uint32_t x0, uint32_t x1, uint32_t x2, int BITS) {
  uint32_t X[3] = {x0, x1, x2}; uint32_t M = 1u << (BITS - 1);
  for (uint32_t Q = M; Q > 1; Q >>= 1) { uint32_t Pm = Q - 1; for (int i = 0; i < 3; i++) { if (X[i] & Q) X[0] ^= Pm; else { uint32_t t = (X[0] ^ X[i]) & Pm; X[0] ^= t; X[i] ^= t; } } }
  X[1] ^= X[0]; X[2] ^= X[1]; uint32_t t = 0; for (uint32_t Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1; X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return (spread16(X[0]) << 2) | (spread16(X[1]) << 1) | spread16(X[2]);
}
struct Tree { int L; uint32_t leaves, P; vector<P3> pts; vector<uint32_t> id; vector<P3> lo, hi; };
static void build_boxes(Tree& t) {
  t.lo.assign(2 * t.P, P3{INFINITY, INFINITY, INFINITY}); t.hi.assign(2 * t.P, P3{-INFINITY, -INFINITY, -INFINITY});
  for (uint32_t l = 0; l < t.leaves; l++) { P3 lo{INFINITY, INFINITY, INFINITY}, hi{-INFINITY, -INFINITY, -INFINITY};
    for (int j = 0; j < t.L; j++) { size_t i = (size_t)l * t.L + j; if (i >= t.pts.size()) break; P3 p = t.pts[i]; lo.x = min(lo.x, p.x); lo.y = min(lo.y, p.y); lo.z = min(lo.z, p.z); hi.x = max(hi.x, p.x); hi.y = max(hi.y, p.y); hi.z = max(hi.z, p.z); }
    t.lo[t.P + l] = lo; t.hi[t.P + l] = hi; }
  for (uint32_t k = t.P - 1; k >= 1; k--) { P3 a = t.lo[2 * k], b = t.lo[2 * k + 1], c = t.hi[2 * k], d = t.hi[2 * k + 1];
    t.lo[k] = P3{min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)}; t.hi[k] = P3{max(c.x, d.x), max(c.y, d.y), max(c.z, d.z)}; }
}
static inline float bd(const Tree& t, uint32_t k, P3 q) {
  float ex = max(max(t.lo[k].x - q.x, q.x - t.hi[k].x), 0.f), ey = max(max(t.lo[k].y - q.y, q.y - t.hi[k].y), 0.f), ez = max(max(t.lo[k].z - q.z, q.z - t.hi[k].z), 0.f);
  return ex * ex + ey * ey + ez * ez;
}
struct Stats { double steps = 0, leaves = 0, pushes = 0, minleaves = 0, found = 0; };
// binary best-first-child DFS like the kernel (counts binary steps; 4-ary ~ half)
static float g_init = -1;
static void query(const Tree& t, P3 q, float mr2, Stats& s) {
  float best = g_init >= 0 ? g_init : mr2; if (!(bd(t, 1, q) <= best)) return;
  vector<pair<float, uint32_t>> st; uint32_t node = 1;
  for (;;) {
    while (node < t.P) { s.steps++; float d0 = bd(t, 2 * node, q), d1 = bd(t, 2 * node + 1, q); bool f0 = d0 <= d1; float dn = f0 ? d0 : d1, df = f0 ? d1 : d0;
      if (!(dn <= best)) { node = 0; break; } if (df <= best) { st.push_back({df, 2 * node + (f0 ? 1 : 0)}); s.pushes++; } node = 2 * node + (f0 ? 0 : 1); }
    if (node) { s.leaves++; uint32_t base = (node - t.P) * t.L; for (int j = 0; j < t.L; j++) { size_t i = base + j; if (i >= t.pts.size()) break; P3 p = t.pts[i]; float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z; float d = dx * dx + dy * dy + dz * dz; if (d < best) best = d; } }
    node = 0; while (!st.empty()) { auto e = st.back(); st.pop_back(); if (e.first <= best) { node = e.second; break; } }
    if (!node) break;
  }
  if (g_init >= 0) return;
  if (best < mr2) { s.found++; // lower bound: leaves with box dist <= best
    // count via traversal with fixed bound
    vector<uint32_t> s2{1}; while (!s2.empty()) { uint32_t k = s2.back(); s2.pop_back(); if (!(bd(t, k, q) <= best)) continue; if (k >= t.P) { s.minleaves++; continue; } s2.push_back(2 * k); s2.push_back(2 * k + 1); } }
}
static Tree make_tree(const vector<P3>& pts, const vector<uint32_t>& order, int L) {
  Tree t; t.L = L; t.pts.resize(pts.size()); t.id = order; for (size_t i = 0; i < pts.size(); i++) t.pts[i] = pts[order[i]];
  t.leaves = (pts.size() + L - 1) / L; t.P = 1; while (t.P < t.leaves) t.P <<= 1; build_boxes(t); return t;
}
// KD order: node at height h covers 2^h leaves; left child takes min(count, 2^(h-1)*L) points, split along widest axis
static void kd_order(const vector<P3>& pts, vector<uint32_t>& idx, size_t b, size_t e, uint32_t cap_leaves, int L, int mode) {
  if (cap_leaves <= 1 || e - b <= (size_t)L) return;
  size_t left = min<size_t>(e - b, (size_t)(cap_leaves / 2) * L);
  if (mode == 1) { // balanced median rounded to a multiple of L (needs pointer tree on GPU; here just to see)
  }
  if (left < e - b) {
    P3 lo{INFINITY, INFINITY, INFINITY}, hi{-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = b; i < e; i++) { P3 p = pts[idx[i]]; lo.x = min(lo.x, p.x); lo.y = min(lo.y, p.y); lo.z = min(lo.z, p.z); hi.x = max(hi.x, p.x); hi.y = max(hi.y, p.y); hi.z = max(hi.z, p.z); }
    float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z; int ax = ex >= ey && ex >= ez ? 0 : (ey >= ez ? 1 : 2);
    nth_element(idx.begin() + b, idx.begin() + b + left, idx.begin() + e, [&](uint32_t a, uint32_t c) { const float* pa = &pts[a].x; const float* pc = &pts[c].x; return pa[ax] < pc[ax]; });
    kd_order(pts, idx, b, b + left, cap_leaves / 2, L, mode); kd_order(pts, idx, b + left, e, cap_leaves / 2, L, mode);
  } else kd_order(pts, idx, b, e, cap_leaves / 2, L, mode);
}
int main(int argc, char** argv) {
  auto pts = readf("/tmp/nnm/tgt.bin"); auto qs = readf("/tmp/nnm/q.bin"); size_t n = pts.size();
  int nq = argc > 1 ? atoi(argv[1]) : 50000; float mr2 = 1.0f;
  P3 lo{INFINITY, INFINITY, INFINITY}, hi{-INFINITY, -INFINITY, -INFINITY};
  for (auto p : pts) { lo.x = min(lo.x, p.x); lo.y = min(lo.y, p.y); lo.z = min(lo.z, p.z); hi.x = max(hi.x, p.x); hi.y = max(hi.y, p.y); hi.z = max(hi.z, p.z); }
  float ext = max(hi.x - lo.x, max(hi.y - lo.y, hi.z - lo.z)); float scale = 65536.f / ext;
  auto run = [&](const char* name, const Tree& t) { Stats s; for (int i = 0; i < nq; i++) query(t, qs[(size_t)i * (qs.size() / nq)], mr2, s);
    if (getenv("PERFECT")) { Stats s2; for (int i = 0; i < nq; i++) { P3 q = qs[(size_t)i * (qs.size() / nq)]; g_init = -1; Stats tmp; // get true nn
        float best = mr2; for (auto& p : t.pts) { float dx=p.x-q.x, dy=p.y-q.y, dz=p.z-q.z; float d=dx*dx+dy*dy+dz*dz; if (d<best) best=d; } g_init = best * 1.0001f; query(t, q, mr2, s2); g_init = -1; }
      printf("   perfect bound: bin-steps/q %.1f leaf scans/q %.1f pushes/q %.1f\n", s2.steps / nq, s2.leaves / nq, s2.pushes / nq); }
    printf("%-28s L=%2d  bin-steps/q %6.1f  leaf scans/q %6.1f  pts tested/q %7.1f  pushes/q %5.1f  min leaves/q %5.1f  found %.3f\n", name, t.L, s.steps / nq, s.leaves / nq, s.leaves * t.L / nq, s.pushes / nq, s.minleaves / max(1.0, s.found), s.found / nq); fflush(stdout); };
  vector<uint64_t> key(n); vector<uint32_t> ord(n);
  for (int variant = 0; variant < 2; variant++) {
    for (size_t i = 0; i < n; i++) { uint32_t u[3]; const float* c = &pts[i].x; const float* l = &lo.x; for (int k = 0; k < 3; k++) u[k] = (uint32_t)min(max((c[k] - l[k]) * scale, 0.f), 65535.f);
      key[i] = variant == 0 ? hilbert(u[0], u[1], u[2], 16) : (spread16(u[0]) | (spread16(u[1]) << 1) | (spread16(u[2]) << 2)); }
    iota(ord.begin(), ord.end(), 0); stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    for (int L : {4, 8, 16, 32}) { Tree t = make_tree(pts, ord, L); run(variant == 0 ? "hilbert" : "morton", t); }
  }
  // per-axis scaled hilbert
  { float sc[3] = {65536.f / (hi.x - lo.x), 65536.f / (hi.y - lo.y), 65536.f / (hi.z - lo.z)};
    for (size_t i = 0; i < n; i++) { uint32_t u[3]; const float* c = &pts[i].x; const float* l = &lo.x; for (int k = 0; k < 3; k++) u[k] = (uint32_t)min(max((c[k] - l[k]) * sc[k], 0.f), 65535.f); key[i] = hilbert(u[0], u[1], u[2], 16); }
    iota(ord.begin(), ord.end(), 0); stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    for (int L : {8}) { Tree t = make_tree(pts, ord, L); run("hilbert per-axis scale", t); } }
  // 2D hilbert on x,y (z ignored), then z as tie-break
  { for (size_t i = 0; i < n; i++) { uint32_t u[3]; const float* c = &pts[i].x; const float* l = &lo.x; for (int k = 0; k < 3; k++) u[k] = (uint32_t)min(max((c[k] - l[k]) * scale, 0.f), 65535.f);
      // 2D hilbert via 3D routine with third coord 0 is not a 2D curve; do simple 2D hilbert
      uint32_t x = u[0], y = u[1], rx, ry; uint64_t d = 0; for (uint32_t s2 = 32768; s2 > 0; s2 >>= 1) { rx = (x & s2) > 0; ry = (y & s2) > 0; d += (uint64_t)s2 * s2 * ((3 * rx) ^ ry); if (ry == 0) { if (rx == 1) { x = 65535 - x; y = 65535 - y; } uint32_t tt = x; x = y; y = tt; } }
      key[i] = d; }
    iota(ord.begin(), ord.end(), 0); stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    for (int L : {8}) { Tree t = make_tree(pts, ord, L); run("hilbert 2D (x,y)", t); } }
  // hybrid: hilbert order, then local KD reorder inside blocks of B points
  for (int B : {512, 2048, 8192, 65536}) {
    for (size_t i = 0; i < n; i++) { uint32_t u[3]; const float* c = &pts[i].x; const float* l = &lo.x; for (int k = 0; k < 3; k++) u[k] = (uint32_t)min(max((c[k] - l[k]) * scale, 0.f), 65535.f); key[i] = hilbert(u[0], u[1], u[2], 16); }
    iota(ord.begin(), ord.end(), 0); stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    int L = 8; for (size_t b = 0; b < n; b += B) kd_order(pts, ord, b, min(n, b + (size_t)B), B / L, L, 0);
    Tree t = make_tree(pts, ord, L); char nm[64]; snprintf(nm, 64, "hilbert + local kd B=%d", B); run(nm, t); }
  for (int L : {4, 8, 16, 32}) { iota(ord.begin(), ord.end(), 0); uint32_t leaves = (n + L - 1) / L, P = 1; while (P < leaves) P <<= 1; kd_order(pts, ord, 0, n, P, L, 0); Tree t = make_tree(pts, ord, L); run("kd (widest axis, pow2 split)", t); }
  return 0;
}
