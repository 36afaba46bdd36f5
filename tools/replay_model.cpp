// Tuning aid (CPU only): exact model of the strict-ICP replay walk (icp.cu, "fast exact replay") on real term streams.
// Counts, per accumulator stream, how many chunks leave the one-add path and why, and what finer summaries would buy.
// Input: tools/replay_model_dump.py -> /tmp/rpm/iter{0,N}.bin ([9][n] float32).
// Build: g++ -O2 -ffp-contract=off -o /tmp/rpm/model tools/replay_model.cpp ; run: /tmp/rpm/model /tmp/rpm/iterN.bin [chunk] [sub]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace std;
typedef long long ll;
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static int fexp(float f) { return (int)((bits(f) >> 23) & 0xff) - 127; }
struct PMap { ll off[2], mn[2], mx[2]; int np; };
static PMap ident() { PMap m; m.off[0] = m.off[1] = 0; m.mn[0] = m.mn[1] = (1ll << 60); m.mx[0] = m.mx[1] = -(1ll << 60); m.np = 2; return m; }
static PMap compose(const PMap& a, const PMap& b) {
  PMap r; r.np = 0;
  for (int p = 0; p < 2; p++) { int mid = (a.np >> p) & 1; r.off[p] = a.off[p] + b.off[mid]; r.mn[p] = min(a.mn[p], a.off[p] + b.mn[mid]); r.mx[p] = max(a.mx[p], a.off[p] + b.mx[mid]); r.np |= ((b.np >> mid) & 1) << p; }
  return r;
}
struct Summary { bool usable; int e; PMap m; };
// summary of x[b, b+len) in units of the ulp of binade e (the kernel's icp_replay_summaries_kernel)
static Summary summarize(const float* x, ll b, ll len, ll n, int e, bool regular) {
  Summary s; s.e = e; s.m = ident();
  const double scale = regular ? ldexp(1.0, 23 - e) : 0.0;
  for (ll i = b; i < b + len; i++) {
    const float xv = i < n ? x[i] : 0.f; const double y = (double)xv * scale; double q = floor(y); const bool sane = fabs(y) < 1.0e12;
    if (!sane) { regular = false; q = 0; }
    const double frac = sane ? y - q : 0.0; const ll qi = (ll)q; PMap el;
    if (frac == 0.5) { el.off[0] = qi + (qi & 1); el.off[1] = qi + ((qi + 1) & 1); el.np = 0; }
    else { const ll r = qi + (frac > 0.5 ? 1 : 0); el.off[0] = el.off[1] = r; el.np = (r & 1) ? 1 : 2; }
    el.mn[0] = el.mx[0] = el.off[0]; el.mn[1] = el.mx[1] = el.off[1]; s.m = compose(s.m, el);
  }
  s.usable = regular && e <= 100; return s;
}
// the walk's interval test, in integers: is the accumulator (in binade s.e) kept strictly inside by every prefix?
static bool fast_ok(const Summary& s, float acc, float* out) {
  if (!s.usable || acc == 0.f || fexp(acc) != s.e) return false;
  const ll kLo = 1ll << 23, kHi = 1ll << 24; const uint32_t u = bits(acc); const int p = u & 1; ll M = (ll)((u & 0x7fffff) | 0x800000); if (u >> 31) M = -M;
  const ll mn = s.m.mn[p], mx = s.m.mx[p], off = s.m.off[p]; if (!(off > -kHi && off < kHi)) return false;
  bool ok;
  if (M > 0) ok = M >= max(kLo + 1 - mn, kLo) && M <= min(kHi - 1 - mx, kHi - 1); else ok = M >= max(-kHi + 1 - mn, -(kHi - 1)) && M <= min(-kLo - 1 - mx, -kLo);
  if (ok) *out = (float)(M + off) * ldexpf(1.f, s.e - 23);
  return ok;
}
static bool guess_regular(float g) { const int e = fexp(g); return g != 0.f && e > -100 && e < 128; }
int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "/tmp/rpm/iterN.bin"; const int CH = argc > 2 ? atoi(argv[2]) : 256; const int SUB = argc > 3 ? atoi(argv[3]) : 32;
  ll n; vector<float> all;
  const char* names_icp[9] = {"Value", "SumW", "G0", "G1", "G2", "G3", "G4", "G5", "R"};
  const char* names_syn[9] = {"unif", "ties", "heavy", "altern", "const", "ints", "zeroX", "grow", "sparse"};
  const char** names = names_icp;
  if (!strcmp(path, "selftest")) {  // adversarial synthetic streams: ties, sign changes, mixed magnitudes, zero crossings
    names = names_syn; n = 70001; all.resize((size_t)n * 9); uint64_t st = argc > 4 ? strtoull(argv[4], 0, 10) : 12345;
    auto rnd = [&]() { st += 0x9e3779b97f4a7c15ull; uint64_t z = st; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); };
    auto uni = [&]() { return (double)(rnd() >> 11) / 9007199254740992.0; };
    auto nrm = [&]() { return sqrt(-2 * log(uni() + 1e-300)) * cos(6.283185307179586 * uni()); };
    for (ll i = 0; i < n; i++) {
      all[0 * n + i] = (float)(2 * uni() - 1);
      all[1 * n + i] = (float)((double)((ll)(rnd() % 8193) - 4096) / 1024.0);           // multiples of 2^-10: exact ties
      all[2 * n + i] = (float)(nrm() * exp(3 * nrm()));
      all[3 * n + i] = (float)((i & 1 ? -1.0 : 1.0) + 1e-4 * nrm());
      all[4 * n + i] = 0.1f;
      all[5 * n + i] = (float)(1e-3 * (double)((ll)(rnd() % 2001) - 1000));
      all[6 * n + i] = (float)(sin(i * 0.001) * (1 + 0.01 * nrm()));                     // sum oscillates through zero
      all[7 * n + i] = (float)(ldexp(1.0, (int)(i / 3000)) * (uni() - 0.3));             // grows over 23 binades
      all[8 * n + i] = (rnd() % 50 == 0) ? (float)(1000 * nrm()) : 0.f;
    }
  } else {
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); return 1; } fseek(f, 0, SEEK_END); const long sz = ftell(f); fseek(f, 0, SEEK_SET);
    n = sz / 4 / 9; all.resize((size_t)n * 9); if (fread(all.data(), 4, all.size(), f) != all.size()) return 1; fclose(f);
  }
  printf("%s: n = %lld, chunk %d, sub-chunk %d\n", path, n, CH, SUB);
  printf("%-6s %6s %6s | slow: %7s %8s %8s %8s | %10s %10s %10s | %s\n", "stream", "chunks", "slow", "unusabl", "binade!=", "crossing", "margin", "B:slow-sub", "C:1-pass", "D:slow-sub", "check");
  ll worst_slow = 0, worst_b = 0, worst_d = 0;
  for (int k = 0; k < 9; k++) {
    const float* x = all.data() + (size_t)k * n; const ll nch = (n + CH - 1) / CH;
    vector<double> csum(nch), psub((size_t)(n + SUB - 1) / SUB + 1, 0.0);
    for (ll c = 0; c < nch; c++) { double s = 0; for (ll i = c * CH; i < min(n, (c + 1) * CH); i++) s += x[i]; csum[c] = s; }
    { double s = 0; for (ll i = 0; i < n; i++) { if (i % SUB == 0) psub[i / SUB] = s; s += x[i]; } }
    float acc = 0.f, seq = 0.f; double g = 0; ll slow = 0, c_unus = 0, c_bin = 0, c_cross = 0, c_margin = 0, b_slow_sub = 0, c_onepass = 0, d_slow_sub = 0, e_slow_sub = 0, pred_miss = 0, pred_extra = 0;
    for (ll c = 0; c < nch; c++) {
      const float guess = (float)g; g += csum[c];
      const Summary s = summarize(x, c * CH, CH, n, fexp(guess), guess_regular(guess));
      float nxt;
      // prediction available when the summaries are built: does the GUESS accumulator pass the interval test?
      float dummy; const bool predicted_fast = fast_ok(s, guess, &dummy);
      if (fast_ok(s, acc, &nxt)) { acc = nxt; if (!predicted_fast) pred_extra++; }
      else {
        slow++;
        if (predicted_fast) pred_miss++;
        // why
        bool crossing = false; { float a = acc; const int e0 = fexp(a); for (ll i = c * CH; i < min(n, (c + 1) * CH); i++) { a += x[i]; if (a == 0.f || fexp(a) != e0) crossing = true; } }
        if (!s.usable) c_unus++; else if (acc == 0.f || fexp(acc) != s.e) c_bin++; else if (crossing) c_cross++; else c_margin++;
        // C: one more warp pass with the TRUE binade of the accumulator settles the chunk in one add
        { const Summary t = summarize(x, c * CH, CH, n, fexp(acc), guess_regular(acc)); float o; if (fast_ok(t, acc, &o)) c_onepass++; }
        // B: sub-chunk summaries with their own f64-prefix guesses; D: sub-chunk summaries in the true binade
        float a = acc, ad = acc, ae = acc;
        for (ll b = c * CH; b < min(n, (c + 1) * (ll)CH); b += SUB) {
          const float gs = (float)psub[b / SUB]; const Summary t = summarize(x, b, SUB, n, fexp(gs), guess_regular(gs)); float o;
          if (fast_ok(t, a, &o)) a = o; else { b_slow_sub++; for (ll i = b; i < min(n, b + SUB); i++) a += x[i]; }
          const Summary td = summarize(x, b, SUB, n, fexp(ad), guess_regular(ad));
          if (fast_ok(td, ad, &o)) ad = o; else { d_slow_sub++; for (ll i = b; i < min(n, b + SUB); i++) ad += x[i]; }
          // E: sub-chunk maps exist only for the chunk's guessed binade and its two neighbours
          const int ee = fexp(ae); const bool have = ae != 0.f && abs(ee - s.e) <= 1;
          const Summary te = summarize(x, b, SUB, n, ee, have && guess_regular(ae));
          if (have && fast_ok(te, ae, &o)) ae = o; else { e_slow_sub++; for (ll i = b; i < min(n, b + SUB); i++) ae += x[i]; }
        }
        for (ll i = c * CH; i < min(n, (c + 1) * (ll)CH); i++) acc += x[i];
        if (bits(a) != bits(acc) || bits(ad) != bits(acc) || bits(ae) != bits(acc)) { printf("MODEL ERROR in sub-chunk path, stream %d chunk %lld\n", k, c); return 2; }
      }
    }
    for (ll i = 0; i < n; i++) seq += x[i];
    printf("%-6s %6lld %6lld | %13lld %8lld %8lld %8lld | %10lld %10lld %10lld | %s  E:slow-sub %lld  predicted-fast-but-slow %lld, predicted-slow-but-fast %lld\n", names[k], nch, slow, c_unus, c_bin, c_cross, c_margin, b_slow_sub, c_onepass, d_slow_sub, bits(seq) == bits(acc) ? "exact" : "MISMATCH", e_slow_sub, pred_miss, pred_extra);
    worst_slow = max(worst_slow, slow); worst_b = max(worst_b, b_slow_sub); worst_d = max(worst_d, d_slow_sub);
  }
  printf("slowest stream: %lld replayed chunks (%lld element adds) now; two-level: %lld sub-chunks (%lld adds) with prefix guesses, %lld (%lld adds) in the true binade\n",
         worst_slow, worst_slow * CH, worst_b, worst_b * SUB, worst_d, worst_d * SUB);
  return 0;
}
