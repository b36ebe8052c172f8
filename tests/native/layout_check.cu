// Host-side check of rc::smem_layout (test infrastructure): arrays that are live together
// never overlap, everything stays inside the reported size, and spill mode keeps exactly
// misc / mbits / mpref in shared memory.  Prints "ok <n>" or the first violation.
#include <cstdio>
#include <vector>
#include "../../gerrychain_b200/csrc/recom_device.cuh"

struct Span { const char* name; size_t lo, hi; };

static bool disjoint(const std::vector<Span>& v, const char* what) {
  for (size_t i = 0; i < v.size(); ++i)
    for (size_t j = i + 1; j < v.size(); ++j)
      if (v[i].lo < v[j].hi && v[j].lo < v[i].hi && v[i].lo != v[i].hi && v[j].lo != v[j].hi) {
        printf("overlap in %s: %s [%zu,%zu) and %s [%zu,%zu)\n", what, v[i].name, v[i].lo, v[i].hi, v[j].name, v[j].lo, v[j].hi);
        return false;
      }
  return true;
}

static bool check(int N, int CN, int CE, int R) {
  using namespace rc;
  static unsigned char dummy[1], coldbuf[1];
  unsigned char* base = dummy;  // offsets only: nothing is dereferenced
  Smem s;
  size_t split = 0, cold = 0;
  const size_t total = smem_layout(&s, base, N, CN, CE, R, nullptr, coldbuf, &split, &cold);
  auto off = [&](const void* p) { return (size_t)((const unsigned char*)p - base); };
  const int NW = (N + 31) / 32, CW = (CN + 31) / 32, EW = (CE + 31) / 32;
#define SP(ptr, bytes) Span{#ptr, off(s.ptr), off(s.ptr) + (size_t)(bytes)}
  std::vector<Span> fixed = {SP(misc, 512), SP(mbits, 4 * NW), SP(mpref, 2 * NW), SP(tflag, 4 * EW), SP(oldb, 4 * CW),
                             SP(newb, 4 * CW)};
  std::vector<Span> tree = fixed;
  for (Span x : {SP(key, 4 * CE), SP(lab, 4 * CE), SP(lj, 2 * CE), SP(best, 4 * CN), SP(bpos, 4 * CN),
                 SP(hook, 2 * CN), SP(vlA, 2 * (CN / 2 + 2)), SP(vlB, 2 * (CN / 2 + 2))}) tree.push_back(x);
  // Euler phase: the sublist ranking (spX, spY) is dead when ent is written; ent aliases spY
  std::vector<Span> euler = fixed, euler2 = fixed;
  for (Span x : {SP(te, 2 * CN), SP(tuv, 4 * CN), SP(head, 4 * CN), SP(nxt, 4 * CN), SP(succ, 4 * CN), SP(P, 4 * CN),
                 SP(W, 8 * CN)}) { euler.push_back(x); euler2.push_back(x); }
  for (Span x : {SP(spX, 8 * s.spCap), SP(spY, 8 * s.spCap)}) euler.push_back(x);
  euler2.push_back(SP(ent, 2 * CN));
#undef SP
  if (!disjoint(tree, "tree phase") || !disjoint(euler, "euler phase") || !disjoint(euler2, "euler phase (ent)")) return false;
  for (const auto& v : {tree, euler, euler2})
    for (const Span& x : v)
      if (x.hi > total) { printf("%s ends at %zu > total %zu\n", x.name, x.hi, total); return false; }
  if ((const void*)s.cur != (const void*)s.succ) { printf("cur must alias succ\n"); return false; }
  if ((const void*)s.ent != (const void*)s.spY) { printf("ent must alias spY\n"); return false; }
  if (split != off(s.tflag)) { printf("split %zu != first array behind the core %zu\n", split, off(s.tflag)); return false; }
  // cold tier: pop, euv, ecls disjoint inside the cold block
  std::vector<Span> cl = {Span{"pop", (size_t)((unsigned char*)s.pop - coldbuf), (size_t)((unsigned char*)s.pop - coldbuf) + 4 * (size_t)CN},
                          Span{"euv", (size_t)((unsigned char*)s.euv - coldbuf), (size_t)((unsigned char*)s.euv - coldbuf) + 4 * (size_t)CE},
                          Span{"ecls", (size_t)((unsigned char*)s.ecls - coldbuf), (size_t)((unsigned char*)s.ecls - coldbuf) + (R > 0 ? (size_t)CE : 0)}};
  if (!disjoint(cl, "cold tier")) return false;
  for (const Span& x : cl) if (x.hi > cold) { printf("%s ends at %zu > cold %zu\n", x.name, x.hi, cold); return false; }
  // spill mode: the shared part is exactly [0, split); the rest is addressed relative to big
  static unsigned char bigbuf[1];
  Smem t;
  smem_layout(&t, base, N, CN, CE, R, bigbuf, coldbuf, nullptr, nullptr);
  if ((unsigned char*)t.misc != base || off(t.mpref) + 2 * (size_t)NW > split) { printf("spill: shared part wrong\n"); return false; }
  if ((unsigned char*)t.tflag != bigbuf) { printf("spill: tflag must start the scratch block\n"); return false; }
  if ((size_t)((unsigned char*)t.ent - bigbuf) + 2 * (size_t)CN > total - split) { printf("spill: ent beyond the scratch block\n"); return false; }
  return true;
}

// rc::wilson_layout: for every placement level the arrays of each tier are disjoint and inside
// the tier's block (shared memory / the CTA's global block / the cold block); wtmp aliases wsum.
static bool check_wilson(int N, int CN, int CE, int R, int W, int sh, int nbr_bytes = 4) {
  using namespace rc;
  const int NW = (N + 31) / 32, CW = (CN + 31) / 32;
  for (int level = 0; level < (nbr_bytes == 2 ? 1 : 3); ++level) {
    static unsigned char sm[1], gl[1], co[1];
    Smem s;
    size_t glob = 0, cold = 0;
    const int msh = N >= 65536 ? 2 : 0;
    const size_t smem = wilson_layout(&s, sm, N, CN, CE, R, W, sh, level, msh, gl, co, &glob, &cold, nbr_bytes);
    std::vector<Span> in_sm, in_gl, in_co;
    auto put = [&](const char* name, const void* p, size_t bytes, int tier) {
      const unsigned char* q = (const unsigned char*)p;
      // tier 0: core (always shared); 1: S; 2: A; 3: cold
      const bool shared = tier == 0 || (tier == 1 && level < 2) || (tier == 2 && level == 0);
      if (tier == 3) in_co.push_back(Span{name, (size_t)(q - co), (size_t)(q - co) + bytes});
      else if (shared) in_sm.push_back(Span{name, (size_t)(q - sm), (size_t)(q - sm) + bytes});
      else in_gl.push_back(Span{name, (size_t)(q - gl), (size_t)(q - gl) + bytes});
    };
    put("misc", s.misc, 512, 0); put("wmeta", s.wmeta, 512, 0); put("wring", s.wring, (size_t)W * 512, 0);
    put("mbits", s.mbits, 4 * (size_t)NW, 0); put("mpref", s.mpref, 2 * (size_t)((NW >> msh) + 1), 0);
    put("oldb", s.oldb, 4 * (size_t)CW, 1); put("newb", s.newb, 4 * (size_t)CW, 1); put("whas", s.whas, 4 * (size_t)CW, 1);
    put("wsum", s.wsum, 4 * (size_t)CN, 1); put("wst", s.wst, 2 * (size_t)W * CN, 1);
    put("nbr", s.nbr, (size_t)nbr_bytes * ((size_t)(CN + CN / 8) << sh), 2); put("wdeg", s.wdeg, 2 * (size_t)CN, 2);
    put("pop", s.pop, 4 * (size_t)CN, 3); put("euv", s.euv, 4 * (size_t)CE, 3);
    put("ecls", s.ecls, R > 0 ? (size_t)CE : 0, 3);
    if (!disjoint(in_sm, "wilson shared") || !disjoint(in_gl, "wilson global") || !disjoint(in_co, "wilson cold")) return false;
    for (const Span& x : in_sm) if (x.hi > smem) { printf("level %d: %s ends at %zu > smem %zu\n", level, x.name, x.hi, smem); return false; }
    for (const Span& x : in_gl) if (x.hi > glob) { printf("level %d: %s ends at %zu > global block %zu\n", level, x.name, x.hi, glob); return false; }
    for (const Span& x : in_co) if (x.hi > cold) { printf("level %d: %s ends at %zu > cold %zu\n", level, x.name, x.hi, cold); return false; }
    if ((const void*)s.wtmp != (const void*)s.wsum) { printf("wtmp must alias wsum\n"); return false; }
    if (level == 0 && glob != 0) { printf("level 0 has a global block\n"); return false; }
  }
  return true;
}

int main() {
  int n = 0;
  const int cases[][4] = {{400, 232, 480, 0}, {10000, 2432, 4992, 0}, {9000, 1232, 3840, 1}, {9000, 1232, 3840, 3},
                          {200000, 9632, 29536, 0}, {64, 64, 112, 2}, {9, 16, 12, 0}, {65000, 32760, 65535, 0}};
  for (const auto& c : cases) {
    if (!check(c[0], c[1], c[2], c[3])) { printf("case N=%d CN=%d CE=%d R=%d\n", c[0], c[1], c[2], c[3]); return 1; }
    ++n;
    for (int W : {1, 4, 8})
      for (int sh : {2, 3}) {
        if (!check_wilson(c[0], c[1], c[2], c[3], W, sh)) { printf("wilson case N=%d CN=%d CE=%d R=%d W=%d sh=%d\n", c[0], c[1], c[2], c[3], W, sh); return 1; }
        ++n;
        if (c[1] + c[1] / 8 <= 2048) {  // 16-bit adjacency entries (level 0)
          if (!check_wilson(c[0], c[1], c[2], c[3], W, sh, 2)) { printf("wilson16 case N=%d CN=%d W=%d sh=%d\n", c[0], c[1], W, sh); return 1; }
          ++n;
        }
      }
  }
  printf("ok %d\n", n);
  return 0;
}
