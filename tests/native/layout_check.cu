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

static bool check(int N, int CN, int CE, int R, int wilson) {
  using namespace rc;
  static unsigned char dummy[1];
  unsigned char* base = dummy;  // offsets only: nothing is dereferenced
  Smem s;
  size_t split = 0;
  const size_t total = smem_layout(&s, base, N, CN, CE, R, wilson, nullptr, &split);
  auto off = [&](const void* p) { return (size_t)((const unsigned char*)p - base); };
  const int NW = (N + 31) / 32, CW = (CN + 31) / 32, EW = (CE + 31) / 32;
#define SP(ptr, bytes) Span{#ptr, off(s.ptr), off(s.ptr) + (size_t)(bytes)}
  std::vector<Span> fixed = {SP(misc, 512), SP(mbits, 4 * NW), SP(mpref, 2 * NW), SP(pop, 4 * CN), SP(eu, 2 * CE),
                             SP(ev, 2 * CE), SP(tflag, 4 * EW), SP(oldb, 4 * CW), SP(newb, 4 * CW),
                             SP(ecls, R > 0 ? CE : 0)};
  if (wilson) { fixed.push_back(SP(loff, 2 * (CN + 1))); fixed.push_back(SP(ladj, 4 * CE)); fixed.push_back(SP(ladje, 4 * CE)); }
  std::vector<Span> tree = fixed;
  if (wilson) {
    for (Span x : {SP(wtmp, 4 * CN), SP(wcnt, 2 * CN), SP(wgA, 2 * CN), SP(wgB, 2 * CN), SP(wmark, 4 * CW),
                   SP(wnx, 2 * CN), SP(wst, 8 * CN), SP(wtab, 16 * CN)}) tree.push_back(x);
  } else {
    for (Span x : {SP(key, 4 * CE), SP(lab, 4 * CE), SP(lj, 2 * CE), SP(best, 4 * CN), SP(bpos, 4 * CN),
                   SP(hook, 2 * CN), SP(vlA, 2 * (CN / 2 + 2)), SP(vlB, 2 * (CN / 2 + 2))}) tree.push_back(x);
  }
  std::vector<Span> euler = fixed;
  for (Span x : {SP(te, 2 * CN), SP(head, 4 * CN), SP(nxt, 4 * CN), SP(succ, 4 * CN), SP(P, 4 * CN), SP(W, 8 * CN),
                 SP(spX, 8 * s.spCap), SP(spY, 8 * s.spCap), SP(ent, 2 * CN)}) euler.push_back(x);
#undef SP
  if (!disjoint(tree, "tree phase") || !disjoint(euler, "euler phase")) return false;
  for (const auto& v : {tree, euler})
    for (const Span& x : v)
      if (x.hi > total) { printf("%s ends at %zu > total %zu\n", x.name, x.hi, total); return false; }
  if ((const void*)s.cur != (const void*)s.succ) { printf("cur must alias succ\n"); return false; }
  // the gather's local->global table borrows best (MST) / wtmp (Wilson): 4 * CN bytes each
  if (split != off(s.pop)) { printf("split %zu != first spilled array %zu\n", split, off(s.pop)); return false; }
  // spill mode: the shared part is exactly [0, split); the rest is addressed relative to big
  static unsigned char bigbuf[1];
  Smem t;
  smem_layout(&t, base, N, CN, CE, R, wilson, bigbuf, nullptr);
  if ((unsigned char*)t.misc != base || off(t.mpref) + 2 * (size_t)NW > split) { printf("spill: shared part wrong\n"); return false; }
  if ((unsigned char*)t.pop != bigbuf) { printf("spill: pop must start the scratch block\n"); return false; }
  if ((size_t)((unsigned char*)t.ent - bigbuf) + 2 * (size_t)CN > total - split) { printf("spill: ent beyond the scratch block\n"); return false; }
  return true;
}

int main() {
  int n = 0;
  const int cases[][4] = {{400, 232, 480, 0}, {10000, 2432, 4992, 0}, {9000, 1232, 3840, 1}, {9000, 1232, 3840, 3},
                          {200000, 9632, 29536, 0}, {64, 64, 112, 2}, {9, 16, 12, 0}, {65000, 32760, 65535, 0}};
  for (const auto& c : cases)
    for (int wilson = 0; wilson < 2; ++wilson) {
      if (!check(c[0], c[1], c[2], c[3], wilson)) { printf("case N=%d CN=%d CE=%d R=%d wilson=%d\n", c[0], c[1], c[2], c[3], wilson); return 1; }
      ++n;
    }
  printf("ok %d\n", n);
  return 0;
}
