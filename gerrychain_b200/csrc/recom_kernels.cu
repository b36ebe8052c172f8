// recom_kernels.cu - the ReCom chain-step kernel and the small state kernels (sm_100a).
//
// Reference behaviour being reproduced (paths relative to /root/reference/gerrychain/):
//   pick pair / subgraph   proposals/tree_proposals.py:103-120
//   random weights + MST   tree.py:60-94 (networkx Kruskal == unique MST == Boruvka here)
//   rooting, subtree pops  tree.py:52-57, :289-322  (Euler tour + list ranking + scan here)
//   balanced cuts          tree.py:411-423 (float64 two-sided test)
//   cut choice             tree.py:501-578
//   retry loop             tree.py:679-718
//   relabel + updaters     tree.py:947-964, partition/assignment.py:76-86,
//                          updaters/tally.py:162-165, updaters/cut_edges.py:118-120
//   validity / accept      constraints/bounds.py:25-28, accept.py:15, chain.py:185-195
#include <type_traits>

#include "recom_device.cuh"

namespace rc {

struct Step {
  // block-uniform registers describing the proposal being built
  int a, b;        // merged district indices, a < b
  int n, m;        // nodes / undirected edges of the merged pair
  int tot;         // total population of the pair
};

struct Chain {
  const Params* p;
  Smem s;
  int chain;            // local chain index
  uint32_t gchain;      // global chain id (RNG subsequence)
  rc_label_t* assign;
  uint32_t* mask;
  int64_t* tally;
  int32_t* nodes;
  int32_t* eids;
  int par;              // block_excl_scan bank parity (toggled in lock step by all threads)
};

// ---- membership of a global node in the merged pair, and its local id ---------------------
__device__ __forceinline__ bool is_member(const Smem& s, int g) { return (s.mbits[g >> 5] >> (g & 31)) & 1u; }
__device__ __forceinline__ int local_id(const Smem& s, int g) {
  const int w = g >> 5;
  int id = (int)s.mpref[w >> s.msh] + __popc(s.mbits[w] & ((1u << (g & 31)) - 1u));
  for (int k = (w >> s.msh) << s.msh; k < w; ++k) id += __popc(s.mbits[k]);  // msh == 0: none
  return id;
}

// 4 assignment bytes -> 4 bits (byte k == a or == b)
__device__ __forceinline__ uint32_t match4(uint32_t x, uint32_t a4, uint32_t b4) {
  const uint32_t r = __vcmpeq4(x, a4) | __vcmpeq4(x, b4);
  return ((((r >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu;
}

__device__ __forceinline__ uint32_t match16(const uint4 q, uint32_t a4, uint32_t b4) {
  return match4(q.x, a4, b4) | match4(q.y, a4, b4) << 4 | match4(q.z, a4, b4) << 8 | match4(q.w, a4, b4) << 12;
}

// the 16-bit-label build: 2 labels per word -> 2 bits, 8 labels per 128-bit load -> 8 bits
__device__ __forceinline__ uint32_t match2(uint32_t x, uint32_t a2, uint32_t b2) {
  const uint32_t r = __vcmpeq2(x, a2) | __vcmpeq2(x, b2);
  return (r & 1u) | ((r >> 15) & 2u);
}

__device__ __forceinline__ uint32_t match8(const uint4 q, uint32_t a2, uint32_t b2) {
  return match2(q.x, a2, b2) | match2(q.y, a2, b2) << 2 | match2(q.z, a2, b2) << 4 | match2(q.w, a2, b2) << 6;
}

// membership bits of nodes 32 w .. 32 w + 31 of an assignment row (label == a or == b); bits past
// N are garbage (the caller masks them).  128-bit loads; rows are padded to 16 bytes.
__device__ __forceinline__ uint32_t row_match32(const rc_label_t* row, int w, int N, uint32_t apat, uint32_t bpat) {
#if RC_LABEL_BITS == 8
  const uint4* src = reinterpret_cast<const uint4*>(row + (size_t)w * 32);
  uint32_t bits = match16(__ldcg(src), apat, bpat);
  if (w * 32 + 16 < N) bits |= match16(__ldcg(src + 1), apat, bpat) << 16;
  return bits;
#else
  const uint4* src = reinterpret_cast<const uint4*>(row + (size_t)w * 32);
  uint4 q[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) q[j] = (w * 32 + 8 * j < N) ? __ldcg(src + j) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) bits |= match8(q[j], apat, bpat) << (8 * j);
  return bits;
#endif
}

// ---- weight key of local edge j for tree `tree_no` (tree.py:78-89) ------------------------
__device__ __forceinline__ uint32_t finish_key(const Chain& c, int j, uint32_t raw) {
  const Params& p = *c.p;
  if (p.g.smult == 0) return raw;
  uint64_t add = 0;
  const uint32_t cls = c.s.ecls[j];
  for (int r = 0; r < p.g.R; ++r)
    if ((cls >> r) & 1u) add += p.g.sfix[r];
  return rc_key_finish(raw, add, p.g.smult);
}

__device__ __forceinline__ uint32_t edge_key(const Chain& c, int j, uint32_t tree_no, uint32_t prop_no,
                                             const uint32_t* rrank) {
  const Params& p = *c.p;
  if (rrank) return rrank[c.eids[j]];
  return finish_key(c, j, rc_edge_raw(p.cfg.seed, c.gchain, (uint32_t)j, tree_no, prop_no));
}

// ---- P0: pick the cut edge (uniform over cut EDGES, tree_proposals.py:106) --------------
// returns edge id (block uniform) or -1
__device__ int pick_cut_edge(Chain& c, int cut_count, uint32_t prop_no, uint32_t retry) {
  const Params& p = *c.p;
  const int tid = threadIdx.x;
  RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_PAIR, retry, 0, prop_no, 0);
  const int kth = (int)rc_below(r.v[0], r.v[1], (uint32_t)cut_count);
  const int words = (p.g.E + 31) >> 5;
  const int per = odd_chunk(words);
  const int lo = min(tid * per, words), hi = min(lo + per, words);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += __popc(__ldcg(c.mask + i));
  if (tid == 0) c.s.misc[M_PICK] = -1;
  int tot;
  const int excl = block_excl_scan(cnt, &tot, c.s.misc, c.par);
  if (kth >= excl && kth < excl + cnt) {
    int k = kth - excl;
    for (int i = lo; i < hi; ++i) {
      const uint32_t x = __ldcg(c.mask + i);
      const int pc = __popc(x);
      if (k < pc) {
        c.s.misc[M_PICK] = i * 32 + (int)__fns(x, 0, k + 1);
        break;
      }
      k -= pc;
    }
  }
  __syncthreads();
  const int e = c.s.misc[M_PICK];
  __syncthreads();
  return e;
}

// Wilson only: adjacency of the merged pair for the walks.  Rows of D = 1 << nbr_shift slots;
// every entry is a neighbour (bits 0-15) TOGETHER WITH THAT NEIGHBOUR'S DEGREE (bits 16-31),
// rows in ascending neighbour id - the order in which a draw indexes the neighbours.  A walk
// step is then ONE dependent load: the entry it fetches names the next node and says how
// many neighbours the next draw chooses from.  Row u < n belongs to node u; the (few) nodes
// of degree > D keep D - 1 entries there and a link in slot D - 1 to a spare row (index >= n)
// with the next ones - linked again if those are more than D.
// entry number idx of node u (degree du): slot idx of a short row, else down the links
// (E: uint32_t entries, neighbour | degree << 16, or - small pairs - uint16_t, neighbour | degree << 11)
template <typename E>
__device__ __forceinline__ E* nbr_slot(const Smem& s, int u, int du, int idx, int sh) {
  const int D = 1 << sh;
  E* const base = reinterpret_cast<E*>(s.nbr);
  E* row = base + ((size_t)u << sh);
  while (du > D) { if (idx < D - 1) return row + idx; row = base + ((size_t)row[D - 1] << sh); idx -= D - 1; du -= D - 1; }
  return row + idx;
}
// neighbour number q of node u (degree d)
__device__ __forceinline__ int nbr_id(const Smem& s, const Params& p, int u, int d, int q) {
  return p.nbr16 ? (int)(*nbr_slot<uint16_t>(s, u, d, q, p.nbr_shift) & 0x7ffu) : (int)(*nbr_slot<uint32_t>(s, u, d, q, p.nbr_shift) & 0xffffu);
}

template <typename E>
__device__ int build_local_adjacency_t(Chain& c, const Step& st) {
  constexpr int kDegShift = sizeof(E) == 2 ? 11 : 16;
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, m = st.m, sh = p.nbr_shift, D = 1 << sh;
  const int rows_cap = p.cap_nodes + p.cap_nodes / 8;
  for (int i = tid; i < n; i += T) s.wtmp[i] = 0;
  if (tid == 0) s.misc[M_CNT] = n;  // next spare row
  __syncthreads();
  for (int j = tid; j < m; j += T) { const uint32_t uv = s.euv[j]; atomicAdd(&s.wtmp[uv & 0xffffu], 1u); atomicAdd(&s.wtmp[uv >> 16], 1u); }
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const int d = (int)s.wtmp[i];
    s.wdeg[i] = (uint16_t)min(d, 0xffff);
    if (d > D) {  // chain of spare rows: each full row but the last keeps D - 1 entries + the link
      // spare rows needed: the d - (D - 1) entries left go into rows of D - 1 (+ link), the last row takes up to D
      int need = 0;
      for (int left = d - (D - 1); left > 0; left -= (left > D ? D - 1 : left)) ++need;
      int r = atomicAdd(&s.misc[M_CNT], need);
      if (r + need <= rows_cap) {
        E* row = reinterpret_cast<E*>(s.nbr) + ((size_t)i << sh);
        for (int left = d - (D - 1); left > 0; left -= (left > D ? D - 1 : left)) {
          row[D - 1] = (E)r;
          row = reinterpret_cast<E*>(s.nbr) + ((size_t)r << sh);
          ++r;
        }
      }
    }
    s.wtmp[i] = 0;  // fill cursor
  }
  __syncthreads();
  if (s.misc[M_CNT] > rows_cap) return RC_ERR_CAPACITY;  // too many high-degree nodes for the spare rows
  for (int j = tid; j < m; j += T) {
    const uint32_t uv = s.euv[j];
    const int u = (int)(uv & 0xffffu), v = (int)(uv >> 16);
    *nbr_slot<E>(s, u, s.wdeg[u], (int)atomicAdd(&s.wtmp[u], 1u), sh) = (E)v;
    *nbr_slot<E>(s, v, s.wdeg[v], (int)atomicAdd(&s.wtmp[v], 1u), sh) = (E)u;
  }
  __syncthreads();
  for (int i = tid; i < n; i += T) {  // insertion sort of a short row, then the neighbours' degrees
    const int d = s.wdeg[i];
    for (int a = 1; a < d; ++a) {
      const E kv = *nbr_slot<E>(s, i, d, a, sh);
      int b = a - 1;
      while (b >= 0 && *nbr_slot<E>(s, i, d, b, sh) > kv) { *nbr_slot<E>(s, i, d, b + 1, sh) = *nbr_slot<E>(s, i, d, b, sh); --b; }
      *nbr_slot<E>(s, i, d, b + 1, sh) = kv;
    }
    for (int a = 0; a < d; ++a) { E* q = nbr_slot<E>(s, i, d, a, sh); const uint32_t v = *q; *q = (E)(v | ((uint32_t)s.wdeg[v] << kDegShift)); }
  }
  __syncthreads();
  return RC_OK;
}
__device__ int build_local_adjacency(Chain& c, const Step& st) {
  return c.p->nbr16 ? build_local_adjacency_t<uint16_t>(c, st) : build_local_adjacency_t<uint32_t>(c, st);
}

// ---- P1: gather the merged pair into shared memory ---------------------------------------
// returns RC_OK / RC_ERR_CAPACITY (block uniform)
template <bool kWilson>
__device__ int gather_pair(Chain& c, Step& st, int* d_sub_out) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int N = p.g.N, NW = (N + 31) >> 5;
#if RC_LABEL_BITS == 8
  const uint32_t a4 = 0x01010101u * (uint32_t)st.a, b4 = 0x01010101u * (uint32_t)st.b;
#else
  const uint32_t a4 = 0x00010001u * (uint32_t)st.a, b4 = 0x00010001u * (uint32_t)st.b;
#endif
  if (tid == 0) { s.misc[M_TOT] = 0; s.misc[M_CNT] = 0; }
  // local id -> global node table, built with the bitmap (a spanning-tree array that is
  // initialised after the gather serves as its home); in spill mode that array lives in
  // global memory and pass 2 searches the shared prefix array instead
  uint32_t* const l2g = kWilson ? s.wtmp : s.best;
  const bool use_tab = __isShared(l2g);
  // pass 1: membership bitmap + exclusive popcount prefix, T words (32 nodes each) per sweep
  int n = 0;
  for (int w0 = 0; w0 < NW; w0 += T) {
    const int w = w0 + tid;
    uint32_t bits = 0;
    if (w < NW) {
      // (the row is the one contiguous slice of the step; a single 1-D bulk copy of it into shared
      // memory - cp.async.bulk + mbarrier - was measured against these loads, scripts/dev/tma_bench.cu:
      // 1142 vs 740 cycles for a 10 KB row alone on the SM, 1576 vs 1697 with 3 CTAs per SM, i.e.
      // +-0.1 % of a step - not worth the staging buffer)
      bits = row_match32(c.assign, w, N, a4, b4);
      const int rem = N - w * 32;
      if (rem < 32) bits &= (1u << rem) - 1u;
      s.mbits[w] = bits;
    }
    int tot;
    const int ex = block_excl_scan(__popc(bits), &tot, s.misc, c.par);
    if (w < NW && (w & ((1 << s.msh) - 1)) == 0) s.mpref[w >> s.msh] = (lid_t)min(n + ex, 0xffff);
    if (use_tab) {
      int o = n + ex;
      for (uint32_t x = bits; x; x &= x - 1, ++o)
        if (o < p.cap_nodes) l2g[o] = (uint32_t)(w * 32 + (__ffs(x) - 1));
    }
    n += tot;
  }
  st.n = n;
  __syncthreads();
  if (n > p.cap_nodes || n > kMaxLocal) return RC_ERR_CAPACITY;
  // pass 2, T local nodes per sweep: local id -> global node, population, old side, and the
  // node's forward edges inside the pair.  Thread order == node order and slots ascend, so
  // edges come out in ascending (u, v) == ascending edge id.  The adjacency row is read
  // once for the count; the qualifying slots are remembered as a bitmask.  The first kRow
  // neighbours of a row (all of them on a grid, most on a triangulation) are requested
  // together and stay in registers for the emit: one L2 round trip instead of one per slot.
  constexpr int kRow = 4;  // measured: 4 beats 6 and 8 (registers held across the block scan)
  const int R = p.g.R;
  int popsum = 0, slots = 0, m = 0;
  for (int i0 = 0; i0 < n; i0 += T) {
    const int i = i0 + tid;
    int g = 0, r0 = 0, r1 = 0, extra = 0;
    int hs[kRow];
    uint32_t qmask = 0;
    bool isb = false;
    if (i < n) {
      if (use_tab) {
        g = (int)l2g[i];
      } else {
        int lo = 0, hi = (NW + (1 << s.msh) - 1) >> s.msh;  // first prefix entry that exceeds i
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((int)s.mpref[mid] > i) hi = mid; else lo = mid + 1;
        }
        int w = (lo - 1) << s.msh, rem = i - (int)s.mpref[lo - 1];
        for (int pc; rem >= (pc = __popc(s.mbits[w])); ++w) rem -= pc;
        g = w * 32 + (int)__fns(s.mbits[w], 0, rem + 1);
      }
      r0 = __ldg(p.g.row_ptr + g);
      r1 = __ldg(p.g.row_ptr + g + 1);
      const int pg = __ldg(p.g.pop + g);
      isb = (int)__ldcg(c.assign + g) == st.b;
#pragma unroll
      for (int k = 0; k < kRow; ++k) hs[k] = r0 + k < r1 ? __ldg(p.g.col_idx + r0 + k) : -1;
      c.nodes[i] = g;
      s.pop[i] = pg;
      popsum += pg;
      slots += r1 - r0;
#pragma unroll
      for (int k = 0; k < kRow; ++k)
        if (hs[k] > g && is_member(s, hs[k])) qmask |= 1u << k;
      for (int sl = r0 + kRow; sl < r1; ++sl) {
        const int h = __ldg(p.g.col_idx + sl);
        if (h > g && is_member(s, h)) {
          if (sl - r0 < 32) qmask |= 1u << (sl - r0); else ++extra;
        }
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, isb);
    if ((tid & 31) == 0 && i < n) s.oldb[i >> 5] = bal;
    const int cnt = __popc(qmask) + extra;
    int tot;
    int j = m + block_excl_scan(cnt, &tot, s.misc, c.par);
    m += tot;
    if (cnt && j + cnt <= p.cap_edges) {
      auto emit = [&](int sl, int h, int eid) {
        s.euv[j] = (uint32_t)i | ((uint32_t)local_id(s, h) << 16);
        c.eids[j] = eid;
        if (R > 0) {  // tree.py:80-87: surcharged when the labels differ or either is None
          uint32_t cls = 0;
          for (int r = 0; r < R; ++r) {
            const int ru = __ldg(p.g.region_ids + (size_t)r * N + g), rv = __ldg(p.g.region_ids + (size_t)r * N + h);
            if (ru != rv || ru < 0 || rv < 0) cls |= 1u << r;
          }
          s.ecls[j] = (uint8_t)cls;
        }
        ++j;
      };
      int se[kRow];
#pragma unroll
      for (int k = 0; k < kRow; ++k) se[k] = (qmask >> k) & 1u ? __ldg(p.g.slot_edge + r0 + k) : 0;
#pragma unroll
      for (int k = 0; k < kRow; ++k)
        if ((qmask >> k) & 1u) emit(r0 + k, hs[k], se[k]);
      qmask &= ~((1u << kRow) - 1u);
      while (qmask) {
        const int sl = r0 + (__ffs(qmask) - 1);
        qmask &= qmask - 1;
        emit(sl, __ldg(p.g.col_idx + sl), __ldg(p.g.slot_edge + sl));
      }
      for (int sl = r0 + 32; extra && sl < r1; ++sl) {
        const int h = __ldg(p.g.col_idx + sl);
        if (h > g && is_member(s, h)) emit(sl, h, __ldg(p.g.slot_edge + sl));
      }
    }
  }
  popsum = warp_sum(popsum);
  slots = warp_sum(slots);
  if ((tid & 31) == 0) { atomicAdd(&s.misc[M_TOT], popsum); atomicAdd(&s.misc[M_CNT], slots); }
  __syncthreads();  // nodes[], eids[] (global) visible to the block; totals complete
  st.tot = s.misc[M_TOT];
  st.m = m;
  *d_sub_out = s.misc[M_CNT];
  if (m > p.cap_edges || m > 65535) return RC_ERR_CAPACITY;
  if (kWilson) return build_local_adjacency(c, st);
  return RC_OK;
}

// Phase timers: compiled in with -DRC_TIMERS (scripts/gpu_phases.sh); thread 0 of every CTA
// adds clock64() deltas per phase to Params.timers (dev tool, not part of the product build).
#ifdef RC_TIMERS
#define RC_PH(i) do { if (threadIdx.x == 0) { unsigned long long* a_ = reinterpret_cast<unsigned long long*>(c.s.misc + 96); \
  const unsigned long long t_ = (unsigned long long)clock64(); a_[i] += t_ - a_[15]; a_[15] = t_; } } while (0)
#else
#define RC_PH(i) do { } while (0)
#endif
enum { PH_OTHER = 0, PH_PICK, PH_GATHER, PH_TREE, PH_EULER, PH_BALANCE, PH_CHOOSE, PH_COMMIT, PH_FLUSH, PH_TICKET, PH_STATE, PH_W1, PH_W2, PH_W3, PH_N };

// te[0..n-2] = the tree edges flagged in tflag, ascending local edge index (8 edges per
// thread step); RC_ERR_DISCONNECTED unless there are exactly n - 1
__device__ int tree_edge_list(Chain& c, const Step& st) {
  const Smem& s = c.s;
  const int tid = threadIdx.x;
  const int n = st.n, m = st.m;
  const int units = (m + 7) >> 3;
  const int per = odd_chunk(units);
  const int lo = min(tid * per, units), hi = min(lo + per, units);
  const uint8_t* fb = reinterpret_cast<const uint8_t*>(s.tflag);
  int cnt = 0;
  for (int u = lo; u < hi; ++u) cnt += __popc((uint32_t)fb[u]);
  int nt;
  int t = block_excl_scan(cnt, &nt, s.misc, c.par);
  __syncthreads();  // te aliases the (now dead) spanning-tree arrays
  if (nt != n - 1) return RC_ERR_DISCONNECTED;
  for (int u = lo; u < hi; ++u) {
    uint32_t x = fb[u];
    while (x) { s.te[t++] = (lid_t)(u * 8 + (__ffs(x) - 1)); x &= x - 1; }
  }
  __syncthreads();
  // endpoints of the tree edges, fetched once (L2, independent loads) for the whole tour
  {
    const uint32_t* __restrict__ euv = s.euv;
    const int T = blockDim.x;
    int q = threadIdx.x;
    for (; q + T < nt; q += 2 * T) {
      const uint32_t e0 = __ldcg(euv + s.te[q]), e1 = __ldcg(euv + s.te[q + T]);
      s.tuv[q] = e0; s.tuv[q + T] = e1;
    }
    if (q < nt) s.tuv[q] = __ldcg(euv + s.te[q]);
  }
  __syncthreads();
  return RC_OK;
}

// ---- P2a: random-weight minimum spanning tree by contracting Boruvka rounds ---------------
// Unique MST under the total order (key32, local edge index) == Kruskal's tree.  Live edges
// carry their endpoints as current component ids and are compacted every round (edges inside
// a component are dropped); the vertex side works on a compacted list of live components,
// which at least halves every round.  Per round: (A) minimum key per component, (B) lowest
// edge index among the edges attaining it, (C) hook, (C') compress the hook forest and list
// the surviving roots, (D) relabel + compact the edges.  On return tflag marks the tree edges
// and te[0..n-2] lists them in ascending local edge index.
constexpr int kCompactK = 8;

__device__ int boruvka_mst(Chain& c, const Step& st, uint32_t tree_no, uint32_t prop_no,
                           const uint32_t* rrank, int* rounds_out) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
  const int n = st.n, m = st.m;
  // weights: one Philox block per four edges
  if (rrank) {
    for (int j = tid; j < m; j += T) s.key[j] = rrank[c.eids[j]];
  } else {
    for (int q = tid; q < (m + 3) >> 2; q += T) {
      const RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_WEIGHT, (uint32_t)q, tree_no, prop_no, 0);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int j = 4 * q + w;
        if (j < m) s.key[j] = finish_key(c, j, r.v[w]);
      }
    }
  }
  {  // the pair's edge list streams in from the cold tier (L2): four loads in flight per thread
    const uint32_t* __restrict__ euv = s.euv;
    int j = tid;
    for (; j + 3 * T < m; j += 4 * T) {
      const uint32_t e0 = __ldcg(euv + j), e1 = __ldcg(euv + j + T), e2 = __ldcg(euv + j + 2 * T), e3 = __ldcg(euv + j + 3 * T);
      s.lab[j] = e0; s.lab[j + T] = e1; s.lab[j + 2 * T] = e2; s.lab[j + 3 * T] = e3;
      s.lj[j] = (lid_t)j; s.lj[j + T] = (lid_t)(j + T); s.lj[j + 2 * T] = (lid_t)(j + 2 * T); s.lj[j + 3 * T] = (lid_t)(j + 3 * T);
    }
    for (; j < m; j += T) { s.lab[j] = __ldcg(euv + j); s.lj[j] = (lid_t)j; }
  }
  for (int i = tid; i < n; i += T) { s.best[i] = 0xffffffffu; s.bpos[i] = 0xffffffffu; s.hook[i] = (lid_t)i; }
  for (int i = tid; i < (m + 31) >> 5; i += T) s.tflag[i] = 0;
  if (tid == 0) { s.misc[M_NLIVE] = 0; s.misc[M_NV] = 0; }
  __syncthreads();
  int nlive = m, nv = n, rounds = 0;
  const lid_t* vcur = nullptr;  // round 0: every node is a component
  lid_t* vnext = s.vlA;
  while (nlive > 0 && nv > 1) {
    ++rounds;
    // A: minimum key per component (native 32-bit shared atomics)
    {  // two adjacent edges per iteration: one 64-bit load of the endpoints, one 32-bit of the ids
      const uint2* lab2 = reinterpret_cast<const uint2*>(s.lab);
      const uint32_t* lj2 = reinterpret_cast<const uint32_t*>(s.lj);
      for (int t = tid; t < (nlive >> 1); t += T) {
        const uint2 ab = lab2[t];
        const uint32_t jj = lj2[t];
        const uint32_t k0 = s.key[jj & 0xffffu], k1 = s.key[jj >> 16];
        atomicMin(&s.best[ab.x & 0xffffu], k0);
        atomicMin(&s.best[ab.x >> 16], k0);
        atomicMin(&s.best[ab.y & 0xffffu], k1);
        atomicMin(&s.best[ab.y >> 16], k1);
      }
      if ((nlive & 1) && tid == 0) {
        const uint32_t ab = s.lab[nlive - 1];
        const uint32_t k = s.key[s.lj[nlive - 1]];
        atomicMin(&s.best[ab & 0xffffu], k);
        atomicMin(&s.best[ab >> 16], k);
      }
    }
    __syncthreads();
    // B: among the (almost always single) edges attaining it, the lowest edge index
    {
      const uint2* lab2 = reinterpret_cast<const uint2*>(s.lab);
      const uint32_t* lj2 = reinterpret_cast<const uint32_t*>(s.lj);
      for (int t = tid; t < (nlive >> 1); t += T) {
        const uint2 ab = lab2[t];
        const uint32_t jj = lj2[t];
        const uint32_t j0 = jj & 0xffffu, j1 = jj >> 16;
        const uint32_t k0 = s.key[j0], k1 = s.key[j1];
        const uint32_t b00 = s.best[ab.x & 0xffffu], b01 = s.best[ab.x >> 16];
        const uint32_t b10 = s.best[ab.y & 0xffffu], b11 = s.best[ab.y >> 16];
        if (k0 == b00) atomicMin(&s.bpos[ab.x & 0xffffu], (j0 << 16) | (uint32_t)(2 * t));
        if (k0 == b01) atomicMin(&s.bpos[ab.x >> 16], (j0 << 16) | (uint32_t)(2 * t));
        if (k1 == b10) atomicMin(&s.bpos[ab.y & 0xffffu], (j1 << 16) | (uint32_t)(2 * t + 1));
        if (k1 == b11) atomicMin(&s.bpos[ab.y >> 16], (j1 << 16) | (uint32_t)(2 * t + 1));
      }
      if ((nlive & 1) && tid == 0) {
        const int q = nlive - 1;
        const uint32_t ab = s.lab[q];
        const uint32_t j = s.lj[q];
        const uint32_t k = s.key[j];
        const uint32_t v = (j << 16) | (uint32_t)q;
        if (k == s.best[ab & 0xffffu]) atomicMin(&s.bpos[ab & 0xffffu], v);
        if (k == s.best[ab >> 16]) atomicMin(&s.bpos[ab >> 16], v);
      }
    }
    __syncthreads();
    // C: every component hooks onto the other side of its lightest edge; of a mutual pair
    // the smaller id stays a root
    for (int q = tid; q < nv; q += T) {
      const int v = vcur ? (int)vcur[q] : q;
      const uint32_t bp = s.bpos[v];
      if (bp == 0xffffffffu) continue;  // a finished component (disconnected input)
      const uint32_t ab = s.lab[bp & 0xffffu];
      const int a = ab & 0xffffu, b = ab >> 16;
      const int other = a == v ? b : a;
      const bool mutual = s.bpos[other] == bp;
      if (!(mutual && v < other)) s.hook[v] = (lid_t)other;
      if (!(mutual && v > other)) atomicOr(&s.tflag[bp >> 21], 1u << ((bp >> 16) & 31u));
    }
    __syncthreads();
    // C': compress the merge forest (benign race: a reader may see a pointer that is already
    // compressed - still an ancestor) and list the roots, the live components of the next round
    for (int q0 = 0; q0 < nv; q0 += T) {
      const int q = q0 + tid;
      bool is_root = false;
      int v = 0;
      if (q < nv) {
        v = vcur ? (int)vcur[q] : q;
        if (s.bpos[v] != 0xffffffffu) {
          int r = v, h;
          while ((h = s.hook[r]) != r) r = h;
          s.hook[v] = (lid_t)r;
          s.best[v] = 0xffffffffu;
          s.bpos[v] = 0xffffffffu;
          is_root = r == v;
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, is_root);
      if (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s.misc[M_NV], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_root) vnext[base + __popc(bal & ((1u << lane) - 1u))] = (lid_t)v;
      }
    }
    __syncthreads();
    nv = s.misc[M_NV];
    vcur = vnext;
    vnext = vnext == s.vlA ? s.vlB : s.vlA;
    if (nv <= 1) break;  // one component left: every remaining edge is inside it
    // D: relabel the live edges to the new roots and drop the ones inside a component;
    // in place: a sweep's reads all precede its writes, and it writes below what it read
    for (int base = 0; base < nlive; base += T * kCompactK) {
      uint32_t nab[kCompactK], nj[kCompactK / 2];  // staged in registers (edge indices two to a word)
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < kCompactK / 2; ++k) nj[k] = 0;
#pragma unroll
      for (int k = 0; k < kCompactK; ++k) {
        const int q = base + k * T + tid;
        nab[k] = 0xffffffffu;
        if (base + k * T >= nlive) break;  // block uniform
        if (q < nlive) {
          const uint32_t ab = s.lab[q];
          const int a = s.hook[ab & 0xffffu], b = s.hook[ab >> 16];
          if (a != b) { nab[k] = (uint32_t)a | ((uint32_t)b << 16); nj[k >> 1] |= (uint32_t)s.lj[q] << (16 * (k & 1)); ++cnt; }
        }
      }
      // warp-aggregated allocation of output slots (their order is immaterial)
      const int incl = warp_incl_scan(cnt);
      __syncthreads();  // every read of this sweep is done
      int wbase = 0;
      if (lane == 31 && incl > 0) wbase = atomicAdd(&s.misc[M_NLIVE], incl);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      int o = wbase + incl - cnt;
#pragma unroll
      for (int k = 0; k < kCompactK; ++k) {
        if (base + k * T >= nlive) break;
        if (nab[k] != 0xffffffffu) { s.lab[o] = nab[k]; s.lj[o] = (lid_t)(nj[k >> 1] >> (16 * (k & 1))); ++o; }
      }
    }
    __syncthreads();
    nlive = s.misc[M_NLIVE];
    __syncthreads();
    if (tid == 0) { s.misc[M_NLIVE] = 0; s.misc[M_NV] = 0; }
  }
  *rounds_out = rounds;
  return tree_edge_list(c, st);
}

__device__ __forceinline__ bool balanced(int tree_pop, int total_pop, double ideal, double thr) {
  // tree.py:412-414, float64 as in the reference
  return fabs(__dsub_rn((double)tree_pop, ideal)) <= thr &&
         fabs(__dsub_rn((double)(total_pop - tree_pop), ideal)) <= thr;
}

// ---- P2b: uniform spanning trees (Wilson, tree.py:97-132): one walker per LANE ------------
// Every node owns a stack of successor draws: draw j at node u is a pure function of (u, j)
// (recom_rng.h), so the tree - and the per-node draw counts - do not depend on the order in
// which walks are made (Wilson 1996), and the reference's loop (tree.py:116-125) can be run
// as it is written: walk from every node not yet in the tree, overwrite next[u] at every
// visit (loop erasure is implicit), retrace the walk to add it to the tree.
//   The walk is a chain of ~8 n dependent steps; more threads do not shorten it (round 1's
// parallel cycle popping needed ~580 dependent rounds per 1000-node tree and kept one warp
// busy for as long as a lone walker needs).  What the SM has in abundance is room for
// INDEPENDENT chains of steps: a step of bipartition_tree (tree.py:679-718) draws trees
// t, t+1, ... until one has a balanced cut, and the trees are independent of each other.
// So lanes 0..W-1 of warp 0 each walk one of the next W trees of this proposal, all on the
// one adjacency the block gathered (same instruction stream, different tree: full SIMT
// efficiency), and the block then examines the trees in order; the lowest tree number with
// a balanced cut wins, exactly as if they had been drawn one after the other.
//   Per walker and node ONE 32-bit word: successor (bits 0-15) | draws consumed (16-30) |
// in the tree (31).  A walk step is one Philox block (the draw), two 16-bit loads (row
// bounds), one load of the successor, one store of the word, one load of the next word.
constexpr uint32_t kInTree = 0x8000u;  // walker state: successor (bits 0-14) | in the tree (bit 15)
constexpr uint32_t kNext = 0x7fffu;
constexpr uint32_t kRing = 128;  // draws per walker's ring (power of two)

// The walks of one batch, all 32 lanes of the walker warp in one loop: lanes 0..n_walk-1
// walk (tree tree_base + lane), every lane helps to refill the draw rings.
//   Draw t of a tree's walk (recom_rng.h RC_DRAW_WALK: t counts the successor draws of the
// tree in the order the loop of tree.py:116-125 makes them - start nodes ascending, every
// walk to its end) does not depend on where the walk is, so the generator is off the
// walk's dependency chain: a ring of 64 draws per walker, refilled 32 at a time by the
// WHOLE warp (one Philox block per lane: 8 lanes per walker) whenever a walker has used up
// its older half.  What is left on the chain of a step is ONE load: the adjacency entry
// picked by the draw names the next node and its degree, and the entry after it is
// requested before the "has the walk reached the tree?" test of the node resolves.
//   The loop is a function of its own (not inlined: its registers are not shared with the
// 200 KB of kernel around it) over accessors: in shared memory the arrays are addressed by
// 32-bit shared-window offsets held in registers with explicit ld/st.shared (the compiler
// otherwise re-derives the window address - a special-register read - in front of the
// loads); in spill mode they are global pointers.  scripts/dev/walk_bench.cu has the loop in
// isolation: 104 cycles per step on B200 (a dependent LDS alone is 34).
struct WalkArgs {
  uint64_t seed;
  uint32_t gchain, tree_base, prop_no;
  int n, walker, sh, cap_nodes;
  void *nbr, *wst, *wring, *wdeg;
  int32_t* wmeta;
  unsigned long long* timers;  // RC_TIMERS builds
};

struct MemShared {  // level 0: 32-bit shared-window addresses
  static constexpr int kFirst = 2, kRest = 8;  // burst lengths (steps of straight-line code)
  static constexpr uint32_t kIdMask = 0xffffu, kDegShift = 16;  // adjacency entry: neighbour | its degree << kDegShift
  uint32_t nbr, S, ring, deg;
  __device__ __forceinline__ uint32_t ld_nbr(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(nbr + 4u * i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_ring(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + 4u * i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_S(uint32_t i) const { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(S + 2u * i) : "memory"); return v; }
  __device__ __forceinline__ void st_S(uint32_t i, uint32_t v) const { asm volatile("st.shared.u16 [%0], %1;" :: "r"(S + 2u * i), "h"((uint16_t)v) : "memory"); }
  __device__ __forceinline__ uint32_t ld_deg(uint32_t i) const { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(deg + 2u * i) : "memory"); return v; }
};

// level 0 with 16-bit adjacency entries (pairs of at most 2048 rows, degrees below 32): half the
// shared memory of the adjacency, one more CTA - four more walks - per SM at config 3
struct MemShared16 {
  static constexpr int kFirst = 2, kRest = 8;
  static constexpr uint32_t kIdMask = 0x7ffu, kDegShift = 11;
  uint32_t nbr, S, ring, deg;
  __device__ __forceinline__ uint32_t ld_nbr(uint32_t i) const { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(nbr + 2u * i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_ring(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + 4u * i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_S(uint32_t i) const { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(S + 2u * i) : "memory"); return v; }
  __device__ __forceinline__ void st_S(uint32_t i, uint32_t v) const { asm volatile("st.shared.u16 [%0], %1;" :: "r"(S + 2u * i), "h"((uint16_t)v) : "memory"); }
  __device__ __forceinline__ uint32_t ld_deg(uint32_t i) const { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(deg + 2u * i) : "memory"); return v; }
};

// level 1: the adjacency is read from the CTA's global block (read-only during the walks: plain
// cached loads, kept in program order by asm volatile), the walker state stays in shared memory
struct MemMixed {
  static constexpr int kFirst = 1, kRest = 2;  // a step is an L2 round trip: a wasted one costs more than a branch
  static constexpr uint32_t kIdMask = 0xffffu, kDegShift = 16;
  const uint32_t* nbr; const uint16_t* deg;
  uint32_t S, ring;
  __device__ __forceinline__ uint32_t ld_nbr(uint32_t i) const { uint32_t v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(nbr + i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_ring(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + 4u * i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_S(uint32_t i) const { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(S + 2u * i) : "memory"); return v; }
  __device__ __forceinline__ void st_S(uint32_t i, uint32_t v) const { asm volatile("st.shared.u16 [%0], %1;" :: "r"(S + 2u * i), "h"((uint16_t)v) : "memory"); }
  __device__ __forceinline__ uint32_t ld_deg(uint32_t i) const { return deg[i]; }
};

struct MemGlobal {  // level 2: state and adjacency in the CTA's global block
  static constexpr int kFirst = 1, kRest = 2;
  static constexpr uint32_t kIdMask = 0xffffu, kDegShift = 16;
  const uint32_t* nbr; uint16_t* S; const uint16_t* deg;
  uint32_t ring;
  __device__ __forceinline__ uint32_t ld_nbr(uint32_t i) const { uint32_t v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(nbr + i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_ring(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + 4u * i) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ld_S(uint32_t i) const { uint16_t v; asm volatile("ld.global.u16 %0, [%1];" : "=h"(v) : "l"(S + i) : "memory"); return v; }
  __device__ __forceinline__ void st_S(uint32_t i, uint32_t v) const { asm volatile("st.global.u16 [%0], %1;" :: "l"(S + i), "h"((uint16_t)v) : "memory"); }
  __device__ __forceinline__ uint32_t ld_deg(uint32_t i) const { return deg[i]; }
};

template <typename Mem, int sh>  // sh: log2 of the adjacency slots per node, a compile-time shift
__device__ __noinline__ void wilson_walk_loop(const WalkArgs a, Mem M) {
  const int lane = threadIdx.x & 31;
  const int n = a.n, w = a.walker;
  constexpr uint32_t D = 1u << sh;
  uint32_t* const wring = (uint32_t*)a.wring + w * kRing;
  // one Philox block per lane: draws [4 b, 4 b + 4) of this walker into its ring
  auto fill = [&](uint32_t b) {
    const RcPhilox r = rc_draw(a.seed, a.gchain, RC_DRAW_WALK, b, a.tree_base + (uint32_t)w, a.prop_no, 0);
    *reinterpret_cast<uint4*>(wring + ((b & (kRing / 4 - 1)) << 2)) = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  };
  fill((uint32_t)lane);  // kRing / 4 == 32 blocks
  __syncwarp();
  // entry of node u (degree du) picked by the random word r
  auto pick = [&](uint32_t u, uint32_t du, uint32_t r) -> uint32_t {
    uint32_t idx = rc_below32(r, du);
    uint32_t row = u << sh;
    while (du > D) {  // a long row: D - 1 entries, then the link to the next part
      if (idx < D - 1u) break;
      row = M.ld_nbr(row + D - 1u) << sh;
      idx -= D - 1u;
      du -= D - 1u;
    }
    return M.ld_nbr(row + idx);
  };
  // Every lane runs the reference's loop as it is written - find the next node outside the
  // tree, walk until the tree is hit, retrace - and comes back to the top of the outer loop,
  // where the warp meets for the refill vote, after kQuota draws at the latest (the ring then
  // still holds them: a refill leaves more than 32 unread).  Lanes in different phases
  // diverge inside a trip; a lone warp pays ~15-20 cycles for every branch it has to resolve,
  // so the phases are tight loops of their own rather than cases of a per-event switch.
  constexpr int kQuota = 58;  // 2 + 7 x 8: below kRing / 2 with the look-ahead of a burst
  bool live = lane == 0 && a.wmeta[WM_ERR + w] == 0;
  bool walking = false, first = false;  // a walk is under way (u, e hold its state); its first burst is due
  int err = 0;
  uint32_t scan = 0, start = 0, u = 0, e = 0, t = 0, filled = kRing;
  int budget = n < (1 << 19) ? 2048 * n + 65536 : 0x7fffffff;  // a walk that cannot reach the tree (disconnected input plan)
#ifdef RC_TIMERS
  unsigned long long tm_iter = 0;
  const unsigned long long tm0 = (unsigned long long)clock64();
#endif
  while (__any_sync(0xffffffffu, live)) {
#ifdef RC_TIMERS
    ++tm_iter;
#endif
    {  // refill: once the older half of the ring is used up, 16 lanes make the next kRing / 2 draws
      const uint32_t fw = __shfl_sync(0xffffffffu, (live && t + kRing / 2 >= filled) ? filled : 0u, 0);
      if (fw) {
        if (lane < 16) fill((fw >> 2) + (uint32_t)lane);
        filled += kRing / 2;
        __syncwarp();
      }
    }
    if (live) {
      int quota = kQuota;
      for (;;) {
        if (!walking) {  // the next node that is not in the tree starts a walk
          while (scan < (uint32_t)n && (M.ld_S(scan) & kInTree)) ++scan;
          if (scan >= (uint32_t)n) { live = false; break; }
          start = u = scan++;
          e = pick(u, M.ld_deg(u), M.ld_ring(t & (kRing - 1u)));  // the entry the first step of the walk takes
          walking = true;
          first = true;
        }
        // state: u = current node (not in the tree), e = the adjacency entry draw t picked at u.
        // A burst is K steps of straight-line code: the steps behind the one that reaches the
        // tree are predicated off instead of branched around (a lone warp waits 15-20 cycles
        // for every branch it resolves), their loads still go out (to valid addresses: e is a
        // valid entry) and are ignored.  Most walks end within two steps (the tree is next
        // door once it has grown), the few long ones make most of the draws: a walk starts
        // with a burst of 2 and goes on in bursts of 8.
        bool hit = false;
        auto burst = [&](auto kc) {
          constexpr int K = decltype(kc)::value;
          uint32_t rr[K];
#pragma unroll
          for (int k = 0; k < K; ++k) rr[k] = M.ld_ring((t + 1u + (uint32_t)k) & (kRing - 1u));  // draws t+1 .. t+K
          bool alive = true;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const uint32_t v = e & Mem::kIdMask;
            if (alive) M.st_S(u, v);
            const uint32_t en = pick(v, e >> Mem::kDegShift, rr[k]);  // the entry the NEXT step takes (draw t + 1 + k)
            const uint32_t wv = M.ld_S(v);
            t += alive ? 1u : 0u;
            const bool go = alive && !(wv & kInTree);
            hit = hit || (alive && (wv & kInTree));
            u = go ? v : u;
            e = go ? en : e;
            alive = go;
          }
        };
        if (first) { first = false; burst(std::integral_constant<int, Mem::kFirst>()); quota -= Mem::kFirst; }
        while (!hit && quota >= Mem::kRest) { burst(std::integral_constant<int, Mem::kRest>()); quota -= Mem::kRest; }
        if (!hit && quota > 0) quota = 0;
        if (!hit) break;  // out of draws for this trip, the walk goes on in the next
        walking = false;
        for (u = start;;) {  // retrace: the loop-erased walk joins the tree
          const uint32_t w = M.ld_S(u);
          if (w & kInTree) break;
          M.st_S(u, w | kInTree);
          u = w & kNext;
        }
        if (quota <= 0) break;
      }
      budget -= kQuota;
      if (budget < 0) { err = RC_ERR_DISCONNECTED; live = false; }
    }
  }
  if (lane == 0 && a.wmeta[WM_ERR + w] == 0) {
    a.wmeta[WM_DRAWS + w] = (int)t;
    a.wmeta[WM_ERR + w] = err;
  }
#ifdef RC_TIMERS
  if (lane == 0 && w == 0) {  // slots 11-12: cycles inside the walk loop, its trips
    atomicAdd(a.timers + 11, (unsigned long long)clock64() - tm0);
    atomicAdd(a.timers + 12, tm_iter);
  }
#endif
}

// walker w of the batch (tree tree_base + w) on the calling warp: lane 0 walks, the other
// lanes help with the draws
__device__ __forceinline__ void wilson_walk_warp(const Chain& c, const Step& st, uint32_t tree_base, int w,
                                                 uint32_t prop_no) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  WalkArgs a;
  a.seed = p.cfg.seed; a.gchain = c.gchain; a.tree_base = tree_base; a.prop_no = prop_no;
  a.n = st.n; a.walker = w; a.sh = p.nbr_shift; a.cap_nodes = p.cap_nodes;
  a.nbr = s.nbr; a.wst = s.wst; a.wring = s.wring; a.wdeg = s.wdeg; a.wmeta = s.wmeta;
  a.timers = reinterpret_cast<unsigned long long*>(s.misc + 96);
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s.wring + w * kRing);
  if (p.wilson_level == 0 && p.nbr16) {
    MemShared16 M;
    M.nbr = (uint32_t)__cvta_generic_to_shared(s.nbr);
    M.S = (uint32_t)__cvta_generic_to_shared(s.wst + (size_t)w * p.cap_nodes);
    M.ring = ring;
    M.deg = (uint32_t)__cvta_generic_to_shared(s.wdeg);
    if (a.sh == 2) wilson_walk_loop<MemShared16, 2>(a, M); else wilson_walk_loop<MemShared16, 3>(a, M);
  } else if (p.wilson_level == 0) {
    MemShared M;
    M.nbr = (uint32_t)__cvta_generic_to_shared(s.nbr);
    M.S = (uint32_t)__cvta_generic_to_shared(s.wst + (size_t)w * p.cap_nodes);
    M.ring = ring;
    M.deg = (uint32_t)__cvta_generic_to_shared(s.wdeg);
    if (a.sh == 2) wilson_walk_loop<MemShared, 2>(a, M); else wilson_walk_loop<MemShared, 3>(a, M);
  } else if (p.wilson_level == 1) {
    MemMixed M;
    M.nbr = s.nbr; M.deg = s.wdeg; M.ring = ring;
    M.S = (uint32_t)__cvta_generic_to_shared(s.wst + (size_t)w * p.cap_nodes);
    if (a.sh == 2) wilson_walk_loop<MemMixed, 2>(a, M); else wilson_walk_loop<MemMixed, 3>(a, M);
  } else {
    MemGlobal M;
    M.nbr = s.nbr; M.S = s.wst + (size_t)w * p.cap_nodes; M.ring = ring; M.deg = s.wdeg;
    if (a.sh == 2) wilson_walk_loop<MemGlobal, 2>(a, M); else wilson_walk_loop<MemGlobal, 3>(a, M);
  }
}

// TEST HOOK: one tree from the reference's recorded stacks (tree.py:119) - visit j of node u
// takes recorded successor j of u; one thread, visit counts in wsum (free until the tree is
// evaluated).  Same loop as the reference's: walk, overwrite, retrace.
__device__ void wilson_walk_recorded(const Chain& c, const Step& st, long long tree_idx) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int n = st.n, sh = p.nbr_shift;
  uint16_t* S = s.wst;
  int32_t* cnt = s.wsum;
  const int64_t* sp = p.rp.stack_ptr + tree_idx * ((long long)p.g.N + 1);
  int err = 0, t = 0;
  for (int node = 0; node < n && !err; ++node) {
    int u = node;
    while (!(S[u] & kInTree)) {
      const int gu = c.nodes[u], cu = cnt[u], d = s.wdeg[u];
      int v = -1;
      if (sp[gu] + (long long)cu < sp[gu + 1]) {
        const int gv = p.rp.stack_val[sp[gu] + cu];
        if (gv >= 0 && gv < p.g.N && is_member(s, gv)) {
          const int lv = local_id(s, gv);
          for (int q = 0; q < d; ++q) if (nbr_id(s, p, u, d, q) == lv) v = lv;
        }
      }
      if (v < 0) { err = RC_ERR_REPLAY; break; }
      cnt[u] = cu + 1;
      ++t;
      S[u] = (uint16_t)v;
      u = v;
    }
    for (u = node; !err && !(S[u] & kInTree);) { const uint32_t w = S[u]; S[u] = (uint16_t)(w | kInTree); u = (int)(w & kNext); }
  }
  s.wmeta[WM_DRAWS] = t;
  s.wmeta[WM_ERR] = err;
}

// Trees tree_base .. tree_base + n_walk - 1 of proposal prop_no, one per lane of the walker
// warp (replay: ONE recorded tree).  On return wst[w] holds tree w as parent pointers (the
// root points at itself), wmeta its root, draw count and status.
__device__ void wilson_walk(Chain& c, const Step& st, uint32_t tree_base, int n_walk, uint32_t prop_no,
                            long long tree_idx_base) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, CN = p.cap_nodes;
  const bool replay = p.rp.n_steps > 0;
  if (tid < n_walk) {
    int root, err = 0;
    if (replay) {
      const int gr = p.rp.wilson_root[tree_idx_base + tid];
      const bool ok = gr >= 0 && gr < p.g.N && is_member(s, gr);
      root = ok ? local_id(s, gr) : 0;
      if (!ok) err = RC_ERR_REPLAY;
    } else {
      const RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_WROOT, 0, tree_base + (uint32_t)tid, prop_no, 0);
      root = (int)rc_below(r.v[0], r.v[1], (uint32_t)n);
    }
    s.wmeta[WM_ROOT + tid] = root;
    s.wmeta[WM_ERR + tid] = err;
    s.wmeta[WM_DRAWS + tid] = 0;
  }
  for (int w = 0; w < n_walk; ++w)
    for (int i = tid; i < n; i += T) s.wst[(size_t)w * CN + i] = 0;
  if (replay) for (int i = tid; i < n; i += T) s.wsum[i] = 0;  // visit counts of the recorded walk
  __syncthreads();
  if (tid < n_walk) {
    const int root = s.wmeta[WM_ROOT + tid];
    s.wst[(size_t)tid * CN + root] = (uint16_t)((uint32_t)root | kInTree);
  }
  for (int i = tid; i < n; i += T)  // a node without a neighbour inside the pair: the input plan is not contiguous
    if (s.wdeg[i] == 0 && n > 1)
      for (int w = 0; w < n_walk; ++w) s.wmeta[WM_ERR + w] = RC_ERR_DISCONNECTED;
  __syncthreads();
  if (replay) {
    if (tid == 0 && s.wmeta[WM_ERR] == 0) wilson_walk_recorded(c, st, tree_idx_base);
  } else if ((tid >> 5) < n_walk) {
    // one walker per warp: warp w issues from scheduler w % 4, the walkers of a CTA (and of
    // the CTAs sharing the SM) spread over the four of them
    wilson_walk_warp(c, st, tree_base, tid >> 5, prop_no);
  }
  __syncthreads();
}

// Subtree populations of tree w (parent pointers): every node adds its population to all
// its proper ancestors.  O(n * depth) shared-memory atomics - a uniform spanning tree of a
// planar pair is ~n^(5/8) deep - against the ~8 n dependent steps of the walk that made it.
// Also: whas (nodes with a child), M_ROOTKIDS, and the number of balanced tree edges
// (tree.py:411-423: the edge above u is balanced when the two sides are).
__device__ int wilson_eval(Chain& c, const Step& st, int w, double ideal, double thr) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n;
  const uint16_t* __restrict__ S = s.wst + (size_t)w * p.cap_nodes;
  const int root = s.wmeta[WM_ROOT + w];
  for (int i = tid; i < n; i += T) s.wsum[i] = s.pop[i];
  for (int i = tid; i < (n + 31) >> 5; i += T) s.whas[i] = 0;
  if (tid == 0) { s.misc[M_ROOTKIDS] = 0; s.misc[M_NBAL] = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    if (i == root) continue;
    const int x = s.pop[i];
    int q = (int)(S[i] & kNext);
    atomicOr(&s.whas[q >> 5], 1u << (q & 31));
    if (q == root) atomicAdd(&s.misc[M_ROOTKIDS], 1);
    while (q != root) {
      atomicAdd(&s.wsum[q], x);
      q = (int)(S[q] & kNext);
    }
  }
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < n; i += T) cnt += (i != root && balanced(s.wsum[i], st.tot, ideal, thr)) ? 1 : 0;
  cnt = warp_sum(cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(&s.misc[M_NBAL], cnt);
  __syncthreads();
  return s.misc[M_NBAL];
}

// ---- P3: root the tree and compute every subtree population ----------------------------
// Euler tour of the tree (arcs 2t: eu->ev, 2t+1: ev->eu of tree edge t), ranked by a
// work-efficient sublist scheme (Helman-JaJa): every kStride-th arc starts a sublist; one
// thread walks each sublist (length and weight), the few hundred sublists are ranked by
// pointer doubling, a second walk writes each arc's tour position and exclusive weight
// prefix.  Arc (y -> x) carries pop[x] when it is the reverse of x's first out-arc (and x is
// not the tour root): every arc entering x lies between the arc that descends into x and
// the one that climbs out of it, so the weights between those two sum to the population
// below the edge, whichever way the edge turns out to be oriented.  After this call:
//   P[a]     = tour position of arc a
//   ent[i]   = position of the arc entering node i (0xffff for the root)
//   cur[t]   = population below tree edge t (the side away from the root)
// The set of balanced edges does not depend on the root, so
// the tour is rooted wherever convenient; the reference's root only decides the labels.
constexpr int kStrideLog = 3, kStride = 1 << kStrideLog;
constexpr uint32_t kNone = 0xffffu;

__device__ void euler_subtree_pops(Chain& c, const Step& st, int root) {
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, nt = n - 1, L = 2 * nt;
  // E1/E2: out-arcs of every node as a linked list (head, nxt)
  for (int i = tid; i < n; i += T) s.head[i] = kNone;
  __syncthreads();
  for (int t = tid; t < nt; t += T) {
    const uint32_t uv = s.tuv[t];
    s.nxt[2 * t] = (lid_t)atomicExch(&s.head[uv & 0xffffu], (uint32_t)(2 * t));
    s.nxt[2 * t + 1] = (lid_t)atomicExch(&s.head[uv >> 16], (uint32_t)(2 * t + 1));
  }
  __syncthreads();
  // E3: successor of arc a = the out-arc after reverse(a) around head(a); arc weights
  const int first = (int)s.head[root];
  for (int a = tid; a < L; a += T) {
    const uint32_t uv = s.tuv[a >> 1];
    const int x = (a & 1) ? (int)(uv & 0xffffu) : (int)(uv >> 16);  // node the arc enters
    const uint32_t hx = s.head[x];
    uint32_t nx = s.nxt[a ^ 1];
    if (nx == kNone) nx = hx;
    s.succ[a] = (lid_t)((int)nx == first ? a : (int)nx);  // the arc that closes the tour ends the list
    if (!((uint32_t)(a ^ 1) == hx && x != root)) s.W[a] = 0;
  }
  // ... and the arcs that carry a population: the reverse of the first out-arc of every node
  // but the root (node order: the populations stream in from the cold tier)
  for (int x = tid; x < n; x += T)
    if (x != root) s.W[s.head[x] ^ 1u] = s.pop[x];
  __syncthreads();
  // E5: walk 1 - length and weight of every sublist.  A walk stops in front of the next
  // splitter (every kStride-th arc, or the tour head) or at the arc that ends the tour.
  const int S = (L + kStride - 1) >> kStrideLog;
  const bool extra = (first & (kStride - 1)) != 0;  // the tour head is not a regular splitter
  const int S1 = S + (extra ? 1 : 0);
  uint32_t* __restrict__ XA = s.spX;
  uint32_t* __restrict__ XB = s.spX + s.spCap;
  int32_t* __restrict__ YA = s.spY;
  int32_t* __restrict__ YB = s.spY + s.spCap;
  for (int k = tid; k < S1; k += T) {
    int a = k < S ? (k << kStrideLog) : first;
    int len = 0, w = 0, nx;
    for (;;) {
      w += s.W[a];
      ++len;
      nx = s.succ[a];
      if (((nx & (kStride - 1)) == 0) | (nx == first) | (nx == a)) break;
      a = nx;
    }
    if (nx == a) {  // terminal sublist: distance 0 to itself, its length is kept aside
      s.misc[M_TAILLEN] = len;
      XA[k] = (uint32_t)k << 16;
      YA[k] = 0;
    } else {
      const int next = (nx == first && extra) ? S : (nx >> kStrideLog);
      XA[k] = ((uint32_t)next << 16) | (uint32_t)len;
      YA[k] = w;
    }
  }
  __syncthreads();
  // E6: rank the sublists (distance / weight from the start of sublist k to the terminal one)
  for (int span = 1; span < S1; span <<= 1) {
    for (int k = tid; k < S1; k += T) {
      const uint32_t x = XA[k];
      const uint32_t y = XA[x >> 16];
      XB[k] = (y & 0xffff0000u) | ((x & 0xffffu) + (y & 0xffffu));
      YB[k] = YA[k] + YA[x >> 16];
    }
    __syncthreads();
    uint32_t* tx = XA; XA = XB; XB = tx;
    int32_t* ty = YA; YA = YB; YB = ty;
  }
  // E7: walk 2 - tour position and exclusive weight prefix (up to a constant) of every arc
  const int tail0 = L - s.misc[M_TAILLEN];
  for (int k = tid; k < S1; k += T) {
    int a = k < S ? (k << kStrideLog) : first;
    int pos = tail0 - (int)(XA[k] & 0xffffu);
    int w = -YA[k];
    for (;;) {
      const int wa = s.W[a];
      const int nx = s.succ[a];
      s.P[a] = (lid_t)pos;
      s.W[a] = w;
      ++pos;
      w += wa;
      if (((nx & (kStride - 1)) == 0) | (nx == first) | (nx == a)) break;
      a = nx;
    }
  }
  __syncthreads();
  // E8: population below every tree edge (cur aliases succ, ent the sublist weights: both dead now)
  if (tid == 0) s.ent[root] = 0xffffu;
  for (int t = tid; t < nt; t += T) {
    const uint32_t uv = s.tuv[t];
    const int p0 = s.P[2 * t], p1 = s.P[2 * t + 1];
    const bool down0 = p0 < p1;  // arc 2t (eu -> ev) points away from the root
    const int child = down0 ? (int)(uv >> 16) : (int)(uv & 0xffffu);
    s.ent[child] = (lid_t)(down0 ? p0 : p1);
    const int w0 = s.W[2 * t], w1 = s.W[2 * t + 1];
    s.cur[t] = down0 ? w1 - w0 : w0 - w1;
  }
  __syncthreads();
}

// k-th element (ascending index) with pred(i) true among [0, count); block uniform result.
template <typename Pred>
__device__ int select_kth(Chain& c, int count, int kth, Pred pred) {
  const Smem& s = c.s;
  const int tid = threadIdx.x;
  const int per = odd_chunk(count);
  const int lo = min(tid * per, count), hi = min(lo + per, count);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += pred(i) ? 1 : 0;
  if (tid == 0) s.misc[M_PICK] = -1;
  int tot;
  const int excl = block_excl_scan(cnt, &tot, s.misc, c.par);
  if (kth >= excl && kth < excl + cnt) {
    int k = kth - excl;
    for (int i = lo; i < hi; ++i)
      if (pred(i) && k-- == 0) { s.misc[M_PICK] = i; break; }
  }
  __syncthreads();
  const int r = s.misc[M_PICK];
  __syncthreads();
  return r;
}

// ---- P5: region-aware cut choice (tree.py:501-578); returns tree edge index --------------
__device__ int choose_cut_by_weight(Chain& c, const Step& st, int policy, uint32_t tree_no,
                                    uint32_t prop_no, const uint32_t* rrank, double ideal, double thr) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int nt = st.n - 1, R = p.g.R;
  // subsets of region columns ordered by (size desc, summed surcharge desc), stable
  int subsets[7], ns = 0;
  if (policy == RC_CUT_REGION_PREF && R >= 1 && R <= 3) {
    double sums[7];
    for (int size = 1; size <= R; ++size)
      for (int q = 1; q < (1 << R); ++q)
        if (__popc(q) == size) {
          double sm = 0.0;
          for (int r = 0; r < R; ++r) if (q >> r & 1) sm = __dadd_rn(sm, p.g.surcharge[r]);
          subsets[ns] = q; sums[ns] = sm; ++ns;
        }
    for (int i = 1; i < ns; ++i) {
      const int q = subsets[i]; const double v = sums[i]; int k = i - 1;
      while (k >= 0) {
        const int ck = __popc(subsets[k]), cq = __popc(q);
        if (!(ck < cq || (ck == cq && sums[k] < v))) break;
        subsets[k + 1] = subsets[k]; sums[k + 1] = sums[k]; --k;
      }
      subsets[k + 1] = q; sums[k + 1] = v;
    }
  }
  unsigned long long* best = reinterpret_cast<unsigned long long*>(&s.misc[M_BEST64]);
  if (tid == 0) { s.misc[M_SUBSET] = 0; *best = 0ull; }
  __syncthreads();
  // tree.py:552-561: a cut "crosses" a region column when its endpoints' labels differ
  // (None == None there, unlike the weight loop)
  auto crossed = [&](int j) {
    const uint32_t uv = s.euv[j];
    const int gu = c.nodes[uv & 0xffffu], gv = c.nodes[uv >> 16];
    int cm = 0;
    for (int r = 0; r < R && r < 3; ++r)
      if (p.g.region_ids[(size_t)r * p.g.N + gu] != p.g.region_ids[(size_t)r * p.g.N + gv]) cm |= 1 << r;
    return cm;
  };
  for (int t = tid; t < nt; t += T) {
    if (!balanced(s.cur[t], st.tot, ideal, thr)) continue;
    const int cm = crossed(s.te[t]);
    int has = 0;
    for (int i = 0; i < ns; ++i) if ((cm & subsets[i]) == subsets[i]) has |= 1 << i;
    if (has) atomicOr(&s.misc[M_SUBSET], has);
  }
  __syncthreads();
  const int has = s.misc[M_SUBSET];
  const int need = has ? subsets[__ffs(has) - 1] : 0;
  // max key, ties -> lowest tree edge: pack (key << 32 | ~t)
  for (int t = tid; t < nt; t += T) {
    if (!balanced(s.cur[t], st.tot, ideal, thr)) continue;
    const int j = s.te[t];
    if (need && (crossed(j) & need) != need) continue;
    const unsigned long long v = ((unsigned long long)edge_key(c, j, tree_no, prop_no, rrank) << 32) |
                                 (unsigned long long)(0xffffffffu - (uint32_t)t);
    atomic_max_u64_smem(best, v);
  }
  __syncthreads();
  const int tstar = (int)(0xffffffffu - (uint32_t)(*best & 0xffffffffull));
  __syncthreads();
  return tstar;
}

struct ProposalResult {
  int status;     // RC_OK or error
  int accepted;   // committed
  int trees;      // trees drawn
  int nbal;       // balanced cuts of the last tree
};

// Wilson: is local edge j an edge of the tree S (parent pointers)?  Returns the child
// endpoint or -1.  (Two adjacent nodes cannot point at each other: the root points at itself.)
__device__ __forceinline__ int wilson_tree_child(const Smem& s, const uint16_t* __restrict__ S, int j) {
  const uint32_t uv = s.euv[j];
  const int a = (int)(uv & 0xffffu), b = (int)(uv >> 16);
  if ((int)(S[a] & kNext) == b) return a;
  if ((int)(S[b] & kNext) == a) return b;
  return -1;
}

// Commit of a proposal whose new side bitmap is in newb (bit i: node i takes the higher id b):
// optional cut-edge acceptance (accept.py:19-35), labels, cut-edge bitmask, Tally.
// Returns true when the proposal was committed.
__device__ bool commit_proposal(Chain& c, const Step& st, int cut_count, long long pop_a, long long pop_b,
                                uint32_t prop_no) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  if (p.cfg.accept_rule == RC_ACCEPT_CUT_EDGE) {
    // accept.py:19-35: the number of cut edges the proposal WOULD have, before anything is
    // written; a rejected proposal leaves the state alone but still counts as a step
    if (tid == 0) s.misc[M_DELTA] = 0;
    __syncthreads();
    int d = 0;
    for (int j = tid; j < st.m; j += T) {
      const uint32_t uv = s.euv[j];
      const int u = (int)(uv & 0xffffu), v = (int)(uv >> 16);
      const int was = ((s.oldb[u >> 5] >> (u & 31)) ^ (s.oldb[v >> 5] >> (v & 31))) & 1u;
      const int now = ((s.newb[u >> 5] >> (u & 31)) ^ (s.newb[v >> 5] >> (v & 31))) & 1u;
      d += now - was;
    }
    d = warp_sum(d);
    if ((tid & 31) == 0 && d) atomicAdd(&s.misc[M_DELTA], d);
    __syncthreads();
    const int new_count = cut_count + s.misc[M_DELTA];
    __syncthreads();
    const double ratio = __ddiv_rn((double)cut_count, (double)new_count);
    const RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_ACCEPT, 0, 0, prop_no, 0);
    if (!(rc_unit_double(r.v[0], r.v[1]) < (ratio < 1.0 ? ratio : 1.0))) return false;  // step consumed, old state kept
  }
  if (tid == 0) s.misc[M_DELTA] = 0;
  for (int i = tid; i < st.n; i += T) {  // partition/assignment.py:76-86
    const bool nb = (s.newb[i >> 5] >> (i & 31)) & 1u, ob = (s.oldb[i >> 5] >> (i & 31)) & 1u;
    if (nb != ob) c.assign[c.nodes[i]] = (rc_label_t)(nb ? st.b : st.a);
  }
  __syncthreads();
  int delta = 0;
  for (int j = tid; j < st.m; j += T) {  // updaters/cut_edges.py:118-120
    const uint32_t uv = s.euv[j];
    const int u = (int)(uv & 0xffffu), v = (int)(uv >> 16);
    const bool was = ((s.oldb[u >> 5] >> (u & 31)) ^ (s.oldb[v >> 5] >> (v & 31))) & 1u;
    const bool now = ((s.newb[u >> 5] >> (u & 31)) ^ (s.newb[v >> 5] >> (v & 31))) & 1u;
    if (now != was) {
      const int eid = c.eids[j];
      atomicXor(c.mask + (eid >> 5), 1u << (eid & 31));
      delta += (int)now - (int)was;
    }
  }
  delta = warp_sum(delta);
  if ((tid & 31) == 0 && delta) atomicAdd(&s.misc[M_DELTA], delta);
  __syncthreads();
  if (tid == 0) {
    c.tally[st.a] = pop_a;  // updaters/tally.py:162-165
    c.tally[st.b] = pop_b;
    p.st.d_cut_count[c.chain] = cut_count + s.misc[M_DELTA];
    if (p.st.d_last_pair) { p.st.d_last_pair[2 * c.chain] = st.a; p.st.d_last_pair[2 * c.chain + 1] = st.b; }
  }
  __syncthreads();
  return true;
}

// This CTA's descriptor, a pure function of the launch parameters: the out-of-line functions
// below build their own instead of taking the caller's (a Chain whose address escapes would
// pin its ~40 pointers in local memory for the whole step path, which otherwise rematerialises
// them from the shared-memory base).
// The launch's Params once more, in constant memory: out-of-line device functions (long_retry,
// help_once) cannot reach the kernel's parameter bank, and through a generic pointer to it every
// field is a load (plus the descriptor set-up) that has to be repeated after each barrier.
// Written by rc_engine_run on the launch's stream before a shared-retry launch; launches that
// use it are ordered by an event (recom_abi.cu).
__constant__ Params g_params;

template <bool kWilson, bool kSpill>
__device__ __forceinline__ void chain_layout(Chain& c, const Params& p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  c.p = &p;
  if constexpr (kWilson)
    wilson_layout(&c.s, smem_raw, p.g.N, p.cap_nodes, p.cap_edges, p.g.R, p.walkers, p.nbr_shift, p.wilson_level, p.mpref_shift,
                  p.spill + (size_t)blockIdx.x * p.spill_stride, p.cold + (size_t)blockIdx.x * p.cold_stride, nullptr, nullptr,
                  p.nbr16 ? 2 : 4);
  else
    smem_layout(&c.s, smem_raw, p.g.N, p.cap_nodes, p.cap_edges, p.g.R,
                kSpill ? p.spill + (size_t)blockIdx.x * p.spill_stride : nullptr,
                p.cold + (size_t)blockIdx.x * p.cold_stride);
  c.nodes = p.nodes + (size_t)blockIdx.x * p.cap_nodes;
  c.eids = p.eids + (size_t)blockIdx.x * p.cap_edges;
}
__device__ __forceinline__ void chain_bind(Chain& c, const Params& p, int chain) {
  c.chain = chain;
  c.gchain = (uint32_t)(chain + p.cfg.chain_offset);
  c.assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
  c.mask = p.st.d_cut_mask + (size_t)chain * p.st.mask_stride;
  c.tally = p.st.d_tally + (size_t)chain * p.cfg.n_districts;
}

// =========================================================================================
// Shared retries.  bipartition_tree (tree.py:679-718) draws trees t = 0, 1, 2, ... until one
// has a balanced cut; the trees are independent of each other (tree t is a pure function of
// chain, proposal and t), so CTAs that have no chain to step - the tail of a launch, or a
// launch with fewer chains than CTAs - draw the later trees of a step another CTA is
// working on.  The lowest tree number with a balanced cut wins, whoever found it: the chain
// is bit-identical to one that drew the trees one after the other.
//   A UNIT is one tree (MST) or one batch of concurrent walks (Wilson).  The owner of the
// step publishes a record (HelpRec) and from then on takes its units from the same atomic
// counter as the helpers.  Everybody reports a unit in ONE 64-bit status word - epoch tag,
// outcome, tree / draw counts - so a word is either valid for the current window or taken
// for "pending"; a pending unit below the lowest success is simply computed again by the
// owner (the result cannot differ), which is why the owner never waits for anybody and no
// helper can stall a chain.  The owner alone touches the chain: when the lowest success is
// final it retires the record, draws the winning tree again if it is not the one in its
// shared memory, and goes on to the cut as if it had found the tree itself.
struct UnitResult { int status, first_ok, nbal; uint32_t value; };  // value: Boruvka rounds / walk draws
constexpr int kHelpAfter = 32;  // failed trees before a step leaves the common path (and is shared when CTAs are idle)
constexpr int kGoLong = -100;  // ProposalResult.status of propose: the proposal is to be finished by long_finish
constexpr int kLagMargin = 2;  // steps behind the average progress of the launch from which a chain keeps its CTA

// status word: epoch (24) | kind (2) | first_ok (6) | value (32)
enum { HK_FAIL = 1, HK_OK = 2, HK_ERR = 3 };
__device__ __forceinline__ unsigned long long help_pack(int epoch, int kind, int first_ok, uint32_t value) {
  return ((unsigned long long)(epoch & 0xffffff) << 40) | ((unsigned long long)kind << 38) |
         ((unsigned long long)(first_ok & 63) << 32) | (unsigned long long)value;
}
__device__ __forceinline__ int ld_acquire(const int32_t* p);
__device__ __forceinline__ void st_release(int32_t* p, int v);
__device__ __forceinline__ void st_release64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Trees tree_first .. tree_first + nw - 1 in order until one has a balanced cut; on success
// shared memory holds that tree, rooted and summed, ready for the cut choice.
template <bool kWilson>
__device__ UnitResult eval_unit(Chain& c, const Step& st, uint32_t tree_first, int nw, uint32_t prop_no,
                                double ideal, double thr) {
  const Smem& s = c.s;
  UnitResult r = {RC_OK, -1, 0, 0u};
  if constexpr (kWilson) {
    wilson_walk(c, st, tree_first, nw, prop_no, 0);
    for (int w = 0; w < nw; ++w) {
      r.status = s.wmeta[WM_ERR + w];
      r.value += (uint32_t)s.wmeta[WM_DRAWS + w];
      if (r.status) return r;
      if (st.n < 3) { r.status = RC_ERR_NO_ROOT; return r; }
      r.nbal = wilson_eval(c, st, w, ideal, thr);
      if (r.nbal > 0) { r.first_ok = w; return r; }
    }
  } else {
    const int tid = threadIdx.x, T = blockDim.x, nt = st.n - 1;
    int rounds = 0;
    r.status = boruvka_mst(c, st, tree_first, prop_no, nullptr, &rounds);
    r.value = (uint32_t)rounds;
    if (r.status) return r;
    if (st.n < 3) { r.status = RC_ERR_NO_ROOT; return r; }
    euler_subtree_pops(c, st, 0);
    if (tid == 0) s.misc[M_NBAL] = 0;
    __syncthreads();
    int cnt = 0;
    for (int t = tid; t < nt; t += T) cnt += balanced(s.cur[t], st.tot, ideal, thr) ? 1 : 0;
    cnt = warp_sum(cnt);
    if ((tid & 31) == 0 && cnt) atomicAdd(&s.misc[M_NBAL], cnt);
    __syncthreads();
    r.nbal = s.misc[M_NBAL];
    __syncthreads();
    if (r.nbal > 0) r.first_ok = 0;
  }
  return r;
}

struct SharedOut {
  int status;        // RC_OK, the error of the first decisive unit, or -1: no unit below the limit has a cut
  uint32_t tree_no;  // the winning tree
  int wsel, nbal;
  long long trees, value;  // trees drawn and rounds / draws from t_first up to and including the winner
  long long units;         // units up to and including the winner's (all of them when none won)
};

// block sum of the value fields of status words [0, cnt) (all valid for the current epoch)
__device__ long long help_value_sum(const Smem& s, const unsigned long long* ST, int cnt) {
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(&s.misc[M_BEST64]);
  if (threadIdx.x == 0) *acc = 0ull;
  __syncthreads();
  unsigned long long v = 0;
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) v += (unsigned long long)(uint32_t)__ldcg(ST + j);
  if (v) atomicAdd(acc, v);
  __syncthreads();
  const long long r = (long long)*acc;
  __syncthreads();
  return r;
}

// The owner's side: trees t_first, t_first + 1, ... of proposal prop_no, `unit` trees per unit;
// only units below `limit` exist (max_attempts), the last of them may hold fewer trees
// (trees_cap trees in all).
template <bool kWilson>
__device__ SharedOut shared_retry(Chain& c, const Step& st, uint32_t t_first, int unit, long long limit_ll,
                                  long long trees_cap, uint32_t prop_no, double ideal, double thr) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  constexpr int kInf = 0x7fffffff;
  HelpRec* R = p.help + c.chain;
  unsigned long long* ST = p.help_status + (size_t)c.chain * kHelpWin;
  const int win = p.help_win;  // units per status window (kHelpWin; smaller only in tests, to exercise the slide)
  const int limit = (int)min(limit_ll, 0x3fffffffLL);
  SharedOut out = {-1, 0u, 0, 0, 0, 0, 0};
  int32_t* bc = s.misc + M_SPARE0;  // broadcast: M_SPARE0, M_SPARE1
  int32_t* lo_ok = &s.misc[M_PICK];
  int32_t* lo_pend = &s.misc[M_CNT];
  int32_t* lo_err = &s.misc[M_SUBSET];
  if (tid == 0) {
    const int e = R->epoch + 1;  // odd from here on
    __threadfence();  // (the retirement of the last session is out before its fields change)
    R->session += 1; R->pair = st.a | (st.b << 16); R->prop_no = prop_no; R->t_first = t_first; R->unit = unit;
    R->limit = limit; R->base = 0; R->next = 0; R->won = kInf; R->want = 0;
    __threadfence();
    st_release(&R->epoch, e);
    atomicOr(p.help_bits + (c.chain >> 5), 1u << (c.chain & 31));  // what the idle CTAs poll
    atomicAdd(p.help_count, 1);
    bc[0] = e;
  }
  __syncthreads();
  int epoch = bc[0];
  __syncthreads();
  int base = 0, have = -1, have_first = -1, have_nbal = 0, winner = -1;
  long long acc_value = 0;  // of the windows slid past (all their units failed)
  auto unit_trees = [&](int u) { return (int)min((long long)unit, trees_cap - (long long)u * unit); };
  int known_dec = kInf;  // lowest deciding unit seen so far (R->won is only a hint for the helpers)
  for (long long spin = 0;; ++spin) {
    if (spin > (1LL << 24)) { winner = -1; out.status = RC_ERR_INVALID_ARG; break; }  // cannot happen; never hang a launch
    // 1. a fresh unit, if there is one worth taking
    if (tid == 0) {
      int u = -1;
      const int nx = *(volatile int32_t*)&R->next;
      if (nx < base + win && nx < limit && nx <= known_dec) {
        u = atomicAdd(&R->next, 1);
        if (!(u >= base && u < base + win && u < limit && u <= known_dec)) u = -1;
      }
      R->won = known_dec;
      R->want = min(nx * p.recruit_mult / p.units_per_recruit, 4096);  // one more CTA off its own chains per kUnitsPerRecruit units that have failed
      bc[0] = u;
      bc[1] = min(min(*(volatile int32_t*)&R->next, base + win), limit) - base;  // units of the window handed out
      *lo_ok = kInf; *lo_pend = kInf; *lo_err = kInf;
    }
    __syncthreads();
    int u = bc[0];
    const int cnt = bc[1];
    // 2. where does the window stand?  lowest success, lowest error, lowest unit not reported
    for (int j = tid; j < cnt; j += T) {
      if (base + j == u) continue;  // about to be drawn here
      const unsigned long long w = __ldcg(ST + j);
      const int kind = ((int)(w >> 40) == (epoch & 0xffffff)) ? (int)((w >> 38) & 3u) : 0;
      if (kind == HK_OK) atomicMin(lo_ok, j);
      else if (kind == HK_ERR) atomicMin(lo_err, j);
      else if (kind == 0) atomicMin(lo_pend, j);
    }
    __syncthreads();
    const int low_pend = *lo_pend, low_dec = min(*lo_ok, *lo_err);  // low_dec: the lowest unit that decides the step
    __syncthreads();
    known_dec = low_dec == kInf ? kInf : base + low_dec;
    if (low_dec != kInf && low_pend > low_dec && (u < 0 || base + low_dec < u)) { winner = base + low_dec; break; }
    if (u < 0) {
      if (low_pend != kInf && low_pend < low_dec) {
        u = base + low_pend;  // handed out but not reported (yet): draw it here, the outcome is the same
      } else if (low_dec == kInf && low_pend == kInf && cnt == min(win, limit - base)) {
        // every unit of the window failed: slide it, or give up at the limit
        acc_value += help_value_sum(s, ST, cnt);
        out.units = (long long)base + cnt;
        if (base + win >= limit) break;
        base += win;
        if (tid == 0) {
          st_release(&R->epoch, epoch + 1);  // nobody reads fields under an even epoch
          __threadfence();
          R->base = base; R->next = base; R->won = kInf;
          __threadfence();
          st_release(&R->epoch, epoch + 2);
        }
        epoch += 2;
        known_dec = kInf;
        __syncthreads();
        continue;
      } else {
        continue;  // a unit was handed out between the two reads of `next`: look again
      }
    }
    // 3. draw unit u and report it
    const UnitResult r = eval_unit<kWilson>(c, st, t_first + (uint32_t)u * (uint32_t)unit, unit_trees(u), prop_no, ideal, thr);
    have = r.first_ok >= 0 ? u : -1;
    have_first = r.first_ok;
    have_nbal = r.nbal;
    if (tid == 0) {
      const int kind = r.status ? HK_ERR : (r.first_ok >= 0 ? HK_OK : HK_FAIL);
      st_release64(ST + (u - base), help_pack(epoch, kind, r.first_ok >= 0 ? r.first_ok : 0, r.value));
      if (kind != HK_FAIL) atomicMin(&R->won, u);
    }
    __syncthreads();
  }
  // retire the record: helpers drop what they are doing for it
  if (tid == 0) {
    atomicAdd(p.help_count, -1);
    atomicAnd(p.help_bits + (c.chain >> 5), ~(1u << (c.chain & 31)));
    __threadfence();
    st_release(&R->epoch, epoch + 1);
  }
  __syncthreads();
  out.value = acc_value;
  out.trees = min((long long)out.units * unit, trees_cap);
  if (winner < 0) return out;  // status -1 (or the spin guard's error)
  // every unit below the winner failed with all its trees
  acc_value += help_value_sum(s, ST, winner - base);
  if (have != winner) {  // found by a helper (or overwritten since): the same tree again, into this CTA's shared memory
    const UnitResult r = eval_unit<kWilson>(c, st, t_first + (uint32_t)winner * (uint32_t)unit, unit_trees(winner), prop_no, ideal, thr);
    out.units = (long long)winner + 1;
    out.trees = (long long)winner * unit + (r.first_ok >= 0 ? r.first_ok + 1 : unit_trees(winner));
    out.value = acc_value + r.value;
    if (r.status) { out.status = r.status; return out; }
    if (r.first_ok < 0) { out.status = RC_ERR_INVALID_ARG; return out; }  // cannot happen: a unit is a pure function
    have_first = r.first_ok;
    have_nbal = r.nbal;
  } else {
    out.value = acc_value + (long long)(uint32_t)__ldcg(ST + (winner - base));
  }
  out.status = RC_OK;
  out.tree_no = t_first + (uint32_t)winner * (uint32_t)unit + (uint32_t)have_first;
  out.wsel = have_first;
  out.nbal = have_nbal;
  out.units = (long long)winner + 1;
  out.trees = (long long)winner * unit + have_first + 1;
  return out;
}

// Everything a step does beyond its first kHelpAfter trees, out of line (the step path of the
// common case keeps its registers): unit after unit here while every CTA has a chain of its
// own, handed to shared_retry as soon as one has not.
// (The descriptor's address must not escape - body and wrapper are separate so that the body's
// `Chain&` dissolves into registers after inlining; with the descriptor in local memory every
// array pointer is re-loaded from the stack and every shared-memory access turns generic.)
template <bool kWilson>
__device__ __forceinline__ SharedOut long_retry_body(const Params& p, Chain& c, const Step st, uint32_t t_first, int unit,
                                                     long long limit, long long trees_cap, uint32_t prop_no, double ideal, double thr) {
  const Smem& s = c.s;
  const int tid = threadIdx.x;
  SharedOut out = {-1, 0u, 0, 0, 0, 0, 0};
  for (long long u = 0; u < limit; ++u) {
    // shared from here on when CTAs are idle (the tail of a launch), or when the step is deep in the
    // heavy tail of the retry distribution (config 3: 4 steps in 10 000 need more than 64 trees, a
    // few per launch more than 1000 = tens of milliseconds): the chain would hold up the launch,
    // so CTAs leave their own chains for it, the more the longer it lasts (HelpRec.want)
    if (tid == 0) s.misc[M_SPARE0] = *(volatile int32_t*)p.idle > 0 || (u + 1) * unit + p.help_after > p.recruit_after;
    __syncthreads();
    const bool idle = s.misc[M_SPARE0] != 0;
    __syncthreads();
    if (idle) {
      SharedOut so = shared_retry<kWilson>(c, st, t_first + (uint32_t)(u * unit), unit, limit - u, trees_cap - u * unit,
                                           prop_no, ideal, thr);
      so.trees += out.trees; so.value += out.value; so.units += out.units;
      return so;
    }
    const int nw = (int)min((long long)unit, trees_cap - u * unit);
    const UnitResult r = eval_unit<kWilson>(c, st, t_first + (uint32_t)(u * unit), nw, prop_no, ideal, thr);
    out.units += 1;
    out.value += r.value;
    out.trees += r.first_ok >= 0 ? r.first_ok + 1 : nw;
    if (r.status) { out.status = r.status; return out; }
    if (r.first_ok >= 0) {
      out.status = RC_OK;
      out.tree_no = t_first + (uint32_t)(u * unit) + (uint32_t)r.first_ok;
      out.wsel = r.first_ok;
      out.nbal = r.nbal;
      return out;
    }
  }
  return out;  // status -1: no tree below max_attempts has a balanced cut
}


// Root, cut choice, sides, validity and commit for the tree in shared memory: tree number tree_no
// of proposal prop_no, with nbal balanced cuts (tree.py:377-424, 446-578, 888-966; chain.py:186-195).
template <bool kWilson>
__device__ __forceinline__ ProposalResult finish_proposal(Chain& c, const Step& st, const RcReplayStep* rs, const int n_rec,
                                                          const uint32_t tree_no, const int nbal, const int wsel, const long long attempts,
                                                          const uint32_t prop_no, const int cut_count, int32_t* ctr, ProposalResult res) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, nt = n - 1;
  const double ideal = p.cfg.pop_target;
  const double thr = __dmul_rn(ideal, p.cfg.epsilon);
  const bool warn_on = !p.cfg.allow_pair_reselection && p.cfg.warn_attempts > 0;
  int tstar = -1, root = -1, cut_child = -1;
  {
    const bool last_rec = rs && (int)tree_no == n_rec - 1;
    if (rs && !last_rec) { res.status = RC_ERR_REPLAY; return res; }
    const uint16_t* __restrict__ S = kWilson ? s.wst + (size_t)wsel * p.cap_nodes : nullptr;
    const int wroot = kWilson ? s.wmeta[WM_ROOT + wsel] : 0;
    const int rootkids = kWilson ? s.misc[M_ROOTKIDS] : 0;
    auto internal = [&](int i) {  // tree degree > 1
      if constexpr (kWilson) return i == wroot ? rootkids >= 2 : (bool)((s.whas[i >> 5] >> (i & 31)) & 1u);
      else return s.nxt[s.head[i]] != kNone;
    };
    // ---- root (tree.py:377): uniform over tree nodes of degree > 1, ascending node index
    if (last_rec) {
      const int gr = rs->root;
      root = (gr >= 0 && gr < p.g.N && is_member(s, gr)) ? local_id(s, gr) : -1;
      if (root < 0 || !internal(root)) { res.status = RC_ERR_REPLAY; return res; }
    } else {
      if (tid == 0) s.misc[M_NINT] = 0;
      __syncthreads();
      int cnt = 0;
      for (int i = tid; i < n; i += T) cnt += internal(i) ? 1 : 0;
      cnt = warp_sum(cnt);
      if ((tid & 31) == 0 && cnt) atomicAdd(&s.misc[M_NINT], cnt);
      __syncthreads();
      const int n_int = s.misc[M_NINT];
      RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_ROOT, 0, tree_no, prop_no, 0);
      root = select_kth(c, n, (int)rc_below(r.v[0], r.v[1], (uint32_t)n_int), internal);
    }
    // ---- P5
    if constexpr (kWilson) {
      // no random_weight and no node attributes on the Wilson tree: every branch of the
      // reference's cut choice degenerates to a uniform pick (tree.py:540-545), taken in
      // ascending edge id like the oracle's canonical order
      auto bal_edge = [&](int j) {
        const int ch = wilson_tree_child(s, S, j);
        return ch >= 0 && balanced(s.wsum[ch], st.tot, ideal, thr);
      };
      int jstar;
      if (rs && rs->cut_edge >= 0) {
        const int want = rs->cut_edge;
        jstar = select_kth(c, st.m, 0, [&](int j) { return c.eids[j] == want && bal_edge(j); });
      } else {
        RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_CUT, 0, tree_no, prop_no, 0);
        jstar = select_kth(c, st.m, (int)rc_below(r.v[0], r.v[1], (uint32_t)nbal), bal_edge);
      }
      if (jstar < 0) { res.status = RC_ERR_REPLAY; return res; }
      cut_child = wilson_tree_child(s, S, jstar);
    } else {
      const uint32_t* rrank = rs ? p.rp.weight_rank + (rs->first_tree + tree_no) * (long long)p.g.E : nullptr;
      auto is_bal = [&](int t) { return balanced(s.cur[t], st.tot, ideal, thr); };
      const int policy = p.cfg.cut_policy;
      if (rs && rs->cut_edge >= 0) {
        const int want = rs->cut_edge;
        tstar = select_kth(c, nt, 0, [&](int t) { return c.eids[s.te[t]] == want && is_bal(t); });
        if (tstar < 0) { res.status = RC_ERR_REPLAY; return res; }
      } else if (policy == RC_CUT_UNIFORM) {
        RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_CUT, 0, tree_no, prop_no, 0);
        tstar = select_kth(c, nt, (int)rc_below(r.v[0], r.v[1], (uint32_t)nbal), is_bal);
      } else {
        tstar = choose_cut_by_weight(c, st, policy, tree_no, prop_no, rrank, ideal, thr);
      }
    }
    if (tid == 0) {
      ctr[RC_CTR_ATTEMPTS] += (int)min(attempts + 1, 0x7fffffffLL);
      if (warn_on && attempts >= p.cfg.warn_attempts) atomicOr(p.st.d_flags + c.chain, RC_FLAG_WARNED);
    }
  }
  RC_PH(PH_CHOOSE);
  // ---- P6: the side holding the reference's root keeps the lower id (tree.py:421, :947-950)
  bool root_below;
  int child_pop;
  if constexpr (kWilson) {
    const uint16_t* __restrict__ S = s.wst + (size_t)wsel * p.cap_nodes;
    const int wroot = s.wmeta[WM_ROOT + wsel];
    // below(i): i lies in the subtree under the cut edge - climb until the cut's child or the root
    auto below = [&](int i) {
      int q = i;
      while (q != cut_child && q != wroot) q = (int)(S[q] & kNext);
      return q == cut_child;
    };
    root_below = below(root);
    child_pop = s.wsum[cut_child];
    for (int i0 = 0; i0 < st.n; i0 += T) {  // side(i): cut away from the root's side -> higher id b
      const int i = i0 + tid;
      const uint32_t bal = __ballot_sync(0xffffffffu, i < st.n && (below(i) != root_below));
      if (i < st.n && (tid & 31) == 0) s.newb[i >> 5] = bal;
    }
  } else {
    const int pa = s.P[2 * tstar], pb = s.P[2 * tstar + 1];
    const int pd = min(pa, pb), pu = max(pa, pb);
    auto below = [&](int i) { const int en = s.ent[i]; return en != 0xffff && en >= pd && en <= pu; };
    root_below = below(root);
    child_pop = s.cur[tstar];
    for (int i0 = 0; i0 < st.n; i0 += T) {
      const int i = i0 + tid;
      const uint32_t bal = __ballot_sync(0xffffffffu, i < st.n && (below(i) != root_below));
      if (i < st.n && (tid & 31) == 0) s.newb[i >> 5] = bal;
    }
  }
  const long long pop_a = root_below ? (long long)child_pop : (long long)st.tot - child_pop;
  const long long pop_b = (long long)st.tot - pop_a;
  __syncthreads();
  bool valid = true;
  if (p.cfg.pop_lower <= p.cfg.pop_upper) {  // constraints/bounds.py:25-28
    const double lo = (double)min(pop_a, pop_b), hi = (double)max(pop_a, pop_b);
    valid = p.cfg.pop_lower <= lo && hi <= p.cfg.pop_upper;
  }
  if (!valid) {
    if (tid == 0) ctr[RC_CTR_INVALID] += 1;
    return res;
  }
  commit_proposal(c, st, cut_count, pop_a, pop_b, prop_no);
  RC_PH(PH_COMMIT);
  res.accepted = 1;
  return res;
}

// The long path of a proposal: the search from tree_no on (long_retry_body: alone, or shared with
// other CTAs) and, when it has found a tree, the rest of the proposal.  Free-running only.
template <bool kWilson, bool kSpill>
__device__ __noinline__ ProposalResult long_finish(int* par) {
  const Params& p = g_params;
  Chain c;
  chain_layout<kWilson, kSpill>(c, p);
  // the arguments come through shared memory (misc[M_HC..], free while this CTA steps a chain of its
  // own): a call with a handful of register arguments costs the caller's loops registers
  const int32_t* arg = c.s.misc + M_HC;
  Step st;
  const int chain = arg[0];
  st.a = arg[1]; st.b = arg[2]; st.n = arg[3]; st.m = arg[4]; st.tot = arg[5];
  uint32_t tree_no = (uint32_t)arg[6];
  const uint32_t prop_no = (uint32_t)arg[7];
  const int cut_count = arg[8];
  __syncthreads();  // (read before thread 0 marks the helper cache empty again)
  if (threadIdx.x == 0) c.s.misc[M_HC] = -1;
  chain_bind(c, p, chain);
  c.par = *par;
  const int tid = threadIdx.x;
  int32_t* ctr = c.s.misc + M_CTR;
  ProposalResult res = {RC_OK, 0, 0, 0};
  const double ideal = p.cfg.pop_target;
  const double thr = __dmul_rn(ideal, p.cfg.epsilon);
  const long long max_attempts = p.cfg.max_attempts > 0 ? p.cfg.max_attempts : 0x7fffffffffffffffLL;
  const long long nr = p.cfg.node_repeats;  // >= 1 (Params.share is 0 otherwise)
  const bool warn_on = !p.cfg.allow_pair_reselection && p.cfg.warn_attempts > 0;
  const int unit = kWilson ? (p.walkers > 0 ? p.walkers : 1) : 1;
  const long long trees_cap = (max_attempts - (long long)tree_no * nr + nr - 1) / nr;  // trees from tree_no on that may still be examined
  const SharedOut so = long_retry_body<kWilson>(p, c, st, tree_no, unit, (trees_cap + unit - 1) / unit, trees_cap, prop_no, ideal, thr);
  if (tid == 0) {
    ctr[RC_CTR_TREES] += (int)so.trees;
    ctr[kWilson ? RC_CTR_WALK : RC_CTR_ROUNDS] += (int)so.value;
    if (kWilson) ctr[RC_CTR_ROUNDS] += (int)so.units;
  }
  res.trees = (int)(tree_no + (uint32_t)so.trees);
  if (so.status > 0) {
    res.status = so.status;
  } else if (so.status < 0) {  // no tree below max_attempts has a balanced cut (tree.py:718; the long path excludes pair reselection)
    if (tid == 0) {
      ctr[RC_CTR_ATTEMPTS] += (int)min(max_attempts, 0x7fffffffLL);
      if (warn_on && max_attempts >= p.cfg.warn_attempts) atomicOr(p.st.d_flags + c.chain, RC_FLAG_WARNED);
    }
    res.status = RC_ERR_MAX_ATTEMPTS;
  } else {
    tree_no = so.tree_no;
    res.trees = (int)tree_no + 1;
    res.nbal = so.nbal;
    res = finish_proposal<kWilson>(c, st, nullptr, 0, tree_no, so.nbal, so.wsel, (long long)tree_no * nr, prop_no, cut_count, ctr, res);
  }
  *par = c.par;  // hand the scan parity back
  return res;
}

// One recom proposal + validity + commit (tree_proposals.py:36-146, chain.py:186-195).
template <bool kWilson, bool kShare, bool kSpill>
__device__ ProposalResult propose(Chain& c, uint32_t prop_no, const RcReplayStep* rs,
                                  int32_t* ctr /* smem counters */) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  ProposalResult res = {RC_OK, 0, 0, 0};
  const int cut_count = __ldcg(p.st.d_cut_count + c.chain);
  if (cut_count <= 0) { res.status = RC_ERR_NO_CUT_EDGES; return res; }
  Step st;
  const int k = p.cfg.n_districts;
  const bool resel = p.cfg.allow_pair_reselection != 0 && !rs;
  uint32_t* bad = p.badpairs ? p.badpairs + (size_t)blockIdx.x * 2048 : nullptr;
  int n_bad = 0;
  if (resel) {
    for (int i = tid; i < 2048; i += T) bad[i] = 0;
    __syncthreads();
  }
  const double ideal = p.cfg.pop_target;
  const double thr = __dmul_rn(ideal, p.cfg.epsilon);
  const long long max_attempts = p.cfg.max_attempts > 0 ? p.cfg.max_attempts : 0x7fffffffffffffffLL;
  const long long nr = p.cfg.node_repeats;
  const bool warn_on = !p.cfg.allow_pair_reselection && p.cfg.warn_attempts > 0;
  const int n_rec = rs ? rs->n_trees : 0;
  const uint32_t t0 = (rs && nr == 0 && n_rec > 0) ? (uint32_t)(n_rec - 1) : 0u;
  const int r0 = 0;  // tour root: local node 0 (any node will do)
  int tstar = -1, root = -1;
  int wsel = 0, cut_child = -1;  // Wilson: winning walker, child endpoint of the chosen cut
  uint32_t tree_no = t0;
  long long attempts = 0;
  int nbal = 0;
  for (uint32_t retry = 0;; ++retry) {  // tree_proposals.py:103-138
  if (resel && (n_bad >= k * (k - 1) / 2 || retry >= 64u * (uint32_t)(k * k))) { res.status = RC_ERR_ALL_PAIRS_BAD; return res; }
  // ---- P0
  int e;
  if (rs) {
    e = rs->pair_edge;
    if (e < 0 || e >= p.g.E || !((__ldcg(c.mask + (e >> 5)) >> (e & 31)) & 1)) { res.status = RC_ERR_REPLAY; return res; }
  } else {
    RC_PH(PH_OTHER);
    e = pick_cut_edge(c, cut_count, prop_no, retry);
    RC_PH(PH_PICK);
    if (e < 0) { res.status = RC_ERR_NO_CUT_EDGES; return res; }
  }
  {
    const int la = __ldcg(c.assign + p.g.edge_uv[2 * e]), lb = __ldcg(c.assign + p.g.edge_uv[2 * e + 1]);
    st.a = min(la, lb); st.b = max(la, lb);  // tree_proposals.py:113
  }
  if (resel && ((bad[(st.a * 256 + st.b) >> 5] >> ((st.a * 256 + st.b) & 31)) & 1)) continue;
  // ---- P1
  int d_sub = 0;
  res.status = gather_pair<kWilson>(c, st, &d_sub);
  RC_PH(PH_GATHER);
  if (tid == 0) { ctr[RC_CTR_NODES] += st.n; ctr[RC_CTR_SLOTS] += d_sub; }
  if (res.status) return res;
  const int n = st.n, nt = n - 1;
  bool reselect = false;

  uint32_t batch_lo = t0, batch_hi = t0;  // Wilson: trees [batch_lo, batch_hi) have been walked
  bool go_long = false;  // shared-retry build: the search leaves the common path (below the loop)
  tree_no = t0;
  attempts = 0;
  nbal = 0;
  for (;; ++tree_no) {
    // tree.py:679-700 attempt accounting (see oracle/recom_oracle.c rco_bipartition)
    attempts = nr > 0 ? (long long)tree_no * nr : (tree_no == t0 ? 0 : max_attempts);
    if (attempts >= max_attempts) {
      if (tid == 0) {
        ctr[RC_CTR_ATTEMPTS] += (int)min(max_attempts, 0x7fffffffLL);
        if (warn_on && max_attempts >= p.cfg.warn_attempts) atomicOr(p.st.d_flags + c.chain, RC_FLAG_WARNED);
      }
      if (resel) { reselect = true; break; }  // ReselectException, tree.py:712-716
      res.status = RC_ERR_MAX_ATTEMPTS;
      return res;
    }
    if (rs && (int)tree_no >= n_rec) { res.status = RC_ERR_REPLAY; return res; }
    // CTAs without a chain to step?  Then the rest of this search is shared with them
    // (shared_retry above); checked once per tree / batch, at a point where nothing is in flight
    bool go_shared = false;
    if constexpr (kShare) {
      // a step that has already failed kHelpAfter trees leaves the common path for good (long_retry):
      // the common step (a handful of trees) is not worth a helper's gather
      go_shared = !rs && !resel && tree_no >= t0 + (uint32_t)p.help_after && (!kWilson || tree_no >= batch_hi);
    }
    if (kShare && go_shared) { go_long = true; break; }
    if constexpr (kWilson) {
      // ---- P2b: the next batch of trees, one per lane of warp 0; examined in order
      if (tree_no >= batch_hi) {
        long long nw = p.walkers > 0 ? p.walkers : 1;
        if (rs) nw = 1;  // a recorded tree is walked by one thread (wilson_walk_recorded)
        if (nr > 0) nw = min(nw, (max_attempts - attempts + nr - 1) / nr);  // trees beyond max_attempts are never examined
        else nw = 1;
        wilson_walk(c, st, tree_no, (int)nw, prop_no, rs ? rs->first_tree + tree_no : 0);
        batch_lo = tree_no;
        batch_hi = tree_no + (uint32_t)nw;
        if (tid == 0) ctr[RC_CTR_ROUNDS] += 1;  // Wilson: batches of concurrent walks
        RC_PH(PH_TREE);
      }
      wsel = (int)(tree_no - batch_lo);
      res.status = s.wmeta[WM_ERR + wsel];
      if (tid == 0) { ctr[RC_CTR_TREES] += 1; ctr[RC_CTR_WALK] += s.wmeta[WM_DRAWS + wsel]; }
      res.trees = (int)tree_no + 1;
      if (res.status) return res;
      if (n < 3) { res.status = RC_ERR_NO_ROOT; return res; }  // tree.py:377: no node of degree > 1
      nbal = wilson_eval(c, st, wsel, ideal, thr);
      RC_PH(PH_EULER);
    } else {
      const uint32_t* rrank = rs ? p.rp.weight_rank + (rs->first_tree + tree_no) * (long long)p.g.E : nullptr;
      // ---- P2a
      int rounds = 0;
      res.status = boruvka_mst(c, st, tree_no, prop_no, rrank, &rounds);
      RC_PH(PH_TREE);
      if (tid == 0) { ctr[RC_CTR_TREES] += 1; ctr[RC_CTR_ROUNDS] += rounds; }
      res.trees = (int)tree_no + 1;
      if (res.status) return res;
      // tree.py:377: the root is drawn among tree nodes of degree > 1 (none: IndexError);
      // only the two-node tree has none
      if (n < 3) { res.status = RC_ERR_NO_ROOT; return res; }
      // ---- P3
      euler_subtree_pops(c, st, r0);
      RC_PH(PH_EULER);
      // ---- P4: balanced tree edges (cur[t] = population below edge t)
      if (tid == 0) s.misc[M_NBAL] = 0;
      __syncthreads();
      int cnt = 0;
      for (int t = tid; t < nt; t += T) cnt += balanced(s.cur[t], st.tot, ideal, thr) ? 1 : 0;
      cnt = warp_sum(cnt);
      if ((tid & 31) == 0 && cnt) atomicAdd(&s.misc[M_NBAL], cnt);
      __syncthreads();
      nbal = s.misc[M_NBAL];
    }
    RC_PH(PH_BALANCE);
    res.nbal = nbal;
    if (nbal > 0) break;
  }
  if constexpr (kShare) {
    // Everything beyond the first help_after trees - the rest of the search AND what follows it, down
    // to the commit - happens out of line, in tail position: nothing of this function is live
    // across the call.  (With the search alone out of line and the step finished here, the values
    // kept for afterwards cost the Boruvka compaction of the common case its staging registers:
    // spills, -8 % at config 2 with nothing shared at all.)
    if (go_long) {
      __syncthreads();
      if (tid == 0) {
        int32_t* arg = s.misc + M_HC;
        arg[0] = c.chain; arg[1] = st.a; arg[2] = st.b; arg[3] = st.n; arg[4] = st.m; arg[5] = st.tot;
        arg[6] = (int)tree_no; arg[7] = (int)prop_no; arg[8] = cut_count;
      }
      __syncthreads();
      res.status = kGoLong;  // step_body makes the call
      return res;
    }
  }
  if (reselect) {  // tree_proposals.py:133-136
    if (tid == 0) { bad[(st.a * 256 + st.b) >> 5] |= 1u << ((st.a * 256 + st.b) & 31); ctr[RC_CTR_RESELECT] += 1; }
    ++n_bad;
    __syncthreads();
    continue;
  }
  break;
  }  // pair loop
  return finish_proposal<kWilson>(c, st, rs, n_rec, tree_no, nbal, wsel, attempts, prop_no, cut_count, ctr, res);
}

// ---- reversible ReCom (proposals/tree_proposals.py:149-267) ------------------------------
// An ordered pair of districts uniformly out of k^2; no shared edge -> self-loop.  One
// uniform spanning tree of the merged pair (no retries), all balanced cuts; more than M ->
// ReversibilityError, none -> self-loop.  The proposal (uniform cut, root side takes the
// label of the picked edge's first endpoint) is taken with probability
// #cuts / (M * seam length of the proposal).  Every call counts as a chain step.
__device__ ProposalResult propose_reversible(Chain& c, uint32_t prop_no, int32_t* ctr) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  ProposalResult res = {RC_OK, 1, 0, 0};  // accepted = 1: self-loops count as steps
  const int k = p.cfg.n_districts;
  const int M = p.cfg.rev_M > 0 ? p.cfg.rev_M : 1;
  const RcPhilox rp = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_PAIR, 0, 0, prop_no, 0);
  const int pr = (int)rc_below(rp.v[0], rp.v[1], (uint32_t)(k * k));
  const int d_out = pr / k, d_in = pr % k;
  if (d_out == d_in) return res;  // tree_proposals.py:227
  // the cut edges between the two districts, ascending edge id
  auto joins = [&](int e) {
    const int la = __ldcg(c.assign + p.g.edge_uv[2 * e]), lb = __ldcg(c.assign + p.g.edge_uv[2 * e + 1]);
    return (la == d_out && lb == d_in) || (la == d_in && lb == d_out);
  };
  const int words = (p.g.E + 31) >> 5;
  const int per = odd_chunk(words);
  const int lo = min(tid * per, words), hi = min(lo + per, words);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) {
    uint32_t x = __ldcg(c.mask + i);
    while (x) { const int e = i * 32 + (__ffs(x) - 1); x &= x - 1; cnt += joins(e) ? 1 : 0; }
  }
  if (tid == 0) s.misc[M_PICK] = -1;
  int n_pair;
  const int excl = block_excl_scan(cnt, &n_pair, s.misc, c.par);
  if (n_pair == 0) { __syncthreads(); return res; }  // not adjacent: self-loop
  const int kth = (int)rc_below(rp.v[2], rp.v[3], (uint32_t)n_pair);
  if (kth >= excl && kth < excl + cnt) {
    int q = kth - excl;
    for (int i = lo; i < hi && q >= 0; ++i) {
      uint32_t x = __ldcg(c.mask + i);
      while (x) {
        const int e = i * 32 + (__ffs(x) - 1);
        x &= x - 1;
        if (joins(e) && q-- == 0) { s.misc[M_PICK] = e; break; }
      }
    }
  }
  __syncthreads();
  const int e = s.misc[M_PICK];
  __syncthreads();
  const int lab_root = __ldcg(c.assign + p.g.edge_uv[2 * e]);       // parts_to_merge[0]
  const int lab_other = __ldcg(c.assign + p.g.edge_uv[2 * e + 1]);  // parts_to_merge[1]
  Step st;
  st.a = min(lab_root, lab_other);
  st.b = max(lab_root, lab_other);
  int d_sub = 0;
  res.status = gather_pair<true>(c, st, &d_sub);
  if (tid == 0) { ctr[RC_CTR_NODES] += st.n; ctr[RC_CTR_SLOTS] += d_sub; }
  if (res.status) return res;
  const int n = st.n;
  if (n < 3) { res.status = RC_ERR_NO_ROOT; return res; }
  wilson_walk(c, st, 0, 1, prop_no, 0);
  res.status = s.wmeta[WM_ERR];
  if (tid == 0) { ctr[RC_CTR_TREES] += 1; ctr[RC_CTR_ATTEMPTS] += 1; ctr[RC_CTR_WALK] += s.wmeta[WM_DRAWS]; }
  res.trees = 1;
  if (res.status) return res;
  const double ideal = p.cfg.pop_target;
  const double thr = __dmul_rn(ideal, p.cfg.epsilon);
  const int nbal = wilson_eval(c, st, 0, ideal, thr);
  const uint16_t* __restrict__ S = s.wst;
  const int wroot = s.wmeta[WM_ROOT];
  const int rootkids = s.misc[M_ROOTKIDS];
  auto internal = [&](int i) { return i == wroot ? rootkids >= 2 : (bool)((s.whas[i >> 5] >> (i & 31)) & 1u); };
  if (tid == 0) s.misc[M_NINT] = 0;
  __syncthreads();
  {
    int ci = 0;
    for (int i = tid; i < n; i += T) ci += internal(i) ? 1 : 0;
    ci = warp_sum(ci);
    if ((tid & 31) == 0 && ci) atomicAdd(&s.misc[M_NINT], ci);
  }
  __syncthreads();
  const int n_int = s.misc[M_NINT];
  res.nbal = nbal;
  if (nbal > M) { res.status = RC_ERR_REVERSIBILITY; return res; }  // tree_proposals.py:205-209
  if (nbal == 0) return res;                                           // self-loop: no balance edge
  const RcPhilox rr = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_ROOT, 0, 0, prop_no, 0);
  const int root = select_kth(c, n, (int)rc_below(rr.v[0], rr.v[1], (uint32_t)n_int), internal);
  const RcPhilox rc_ = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_CUT, 0, 0, prop_no, 0);
  const int jstar = select_kth(c, st.m, (int)rc_below(rc_.v[0], rc_.v[1], (uint32_t)nbal), [&](int j) {
    const int ch = wilson_tree_child(s, S, j);
    return ch >= 0 && balanced(s.wsum[ch], st.tot, ideal, thr);
  });
  const int cut_child = wilson_tree_child(s, S, jstar);
  auto below = [&](int i) {
    int q = i;
    while (q != cut_child && q != wroot) q = (int)(S[q] & kNext);
    return q == cut_child;
  };
  const bool root_below = below(root);
  const int child_pop = s.wsum[cut_child];
  const long long pop_root = root_below ? (long long)child_pop : (long long)st.tot - child_pop;
  const long long pop_other = (long long)st.tot - pop_root;
  // new side bitmap: bit = node carries st.b afterwards
  const bool root_is_b = lab_root == st.b;
  auto new_is_b = [&](int i) { return (below(i) != root_below) != root_is_b; };  // cut-away side takes lab_other
  if (tid == 0) { s.misc[M_DELTA] = 0; s.misc[M_CNT] = 0; }
  for (int i0 = 0; i0 < n; i0 += T) {
    const int i = i0 + tid;
    const uint32_t bal = __ballot_sync(0xffffffffu, i < n && new_is_b(i));
    if (i < n && (tid & 31) == 0) s.newb[i >> 5] = bal;
  }
  __syncthreads();
  int seam = 0, delta = 0;
  for (int j = tid; j < st.m; j += T) {
    const uint32_t uv = s.euv[j];
    const int u = (int)(uv & 0xffffu), v = (int)(uv >> 16);
    const int was = ((s.oldb[u >> 5] >> (u & 31)) ^ (s.oldb[v >> 5] >> (v & 31))) & 1u;
    const int now = ((s.newb[u >> 5] >> (u & 31)) ^ (s.newb[v >> 5] >> (v & 31))) & 1u;
    seam += now;
    delta += now - was;
  }
  seam = warp_sum(seam);
  delta = warp_sum(delta);
  if ((tid & 31) == 0) { if (seam) atomicAdd(&s.misc[M_CNT], seam); if (delta) atomicAdd(&s.misc[M_DELTA], delta); }
  __syncthreads();
  const int seam_len = s.misc[M_CNT], dcut = s.misc[M_DELTA];
  const double prob = __ddiv_rn((double)nbal, (double)((long long)M * seam_len));  // tree_proposals.py:256
  if (prob > 1.0) { res.status = RC_ERR_REVERSIBILITY; return res; }
  const RcPhilox ra = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_ACCEPT, 0, 0, prop_no, 0);
  if (!(rc_unit_double(ra.v[0], ra.v[1]) < prob)) return res;  // self-loop
  if (p.cfg.pop_lower <= p.cfg.pop_upper) {  // the chain's Bounds (constraints/bounds.py:25-28)
    const double lo2 = (double)min(pop_root, pop_other), hi2 = (double)max(pop_root, pop_other);
    if (!(p.cfg.pop_lower <= lo2 && hi2 <= p.cfg.pop_upper)) {
      if (tid == 0) ctr[RC_CTR_INVALID] += 1;
      res.accepted = 0;
      return res;
    }
  }
  for (int i = tid; i < n; i += T) {
    const bool nb = (s.newb[i >> 5] >> (i & 31)) & 1u, ob = (s.oldb[i >> 5] >> (i & 31)) & 1u;
    if (nb != ob) c.assign[c.nodes[i]] = (rc_label_t)(nb ? st.b : st.a);
  }
  for (int j = tid; j < st.m; j += T) {
    const uint32_t uv = s.euv[j];
    const int u = (int)(uv & 0xffffu), v = (int)(uv >> 16);
    const int was = ((s.oldb[u >> 5] >> (u & 31)) ^ (s.oldb[v >> 5] >> (v & 31))) & 1u;
    const int now = ((s.newb[u >> 5] >> (u & 31)) ^ (s.newb[v >> 5] >> (v & 31))) & 1u;
    if (now != was) { const int eid = c.eids[j]; atomicXor(c.mask + (eid >> 5), 1u << (eid & 31)); }
  }
  if (tid == 0) {
    c.tally[lab_root] = pop_root;
    c.tally[lab_other] = pop_other;
    p.st.d_cut_count[c.chain] = __ldcg(p.st.d_cut_count + c.chain) + dcut;
    if (p.st.d_last_pair) { p.st.d_last_pair[2 * c.chain] = st.a; p.st.d_last_pair[2 * c.chain + 1] = st.b; }
  }
  __syncthreads();
  return res;
}

// =========================================================================================
__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The helper's side of a shared retry (see shared_retry): one unit of somebody else's step.
// `hc` remembers the pair gathered last, so that staying with a chain costs one gather.
// (kept in shared memory, misc[M_HC..]: registers held across the step loop cost the step path)
struct HelpCache {
  int32_t* m;
  __device__ __forceinline__ int chain() const { return m[0]; }
  __device__ __forceinline__ int session() const { return m[1]; }
  __device__ __forceinline__ Step st() const { Step s; s.a = m[2]; s.b = m[3]; s.n = m[4]; s.m = m[5]; s.tot = m[6]; return s; }
  __device__ __forceinline__ void set(int chain, int session, const Step& s) {  // thread 0, between barriers
    m[0] = chain; m[1] = session; m[2] = s.a; m[3] = s.b; m[4] = s.n; m[5] = s.m; m[6] = s.tot;
  }
};

// a chain that wants help: the first bit of the help bitmap at or after a CTA-specific start
// (spreads the helpers), or -1
__device__ __forceinline__ int help_pick(const Params& p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* const misc = reinterpret_cast<int32_t*>(smem_raw);  // Smem.misc is the first array of every layout
  const int tid = threadIdx.x, T = blockDim.x;
  const int n_chains = p.cfg.n_chains;
  constexpr int kInf = 0x7fffffff;
  if (tid == 0) misc[M_PICK] = kInf;
  __syncthreads();
  const int nw = (n_chains + 31) >> 5;
  const int start = (int)(blockIdx.x % (unsigned)n_chains), sw = start >> 5, sb = start & 31;
  for (int i = tid; i <= nw; i += T) {  // the start word twice: its high bits first, its low bits last
    int w = sw + i;
    if (w >= nw) w -= nw;
    uint32_t bits = __ldcg(p.help_bits + w);
    if (i == 0) bits &= ~0u << sb;
    if (i == nw) bits &= (1u << sb) - 1u;
    if (bits) { atomicMin(&misc[M_PICK], (i << 5) | (__ffs((int)bits) - 1)); break; }  // (words after the start word) << 5 | bit
  }
  __syncthreads();
  const int pick = misc[M_PICK];
  __syncthreads();
  if (pick == kInf) return -1;
  int wsel = sw + (pick >> 5);
  if (wsel >= nw) wsel -= nw;
  return (wsel << 5) | (pick & 31);
}

template <bool kWilson>
__device__ __forceinline__ int help_body(const Params& p, Chain& c, const int ch, const bool busy) {
  const Smem& s = c.s;
  HelpCache hc;
  hc.m = s.misc + M_HC;
  const int tid = threadIdx.x;
  int32_t* bc = s.misc + M_SPARE0;
  HelpRec* R = p.help + ch;
  // The record is read under ONE epoch value: every change of it (publish, window slide,
  // retirement) bumps the epoch, the owner fences between the bump and the next change of a
  // field, so fields read between two equal reads of the epoch belong to that epoch.  All that
  // follows - the unit taken, the status word written - is tagged with it and dropped as soon
  // as the epoch has moved on.
  if (tid == 0) {
    const int e = ld_acquire(&R->epoch);
    s.misc[M_CNT] = __ldcg(&R->pair);
    s.misc[M_SUBSET] = (int)__ldcg(&R->prop_no);
    s.misc[M_DELTA] = (int)__ldcg(&R->t_first);
    s.misc[M_NINT] = __ldcg(&R->unit);
    s.misc[M_NLIVE] = __ldcg(&R->limit);
    s.misc[M_NBAL] = __ldcg(&R->base);
    bc[1] = __ldcg(&R->session);
    // a CTA that has chains to step joins only when the step has failed often enough to want it
    // (stateless admission: the first `want` CTAs counted from a record-specific offset)
    const bool admitted = !busy || (int)((blockIdx.x + (unsigned)ch * 7919u) % gridDim.x) < __ldcg(&R->want);
    __threadfence();
    bc[0] = (ld_acquire(&R->epoch) == e && (e & 1) && admitted) ? e : 0;
  }
  __syncthreads();
  const int epoch = bc[0], session = bc[1], pair = s.misc[M_CNT], unit = s.misc[M_NINT];
  const int limit = s.misc[M_NLIVE], base = s.misc[M_NBAL];
  const uint32_t prop_no = (uint32_t)s.misc[M_SUBSET], t_first = (uint32_t)s.misc[M_DELTA];
  __syncthreads();
  if (epoch == 0) return 0;
  chain_bind(c, p, ch);
  if (hc.chain() != ch || hc.session() != session) {
    // the chain's plan is stable while its record is live (the owner retires it before the commit)
    __syncthreads();
    if (tid == 0) hc.m[0] = -1;
    Step st;
    st.a = pair & 0xffff; st.b = pair >> 16;
    int d_sub = 0;
    const int rc = gather_pair<kWilson>(c, st, &d_sub);
    __threadfence();
    if (tid == 0) bc[0] = ld_acquire(&R->epoch);
    __syncthreads();
    const int still = bc[0];
    __syncthreads();
    if (rc != RC_OK || still != epoch) return 0;  // retired (or slid) meanwhile: the plan may have moved under the gather
    if (tid == 0) hc.set(ch, session, st);
    __syncthreads();
  }
  const Step hst = hc.st();
  // take a unit
  if (tid == 0) {
    int u = -1;
    const int nx = *(volatile int32_t*)&R->next, won = *(volatile int32_t*)&R->won;
    if (ld_acquire(&R->epoch) == epoch && nx < base + p.help_win && nx < limit && nx <= won) {
      u = atomicAdd(&R->next, 1);
      if (!(u >= base && u < base + p.help_win && u < limit)) u = -1;
    }
    bc[0] = u;
  }
  __syncthreads();
  const int u = bc[0];
  __syncthreads();
  if (u < 0) {
    if (!busy) __nanosleep(1000);
    return 0;
  }
  const double ideal = p.cfg.pop_target;
  const double thr = __dmul_rn(ideal, p.cfg.epsilon);
  const UnitResult r = eval_unit<kWilson>(c, hst, t_first + (uint32_t)u * (uint32_t)unit, unit, prop_no, ideal, thr);
  if (tid == 0 && ld_acquire(&R->epoch) == epoch) {
    const int kind = r.status ? HK_ERR : (r.first_ok >= 0 ? HK_OK : HK_FAIL);
    st_release64(p.help_status + (size_t)ch * kHelpWin + (u - base), help_pack(epoch, kind, r.first_ok >= 0 ? r.first_ok : 0, r.value));
    if (kind != HK_FAIL) atomicMin(&R->won, u);
  }
  __syncthreads();
  return 1;
}

// returns 1 when a unit was drawn; busy: the CTA has chains to step (see help_body)
template <bool kWilson, bool kSpill>
__device__ __noinline__ int help_once(int* par, const int busy) {
  const Params& p = g_params;
  const int ch = help_pick(p);
  if (ch < 0) {
    if (!busy) __nanosleep(2000);
    return 0;
  }
  Chain c;
  chain_layout<kWilson, kSpill>(c, p);
  c.par = *par;
  const int did = help_body<kWilson>(p, c, ch, busy != 0);
  *par = c.par;  // hand the scan parity back
  return did;
}

// One accepted step of `chain` (chain.py:185-195) and its bookkeeping; returns the chain's status.
template <bool kWilson, bool kSpill, bool kShare>
__device__ __forceinline__ int step_body(const Params& p, Chain& c, const int chain, const unsigned long long t,
                                         const bool replay, const int streak_cap, int32_t* ctr) {
  const int tid = threadIdx.x;
  c.chain = chain;
  c.gchain = (uint32_t)(chain + p.cfg.chain_offset);
  c.assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
  c.mask = p.st.d_cut_mask + (size_t)chain * p.st.mask_stride;
  c.tally = p.st.d_tally + (size_t)chain * p.cfg.n_districts;
  int status = __ldcg(p.st.d_status + chain);
  if (status == RC_OK) {
    // one accepted step (chain.py:185-195): propose until valid, accept always
    uint32_t prop_no = __ldcg(p.st.d_proposal_no + chain);
    RC_PH(PH_STATE);
    const RcReplayStep* rs = replay ? p.rp.steps + t : nullptr;
    ProposalResult r = {RC_OK, 0, 0, 0};
    for (int streak = 0; streak < streak_cap; ++streak) {
      if constexpr (kWilson) {
        if (p.cfg.proposal_kind == RC_PROPOSAL_REVERSIBLE && !replay) r = propose_reversible(c, prop_no, ctr);
        else r = propose<kWilson, kShare, kSpill>(c, replay ? (uint32_t)t : prop_no, rs, ctr);
      } else {
        r = propose<kWilson, kShare, kSpill>(c, replay ? (uint32_t)t : prop_no, rs, ctr);
      }
      if constexpr (kShare) {
        if (r.status == kGoLong) {  // the proposal goes on out of line (its arguments are in shared memory)
          int par = c.par;
          r = long_finish<kWilson, kSpill>(&par);
          c.par = par;
        }
      }
      ++prop_no;
      __syncthreads();
      if (r.status || r.accepted || replay) break;
    }
    status = r.status ? r.status : (r.accepted ? RC_OK : RC_ERR_INVALID_STREAK);
    if (tid == 0) {
      p.st.d_proposal_no[chain] = prop_no;
      if (status == RC_OK) p.st.d_step_no[chain] = __ldcg(p.st.d_step_no + chain) + 1;
      else p.st.d_status[chain] = status;
    }
    if (tid < RC_N_COUNTERS) {
      int64_t* g = p.st.d_counters + (size_t)chain * RC_N_COUNTERS + tid;
      *g = __ldcg(g) + ctr[tid];
      ctr[tid] = 0;
    }
    if (replay) {
      if (p.rp.info_trace && tid == 0) {
        int32_t* inf = p.rp.info_trace + 4 * t;
        inf[0] = r.trees; inf[1] = r.nbal; inf[2] = p.st.d_cut_count[chain]; inf[3] = status;
      }
      if (p.rp.assign_trace)
        for (int i = tid; i < p.g.N; i += blockDim.x) p.rp.assign_trace[(size_t)t * p.g.N + i] = c.assign[i];
    }
  }
  return status;
}

// Free-running mode is a persistent grid over a ring of READY chains (see the loop below):
// every SM stays busy whatever the number of chains and however unevenly the retries fall,
// and a slow step delays only its own chain.  Replay mode (TEST HOOK) is one CTA walking one
// chain through the record.
template <bool kWilson, bool kSpill, bool kShare>
__global__ void __launch_bounds__(kWilson ? 256 : 512, kWilson ? 4 : 2) recom_step_kernel(const __grid_constant__ Params p) {
  Chain c;
  chain_layout<kWilson, kSpill>(c, p);
  const int tid = threadIdx.x;
  c.par = 0;
#ifdef RC_TIMERS
  if (tid < 30) c.s.misc[96 + tid] = 0;  // misc[96..127]: per-CTA phase cycles (+ last stamp), flushed at exit
  if (tid == 0) reinterpret_cast<unsigned long long*>(c.s.misc + 96)[15] = (unsigned long long)clock64();
#endif
  int32_t* ctr = c.s.misc + M_CTR;  // [RC_N_COUNTERS]
  if (tid < RC_N_COUNTERS) ctr[tid] = 0;
  if (kWilson && tid == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int slot = atomicAdd(p.sm_slots + (smid & 1023u), 1);
    const int n_warps = (int)blockDim.x >> 5;
    c.s.misc[M_WALKER] = slot % (n_warps < 4 ? n_warps : 4);
  }
  __syncthreads();
  const bool replay = p.rp.n_steps > 0;
  const long long n_chains = p.cfg.n_chains;
  const unsigned long long total =
      replay ? (unsigned long long)p.rp.n_steps : (unsigned long long)n_chains * (unsigned long long)p.n_steps;
  const int streak_cap = p.cfg.max_invalid_streak > 0 ? p.cfg.max_invalid_streak : 1000;
  // Work queue of READY chains (free-running mode).  Pop index h < n_chains takes chain h;
  // later pops take ring slot h - n_chains, which the CTA that finished a step of some chain
  // fills (release) once that chain's state is in L2.  A chain is in the ring at most once,
  // so a slow step (many retries) delays only its own chain: no CTA ever waits while another
  // chain is ready.
  int32_t* tslot = &c.s.misc[M_TICKET];
  unsigned long long rt = 0;  // replay: steps are taken in order
  const unsigned long long cap = (unsigned long long)p.ring_cap;
  int flip = 0;
  // A CTA whose ticket has no ready chain yet (the tail of a launch, or fewer chains than CTAs)
  // does not sleep on its ring slot: between two looks at the slot it draws one unit of a
  // step some other CTA is retrying (help_once).  A ticket beyond the last step turns the
  // CTA into a helper for the rest of the launch, i.e. until every chain-step is done.
  const bool share = kShare && p.share != 0 && !replay;
  HelpCache hc;
  hc.m = c.s.misc + M_HC;
  unsigned long long* const ticket_p = reinterpret_cast<unsigned long long*>(c.s.misc + M_TICKET64);  // thread 0's
  if (tid == 0) { hc.m[0] = -1; hc.m[1] = 0; *ticket_p = ~0ull; c.s.misc[M_KEEP] = -1; c.s.misc[M_HELPMODE] = 0; }  // ~0: no ticket
  __syncthreads();
  bool idle = false;
  for (;;) {
    unsigned long long t;
    int chain;
    if (replay) {
      t = rt++;
      chain = p.rp.chain;
    } else {
      if (tid == 0) {
        int ch = c.s.misc[M_KEEP];  // a chain behind the field goes on without queueing (below)
        if (ch >= 0) {
          c.s.misc[M_KEEP] = -1;
        } else if (kShare && c.s.misc[M_HELPMODE]) {
          ch = -4;  // a long step somewhere wants CTAs (below)
        } else {
          unsigned long long ticket = *ticket_p;
          if (ticket == ~0ull) ticket = atomicAdd(p.work_counter, 1ull);
          ch = -1;  // -1: nothing to step right now, -2: the launch is over
          if (ticket < total) {
            if (ticket < (unsigned long long)n_chains) {
              ch = (int)ticket;
            } else {
              // (kept chains are not pushed: the last tickets wait for pushes that never come,
              // until the count of finished steps says that the launch is over)
              int32_t* slot = p.ring + (ticket - (unsigned long long)n_chains) % cap;
              int v = ld_acquire(slot);
              if (!share) {
                for (int spin = 1; v == 0; ++spin) {
                  __nanosleep(32);
                  v = ld_acquire(slot);
                  if (v == 0 && (spin & 15) == 0 && *(volatile unsigned long long*)p.steps_done >= total) { ch = -2; break; }
                }
              } else if (v == 0 && *(volatile unsigned long long*)p.steps_done >= total) {
                ch = -2;
              }
              if (v != 0) {
                st_release(slot, 0);  // the slot may be refilled cap pushes later
                ch = v - 1;
              }
            }
          } else if (!share || *(volatile unsigned long long*)p.steps_done >= total) {
            ch = -2;
          }
          *ticket_p = ch >= 0 ? ~0ull : ticket;
          if (share && (ch == -1) != idle) atomicAdd(p.idle, ch == -1 ? 1 : -1);
          if (share && ch == -1 && *(volatile int32_t*)p.help_count <= 0) ch = -3;  // nobody to help right now
        }
        tslot[flip] = ch;
      }
      __syncthreads();
      chain = tslot[flip];
      flip ^= 1;
      RC_PH(PH_TICKET);
      if (chain == -3) {  // (one word polled by one thread: the CTAs at work keep the L2 to themselves)
        idle = true;
        __nanosleep(2000);
        continue;
      }
      if (chain == -1 || chain == -4) {
        if (chain == -1) idle = true;
        if constexpr (kShare) {
          int par = c.par;
#ifndef RC_X_NOHELP
          const int did = help_once<kWilson, kSpill>(&par, chain == -4);  // out of line, with a descriptor of its own
#else
          const int did = 0;
#endif
          c.par = par;
          if (chain == -4 && !did && tid == 0) c.s.misc[M_HELPMODE] = 0;  // back to the ring of ready chains
        }
        continue;
      }
      idle = false;
      if (kShare && tid == 0) hc.m[0] = -1;  // this CTA's shared memory is about to hold another pair
      t = chain < 0 ? total : 0;
    }
    if (t >= total) break;
    const int status = step_body<kWilson, kSpill, kShare>(p, c, chain, t, replay, streak_cap, ctr);
    __threadfence();
    __syncthreads();
    RC_PH(PH_FLUSH);
    if (!replay && tid == 0) {
      const int done = __ldcg(p.progress + chain) + 1;  // this CTA owns the chain until it is pushed back
      p.progress[chain] = done;
      const unsigned long long all_done = atomicAdd(p.steps_done, 1ull) + 1ull;
      if (kShare && share && *(volatile int32_t*)p.help_count > 0) c.s.misc[M_HELPMODE] = 1;  // look at it before the next ticket
      // The launch ends with its slowest chain, and a chain that has fallen behind (a step with
      // hundreds of retries) loses more time yet queueing between its steps.  So a chain more than
      // kLagMargin steps behind the average keeps its CTA until it has caught up: with more chains
      // than CTAs the field queues ~ n_chains / grid step times per step, the laggard none.
      if (done < p.n_steps && (unsigned long long)(done + kLagMargin) * (unsigned long long)n_chains <= all_done) {
        c.s.misc[M_KEEP] = chain;
      } else if (done < p.n_steps) {
        const unsigned long long q = atomicAdd(p.work_counter + 1, 1ull);
        int32_t* slot = p.ring + q % cap;
        while (ld_acquire(slot) != 0) __nanosleep(32);  // its previous content has been taken
        st_release(slot, chain + 1);
      }
    }
    if (replay && status != RC_OK) break;
  }
#ifdef RC_TIMERS
  __syncthreads();
  if (tid <= PH_N) atomicAdd(p.timers + tid, reinterpret_cast<unsigned long long*>(c.s.misc + 96)[tid]);  // slots 0..14
#endif
}

// =========================================================================================
// Seed plans: recursive_tree_part (tree.py:971-1085) with method = bipartition_tree on
// random_spanning_tree, node_repeats = 1.  One CTA builds one plan: k - 2 one-sided splits
// of the shrinking remainder (label RC_SEED_REMAINING) with the debt-tightened target
// (tree.py:1021-1050), then one two-sided split (tree.py:1054-1061).  The merged "pair" is
// the whole remainder, so the engine must have been created with caps covering the graph
// (the working arrays then usually live in the global scratch: spill mode).
template <bool kSpill>
__global__ void __launch_bounds__(512, 2) seed_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Chain c;
  c.p = &p;
  smem_layout(&c.s, smem_raw, p.g.N, p.cap_nodes, p.cap_edges, p.g.R,
              kSpill ? p.spill + (size_t)blockIdx.x * p.spill_stride : nullptr,
              p.cold + (size_t)blockIdx.x * p.cold_stride);
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  c.nodes = p.nodes + (size_t)blockIdx.x * p.cap_nodes;
  c.eids = p.eids + (size_t)blockIdx.x * p.cap_edges;
  c.par = 0;
  int32_t* ctr = s.misc + M_CTR;
  const bool replay = p.rp.n_steps > 0;
  const int k = p.cfg.n_districts;
  const double pt = p.cfg.pop_target, eps = p.cfg.epsilon;
  const double lb_pop = __dmul_rn(pt, __dsub_rn(1.0, eps)), ub_pop = __dmul_rn(pt, __dadd_rn(1.0, eps));
  const long long max_attempts = p.seed_max_attempts > 0 ? p.seed_max_attempts : 0x7fffffffffffffffLL;
  for (int chain = replay ? p.rp.chain : (int)blockIdx.x; chain < p.cfg.n_chains; chain += gridDim.x) {
    c.chain = chain;
    c.gchain = (uint32_t)(chain + p.cfg.chain_offset);
    c.assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
    c.mask = p.st.d_cut_mask + (size_t)chain * p.st.mask_stride;
    c.tally = p.st.d_tally + (size_t)chain * k;
    if (tid < RC_N_COUNTERS) ctr[tid] = 0;
    for (int i = tid; i < (int)p.st.assign_stride; i += T) c.assign[i] = RC_SEED_REMAINING;
    for (int d = tid; d < k; d += T) c.tally[d] = 0;
    __syncthreads();
    int status = RC_OK;
    double debt = 0.0;
    for (int part = 0; part < k - 1 && status == RC_OK; ++part) {
      const bool last = part == k - 2;
      double ideal = pt, e1 = eps;
      if (!last) {  // tree.py:1021-1023
        const double min_pop = fmax(lb_pop, __dsub_rn(lb_pop, debt));
        const double max_pop = fmin(ub_pop, __dsub_rn(ub_pop, debt));
        ideal = __ddiv_rn(__dadd_rn(min_pop, max_pop), 2.0);
        e1 = __ddiv_rn(__dsub_rn(max_pop, min_pop), __dmul_rn(2.0, ideal));
      }
      const double thr = __dmul_rn(ideal, e1);
      const uint32_t prop_no = RC_SEED_PROPOSAL | (uint32_t)part;
      const RcReplayStep* rs = replay ? p.rp.steps + part : nullptr;
      const int n_rec = rs ? rs->n_trees : 0;
      Step st;
      st.a = st.b = RC_SEED_REMAINING;
      int d_sub = 0;
      status = gather_pair<false>(c, st, &d_sub);
      if (tid == 0) { ctr[RC_CTR_NODES] += st.n; ctr[RC_CTR_SLOTS] += d_sub; }
      if (status) break;
      const int n = st.n, nt = n - 1;
      if (n < 3) { status = RC_ERR_NO_ROOT; break; }
      for (uint32_t tree_no = 0;; ++tree_no) {
        if ((long long)tree_no >= max_attempts) { status = RC_ERR_MAX_ATTEMPTS; break; }
        if (rs && (int)tree_no >= n_rec) { status = RC_ERR_REPLAY; break; }
        const uint32_t* rrank = rs ? p.rp.weight_rank + (rs->first_tree + tree_no) * (long long)p.g.E : nullptr;
        int rounds = 0;
        status = boruvka_mst(c, st, tree_no, prop_no, rrank, &rounds);
        if (tid == 0) { ctr[RC_CTR_TREES] += 1; ctr[RC_CTR_ROUNDS] += rounds; }
        if (status) break;
        const int r0 = 0;
        euler_subtree_pops(c, st, r0);
        // the reference's root (tree.py:377): uniform over tree nodes of degree > 1
        auto internal = [&](int i) { return s.nxt[s.head[i]] != kNone; };
        int root;
        const bool last_rec = rs && (int)tree_no == n_rec - 1;
        if (last_rec) {
          const int gr = rs->root;
          root = (gr >= 0 && gr < p.g.N && is_member(s, gr)) ? local_id(s, gr) : -1;
          if (root < 0 || !internal(root)) { status = RC_ERR_REPLAY; break; }
        } else {
          if (tid == 0) s.misc[M_NINT] = 0;
          __syncthreads();
          int cnt = 0;
          for (int i = tid; i < n; i += T) cnt += internal(i) ? 1 : 0;
          cnt = warp_sum(cnt);
          if ((tid & 31) == 0 && cnt) atomicAdd(&s.misc[M_NINT], cnt);
          __syncthreads();
          const int n_int = s.misc[M_NINT];
          RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_ROOT, 0, tree_no, prop_no, 0);
          root = select_kth(c, n, (int)rc_below(r.v[0], r.v[1], (uint32_t)n_int), internal);
        }
        // population below edge t as seen from the reference's root
        const int ent_root = s.ent[root];
        auto below_pop = [&](int t) {
          const int pa = s.P[2 * t], pb = s.P[2 * t + 1];
          const bool root_in_child = ent_root != 0xffff && ent_root >= min(pa, pb) && ent_root <= max(pa, pb);
          return root_in_child ? st.tot - s.cur[t] : s.cur[t];
        };
        auto fits = [&](int pop) { return fabs(__dsub_rn((double)pop, ideal)) <= thr; };
        auto is_cut = [&](int t) {  // tree.py:386-424
          const int pb = below_pop(t);
          return last ? (fits(pb) && fits(st.tot - pb)) : (fits(pb) || fits(st.tot - pb));
        };
        if (tid == 0) s.misc[M_NBAL] = 0;
        __syncthreads();
        {
          int cnt = 0;
          for (int t = tid; t < nt; t += T) cnt += is_cut(t) ? 1 : 0;
          cnt = warp_sum(cnt);
          if ((tid & 31) == 0 && cnt) atomicAdd(&s.misc[M_NBAL], cnt);
        }
        __syncthreads();
        const int nbal = s.misc[M_NBAL];
        if (nbal == 0) continue;
        if (rs && !last_rec) { status = RC_ERR_REPLAY; break; }
        int tstar;
        if (rs && rs->cut_edge >= 0) {
          const int want = rs->cut_edge;
          tstar = select_kth(c, nt, 0, [&](int t) { return c.eids[s.te[t]] == want && is_cut(t); });
          if (tstar < 0) { status = RC_ERR_REPLAY; break; }
        } else {
          RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_CUT, 0, tree_no, prop_no, 0);
          tstar = select_kth(c, nt, (int)rc_below(r.v[0], r.v[1], (uint32_t)nbal), is_cut);
        }
        if (tid == 0) ctr[RC_CTR_ATTEMPTS] += (int)tree_no + 1;
        // labels: `below` = the side of the cut that does not hold the reference's root
        const int pa = s.P[2 * tstar], pb = s.P[2 * tstar + 1];
        const int pd = min(pa, pb), pu = max(pa, pb);
        auto child_side = [&](int i) { const int en = s.ent[i]; return en != 0xffff && en >= pd && en <= pu; };
        const bool root_cs = child_side(root);
        const int pop_below = below_pop(tstar);
        long long part_pop;
        if (!last) {
          const bool take_below = fits(pop_below);  // tree.py:388 first, :399 otherwise
          part_pop = take_below ? pop_below : st.tot - pop_below;
          for (int i = tid; i < n; i += T) {
            const bool below = child_side(i) != root_cs;
            if (below == take_below) c.assign[c.nodes[i]] = (rc_label_t)part;
          }
          if (tid == 0) c.tally[part] = part_pop;
          if (!(lb_pop <= (double)part_pop && (double)part_pop <= ub_pop)) status = RC_ERR_BALANCE;
          debt = __dadd_rn(debt, __dsub_rn((double)part_pop, pt));
        } else {  // root side -> parts[-2], the rest -> parts[-1] (tree.py:1054-1075)
          for (int i = tid; i < n; i += T)
            c.assign[c.nodes[i]] = (rc_label_t)((child_side(i) != root_cs) ? k - 1 : k - 2);
          const long long p_hi = pop_below, p_lo = (long long)st.tot - pop_below;
          if (tid == 0) { c.tally[k - 1] = p_hi; c.tally[k - 2] = p_lo; }
          if (!(lb_pop <= (double)p_hi && (double)p_hi <= ub_pop && lb_pop <= (double)p_lo && (double)p_lo <= ub_pop))
            status = RC_ERR_BALANCE;
        }
        break;
      }
      __syncthreads();  // the labels written above are read by the next gather
    }
    __syncthreads();
    if (tid == 0) {
      p.st.d_status[chain] = status;
      for (int i = 0; i < RC_N_COUNTERS; ++i) p.st.d_counters[(size_t)chain * RC_N_COUNTERS + i] = ctr[i];
    }
    __syncthreads();
    if (replay) break;
  }
}

// ---- state kernels -----------------------------------------------------------------------
// Tally._initialize_tally (updaters/tally.py:118-141) + first cut_edges scan
// (updaters/cut_edges.py:110-114); one CTA per chain.
__global__ void init_state_kernel(const Params p) {
  __shared__ int cnt;
  const int chain = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const rc_label_t* assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
  uint32_t* mask = p.st.d_cut_mask + (size_t)chain * p.st.mask_stride;
  int64_t* tally = p.st.d_tally + (size_t)chain * p.cfg.n_districts;
  const int k = p.cfg.n_districts;
  if (tid == 0) cnt = 0;
  for (int d = tid; d < k; d += T) tally[d] = 0;
  __syncthreads();
  bool bad = false;
  for (int v = tid; v < p.g.N; v += T) {
    const int d = assign[v];
    if (d >= k) { bad = true; continue; }
    atomicAdd(reinterpret_cast<unsigned long long*>(tally + d), (unsigned long long)(long long)p.g.pop[v]);
  }
  const int words = (p.g.E + 31) >> 5;
  int local = 0;
  for (int w = tid; w < (int)p.st.mask_stride; w += T) {
    uint32_t bits = 0;
    if (w < words)
      for (int b = 0; b < 32; ++b) {
        const int e = w * 32 + b;
        if (e < p.g.E && assign[p.g.edge_uv[2 * e]] != assign[p.g.edge_uv[2 * e + 1]]) bits |= 1u << b;
      }
    mask[w] = bits;
    local += __popc(bits);
  }
  atomicAdd(&cnt, local);
  __syncthreads();
  if (tid == 0) {
    p.st.d_cut_count[chain] = cnt;
    if (p.st.d_last_pair) { p.st.d_last_pair[2 * chain] = -1; p.st.d_last_pair[2 * chain + 1] = -1; }
    if (!p.keep_counters) {  // a re-upload of plans keeps the RNG position and the statistics
      p.st.d_step_no[chain] = 0;
      p.st.d_proposal_no[chain] = 0;
      p.st.d_flags[chain] = 0;
      for (int i = 0; i < RC_N_COUNTERS; ++i) p.st.d_counters[(size_t)chain * RC_N_COUNTERS + i] = 0;
    }
  }
  const bool any_bad = __syncthreads_or(bad);
  if (tid == 0) {
    if (p.keep_counters == 2 && p.st.d_status[chain] != RC_OK) { /* seed plan failed: keep why */ }
    else p.st.d_status[chain] = any_bad ? RC_ERR_INVALID_ARG : RC_OK;
  }
}

// Tally of an arbitrary integer node column for every chain (updaters/tally.py:69-170 for
// columns other than the population the proposal balances - e.g. the vote counts of an
// Election updater): out[chain][d] = sum of column[v] over nodes v of district d.  One CTA
// per chain, shared-memory bins, 128-bit loads of the assignment row.
__global__ void tally_column_kernel(const Params p, const int32_t* __restrict__ column, long long* __restrict__ out) {
  const int chain = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int k = p.cfg.n_districts, N = p.g.N;
  const rc_label_t* assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
#if RC_LABEL_BITS == 8
  __shared__ unsigned long long bins[256];
  for (int d = tid; d < 256; d += T) bins[d] = 0;
  __syncthreads();
  for (int v0 = tid * 16; v0 < N; v0 += T * 16) {
    const uint4 q = __ldcg(reinterpret_cast<const uint4*>(assign + v0));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int v = v0 + i;
      if (v < N) {
        const uint32_t d = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
        atomicAdd(&bins[d], (unsigned long long)(long long)__ldg(column + v));
      }
    }
  }
  __syncthreads();
  for (int d = tid; d < k; d += T) out[(size_t)chain * k + d] = (long long)bins[d];
#else
  // up to 32 767 districts: the bins are the chain's output row itself (L2 atomics)
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(out + (size_t)chain * k);
  for (int d = tid; d < k; d += T) bins[d] = 0;
  __syncthreads();
  for (int v0 = tid * 8; v0 < N; v0 += T * 8) {
    const uint4 q = __ldcg(reinterpret_cast<const uint4*>(assign + v0));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = v0 + i;
      if (v < N) {
        const uint32_t d = (w[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
        if (d < (uint32_t)k) atomicAdd(&bins[d], (unsigned long long)(long long)__ldg(column + v));
      }
    }
  }
#endif
}

// Ensemble histograms (one thread per chain).
__global__ void histogram_kernel(const Params p, unsigned long long* cut_hist, int n_cut_bins,
                                 unsigned long long* dev_hist, int n_dev_bins, double dev_bin_width) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= p.cfg.n_chains) return;
  if (cut_hist) {
    const int cc = p.st.d_cut_count[chain];
    atomicAdd(cut_hist + min(max(cc, 0), n_cut_bins - 1), 1ull);
  }
  if (dev_hist) {
    const int64_t* tally = p.st.d_tally + (size_t)chain * p.cfg.n_districts;
    double worst = 0.0;
    for (int d = 0; d < p.cfg.n_districts; ++d)
      worst = fmax(worst, fabs(((double)tally[d] - p.cfg.pop_target) / p.cfg.pop_target));
    const int bin = (int)fmin(floor(worst / dev_bin_width), (double)(n_dev_bins - 1));
    atomicAdd(dev_hist + bin, 1ull);
  }
}

}  // namespace rc
