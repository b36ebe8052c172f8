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
  uint8_t* assign;
  uint32_t* mask;
  int64_t* tally;
  int32_t* nodes;
  int32_t* eids;
  int par;              // block_excl_scan bank parity (toggled in lock step by all threads)
};

// ---- membership of a global node in the merged pair, and its local id ---------------------
__device__ __forceinline__ bool is_member(const Smem& s, int g) { return (s.mbits[g >> 5] >> (g & 31)) & 1u; }
__device__ __forceinline__ int local_id(const Smem& s, int g) {
  return (int)s.mpref[g >> 5] + __popc(s.mbits[g >> 5] & ((1u << (g & 31)) - 1u));
}

// 4 assignment bytes -> 4 bits (byte k == a or == b)
__device__ __forceinline__ uint32_t match4(uint32_t x, uint32_t a4, uint32_t b4) {
  const uint32_t r = __vcmpeq4(x, a4) | __vcmpeq4(x, b4);
  return ((((r >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu;
}

__device__ __forceinline__ uint32_t match16(const uint4 q, uint32_t a4, uint32_t b4) {
  return match4(q.x, a4, b4) | match4(q.y, a4, b4) << 4 | match4(q.z, a4, b4) << 8 | match4(q.w, a4, b4) << 12;
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

// Wilson only: adjacency lists of the merged pair (loff / ladj / ladje), every row in
// ascending neighbour id - the order in which the walk's draw indexes the neighbours.
__device__ void build_local_adjacency(Chain& c, const Step& st) {
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, m = st.m;
  for (int i = tid; i < n; i += T) s.wtmp[i] = 0;
  __syncthreads();
  for (int j = tid; j < m; j += T) { atomicAdd(&s.wtmp[s.eu[j]], 1u); atomicAdd(&s.wtmp[s.ev[j]], 1u); }
  __syncthreads();
  {
    const int per = odd_chunk(n);
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += (int)s.wtmp[i];
    int tot;
    int run = block_excl_scan(sum, &tot, s.misc, c.par);
    for (int i = lo; i < hi; ++i) { const int d = (int)s.wtmp[i]; s.loff[i] = (lid_t)run; s.wtmp[i] = (uint32_t)run; run += d; }
    if (tid == 0) s.loff[n] = (lid_t)(2 * m);
  }
  __syncthreads();
  for (int j = tid; j < m; j += T) {
    const int u = s.eu[j], v = s.ev[j];
    const int pu = (int)atomicAdd(&s.wtmp[u], 1u);
    s.ladj[pu] = (lid_t)v; s.ladje[pu] = (lid_t)j;
    const int pv = (int)atomicAdd(&s.wtmp[v], 1u);
    s.ladj[pv] = (lid_t)u; s.ladje[pv] = (lid_t)j;
  }
  __syncthreads();
  for (int i = tid; i < n; i += T) {  // insertion sort of a short row
    const int r0 = s.loff[i], r1 = s.loff[i + 1];
    for (int a = r0 + 1; a < r1; ++a) {
      const lid_t kv = s.ladj[a], ke = s.ladje[a];
      int b = a - 1;
      while (b >= r0 && s.ladj[b] > kv) { s.ladj[b + 1] = s.ladj[b]; s.ladje[b + 1] = s.ladje[b]; --b; }
      s.ladj[b + 1] = kv; s.ladje[b + 1] = ke;
    }
  }
  __syncthreads();
}

// ---- P1: gather the merged pair into shared memory ---------------------------------------
// returns RC_OK / RC_ERR_CAPACITY (block uniform)
template <bool kWilson>
__device__ int gather_pair(Chain& c, Step& st, int* d_sub_out) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int N = p.g.N, NW = (N + 31) >> 5;
  const uint32_t a4 = 0x01010101u * (uint32_t)st.a, b4 = 0x01010101u * (uint32_t)st.b;
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
      const uint4* src = reinterpret_cast<const uint4*>(c.assign + (size_t)w * 32);
      bits = match16(__ldcg(src), a4, b4);
      if (w * 32 + 16 < N) bits |= match16(__ldcg(src + 1), a4, b4) << 16;
      const int rem = N - w * 32;
      if (rem < 32) bits &= (1u << rem) - 1u;
      s.mbits[w] = bits;
    }
    int tot;
    const int ex = block_excl_scan(__popc(bits), &tot, s.misc, c.par);
    if (w < NW) s.mpref[w] = (lid_t)min(n + ex, 0xffff);
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
        int lo = 0, hi = NW;  // first word whose prefix exceeds i
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((int)s.mpref[mid] > i) hi = mid; else lo = mid + 1;
        }
        const int w = lo - 1;
        g = w * 32 + (int)__fns(s.mbits[w], 0, i - (int)s.mpref[w] + 1);
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
        s.eu[j] = (lid_t)i;
        s.ev[j] = (lid_t)local_id(s, h);
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
  if (kWilson) build_local_adjacency(c, st);
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
  for (int j = tid; j < m; j += T) { s.lab[j] = (uint32_t)s.eu[j] | ((uint32_t)s.ev[j] << 16); s.lj[j] = (lid_t)j; }
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
      uint32_t nab[kCompactK], nj[kCompactK];
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < kCompactK; ++k) {
        const int q = base + k * T + tid;
        nab[k] = 0xffffffffu;
        nj[k] = 0;
        if (base + k * T >= nlive) break;  // block uniform
        if (q < nlive) {
          const uint32_t ab = s.lab[q];
          const int a = s.hook[ab & 0xffffu], b = s.hook[ab >> 16];
          if (a != b) { nab[k] = (uint32_t)a | ((uint32_t)b << 16); nj[k] = s.lj[q]; ++cnt; }
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
        if (nab[k] != 0xffffffffu) { s.lab[o] = nab[k]; s.lj[o] = (lid_t)nj[k]; ++o; }
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

// ---- P2b: uniform spanning tree (Wilson, tree.py:97-132) by parallel cycle popping -------
// Every node owns a stack of successor draws: draw j at node u is a pure function of (u, j)
// (recom_rng.h).  Pointing every node at the top of its stack gives a functional graph; its
// cycles are popped (every node on a cycle takes its next draw) until none is left.  The
// popped cycles and the final tree do not depend on the popping order (Wilson 1996), so
// this is the tree - and the per-node draw counts - of the reference's sequential
// loop-erased walks fed with the same stacks.  Cycles of up to kHop nodes are found by a
// look-ahead from the nodes whose pointer just changed; longer ones by pointer doubling
// over the functional graph.
//
// Draws are tabulated: wtab[u] is a ring of 8 successors, draws [hi - 8, hi) of node u at
// slot (j & 7), filled four at a time (one Philox block) - the first 8 by everybody at the
// start, the later ones by whoever consumes draw j == hi.  A pop is then one table look-up;
// the generator runs once per four pops, off the pointer-chasing paths.
constexpr int kHop = 8;      // a worklist node looks this many pointers ahead for a cycle through itself
constexpr int kHandoff = 1;  // default RcConfig.wilson_handoff (measured: the parallel rounds beat the lone thread to the very end)
constexpr int kMaxPasses = 4096;    // safety (disconnected input plans): then phase 3's budget applies
constexpr int kTab = 8;  // ring of tabulated draws per node
constexpr int kWarpRoundsEnter = 32;  // worklists this short are handled by warp 0 alone ...
constexpr int kWarpRoundsExit = 64;   // ... until they outgrow this (multiple of 32)
constexpr uint32_t kInTree = 0xffffu;  // walk stamp of a node that leads to the root


__device__ int wilson_tree(Chain& c, const Step& st, uint32_t tree_no, uint32_t prop_no, long long tree_idx,
                           int32_t* ctr) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, m = st.m;
  const bool replay = p.rp.n_steps > 0;
  int root;
  if (replay) {
    const int gr = p.rp.wilson_root[tree_idx];
    root = (gr >= 0 && gr < p.g.N && is_member(s, gr)) ? local_id(s, gr) : -1;
    if (root < 0) return RC_ERR_REPLAY;
  } else {
    const RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_WROOT, 0, tree_no, prop_no, 0);
    root = (int)rc_below(r.v[0], r.v[1], (uint32_t)n);
  }
  const int64_t* sp = replay ? p.rp.stack_ptr + tree_idx * ((long long)p.g.N + 1) : nullptr;
  // replay: successor draw j of node u from the recorded stacks (tree.py:119)
  auto recorded = [&](int u, uint32_t j) {
    const int r0 = s.loff[u], d = s.loff[u + 1] - r0;
    const int gu = c.nodes[u];
    int nn = -1;
    if (sp[gu] + (long long)j < sp[gu + 1]) {
      const int gv = p.rp.stack_val[sp[gu] + j];
      if (gv >= 0 && gv < p.g.N && is_member(s, gv)) {
        const int lv = local_id(s, gv);
        for (int q = 0; q < d; ++q) if (s.ladj[r0 + q] == lv) nn = lv;
      }
    }
    if (nn < 0) { s.misc[M_ERR] = RC_ERR_REPLAY; nn = s.ladj[r0]; }
    return nn;
  };
  // free-running: draws 4b .. 4b + 3 of node u into their ring slots (one 64-bit store)
  auto fill_block = [&](int u, uint32_t b) {
    const int r0 = s.loff[u];
    const uint32_t d = (uint32_t)(s.loff[u + 1] - r0);
    const RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_WALK, (uint32_t)u, tree_no, prop_no, b);
    const uint32_t a0 = s.ladj[r0 + (int)rc_below32(r.v[0], d)], a1 = s.ladj[r0 + (int)rc_below32(r.v[1], d)];
    const uint32_t a2 = s.ladj[r0 + (int)rc_below32(r.v[2], d)], a3 = s.ladj[r0 + (int)rc_below32(r.v[3], d)];
    *reinterpret_cast<uint2*>(s.wtab + u * kTab + ((b & 1u) << 2)) = make_uint2(a0 | (a1 << 16), a2 | (a3 << 16));
  };
  // node u takes its next draw (one thread per node and round).  The end of the tabulated
  // draws is implied in phase 1 - max(8, 4 * ceil(j / 4)): blocks beyond the first two are
  // filled exactly when their first draw is taken
  auto pop1 = [&](int u) {
    const uint32_t j = s.wcnt[u];
    int nn;
    if (replay) {
      nn = recorded(u, j);
    } else {
      if (j >= (uint32_t)kTab && (j & 3u) == 0) fill_block(u, j >> 2);
      nn = s.wtab[u * kTab + (j & 7u)];
    }
    s.wnx[u] = (lid_t)nn;
    s.wcnt[u] = (lid_t)(j + 1);
  };
  auto succ_of = [&](int u) { return (int)s.wnx[u]; };  // wnx[root] == root
  if (tid == 0) s.misc[M_ERR] = 0;
  for (int u = tid; u < n; u += T) {
    if (u != root) {
      if (s.loff[u + 1] == s.loff[u]) { s.misc[M_ERR] = RC_ERR_DISCONNECTED; continue; }
      if (!replay) { fill_block(u, 0); fill_block(u, 1); }
      s.wnx[u] = (lid_t)(replay ? recorded(u, 0) : (int)s.wtab[u * kTab]);
      s.wcnt[u] = 1;
    } else {
      s.wnx[u] = (lid_t)u;
      s.wcnt[u] = 0;
    }
  }
  for (int i = tid; i < (m + 31) >> 5; i += T) s.tflag[i] = 0;
  for (int i = tid; i < (n + 31) >> 5; i += T) s.wmark[i] = 0;
  __syncthreads();
  if (s.misc[M_ERR]) return s.misc[M_ERR];
  // Phase 1 - pop cycles in parallel, driven by a worklist.  A cycle that was not there
  // before must run through a node whose pointer just changed, so only the nodes popped in
  // the previous round are examined: the thread that holds worklist entry u follows the
  // pointers for up to kHop steps; if that leads back to u it has found a cycle (its members
  // in registers).  Every worklist member of the cycle finds it; the one that first sets the
  // bit of the cycle's lowest node pops it - all members take their next draw and form the
  // next worklist.  Round one examines every node.  Work per round is O(popped), not O(n):
  // late rounds keep a warp or two busy and leave the SM to the co-resident CTA.
  //   When the worklist runs dry, what is left are cycles longer than kHop (or none): one
  // global detection (pointer doubling: after >= n steps every node sits on a cycle or at the
  // root, and every cycle node is somebody's landing point) pops them and refills the
  // worklist.  A global detection that finds nothing ends the phase: the tree is complete.
  // RcConfig.wilson_handoff > 1 ends it earlier (worklist / detection yield below the
  // threshold) and leaves the rest to the sequential finish - kept as a tuning and test knob.
  const int kLocalThreshold = p.cfg.wilson_handoff > 0 ? p.cfg.wilson_handoff : kHandoff;
  const int kFullThreshold = 2 * kLocalThreshold;
  int walk = 0, passes = 0;
  lid_t* __restrict__ P = s.wgA;  // worklists alias the pointer-doubling buffers (never live together)
  lid_t* __restrict__ Q = s.wgB;
  int32_t* const wq = s.misc + M_WQ;  // wq[0], wq[1]: length of the list being built, by round parity
  for (int i = tid; i < n - 1; i += T) P[i] = (lid_t)(i + (i >= root));
  if (tid == 0) { wq[0] = 0; wq[1] = 0; }
  int np = n - 1, par = 0;
  __syncthreads();
  // one worklist entry: look ahead from P[i]; the owner of a cycle appends its members to Q
  // and returns the cycle's lowest node (whose mark it clears after the round), else kNoCycle
  constexpr uint32_t kNoCycle = 0xffffffffu;
  auto detect = [&](const lid_t* __restrict__ Pl, lid_t* __restrict__ Ql, int i, int parity) -> uint32_t {
    const int u = Pl[i];
    lid_t mem[kHop];
    int L = 0, v = u;
    uint32_t mn = (uint32_t)u;
#pragma unroll
    for (int h = 0; h < kHop; ++h) {
      mem[h] = (lid_t)v;
      v = succ_of(v);
      if (v == u) { L = h + 1; break; }
      if (v == root) break;
      mn = min(mn, (uint32_t)v);
    }
    if (L && !((atomicOr(&s.wmark[mn >> 5], 1u << (mn & 31u)) >> (mn & 31u)) & 1u)) {
      const int slot = atomicAdd(&wq[parity], L);
#pragma unroll
      for (int h = 0; h < kHop; ++h) if (h < L) Ql[slot + h] = mem[h];
      return mn;
    }
    return kNoCycle;
  };
  int32_t* const wbc = s.misc + M_WBC;  // what warp 0 hands back to the block: np, rounds done
  for (;;) {
    while (np >= kLocalThreshold && passes <= kMaxPasses) {
      if (np <= kWarpRoundsEnter) {
        // Short worklists (nearly every round after the first few): warp 0 runs the rounds on
        // its own, separated by __syncwarp() instead of block barriers, while the other warps
        // wait at ONE barrier for the whole stretch and issue nothing.  It hands the block back
        // a worklist that outgrew it (a popped cycle can be longer than the list that found it)
        // or ran dry.
        RC_PH(PH_W1);  // timers build: W1 = block-wide rounds, W2 / W3 = detect / pop of warp-0 rounds, slot 14 = global detections
        if (tid < 32) {
          int rounds = 0;
          int npw = np, parw = par;
          lid_t* Pw = P; lid_t* Qw = Q;
          while (npw >= kLocalThreshold && npw <= kWarpRoundsExit && passes + rounds <= kMaxPasses) {
            ++rounds;
            uint32_t own[kWarpRoundsExit / 32];
#pragma unroll
            for (int r = 0; r < kWarpRoundsExit / 32; ++r) {
              const int i = r * 32 + tid;
              own[r] = i < npw ? detect(Pw, Qw, i, parw) : kNoCycle;
            }
            __syncwarp();
            RC_PH(PH_W2);
            const int nq = *(volatile int32_t*)&wq[parw];
            for (int i = tid; i < nq; i += 32) { pop1(Qw[i]); ++walk; }
#pragma unroll
            for (int r = 0; r < kWarpRoundsExit / 32; ++r)
              if (own[r] != kNoCycle) atomicAnd(&s.wmark[own[r] >> 5], ~(1u << (own[r] & 31u)));
            if (tid == 0) wq[parw ^ 1] = 0;
            __syncwarp();
            RC_PH(PH_W3);
            npw = nq;
            parw ^= 1;
            lid_t* t_ = Pw; Pw = Qw; Qw = t_;
          }
          if (tid == 0) { wbc[0] = npw; wbc[1] = rounds; }
        }
        __syncthreads();
        const int rounds = wbc[1];
        np = wbc[0];
        passes += rounds;
        if (rounds & 1) { par ^= 1; lid_t* t_ = P; P = Q; Q = t_; }
        __syncthreads();  // wbc is rewritten only after everybody has read it
        if (rounds > 0) continue;
      }
      ++passes;
      for (int i = tid; i < np; i += T) detect(P, Q, i, par);
      __syncthreads();  // every pointer has been read; Q is complete (a cycle's nodes appear once)
      const int nq = wq[par];
      for (int i = tid; i < nq; i += T) { pop1(Q[i]); ++walk; }
      for (int i = tid; i < (n + 31) >> 5; i += T) s.wmark[i] = 0;  // the owners' marks
      if (tid == 0) wq[par ^ 1] = 0;  // the next round's counter: everybody has read it by now
      __syncthreads();
      np = nq;
      par ^= 1;
      lid_t* t_ = P; P = Q; Q = t_;
    }
    if (passes > kMaxPasses) break;
    ++passes;
    RC_PH(PH_W1);
    // global detection (the worklist is dead: P and Q are free)
    lid_t* __restrict__ A = s.wgA;
    lid_t* __restrict__ B = s.wgB;
    for (int u = tid; u < n; u += T) A[u] = (lid_t)succ_of(u);
    __syncthreads();
    if (tid == 0) wq[par ^ 1] = 0;
    for (int span = 1; span < n; span <<= 1) {
      for (int u = tid; u < n; u += T) B[u] = A[A[u]];
      __syncthreads();
      lid_t* tmp = A; A = B; B = tmp;
    }
    for (int u = tid; u < n; u += T) {
      const int v = A[u];
      if (v != root) atomicOr(&s.wmark[v >> 5], 1u << (v & 31));
    }
    __syncthreads();  // A and B are dead from here on
    P = s.wgA;
    Q = s.wgB;
    for (int u0 = 0; u0 < n; u0 += T) {
      const int u = u0 + tid;
      const bool hit = u < n && ((s.wmark[u >> 5] >> (u & 31)) & 1u);
      if (hit) { pop1(u); ++walk; }
      // warp-aggregated append to the worklist
      const uint32_t bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        int slot = 0;
        if ((tid & 31) == 0) slot = atomicAdd(&wq[par], __popc(bal));
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (hit) P[slot + __popc(bal & ((1u << (tid & 31)) - 1u))] = (lid_t)u;
      }
    }
    __syncthreads();
    np = wq[par];
    par ^= 1;
    for (int i = tid; i < (n + 31) >> 5; i += T) s.wmark[i] = 0;
    __syncthreads();
    RC_PH(PH_N);  // slot 14
    if (np < kFullThreshold) break;
  }
  RC_PH(PH_W1);
  // A global detection that found no cycle at all means the pointers ARE the tree (the
  // normal end with the default hand-off threshold): phases 2-3 have nothing to do.
  const bool complete = np == 0 && passes <= kMaxPasses;
  if (!complete) {
  // Phase 2 - which nodes already lead to the root?  After >= n steps of the functional
  // graph every node sits on a cycle or at the root.
  for (int u = tid; u < n; u += T) s.wgA[u] = (lid_t)succ_of(u);
  __syncthreads();
  lid_t* __restrict__ A = s.wgA;
  lid_t* __restrict__ B = s.wgB;
  for (int span = 1; span < n; span <<= 1) {
    for (int u = tid; u < n; u += T) B[u] = A[A[u]];
    __syncthreads();
    lid_t* tmp = A; A = B; B = tmp;
  }
  // ... and, for the sequential part, one packed state word per node - successor (bits
  // 0-15), walk stamp (16-31; kInTree: leads to the root), draws consumed (32-47), end of the
  // tabulated draws hi (48-63) - with every ring topped up to at least 5 unread draws by all
  // threads.  One 64-bit load per walk step, one 16-bit look-up per pop.
  unsigned long long* __restrict__ wst = s.wst;
  for (int u = tid; u < n; u += T) {
    const uint32_t cnt = s.wcnt[u];
    uint32_t hi = cnt > (uint32_t)kTab ? ((cnt + 3u) & ~3u) : (uint32_t)kTab;
    if (!replay && u != root)
      while (hi - cnt <= 4u) { fill_block(u, hi >> 2); hi += 4u; }
    const uint32_t stamp = (int)A[u] == root ? kInTree : 0u;
    wst[u] = (unsigned long long)s.wnx[u] | ((unsigned long long)stamp << 16) |
             ((unsigned long long)cnt << 32) | ((unsigned long long)hi << 48);
  }
  __syncthreads();
  RC_PH(PH_W2);
  // Phase 3 - the remaining cycles hang on a handful of components that pop one after the
  // other (the long tail of Wilson's algorithm is sequential): one thread finishes them as
  // loop-erased walks over the current pointers, drawing only where a loop closes.  The
  // loops are written for a lone thread: the next node's word is requested before the
  // current one is dealt with, and only the half of a word that changes is stored.
  if (tid == s.misc[M_WALKER]) {
    // the walk proper, over accessors: M.ld(u) / M.st(u, w) the 64-bit word of node u,
    // M.ldlo / M.stlo its low half (successor | stamp), M.tab(i) entry i of the draw rings
    auto finish = [&](auto M) {
      auto pop = [&](int v, unsigned long long w) {  // node v takes its next draw; returns new word
        const uint32_t cnt = (uint32_t)(w >> 32) & 0xffffu;
        uint32_t hi = (uint32_t)(w >> 48);
        uint32_t nn;
        if (replay) {
          nn = (uint32_t)recorded(v, cnt);
        } else {
          if (cnt == hi) { fill_block(v, hi >> 2); hi += 4u; }
          nn = M.tab(v * kTab + (int)(cnt & 7u));
        }
        return (unsigned long long)nn | ((unsigned long long)((cnt + 1u) & 0xffffu) << 32) |
               ((unsigned long long)(hi & 0xffffu) << 48);  // stamp cleared
      };
      // a connected pair needs ~10-20 draws per node; a walk that cannot reach the root
      // (disconnected input plan) is cut off instead of spinning for ever
      int left = 2048 * n + 65536;  // n < 65536
      int pops = 0;
      for (int node = 0; node < n && left >= 0; ++node) {
        unsigned long long w = M.ld(node);
        if ((((uint32_t)w) >> 16) == kInTree) continue;
        const uint32_t id = (uint32_t)(node % 0xfffe) + 1u;  // 1..0xfffe, never kInTree
        const uint32_t idbits = id << 16;
        int u = node;
        for (;;) {  // w == word of u, u does not lead to the root yet
          uint32_t lo = (uint32_t)w;
          if ((lo >> 16) == id) {  // the walk closed a loop at u: pop it
            int v = u;
            unsigned long long wv = w;
            do {
              const int vn = (int)((uint32_t)wv & 0xffffu);
              const unsigned long long wn = M.ld(vn);
              M.st(v, pop(v, wv));
              ++pops;
              v = vn;
              wv = wn;
            } while (v != u);
            lo = M.ldlo(u);
          }
          const int un = (int)(lo & 0xffffu);
          w = M.ld(un);
          M.stlo(u, (lo & 0xffffu) | idbits);
          u = un;
          if (--left < 0 || (((uint32_t)w) >> 16) == kInTree) break;
        }
        if (left < 0) break;
        for (u = node;;) {
          const uint32_t lo = M.ldlo(u);
          if ((lo >> 16) == kInTree) break;
          M.stlo(u, lo | 0xffff0000u);
          u = (int)(lo & 0xffffu);
        }
      }
      walk += pops;
      if (left < 0 && !s.misc[M_ERR]) s.misc[M_ERR] = RC_ERR_DISCONNECTED;
    };
    if (__isShared(wst)) {
      // shared-window offsets pinned in registers and explicit ld/st.shared: the loop must not
      // re-derive the window address (a special-register read) in front of every load
      struct Sh {
        uint32_t w, t;
        __device__ __forceinline__ unsigned long long ld(int u) const {
          unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(w + 8u * (uint32_t)u) : "memory"); return v; }
        __device__ __forceinline__ void st(int u, unsigned long long v) const {
          asm volatile("st.shared.u64 [%0], %1;" :: "r"(w + 8u * (uint32_t)u), "l"(v) : "memory"); }
        __device__ __forceinline__ uint32_t ldlo(int u) const {
          uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(w + 8u * (uint32_t)u) : "memory"); return v; }
        __device__ __forceinline__ void stlo(int u, uint32_t v) const {
          asm volatile("st.shared.u32 [%0], %1;" :: "r"(w + 8u * (uint32_t)u), "r"(v) : "memory"); }
        __device__ __forceinline__ uint32_t tab(int i) const {
          uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(t + 2u * (uint32_t)i) : "memory"); return v; }
      } M;
      M.w = (uint32_t)__cvta_generic_to_shared(wst);
      M.t = (uint32_t)__cvta_generic_to_shared(s.wtab);
      asm volatile("" : "+r"(M.w), "+r"(M.t));
      finish(M);
    } else {  // spill mode: the arrays live in this CTA's global scratch
      struct Gl {
        unsigned long long* w; const lid_t* t;
        __device__ __forceinline__ unsigned long long ld(int u) const { return w[u]; }
        __device__ __forceinline__ void st(int u, unsigned long long v) const { w[u] = v; }
        __device__ __forceinline__ uint32_t ldlo(int u) const { return reinterpret_cast<const uint32_t*>(w)[2 * u]; }
        __device__ __forceinline__ void stlo(int u, uint32_t v) const { reinterpret_cast<uint32_t*>(w)[2 * u] = v; }
        __device__ __forceinline__ uint32_t tab(int i) const { return t[i]; }
      } M;
      M.w = wst;
      M.t = s.wtab;
      finish(M);
    }
  }
  RC_PH(PH_W3);
  __syncthreads();
  if (s.misc[M_ERR]) return s.misc[M_ERR];
  }  // !complete
  // the tree: edge (u, successor of u) for every node but the root; the slot of the
  // successor in u's row gives the edge
  for (int u = tid; u < n; u += T) {
    walk += (u != root);  // the first draw of every node
    if (u == root) continue;
    const int nn = complete ? succ_of(u) : (int)(s.wst[u] & 0xffffu);
    int j = -1;
    for (int q = s.loff[u]; q < s.loff[u + 1]; ++q)
      if (s.ladj[q] == nn) { j = s.ladje[q]; break; }
    if (j >= 0) atomicOr(&s.tflag[j >> 5], 1u << (j & 31));  // else: tree_edge_list reports it
  }
  walk = warp_sum(walk);
  if ((tid & 31) == 0 && walk) atomicAdd(&ctr[RC_CTR_WALK], walk);
  __syncthreads();
  return tree_edge_list(c, st);
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
    const int j = s.te[t];
    s.nxt[2 * t] = (lid_t)atomicExch(&s.head[s.eu[j]], (uint32_t)(2 * t));
    s.nxt[2 * t + 1] = (lid_t)atomicExch(&s.head[s.ev[j]], (uint32_t)(2 * t + 1));
  }
  __syncthreads();
  // E3: successor of arc a = the out-arc after reverse(a) around head(a); arc weights
  const int first = (int)s.head[root];
  for (int a = tid; a < L; a += T) {
    const int j = s.te[a >> 1];
    const int x = (a & 1) ? s.eu[j] : s.ev[j];  // node the arc enters
    const uint32_t hx = s.head[x];
    uint32_t nx = s.nxt[a ^ 1];
    if (nx == kNone) nx = hx;
    s.succ[a] = (lid_t)((int)nx == first ? a : (int)nx);  // the arc that closes the tour ends the list
    s.W[a] = ((uint32_t)(a ^ 1) == hx && x != root) ? s.pop[x] : 0;
  }
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
  if (tid == 0) s.ent[root] = 0xffffu;
  __syncthreads();
  // E8: population below every tree edge (cur aliases succ, which is dead now)
  for (int t = tid; t < nt; t += T) {
    const int j = s.te[t];
    const int p0 = s.P[2 * t], p1 = s.P[2 * t + 1];
    const bool down0 = p0 < p1;  // arc 2t (eu -> ev) points away from the root
    const int child = down0 ? s.ev[j] : s.eu[j];
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

__device__ __forceinline__ bool balanced(int tree_pop, int total_pop, double ideal, double thr) {
  // tree.py:412-414, float64 as in the reference
  return fabs(__dsub_rn((double)tree_pop, ideal)) <= thr &&
         fabs(__dsub_rn((double)(total_pop - tree_pop), ideal)) <= thr;
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
    const int gu = c.nodes[s.eu[j]], gv = c.nodes[s.ev[j]];
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

// One recom proposal + validity + commit (tree_proposals.py:36-146, chain.py:186-195).
template <bool kWilson>
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
  for (uint32_t retry = 0;; ++retry) {  // tree_proposals.py:103-138
  if (resel && n_bad >= k * (k - 1) / 2) { res.status = RC_ERR_ALL_PAIRS_BAD; return res; }
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
  for (uint32_t tree_no = t0;; ++tree_no) {
    // tree.py:679-700 attempt accounting (see oracle/recom_oracle.c rco_bipartition)
    const long long attempts = nr > 0 ? (long long)tree_no * nr : (tree_no == t0 ? 0 : max_attempts);
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
    constexpr bool wilson = kWilson;
    const uint32_t* rrank = (rs && !wilson) ? p.rp.weight_rank + (rs->first_tree + tree_no) * (long long)p.g.E : nullptr;
    // ---- P2
    int rounds = 0;
    if (wilson) res.status = wilson_tree(c, st, tree_no, prop_no, rs ? rs->first_tree + tree_no : 0, ctr);
    else res.status = boruvka_mst(c, st, tree_no, prop_no, rrank, &rounds);
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
    auto is_bal = [&](int t) { return balanced(s.cur[t], st.tot, ideal, thr); };
    {
      if (tid == 0) s.misc[M_NBAL] = 0;
      __syncthreads();
      int cnt = 0;
      for (int t = tid; t < nt; t += T) cnt += is_bal(t) ? 1 : 0;
      cnt = warp_sum(cnt);
      if ((tid & 31) == 0 && cnt) atomicAdd(&s.misc[M_NBAL], cnt);
      __syncthreads();
    }
    const int nbal = s.misc[M_NBAL];
    RC_PH(PH_BALANCE);
    res.nbal = nbal;
    if (nbal > 0) {
      const bool last_rec = rs && (int)tree_no == n_rec - 1;
      if (rs && !last_rec) { res.status = RC_ERR_REPLAY; return res; }
      auto internal = [&](int i) { return s.nxt[s.head[i]] != kNone; };  // tree degree > 1
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
      const int policy = kWilson ? RC_CUT_UNIFORM : p.cfg.cut_policy;
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
      if (tid == 0) {
        ctr[RC_CTR_ATTEMPTS] += (int)min(attempts + 1, 0x7fffffffLL);
        if (warn_on && attempts >= p.cfg.warn_attempts) atomicOr(p.st.d_flags + c.chain, RC_FLAG_WARNED);
      }
      break;
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
  RC_PH(PH_CHOOSE);
  // ---- P6: the side holding the reference's root keeps the lower id (tree.py:421, :947-950)
  const int pa = s.P[2 * tstar], pb = s.P[2 * tstar + 1];
  const int pd = min(pa, pb), pu = max(pa, pb);
  auto below = [&](int i) { const int en = s.ent[i]; return en != 0xffff && en >= pd && en <= pu; };
  const bool root_below = below(root);
  const int child_pop = s.cur[tstar];
  const long long pop_a = root_below ? (long long)child_pop : (long long)st.tot - child_pop;
  const long long pop_b = (long long)st.tot - pop_a;
  bool valid = true;
  if (p.cfg.pop_lower <= p.cfg.pop_upper) {  // constraints/bounds.py:25-28
    const double lo = (double)min(pop_a, pop_b), hi = (double)max(pop_a, pop_b);
    valid = p.cfg.pop_lower <= lo && hi <= p.cfg.pop_upper;
  }
  if (!valid) {
    if (tid == 0) ctr[RC_CTR_INVALID] += 1;
    return res;
  }
  // side(i): node i is cut away from the root's side and takes the higher id b
  auto side = [&](int i) { return below(i) != root_below; };
  if (p.cfg.accept_rule == RC_ACCEPT_CUT_EDGE) {
    // accept.py:19-35: the number of cut edges the proposal WOULD have, before anything is
    // written; a rejected proposal leaves the state alone but still counts as a step
    if (tid == 0) s.misc[M_DELTA] = 0;
    for (int i0 = 0; i0 < st.n; i0 += T) {
      const int i = i0 + tid;
      const uint32_t bal = __ballot_sync(0xffffffffu, i < st.n && side(i));
      if (i < st.n && (tid & 31) == 0) s.newb[i >> 5] = bal;
    }
    __syncthreads();
    int d = 0;
    for (int j = tid; j < st.m; j += T) {
      const int u = s.eu[j], v = s.ev[j];
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
    if (!(rc_unit_double(r.v[0], r.v[1]) < (ratio < 1.0 ? ratio : 1.0))) {
      res.accepted = 1;  // step consumed, old state kept
      return res;
    }
  }
  if (tid == 0) s.misc[M_DELTA] = 0;
  for (int i0 = 0; i0 < st.n; i0 += T) {  // partition/assignment.py:76-86
    const int i = i0 + tid;
    const bool sd = i < st.n && side(i);
    const uint32_t bal = __ballot_sync(0xffffffffu, sd);
    if (i < st.n) {
      if ((tid & 31) == 0) s.newb[i >> 5] = bal;
      if (sd != (bool)((s.oldb[i >> 5] >> (i & 31)) & 1u)) c.assign[c.nodes[i]] = (uint8_t)(sd ? st.b : st.a);
    }
  }
  __syncthreads();
  int delta = 0;
  for (int j = tid; j < st.m; j += T) {  // updaters/cut_edges.py:118-120
    const int u = s.eu[j], v = s.ev[j];
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
  RC_PH(PH_COMMIT);
  res.accepted = 1;
  return res;
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
  const int n = st.n, nt = n - 1;
  if (n < 3) { res.status = RC_ERR_NO_ROOT; return res; }
  res.status = wilson_tree(c, st, 0, prop_no, 0, ctr);
  if (tid == 0) { ctr[RC_CTR_TREES] += 1; ctr[RC_CTR_ATTEMPTS] += 1; }
  res.trees = 1;
  if (res.status) return res;
  euler_subtree_pops(c, st, 0);
  const double ideal = p.cfg.pop_target;
  const double thr = __dmul_rn(ideal, p.cfg.epsilon);
  auto is_bal = [&](int t) { return balanced(s.cur[t], st.tot, ideal, thr); };
  if (tid == 0) { s.misc[M_NBAL] = 0; s.misc[M_NINT] = 0; }
  __syncthreads();
  auto internal = [&](int i) { return s.nxt[s.head[i]] != kNone; };
  {
    int cb = 0, ci = 0;
    for (int t = tid; t < nt; t += T) cb += is_bal(t) ? 1 : 0;
    for (int i = tid; i < n; i += T) ci += internal(i) ? 1 : 0;
    cb = warp_sum(cb);
    ci = warp_sum(ci);
    if ((tid & 31) == 0) { if (cb) atomicAdd(&s.misc[M_NBAL], cb); if (ci) atomicAdd(&s.misc[M_NINT], ci); }
  }
  __syncthreads();
  const int nbal = s.misc[M_NBAL], n_int = s.misc[M_NINT];
  res.nbal = nbal;
  if (nbal > M) { res.status = RC_ERR_REVERSIBILITY; return res; }  // tree_proposals.py:205-209
  if (nbal == 0) return res;                                           // self-loop: no balance edge
  const RcPhilox rr = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_ROOT, 0, 0, prop_no, 0);
  const int root = select_kth(c, n, (int)rc_below(rr.v[0], rr.v[1], (uint32_t)n_int), internal);
  const RcPhilox rc_ = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_CUT, 0, 0, prop_no, 0);
  const int tstar = select_kth(c, nt, (int)rc_below(rc_.v[0], rc_.v[1], (uint32_t)nbal), is_bal);
  const int pa = s.P[2 * tstar], pb = s.P[2 * tstar + 1];
  const int pd = min(pa, pb), pu = max(pa, pb);
  auto below = [&](int i) { const int en = s.ent[i]; return en != 0xffff && en >= pd && en <= pu; };
  const bool root_below = below(root);
  const int child_pop = s.cur[tstar];
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
    const int u = s.eu[j], v = s.ev[j];
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
    if (nb != ob) c.assign[c.nodes[i]] = (uint8_t)(nb ? st.b : st.a);
  }
  for (int j = tid; j < st.m; j += T) {
    const int u = s.eu[j], v = s.ev[j];
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

// Free-running mode is a persistent grid over a ring of READY chains (see the loop below):
// every SM stays busy whatever the number of chains and however unevenly the retries fall,
// and a slow step delays only its own chain.  Replay mode (TEST HOOK) is one CTA walking one
// chain through the record.
template <bool kWilson, bool kSpill>
__global__ void __launch_bounds__(512, 2) recom_step_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Chain c;
  c.p = &p;
  smem_layout(&c.s, smem_raw, p.g.N, p.cap_nodes, p.cap_edges, p.g.R, kWilson,
              kSpill ? p.spill + (size_t)blockIdx.x * p.spill_stride : nullptr);
  const int tid = threadIdx.x;
  c.nodes = p.nodes + (size_t)blockIdx.x * p.cap_nodes;
  c.eids = p.eids + (size_t)blockIdx.x * p.cap_edges;
  c.par = 0;
#ifdef RC_TIMERS
  if (tid < 30) c.s.misc[96 + tid] = 0;  // misc[96..127]: per-CTA phase cycles (+ last stamp), flushed at exit
  if (tid == 0) reinterpret_cast<unsigned long long*>(c.s.misc + 96)[15] = (unsigned long long)clock64();
#endif
  int32_t* ctr = c.s.misc + M_CTR;  // [RC_N_COUNTERS]
  if (tid < RC_N_COUNTERS) ctr[tid] = 0;
  if (tid == 0) {
    // Long single-thread sections (Wilson's sequential tail) run in a different warp in each
    // of the CTAs that share an SM: warp w issues from scheduler w % 4, and lone threads of
    // co-resident CTAs that all sit in warp 0 would queue on one scheduler.
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int slot = atomicAdd(p.sm_slots + (smid & 1023u), 1);
    const int n_warps = (int)blockDim.x >> 5;
    c.s.misc[M_WALKER] = 32 * (slot % (n_warps < 4 ? n_warps : 4));
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
  for (;;) {
    unsigned long long t;
    int chain;
    if (replay) {
      t = rt++;
      chain = p.rp.chain;
    } else {
      if (tid == 0) {
        const unsigned long long h = atomicAdd(p.work_counter, 1ull);
        int ch = -1;
        if (h < total) {
          if (h < (unsigned long long)n_chains) {
            ch = (int)h;
          } else {
            int32_t* slot = p.ring + (h - (unsigned long long)n_chains) % cap;
            int v;
            while ((v = ld_acquire(slot)) == 0) __nanosleep(32);
            st_release(slot, 0);  // the slot may be refilled cap pushes later
            ch = v - 1;
          }
        }
        tslot[flip] = ch;
      }
      __syncthreads();
      chain = tslot[flip];
      flip ^= 1;
      RC_PH(PH_TICKET);
      t = chain < 0 ? total : 0;
    }
    if (t >= total) break;
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
          else r = propose<kWilson>(c, replay ? (uint32_t)t : prop_no, rs, ctr);
        } else {
          r = propose<kWilson>(c, replay ? (uint32_t)t : prop_no, rs, ctr);
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
    __threadfence();
    __syncthreads();
    RC_PH(PH_FLUSH);
    if (!replay && tid == 0) {
      const int done = __ldcg(p.progress + chain) + 1;  // this CTA owns the chain until it is pushed back
      p.progress[chain] = done;
      if (done < p.n_steps) {
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
  smem_layout(&c.s, smem_raw, p.g.N, p.cap_nodes, p.cap_edges, p.g.R, false,
              kSpill ? p.spill + (size_t)blockIdx.x * p.spill_stride : nullptr);
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
            if (below == take_below) c.assign[c.nodes[i]] = (uint8_t)part;
          }
          if (tid == 0) c.tally[part] = part_pop;
          if (!(lb_pop <= (double)part_pop && (double)part_pop <= ub_pop)) status = RC_ERR_BALANCE;
          debt = __dadd_rn(debt, __dsub_rn((double)part_pop, pt));
        } else {  // root side -> parts[-2], the rest -> parts[-1] (tree.py:1054-1075)
          for (int i = tid; i < n; i += T)
            c.assign[c.nodes[i]] = (uint8_t)((child_side(i) != root_cs) ? k - 1 : k - 2);
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
  const uint8_t* assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
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
  __shared__ unsigned long long bins[256];
  const int chain = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int k = p.cfg.n_districts, N = p.g.N;
  const uint8_t* assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
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
