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
  uint16_t* g2l;
  int32_t* nodes;
  int32_t* eids;
};

__device__ __forceinline__ bool in_pair(uint8_t x, int a, int b) { return x == a || x == b; }

// ---- weight key of local edge j for tree `tree_no` (tree.py:78-89) ---------------------
__device__ __forceinline__ unsigned long long edge_key(const Chain& c, int j, uint32_t tree_no,
                                                       uint32_t prop_no, const double* rweights) {
  const Params& p = *c.p;
  const int eid = c.eids[j];
  if (rweights) return dbl_key(rweights[eid]);
  double w = rc_edge_unit_weight(p.cfg.seed, c.gchain, (uint32_t)eid, tree_no, prop_no);
  if (p.g.R > 0) {
    const int u = p.g.edge_uv[2 * eid], v = p.g.edge_uv[2 * eid + 1];
    for (int r = 0; r < p.g.R; ++r) {
      const int ru = p.g.region_ids[(size_t)r * p.g.N + u];
      const int rv = p.g.region_ids[(size_t)r * p.g.N + v];
      if (ru != rv || ru < 0 || rv < 0) w = __dadd_rn(w, p.g.surcharge[r]);
    }
  }
  return dbl_key(w);
}

// ---- P0: pick the cut edge (uniform over cut EDGES, tree_proposals.py:106) --------------
// returns edge id (block uniform) or -1
__device__ int pick_cut_edge(const Chain& c, int cut_count, uint32_t prop_no, uint32_t retry) {
  const Params& p = *c.p;
  const int tid = threadIdx.x, T = blockDim.x;
  RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_PAIR, retry, 0, prop_no, 0);
  const int kth = (int)rc_below(r.v[0], r.v[1], (uint32_t)cut_count);
  const int words = (p.g.E + 31) >> 5;
  const int per = (words + T - 1) / T;
  const int lo = min(tid * per, words), hi = min(lo + per, words);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += __popc(__ldcg(c.mask + i));
  int tot;
  const int excl = block_excl_scan(cnt, &tot, c.s.misc);
  if (tid == 0) c.s.misc[M_PICK] = -1;
  __syncthreads();
  if (kth >= excl && kth < excl + cnt) {
    int k = kth - excl;
    for (int i = lo; i < hi; ++i) {
      uint32_t x = __ldcg(c.mask + i);
      const int pc = __popc(x);
      if (k < pc) {
        while (k--) x &= x - 1;
        c.s.misc[M_PICK] = i * 32 + (__ffs(x) - 1);
        break;
      }
      k -= pc;
    }
  }
  __syncthreads();
  return c.s.misc[M_PICK];
}

// ---- P1: gather the merged pair into shared memory ---------------------------------------
// returns RC_OK / RC_ERR_CAPACITY (block uniform)
__device__ int gather_pair(const Chain& c, Step& st, int* d_sub_out) {
  const Params& p = *c.p;
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int a = st.a, b = st.b;
  const int N = p.g.N;
  const int tiles = (N + 15) >> 4;
  if (tid == 0) { s.misc[M_TOT] = 0; s.misc[M_CNT] = 0; }
  __syncthreads();
  int nbase = 0, popsum = 0, slots = 0;
  for (int t0 = 0; t0 < tiles; t0 += T) {
    const int tile = t0 + tid;
    uint32_t fm = 0;
    if (tile < tiles) {
      const uint4 q = __ldcg(reinterpret_cast<const uint4*>(c.assign) + tile);
      const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t x = (wv[k >> 2] >> ((k & 3) * 8)) & 0xffu;
        if ((int)x == a || (int)x == b) fm |= 1u << k;
      }
      const int rem = N - tile * 16;
      if (rem < 16) fm &= (1u << rem) - 1u;
    }
    int tot;
    int local = nbase + block_excl_scan(__popc(fm), &tot, s.misc);
    while (fm) {
      const int k = __ffs(fm) - 1;
      fm &= fm - 1;
      const int g = tile * 16 + k;
      if (local < p.cap_nodes) {
        c.nodes[local] = g;
        c.g2l[g] = (uint16_t)local;
        const int pg = p.g.pop[g];
        s.pop[local] = pg;
        popsum += pg;
        slots += p.g.row_ptr[g + 1] - p.g.row_ptr[g];
      }
      ++local;
    }
    nbase += tot;
  }
  atomicAdd(&s.misc[M_TOT], popsum);
  atomicAdd(&s.misc[M_CNT], slots);
  __syncthreads();  // nodes[], g2l[] (global) and pop[] visible to the block
  st.n = nbase;
  st.tot = s.misc[M_TOT];
  *d_sub_out = s.misc[M_CNT];
  if (st.n > p.cap_nodes || st.n > kMaxLocal) return RC_ERR_CAPACITY;
  // edges: each thread owns a contiguous run of local nodes so that edges come out in
  // ascending (u, v) == ascending edge id.
  const int n = st.n;
  const int per = (n + T - 1) / T;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) {
    const int g = c.nodes[i];
    const int r1 = p.g.row_ptr[g + 1];
    for (int sl = p.g.row_ptr[g]; sl < r1; ++sl) {
      const int h = p.g.col_idx[sl];
      if (h > g && in_pair(c.assign[h], a, b)) ++cnt;
    }
  }
  int m;
  int j = block_excl_scan(cnt, &m, s.misc);
  st.m = m;
  if (m > p.cap_edges || m > 65535) return RC_ERR_CAPACITY;
  for (int i = lo; i < hi; ++i) {
    const int g = c.nodes[i];
    const int r1 = p.g.row_ptr[g + 1];
    for (int sl = p.g.row_ptr[g]; sl < r1; ++sl) {
      const int h = p.g.col_idx[sl];
      if (h > g && in_pair(c.assign[h], a, b)) {
        s.eu[j] = (lid_t)i;
        s.ev[j] = (lid_t)c.g2l[h];
        c.eids[j] = p.g.slot_edge[sl];
        ++j;
      }
    }
  }
  __syncthreads();
  return RC_OK;
}

// ---- P2a: random-weight minimum spanning tree by Boruvka rounds --------------------------
// Unique MST under the total order (weight, local edge index) == networkx Kruskal's tree.
// On return te[0..n-2] lists the tree edges in ascending local edge index.
__device__ int boruvka_mst(const Chain& c, const Step& st, uint32_t tree_no, uint32_t prop_no,
                           const double* rweights, int* rounds_out) {
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, m = st.m;
  for (int j = tid; j < m; j += T) s.khi[j] = (uint32_t)(edge_key(c, j, tree_no, prop_no, rweights) >> 32);
  for (int i = tid; i < n; i += T) s.comp[i] = (lid_t)i;
  for (int i = tid; i < (m + 31) / 32; i += T) s.tflag[i] = 0;
  if (tid == 0) s.misc[M_ERR] = 0;
  int ncomp = n, rounds = 0;
  while (ncomp > 1) {
    ++rounds;
    for (int i = tid; i < n; i += T) { s.best_hi[i] = 0xffffffffu; s.best_lo[i] = ~0ull; }
    __syncthreads();
    // A: minimum high word per component (native 32-bit shared atomics)
    for (int j = tid; j < m; j += T) {
      const int cu = s.comp[s.eu[j]], cv = s.comp[s.ev[j]];
      if (cu != cv) {
        const uint32_t h = s.khi[j];
        atomicMin(&s.best_hi[cu], h);
        atomicMin(&s.best_hi[cv], h);
      }
    }
    __syncthreads();
    // B: among the (almost always single) edges matching the high word, the full key
    for (int j = tid; j < m; j += T) {
      const int cu = s.comp[s.eu[j]], cv = s.comp[s.ev[j]];
      if (cu != cv) {
        const uint32_t h = s.khi[j];
        const bool mu = h == s.best_hi[cu], mv = h == s.best_hi[cv];
        if (mu || mv) {
          const unsigned long long key = edge_key(c, j, tree_no, prop_no, rweights);
          const unsigned long long v = (key << 32) | (unsigned long long)j;
          if (mu) atomic_min_u64_smem(&s.best_lo[cu], v);
          if (mv) atomic_min_u64_smem(&s.best_lo[cv], v);
        }
      }
    }
    __syncthreads();
    // C: hook every component onto the other side of its lightest edge
    for (int i = tid; i < n; i += T) {
      if (s.comp[i] != i) continue;
      if (s.best_hi[i] == 0xffffffffu && s.best_lo[i] == ~0ull) { s.misc[M_ERR] = RC_ERR_DISCONNECTED; s.par[i] = (lid_t)i; continue; }
      const int j = (int)(s.best_lo[i] & 0xffffffffull);
      const int cu = s.comp[s.eu[j]], cv = s.comp[s.ev[j]];
      const int other = cu == i ? cv : cu;
      const bool mutual = s.best_hi[other] == s.best_hi[i] && s.best_lo[other] == s.best_lo[i];
      s.par[i] = (lid_t)((mutual && i < other) ? i : other);
      atomicOr(&s.tflag[j >> 5], 1u << (j & 31));
    }
    __syncthreads();
    if (s.misc[M_ERR]) return s.misc[M_ERR];
    for (int i = tid; i < n; i += T) if (s.comp[i] == i) s.comp[i] = s.par[i];
    __syncthreads();
    // D: flatten (no writes to comp[] while it is being walked)
    for (int i = tid; i < n; i += T) {
      int r = s.comp[i];
      while (s.comp[r] != r) r = s.comp[r];
      s.par[i] = (lid_t)r;
    }
    __syncthreads();
    int roots = 0;
    for (int i = tid; i < n; i += T) { const int r = s.par[i]; s.comp[i] = (lid_t)r; roots += (r == i); }
    if (tid == 0) s.misc[M_NCOMP] = 0;
    __syncthreads();
    if (roots) atomicAdd(&s.misc[M_NCOMP], roots);
    __syncthreads();
    ncomp = s.misc[M_NCOMP];
  }
  *rounds_out = rounds;
  // tree edges in ascending local edge index
  const int per = (m + T - 1) / T;
  const int lo = min(tid * per, m), hi = min(lo + per, m);
  int cnt = 0;
  for (int j = lo; j < hi; ++j) cnt += (s.tflag[j >> 5] >> (j & 31)) & 1;
  int nt;
  int t = block_excl_scan(cnt, &nt, s.misc);
  for (int j = lo; j < hi; ++j)
    if ((s.tflag[j >> 5] >> (j & 31)) & 1) s.te[t++] = (lid_t)j;
  __syncthreads();
  return nt == n - 1 ? RC_OK : RC_ERR_DISCONNECTED;
}

// ---- P3: root the tree and compute every subtree population ----------------------------
// Euler tour of the tree (arcs 2t: eu->ev, 2t+1: ev->eu of tree edge t), list ranking by
// pointer doubling, prefix sum in tour order.  After this call:
//   apos[a]  = tour position of arc a
//   ent[i]   = position of the arc entering node i (0xffff for the root)
//   cur[t]   = population below tree edge t (the side away from the root)
__device__ void euler_subtree_pops(const Chain& c, const Step& st, int root) {
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, nt = n - 1, L = 2 * nt;
  // nxt + ranks
  for (int a = tid; a < L; a += T) {
    const int t = a >> 1, j = s.te[t];
    const int head = (a & 1) ? s.eu[j] : s.ev[j];
    const int r = a ^ 1;
    int pn = s.apos[r] + 1;
    if (pn == s.toff[head + 1]) pn = s.toff[head];
    s.lrA[a] = ((uint32_t)s.tadj[pn] << 16) | 1u;
  }
  __syncthreads();
  if (tid == 0) {
    const int last = s.tadj[s.toff[root + 1] - 1] ^ 1;
    s.lrA[last] = ((uint32_t)last << 16) | 0u;
  }
  __syncthreads();
  uint32_t* A = s.lrA;
  uint32_t* B = s.lrB;
  for (int span = 1; span < L; span <<= 1) {
    for (int a = tid; a < L; a += T) {
      const uint32_t x = A[a];
      const uint32_t y = A[x >> 16];
      B[a] = (y & 0xffff0000u) | ((x & 0xffffu) + (y & 0xffffu));
    }
    __syncthreads();
    uint32_t* tmp = A; A = B; B = tmp;
  }
  // tour positions; B is free and becomes the weight / prefix-sum array
  int32_t* W = reinterpret_cast<int32_t*>(B);
  for (int a = tid; a < L; a += T) s.apos[a] = (lid_t)(L - 1 - (int)(A[a] & 0xffffu));
  if (tid == 0) s.ent[root] = 0xffffu;
  __syncthreads();
  for (int t = tid; t < nt; t += T) {
    const int j = s.te[t];
    const int p0 = s.apos[2 * t], p1 = s.apos[2 * t + 1];
    const bool down0 = p0 < p1;  // arc 2t (eu -> ev) points away from the root
    const int child = down0 ? s.ev[j] : s.eu[j];
    const int pd = down0 ? p0 : p1, pu = down0 ? p1 : p0;
    W[pd] = s.pop[child];
    W[pu] = 0;
    s.ent[child] = (lid_t)pd;
  }
  __syncthreads();
  // inclusive prefix sum of W[0..L)
  {
    const int per = (L + T - 1) / T;
    const int lo = min(tid * per, L), hi = min(lo + per, L);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += W[i];
    int tot;
    int run = block_excl_scan(sum, &tot, s.misc);
    for (int i = lo; i < hi; ++i) { run += W[i]; W[i] = run; }
  }
  __syncthreads();
  for (int t = tid; t < nt; t += T) {
    const int j = s.te[t];
    const int p0 = s.apos[2 * t], p1 = s.apos[2 * t + 1];
    const bool down0 = p0 < p1;
    const int child = down0 ? s.ev[j] : s.eu[j];
    const int pd = down0 ? p0 : p1, pu = down0 ? p1 : p0;
    s.cur[t] = W[pu] - W[pd] + s.pop[child];
  }
  __syncthreads();
}

// tree adjacency (arcs grouped by tail) from te[]; fills toff, tadj, apos(slot)
__device__ void build_tree_adjacency(const Chain& c, const Step& st) {
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int n = st.n, nt = n - 1;
  for (int i = tid; i < n; i += T) s.cur[i] = 0;
  __syncthreads();
  for (int t = tid; t < nt; t += T) {
    const int j = s.te[t];
    atomicAdd(&s.cur[s.eu[j]], 1);
    atomicAdd(&s.cur[s.ev[j]], 1);
  }
  __syncthreads();
  {
    const int per = (n + T - 1) / T;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += s.cur[i];
    int tot;
    int run = block_excl_scan(sum, &tot, s.misc);
    for (int i = lo; i < hi; ++i) { const int d = s.cur[i]; s.toff[i] = run; s.cur[i] = run; run += d; }
    if (tid == T - 1 || hi == n) s.toff[n] = 2 * nt;
  }
  __syncthreads();
  for (int t = tid; t < nt; t += T) {
    const int j = s.te[t];
    const int pu = atomicAdd(&s.cur[s.eu[j]], 1);
    s.tadj[pu] = (lid_t)(2 * t);
    s.apos[2 * t] = (lid_t)pu;
    const int pv = atomicAdd(&s.cur[s.ev[j]], 1);
    s.tadj[pv] = (lid_t)(2 * t + 1);
    s.apos[2 * t + 1] = (lid_t)pv;
  }
  __syncthreads();
}

// k-th element (ascending index) with pred(i) true among [0, count); block uniform result.
template <typename Pred>
__device__ int select_kth(const Chain& c, int count, int kth, Pred pred, int* total_out) {
  const Smem& s = c.s;
  const int tid = threadIdx.x, T = blockDim.x;
  const int per = (count + T - 1) / T;
  const int lo = min(tid * per, count), hi = min(lo + per, count);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += pred(i) ? 1 : 0;
  int tot;
  const int excl = block_excl_scan(cnt, &tot, s.misc);
  if (tid == 0) s.misc[M_PICK] = -1;
  __syncthreads();
  if (kth >= excl && kth < excl + cnt) {
    int k = kth - excl;
    for (int i = lo; i < hi; ++i)
      if (pred(i) && k-- == 0) { s.misc[M_PICK] = i; break; }
  }
  __syncthreads();
  *total_out = tot;
  return s.misc[M_PICK];
}

__device__ __forceinline__ bool balanced(int tree_pop, int total_pop, double ideal, double thr) {
  // tree.py:412-414, float64 as in the reference
  return fabs(__dsub_rn((double)tree_pop, ideal)) <= thr &&
         fabs(__dsub_rn((double)(total_pop - tree_pop), ideal)) <= thr;
}

// ---- P5: region-aware cut choice (tree.py:501-578); returns tree edge index --------------
__device__ int choose_cut_by_weight(const Chain& c, const Step& st, int policy, uint32_t tree_no,
                                    uint32_t prop_no, const double* rweights, double ideal, double thr) {
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
  // which subsets have a suitable cut
  for (int t = tid; t < nt; t += T) {
    if (!balanced(s.cur[t], st.tot, ideal, thr)) continue;
    const int j = s.te[t];
    const int gu = c.nodes[s.eu[j]], gv = c.nodes[s.ev[j]];
    int cm = 0;
    for (int r = 0; r < R && r < 3; ++r)
      if (p.g.region_ids[(size_t)r * p.g.N + gu] != p.g.region_ids[(size_t)r * p.g.N + gv]) cm |= 1 << r;
    int has = 0;
    for (int i = 0; i < ns; ++i) if ((cm & subsets[i]) == subsets[i]) has |= 1 << i;
    if (has) atomicOr(&s.misc[M_SUBSET], has);
  }
  __syncthreads();
  const int has = s.misc[M_SUBSET];
  const int need = has ? subsets[__ffs(has) - 1] : 0;
  for (int t = tid; t < nt; t += T) {
    if (!balanced(s.cur[t], st.tot, ideal, thr)) continue;
    const int j = s.te[t];
    if (need) {
      const int gu = c.nodes[s.eu[j]], gv = c.nodes[s.ev[j]];
      int cm = 0;
      for (int r = 0; r < R && r < 3; ++r)
        if (p.g.region_ids[(size_t)r * p.g.N + gu] != p.g.region_ids[(size_t)r * p.g.N + gv]) cm |= 1 << r;
      if ((cm & need) != need) continue;
    }
    // max weight, ties -> lowest edge index.  Two rounds keep it exact with one 64-bit CAS:
    atomic_max_u64_smem(best, edge_key(c, j, tree_no, prop_no, rweights));
  }
  __syncthreads();
  const unsigned long long bw = *best;
  if (tid == 0) s.misc[M_TSTAR] = 0x7fffffff;
  __syncthreads();
  for (int t = tid; t < nt; t += T) {
    if (!balanced(s.cur[t], st.tot, ideal, thr)) continue;
    const int j = s.te[t];
    if (need) {
      const int gu = c.nodes[s.eu[j]], gv = c.nodes[s.ev[j]];
      int cm = 0;
      for (int r = 0; r < R && r < 3; ++r)
        if (p.g.region_ids[(size_t)r * p.g.N + gu] != p.g.region_ids[(size_t)r * p.g.N + gv]) cm |= 1 << r;
      if ((cm & need) != need) continue;
    }
    if (edge_key(c, j, tree_no, prop_no, rweights) == bw) atomicMin(&s.misc[M_TSTAR], t);
  }
  __syncthreads();
  return s.misc[M_TSTAR];
}

struct ProposalResult {
  int status;     // RC_OK or error
  int accepted;   // committed
  int trees;      // trees drawn
  int nbal;       // balanced cuts of the last tree
};

// One recom proposal + validity + commit (tree_proposals.py:36-146, chain.py:186-195).
__device__ ProposalResult propose(const Chain& c, uint32_t prop_no, const RcReplayStep* rs,
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
  int tstar = -1;
  for (uint32_t retry = 0;; ++retry) {  // tree_proposals.py:103-138
  if (resel && n_bad >= k * (k - 1) / 2) { res.status = RC_ERR_ALL_PAIRS_BAD; return res; }
  // ---- P0
  int e;
  if (rs) {
    e = rs->pair_edge;
    if (e < 0 || e >= p.g.E || !((__ldcg(c.mask + (e >> 5)) >> (e & 31)) & 1)) { res.status = RC_ERR_REPLAY; return res; }
  } else {
    e = pick_cut_edge(c, cut_count, prop_no, retry);
    if (e < 0) { res.status = RC_ERR_NO_CUT_EDGES; return res; }
  }
  {
    const int la = c.assign[p.g.edge_uv[2 * e]], lb = c.assign[p.g.edge_uv[2 * e + 1]];
    st.a = min(la, lb); st.b = max(la, lb);  // tree_proposals.py:113
  }
  if (resel && ((bad[(st.a * 256 + st.b) >> 5] >> ((st.a * 256 + st.b) & 31)) & 1)) continue;
  // ---- P1
  int d_sub = 0;
  res.status = gather_pair(c, st, &d_sub);
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
    const double* rweights = rs ? p.rp.weights + (rs->first_tree + tree_no) * (long long)p.g.E : nullptr;
    // ---- P2
    int rounds = 0;
    res.status = boruvka_mst(c, st, tree_no, prop_no, rweights, &rounds);
    if (tid == 0) { ctr[RC_CTR_TREES] += 1; ctr[RC_CTR_ROUNDS] += rounds; }
    res.trees = (int)tree_no + 1;
    if (res.status) return res;
    build_tree_adjacency(c, st);
    // ---- root (tree.py:377): uniform over tree nodes of degree > 1
    int root;
    const bool last_rec = rs && (int)tree_no == n_rec - 1;
    if (last_rec) {
      const int gr = rs->root;
      root = (gr >= 0 && gr < p.g.N && in_pair(c.assign[gr], st.a, st.b)) ? (int)c.g2l[gr] : -1;
      if (root < 0 || s.toff[root + 1] - s.toff[root] < 2) { res.status = RC_ERR_REPLAY; return res; }
    } else {
      RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_ROOT, 0, tree_no, prop_no, 0);
      int n_int = 0;
      auto internal = [&](int i) { return s.toff[i + 1] - s.toff[i] > 1; };
      // two passes: count, then select (the draw needs the count)
      (void)select_kth(c, n, -1, internal, &n_int);
      if (n_int == 0) { res.status = RC_ERR_NO_ROOT; return res; }
      root = select_kth(c, n, (int)rc_below(r.v[0], r.v[1], (uint32_t)n_int), internal, &n_int);
    }
    // ---- P3
    euler_subtree_pops(c, st, root);
    // ---- P4: balanced tree edges (cur[t] = population below edge t)
    auto is_bal = [&](int t) { return balanced(s.cur[t], st.tot, ideal, thr); };
    int nbal = 0;
    (void)select_kth(c, nt, -1, is_bal, &nbal);
    res.nbal = nbal;
    if (nbal > 0) {
      if (rs && !last_rec) { res.status = RC_ERR_REPLAY; return res; }
      // ---- P5
      const int policy = p.cfg.tree_kind == RC_TREE_WILSON ? RC_CUT_UNIFORM : p.cfg.cut_policy;
      if (rs && rs->cut_edge >= 0) {
        const int want = rs->cut_edge;
        tstar = select_kth(c, nt, 0, [&](int t) { return c.eids[s.te[t]] == want && is_bal(t); }, &nbal);
        nbal = res.nbal;
        if (tstar < 0) { res.status = RC_ERR_REPLAY; return res; }
      } else if (policy == RC_CUT_UNIFORM) {
        RcPhilox r = rc_draw(p.cfg.seed, c.gchain, RC_DRAW_CUT, 0, tree_no, prop_no, 0);
        tstar = select_kth(c, nt, (int)rc_below(r.v[0], r.v[1], (uint32_t)nbal), is_bal, &nbal);
      } else {
        tstar = choose_cut_by_weight(c, st, policy, tree_no, prop_no, rweights, ideal, thr);
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
  // ---- P6: the side holding the root keeps the lower id (tree.py:421, :947-950)
  const int pa = s.apos[2 * tstar], pb = s.apos[2 * tstar + 1];
  const int pd = min(pa, pb), pu = max(pa, pb);
  const int child_pop = s.cur[tstar];
  const long long pop_b = child_pop, pop_a = (long long)st.tot - child_pop;
  bool valid = true;
  if (p.cfg.pop_lower <= p.cfg.pop_upper) {  // constraints/bounds.py:25-28
    const double lo = (double)min(pop_a, pop_b), hi = (double)max(pop_a, pop_b);
    valid = p.cfg.pop_lower <= lo && hi <= p.cfg.pop_upper;
  }
  if (!valid) {
    if (tid == 0) ctr[RC_CTR_INVALID] += 1;
    return res;
  }
  auto side = [&](int i) { const int en = s.ent[i]; return en != 0xffff && en >= pd && en <= pu; };
  if (tid == 0) s.misc[M_DELTA] = 0;
  __syncthreads();
  int delta = 0;
  for (int j = tid; j < st.m; j += T) {
    const bool now = side(s.eu[j]) != side(s.ev[j]);
    const int eid = c.eids[j];
    const uint32_t bit = 1u << (eid & 31);
    const uint32_t old = now ? atomicOr(c.mask + (eid >> 5), bit) : atomicAnd(c.mask + (eid >> 5), ~bit);
    delta += (int)now - (int)((old & bit) != 0);
  }
  if (delta) atomicAdd(&s.misc[M_DELTA], delta);
  for (int i = tid; i < st.n; i += T) c.assign[c.nodes[i]] = (uint8_t)(side(i) ? st.b : st.a);
  __syncthreads();
  if (tid == 0) {
    c.tally[st.a] = pop_a;  // updaters/tally.py:162-165
    c.tally[st.b] = pop_b;
    p.st.d_cut_count[c.chain] = cut_count + s.misc[M_DELTA];
    if (p.st.d_last_pair) { p.st.d_last_pair[2 * c.chain] = st.a; p.st.d_last_pair[2 * c.chain + 1] = st.b; }
  }
  __syncthreads();
  res.accepted = 1;
  return res;
}

// =========================================================================================
__global__ void __launch_bounds__(512) recom_step_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Chain c;
  c.p = &p;
  smem_layout(&c.s, smem_raw, p.cap_nodes, p.cap_edges);
  const int tid = threadIdx.x;
  const bool replay = p.rp.n_steps > 0;
  c.chain = replay ? p.rp.chain : (int)blockIdx.x;
  c.gchain = (uint32_t)(c.chain + p.cfg.chain_offset);
  c.assign = p.st.d_assignment + (size_t)c.chain * p.st.assign_stride;
  c.mask = p.st.d_cut_mask + (size_t)c.chain * p.st.mask_stride;
  c.tally = p.st.d_tally + (size_t)c.chain * p.cfg.n_districts;
  c.g2l = p.g2l + (size_t)blockIdx.x * p.g2l_stride;
  c.nodes = p.nodes + (size_t)blockIdx.x * p.cap_nodes;
  c.eids = p.eids + (size_t)blockIdx.x * p.cap_edges;
  int32_t* ctr = c.s.misc + M_CTR;  // [RC_N_COUNTERS]
  if (tid < RC_N_COUNTERS) ctr[tid] = 0;
  __syncthreads();
  if (p.st.d_status[c.chain] != RC_OK) return;
  uint32_t prop_no = p.st.d_proposal_no[c.chain];
  const int streak_cap = p.cfg.max_invalid_streak > 0 ? p.cfg.max_invalid_streak : 1000;
  const long long n_steps = replay ? p.rp.n_steps : p.n_steps;
  int status = RC_OK;
  long long done = 0;
  for (long long sidx = 0; sidx < n_steps && status == RC_OK; ++sidx) {
    const RcReplayStep* rs = replay ? p.rp.steps + sidx : nullptr;
    ProposalResult r = {RC_OK, 0, 0, 0};
    int streak = 0;
    for (; streak < streak_cap; ++streak) {
      r = propose(c, replay ? (uint32_t)sidx : prop_no, rs, ctr);
      ++prop_no;
      if (r.status || r.accepted || replay) break;
    }
    status = r.status;
    if (status == RC_OK && !r.accepted) status = RC_ERR_INVALID_STREAK;
    if (status == RC_OK) ++done;
    __syncthreads();
    if (replay) {
      if (p.rp.info_trace && tid == 0) {
        int32_t* inf = p.rp.info_trace + 4 * sidx;
        inf[0] = r.trees; inf[1] = r.nbal; inf[2] = p.st.d_cut_count[c.chain]; inf[3] = status;
      }
      if (p.rp.assign_trace)
        for (int i = tid; i < p.g.N; i += blockDim.x) p.rp.assign_trace[(size_t)sidx * p.g.N + i] = c.assign[i];
    }
    __syncthreads();
  }
  if (tid == 0) {
    p.st.d_proposal_no[c.chain] = prop_no;
    p.st.d_step_no[c.chain] += done;
    if (status != RC_OK) p.st.d_status[c.chain] = status;
    for (int i = 0; i < RC_N_COUNTERS; ++i) p.st.d_counters[(size_t)c.chain * RC_N_COUNTERS + i] += ctr[i];
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
  if (__syncthreads_or(bad) && tid == 0) p.st.d_status[chain] = RC_ERR_INVALID_ARG;
  else if (tid == 0) p.st.d_status[chain] = RC_OK;
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
