/*
 * recom_oracle.c - CPU restatement of GerryChain's ReCom proposal step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gerrychain_b200/, the CUDA
 * library) may call, link or import this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do, and there only as the checker
 * or the timed CPU baseline.
 *
 * It follows the REFERENCE's algorithm, sequentially, function by function
 * (paths relative to /root/reference/gerrychain/):
 *   rco_pick_pair        proposals/tree_proposals.py:103-120   recom: cut edge -> pair -> subgraph
 *   rco_random_weights   tree.py:78-89                        random_spanning_tree weight loop
 *   rco_kruskal          tree.py:91-93 + networkx/algorithms/tree/mst.py kruskal_mst_edges
 *                        (sort by weight, union-find)
 *   rco_wilson           tree.py:112-132                      uniform_spanning_tree
 *   rco_bfs              tree.py:52-57, :378-379              predecessors / successors
 *   rco_subtree_pops     tree.py:289-322                      _calc_pops
 *   rco_balanced_cuts    tree.py:411-423                      two-sided epsilon test
 *   rco_choose_cut       tree.py:501-578, :446-494            cut choice policies
 *   rco_bipartition      tree.py:670-718                      retry loop of bipartition_tree
 *   rco_apply            tree.py:947-964, partition/assignment.py:76-86,
 *                        updaters/tally.py:162-165, updaters/cut_edges.py:118-120
 *   rco_step             chain.py:185-195 + constraints/bounds.py:25-28 + accept.py:15
 *   rco_init_state       updaters/tally.py:118-141, updaters/cut_edges.py:110-114
 *
 * Randomness: replay mode consumes the reference's own recorded draws (by identity);
 * free-running mode uses the counter-based schedule of include/recom_rng.h.  Canonical
 * orders wherever the reference's order is an artefact of Python set/dict iteration:
 * cut edges, balanced cuts -> ascending edge id; candidate roots, nodes -> ascending
 * node index.
 *
 * Pinned against (tests/test_oracle_vs_reference.py): the reference's known-answer trees
 * (tests/test_tree.py:354-385) and recorded reference steps (tests/golden/replay_*.npz,
 * written by oracle/record_reference.py from the unmodified reference).
 */
#include "recom_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/recom_rng.h"

struct RcoWork {
  int32_t N, E;
  int32_t *g2l;                 /* [N] global -> local node, -1 outside the pair      */
  int32_t *nodes;               /* [N] local -> global                                 */
  int32_t *eu, *ev, *eid;       /* [E] subgraph edges (local endpoints, global id)     */
  double *w;                    /* [E] random_weight per subgraph edge                 */
  int32_t *order;               /* [E] Kruskal sort permutation                        */
  int32_t *uf;                  /* [N] union-find parent                               */
  int32_t *te;                  /* [N] tree edges (index into eu/ev/eid), or Wilson    */
  int32_t *tu, *tv;             /* [N] tree edge endpoints (local)                     */
  int32_t *tw_eid;              /* [N] global edge id of tree edge                     */
  double *tw;                   /* [N] weight of tree edge (NaN for Wilson)            */
  int32_t *tdeg, *toff, *tadj, *tadj_t; /* tree adjacency (CSR over local nodes)       */
  int32_t *parent, *parent_t, *bfs;
  int64_t *subpop;
  int32_t *bal_t;               /* balanced tree-edge indices                          */
  int32_t *bal_child;           /* child node (subtree side) per balanced cut          */
  int32_t *next;                /* Wilson successor                                    */
  uint8_t *in_tree;
  uint32_t *draws;              /* Wilson per-node draw counter                        */
  uint8_t *side;                /* [N] 1 = subtree (non-root) side of the chosen cut   */
  uint64_t *bad_pairs;          /* k*k bitset for pair reselection                     */
  int32_t n, m, nt;
  int64_t totpop;
};

static void* xcalloc(size_t n, size_t s) {
  void* p = calloc(n ? n : 1, s);
  if (!p) { fprintf(stderr, "recom_oracle: out of memory\n"); abort(); }
  return p;
}

RcoWork* rco_work_create(const RcGraph* g) {
  RcoWork* w = (RcoWork*)xcalloc(1, sizeof(RcoWork));
  int32_t N = g->n_nodes, E = g->n_edges;
  w->N = N; w->E = E;
  w->g2l = (int32_t*)xcalloc(N, 4);
  for (int i = 0; i < N; ++i) w->g2l[i] = -1;
  w->nodes = (int32_t*)xcalloc(N, 4);
  w->eu = (int32_t*)xcalloc(E, 4); w->ev = (int32_t*)xcalloc(E, 4);
  w->eid = (int32_t*)xcalloc(E, 4);
  w->w = (double*)xcalloc(E, 8);
  w->order = (int32_t*)xcalloc(E, 4);
  w->uf = (int32_t*)xcalloc(N, 4);
  w->te = (int32_t*)xcalloc(N, 4);
  w->tu = (int32_t*)xcalloc(N, 4); w->tv = (int32_t*)xcalloc(N, 4);
  w->tw_eid = (int32_t*)xcalloc(N, 4);
  w->tw = (double*)xcalloc(N, 8);
  w->tdeg = (int32_t*)xcalloc(N + 1, 4); w->toff = (int32_t*)xcalloc(N + 1, 4);
  w->tadj = (int32_t*)xcalloc(2 * (size_t)N, 4);
  w->tadj_t = (int32_t*)xcalloc(2 * (size_t)N, 4);
  w->parent = (int32_t*)xcalloc(N, 4); w->parent_t = (int32_t*)xcalloc(N, 4);
  w->bfs = (int32_t*)xcalloc(N, 4);
  w->subpop = (int64_t*)xcalloc(N, 8);
  w->bal_t = (int32_t*)xcalloc(N, 4); w->bal_child = (int32_t*)xcalloc(N, 4);
  w->next = (int32_t*)xcalloc(N, 4);
  w->in_tree = (uint8_t*)xcalloc(N, 1);
  w->draws = (uint32_t*)xcalloc(N, 4);
  w->side = (uint8_t*)xcalloc(N, 1);
  w->bad_pairs = (uint64_t*)xcalloc(1024, 8); /* 256*256 bits */
  return w;
}

void rco_work_destroy(RcoWork* w) {
  if (!w) return;
  free(w->g2l); free(w->nodes); free(w->eu); free(w->ev); free(w->eid); free(w->w);
  free(w->order); free(w->uf); free(w->te); free(w->tu); free(w->tv); free(w->tw_eid);
  free(w->tw); free(w->tdeg); free(w->toff); free(w->tadj); free(w->tadj_t);
  free(w->parent); free(w->parent_t); free(w->bfs); free(w->subpop); free(w->bal_t);
  free(w->bal_child); free(w->next); free(w->in_tree); free(w->draws); free(w->side);
  free(w->bad_pairs); free(w);
}

/* ------------------------------------------------------------------------------------ */
/* updaters/tally.py:118-141 + updaters/cut_edges.py:110-114: first-time scans.         */
int rco_init_state(const RcGraph* g, const RcConfig* c, const rc_label_t* assign,
                   int64_t* tally, uint32_t* mask, int32_t* cut_count) {
  int k = c->n_districts;
  for (int d = 0; d < k; ++d) tally[d] = 0;
  for (int v = 0; v < g->n_nodes; ++v) {
    if (assign[v] >= k) return RC_ERR_INVALID_ARG;
    tally[assign[v]] += g->pop[v];
  }
  int words = (g->n_edges + 31) / 32;
  memset(mask, 0, (size_t)words * 4);
  int cnt = 0;
  for (int e = 0; e < g->n_edges; ++e) {
    int u = g->edge_uv[2 * e], v = g->edge_uv[2 * e + 1];
    if (assign[u] != assign[v]) { mask[e >> 5] |= 1u << (e & 31); ++cnt; }
  }
  *cut_count = cnt;
  return RC_OK;
}

/* k-th set bit of the cut mask = k-th cut edge in ascending edge id. */
static int32_t kth_cut_edge(const uint32_t* mask, int words, uint32_t kth) {
  for (int i = 0; i < words; ++i) {
    uint32_t x = mask[i];
    uint32_t pc = (uint32_t)__builtin_popcount(x);
    if (kth < pc) {
      for (;;) {
        int b = __builtin_ctz(x);
        if (kth == 0) return i * 32 + b;
        x &= x - 1; --kth;
      }
    }
    kth -= pc;
  }
  return -1;
}

/* proposals/tree_proposals.py:118-120 + tree.py:670: induced subgraph of the pair.      */
static void rco_gather(const RcGraph* g, RcoWork* w, const rc_label_t* assign, int a, int b) {
  int n = 0;
  int64_t tot = 0;
  for (int v = 0; v < g->n_nodes; ++v)
    if (assign[v] == a || assign[v] == b) {
      w->g2l[v] = n; w->nodes[n++] = v; tot += g->pop[v];
    }
  int m = 0;
  for (int i = 0; i < n; ++i) {
    int u = w->nodes[i];
    for (int s = g->row_ptr[u]; s < g->row_ptr[u + 1]; ++s) {
      int v = g->col_idx[s];
      if (v > u && w->g2l[v] >= 0) {
        w->eu[m] = i; w->ev[m] = w->g2l[v]; w->eid[m] = g->slot_edge[s]; ++m;
      }
    }
  }
  w->n = n; w->m = m; w->totpop = tot;
}

static void rco_ungather(RcoWork* w) {
  for (int i = 0; i < w->n; ++i) w->g2l[w->nodes[i]] = -1;
}

/* tree.py:78-89: weight = random.random() + surcharges for region-crossing edges, in the
 * 32-bit fixed point of include/recom_rng.h (only the order of the weights matters). */
static uint64_t surcharge_fix(const RcGraph* g, int u, int v) {
  uint64_t add = 0;
  for (int r = 0; r < g->n_region_cols; ++r) {
    int ru = g->region_ids[(size_t)r * g->n_nodes + u];
    int rv = g->region_ids[(size_t)r * g->n_nodes + v];
    if (ru != rv || ru < 0 || rv < 0) add += rc_surcharge_fix(g->surcharges[r]);
  }
  return add;
}

static void rco_random_weights(const RcGraph* g, const RcConfig* c, RcoWork* w,
                               uint32_t chain, uint32_t tree_no, uint32_t prop_no) {
  uint64_t tot = 0;
  for (int r = 0; r < g->n_region_cols; ++r) tot += rc_surcharge_fix(g->surcharges[r]);
  uint64_t mult = rc_key_mult(tot, g->n_region_cols);
  for (int j = 0; j < w->m; ++j) {   /* j-th edge in ascending edge id (rco_gather order) */
    uint32_t raw = rc_edge_raw(c->seed, chain, (uint32_t)j, tree_no, prop_no);
    uint64_t add = mult ? surcharge_fix(g, w->nodes[w->eu[j]], w->nodes[w->ev[j]]) : 0;
    w->w[j] = (double)rc_key_finish(raw, add, mult);
  }
}

static __thread RcoWork* t_sort_ctx; /* qsort has no context argument */
static int cmp_edges(const void* pa, const void* pb) {
  int a = *(const int32_t*)pa, b = *(const int32_t*)pb;
  const RcoWork* w = t_sort_ctx;
  if (w->w[a] < w->w[b]) return -1;
  if (w->w[a] > w->w[b]) return 1;
  return (w->eid[a] > w->eid[b]) - (w->eid[a] < w->eid[b]);
}

static int uf_find(int32_t* uf, int x) {
  while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; }
  return x;
}

/* networkx kruskal_mst_edges: edges sorted by weight, joined through a union-find.       */
static int rco_kruskal(RcoWork* w) {
  for (int j = 0; j < w->m; ++j) w->order[j] = j;
  t_sort_ctx = w;
  qsort(w->order, (size_t)w->m, sizeof(int32_t), cmp_edges);
  for (int i = 0; i < w->n; ++i) w->uf[i] = i;
  int nt = 0;
  for (int j = 0; j < w->m && nt < w->n - 1; ++j) {
    int e = w->order[j];
    int ru = uf_find(w->uf, w->eu[e]), rv = uf_find(w->uf, w->ev[e]);
    if (ru != rv) {
      w->uf[ru] = rv;
      w->tu[nt] = w->eu[e]; w->tv[nt] = w->ev[e];
      w->tw_eid[nt] = w->eid[e]; w->tw[nt] = w->w[e];
      ++nt;
    }
  }
  w->nt = nt;
  return nt == w->n - 1 ? RC_OK : RC_ERR_DISCONNECTED;
}

/* Neighbours of local node i inside the pair, ascending node index. */
static int sub_neighbors(const RcGraph* g, const RcoWork* w, int i, int32_t* out_local,
                         int32_t* out_eid) {
  int u = w->nodes[i], d = 0;
  for (int s = g->row_ptr[u]; s < g->row_ptr[u + 1]; ++s) {
    int l = w->g2l[g->col_idx[s]];
    if (l >= 0) { out_local[d] = l; if (out_eid) out_eid[d] = g->slot_edge[s]; ++d; }
  }
  return d;
}

/* tree.py:112-132 uniform_spanning_tree (Wilson).  Free-running: successor draw t of the
 * tree (recom_rng.h RC_DRAW_WALK), start nodes in ascending node index. */
static int rco_wilson(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t chain,
                      uint32_t tree_no, uint32_t prop_no, const RcReplay* rp, int64_t tree_idx,
                      int64_t* walk_draws) {
  int n = w->n;
  int32_t nb[512], nbe[512];
  int root;
  if (rp) {
    root = w->g2l[rp->wilson_root[tree_idx]];
    if (root < 0) return RC_ERR_REPLAY;
  } else {
    RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_WROOT, 0, tree_no, prop_no, 0);
    root = (int)rc_below(p.v[0], p.v[1], (uint32_t)n);
  }
  memset(w->in_tree, 0, (size_t)n);
  memset(w->draws, 0, (size_t)n * 4);
  w->in_tree[root] = 1; w->next[root] = -1;
  uint32_t t = 0;
  for (int node = 0; node < n; ++node) {
    int u = node;
    while (!w->in_tree[u]) {
      uint32_t j = w->draws[u]++;
      int v;
      if (rp) {
        const int64_t* sp = rp->stack_ptr + tree_idx * ((int64_t)g->n_nodes + 1);
        int gu = w->nodes[u];
        if (sp[gu] + j >= sp[gu + 1]) return RC_ERR_REPLAY;
        v = w->g2l[rp->stack_val[sp[gu] + j]];
        if (v < 0) return RC_ERR_REPLAY;
      } else {
        int d = sub_neighbors(g, w, u, nb, NULL);
        if (d > 512) return RC_ERR_CAPACITY;
        RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_WALK, t >> 2, tree_no, prop_no, 0);
        v = nb[rc_below32(p.v[t & 3], (uint32_t)d)];
        ++t;
      }
      ++*walk_draws;
      w->next[u] = v;
      u = v;
    }
    u = node;
    while (!w->in_tree[u]) { w->in_tree[u] = 1; u = w->next[u]; }
  }
  int nt = 0;
  for (int i = 0; i < n; ++i)
    if (w->next[i] >= 0) {
      int d = sub_neighbors(g, w, i, nb, nbe), eid = -1;
      for (int q = 0; q < d; ++q) if (nb[q] == w->next[i]) eid = nbe[q];
      if (eid < 0) return RC_ERR_REPLAY;
      w->tu[nt] = i; w->tv[nt] = w->next[i]; w->tw_eid[nt] = eid; w->tw[nt] = NAN;
      ++nt;
    }
  w->nt = nt;
  return nt == n - 1 ? RC_OK : RC_ERR_DISCONNECTED;
}

static void build_tree_adj(RcoWork* w) {
  int n = w->n;
  for (int i = 0; i <= n; ++i) w->tdeg[i] = 0;
  for (int t = 0; t < w->nt; ++t) { w->tdeg[w->tu[t]]++; w->tdeg[w->tv[t]]++; }
  w->toff[0] = 0;
  for (int i = 0; i < n; ++i) w->toff[i + 1] = w->toff[i] + w->tdeg[i];
  for (int i = 0; i < n; ++i) w->tdeg[i] = 0;
  for (int t = 0; t < w->nt; ++t) {
    int u = w->tu[t], v = w->tv[t];
    w->tadj[w->toff[u] + w->tdeg[u]] = v; w->tadj_t[w->toff[u] + w->tdeg[u]++] = t;
    w->tadj[w->toff[v] + w->tdeg[v]] = u; w->tadj_t[w->toff[v] + w->tdeg[v]++] = t;
  }
}

/* tree.py:52-57: BFS predecessors (parent per node) from the root.                      */
static void rco_bfs(RcoWork* w, int root) {
  int n = w->n, head = 0, tail = 0;
  for (int i = 0; i < n; ++i) w->parent[i] = -2;
  w->parent[root] = -1; w->parent_t[root] = -1;
  w->bfs[tail++] = root;
  while (head < tail) {
    int u = w->bfs[head++];
    for (int s = w->toff[u]; s < w->toff[u + 1]; ++s) {
      int v = w->tadj[s];
      if (w->parent[v] == -2) { w->parent[v] = u; w->parent_t[v] = w->tadj_t[s]; w->bfs[tail++] = v; }
    }
  }
}

/* tree.py:289-322 _calc_pops: population of the subtree below every node.               */
static void rco_subtree_pops(const RcGraph* g, RcoWork* w) {
  int n = w->n;
  for (int i = 0; i < n; ++i) w->subpop[i] = g->pop[w->nodes[i]];
  for (int q = n - 1; q > 0; --q) {
    int v = w->bfs[q];
    w->subpop[w->parent[v]] += w->subpop[v];
  }
}

/* tree.py:411-423: two-sided test, evaluated in float64 exactly as the reference.       */
static int is_balanced(int64_t tree_pop, int64_t total_pop, double ideal, double eps) {
  return fabs((double)tree_pop - ideal) <= ideal * eps &&
         fabs((double)(total_pop - tree_pop) - ideal) <= ideal * eps;
}

static int cmp_bal(const void* pa, const void* pb) {
  int a = *(const int32_t*)pa, b = *(const int32_t*)pb;
  const RcoWork* w = t_sort_ctx;
  return (w->tw_eid[a] > w->tw_eid[b]) - (w->tw_eid[a] < w->tw_eid[b]);
}

static int rco_balanced_cuts(RcoWork* w, double ideal, double eps, int root) {
  int nb = 0;
  for (int v = 0; v < w->n; ++v) {
    if (v == root) continue;
    if (is_balanced(w->subpop[v], w->totpop, ideal, eps)) w->bal_t[nb++] = w->parent_t[v];
  }
  t_sort_ctx = w;
  qsort(w->bal_t, (size_t)nb, sizeof(int32_t), cmp_bal); /* canonical: ascending edge id */
  for (int i = 0; i < nb; ++i) {
    int t = w->bal_t[i];
    w->bal_child[i] = (w->parent[w->tu[t]] == w->tv[t] && w->parent_t[w->tu[t]] == t)
                          ? w->tu[t] : w->tv[t];
  }
  return nb;
}

/* tree.py:552-561: a cut "crosses" region column r when its endpoints' labels differ
 * (None == None there, unlike the weight loop). */
static int cut_crosses(const RcGraph* g, const RcoWork* w, int t, int r) {
  int ru = g->region_ids[(size_t)r * g->n_nodes + w->nodes[w->tu[t]]];
  int rv = g->region_ids[(size_t)r * g->n_nodes + w->nodes[w->tv[t]]];
  return ru != rv;
}

static int max_weight(const RcoWork* w, const int* cand, int nc) {
  int best = cand[0];
  for (int i = 1; i < nc; ++i)
    if (w->tw[w->bal_t[cand[i]]] > w->tw[w->bal_t[best]]) best = cand[i];
  return best;
}

/* tree.py:501-578 _region_preferred_max_weight_choice / :446-480 _max_weight_choice /
 * :540-545 random.choice.  Returns an index into the balanced list. */
static int rco_choose_cut(const RcGraph* g, const RcConfig* c, RcoWork* w, int nb,
                          uint32_t chain, uint32_t tree_no, uint32_t prop_no) {
  int policy = c->cut_policy;
  if (c->tree_kind == RC_TREE_WILSON) policy = RC_CUT_UNIFORM; /* no random_weight, no
      node attrs on the Wilson tree: every branch degenerates to a uniform pick */
  if (policy == RC_CUT_UNIFORM) {
    RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_CUT, 0, tree_no, prop_no, 0);
    return (int)rc_below(p.v[0], p.v[1], (uint32_t)nb);
  }
  int* cand = (int*)w->order; /* reuse: nb <= n <= E entries */
  int R = g->n_region_cols;
  if (policy == RC_CUT_REGION_PREF && R >= 1 && R <= 3) {
    /* tree.py:483-494: subsets by (size, summed surcharge) descending, stable. */
    int subsets[7], ns = 0;
    for (int size = 1; size <= R; ++size)
      for (int s = 1; s < (1 << R); ++s)
        if (__builtin_popcount((unsigned)s) == size) subsets[ns++] = s;
    /* for R <= 3 ascending bitmasks within a size == itertools.combinations order */
    double sums[7];
    for (int i = 0; i < ns; ++i) {
      double s = 0.0;
      for (int r = 0; r < R; ++r) if (subsets[i] >> r & 1) s += g->surcharges[r];
      sums[i] = s;
    }
    /* stable sort, descending by (size, sum): insertion sort keeps ties in order */
    for (int i = 1; i < ns; ++i) {
      int s = subsets[i]; double v = sums[i]; int j = i - 1;
      while (j >= 0) {
        int cj = __builtin_popcount((unsigned)subsets[j]), cs = __builtin_popcount((unsigned)s);
        int less = cj < cs || (cj == cs && sums[j] < v);
        if (!less) break;
        subsets[j + 1] = subsets[j]; sums[j + 1] = sums[j]; --j;
      }
      subsets[j + 1] = s; sums[j + 1] = v;
    }
    for (int i = 0; i < ns; ++i) {
      int nc = 0;
      for (int q = 0; q < nb; ++q) {
        int ok = 1;
        for (int r = 0; r < R && ok; ++r)
          if ((subsets[i] >> r & 1) && !cut_crosses(g, w, w->bal_t[q], r)) ok = 0;
        if (ok) cand[nc++] = q;
      }
      if (nc) return max_weight(w, cand, nc);
    }
  }
  for (int q = 0; q < nb; ++q) cand[q] = q;
  return max_weight(w, cand, nb);
}

/* Mark the subtree hanging below `child` (tree.py:325-348 _part_nodes). */
static void mark_subtree(RcoWork* w, int child) {
  memset(w->side, 0, (size_t)w->n);
  int top = 0;
  w->bfs[top++] = child; /* bfs[] is free once the subtree sums exist */
  while (top) {
    int u = w->bfs[--top];
    w->side[u] = 1;
    for (int s = w->toff[u]; s < w->toff[u + 1]; ++s) {
      int v = w->tadj[s];
      if (w->parent[v] == u && w->parent_t[v] == w->tadj_t[s]) w->bfs[top++] = v;
    }
  }
}

/* tree.py:670-718 bipartition_tree.  On success w->side holds the split. */
static int rco_bipartition(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t chain,
                           uint32_t prop_no, const RcReplay* rp, const RcReplayStep* rs,
                           int64_t* counters, uint32_t* flags, int64_t* child_pop,
                           int32_t* info) {
  int64_t max_attempts = c->max_attempts > 0 ? c->max_attempts : INT64_MAX;
  int64_t r = c->node_repeats;
  int n_trees_rec = rs ? rs->n_trees : 0;
  int warn_on = !c->allow_pair_reselection && c->warn_attempts > 0;
  /* tree.py:679-700: tree t is first tested after t*node_repeats failed attempts (the
     balanced-cut set does not depend on the root, so repeats of a failed tree fail too).
     node_repeats == 0: the tree drawn before the loop is replaced at once (restarts == 0),
     and that second tree is then retried until max_attempts. */
  uint32_t t0 = (rs && r == 0 && n_trees_rec > 0) ? (uint32_t)(n_trees_rec - 1) : 0;
  for (uint32_t t = t0;; ++t) {
    int64_t attempts = r > 0 ? (int64_t)t * r : (t == t0 ? 0 : max_attempts);
    if (attempts >= max_attempts) {
      counters[RC_CTR_ATTEMPTS] += max_attempts;
      if (warn_on && max_attempts >= c->warn_attempts) *flags |= RC_FLAG_WARNED;
      return c->allow_pair_reselection ? -1 /* ReselectException */ : RC_ERR_MAX_ATTEMPTS;
    }
    if (rs && (int)t >= n_trees_rec) return RC_ERR_REPLAY;
    int rc;
    int64_t tree_idx = rs ? rs->first_tree + t : 0;
    if (c->tree_kind == RC_TREE_MST) {
      if (rp) {
        const double* rw = rp->weights + tree_idx * (int64_t)g->n_edges;
        for (int j = 0; j < w->m; ++j) w->w[j] = rw[w->eid[j]];
      } else {
        rco_random_weights(g, c, w, chain, t, prop_no);
      }
      rc = rco_kruskal(w);
    } else {
      rc = rco_wilson(g, c, w, chain, t, prop_no, rp, tree_idx, &counters[RC_CTR_WALK]);
    }
    if (rc) return rc;
    counters[RC_CTR_TREES]++;
    build_tree_adj(w);
    /* tree.py:377 root = choice([x for x in h if h.degree(x) > 1]) */
    int root = -1;
    if (rs && (int)t == n_trees_rec - 1) {
      root = w->g2l[rs->root];
      if (root < 0 || w->toff[root + 1] - w->toff[root] < 2) return RC_ERR_REPLAY;
    } else {
      int n_int = 0;
      for (int i = 0; i < w->n; ++i) n_int += (w->toff[i + 1] - w->toff[i]) > 1;
      if (!n_int) return RC_ERR_NO_ROOT;
      RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_ROOT, 0, t, prop_no, 0);
      int kth = (int)rc_below(p.v[0], p.v[1], (uint32_t)n_int);
      for (int i = 0; i < w->n; ++i)
        if ((w->toff[i + 1] - w->toff[i]) > 1 && kth-- == 0) { root = i; break; }
    }
    rco_bfs(w, root);
    rco_subtree_pops(g, w);
    int nb = rco_balanced_cuts(w, c->pop_target, c->epsilon, root);
    if (info) { info[0] = (int)t + 1; info[1] = nb; }
    if (nb) {
      if (rs && (int)t != n_trees_rec - 1) return RC_ERR_REPLAY;
      int pick = -1;
      if (rs && rs->cut_edge >= 0) {
        for (int q = 0; q < nb; ++q) if (w->tw_eid[w->bal_t[q]] == rs->cut_edge) pick = q;
        if (pick < 0) return RC_ERR_REPLAY;
      } else {
        pick = rco_choose_cut(g, c, w, nb, chain, t, prop_no);
      }
      int child = w->bal_child[pick];
      *child_pop = w->subpop[child];
      counters[RC_CTR_ATTEMPTS] += attempts + 1;
      if (warn_on && attempts >= c->warn_attempts) *flags |= RC_FLAG_WARNED;
      mark_subtree(w, child);
      return RC_OK;
    }
  }
}

/* One call of recom (proposals/tree_proposals.py:36-146) followed by the chain's validity
 * check; returns RC_OK with *accepted = 1 when the proposal was committed. */
static int rco_propose(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t chain,
                       rc_label_t* assign, int64_t* tally, uint32_t* mask, int32_t* cut_count,
                       uint32_t prop_no, int64_t* counters, uint32_t* flags,
                       const RcReplay* rp, const RcReplayStep* rs, int32_t* info,
                       int* accepted, int32_t* last_pair) {
  int k = c->n_districts, words = (g->n_edges + 31) / 32;
  *accepted = 0;
  if (*cut_count <= 0) return RC_ERR_NO_CUT_EDGES;
  memset(w->bad_pairs, 0, 1024 * 8);
  int n_bad = 0, tot_pairs = k * (k - 1) / 2;
  int a, b, rc;
  int64_t child_pop = 0;
  for (uint32_t retry = 0;; ++retry) {
    if (n_bad >= tot_pairs) return RC_ERR_ALL_PAIRS_BAD;
    int32_t e;
    if (rs) {
      e = rs->pair_edge;
      if (e < 0 || e >= g->n_edges || !(mask[e >> 5] >> (e & 31) & 1)) return RC_ERR_REPLAY;
    } else {
      RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_PAIR, retry, 0, prop_no, 0);
      e = kth_cut_edge(mask, words, rc_below(p.v[0], p.v[1], (uint32_t)*cut_count));
    }
    a = assign[g->edge_uv[2 * e]]; b = assign[g->edge_uv[2 * e + 1]];
    if (a > b) { int t = a; a = b; b = t; }               /* tree_proposals.py:113 */
    /* (256 x 256 bits: the engine accepts allow_pair_reselection up to 256 districts) */
    if (c->allow_pair_reselection && (w->bad_pairs[(a * 256 + b) >> 6] >> ((a * 256 + b) & 63) & 1)) continue;
    rco_gather(g, w, assign, a, b);
    counters[RC_CTR_NODES] += w->n;
    for (int i = 0; i < w->n; ++i)
      counters[RC_CTR_SLOTS] += g->row_ptr[w->nodes[i] + 1] - g->row_ptr[w->nodes[i]];
    rc = rco_bipartition(g, c, w, chain, prop_no, rp, rs, counters, flags, &child_pop, info);
    if (rc == -1) {                                       /* tree_proposals.py:133-136 */
      w->bad_pairs[(a * 256 + b) >> 6] |= 1ull << ((a * 256 + b) & 63);
      ++n_bad; counters[RC_CTR_RESELECT]++;
      rco_ungather(w);
      continue;
    }
    if (rc) { rco_ungather(w); return rc; }
    break;
  }
  /* tree.py:947-964: the returned subset (root side) takes parts[-2] = a, the rest b. */
  int64_t pop_b = child_pop, pop_a = w->totpop - child_pop;
  int valid = 1;
  if (c->pop_lower <= c->pop_upper) {                     /* constraints/bounds.py:25-28 */
    double lo = (double)(pop_a < pop_b ? pop_a : pop_b), hi = (double)(pop_a < pop_b ? pop_b : pop_a);
    valid = (c->pop_lower <= lo) && (hi <= c->pop_upper);
  }
  if (valid && c->accept_rule == RC_ACCEPT_CUT_EDGE) {      /* accept.py:19-35 */
    int delta = 0;
    for (int j = 0; j < w->m; ++j) {
      int e = w->eid[j];
      int was = mask[e >> 5] >> (e & 31) & 1;
      int now = w->side[w->eu[j]] != w->side[w->ev[j]];
      delta += now - was;
    }
    double ratio = (double)*cut_count / (double)(*cut_count + delta);
    RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_ACCEPT, 0, 0, prop_no, 0);
    if (!(rc_unit_double(p.v[0], p.v[1]) < (ratio < 1.0 ? ratio : 1.0))) {
      *accepted = 1;                                         /* chain.py:191-195: counts, state kept */
      rco_ungather(w);
      return RC_OK;
    }
  }
  if (valid) {
    for (int i = 0; i < w->n; ++i) assign[w->nodes[i]] = (rc_label_t)(w->side[i] ? b : a);
    tally[a] = pop_a; tally[b] = pop_b;                   /* updaters/tally.py:162-165 */
    int delta = 0;
    for (int j = 0; j < w->m; ++j) {                      /* updaters/cut_edges.py:118-120 */
      int e = w->eid[j];
      int was = mask[e >> 5] >> (e & 31) & 1;
      int now = w->side[w->eu[j]] != w->side[w->ev[j]];
      if (was != now) { mask[e >> 5] ^= 1u << (e & 31); delta += now - was; }
    }
    *cut_count += delta;
    *accepted = 1;
    if (last_pair) { last_pair[0] = a; last_pair[1] = b; }
  } else {
    counters[RC_CTR_INVALID]++;
  }
  rco_ungather(w);
  return RC_OK;
}

/* proposals/tree_proposals.py:149-267 reversible_recom; *accepted = 1 also for self-loops
 * (the chain counts every call as a step), 0 only when the proposal fails the Bounds. */
/* rv / rw (TEST HOOK, optional): the recorded randomness of one reference call - the ordered pair,
 * the edge between the two districts, the walk stacks of the tree (rw, tree index rv->tree_idx),
 * the BFS root, the chosen cut and the acceptance draw - replayed by identity like rco_propose's
 * RcReplayStep; info (optional): {trees, balanced cuts, -, outcome RCO_REV_*}. */
static int rco_propose_reversible(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t chain,
                                  rc_label_t* assign, int64_t* tally, uint32_t* mask, int32_t* cut_count,
                                  uint32_t prop_no, int64_t* counters, int* accepted, int32_t* last_pair,
                                  const RcoRevStep* rv, const RcReplay* rw, int32_t* info) {
  int k = c->n_districts, M = c->rev_M > 0 ? c->rev_M : 1;
  *accepted = 1;
  if (info) info[3] = RCO_REV_NO_PAIR;
  RcPhilox rp = rc_draw(c->seed, chain, RC_DRAW_PAIR, 0, 0, prop_no, 0);
  int pr = (int)rc_below(rp.v[0], rp.v[1], (uint32_t)(k * k));
  int d_out = pr / k, d_in = pr % k;
  if (rv) {
    d_out = rv->d_out; d_in = rv->d_in;
    if (d_out < 0 || d_out >= k || d_in < 0 || d_in >= k) return RC_ERR_REPLAY;
  }
  if (d_out == d_in) return RC_OK;
  int n_pair = 0;
  for (int e = 0; e < g->n_edges; ++e) {
    int la = assign[g->edge_uv[2 * e]], lb = assign[g->edge_uv[2 * e + 1]];
    if ((la == d_out && lb == d_in) || (la == d_in && lb == d_out)) ++n_pair;
  }
  if (!n_pair) return (rv && rv->pair_edge >= 0) ? RC_ERR_REPLAY : RC_OK;
  int kth = (int)rc_below(rp.v[2], rp.v[3], (uint32_t)n_pair), e_star = -1;
  for (int e = 0; e < g->n_edges; ++e) {
    int la = assign[g->edge_uv[2 * e]], lb = assign[g->edge_uv[2 * e + 1]];
    if (((la == d_out && lb == d_in) || (la == d_in && lb == d_out)) && kth-- == 0) { e_star = e; break; }
  }
  if (rv) {  /* the recorded edge must join the recorded pair */
    e_star = rv->pair_edge;
    if (e_star < 0 || e_star >= g->n_edges) return RC_ERR_REPLAY;
    int la = assign[g->edge_uv[2 * e_star]], lb = assign[g->edge_uv[2 * e_star + 1]];
    if (!((la == d_out && lb == d_in) || (la == d_in && lb == d_out))) return RC_ERR_REPLAY;
  }
  int lab_root = assign[g->edge_uv[2 * e_star]], lab_other = assign[g->edge_uv[2 * e_star + 1]];
  int a = lab_root < lab_other ? lab_root : lab_other, b = lab_root < lab_other ? lab_other : lab_root;
  rco_gather(g, w, assign, a, b);
  counters[RC_CTR_NODES] += w->n;
  for (int i = 0; i < w->n; ++i) counters[RC_CTR_SLOTS] += g->row_ptr[w->nodes[i] + 1] - g->row_ptr[w->nodes[i]];
  if (w->n < 3) { rco_ungather(w); return RC_ERR_NO_ROOT; }
  int rc = rco_wilson(g, c, w, chain, 0, prop_no, rv ? rw : NULL, rv ? rv->tree_idx : 0, &counters[RC_CTR_WALK]);
  counters[RC_CTR_TREES]++; counters[RC_CTR_ATTEMPTS]++;
  if (info) { info[0] = 1; info[3] = RCO_REV_NO_CUT; }
  if (rc) { rco_ungather(w); return rc; }
  build_tree_adj(w);
  int n_int = 0, root = -1;
  for (int i = 0; i < w->n; ++i) n_int += (w->toff[i + 1] - w->toff[i]) > 1;
  if (!n_int) { rco_ungather(w); return RC_ERR_NO_ROOT; }
  {
    RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_ROOT, 0, 0, prop_no, 0);
    int q = (int)rc_below(p.v[0], p.v[1], (uint32_t)n_int);
    for (int i = 0; i < w->n; ++i)
      if ((w->toff[i + 1] - w->toff[i]) > 1 && q-- == 0) { root = i; break; }
  }
  if (rv) {  /* tree.py:377: the recorded root must be a node of tree degree > 1 */
    root = (rv->root >= 0 && rv->root < g->n_nodes) ? w->g2l[rv->root] : -1;
    if (root < 0 || (w->toff[root + 1] - w->toff[root]) <= 1) { rco_ungather(w); return RC_ERR_REPLAY; }
  }
  rco_bfs(w, root);
  rco_subtree_pops(g, w);
  int nb = rco_balanced_cuts(w, c->pop_target, c->epsilon, root);
  if (info) info[1] = nb;
  if (nb > M) { rco_ungather(w); return RC_ERR_REVERSIBILITY; }
  if (!nb) { rco_ungather(w); return (rv && rv->cut_edge >= 0) ? RC_ERR_REPLAY : RC_OK; }
  RcPhilox pc = rc_draw(c->seed, chain, RC_DRAW_CUT, 0, 0, prop_no, 0);
  int pick = (int)rc_below(pc.v[0], pc.v[1], (uint32_t)nb);
  if (rv) {  /* the recorded Cut.edge must be one of the balanced cuts found here */
    pick = -1;
    for (int q = 0; q < nb; ++q) {
      int ch = w->bal_child[q], par = w->parent[ch], eid = -1;
      for (int t = 0; t < w->nt; ++t)
        if ((w->tu[t] == ch && w->tv[t] == par) || (w->tu[t] == par && w->tv[t] == ch)) eid = w->tw_eid[t];
      if (eid == rv->cut_edge) pick = q;
    }
    if (pick < 0) { rco_ungather(w); return RC_ERR_REPLAY; }
  }
  int child = w->bal_child[pick];
  int64_t pop_other = w->subpop[child], pop_root = w->totpop - pop_other;
  mark_subtree(w, child);  /* side = 1: cut away from the root -> lab_other */
  int seam = 0, delta = 0;
  for (int j = 0; j < w->m; ++j) {
    int e = w->eid[j];
    int was = mask[e >> 5] >> (e & 31) & 1;
    int now = w->side[w->eu[j]] != w->side[w->ev[j]];
    seam += now; delta += now - was;
  }
  double prob = (double)nb / (double)((int64_t)M * seam);
  if (prob > 1.0) { rco_ungather(w); return RC_ERR_REVERSIBILITY; }
  RcPhilox pa = rc_draw(c->seed, chain, RC_DRAW_ACCEPT, 0, 0, prop_no, 0);
  double u_acc = rv ? rv->accept_u : rc_unit_double(pa.v[0], pa.v[1]);
  if (info) { info[2] = seam; info[3] = RCO_REV_REJECTED; }
  if (!(u_acc < prob)) { rco_ungather(w); return RC_OK; }
  if (info) info[3] = RCO_REV_MOVED;
  if (c->pop_lower <= c->pop_upper) {
    double lo = (double)(pop_root < pop_other ? pop_root : pop_other), hi = (double)(pop_root < pop_other ? pop_other : pop_root);
    if (!(c->pop_lower <= lo && hi <= c->pop_upper)) { counters[RC_CTR_INVALID]++; *accepted = 0; rco_ungather(w); return RC_OK; }
  }
  for (int i = 0; i < w->n; ++i) assign[w->nodes[i]] = (rc_label_t)(w->side[i] ? lab_other : lab_root);
  tally[lab_root] = pop_root; tally[lab_other] = pop_other;
  for (int j = 0; j < w->m; ++j) {
    int e = w->eid[j];
    int was = mask[e >> 5] >> (e & 31) & 1;
    int now = w->side[w->eu[j]] != w->side[w->ev[j]];
    if (was != now) mask[e >> 5] ^= 1u << (e & 31);
  }
  *cut_count += delta;
  if (last_pair) { last_pair[0] = a; last_pair[1] = b; }
  rco_ungather(w);
  return RC_OK;
}

/* chain.py:185-195: propose until a valid state, accept (always), count the step. */
int rco_step(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t global_chain,
             rc_label_t* assign, int64_t* tally, uint32_t* mask, int32_t* cut_count,
             uint32_t* proposal_no, int64_t* counters, uint32_t* flags, int32_t* last_pair) {
  int streak_cap = c->max_invalid_streak > 0 ? c->max_invalid_streak : 1000;
  for (int streak = 0; streak < streak_cap; ++streak) {
    int accepted = 0;
    uint32_t p = (*proposal_no)++;
    int rc = c->proposal_kind == RC_PROPOSAL_REVERSIBLE
                 ? rco_propose_reversible(g, c, w, global_chain, assign, tally, mask, cut_count, p, counters,
                                          &accepted, last_pair, NULL, NULL, NULL)
                 : rco_propose(g, c, w, global_chain, assign, tally, mask, cut_count, p, counters,
                               flags, NULL, NULL, NULL, &accepted, last_pair);
    if (rc) return rc;
    if (accepted) return RC_OK;
  }
  return RC_ERR_INVALID_STREAK;
}

int rco_replay(const RcGraph* g, const RcConfig* c, RcoWork* w, const RcReplay* rp,
               rc_label_t* assign, int64_t* tally, uint32_t* mask, int32_t* cut_count,
               int64_t* counters, uint32_t* flags) {
  for (int s = 0; s < rp->n_steps; ++s) {
    int accepted = 0;
    int32_t info[4] = {0, 0, 0, 0};
    int rc = rco_propose(g, c, w, 0, assign, tally, mask, cut_count, (uint32_t)s, counters,
                         flags, rp, &rp->steps[s], info, &accepted, NULL);
    info[2] = *cut_count; info[3] = rc ? rc : (accepted ? RC_OK : RC_ERR_INVALID_STREAK);
    if (rp->info_trace) memcpy(rp->info_trace + 4 * s, info, sizeof info);
    if (rp->assign_trace) memcpy(rp->assign_trace + (size_t)s * g->n_nodes, assign, (size_t)g->n_nodes);
    if (rc) return rc;
  }
  return RC_OK;
}

/* TEST HOOK: recorded reference calls of reversible_recom (oracle/record_reversible.py) applied to one
 * plan.  rw carries the Wilson arrays of the recorded trees (wilson_root, stack_ptr, stack_val);
 * assign_trace [n_steps][N] and info_trace [n_steps][4] are optional outputs. */
int rco_replay_reversible(const RcGraph* g, const RcConfig* c, RcoWork* w, int32_t n_steps,
                          const RcoRevStep* steps, const RcReplay* rw, rc_label_t* assign, int64_t* tally,
                          uint32_t* mask, int32_t* cut_count, int64_t* counters,
                          rc_label_t* assign_trace, int32_t* info_trace) {
  for (int s = 0; s < n_steps; ++s) {
    int accepted = 0;
    int32_t info[4] = {0, 0, 0, 0};
    int rc = rco_propose_reversible(g, c, w, 0, assign, tally, mask, cut_count, (uint32_t)s, counters,
                                    &accepted, NULL, &steps[s], rw, info);
    if (info_trace) memcpy(info_trace + 4 * s, info, sizeof info);
    if (assign_trace) memcpy(assign_trace + (size_t)s * g->n_nodes, assign, (size_t)g->n_nodes * sizeof(rc_label_t));
    if (rc) return rc;
    if (!accepted) return RC_ERR_INVALID_STREAK;  /* the recorded chains run without Bounds */
  }
  return RC_OK;
}

/* ------------------------------------------------------------------------------------ */
/* Seed plans: tree.py:971-1085 recursive_tree_part with method = bipartition_tree
 * (random_spanning_tree, node_repeats = 1): k - 2 one-sided splits of the shrinking
 * remainder with a debt-tightened target (tree.py:1021-1050), then one two-sided split
 * (tree.py:1054-1061).  Nodes not yet assigned carry RC_SEED_REMAINING.  One-sided cuts
 * (tree.py:386-409): the subtree below a node if ITS population fits, else everything but
 * the subtree if THAT fits.  rp (optional): recorded reference randomness, one
 * RcReplayStep per split (pair_edge unused). */
int rco_seed_plan(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t chain,
                  int32_t seed_max_attempts, rc_label_t* assign, int64_t* tally,
                  int64_t* counters, const RcReplay* rp) {
  const int k = c->n_districts;
  const double pt = c->pop_target, eps = c->epsilon;
  const double lb_pop = pt * (1 - eps), ub_pop = pt * (1 + eps);
  const int64_t max_attempts = seed_max_attempts > 0 ? seed_max_attempts : INT64_MAX;
  for (int v = 0; v < g->n_nodes; ++v) assign[v] = RC_SEED_REMAINING;
  for (int d = 0; d < k; ++d) tally[d] = 0;
  double debt = 0;
  for (int part = 0; part < k - 2; ++part) {
    double min_pop = fmax(pt * (1 - eps), pt * (1 - eps) - debt);
    double max_pop = fmin(pt * (1 + eps), pt * (1 + eps) - debt);
    double ideal = (min_pop + max_pop) / 2;
    double e1 = (max_pop - min_pop) / (2 * ideal);
    const uint32_t prop_no = RC_SEED_PROPOSAL | (uint32_t)part;
    const RcReplayStep* rs = rp ? &rp->steps[part] : NULL;
    rco_gather(g, w, assign, RC_SEED_REMAINING, RC_SEED_REMAINING);
    int done = 0;
    for (uint32_t t = 0; !done; ++t) {
      if ((int64_t)t >= max_attempts) { rco_ungather(w); return RC_ERR_MAX_ATTEMPTS; }
      if (rs && (int)t >= rs->n_trees) { rco_ungather(w); return RC_ERR_REPLAY; }
      if (rp) {
        const double* rw = rp->weights + (rs->first_tree + t) * (int64_t)g->n_edges;
        for (int j = 0; j < w->m; ++j) w->w[j] = rw[w->eid[j]];
      } else {
        rco_random_weights(g, c, w, chain, t, prop_no);
      }
      int rc = rco_kruskal(w);
      if (rc) { rco_ungather(w); return rc; }
      counters[RC_CTR_TREES]++;
      build_tree_adj(w);
      int root = -1;
      if (rs && (int)t == rs->n_trees - 1) {
        root = w->g2l[rs->root];
        if (root < 0 || w->toff[root + 1] - w->toff[root] < 2) { rco_ungather(w); return RC_ERR_REPLAY; }
      } else {
        int n_int = 0;
        for (int i = 0; i < w->n; ++i) n_int += (w->toff[i + 1] - w->toff[i]) > 1;
        if (!n_int) { rco_ungather(w); return RC_ERR_NO_ROOT; }
        RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_ROOT, 0, t, prop_no, 0);
        int kth = (int)rc_below(p.v[0], p.v[1], (uint32_t)n_int);
        for (int i = 0; i < w->n; ++i)
          if ((w->toff[i + 1] - w->toff[i]) > 1 && kth-- == 0) { root = i; break; }
      }
      rco_bfs(w, root);
      rco_subtree_pops(g, w);
      /* one-sided cuts, canonical order: ascending edge id; bal_child < 0 encodes "ABOVE" */
      int nb = 0;
      for (int v = 0; v < w->n; ++v) {
        if (v == root) continue;
        int64_t p = w->subpop[v];
        if (fabs((double)p - ideal) <= ideal * e1) w->bal_t[nb++] = w->parent_t[v];
        else if (fabs((double)(w->totpop - p) - ideal) <= ideal * e1) w->bal_t[nb++] = w->parent_t[v];
      }
      if (!nb) continue;
      if (rs && (int)t != rs->n_trees - 1) { rco_ungather(w); return RC_ERR_REPLAY; }
      t_sort_ctx = w;
      qsort(w->bal_t, (size_t)nb, sizeof(int32_t), cmp_bal);
      int pick = -1;
      if (rs && rs->cut_edge >= 0) {
        for (int q = 0; q < nb; ++q) if (w->tw_eid[w->bal_t[q]] == rs->cut_edge) pick = q;
        if (pick < 0) { rco_ungather(w); return RC_ERR_REPLAY; }
      } else {
        RcPhilox p = rc_draw(c->seed, chain, RC_DRAW_CUT, 0, t, prop_no, 0);
        pick = (int)rc_below(p.v[0], p.v[1], (uint32_t)nb);
      }
      int te = w->bal_t[pick];
      int child = (w->parent[w->tu[te]] == w->tv[te] && w->parent_t[w->tu[te]] == te) ? w->tu[te] : w->tv[te];
      int64_t below = w->subpop[child];
      int take_below = fabs((double)below - ideal) <= ideal * e1;
      mark_subtree(w, child);
      int64_t part_pop = take_below ? below : w->totpop - below;
      for (int i = 0; i < w->n; ++i)
        if ((w->side[i] != 0) == (take_below != 0)) assign[w->nodes[i]] = (rc_label_t)part;
      tally[part] = part_pop;
      counters[RC_CTR_ATTEMPTS] += (int64_t)t + 1;
      if (!(lb_pop <= (double)part_pop && (double)part_pop <= ub_pop)) { rco_ungather(w); return RC_ERR_BALANCE; }
      debt += (double)part_pop - pt;
      done = 1;
    }
    rco_ungather(w);
  }
  /* the last two districts: a two-sided split of what remains (tree.py:1054-1061) */
  {
    RcConfig c2 = *c;
    c2.max_attempts = seed_max_attempts;
    c2.allow_pair_reselection = 0;
    c2.node_repeats = 1;
    c2.tree_kind = RC_TREE_MST;
    c2.cut_policy = RC_CUT_UNIFORM;
    const uint32_t prop_no = RC_SEED_PROPOSAL | (uint32_t)(k - 2);
    const RcReplayStep* rs = rp ? &rp->steps[k - 2] : NULL;
    rco_gather(g, w, assign, RC_SEED_REMAINING, RC_SEED_REMAINING);
    int64_t child_pop = 0;
    uint32_t flags = 0;
    int rc = rco_bipartition(g, &c2, w, chain, prop_no, rp, rs, counters, &flags, &child_pop, NULL);
    if (rc) { rco_ungather(w); return rc == -1 ? RC_ERR_MAX_ATTEMPTS : rc; }
    for (int i = 0; i < w->n; ++i) assign[w->nodes[i]] = (rc_label_t)(w->side[i] ? k - 1 : k - 2);
    tally[k - 1] = child_pop;
    tally[k - 2] = w->totpop - child_pop;
    rco_ungather(w);
    for (int d = k - 2; d < k; ++d)
      if (!(lb_pop <= (double)tally[d] && (double)tally[d] <= ub_pop)) return RC_ERR_BALANCE;
  }
  return RC_OK;
}

/* Known-answer entry (reference tests/test_tree.py:354-385): balanced cut edges of a GIVEN
 * tree.  tree_uv [n-1][2] node indices, pops [n]; returns count, fills out_uv [..][2]
 * with (child, parent) pairs in ascending (min,max) order. */
int rco_balanced_cuts_of_tree(int32_t n, const int32_t* tree_uv, const int32_t* pops,
                              double ideal, double eps, int32_t root, int32_t* out_uv) {
  RcGraph g; memset(&g, 0, sizeof g);
  g.n_nodes = n; g.n_edges = n - 1; g.pop = pops;
  RcoWork* w = rco_work_create(&g);
  w->n = n; w->nt = n - 1; w->totpop = 0;
  for (int i = 0; i < n; ++i) { w->nodes[i] = i; w->totpop += pops[i]; }
  for (int t = 0; t < n - 1; ++t) {
    int u = tree_uv[2 * t], v = tree_uv[2 * t + 1];
    w->tu[t] = u < v ? u : v; w->tv[t] = u < v ? v : u; w->tw[t] = 0; w->tw_eid[t] = 0;
  }
  /* canonical edge ids: rank of (min,max) */
  for (int t = 0; t < n - 1; ++t) {
    int r = 0;
    for (int s = 0; s < n - 1; ++s)
      if (w->tu[s] < w->tu[t] || (w->tu[s] == w->tu[t] && w->tv[s] < w->tv[t])) ++r;
    w->tw_eid[t] = r;
  }
  build_tree_adj(w);
  rco_bfs(w, root);
  rco_subtree_pops(&g, w);
  int nb = rco_balanced_cuts(w, ideal, eps, root);
  for (int i = 0; i < nb; ++i) {
    int child = w->bal_child[i];
    out_uv[2 * i] = child; out_uv[2 * i + 1] = w->parent[child];
  }
  rco_work_destroy(w);
  return nb;
}

/* Raw MST entry for differential tests against scipy/networkx: weights per edge of an
 * arbitrary connected graph -> tree edge ids (ascending). */
int rco_mst_of_graph(const RcGraph* g, const double* weights, int32_t* out_eids) {
  RcoWork* w = rco_work_create(g);
  rc_label_t* assign = (rc_label_t*)xcalloc((size_t)g->n_nodes, sizeof(rc_label_t));
  rco_gather(g, w, assign, 0, 0);
  for (int j = 0; j < w->m; ++j) w->w[j] = weights[w->eid[j]];
  int rc = rco_kruskal(w);
  for (int t = 0; t < w->nt; ++t) out_eids[t] = w->tw_eid[t];
  int nt = w->nt;
  rco_ungather(w);
  free(assign);
  rco_work_destroy(w);
  return rc ? -rc : nt;
}

/* ---- multi-chain driver for the CPU baseline: one chain per task, T host threads ---- */
typedef struct {
  const RcGraph* g; const RcConfig* c; int64_t n_steps;
  rc_label_t* assign; int64_t assign_stride; int64_t* tally; uint32_t* mask; int64_t mask_stride;
  int32_t* cut_count; uint32_t* proposal_no; int32_t* status; uint32_t* flags; int64_t* counters;
  int64_t* step_no; int32_t* last_pair;
  int n_chains; volatile int next; pthread_mutex_t mu;
} RcoJob;

static void* rco_worker(void* arg) {
  RcoJob* j = (RcoJob*)arg;
  RcoWork* w = rco_work_create(j->g);
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int ch = j->next < j->n_chains ? j->next++ : -1;
    pthread_mutex_unlock(&j->mu);
    if (ch < 0) break;
    if (j->status[ch] != RC_OK) continue;
    for (int64_t s = 0; s < j->n_steps; ++s) {
      int rc = rco_step(j->g, j->c, w, (uint32_t)(ch + j->c->chain_offset),
                        j->assign + ch * j->assign_stride, j->tally + (int64_t)ch * j->c->n_districts,
                        j->mask + ch * j->mask_stride, &j->cut_count[ch], &j->proposal_no[ch],
                        j->counters + (int64_t)ch * RC_N_COUNTERS, &j->flags[ch],
                        j->last_pair ? j->last_pair + 2 * ch : NULL);
      if (rc) { j->status[ch] = rc; break; }
      j->step_no[ch]++;
    }
  }
  rco_work_destroy(w);
  return NULL;
}

int rco_run_chains(const RcGraph* g, const RcConfig* c, int64_t n_steps, const RcState* st,
                   int n_threads) {
  RcoJob j;
  memset(&j, 0, sizeof j);
  j.g = g; j.c = c; j.n_steps = n_steps;
  j.assign = st->d_assignment; j.assign_stride = st->assign_stride;
  j.tally = st->d_tally; j.mask = st->d_cut_mask; j.mask_stride = st->mask_stride;
  j.cut_count = st->d_cut_count; j.proposal_no = st->d_proposal_no; j.status = st->d_status;
  j.flags = st->d_flags; j.counters = st->d_counters; j.step_no = st->d_step_no;
  j.last_pair = st->d_last_pair;
  j.n_chains = c->n_chains; j.next = 0;
  pthread_mutex_init(&j.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)xcalloc((size_t)n_threads, sizeof(pthread_t));
  for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, rco_worker, &j);
  rco_worker(&j);
  for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&j.mu);
  return RC_OK;
}

int rco_init_chains(const RcGraph* g, const RcConfig* c, const RcState* st) {
  for (int ch = 0; ch < c->n_chains; ++ch) {
    int rc = rco_init_state(g, c, st->d_assignment + ch * st->assign_stride,
                            st->d_tally + (int64_t)ch * c->n_districts,
                            st->d_cut_mask + ch * st->mask_stride, &st->d_cut_count[ch]);
    if (rc) return rc;
    st->d_step_no[ch] = 0; st->d_proposal_no[ch] = 0; st->d_status[ch] = RC_OK;
    st->d_flags[ch] = 0;
    if (st->d_last_pair) { st->d_last_pair[2 * ch] = -1; st->d_last_pair[2 * ch + 1] = -1; }
    memset(st->d_counters + (int64_t)ch * RC_N_COUNTERS, 0, RC_N_COUNTERS * 8);
  }
  return RC_OK;
}

/* Philox known-answer access for tests. */
void rco_philox(uint32_t k0, uint32_t k1, const uint32_t* ctr, uint32_t* out) {
  RcPhilox p = rc_philox4x32_10(k0, k1, ctr[0], ctr[1], ctr[2], ctr[3]);
  memcpy(out, p.v, 16);
}

/* ---- per-plan statistics --------------------------------- */

/* updaters/county_splits.py:145-161 total_reg_splits: walk the cut edges, mark a region when
 * both ends of a cut edge lie in it, count the marked regions -> out[0];
 * updaters/county_splits.py:82-99 compute_county_splits (no parent): the set of districts
 * seen in every county; counties with more than one -> out[1]; total set sizes -> out[2]. */
int rco_region_splits(const RcGraph* g, int32_t k, const rc_label_t* assign, const int32_t* region,
                      int32_t n_regions, int32_t* out) {
  uint8_t* marked = (uint8_t*)xcalloc((size_t)n_regions, 1);
  uint8_t* seen = (uint8_t*)xcalloc((size_t)n_regions * (size_t)k, 1);
  for (int e = 0; e < g->n_edges; ++e) {
    const int u = g->edge_uv[2 * e], v = g->edge_uv[2 * e + 1];
    if (assign[u] == assign[v]) continue; /* not a cut edge */
    if (region[u] == region[v] && region[u] >= 0 && region[u] < n_regions) marked[region[u]] = 1;
  }
  for (int v = 0; v < g->n_nodes; ++v)
    if (region[v] >= 0 && region[v] < n_regions) seen[(size_t)region[v] * k + assign[v]] = 1;
  out[0] = out[1] = out[2] = 0;
  for (int r = 0; r < n_regions; ++r) {
    int nd = 0;
    for (int d = 0; d < k; ++d) nd += seen[(size_t)r * k + d];
    out[0] += marked[r];
    out[1] += nd > 1;
    out[2] += nd;
  }
  free(marked);
  free(seen);
  return RC_OK;
}

/* constraints/contiguity.py:169-181 contiguous -> is_connected_bfs per district
 * (contiguity.py:184-204, :244-270) and contiguous_components (:226-241): here one breadth-first sweep
 * that never crosses a cut edge counts the pieces of every district.  out[d] = pieces of d. */
int rco_contiguity(const RcGraph* g, int32_t k, const rc_label_t* assign, int32_t* out) {
  const int n = g->n_nodes;
  uint8_t* visited = (uint8_t*)xcalloc((size_t)n, 1);
  int32_t* queue = (int32_t*)xcalloc((size_t)n, sizeof(int32_t));
  for (int d = 0; d < k; ++d) out[d] = 0;
  for (int s = 0; s < n; ++s) {
    if (visited[s]) continue;
    if (assign[s] < k) out[assign[s]]++;
    int head = 0, tail = 0;
    queue[tail++] = s;
    visited[s] = 1;
    while (head < tail) {
      const int u = queue[head++];
      for (int q = g->row_ptr[u]; q < g->row_ptr[u + 1]; ++q) {
        const int v = g->col_idx[q];
        if (!visited[v] && assign[v] == assign[u]) { visited[v] = 1; queue[tail++] = v; }
      }
    }
  }
  free(visited);
  free(queue);
  return RC_OK;
}
