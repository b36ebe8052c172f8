// recom_abi.cu - host side of the C ABI declared in include/recom_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "recom_kernels.cu"
#include "recom_stats.cu"

namespace {
std::mutex g_const_mutex;            // guards g_const_event and the order of writes to rc::g_params
cudaEvent_t g_const_event[64] = {};  // per device: the last launch that reads rc::g_params

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_OK(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(RC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                  __FILE__, __LINE__);                                                  \
  } while (0)

// inside rc_engine_create after the engine exists: free it before reporting the error
#define CUDA_OK_E(call)                                                                 \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      rc_engine_destroy(e);                                                             \
      return fail(RC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                  __FILE__, __LINE__);                                                  \
    }                                                                                   \
  } while (0)

template <typename T>
int upload(const T* h, size_t n, T** d) {
  *d = nullptr;
  if (n == 0) return RC_OK;
  CUDA_OK(cudaMalloc((void**)d, n * sizeof(T)));
  CUDA_OK(cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return RC_OK;
}

}  // namespace

struct RcEngine {
  int device = 0;
  rc::Params p{};
  bool bound = false;
  int grid = 0, capacity = 0, threads = 0, ctas_per_sm = 0, n_sms = 0;
  size_t smem = 0, scratch_bytes = 0, sched_bytes = 0, spill_bytes = 0, cold_bytes = 0;
  bool spill = false;
  const void* kernel = nullptr;
  const void* kernel_share = nullptr;  // the build with shared retries
  int share_mode = 2;                  // RECOM_B200_SHARE: 0 never, 1 always, 2 (default) whichever build measures faster
  double ms_avg[2] = {0, 0};           // launch time of the plain / shared-retry build at stat_steps steps per launch
  int n_obs[2] = {0, 0}, n_tried[2] = {0, 0}, since_explore = 0;
  int64_t stat_steps = -1;
  struct Obs { cudaEvent_t a = nullptr, b = nullptr; bool share = false, pending = false; int64_t steps = 0; } obs[8];
  int64_t launches = 0;
  // owned device memory
  int32_t *d_row_ptr = nullptr, *d_col_idx = nullptr, *d_slot_edge = nullptr, *d_edge_uv = nullptr,
          *d_pop = nullptr, *d_region = nullptr;
  void* d_scratch = nullptr;
  int* d_cc_scratch = nullptr;  // union-find parents of rc_engine_contiguity when N words exceed shared memory
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
};

static int wilson_spill_level(const RcEngine* e) { return e->p.cfg.tree_kind == RC_TREE_WILSON ? e->p.wilson_level : (e->spill ? 2 : 0); }

extern "C" {

const char* rc_last_error(void) { return g_err; }
int rc_abi_version(void) { return RC_ABI_VERSION; }
int rc_label_bits(void) { return RC_LABEL_BITS; }

int rc_state_strides(const RcGraph* g, const RcConfig* c, int64_t* assign_stride, int64_t* mask_stride) {
  (void)c;
  if (!g || !assign_stride || !mask_stride) return fail(RC_ERR_INVALID_ARG, "null argument");
  *assign_stride = ((int64_t)g->n_nodes + 15) / 16 * 16;
  *mask_stride = (((int64_t)g->n_edges + 31) / 32 + 3) / 4 * 4;
  return RC_OK;
}

int rc_engine_destroy(RcEngine* e) {
  if (!e) return RC_OK;
  cudaSetDevice(e->device);
  cudaFree(e->d_row_ptr); cudaFree(e->d_col_idx); cudaFree(e->d_slot_edge); cudaFree(e->d_edge_uv);
  cudaFree(e->d_pop); cudaFree(e->d_region); cudaFree(e->d_scratch); cudaFree(e->d_cc_scratch);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  for (auto& o : e->obs) { if (o.a) cudaEventDestroy(o.a); if (o.b) cudaEventDestroy(o.b); }
  delete e;
  return RC_OK;
}

int rc_engine_create(const RcGraph* g, const RcConfig* cfg, int device, RcEngine** out) {
  if (!g || !cfg || !out) return fail(RC_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (g->n_nodes < 2 || g->n_edges < 1) return fail(RC_ERR_INVALID_ARG, "empty graph");
  if (cfg->n_districts < 2 || cfg->n_districts > RC_MAX_DISTRICTS)
    return fail(RC_ERR_INVALID_ARG, "n_districts must be in [2, %d] (%d-bit labels; librecom_b200_l16.so carries 16-bit labels)",
                RC_MAX_DISTRICTS, RC_LABEL_BITS);
  if (cfg->allow_pair_reselection && cfg->n_districts > 256)
    return fail(RC_ERR_INVALID_ARG, "allow_pair_reselection keeps a 256 x 256 bitset of failed pairs: n_districts <= 256");
  if (cfg->n_chains < 1) return fail(RC_ERR_INVALID_ARG, "n_chains must be >= 1");
  if (g->n_region_cols < 0 || g->n_region_cols > 8)
    return fail(RC_ERR_INVALID_ARG, "at most 8 region columns");
  if (cfg->tree_kind != RC_TREE_MST && cfg->tree_kind != RC_TREE_WILSON)
    return fail(RC_ERR_INVALID_ARG, "tree_kind %d is neither RC_TREE_MST nor RC_TREE_WILSON", cfg->tree_kind);
  if (cfg->proposal_kind == RC_PROPOSAL_REVERSIBLE && cfg->tree_kind != RC_TREE_WILSON)
    return fail(RC_ERR_INVALID_ARG, "reversible_recom draws uniform spanning trees: set tree_kind = RC_TREE_WILSON");
  if (!(cfg->epsilon > 0.0) || !(cfg->pop_target > 0.0))
    return fail(RC_ERR_INVALID_ARG, "epsilon and pop_target must be positive");
  CUDA_OK(cudaSetDevice(device));
  RcEngine* e = new (std::nothrow) RcEngine();
  if (!e) return fail(RC_ERR_INVALID_ARG, "out of host memory");
  e->device = device;
  const int N = g->n_nodes, E = g->n_edges, R = g->n_region_cols;
  int rc;
#define UP(field, src, n) if ((rc = upload(src, (size_t)(n), &e->field)) != RC_OK) { rc_engine_destroy(e); return rc; }
  UP(d_row_ptr, g->row_ptr, N + 1)
  UP(d_col_idx, g->col_idx, 2 * (size_t)E)
  UP(d_slot_edge, g->slot_edge, 2 * (size_t)E)
  UP(d_edge_uv, g->edge_uv, 2 * (size_t)E)
  UP(d_pop, g->pop, N)
  UP(d_region, g->region_ids, (size_t)R * N)
#undef UP
  rc::Params& p = e->p;
  p.g.N = N; p.g.E = E; p.g.R = R;
  p.g.row_ptr = e->d_row_ptr; p.g.col_idx = e->d_col_idx; p.g.slot_edge = e->d_slot_edge;
  p.g.edge_uv = e->d_edge_uv; p.g.pop = e->d_pop; p.g.region_ids = e->d_region;
  uint64_t sfix_total = 0;
  for (int r = 0; r < R; ++r) {
    // Free-running keys are (raw + surcharges) rescaled into 32 bits: their resolution shrinks by
    // 1 + the summed surcharge, and exact ties fall to the lower edge id (p(tie in a tree) ~
    // m^2 (1 + sum) / 2^33).  The reference's tutorials use surcharges <= 1; beyond 256 the
    // ties would no longer be negligible, so such values are refused rather than biased.
    if (!(g->surcharges[r] >= 0.0) || g->surcharges[r] > 256.0) {
      rc_engine_destroy(e);
      return fail(RC_ERR_INVALID_ARG, "region surcharges must be in [0, 256]");
    }
    p.g.surcharge[r] = g->surcharges[r];
    p.g.sfix[r] = rc_surcharge_fix(g->surcharges[r]);
    sfix_total += p.g.sfix[r];
  }
  p.g.smult = rc_key_mult(sfix_total, R);
  p.cfg = *cfg;
  // caps: merged pair of two districts, with head-room (RC_ERR_CAPACITY if exceeded)
  const int k = cfg->n_districts;
  int cap_nodes = cfg->cap_nodes > 0 ? cfg->cap_nodes : (int)(2.4 * N / k) + 32;
  int maxdeg = 0;
  for (int v = 0; v < N; ++v) maxdeg = g->row_ptr[v + 1] - g->row_ptr[v] > maxdeg ? g->row_ptr[v + 1] - g->row_ptr[v] : maxdeg;
  bool tight = false;
  if (cfg->cap_nodes <= 0 && cfg->pop_lower <= cfg->pop_upper) {
    // With the chain's population Bounds in force no district ever holds more nodes than fit
    // under pop_upper when the smallest populations are taken first: a rigorous cap (it is the
    // one that binds on a grid of unit populations: 2 x 1050 nodes at BASELINE config 2, where
    // the 2.4 N / k rule allows 2432), and every KB of shared memory decides how many CTAs fit
    std::vector<int32_t> ps(g->pop, g->pop + N);
    std::sort(ps.begin(), ps.end());
    double acc = 0.0;
    int cnt = 0;
    while (cnt < N && acc + (double)ps[cnt] <= cfg->pop_upper) acc += (double)ps[cnt++];
    if (2 * cnt + 8 < cap_nodes) { cap_nodes = 2 * cnt + 8; tight = true; }
  }
  if (cap_nodes > N) cap_nodes = N;
  if (cap_nodes > rc::kMaxLocal) cap_nodes = rc::kMaxLocal;
  int cap_edges = cfg->cap_edges > 0 ? cfg->cap_edges
                                     : (int)(1.02 * cap_nodes * ((double)E / N)) + 64;
  if (cfg->cap_edges <= 0 && tight) {  // a pair of cap_nodes nodes has at most cap_nodes * maxdeg / 2 edges
    const long long by_deg = (long long)cap_nodes * maxdeg / 2;
    if (by_deg < cap_edges) cap_edges = (int)by_deg;
  }
  if (cap_edges > E) cap_edges = E;
  if (cap_edges > 65535) cap_edges = 65535;
  cap_nodes = (cap_nodes + 7) & ~7;
  cap_edges = (cap_edges + 31) & ~31;
  p.cap_nodes = cap_nodes; p.cap_edges = cap_edges;
  const bool wilson = cfg->tree_kind == RC_TREE_WILSON;  // one kernel instantiation per tree kind
  cudaDeviceProp prop;
  CUDA_OK_E(cudaGetDeviceProperties(&prop, device));
  e->n_sms = prop.multiProcessorCount;
  const size_t smem_max = (size_t)prop.sharedMemPerBlockOptin;
  size_t split = 0, total = 0, cold = 0;
  if (wilson) {
    if (cfg->wilson_walkers < 0 || cfg->wilson_walkers > 8) {
      rc_engine_destroy(e);
      return fail(RC_ERR_INVALID_ARG, "wilson_walkers must be in [0, 8] (0 = automatic; one warp per walker)");
    }
    // Concurrent trees per chain-step.  Every walker is one more independent chain of
    // dependent steps on the SM (the walks are latency bound), but a tree drawn beyond the
    // first success is wasted, and each costs 4 B per node of shared memory: 4 is where
    // trees in flight per SM x P(tree is needed) peaks for the BASELINE configurations
    // (P(a tree has a balanced cut) ~ 0.25-0.5).  reversible_recom draws ONE tree per proposal.
    // adjacency slots per node kept in the hot tier: enough for every node of a grid (4), else 8
    // (a planar triangulation has mean degree < 6; longer rows are read from the cold tier)
    const int sh = maxdeg <= 4 ? 2 : 3;
    p.nbr_shift = sh;
    int W = cfg->wilson_walkers > 0 ? cfg->wilson_walkers : 4;
    if (cfg->proposal_kind == RC_PROPOSAL_REVERSIBLE) W = 1;
    // placement: everything in shared memory if it fits (with fewer walkers if need be); else
    // the walkers' state in shared memory and the adjacency in L2; else both in L2
    size_t glob = 0;
    int level = 0;
    if (const char* v = getenv("RECOM_B200_WILSON_LEVEL")) level = std::min(2, std::max(0, atoi(v)));  // dev: lowest placement level tried
    const int msh = N >= 65536 ? 2 : 0;  // large graphs: a popcount prefix per four bitmap words
    p.mpref_shift = msh;
    const int W0 = W;
    // small pairs (row index in 11 bits, degree in 5): 16-bit adjacency entries, level 0 only
    const bool can16 = cap_nodes + cap_nodes / 8 <= 2048 && maxdeg <= 31 && !getenv("RECOM_B200_NBR32");
    for (;;) {
      p.nbr16 = (can16 && level == 0) ? 1 : 0;
      total = rc::wilson_layout(nullptr, nullptr, N, cap_nodes, cap_edges, R, W, sh, level, msh, nullptr, nullptr, &glob, &cold,
                                p.nbr16 ? 2 : 4);
      // Level 0 takes what fits (with fewer walkers if need be).  Level 1 (walker state in shared
      // memory, adjacency in L2) only where it still leaves room for 3 CTAs per SM: once a walk
      // step is a round trip to L2 / HBM anyway, what counts is walks in flight, and level 2
      // (shared memory holds the membership bitmap alone) runs 4 CTAs of 256 threads per SM.
      // Measured at config 4 (200 k nodes, 8 k-node pairs, steps/s): level 1, 1 CTA x 4 walkers
      // 5.6 k; level 1, 2 x 2: 7.7 k; level 2, 4 x 4: 8.0-8.2 k; 5 x 192 threads 7.8 k;
      // 6 x 160: 5.8 k; 7 x 128: 5.0 k (the blocks of that many CTAs thrash L2); 8 walkers 6.6 k.
      if (level == 0 && total <= smem_max) break;
      if (level == 1 && total <= smem_max && 3 * (total + 1024) <= (size_t)prop.sharedMemPerMultiprocessor) break;
      if (level == 2) break;
      if (level == 0 && cfg->wilson_walkers == 0 && W > 2) { W >>= 1; continue; }
      ++level;
      W = W0;
    }
    p.walkers = W;
    p.wilson_level = level;
    split = total;
    total += glob;
    e->cold_bytes = (cold + 255) & ~size_t(255);
  } else {
    total = rc::smem_layout(nullptr, nullptr, N, cap_nodes, cap_edges, R, nullptr, nullptr, &split, &cold);
    e->cold_bytes = (cold + 255) & ~size_t(255);
  }
  // a merged pair that does not fit an SM's shared memory keeps its working arrays in a
  // per-CTA global scratch block instead (L2 resident): same code, generic pointers
  e->spill = wilson ? p.wilson_level > 0 : total > smem_max;
  e->smem = e->spill ? split : total;
  e->spill_bytes = e->spill ? ((total - split + 255) & ~size_t(255)) : 0;
  if (e->smem > smem_max) {
    rc_engine_destroy(e);
    return fail(RC_ERR_CAPACITY, "graph of %d nodes needs %zu B of shared memory for the membership bitmap (> %zu)",
                N, e->smem, smem_max);
  }
  e->threads = cfg->threads_per_cta;
  if (e->threads == 0) {
    // Auto: the step is latency bound, so small merged pairs want MORE resident CTAs rather
    // than more threads each (measured: Grid 20x20 runs 2.2x faster with 8 CTAs x 128 threads
    // than with 2 x 512).  About 4 edges per thread, at most 8 CTAs per SM, and no more
    // threads than the register file (64 regs x 1024 threads) leaves for that many CTAs.
    e->threads = wilson ? 256 : 512;
    if (!e->spill) {
      int c_smem = (int)((size_t)prop.sharedMemPerMultiprocessor / (e->smem + 1024));
      if (c_smem > 8) c_smem = 8;
      if (c_smem < 1) c_smem = 1;
      int by_regs = (1024 / c_smem) / 32 * 32;
      int by_work = ((cap_edges / 4) + 31) / 32 * 32;
      int t = by_work < by_regs ? by_work : by_regs;
      if (t < 64) t = 64;
      if (t > (wilson ? 256 : 512)) t = wilson ? 256 : 512;
      e->threads = t;
    }
    if (wilson && e->threads < 32 * p.walkers) e->threads = 32 * p.walkers;  // one warp per walker
  }
  if (wilson && e->threads < 32 * p.walkers) {
    rc_engine_destroy(e);
    return fail(RC_ERR_INVALID_ARG, "threads_per_cta must be at least 32 x wilson_walkers (%d)", 32 * p.walkers);
  }
  if (e->threads < 64 || e->threads > (wilson ? 256 : 512) || e->threads % 32) {
    rc_engine_destroy(e);
    return fail(RC_ERR_INVALID_ARG, "threads_per_cta must be a multiple of 32 in [64, %d] (0 = automatic)", wilson ? 256 : 512);
  }
  // two builds of the chain-step kernel: the plain one, and one with shared retries (CTAs without a
  // chain draw the later trees of steps in flight), which costs the step path ~20 % in registers
  // and barriers and is launched where it pays: few steps per launch or fewer chains than CTAs
  e->kernel = wilson ? (const void*)rc::recom_step_kernel<true, false, false>  // the placement level is a run-time parameter
                     : (e->spill ? (const void*)rc::recom_step_kernel<false, true, false> : (const void*)rc::recom_step_kernel<false, false, false>);
  e->kernel_share = wilson ? (const void*)rc::recom_step_kernel<true, false, true>
                           : (e->spill ? (const void*)rc::recom_step_kernel<false, true, true> : (const void*)rc::recom_step_kernel<false, false, true>);
  CUDA_OK_E(cudaFuncSetAttribute(e->kernel_share, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
  if (const char* sm = getenv("RECOM_B200_SHARE")) e->share_mode = sm[0] == '0' ? 0 : sm[0] == '1' ? 1 : sm[0] == '3' ? 3 : 2;
  p.help_win = rc::kHelpWin;
  if (const char* hw = getenv("RECOM_B200_HELP_WIN")) {  // test knob: a tiny window makes every shared step slide it
    const int v = atoi(hw);
    if (v >= 1 && v <= rc::kHelpWin) p.help_win = v;
  }
  p.recruit_after = 128;     // tuning knobs of the shared retries (recom_kernels.cu long_retry / shared_retry)
  p.units_per_recruit = 1;
  p.help_after = rc::kHelpAfter;
  p.recruit_mult = 1;
  if (const char* v = getenv("RECOM_B200_HELP_AFTER")) p.help_after = std::max(1, atoi(v));
  if (const char* v = getenv("RECOM_B200_RECRUIT_MULT")) p.recruit_mult = std::max(1, atoi(v));
  if (const char* v = getenv("RECOM_B200_RECRUIT_AFTER")) p.recruit_after = std::max(1, atoi(v));
  if (const char* v = getenv("RECOM_B200_UNITS_PER_RECRUIT")) p.units_per_recruit = std::max(1, atoi(v));
  CUDA_OK_E(cudaFuncSetAttribute(e->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
  CUDA_OK_E(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->ctas_per_sm, e->kernel, e->threads, e->smem));
  if (e->ctas_per_sm < 1) { rc_engine_destroy(e); return fail(RC_ERR_CAPACITY, "kernel does not fit an SM"); }
  // persistent grid: every CTA resident; they pop ready chains from a ring (recom_step_kernel)
  e->grid = e->ctas_per_sm * e->n_sms;
  if (e->spill && !wilson) {
    // MST spill mode: the Boruvka rounds stream the CTA's scratch block again and again; keep the
    // blocks of the resident CTAs inside L2 (126 MB) rather than have more CTAs wait on DRAM.
    // (Not for Wilson: its walks are latency bound, more of them in flight win - see above.)
    const size_t per_cta = e->spill_bytes + (e->cold_bytes + (size_t)(cap_nodes + cap_edges) * 4) / 4;  // the streamed arrays are touched once
    size_t l2_budget = (size_t)96 << 20;
    if (const char* v = getenv("RECOM_B200_L2_FIT_MB")) l2_budget = (size_t)std::max(1, atoi(v)) << 20;  // dev
    long long fit = (long long)l2_budget / (long long)(per_cta ? per_cta : 1);
    fit = fit / e->n_sms * e->n_sms;
    if (fit < e->n_sms) fit = e->n_sms;
    if (e->grid > fit) e->grid = (int)fit;
  }
  {
    // fewer chains than CTAs: the spare CTAs have no chain of their own, but with shared retries
    // (recom_kernels.cu) they draw the later spanning trees of the steps in flight - what a
    // single-chain MarkovChain driver gets out of the GPU; without them they would only wait
    const bool share = !(cfg->allow_pair_reselection || cfg->proposal_kind == RC_PROPOSAL_REVERSIBLE || cfg->node_repeats < 1);
    e->capacity = e->grid;
    if (!share && e->grid > cfg->n_chains) e->grid = cfg->n_chains;
  }
  // scratch
  size_t b_nodes = (size_t)e->grid * cap_nodes * sizeof(int32_t);
  size_t b_eids = (size_t)e->grid * cap_edges * sizeof(int32_t);
  size_t b_bad = cfg->allow_pair_reselection ? (size_t)e->grid * 2048 * sizeof(uint32_t) : 0;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  // pop/push indices, phase timers, progress words, ring of ready chains, CTAs counted per SM
  // the ring of ready chains has a slot for every chain AND for every waiting CTA: two tickets
  // must never wait on the same slot (with shared retries the grid may exceed the chain count)
  const int ring_cap = cfg->n_chains > e->grid ? cfg->n_chains : e->grid;
  size_t b_prog = 512 + ((size_t)cfg->n_chains + (size_t)ring_cap) * sizeof(int32_t) + 1024 * sizeof(int32_t);
  // + shared-retry board: one record and one window of status words per chain
  const size_t o_help = (b_prog + 63) & ~size_t(63);
  const size_t o_hstat = o_help + (size_t)cfg->n_chains * sizeof(rc::HelpRec);
  const size_t o_hbits = o_hstat + (size_t)cfg->n_chains * rc::kHelpWin * sizeof(unsigned long long);
  b_prog = o_hbits + ((size_t)cfg->n_chains / 32 + 1) * sizeof(uint32_t);
  size_t b_spill = (size_t)e->grid * e->spill_bytes;
  size_t b_cold = (size_t)e->grid * e->cold_bytes;
  e->scratch_bytes = al(b_nodes) + al(b_eids) + al(b_bad) + al(b_prog) + al(b_spill) + al(b_cold);
  CUDA_OK_E(cudaMalloc(&e->d_scratch, e->scratch_bytes));
  char* base = (char*)e->d_scratch;
  p.nodes = (int32_t*)base; base += al(b_nodes);
  p.eids = (int32_t*)base; base += al(b_eids);
  p.badpairs = b_bad ? (uint32_t*)base : nullptr; base += al(b_bad);
  p.work_counter = (unsigned long long*)base;
  p.timers = (unsigned long long*)(base + 256);
  p.progress = (int32_t*)(base + 512);
  p.ring = p.progress + cfg->n_chains;
  p.ring_cap = ring_cap;
  p.sm_slots = p.ring + ring_cap;
  p.idle = (int32_t*)(base + 64);
  p.steps_done = (unsigned long long*)(base + 128);
  p.help = (rc::HelpRec*)(base + o_help);
  p.help_status = (unsigned long long*)(base + o_hstat);
  p.help_bits = (uint32_t*)(base + o_hbits);
  p.help_count = (int32_t*)(base + 192);
  p.share = (cfg->allow_pair_reselection || cfg->proposal_kind == RC_PROPOSAL_REVERSIBLE || cfg->node_repeats < 1) ? 0 : 1;
  base += al(b_prog);
  p.spill = b_spill ? (unsigned char*)base : nullptr;
  p.spill_stride = (int64_t)e->spill_bytes;
  base += al(b_spill);
  p.cold = b_cold ? (unsigned char*)base : nullptr;
  p.cold_stride = (int64_t)e->cold_bytes;
  e->sched_bytes = b_prog;
  CUDA_OK_E(cudaEventCreate(&e->ev0));
  CUDA_OK_E(cudaEventCreate(&e->ev1));
  *out = e;
  return RC_OK;
}

int rc_engine_bind_state(RcEngine* e, const RcState* st) {
  if (!e || !st) return fail(RC_ERR_INVALID_ARG, "null argument");
  if (!st->d_assignment || !st->d_tally || !st->d_cut_mask || !st->d_cut_count || !st->d_step_no ||
      !st->d_proposal_no || !st->d_status || !st->d_flags || !st->d_counters)
    return fail(RC_ERR_INVALID_ARG, "RcState has a null buffer");
  int64_t as, ms;
  RcGraph g{}; g.n_nodes = e->p.g.N; g.n_edges = e->p.g.E;
  rc_state_strides(&g, &e->p.cfg, &as, &ms);
  if (st->assign_stride < as || st->assign_stride % 16 || st->mask_stride < ms || st->mask_stride % 4)
    return fail(RC_ERR_INVALID_ARG, "state strides must be >= rc_state_strides() and 16-byte multiples");
  if (((uintptr_t)st->d_assignment & 15) || ((uintptr_t)st->d_cut_mask & 15))
    return fail(RC_ERR_INVALID_ARG, "d_assignment and d_cut_mask must be 16-byte aligned");
  e->p.st = *st;
  e->bound = true;
  return RC_OK;
}

static int init_state_impl(RcEngine* e, void* stream, int keep_counters) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  CUDA_OK(cudaSetDevice(e->device));
  rc::Params p = e->p;
  p.rp = RcReplay{};
  p.keep_counters = keep_counters;
  rc::init_state_kernel<<<e->p.cfg.n_chains, 256, 0, (cudaStream_t)stream>>>(p);
  e->launches++;
  CUDA_OK(cudaGetLastError());
  return RC_OK;
}

int rc_engine_init_state(RcEngine* e, void* stream) { return init_state_impl(e, stream, 0); }

int rc_engine_run(RcEngine* e, int64_t n_steps, void* stream) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  if (n_steps <= 0) return RC_OK;
  CUDA_OK(cudaSetDevice(e->device));
  rc::Params p = e->p;
  p.rp = RcReplay{};
  p.n_steps = n_steps;
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_OK(cudaEventRecord(e->ev0, s));
  CUDA_OK(cudaMemsetAsync(p.work_counter, 0, e->sched_bytes, s));
  void* args[] = {(void*)&p};
  // shared retries where the tail of the launch matters: few steps per chain, or spare CTAs
  const int n_chains = e->p.cfg.n_chains;
  // Which build?  Shared retries cut the tail of a launch (config 3, 1024 chains x 32 steps: 0.26 M ->
  // 1.1 M steps/s; the slowest of 5 launches of 512 steps: -30 %) and cost the step path 12-17 %
  // (a bigger kernel, out-of-line retries); with fewer chains than CTAs the polling helpers cost
  // more than the rare stuck step returns.  How heavy the tail is depends on the workload (how
  // rare balanced cuts are), so the engine measures: launches of the same length are timed with
  // CUDA events, each build is tried twice, then the faster one is used and the other one tried
  // again every 16th launch.  The chains are bit-identical either way.  RECOM_B200_SHARE=0/1 forces it.
  for (auto& o : e->obs) {  // harvest the launches that have finished since
    if (!o.pending || cudaEventQuery(o.b) != cudaSuccess) continue;
    float ms = 0.f;
    if (o.steps == e->stat_steps && cudaEventElapsedTime(&ms, o.a, o.b) == cudaSuccess) {
      const int v = o.share ? 1 : 0;
      e->ms_avg[v] = e->n_obs[v] == 0 ? ms : 0.7 * e->ms_avg[v] + 0.3 * ms;
      e->n_obs[v]++;
    }
    o.pending = false;
  }
  if (e->stat_steps != n_steps) {
    e->stat_steps = n_steps;
    e->n_obs[0] = e->n_obs[1] = e->n_tried[0] = e->n_tried[1] = e->since_explore = 0;
  }
  bool share = false;
  if (p.share && e->share_mode != 0 && n_chains >= e->grid) {
    if (e->share_mode == 1) share = true;
    else if (e->n_tried[0] < 2) share = false;
    else if (e->n_tried[1] < 2) share = true;
    else if (e->n_obs[0] > 0 && e->n_obs[1] > 0) {
      share = e->ms_avg[1] < e->ms_avg[0];
      if (++e->since_explore >= 16) { e->since_explore = 0; share = !share; }
    }
  } else if (p.share && e->share_mode == 1) {
    share = true;
  }
  if (e->share_mode == 3) { share = true; p.share = 0; }  // dev: the shared-retry build with the protocol switched off
  e->n_tried[share ? 1 : 0]++;
  RcEngine::Obs* slot = nullptr;
  for (auto& o : e->obs)
    if (!o.pending) { slot = &o; break; }
  if (slot) {
    if (!slot->a) { CUDA_OK(cudaEventCreate(&slot->a)); CUDA_OK(cudaEventCreate(&slot->b)); }
    CUDA_OK(cudaEventRecord(slot->a, s));
  }
  const int grid = share ? e->grid : (n_chains < e->grid ? n_chains : e->grid);
  if (share) {
    // the shared-retry build reads Params from constant memory in its out-of-line parts
    // (rc::g_params); launches that write it are ordered by an event per device
    std::lock_guard<std::mutex> lock(g_const_mutex);
    cudaEvent_t& ev = g_const_event[e->device & 63];
    if (!ev) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    else CUDA_OK(cudaStreamWaitEvent(s, ev, 0));
    CUDA_OK(cudaMemcpyToSymbolAsync(rc::g_params, &p, sizeof(p), 0, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaLaunchKernel(e->kernel_share, dim3(grid), dim3(e->threads), args, e->smem, s));
    CUDA_OK(cudaEventRecord(ev, s));
  } else {
    CUDA_OK(cudaLaunchKernel(e->kernel, dim3(grid), dim3(e->threads), args, e->smem, s));
  }
  e->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(e->ev1, s));
  if (slot) {
    CUDA_OK(cudaEventRecord(slot->b, s));
    slot->share = share; slot->steps = n_steps; slot->pending = true;
  }
  e->timed = true;
  return RC_OK;
}

int rc_engine_step_host(RcEngine* e, const rc_label_t* h_assignment, int64_t h_stride, int64_t n_steps,
                        rc_label_t* h_assignment_out, int64_t* h_tally_out, int32_t* h_cut_count_out,
                        int32_t* h_status_out, void* stream) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  if (!h_assignment) return fail(RC_ERR_INVALID_ARG, "null h_assignment");
  const int N = e->p.g.N, n = e->p.cfg.n_chains, k = e->p.cfg.n_districts;
  if (h_stride < N) return fail(RC_ERR_INVALID_ARG, "h_stride must be >= n_nodes");
  CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  const RcState& st = e->p.st;
  constexpr size_t LB = sizeof(rc_label_t);
  if (h_stride == st.assign_stride)  // same pitch on both sides: one contiguous copy
    CUDA_OK(cudaMemcpyAsync(st.d_assignment, h_assignment, LB * (size_t)h_stride * n, cudaMemcpyHostToDevice, s));
  else
    CUDA_OK(cudaMemcpy2DAsync(st.d_assignment, LB * (size_t)st.assign_stride, h_assignment, LB * (size_t)h_stride,
                              LB * (size_t)N, (size_t)n, cudaMemcpyHostToDevice, s));
  int rc = init_state_impl(e, stream, 1);
  if (rc) return rc;
  rc = rc_engine_run(e, n_steps, stream);
  if (rc) return rc;
  if (h_assignment_out && h_stride == st.assign_stride)
    CUDA_OK(cudaMemcpyAsync(h_assignment_out, st.d_assignment, LB * (size_t)h_stride * n, cudaMemcpyDeviceToHost, s));
  else if (h_assignment_out)
    CUDA_OK(cudaMemcpy2DAsync(h_assignment_out, LB * (size_t)h_stride, st.d_assignment, LB * (size_t)st.assign_stride,
                              LB * (size_t)N, (size_t)n, cudaMemcpyDeviceToHost, s));
  if (h_tally_out)
    CUDA_OK(cudaMemcpyAsync(h_tally_out, st.d_tally, sizeof(int64_t) * (size_t)n * k, cudaMemcpyDeviceToHost, s));
  if (h_cut_count_out)
    CUDA_OK(cudaMemcpyAsync(h_cut_count_out, st.d_cut_count, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (h_status_out)
    CUDA_OK(cudaMemcpyAsync(h_status_out, st.d_status, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return RC_OK;
}

static int seed_impl(RcEngine* e, int32_t max_attempts, const RcReplay* rp, void* stream) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  if (e->p.cap_nodes < e->p.g.N || e->p.cap_edges < e->p.g.E)
    return fail(RC_ERR_CAPACITY, "seed plans split the whole graph: create the engine with cap_nodes = %d, cap_edges = %d",
                e->p.g.N, e->p.g.E);
  if (e->p.cfg.n_districts >= RC_SEED_REMAINING)
    return fail(RC_ERR_INVALID_ARG, "seed plans need n_districts <= %d", RC_SEED_REMAINING - 1);
  if (e->p.g.R > 0) return fail(RC_ERR_INVALID_ARG, "seed plans are drawn without region surcharges");
  if (rp && (rp->n_steps != e->p.cfg.n_districts - 1 || !rp->steps || !rp->weight_rank ||
             rp->chain < 0 || rp->chain >= e->p.cfg.n_chains))
    return fail(RC_ERR_INVALID_ARG, "seed replay needs n_districts - 1 recorded splits with weight ranks");
  CUDA_OK(cudaSetDevice(e->device));
  rc::Params p = e->p;
  p.rp = rp ? *rp : RcReplay{};
  p.seed_max_attempts = max_attempts;
  const void* kern = e->spill ? (const void*)rc::seed_kernel<true> : (const void*)rc::seed_kernel<false>;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
  void* args[] = {(void*)&p};
  CUDA_OK(cudaLaunchKernel(kern, dim3(rp ? 1 : e->grid), dim3(e->threads), args, e->smem, (cudaStream_t)stream));
  e->launches++;
  CUDA_OK(cudaGetLastError());
  return init_state_impl(e, stream, 2);
}

int rc_engine_seed(RcEngine* e, int32_t max_attempts, void* stream) { return seed_impl(e, max_attempts, nullptr, stream); }

int rc_engine_seed_replay(RcEngine* e, int32_t max_attempts, const RcReplay* rp, void* stream) {
  if (!rp) return fail(RC_ERR_INVALID_ARG, "null replay");
  return seed_impl(e, max_attempts, rp, stream);
}

int rc_engine_replay(RcEngine* e, const RcReplay* rp, void* stream) {
  if (!e || !e->bound || !rp) return fail(RC_ERR_INVALID_ARG, "engine has no bound state / null replay");
  if (rp->n_steps <= 0 || !rp->steps) return fail(RC_ERR_INVALID_ARG, "empty replay");
  if (rp->chain < 0 || rp->chain >= e->p.cfg.n_chains) return fail(RC_ERR_INVALID_ARG, "bad chain index");
  if (e->p.cfg.tree_kind == RC_TREE_MST && !rp->weight_rank) return fail(RC_ERR_INVALID_ARG, "MST replay needs weight_rank");
  if (e->p.cfg.tree_kind == RC_TREE_WILSON && (!rp->wilson_root || !rp->stack_ptr || !rp->stack_val))
    return fail(RC_ERR_INVALID_ARG, "Wilson replay needs wilson_root, stack_ptr and stack_val");
  CUDA_OK(cudaSetDevice(e->device));
  rc::Params p = e->p;
  p.rp = *rp;
  p.n_steps = rp->n_steps;
  void* args[] = {(void*)&p};
  CUDA_OK(cudaLaunchKernel(e->kernel, dim3(1), dim3(e->threads), args, e->smem, (cudaStream_t)stream));
  e->launches++;
  CUDA_OK(cudaGetLastError());
  return RC_OK;
}

int rc_engine_histograms(RcEngine* e, int64_t* d_cut_hist, int32_t n_cut_bins, int64_t* d_dev_hist,
                         int32_t n_dev_bins, double dev_bin_width, void* stream) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  if ((d_cut_hist && n_cut_bins < 1) || (d_dev_hist && (n_dev_bins < 1 || !(dev_bin_width > 0))))
    return fail(RC_ERR_INVALID_ARG, "bad histogram shape");
  CUDA_OK(cudaSetDevice(e->device));
  const int n = e->p.cfg.n_chains;
  rc::histogram_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      e->p, (unsigned long long*)d_cut_hist, n_cut_bins, (unsigned long long*)d_dev_hist, n_dev_bins, dev_bin_width);
  e->launches++;
  CUDA_OK(cudaGetLastError());
  return RC_OK;
}

int rc_engine_tally(RcEngine* e, const int32_t* d_column, int64_t* d_out, void* stream) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  if (!d_column || !d_out) return fail(RC_ERR_INVALID_ARG, "null argument");
  CUDA_OK(cudaSetDevice(e->device));
  rc::tally_column_kernel<<<e->p.cfg.n_chains, 256, 0, (cudaStream_t)stream>>>(e->p, d_column, (long long*)d_out);
  e->launches++;
  CUDA_OK(cudaGetLastError());
  return RC_OK;
}

int rc_engine_region_splits(RcEngine* e, const int32_t* d_region, int32_t n_regions, int32_t* d_out, void* stream) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  if (!d_region || !d_out || n_regions < 1) return fail(RC_ERR_INVALID_ARG, "null argument or no regions");
  CUDA_OK(cudaSetDevice(e->device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, e->device));
  const size_t words = (size_t)n_regions * (1 + (e->p.cfg.n_districts + 31) / 32);
  const size_t smem = words * sizeof(uint32_t);
  if (smem + 1024 > prop.sharedMemPerBlockOptin)
    return fail(RC_ERR_CAPACITY, "%d regions x %d districts need %zu bytes of shared memory", n_regions,
                e->p.cfg.n_districts, smem);
  CUDA_OK(cudaFuncSetAttribute(rc::region_splits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = e->p.cfg.n_chains < 4 * e->n_sms ? e->p.cfg.n_chains : 4 * e->n_sms;
  rc::region_splits_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(e->p, d_region, n_regions, d_out);
  e->launches++;
  CUDA_OK(cudaGetLastError());
  return RC_OK;
}

int rc_engine_contiguity(RcEngine* e, int32_t* d_out, void* stream) {
  if (!e || !e->bound) return fail(RC_ERR_INVALID_ARG, "engine has no bound state");
  if (!d_out) return fail(RC_ERR_INVALID_ARG, "null argument");
  CUDA_OK(cudaSetDevice(e->device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, e->device));
  const size_t bytes = (size_t)e->p.g.N * sizeof(int);
  const bool in_smem = bytes + 1024 <= prop.sharedMemPerBlockOptin;
  int grid = e->p.cfg.n_chains < 2 * e->n_sms ? e->p.cfg.n_chains : 2 * e->n_sms;
  if (in_smem) {
    CUDA_OK(cudaFuncSetAttribute(rc::contiguity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  } else if (!e->d_cc_scratch) {
    CUDA_OK(cudaMalloc((void**)&e->d_cc_scratch, (size_t)2 * e->n_sms * bytes));
  }
  rc::contiguity_kernel<<<grid, 512, in_smem ? bytes : 0, (cudaStream_t)stream>>>(e->p, e->d_cc_scratch, in_smem ? 1 : 0, d_out);
  e->launches++;
  CUDA_OK(cudaGetLastError());
  return RC_OK;
}

int rc_engine_info(RcEngine* e, RcEngineInfo* out) {
  if (!e || !out) return fail(RC_ERR_INVALID_ARG, "null argument");
  out->cap_nodes = e->p.cap_nodes; out->cap_edges = e->p.cap_edges;
  out->threads_per_cta = e->threads; out->ctas_per_sm = e->ctas_per_sm; out->n_sms = e->n_sms;
  out->grid = e->grid; out->smem_bytes = (int64_t)e->smem; out->scratch_bytes = (int64_t)e->scratch_bytes;
  out->kernel_launches = e->launches;
  out->wilson_walkers = e->p.walkers; out->spill = wilson_spill_level(e);
  return RC_OK;
}

/* dev tool (not in the public header): cycles per phase accumulated by an RC_TIMERS build
 * during the last rc_engine_run */
int rc_debug_phase_cycles(RcEngine* e, unsigned long long* out16) {
  if (!e || !out16) return fail(RC_ERR_INVALID_ARG, "null argument");
  CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaMemcpy(out16, e->p.timers, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return RC_OK;
}

int rc_engine_last_run_ms(RcEngine* e, float* ms) {
  if (!e || !ms) return fail(RC_ERR_INVALID_ARG, "null argument");
  if (!e->timed) return fail(RC_ERR_INVALID_ARG, "no rc_engine_run has been issued");
  CUDA_OK(cudaSetDevice(e->device));
  CUDA_OK(cudaEventSynchronize(e->ev1));
  CUDA_OK(cudaEventElapsedTime(ms, e->ev0, e->ev1));
  return RC_OK;
}

}  // extern "C"
