/*
 * recom_b200.h - C ABI of the B200 batched ReCom engine.
 *
 * GerryChain has no FFI: its plug points are Python callables.  This
 * header is what a ctypes stub on the reference side binds (INTEGRATION.md shows the
 * stub).  Every entry point cites the reference interface it replaces, relative to
 * /root/reference/gerrychain/.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ or torch types; no exceptions cross the ABI
 *  - every function returns an int status (RC_OK == 0); rc_last_error() is thread local
 *  - pointers named d_* are DEVICE pointers owned by the caller (torch tensors'
 *    data_ptr() on the Python side); pointers named h_* are HOST pointers;
 *    the library owns only what rc_engine_create allocates (graph copy + scratch)
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    kernels are asynchronous on it unless the function says it synchronises
 *  - one host thread per engine
 *
 * The same structs are used by the CPU oracle (oracle/recom_oracle.h) so that a replay
 * record or a configuration means the same thing on both sides.
 */
#ifndef RECOM_B200_H
#define RECOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_ABI_VERSION 1

/* ---- district labels ----------------------------------------------------------------
 * One label per node in the assignment rows.  The default build (librecom_b200.so) keeps
 * them in one byte (plans with up to 255 districts: a 10 KB row for a 10k-node graph);
 * the same sources compiled with -DRC_LABEL_BITS=16 (librecom_b200_l16.so, identical entry
 * points) carry 16-bit labels for plans with up to 32 767 districts.  rc_label_bits() says
 * which build is loaded; the host side (gerrychain_b200/_abi.py) picks the library by k. */
#ifndef RC_LABEL_BITS
#define RC_LABEL_BITS 8
#endif
#if RC_LABEL_BITS == 16
typedef uint16_t rc_label_t;
#define RC_MAX_DISTRICTS 32767
#elif RC_LABEL_BITS == 8
typedef uint8_t rc_label_t;
#define RC_MAX_DISTRICTS 255
#else
#error "RC_LABEL_BITS must be 8 or 16"
#endif

/* ---- status codes (per call and per chain) -------------------------------------- */
enum {
  RC_OK = 0,
  RC_ERR_INVALID_ARG = 1,   /* bad pointer / size / unsupported configuration         */
  RC_ERR_CUDA = 2,          /* a CUDA runtime call failed; see rc_last_error()        */
  RC_ERR_CAPACITY = 3,      /* merged pair larger than the engine's node/edge caps    */
  RC_ERR_MAX_ATTEMPTS = 4,  /* tree.py:718  RuntimeError("Could not find a possible cut after N attempts.") */
  RC_ERR_ALL_PAIRS_BAD = 5, /* proposals/tree_proposals.py:140-144  MetagraphError    */
  RC_ERR_DISCONNECTED = 6,  /* merged subgraph not connected (invalid input plan)     */
  RC_ERR_NO_CUT_EDGES = 7,  /* tree_proposals.py:106 random.choice(()) -> IndexError  */
  RC_ERR_REPLAY = 8,        /* replay record inconsistent with the state              */
  RC_ERR_NO_ROOT = 9,       /* tree.py:377 choice([]) : no tree node of degree > 1    */
  RC_ERR_INVALID_STREAK = 10, /* proposals kept failing the population Bounds          */
  RC_ERR_BALANCE = 11,      /* tree.py:1441-1446 BalanceError / PopulationBalanceError (seed plans) */
  RC_ERR_REVERSIBILITY = 12 /* tree_proposals.py:205-209, :257-261 ReversibilityError   */
};

/* seed plans (tree.py:971-1085): label of nodes not yet assigned, and the proposal-number
 * space of the seed splits in the random-draw schedule */
#define RC_SEED_REMAINING ((1 << RC_LABEL_BITS) - 1)
#define RC_SEED_PROPOSAL 0x40000000u

/* per-chain flag bits (RcState.d_flags), sticky */
enum {
  RC_FLAG_WARNED = 1 /* tree.py:703-710 BipartitionWarning threshold was reached     */
};

enum { RC_TREE_MST = 0,   /* tree.py:60-94  random_spanning_tree (+ networkx Kruskal) */
       RC_TREE_WILSON = 1 /* tree.py:97-132 uniform_spanning_tree                     */ };

enum { RC_PROPOSAL_RECOM = 0,      /* proposals/tree_proposals.py:36-146  recom              */
       RC_PROPOSAL_REVERSIBLE = 1  /* proposals/tree_proposals.py:149-267 reversible_recom
                                      (needs tree_kind = RC_TREE_WILSON)                  */ };

enum { RC_ACCEPT_ALWAYS = 0,   /* accept.py:15 always_accept                            */
       RC_ACCEPT_CUT_EDGE = 1  /* accept.py:19-35 cut_edge_accept: a valid proposal is taken
                                  with probability min(1, #cut edges before / after); a
                                  rejected one still counts as a step (chain.py:191-195)  */ };

enum { RC_CUT_UNIFORM = 0,      /* tree.py:540-545 random.choice(cuts)                 */
       RC_CUT_REGION_PREF = 1,  /* tree.py:548-578 power-set preference, then max weight */
       RC_CUT_MAX_WEIGHT = 2    /* tree.py:446-480 _max_weight_choice                  */ };

/* ---- graph: replaces Graph / FrozenGraph (graph/graph.py:51, :508) as input -------- */
typedef struct RcGraph {
  int32_t n_nodes;            /* N: node index = position in graph.nodes order         */
  int32_t n_edges;            /* E: undirected; edge id = rank of (min,max) pair       */
  const int32_t* row_ptr;     /* [N+1]                                                 */
  const int32_t* col_idx;     /* [2E] neighbours, each row ascending                   */
  const int32_t* slot_edge;   /* [2E] undirected edge id of each directed slot         */
  const int32_t* edge_uv;     /* [E][2] u < v                                          */
  const int32_t* pop;         /* [N] graph.nodes[n][pop_col], integers                 */
  int32_t n_region_cols;      /* R = len(region_surcharge) (tree.py:80)                */
  const int32_t* region_ids;  /* [R][N] dense code per region label; -1 == None        */
  const double* surcharges;   /* [R] region_surcharge.values() in dict order           */
} RcGraph;

/* ---- configuration: the kwargs of recom / bipartition_tree / the Bounds constraint - */
typedef struct RcConfig {
  int32_t n_districts;        /* k = len(partition)                                    */
  int32_t n_chains;           /* independent chains resident on this device            */
  double pop_target;          /* recom(pop_target=)   tree_proposals.py:39             */
  double epsilon;             /* recom(epsilon=)      tree_proposals.py:40             */
  int32_t node_repeats;       /* recom(node_repeats=) tree.py:586, :680-682            */
  int32_t max_attempts;       /* bipartition_tree(max_attempts=) tree.py:593; <=0: None */
  int32_t warn_attempts;      /* tree.py:594                                           */
  int32_t allow_pair_reselection; /* tree.py:595, tree_proposals.py:133-138            */
  int32_t tree_kind;          /* RC_TREE_*                                             */
  int32_t cut_policy;         /* RC_CUT_*                                              */
  double pop_lower;           /* Bounds lower: constraints/validity.py:88-91, bounds.py:25-28 */
  double pop_upper;           /* Bounds upper; lower > upper disables the check        */
  uint64_t seed;              /* Philox key; chain c uses subsequence c + chain_offset */
  int32_t chain_offset;       /* global id of local chain 0 (multi-GPU sharding)       */
  int32_t cap_nodes;          /* max nodes of a merged pair (0: engine picks)          */
  int32_t cap_edges;          /* max undirected edges of a merged pair (0: picks)      */
  int32_t threads_per_cta;    /* 0: engine picks                                       */
  int32_t max_invalid_streak; /* consecutive invalid proposals before RC_ERR_INVALID_STREAK (0: 1000) */
  int32_t accept_rule;        /* RC_ACCEPT_*: MarkovChain(accept=)  chain.py:192        */
  int32_t proposal_kind;      /* RC_PROPOSAL_*                                          */
  int32_t rev_M;              /* reversible_recom(M=): bound on the number of balanced cuts */
  int32_t wilson_walkers;     /* RC_TREE_WILSON: spanning trees t, t+1, ... of one bipartition_tree
                               * call (tree.py:679-718) drawn concurrently, one random walk per warp;
                               * the lowest tree number with a balanced cut wins, so the result does
                               * not depend on it.  1..8; 0 = engine picks                        */
} RcConfig;

/* ---- chain state, resident in HBM, owned by the caller --------------------------- */
typedef struct RcState {
  rc_label_t* d_assignment; /* [n_chains][assign_stride] district index per node
                              (partition/assignment.py:9 Assignment.mapping)          */
  int64_t* d_tally;        /* [n_chains][k]  updaters/tally.py:69 Tally               */
  uint32_t* d_cut_mask;    /* [n_chains][mask_stride] bit e set <=> edge e is cut
                              (updaters/cut_edges.py:99)                              */
  int32_t* d_cut_count;    /* [n_chains] len(partition["cut_edges"])                  */
  int64_t* d_step_no;      /* [n_chains] accepted steps so far (chain.py:193)         */
  uint32_t* d_proposal_no; /* [n_chains] proposals so far, valid or not (RNG counter) */
  int32_t* d_status;       /* [n_chains] RC_OK or the RC_ERR_* that stopped the chain */
  uint32_t* d_flags;       /* [n_chains] RC_FLAG_*                                    */
  int64_t* d_counters;     /* [n_chains][RC_N_COUNTERS] see below                     */
  int64_t assign_stride;   /* labels between chains in d_assignment (rows 16-byte aligned) */
  int64_t mask_stride;     /* uint32 words between chains in d_cut_mask (multiple of 4) */
  int32_t* d_last_pair;    /* [n_chains][2] optional (may be NULL): the district pair (a < b)
                              merged by the chain's last committed proposal, i.e. the keys of
                              its `flips` (tree.py:947-950); {-1,-1} before the first one   */
} RcState;

enum {
  RC_CTR_TREES = 0,      /* spanning trees drawn                                      */
  RC_CTR_ATTEMPTS = 1,   /* bipartition_tree attempts (tree.py:700)                   */
  RC_CTR_NODES = 2,      /* sum of n_sub over proposals (roofline bytes)  */
  RC_CTR_SLOTS = 3,      /* sum of d_sub (directed CSR slots of merged nodes)         */
  RC_CTR_INVALID = 4,    /* proposals rejected by the population bounds               */
  RC_CTR_RESELECT = 5,   /* ReselectException count                                   */
  RC_CTR_WALK = 6,       /* Wilson successor draws                                    */
  RC_CTR_ROUNDS = 7,     /* Boruvka rounds                                            */
  RC_N_COUNTERS = 8
};

/* ---- replay (TEST HOOK): identical injected randomness ---------------------------- */
typedef struct RcReplayStep {
  int32_t pair_edge;   /* edge id picked at tree_proposals.py:106                      */
  int32_t n_trees;     /* trees drawn by bipartition_tree in this step                 */
  int64_t first_tree;  /* index of this step's first tree in the per-tree arrays       */
  int32_t root;        /* node index chosen at tree.py:377 in the successful attempt   */
  int32_t cut_edge;    /* edge id of the chosen Cut.edge; -1: apply cfg.cut_policy     */
} RcReplayStep;

typedef struct RcReplay {
  int32_t n_steps;
  int32_t chain;                 /* local chain index the record is applied to          */
  const RcReplayStep* steps;     /* [n_steps]                                           */
  const double* weights;         /* MST: [n_trees_total][E] random_weight per edge id
                                    (read by the CPU oracle)                            */
  const uint32_t* weight_rank;   /* MST: [n_trees_total][E] rank of that weight among the
                                    tree's recorded weights, ties by edge id (read by the
                                    device engine: only the order matters, tree.py:91-93) */
  const int32_t* wilson_root;    /* Wilson: [n_trees_total] root (tree.py:112)          */
  const int64_t* stack_ptr;      /* Wilson: [n_trees_total][N+1] into stack_val         */
  const int32_t* stack_val;      /* Wilson: chosen neighbour node index per draw        */
  rc_label_t* assign_trace;      /* out, optional: [n_steps][N] assignment after step   */
  int32_t* info_trace;           /* out, optional: [n_steps][4] {trees used, #balanced
                                    cuts of the last tree, cut count after, status}    */
} RcReplay;

/* ---- engine ---------------------------------------------------------------------- */
typedef struct RcEngine RcEngine;

const char* rc_last_error(void);
int rc_abi_version(void);
int rc_label_bits(void); /* 8 or 16: width of rc_label_t in the loaded build */

/* Device-side sizes the caller must honour when allocating RcState buffers. */
int rc_state_strides(const RcGraph* g, const RcConfig* c, int64_t* assign_stride,
                     int64_t* mask_stride);

/* Copies the (host) graph to the device, sizes caps and scratch.  Replaces the implicit
 * graph capture of Partition.__init__ (partition/partition.py:122-151). */
int rc_engine_create(const RcGraph* h_graph, const RcConfig* cfg, int device,
                     RcEngine** out);
int rc_engine_destroy(RcEngine* e);

/* Remember the caller's state buffers (no copy). */
int rc_engine_bind_state(RcEngine* e, const RcState* state);

/* Full recompute of tally / cut mask / cut count from d_assignment and reset of the
 * counters: Tally._initialize_tally (updaters/tally.py:118-141) and the first
 * cut_edges scan (updaters/cut_edges.py:110-114). */
int rc_engine_init_state(RcEngine* e, void* stream);

/* Advance every chain by n_steps accepted steps: the body of MarkovChain.__next__
 * (chain.py:185-195) with proposal = recom (proposals/tree_proposals.py:36-146),
 * constraints = population Bounds, accept = always_accept (accept.py:15). */
int rc_engine_run(RcEngine* e, int64_t n_steps, void* stream);

/* The same call with HOST buffers: the whole reference-facing round trip of a batch of
 * chains.  Copies h_assignment ([n_chains][h_stride] labels, district index per node, only
 * the first N labels of a row are read) to the device, recomputes Tally / cut_edges
 * (rc_engine_init_state), advances every chain by n_steps accepted steps (rc_engine_run)
 * and copies back: the assignment (same layout, may alias h_assignment), Tally
 * [n_chains][k] int64, len(cut_edges) [n_chains] int32 and the per-chain status
 * [n_chains] int32 (any of the outputs may be NULL).  Host buffers should be pinned
 * (cudaHostAlloc / torch pin_memory) for the copies to be asynchronous; with h_stride equal
 * to the device pitch (rc_state_strides) each direction is one contiguous copy.  Synchronises
 * `stream` before returning.  This is what one MarkovChain batch costs a caller who keeps
 * its partitions on the host, as the reference does (chain.py:166-196). */
int rc_engine_step_host(RcEngine* e, const rc_label_t* h_assignment, int64_t h_stride,
                        int64_t n_steps, rc_label_t* h_assignment_out, int64_t* h_tally_out,
                        int32_t* h_cut_count_out, int32_t* h_status_out, void* stream);

/* Random seed plans for every chain: recursive_tree_part (tree.py:971-1085) with
 * method = partial(bipartition_tree, max_attempts=max_attempts) (the reference's default is
 * 10000) - k - 2 one-sided splits with a debt-tightened target, then a two-sided one -
 * followed by rc_engine_init_state.  Replaces Partition.from_random_assignment
 * (partition/partition.py:67-120).  The engine must have been created with cap_nodes >=
 * n_nodes and cap_edges >= n_edges (every split works on the whole remainder).  A chain
 * whose plan could not be completed carries the status word (RC_ERR_MAX_ATTEMPTS,
 * RC_ERR_BALANCE, ...). */
int rc_engine_seed(RcEngine* e, int32_t max_attempts, void* stream);

/* TEST HOOK: the same for chain rp->chain with recorded reference randomness, one
 * RcReplayStep per split (pair_edge unused). */
int rc_engine_seed_replay(RcEngine* e, int32_t max_attempts, const RcReplay* rp, void* stream);

/* TEST HOOK: apply recorded reference steps to one chain (device pointers inside). */
int rc_engine_replay(RcEngine* e, const RcReplay* d_replay_host_struct, void* stream);

/* Ensemble histograms over this device's chains (int64 counts, caller zeroes/accumulates):
 * cut_hist[min(cut_count, n_cut_bins-1)]++ ;
 * dev_hist[min(floor(max_d |pop_d - ideal| / ideal / dev_bin_width), n_dev_bins-1)]++
 * (constraints/validity.py:96-122 deviation_from_ideal). */
int rc_engine_histograms(RcEngine* e, int64_t* d_cut_hist, int32_t n_cut_bins,
                         int64_t* d_dev_hist, int32_t n_dev_bins, double dev_bin_width,
                         void* stream);

/* Per-district sums of an arbitrary integer node column for every chain:
 * d_out[chain][d] = sum of d_column[v] over the nodes v of district d (int64).  Replaces
 * Tally (updaters/tally.py:69-170) for columns other than the balanced population, e.g. the
 * vote totals of an Election updater (updaters/election.py). */
int rc_engine_tally(RcEngine* e, const int32_t* d_column, int64_t* d_out, void* stream);

/* How a region column (counties, municipalities ...) is split by every chain's plan.
 * d_region[v] = dense region id in [0, n_regions) of node v, negative = none.
 * d_out[chain] = { regions with a cut edge whose two ends lie in the region - the count
 *                  tally_region_splits / total_reg_splits report (updaters/county_splits.py:120-161),
 *                  regions whose nodes lie in more than one district - the counties
 *                  compute_county_splits marks other than NOT_SPLIT (updaters/county_splits.py:58-117),
 *                  sum over regions of the number of districts present in them }  (int32 x 3).
 * RC_ERR_CAPACITY when n_regions * (1 + ceil(k/32)) words exceed one CTA's shared memory. */
int rc_engine_region_splits(RcEngine* e, const int32_t* d_region, int32_t n_regions, int32_t* d_out,
                            void* stream);

/* Contiguity of every chain's plan (constraints/contiguity.py:169-181 contiguous, :226-241
 * contiguous_components): d_out[chain][d] = number of connected pieces of district d in the
 * graph without its cut edges (0: the district owns no node), int32 [n_chains][n_districts].
 * The plan is contiguous iff no entry exceeds 1.  A debug assertion for the step kernel too:
 * ReCom keeps plans contiguous by construction. */
int rc_engine_contiguity(RcEngine* e, int32_t* d_out, void* stream);

/* Introspection: caps chosen, shared memory per CTA, resident CTAs per SM, grid. */
typedef struct RcEngineInfo {
  int32_t cap_nodes, cap_edges, threads_per_cta, ctas_per_sm, n_sms, grid;
  int64_t smem_bytes, scratch_bytes;
  int64_t kernel_launches; /* launches of this library's kernels so far              */
  int32_t wilson_walkers;  /* RC_TREE_WILSON: trees walked concurrently per chain-step  */
  int32_t spill;           /* 1: a merged pair's working arrays live in global scratch  */
} RcEngineInfo;
int rc_engine_info(RcEngine* e, RcEngineInfo* out);

/* Device time (ms) of the last rc_engine_run, measured with CUDA events on its stream;
 * synchronises on the end event. */
int rc_engine_last_run_ms(RcEngine* e, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* RECOM_B200_H */
