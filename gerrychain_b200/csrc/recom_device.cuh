// recom_device.cuh - device-side building blocks of the batched ReCom engine (sm_100a).
//
// One CTA owns one chain-step.  Everything between "gather the merged pair" and "commit"
// lives in shared memory; HBM holds only the chain state and the read-only CSR.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/recom_b200.h"
#include "../../include/recom_rng.h"

namespace rc {

typedef uint16_t lid_t;  // local node / arc / edge-slot id inside one merged pair
constexpr int kMaxLocal = 32760;  // 2*(n-1) arcs must fit 16 bits

struct DevGraph {
  int32_t N, E, R;
  const int32_t* row_ptr;
  const int32_t* col_idx;
  const int32_t* slot_edge;
  const int32_t* edge_uv;
  const int32_t* pop;
  const int32_t* region_ids;  // [R][N]
  double surcharge[8];        // region_surcharge values
  uint64_t sfix[8];           // rc_surcharge_fix(surcharge[r])
  uint64_t smult;             // rc_key_mult(sum sfix, R); 0 when R == 0
};

// Shared retries (recom_kernels.cu shared_retry / help_once): CTAs with no ready chain compute
// spanning trees t, t + 1, ... of a chain-step another CTA is retrying (tree.py:679-718).
struct HelpRec {
  int32_t epoch;     // odd: help wanted; bumped by the owner on publish, window slide and retirement
  int32_t session;   // bumped per publish: same session = same merged pair (a helper's gather is still good)
  int32_t pair;      // a | b << 16
  uint32_t prop_no;
  uint32_t t_first;  // tree number of unit 0
  int32_t unit;      // trees per unit (1 for MST, the walker count for Wilson)
  int32_t limit;     // units >= limit lie beyond max_attempts
  int32_t base;      // first unit of the status window
  int32_t next;      // next unit to hand out (atomic)
  int32_t won;       // lowest unit with a balanced cut so far (atomicMin)
  int32_t want;      // CTAs that should leave their own chains for this step (grows with the units that have failed)
  int32_t pad[5];
};
constexpr int kHelpWin = 256;  // status words per chain: units [base, base + kHelpWin)

struct Params {
  DevGraph g;
  RcConfig cfg;
  RcState st;
  RcReplay rp;        // rp.n_steps > 0: replay mode (device pointers inside)
  int64_t n_steps;    // free-running: accepted steps per chain for this launch
  int32_t cap_nodes, cap_edges;
  // per-CTA global scratch (L2 resident)
  int32_t* nodes;     // [grid][cap_nodes] local -> global node
  int32_t* eids;      // [grid][cap_edges] local edge -> global edge id
  uint32_t* badpairs; // [grid][2048] bitset of (a*256+b), pair reselection only
  unsigned long long* work_counter;  // [2] persistent scheduling: pop index, push index
  int32_t* progress;      // [n_chains] steps of this launch completed per chain
  int32_t* ring;          // [ring_cap] ready chains (chain id + 1; 0 = empty)
  int32_t ring_cap;
  int32_t* sm_slots;      // [1024] CTAs of this launch counted per SM (%smid): spreads lone-thread work over the SM's schedulers
  int32_t keep_counters;  // init_state_kernel: 1 = leave step/proposal numbers and counters
                          // alone; 2 = after seeding: also keep a failed chain's status
  int32_t seed_max_attempts;
  unsigned long long* timers;  // RC_TIMERS builds: cycles per phase (thread 0 of every CTA)
  unsigned char* spill;   // spill mode: [grid][spill_stride] working arrays in global memory
  int64_t spill_stride;
  unsigned char* cold;    // Wilson: [grid][cold_stride] arrays only streamed through (pop, eu, ev, ecls)
  int64_t cold_stride;
  int32_t walkers;        // Wilson: spanning trees drawn concurrently per chain-step (lanes of one warp)
  int32_t nbr_shift;      // Wilson: log2 of the adjacency slots per node kept in the hot tier
  int32_t nbr16;          // Wilson: 16-bit adjacency entries (neighbour 11 bits | its degree << 11): small pairs, level 0
  int32_t wilson_level;   // Wilson: 0 = state and adjacency in shared memory, 1 = adjacency in L2, 2 = both in L2
  int32_t mpref_shift;    // Wilson: Smem.msh
  HelpRec* help;                       // [n_chains]
  unsigned long long* help_status;     // [n_chains][kHelpWin]
  uint32_t* help_bits;                 // [n_chains / 32 + 1] bit c: chain c has a live help record
  int32_t* help_count;                 // live help records
  int32_t* idle;                       // CTAs of this launch that have no chain to step
  unsigned long long* steps_done;      // chain-steps completed in this launch
  int32_t share;                       // 0: no shared retries
  int32_t help_win;                    // units per status window (<= kHelpWin)
  int32_t help_after;                  // failed trees after which a step leaves the common path (long_retry)
  int32_t recruit_mult;                // CTAs recruited per units_per_recruit failed units
  int32_t recruit_after;               // failed trees from which a step is shared although every CTA has a chain of its own
  int32_t units_per_recruit;           // ... and takes one more CTA off its own chains per this many failed units
};

// ---- shared-memory carve-up -------------------------------------------------------
constexpr int kMiscInts = 160;
struct Smem {
  int32_t* misc;     // [kMiscInts]
  // whole step
  uint32_t* mbits;   // [NW] membership bitmap of the merged pair over global node ids
  lid_t* mpref;      // [NW >> msh] exclusive prefix popcount of every (1 << msh)-th word: local id = prefix + popc(below)
  int msh;           // 0: a prefix per bitmap word; 2: per four words (Wilson on large graphs: 3/4 of the array saved)
  int32_t* pop;      // [CN]   population of local node i                      (cold tier)
  uint32_t* euv;     // [CE]   local edge j = (u | v << 16), u < v, ascending edge id  (cold tier)
  uint32_t* tflag;   // [CE/32] tree-edge bitmap over local edges
  uint32_t* oldb;    // [CN/32] bit i: local node i carries the higher district id b
  uint32_t* newb;    // [CN/32] the same after the proposed cut
  uint8_t* ecls;     // [CE] bitmask of region columns the edge crosses (R > 0 only)
  // Wilson only: adjacency of the merged pair, rows in ascending neighbour id
  uint32_t* nbr;     // (uint16_t entries, neighbour | degree << 11, when Params.nbr16)
                     // [(CN + CN/8) << nbr_shift] rows of D = 1 << nbr_shift slots: row u < n belongs to
                     // node u; an entry is a neighbour (bits 0-15) | degree OF THAT NEIGHBOUR << 16,
                     // ascending; a node of degree > D keeps D - 1 entries and a link (slot D - 1) to
                     // a spare row (index >= n) that holds the next ones, and so on
  uint16_t* wdeg;    // [CN] degree inside the pair
  // spanning-tree phase
  uint32_t* key;     // [CE] 32-bit weight key of local edge j
  uint32_t* lab;     // [CE] live edge list: endpoints as current component ids (a | b << 16)
  lid_t* lj;         // [CE] live edge list: local edge index
  // uniform-spanning-tree phase (Wilson): one state word per node and walker, see wilson_walk
  uint16_t* wst;     // [W][CN] successor (bits 0-14) | in the tree (15)
  uint32_t* wtmp;    // [CN] scratch of the gather / adjacency build; aliases wsum
  int32_t* wsum;     // [CN] population of the subtree below node u (tree being evaluated)
  uint32_t* whas;    // [CN/32] bit u: node u has a child in the tree being evaluated
  int32_t* wmeta;    // [4][32] per walker: root, draws, error, spare
  uint32_t* wring;   // [W][128] per walker: ring of successor draws (random words), see wilson_walk_warp
  uint32_t* best;    // [CN] minimum key over the live edges of a component
  uint32_t* bpos;    // [CN] (local edge << 16 | live-list position) of that minimum
  lid_t* hook;       // [CN] component merge forest
  lid_t* vlA;        // [CN/2+2] live components, ping
  lid_t* vlB;        // [CN/2+2] live components, pong
  // Euler-tour phase (aliases the spanning-tree phase)
  lid_t* te;         // [CN] local edge index of tree edge t (ascending)
  uint32_t* tuv;     // [CN] its endpoints (u | v << 16): the tour never goes back to the edge list
  uint32_t* head;    // [CN] first out-arc of a node (linked adjacency)
  lid_t* nxt;        // [2CN] next out-arc of the same node
  lid_t* succ;       // [2CN] tour successor of an arc
  int32_t* cur;      // [CN] population below tree edge t (aliases succ)
  lid_t* P;          // [2CN] tour position of arc a
  int32_t* W;        // [2CN] arc weight, later exclusive weight prefix in tour order
  uint32_t* spX;     // [2][spCap] sublist ranking: (next << 16 | dist), ping-pong
  int32_t* spY;      // [2][spCap] sublist ranking: weight
  lid_t* ent;        // [CN] tour position of the arc entering the node
  int spCap;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// N: nodes of the whole graph, CN/CE: caps of a merged pair, R: region columns.  (MST kernels.)
// Tiers: `core` (misc, membership bitmap + prefix) always in shared memory at `base`; `hot` -
// the spanning-tree and Euler-tour arrays, which alias each other - in shared memory behind
// the core, or - when a merged pair does not fit the 227 KB of an SM (spill mode) - in a
// per-CTA global scratch block at `big` (L2 resident); `cold` - what is only streamed
// through, in edge or node order (populations, the pair's edge list, region classes) - in the
// CTA's global block at `cold`, read through L2: every KB of shared memory not spent on
// them is a step closer to one more resident CTA, and the step is latency bound.
// Returns core + hot; *split_out = core, *cold_out = the cold bytes.
__host__ __device__ inline size_t smem_layout(Smem* s, unsigned char* base, int N, int CN, int CE, int R,
                                              unsigned char* big, unsigned char* cold,
                                              size_t* split_out = nullptr, size_t* cold_out = nullptr) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align16(o + bytes); return r; };
  const int NW = (N + 31) / 32;
  size_t o_misc = take(kMiscInts * 4);
  size_t o_mbits = take(size_t(NW) * 4);
  size_t o_mpref = take(size_t(NW) * 2);
  const size_t split = o;
  if (split_out) *split_out = split;
  size_t o_tfl = take(size_t((CE + 31) / 32) * 4);
  size_t o_oldb = take(size_t((CN + 31) / 32) * 4);
  size_t o_newb = take(size_t((CN + 31) / 32) * 4);
  size_t phase = o;
  // spanning-tree phase
  size_t o_key = take(size_t(CE) * 4);
  size_t o_lab = take(size_t(CE) * 4);
  size_t o_lj = take(size_t(CE) * 2);
  size_t o_best = take(size_t(CN) * 4);
  size_t o_bpos = take(size_t(CN) * 4);
  size_t o_hook = take(size_t(CN) * 2);
  size_t o_vlA = take(size_t(CN / 2 + 2) * 2);
  size_t o_vlB = take(size_t(CN / 2 + 2) * 2);
  size_t end_a = o;
  // Euler phase
  o = phase;
  const int spCap = (2 * CN) / 8 + 4;
  size_t o_te = take(size_t(CN) * 2);
  size_t o_tuv = take(size_t(CN) * 4);
  size_t o_head = take(size_t(CN) * 4);
  size_t o_nxt = take(size_t(2 * CN) * 2);
  size_t o_succ = take(size_t(2 * CN) * 2);
  size_t o_P = take(size_t(2 * CN) * 2);
  size_t o_W = take(size_t(2 * CN) * 4);
  size_t o_spX = take(size_t(2 * spCap) * 4);
  size_t o_spY = take(size_t(2 * spCap) * 4);  // ent (2 CN bytes, written when the ranking is done) aliases it
  size_t end_b = o;
  o = 0;
  size_t c_pop = take(size_t(CN) * 4);
  size_t c_euv = take(size_t(CE) * 4);
  size_t c_ecls = take(R > 0 ? size_t(CE) : 0);
  if (cold_out) *cold_out = o;
  if (s) {
    *s = Smem{};
    s->misc = (int32_t*)(base + o_misc);
    s->mbits = (uint32_t*)(base + o_mbits);
    s->mpref = (lid_t*)(base + o_mpref);
    if (big) base = big - split;  // the offsets below are relative to the start of the layout
    s->tflag = (uint32_t*)(base + o_tfl);
    s->oldb = (uint32_t*)(base + o_oldb);
    s->newb = (uint32_t*)(base + o_newb);
    s->key = (uint32_t*)(base + o_key);
    s->lab = (uint32_t*)(base + o_lab);
    s->lj = (lid_t*)(base + o_lj);
    s->best = (uint32_t*)(base + o_best);
    s->bpos = (uint32_t*)(base + o_bpos);
    s->hook = (lid_t*)(base + o_hook);
    s->vlA = (lid_t*)(base + o_vlA);
    s->vlB = (lid_t*)(base + o_vlB);
    s->te = (lid_t*)(base + o_te);
    s->tuv = (uint32_t*)(base + o_tuv);
    s->head = (uint32_t*)(base + o_head);
    s->nxt = (lid_t*)(base + o_nxt);
    s->succ = (lid_t*)(base + o_succ);
    s->cur = (int32_t*)(base + o_succ);
    s->P = (lid_t*)(base + o_P);
    s->W = (int32_t*)(base + o_W);
    s->spX = (uint32_t*)(base + o_spX);
    s->spY = (int32_t*)(base + o_spY);
    s->ent = (lid_t*)(base + o_spY);
    s->spCap = spCap;
    s->pop = (int32_t*)(cold + c_pop);
    s->euv = (uint32_t*)(cold + c_euv);
    s->ecls = (uint8_t*)(cold + c_ecls);
  }
  return end_a > end_b ? end_a : end_b;
}

// Wilson kernels.  Tiers: `core` (misc, per-walker bookkeeping and draw rings, membership
// bitmap + prefix) always in shared memory; `S` - what a walk WRITES at random: one state
// word per node and walker, and the subtree sums / side bitmaps of the evaluation; `A` - what
// it READS at random: the pair's adjacency; `cold` - arrays that are only streamed through
// (populations, the edge list) in the CTA's global block, read through L2.
//   level 0: core + S + A in shared memory;  level 1 (a pair whose adjacency does not fit):
// core + S in shared memory, A in the CTA's global block (L2);  level 2: S and A global.
// Returns the bytes of shared memory for `level`; *glob_out the bytes of the global block
// (S and/or A), *cold_out the cold bytes.
__host__ __device__ inline size_t wilson_layout(Smem* s, unsigned char* base, int N, int CN, int CE, int R, int W,
                                                int nbr_shift, int level, int msh, unsigned char* glob, unsigned char* cold,
                                                size_t* glob_out = nullptr, size_t* cold_out = nullptr, int nbr_bytes = 4) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align16(o + bytes); return r; };
  const int NW = (N + 31) / 32;
  size_t o_misc = take(kMiscInts * 4);
  size_t o_meta = take(4 * 32 * 4);
  size_t o_ring = take(size_t(W) * 128 * 4);
  size_t o_mbits = take(size_t(NW) * 4);
  size_t o_mpref = take(size_t((NW >> msh) + 1) * 2);
  const size_t core = o;
  size_t o_oldb = take(size_t((CN + 31) / 32) * 4);
  size_t o_newb = take(size_t((CN + 31) / 32) * 4);
  size_t o_whas = take(size_t((CN + 31) / 32) * 4);
  size_t o_wsum = take(size_t(CN) * 4);
  size_t o_wst = take(size_t(W) * size_t(CN) * 2);
  const size_t s_end = o;
  size_t o_nbr = take((size_t(CN + CN / 8) << nbr_shift) * (size_t)nbr_bytes);
  size_t o_wdeg = take(size_t(CN) * 2);
  const size_t a_end = o;
  const size_t smem = level == 0 ? a_end : level == 1 ? s_end : core;
  if (glob_out) *glob_out = a_end - smem;
  o = 0;
  size_t c_pop = take(size_t(CN) * 4);
  size_t c_euv = take(size_t(CE) * 4);
  size_t c_ecls = take(R > 0 ? size_t(CE) : 0);
  if (cold_out) *cold_out = o;
  if (s) {
    *s = Smem{};
    s->misc = (int32_t*)(base + o_misc);
    s->wmeta = (int32_t*)(base + o_meta);
    s->wring = (uint32_t*)(base + o_ring);
    s->mbits = (uint32_t*)(base + o_mbits);
    s->mpref = (lid_t*)(base + o_mpref);
    s->msh = msh;
    // the global block starts where shared memory ends: offsets beyond `smem` move there
    unsigned char* hs = level == 2 ? glob - smem : base;
    unsigned char* ha = level == 0 ? base : glob - smem;
    s->oldb = (uint32_t*)(hs + o_oldb);
    s->newb = (uint32_t*)(hs + o_newb);
    s->whas = (uint32_t*)(hs + o_whas);
    s->wsum = (int32_t*)(hs + o_wsum);
    s->wtmp = (uint32_t*)(hs + o_wsum);
    s->wst = (uint16_t*)(hs + o_wst);
    s->nbr = (uint32_t*)(ha + o_nbr);
    s->wdeg = (uint16_t*)(ha + o_wdeg);
    s->pop = (int32_t*)(cold + c_pop);
    s->euv = (uint32_t*)(cold + c_euv);
    s->ecls = (uint8_t*)(cold + c_ecls);
  }
  return smem;
}

// misc[] slots
enum { M_WSUM = 0 /* 0..63: two alternating banks of 32 warp sums */,
       M_ERR = 65, M_PICK, M_CNT, M_TOT, M_NINT, M_SUBSET, M_DELTA, M_NLIVE, M_NBAL, M_TAILLEN, M_NV, M_WALKER /* Wilson: the warp whose lanes walk */, M_SPARE0, M_SPARE1 /* 78..79 */,
       M_BEST64 = 80 /* 80..81, 8-byte aligned */, M_TICKET = 82 /* 82..83 */,
       M_ROOTKIDS = 86 /* Wilson: children of the tree root */, M_CTR = 88 /* 88..95 counters */,
       /* 96..127: phase timers of RC_TIMERS builds */
       M_HC = 128 /* 128..134: the pair a helper gathered last (chain, session, a, b, n, m, tot); 128..136: arguments of long_finish */,
       M_TICKET64 = 140 /* 140..141: the ticket this CTA holds (8-byte aligned) */,
       M_KEEP = 138 /* the chain this CTA goes on with (thread 0's), -1: none */,
       M_HELPMODE = 139 /* thread 0's: this CTA, though chains are ready, draws trees for a long step */ };
// wmeta[] rows (Wilson, per walker)
enum { WM_ROOT = 0, WM_DRAWS = 32, WM_ERR = 64 };

#ifdef __CUDACC__

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exclusive block scan of one int per thread; *total = block sum.  ONE barrier: the warp
// totals go to one of two alternating banks, so a slow reader of call k is never
// overtaken by the writers of call k+2 (they are separated by the barrier of call k+1).
// `parity` is a per-thread register that every thread toggles in lock step.
__device__ __forceinline__ int block_excl_scan(int v, int* total, int32_t* misc, int& parity) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  int32_t* bank = misc + M_WSUM + 32 * parity;
  parity ^= 1;
  const int inc = warp_incl_scan(v);
  if (lane == 31) bank[w] = inc;
  __syncthreads();
  const int ws = lane < nw ? bank[lane] : 0;
  const int wi = warp_incl_scan(ws);
  const int wbase = __shfl_sync(0xffffffffu, wi - ws, w);
  *total = __shfl_sync(0xffffffffu, wi, 31);
  return wbase + inc - v;
}

// odd chunk length for thread-contiguous loops: stride-odd shared-memory accesses are
// bank-conflict free
__device__ __forceinline__ int odd_chunk(int count) {
  return ((count + (int)blockDim.x - 1) / (int)blockDim.x) | 1;
}

__device__ __forceinline__ void atomic_max_u64_smem(unsigned long long* a, unsigned long long v) {
  unsigned long long old = *a;
  while (v > old) {
    unsigned long long seen = atomicCAS(a, old, v);
    if (seen == old) break;
    old = seen;
  }
}

#endif  // __CUDACC__

}  // namespace rc
