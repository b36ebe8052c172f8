/*
 * recom_oracle.h - CPU oracle for the ReCom hot path.  TEST INFRASTRUCTURE ONLY:
 * see the header of recom_oracle.c.  Shares the plain-C structs of the product ABI
 * (include/recom_b200.h) so one replay record / configuration drives both sides; in the
 * oracle every RcState / RcReplay pointer is a HOST pointer.
 */
#ifndef RECOM_ORACLE_H
#define RECOM_ORACLE_H

#include "../include/recom_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RcoWork RcoWork;

RcoWork* rco_work_create(const RcGraph* g);
void rco_work_destroy(RcoWork* w);

int rco_init_state(const RcGraph* g, const RcConfig* c, const rc_label_t* assign,
                   int64_t* tally, uint32_t* mask, int32_t* cut_count);
int rco_step(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t global_chain,
             rc_label_t* assign, int64_t* tally, uint32_t* mask, int32_t* cut_count,
             uint32_t* proposal_no, int64_t* counters, uint32_t* flags, int32_t* last_pair);
int rco_replay(const RcGraph* g, const RcConfig* c, RcoWork* w, const RcReplay* rp,
               rc_label_t* assign, int64_t* tally, uint32_t* mask, int32_t* cut_count,
               int64_t* counters, uint32_t* flags);
/* TEST HOOK: the recorded randomness of one reference reversible_recom call
 * (proposals/tree_proposals.py:149-267), replayed by identity (oracle/record_reversible.py). */
typedef struct RcoRevStep {
  int32_t d_out, d_in;   /* random.choice(dist_pairs), tree_proposals.py:224 (district indices)   */
  int32_t pair_edge;     /* random.choice(list(pair_edges)), :230; -1: self-loop before any tree  */
  int32_t root;          /* node index chosen at tree.py:377; -1: no tree                         */
  int32_t cut_edge;      /* edge id of choice(all_cuts).edge, :244; -1: no balanced cut            */
  int32_t pad;
  int64_t tree_idx;      /* index of this call's tree in the Wilson arrays of the RcReplay         */
  double accept_u;       /* random.random() at :262 (NaN when the call never got there)            */
} RcoRevStep;
enum { RCO_REV_NO_PAIR = 0,   /* :227 self-loop: equal districts or not adjacent                  */
       RCO_REV_NO_CUT = 1,    /* :241 self-loop: the tree has no balanced edge                    */
       RCO_REV_REJECTED = 2,  /* :265 self-loop: the acceptance draw failed                       */
       RCO_REV_MOVED = 3 };   /* :263 the new partition                                           */
int rco_replay_reversible(const RcGraph* g, const RcConfig* c, RcoWork* w, int32_t n_steps,
                          const RcoRevStep* steps, const RcReplay* rw, rc_label_t* assign, int64_t* tally,
                          uint32_t* mask, int32_t* cut_count, int64_t* counters,
                          rc_label_t* assign_trace, int32_t* info_trace);
int rco_seed_plan(const RcGraph* g, const RcConfig* c, RcoWork* w, uint32_t chain,
                  int32_t seed_max_attempts, rc_label_t* assign, int64_t* tally,
                  int64_t* counters, const RcReplay* rp);
int rco_balanced_cuts_of_tree(int32_t n, const int32_t* tree_uv, const int32_t* pops,
                              double ideal, double eps, int32_t root, int32_t* out_uv);
int rco_mst_of_graph(const RcGraph* g, const double* weights, int32_t* out_eids);
int rco_init_chains(const RcGraph* g, const RcConfig* c, const RcState* st);
int rco_run_chains(const RcGraph* g, const RcConfig* c, int64_t n_steps, const RcState* st,
                   int n_threads);
int rco_region_splits(const RcGraph* g, int32_t k, const rc_label_t* assign, const int32_t* region,
                      int32_t n_regions, int32_t* out);
int rco_contiguity(const RcGraph* g, int32_t k, const rc_label_t* assign, int32_t* out);
void rco_philox(uint32_t k0, uint32_t k1, const uint32_t* ctr, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
