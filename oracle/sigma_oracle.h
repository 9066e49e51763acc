/* sigma_oracle.h - CPU ORACLE of the SIGmA pass.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the reference algorithm (muhos/SeqFROST v1.0.2,
 * simplify.cpp:180-210 and the stage files it calls), written over the same structure-of-arrays
 * clause store the CUDA path uses so that every intermediate array can be compared 1:1.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libsigma_b200.so) never does and has no CPU fallback.
 *
 * Parity pinning: the reference ships no golden vectors.  This port is pinned
 * against the UNMODIFIED reference compiled by oracle/Makefile (`make ref`, oracle/_ref/ref_sigma)
 * on seeded instances - tests/test_oracle.py - and against the fixtures generated
 * from that binary and committed under tests/golden/ (tests/golden/make_golden.py).
 *
 * The entry points mirror include/sigma_b200.h one for one (sgo_* instead of sigma_*).
 */
#ifndef SIGMA_ORACLE_H
#define SIGMA_ORACLE_H
#include "../include/sigma_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sgo_ctx sgo_ctx;
int  sgo_create(sgo_ctx **out);
void sgo_destroy(sgo_ctx *ctx);
int  sgo_upload(sgo_ctx *ctx, const sigma_opts *o, const sigma_formula *in);
int  sgo_execute(sgo_ctx *ctx);
int  sgo_result_sizes(sgo_ctx *ctx, sigma_formula *sizes);
int  sgo_download(sgo_ctx *ctx, sigma_formula *out);
int  sgo_get_report(sgo_ctx *ctx, sigma_report *rep);
int  sgo_set_stop(sgo_ctx *ctx, int phase, int stage);
int  sgo_snapshot(sgo_ctx *ctx, const char *name, void *dst, uint64_t cap_bytes, uint64_t *nbytes);
/* the hand-off stages either side of the pass, stand-alone:
 * extract(orgs); extract(learnts) of awaken() (simplify.cpp:118-133,146-148) into caller-sized arena / refs
 * (upper bounds: cnf->cm_buckets buckets, n_orgs + n_learnts refs); returns the clause count */
uint64_t sgo_extract(const sigma_cnf *cnf, uint32_t *arena, uint64_t *refs, uint64_t *n_buckets);
/* newBeginning() without variable mapping (simplify.cpp:249-268, sclause.cpp:22-62) from a compacted SCNF */
void sgo_new_beginning(const uint32_t *arena, const uint64_t *refs, uint64_t n_clauses, int lbd_tier1, int later_call,
                       sigma_cnf_out *out);
/* Solver::parser() + makeClause() (dimacs.cpp:25-177) over a file image; output buffers and counters as sigma_parse_download */
int  sgo_parse_dimacs(const char *text, uint64_t bytes, sigma_dimacs *out);
#ifdef __cplusplus
}
#endif
#endif
