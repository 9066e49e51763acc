/* sigma_backend.h - the launch layer between the pass driver (sigma_pass.cpp) and the device.
 *
 * libsigma_b200.so links sigma_pass.cpp against sigma_kernels.cu, which implements every
 * function below as CUDA kernels for sm_100a.  There is no other implementation in the product.
 * (tests/emul/ links the same driver against a serial host stand-in so that the host logic and
 * the per-item bodies of sigma_items.cuh can be exercised in a container without a GPU; that
 * build is test infrastructure, is never installed and cannot be loaded by the library.)
 */
#ifndef SIGMA_BACKEND_H
#define SIGMA_BACKEND_H

#include <stddef.h>
#include <stdint.h>
#include "sigma_types.h"

/* one work item per thread, bodies in sigma_items.cuh */
enum SgItemKind {
    SGK_IDSORT = 0,     /* n = nlit      : order fix of every occurrence list                     */
    SGK_REDUCE,         /* n = nlit      : reduceOT                                                */
    SGK_ELECT_PREP,     /* n = nvars     : static election classes                                 */
    SGK_SORTOT,         /* n = 2*elected : sortOT, lists of up to 24 entries (one warp each)       */
    SGK_SCRATCH_LEN,    /* n = elected   : scratch words per elected variable                      */
    SGK_SUB,            /* n = elected                                                             */
    SGK_VE_COUNT,       /* n = elected                                                             */
    SGK_VE_EMIT,        /* n = elected                                                             */
    SGK_NBH,            /* n = elected   : sortOT + SUB + BVE decision of one variable by one warp */
    SGK_BCE_COUNT,      /* n = elected                                                             */
    SGK_BCE_EMIT,       /* n = elected                                                             */
    SGK_ERE_CHUNKS,     /* n = elected   : 256-pair work items per variable                        */
    SGK_ERE_PROPOSE,    /* n = work items: one warp each                                           */
    SGK_ERE_DIRTY,      /* n = elected                                                             */
    SGK_ERE_LIST,       /* n = elected   : positions with recorded pairs or dirty -> v.acting (ordered)    */
    SGK_ERE_CLAIM,      /* n = recorded pairs before the first dirty variable: one warp each       */
    SGK_ERE_SAFE,
    SGK_ERE_APPLY,
    SGK_ERE_FLAG01,     /* n = elected   : ve_ncls[e] = (ere_flags[e] != 0)                               */
    SGK_SORTOT_LONG,    /* n = 2*elected : sortOT, lists of more than 24 entries (one thread each)  */
    SGK_IDSORT_LONG,    /* n = literals  : createOT order fix for the lists of more than 64 entries */
    SGK_N
};
/* sequential stages, one thread */
enum SgSerialKind { SGS_PROP = 0, SGS_MARKS, SGS_APPLY_UNITS, SGS_ERE_REPLAY /* one warp */, SGS_ERE_SPLIT, SGS_N };

#ifdef __cplusplus
extern "C++" {
#endif

/* ---- device lifetime / memory ---------------------------------------------------------------- */
int   be_init(int device, void **stream_out, char *err, size_t errlen);   /* 0 = ok */
void  be_shutdown(void *stream);
void  be_bind(int device);                     /* make the context's device current for the calling host thread */
void *be_malloc(size_t bytes);                 /* NULL on failure */
void  be_free(void *p);
void *be_host_alloc(size_t bytes);             /* pinned host memory */
void  be_host_free(void *p);
int   be_host_register(void *p, size_t bytes);   /* pin caller-owned pageable memory; 0 = ok */
int   be_host_unregister(void *p);
int   be_h2d(void *stream, void *dst, const void *src, size_t bytes);
int   be_d2h(void *stream, void *dst, const void *src, size_t bytes);
int   be_d2d(void *stream, void *dst, const void *src, size_t bytes);
int   be_memset(void *stream, void *dst, int byte, size_t bytes);
int   be_sync(void *stream, char *err, size_t errlen);                      /* 0 = ok */
void *be_event_create(void);
void  be_event_destroy(void *ev);
void  be_event_record(void *stream, void *ev);
float be_event_ms(void *a, void *b);
uint64_t be_launch_count(void *stream);        /* kernels launched so far on this context's stream */
/* small control blocks between device memory and mapped pinned host memory, without the copy engine (bytes % 4 == 0) */
void  be_copy_words(void *stream, void *dst, const void *src, size_t bytes);

/* ---- streaming kernels ------------------------------------------------------------------------ */
/* AoS SCNF arena -> device store (reference hand-off layout, sclause.hpp:45-49) */
/* *bad != 0 afterwards: a ref or a size word points outside the n_buckets words of the arena (sizes[] are then 0 there) */
void be_ingest_sizes(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, uint32_t *sizes, uint32_t *bad);
void be_ingest_copy(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, const uint32_t *offs,
                    sg_u4 *hdr, uint32_t *lits);
/* one-kernel ingest for a gap-free arena (what awaken() leaves); *gap != 0 afterwards: use the general path */
void be_ingest_fused(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, sg_u4 *hdr, uint32_t *lits, uint32_t *gap);
/* extract() on the device (simplify.cpp:118-133): sizes[i] = SCNF buckets of source clause i (0 = deleted), live[i] = 0/1;
 * then, with the two exclusive scans, the SCLAUSE records and their refs */
void be_extract_sizes(void *s, const uint32_t *cm, uint64_t cm_buckets, const uint64_t *crefs, uint32_t n, const uint8_t *stencil, uint64_t stencil_size,
                      uint32_t *sizes, uint32_t *live, uint32_t *bad);   /* *bad != 0: a ref / record leaves the cm_buckets words */
void be_extract_write(void *s, const uint32_t *cm, const uint64_t *crefs, uint32_t n, const uint32_t *sizes, const uint32_t *newoff,
                      const uint32_t *newid, uint32_t *arena, uint64_t *refs);
/* newBeginning() on the device (simplify.cpp:249-268, sclause.cpp:22-62): exported SCNF -> CLAUSE records + orgs / learnts.
 * learnt[i] = 0/1 first; then, with its exclusive scan, the records (clause i at bucket refs[i] + i), the two ref lists,
 * counts[0] / counts[1] = literals of original / learnt clauses and, when subsume != NULL, the mark_ssubsume bytes */
void be_cm_flags(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t *learnt);
void be_cm_write(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, const uint32_t *learnt_before,
                 int lbd_tier1, uint32_t *cm, uint64_t *orgs, uint64_t *learnts, unsigned long long *counts, uint8_t *subsume, const uint32_t *vmap /* device, may be NULL: variable mapping (map.hpp:52-76) */);
/* ERE proposal pass: filt[2c] = size of clause c (0 when deleted), filt[2c+1] = its signature */
void be_ere_filt(void *s, const SgView &v, uint32_t *filt, uint8_t *size8);   /* size8[c] = min(size, 255), 0 when deleted */
void be_ere_maxsz(void *s, const SgView &v, const uint8_t *size8, uint8_t *maxsz);   /* maxsz[lit] = largest size8 in the list of lit */
/* createOT: every list of the freshly built (packed) table into ascending clause id order (simplify.cpp:50-58) */
void be_ot_order(void *s, const SgView &v);
/* `count` (<= 4) independent exclusive prefix sums of the same length with the launches of one; tmp: count x be_scan_tmp_words(n) */
struct SgScanSet { const uint32_t *in[4]; uint32_t *out[4]; uint32_t *total[4]; };
void be_scan_multi(void *s, const SgScanSet &ss, uint32_t count, uint32_t n, uint32_t *tmp);
/* exclusive prefix sum; out may alias in; *total (device) receives the sum when non-NULL */
void be_scan_u32(void *s, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total, uint32_t *tmp);
size_t be_scan_tmp_words(uint32_t n);
/* per-literal occurrence histogram over live clauses (histogram.hpp:43-50, simplify.cpp:50-58) */
void be_hist(void *s, const SgView &v, uint32_t *counts);
/* createOT (simplify.cpp:39-61) as a two-level partition (sigma_kernels.cu): ot_off / ot_len / ot_data from the store, every
 * list in ascending clause id order (= the reference's push order).
 *   be_ot_build_plan    geometry + layout of tmp (be_ot_build_tmp_words words); `lits` = live literals (an upper bound is fine);
 *                       non-zero / 0 words: the formula does not fit (caller: be_hist / scan / be_scatter / be_ot_order)
 *   be_ot_build_count   per (bucket, tile) literal counts into plan.cnt; bits (may be NULL): bitmap of the live clauses
 *   (caller)            exclusive scan of plan.cnt[0 .. plan.ncnt) in place, total -> device word
 *   be_ot_build_scatter (literal, clause) pairs, bucket by bucket, tile by tile
 *   be_ot_build_lists   one block per bucket: ot_len, ot_off, ordered lists */
struct SgOtBuild { uint32_t shift, NB, T, per_tile; size_t ncnt; uint32_t *cnt, *scantmp; void *pairs; };
size_t be_ot_build_tmp_words(uint32_t nlit, uint32_t n_cls, uint64_t lits);
int be_ot_build_plan(uint32_t nlit, uint32_t n_cls, uint64_t lits, uint32_t *tmp, SgOtBuild *plan);
void be_ot_build_count(void *s, const SgView &v, const SgOtBuild &plan, uint32_t *bits);
void be_ot_build_scatter(void *s, const SgView &v, const SgOtBuild &plan);
void be_ot_build_lists(void *s, const SgView &v, const SgOtBuild &plan, const uint32_t *total);
/* scatter clause ids into the lists; cursor must hold a copy of ot_off (absolute write positions) */
void be_scatter(void *s, const SgView &v, uint32_t *cursor);
/* Incremental createOT for phases after the first.  The occurrence table the reference rebuilds from
 * scratch (simplify.cpp:39-61) differs from the table at the end of the previous phase only in the lists that
 * hold a clause deleted since (entries dropped), the lists of the last elected variables (sortOT order, or
 * emptied when eliminated), and the lists that gain entries: literals of the clauses with id >= first_new
 * (resolvents) and the (literal, clause) pairs in v.ot_add (equivalence substitution).  Only those lists are
 * rewritten, in place; a list that gains entries moves to fresh room at the end of ot_data (lists are
 * (offset, length) pairs and need not be packed).  Clause ids must be unchanged since the table was built.
 *   be_ot_bitmap      : after a full build: bits = clauses represented in the table
 *   be_ot_update_plan : marks the lists to rewrite; *total receives the words of fresh room needed
 *   be_ot_update      : rewrites them (ot_data must have *bump + needed words of capacity) */
struct SgOtUpdate {
    uint32_t *bits;             /* clauses represented in the table (updated) */
    uint32_t first_new;         /* first clause id that has no entry in the table */
    uint32_t n_add;             /* pairs in v.ot_add */
    uint32_t n_elected;         /* entries of v.elected from the previous phase */
    uint32_t *tmp, *total, *bump;  /* total: words the growing lists need (NULL = not wanted); bump: end of the table, advanced */
};
size_t be_ot_update_tmp_words(uint32_t nlit, uint32_t n_cls);
size_t be_ot_bits_words(uint32_t n_cls);
void be_ot_bitmap(void *s, const SgView &v, uint32_t *bits);
void be_ot_update_plan(void *s, const SgView &v, const SgOtUpdate &u);
void be_ot_update(void *s, const SgView &v, const SgOtUpdate &u);
/* compare two occurrence tables entry by entry (test hook); *mismatch receives the number of differences */
void be_ot_compare(void *s, const SgView &v, const uint32_t *off2, const uint32_t *len2, const uint32_t *data2, uint32_t *mismatch);

/* scores[v] = ps*ns (uint32 wrap, histogram.hpp:23), order[v-1] = v */
void be_scores(void *s, const SgView &v);
/* stable ascending sort of (keys, vals); result in keys/vals; tmp arrays of n each */
void be_sort_pairs(void *s, uint32_t *keys, uint32_t *vals, uint32_t n, uint32_t *tmpk, uint32_t *tmpv, uint32_t *tmp,
                   uint32_t key_bits);
size_t be_sort_tmp_words(uint32_t n);
void be_rank(void *s, const SgView &v);                                           /* rank[order[r]] = r */
/* election rounds */
/* worklist of the still-undecided candidates with rank in [r0, r1): fills wl, sc->wl_count[0] (must be 0) */
void be_elect_batch(void *s, const SgView &v, uint32_t r0, uint32_t r1, uint32_t *wl);
void be_elect_round(void *s, const SgView &v, const uint32_t *wl_in, uint32_t *wl_out, int parity);
/* every batch and round of one election in a single launch; returns non-zero when unavailable (the caller then
 * drives be_elect_batch / be_elect_round itself).  sc->wl_count[0] receives the number of rounds. */
int be_elect_all(void *s, const SgView &v, uint32_t *wl0, uint32_t *wl1, uint32_t b0, uint32_t grow);
/* the whole election in one launch without rounds: every warp takes its candidates in rank order, parks the blocked ones
 * and polls their blockers (sigma_kernels.cu, k_elect_async).  Same return convention; sc->wl_count[0] = looks of the busiest warp. */
int be_elect_async(void *s, const SgView &v);
/* scores + stable sort by (score, variable) + rank[] in one cooperative launch; forced_bits = 0: key width from the largest
 * score, on the device.  Nonzero return: not available, use be_scores / be_sort_pairs / be_rank. */
int be_vo_sort(void *s, const SgView &v, uint32_t *tmpk, uint32_t *tmpv, uint32_t *tmp, uint32_t forced_bits);
void be_acting_flags(void *s, const SgView &v, uint32_t *flags_acting, uint32_t *flags_elected);   /* per rank position */
void be_acting_scatter(void *s, const SgView &v, const uint32_t *fa, const uint32_t *pa, const uint32_t *fe, const uint32_t *pe);
/* generic item / serial launches */
void be_items(void *s, int kind, const SgView &v, uint32_t n);
void be_serial(void *s, int kind, const SgView &v, uint32_t arg, uint32_t arg2 = 0);
void be_sort_units(void *s, const SgView &v, uint32_t n);
/* prop by frontiers (propelim.cpp:46-81; method in sigma_items.cuh).  be_has_prop_frontier() == 0: only be_serial(SGS_PROP).
 * mark: fpos[] of the frontier's falsified literals, counters cleared.  round: every visit with the guess pp.disc, reports into
 * pp.disc_new; diff != 0: ctl[2] = the report differs from the guess.  next: the report becomes the guess (the caller swaps the two
 * pointer pairs afterwards).  count: cnt[] per position, ctl[3] = a visit ends in a conflict, base[] = prefix sum, ctl[4] =
 * new units.  commit: units on the trail in visit order, clauses rewritten (claim[] zeroed by the caller), lists of the
 * frontier emptied, values / states / scalars set, all side tables back to their idle state.  conflict: idle state + unsat. */
int  be_has_prop_frontier(void);
/* the n units in v.ubuf (sorted) applied at once, same result as be_serial(SGS_APPLY_UNITS): first = one word per literal, all
 * ones on entry and on return; flags / pos: n words each; total: one word */
void be_apply_units(void *s, const SgView &v, uint32_t n, uint32_t *first, uint32_t *flags, uint32_t *pos, uint32_t *total, uint32_t *scan_tmp);
void be_pp_mark(void *s, const SgView &v, const SgProp &pp);
void be_pp_round(void *s, const SgView &v, const SgProp &pp, int diff);
void be_pp_next(void *s, const SgProp &pp);
void be_pp_count(void *s, const SgView &v, const SgProp &pp, uint32_t *scan_tmp);
void be_pp_commit(void *s, const SgView &v, const SgProp &pp, uint32_t n_new);
void be_pp_conflict(void *s, const SgView &v, const SgProp &pp);
/* sortOT of the lists with 25 .. SG_SOT_WIDE_MAX entries by one warp each (work list left by SGK_SORTOT_LONG when v.sot_wl is set) */
int be_has_sortot_wide(void);
int be_sortot_wide(void *s, const SgView &v);
/* the ERE replay by one block of 32 warps: parallel claim / apply rounds for the recorded pairs between consecutive dirty
 * variables (v.ere_owner all ones, v.ere_done zero on entry; claims: 8 words per recorded pair, ccnt: one).  Nonzero return:
 * not available (be_serial(SGS_ERE_REPLAY) instead) */
int be_ere_replay_block(void *s, const SgView &v, uint32_t n_events, uint32_t n_list, uint32_t *claims, uint32_t *ccnt);                          /* by (eidx, seq) */
/* live clause / literal count (solve.hpp:661-671) and melted variable count (solve.hpp:645-654) */
void be_count_live(void *s, const SgView &v);
void be_count_melted(void *s, const SgView &v, uint32_t *out);
/* compaction (simplify.cpp:235-247): scan of (live, size) over the clause ids straight from the headers
 * (totals[0] = live clauses, totals[1] = live literals), then the move fused with the down-sweep */
size_t be_live_tmp_words(uint32_t n_cls);
void be_live_scan(void *s, const SgView &v, uint32_t *tmp, uint32_t *totals);
void be_compact_apply(void *s, const SgView &v, const uint32_t *tmp, sg_u4 *hdr2, uint32_t *lits2);
/* device store -> AoS SCNF arena + refs in the reference's layout */
void be_export_apply(void *s, const SgView &v, const uint32_t *tmp, uint32_t *arena, uint64_t *refs);

void be_iota_u32(void *s, uint32_t *p, uint32_t n);       /* p[i] = i */
/* ---- DIMACS parser (dimacs.cpp:25-177; method and per-item bodies in sigma_dimacs.cuh) ------------------------------ */
/* start state of every 64-byte chunk of text[0, n) (the clause part of the file, after the header) + the state at the end */
void be_dm_chunk_states(void *s, const uint8_t *text, uint64_t n, uint32_t nchunks, uint32_t *fn, uint32_t *tilefn, uint32_t *tstate, uint32_t *cstate,
                        uint32_t *final_state);
/* emit = 0: integers / terminators that start in each chunk (ntok, nterm); *bad = 2 on a byte the reference rejects.
 * emit = 1: tok[] (0 = terminator, else 2 var + sign) and term[] (token index of every terminator); *bad = 3: variable > nvars */
void be_dm_tokens(void *s, const uint8_t *text, uint64_t n, uint32_t nchunks, const uint32_t *cstate, uint32_t nvars, int emit, uint32_t *ntok,
                  uint32_t *nterm, const uint32_t *tokoff, const uint32_t *termoff, uint32_t *tok, uint32_t *term, uint32_t *bad);
/* one round of makeClause for all clauses: csize, units into utime_new (atomic min of the clause index), *unsat = first empty clause */
void be_dm_clauses(void *s, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, uint32_t *utime_new, uint32_t *csize, uint32_t *unsat);
void be_dm_diff(void *s, const uint32_t *a, const uint32_t *b, uint32_t n, uint32_t *changed);
void be_dm_sizes(void *s, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, const uint32_t *csize, uint32_t *words,
                 uint32_t *kept, uint32_t *isunit);
void be_dm_write(void *s, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, const uint32_t *csize, const uint32_t *wordoff,
                 const uint32_t *keptoff, const uint32_t *unitoff, const uint32_t *isunit, uint32_t *cm, unsigned long long *orgs, uint32_t *trail,
                 int8_t *value, uint8_t *state, unsigned long long *stats, uint32_t header_clauses, uint32_t *bad /* 4: too many clauses */);

/* ---- rebuildWT (watch.cpp:135-157) on the device --------------------------------------------------------------------- */
/* group of every clause of the exported SCNF in the attach order (0/1 flags per group) */
void be_wt_groups(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t code, uint32_t *f0, uint32_t *f1, uint32_t *f2, uint32_t *f3);
/* sortClause on the CLAUSE records + the scratch store of watched-literal pairs (one two-literal entry per clause, in
 * attach order) + {ref, size, first ^ second} per entry; totals = the four group sizes (device) */
void be_wt_fill(void *s, uint32_t *cm, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t code, const uint32_t *p0, const uint32_t *p1,
                const uint32_t *p2, const uint32_t *p3, const uint32_t *totals, const int8_t *value, sg_u4 *hdr, uint32_t *lits, sg_u4 *winfo);
/* WATCH records of every list from the occurrence table of the scratch store */
void be_wt_records(void *s, uint32_t nlit, const uint32_t *ot_off, const uint32_t *ot_len, const uint32_t *ot_data, const sg_u4 *winfo, sg_u4 *rec);

#ifdef __cplusplus
}
#endif
#endif
