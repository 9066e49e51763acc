/* sigma_b200.h - C ABI of the B200-native SIGmA inprocessing pass (libsigma_b200.so).
 *
 * The reference (muhos/SeqFROST) has no plugin/FFI interface: the pass is a set of Solver
 * member functions.  The boundary this library sits behind is the narrowest cut through
 * Solver::simplifying() (reference src/simplify.cpp:155-233):
 *
 *     ... rootify(); shrinkTop(false); awaken();            simplify.cpp:161-171  (host, kept)
 *     >>> while (litsbefore) { resizeCNF .. filterElected }  simplify.cpp:180-197  (REPLACED)
 *     >>> prop();                                            simplify.cpp:204      (REPLACED)
 *     >>> clearMapFrozen(); countFinal(); shrinkSimp();      simplify.cpp:208-210  (REPLACED)
 *     ... SAT check; map()/newBeginning(); rebuildWT(); ..   simplify.cpp:218-229  (host, kept)
 *
 * The hand-off object on both sides is the reference's own SCNF (simptypes.hpp:48-95): a bump
 * arena of uint32 buckets holding SCLAUSE records (sclause.hpp:45-49: word0 bitfield
 * st:2 f:1 a:1 u:2 lbd:26, word1 sig, word2 size, then `size` ascending literals, literal =
 * 2*var+sign, constants.hpp:77-82) plus the `refs` vector (bucket index per clause, allocation
 * order), together with the solver state the pass reads and writes.
 *
 * Plain C: pointers + sizes only, no C++ types, no exceptions across the boundary.  One
 * sigma_ctx per GPU; a ctx is not thread-safe; different ctxs are independent.
 * There is NO CPU fallback: every entry point fails with SIGMA_E_CUDA when no device is usable.
 */
#ifndef SIGMA_B200_H
#define SIGMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (reference: prop()/VE()/strengthen() end in learnEmpty();killSolver(),
 *      control.cpp:221; allocators throw MEMOUTEXCEPTION, malloc.hpp:28-58) ---------------- */
enum {
    SIGMA_OK = 0,
    SIGMA_UNSAT = 1,          /* the pass derived the empty clause                            */
    SIGMA_INTERRUPTED = 2,    /* *interrupt_flag became non-zero between stages               */
    SIGMA_E_NOMEM = 3,        /* device or host allocation failed                             */
    SIGMA_E_CUDA = 4,         /* no device / launch failure (see sigma_last_error)            */
    SIGMA_E_NOSPACE = 5,      /* an output buffer is too small: required sizes are filled in  */
    SIGMA_E_ARG = 6,          /* malformed input                                              */
    SIGMA_E_UNSUPPORTED = 7   /* option combination outside the implemented path              */
};

/* ---- options: OPTION fields that parameterise the pass (options.hpp:108-130, defaults
 *      options.cpp:24-51,146; derived rules options.cpp:315-351).  32 x int32, the same order
 *      as the "opts" section written by oracle/ref_harness.cpp. ----------------------------- */
typedef struct sigma_opts {
    int32_t phases;            /* --phases (5), decremented by the caller after each call     */
    int32_t phase_lits_min;    /* --eliminationphasemin (500)                                  */
    int32_t mu_pos, mu_neg;    /* --mupos/--muneg (32/32)                                      */
    int32_t lcve_max_occurs;   /* --electionsmax (3000)                                        */
    int32_t lcve_clause_max;   /* --electionsclausemax (30000)                                 */
    int32_t lcve_min_vars;     /* --electionsmin (1)                                           */
    int32_t sub_max_occurs;    /* --subsumemaxoccurs (1000)                                    */
    int32_t sub_clause_max;    /* --subsumeclausemax (100)                                     */
    int32_t ve_clause_max;     /* --boundedclausemax (100, 0 = unlimited)                      */
    int32_t xor_max_arity;     /* --xormaxarity (10)                                           */
    int32_t bce_max_occurs;    /* --blockedmaxoccurs (1000)                                    */
    int32_t ere_max_occurs;    /* --redundancymaxoccurs (1000)                                 */
    int32_t ere_clause_max;    /* --redundancyclausemax (200)                                  */
    int32_t collect_freq;      /* --collectfrequency (2)                                       */
    int32_t lbd_tier1;         /* --lbdtier1 (2), used by bumpShrunken, solve.hpp:725          */
    int32_t sub_en, ve_en, ve_plus_en, ve_fun_en, ve_lbound_en;
    int32_t bce_en, ere_en, ere_extend, all_en, proof_en;
    int32_t reserved[6];
} sigma_opts;

/* ---- the formula + solver state at the cut (used for input and output) ------------------- */
typedef struct sigma_formula {
    uint32_t  nvars;           /* inf.maxVar                                                   */
    uint32_t  n_clauses;       /* scnf.size(): number of refs                                  */
    uint64_t  n_buckets;       /* used uint32 buckets of the arena                             */
    uint64_t  n_literals;      /* inf.nLiterals                                                */
    uint32_t *arena;           /* SCLAUSE records, sclause.hpp:45-49                           */
    uint64_t *refs;            /* S_REF per clause, simptypes.hpp:43                           */
    uint8_t  *state;           /* sp->state[v].state, nvars+1 bytes (state.hpp:26-41)          */
    int8_t   *value;           /* sp->value[lit], 2*nvars+2 bytes (-1 unassigned, 0, 1)        */
    uint32_t *trail;           /* root trail                                                   */
    uint32_t  trail_size;
    uint32_t  propagated;      /* sp->propagated                                               */
    uint32_t *vorg;            /* current -> original variable id, nvars+1 (model.hpp:65)      */
    uint8_t  *ifrozen;         /* incremental freeze mask or NULL (solve.hpp:816)              */
    uint32_t *model;           /* MODEL::resolved (model.hpp:31)                               */
    uint64_t  model_size;
    uint32_t  unassigned;      /* inf.unassigned                                               */
    uint32_t  max_frozen;      /* inf.maxFrozen                                                */
    uint32_t  max_melted;      /* inf.maxMelted                                                */
    uint32_t  simplify_calls;  /* stats.simplify.calls (1 = first call, bounded.cpp:39)        */
    /* capacities of the caller-owned OUTPUT buffers (ignored on input) */
    uint64_t  arena_cap;       /* in uint32 buckets                                            */
    uint32_t  refs_cap;
    uint32_t  trail_cap;
    uint64_t  model_cap;
} sigma_formula;

/* ---- what a run did: counters mirror stats.simplify.* (statistics.hpp:42-52) and the
 *      per-stage timers timer.{vo,cot,sot,rot,ve,sub,gc} (timer.hpp:35-51).  Every stage is a
 *      group of kernels bracketed by CUDA events on the library's stream; ms / launches / bytes
 *      are summed over the phases of the last sigma_execute. ---------------------------------- */
enum { SIGMA_T_H2D = 0,      /* sigma_upload: host -> device copies                                   */
       SIGMA_T_INGEST,       /* AoS arena -> device store                                             */
       SIGMA_T_OT_HIST,      /* createOT: per-literal histogram            (simplify.cpp:39-61)       */
       SIGMA_T_OT_SCAN,      /* createOT: exclusive scan of the histogram                              */
       SIGMA_T_OT_SCATTER,   /* createOT: clause ids into the lists                                    */
       SIGMA_T_OT_ORDER,     /* createOT: per-list order fix (push order of the reference)             */
       SIGMA_T_PROP,         /* prop()                                      (propelim.cpp:46-81)       */
       SIGMA_T_VO,           /* histCNF + scores + stable sort              (lcve.cpp:34-40)           */
       SIGMA_T_ELECT,        /* election + marks                            (lcve.cpp:52-102)          */
       SIGMA_T_SOT,          /* sortOT                                      (simplify.cpp:85-100)      */
       SIGMA_T_SUB,          /* SUB + its units                             (subsume.cpp:23-43)        */
       SIGMA_T_VE,           /* VE count + scan + emit + its units          (bounded.cpp:27-185)       */
       SIGMA_T_COUNT,        /* countAll / countFinal                       (solve.hpp:645-671)        */
       SIGMA_T_GC,           /* shrinkSimp between phases                   (simplify.cpp:235-247)     */
       SIGMA_T_EXPORT,       /* final shrinkSimp fused with the device store -> AoS arena export       */
       SIGMA_T_D2H,          /* sigma_download: device -> host copies                                  */
       SIGMA_T_TOTAL_DEVICE, /* first kernel of sigma_execute to its last                              */
       SIGMA_T_ERE,          /* ERE: propose + ordered replay              (redundancy.cpp:26-123)    */
       SIGMA_T_OT_UPDATE,    /* createOT of a later phase as an update of the previous table          */
       SIGMA_T_NBH,          /* sortOT + SUB + BVE decision fused per elected variable (one warp each) */
       SIGMA_T_N = 20 };

typedef struct sigma_report {
    int32_t  status;
    int32_t  phases_run;
    uint32_t n_clauses, max_melted_delta;
    uint64_t n_literals;
    uint64_t literals_in;           /* inf.nLiterals after extract(): the "literals/s" numerator */
    int64_t  pures, inverters, andors, ites, xors, funs, resolutions, subsumed, strengthened;
    int64_t  units, sort_ties_gt24;
    float    stage_ms[SIGMA_T_N];        /* CUDA-event milliseconds per stage                       */
    uint32_t stage_launches[SIGMA_T_N];  /* kernels launched per stage                              */
    uint64_t stage_bytes[SIGMA_T_N];     /* algorithmic bytes per stage, 0 = n/a  */
    uint64_t kernel_launches;       /* kernels of this library launched by the last execute       */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t elected_per_phase[8];
    uint32_t election_rounds[8];
    uint64_t reserved[8];
    uint32_t ot_build_mode;         /* full occurrence-table builds of this run: 2 = two-level partition (k_otb_*), 1 = one-level
                                       histogram + scatter + order fix (k_hist, k_scatter, k_ot_order_staged), 0 = none  */
    uint32_t reserved32[15];
} sigma_report;

typedef struct sigma_ctx sigma_ctx;

/* lifecycle */
int  sigma_create(int device, sigma_ctx **out);
void sigma_destroy(sigma_ctx *ctx);
void sigma_default_opts(sigma_opts *o);               /* reference defaults, ERE as options.cpp:28 */
const char *sigma_strerror(int code);
const char *sigma_last_error(sigma_ctx *ctx);

/* one-shot: replaces simplify.cpp:180-210.  `out` holds caller-owned buffers with capacities;
 * on SIGMA_E_NOSPACE the required n_buckets / n_clauses / trail_size / model_size are set and
 * the device result is kept so that sigma_download() can be retried with larger buffers. */
int  sigma_run(sigma_ctx *ctx, const sigma_opts *o, const sigma_formula *in, sigma_formula *out,
               sigma_report *rep);

/* split form (what sigma_run does internally; lets the host overlap, re-time or diff stages) */
int  sigma_upload(sigma_ctx *ctx, const sigma_opts *o, const sigma_formula *in);   /* H2D only       */
int  sigma_execute(sigma_ctx *ctx);                          /* ingest + phases + export, on device */
int  sigma_result_sizes(sigma_ctx *ctx, sigma_formula *sizes);                     /* after execute  */
int  sigma_download(sigma_ctx *ctx, sigma_formula *out);                           /* D2H only       */
int  sigma_get_report(sigma_ctx *ctx, sigma_report *rep);
/* sigma_execute always starts from the uploaded formula, so it can be repeated for timing.     */

/* ---- widening #1: extract() on the device ------------------------------------
 * The solver's clause database as it stands when awaken() is entered (simplify.cpp:135-153), i.e.
 * BEFORE the reference's extract(orgs); extract(learnts) (simplify.cpp:118-133) has built the SCNF.
 * sigma_upload_cnf copies it to the device and builds there what extract() builds on the host: one
 * SCLAUSE per clause that cm.deleted(ref) does not flag, originals first, then learnts, each in list
 * order; status / usage / lbd taken over as SCLAUSE(const CLAUSE&) does (sclause.hpp:61-75), literals
 * sorted ascending (SORT, sort.hpp:44-52), signature = calcSig (sclause.hpp:158-162).  Afterwards the
 * context is in the state sigma_upload would have left it in; `in->arena / refs / n_clauses /
 * n_buckets / n_literals` are ignored (the counts are what extract() would leave in inf.nClauses /
 * inf.nLiterals and are returned by sigma_extracted_sizes). */
typedef struct sigma_cnf {
    const uint32_t *cm;        /* CMM arena (solvetypes.hpp:57): CLAUSE records, clause.hpp:47-59 -
                                  4 header buckets {flags|_l|_v|_used, _sz, _pos, _lbd} + _sz literals */
    uint64_t  cm_buckets;      /* cm.size(): used buckets                                          */
    const uint64_t *orgs;      /* Solver::orgs (BCNF = Vec<C_REF>)                                  */
    uint64_t  n_orgs;
    const uint64_t *learnts;   /* Solver::learnts                                                  */
    uint64_t  n_learnts;
    const uint8_t *stencil;    /* cm.stencil(): 1 = deleted, indexed by C_REF (solvetypes.hpp:59,81) */
    uint64_t  stencil_size;    /* valid entries; refs at or beyond it count as not deleted         */
} sigma_cnf;
int  sigma_upload_cnf(sigma_ctx *ctx, const sigma_opts *o, const sigma_cnf *cnf, const sigma_formula *in);
int  sigma_extracted_sizes(sigma_ctx *ctx, uint32_t *n_clauses, uint64_t *n_buckets, uint64_t *n_literals);

/* ---- widening #2: newBeginning() on the device ---------------------------------
 * The simplified formula handed back as the solver's own clause database instead of an SCNF: what
 * newBeginning() (simplify.cpp:249-268) builds by calling addClause(SCLAUSE&) (sclause.cpp:22-62) for
 * every clause of the compacted SCNF, WITHOUT variable mapping (`mapped == false`) and without
 * --aggresivesort: one CLAUSE record (clause.hpp:47-59, constructor :77-92) per clause, back to back in
 * refs order - _k = 1 except learnt clauses with size > 2 and lbd > lbd_tier1, _b = (size == 2), _l / _lbd /
 * _used from the SCLAUSE, _pos = 2, literals copied -, the ref of every ORIGINAL clause in `orgs` and of
 * every LEARNT clause in `learnts` (list order = clause order), stats.literals.{original,learnt}, and,
 * for simplify() calls after the first, the variables of added clauses (mark_ssubsume, sclause.hpp:209-215)
 * as one byte per variable.  The two padding bits of a CLAUSE's first byte are written as 0 (the reference
 * leaves them uninitialised).  `rest` receives state / value / trail / model exactly like sigma_download
 * (its arena / refs are ignored). */
typedef struct sigma_cnf_out {
    uint32_t *cm;       uint64_t cm_cap;        /* buckets                                            */
    uint64_t *orgs;     uint64_t orgs_cap;
    uint64_t *learnts;  uint64_t learnts_cap;
    uint8_t  *subsume;                          /* nvars + 1 bytes, may be NULL                       */
    /* filled in (also by sigma_cnf_out_sizes, and on SIGMA_E_NOSPACE) */
    uint64_t  cm_buckets, n_orgs, n_learnts, lits_original, lits_learnt;
    uint64_t  last_ref;                         /* C_REF of the last record (its size = cm_buckets - last_ref - 4) */
} sigma_cnf_out;
int  sigma_cnf_out_sizes(sigma_ctx *ctx, sigma_cnf_out *sizes);                       /* after execute */
int  sigma_download_cnf(sigma_ctx *ctx, sigma_cnf_out *out, sigma_formula *rest);

/* ---- widening #3: the DIMACS parser on the device -------------------------------
 * Replaces Solver::parser() + makeClause() (reference dimacs.cpp:25-177, dimacs.hpp:57-74) for a whole file held in host
 * memory (the reference mmaps it, dimacs.cpp:37-39).  The header is read on the host; the clause part is copied to the
 * device and parsed there (byte-level state machine evaluated by a scan of per-chunk transition functions, integers, clause
 * boundaries, makeClause's literal checks, parse-time units acting on later clauses, CLAUSE records scan-allocated in file
 * order - seqfrost_b200/csrc/sigma_dimacs.cuh).  The result is what the reference's solver holds when parser() returns:
 *   cm / orgs   the CMM arena of CLAUSE records (clause.hpp:47-59; 4 header buckets {_k|_b flags, _sz, _pos = 2, _lbd = 0} +
 *               literals in file order) and Solver::orgs, word for word
 *   trail       the root units in the order parser() enqueued them; value[] / state[] as enqueueUnit leaves them
 *               (solve.hpp:373-398: value[lit] = 1, value[~lit] = 0, state = FROZEN_M)
 *   counters    formula.{units, binaries, ternaries, large, maxClauseSize}, stats.literals.original
 * Returns SIGMA_OK, SIGMA_UNSAT (an empty clause: parser() returns false, solve.cpp:52) or SIGMA_E_ARG with
 * sigma_last_error() naming what the reference would have stopped on (LOGERR: bad header, "expected a digit", "too many
 * variables", "too many clauses").  Not supported (SIGMA_E_ARG): a second header line, -parseincr.
 * Watches are not built here: the caller runs rebuildWT() (watch.cpp:135) or goes straight to sigma_upload_cnf. */
typedef struct sigma_dimacs {
    /* caller-owned output buffers for sigma_parse_download (ignored by sigma_parse_dimacs) */
    uint32_t *cm;     uint64_t cm_cap;       /* buckets */
    uint64_t *orgs;   uint64_t orgs_cap;
    uint32_t *trail;  uint64_t trail_cap;
    int8_t   *value;                          /* 2 nvars + 2 bytes, may be NULL */
    uint8_t  *state;                          /* nvars + 1 bytes, may be NULL */
    /* filled in by both calls */
    uint32_t  nvars, header_clauses;          /* inf.orgVars, inf.orgCls */
    uint64_t  cm_buckets, n_orgs, n_units;
    uint64_t  binaries, ternaries, large, literals;
    uint32_t  max_clause_size, unit_rounds;   /* unit_rounds: makeClause rounds needed (1 = no parse-time unit changed anything) */
    uint64_t  body_offset;                    /* first byte after the header */
    float     h2d_ms, device_ms;              /* copy of the file / kernels, CUDA events on the context's stream */
} sigma_dimacs;
int  sigma_parse_dimacs(sigma_ctx *ctx, const char *text, uint64_t bytes, sigma_dimacs *sizes);
int  sigma_parse_download(sigma_ctx *ctx, sigma_dimacs *out);
/* file -> pass without leaving the device: the parsed clause database becomes the input of extract() (as sigma_upload_cnf
 * would build it from host arrays: simplify.cpp:118-153), with the solver state of a fresh solver (nothing assigned, vorg =
 * identity, empty model, first simplify() call).  Only for files without unit clauses - otherwise the solver's own BCP() and
 * shrinkTop() come between parser() and simplify() (solve.cpp:52, simplify.cpp:161-162): SIGMA_E_UNSUPPORTED.  Afterwards
 * sigma_execute / sigma_download / sigma_download_cnf as usual. */
int  sigma_upload_parsed(sigma_ctx *ctx, const sigma_opts *o);

/* ---- newBeginning() under variable mapping: map(true) (map.cpp:24-46) --------------------------------------------------
 * When canMap() holds (can.hpp:43: more than map_perc of the variables are inactive) the reference calls map(true) instead of
 * newBeginning(): vmap.initiate(sp) numbers the active variables 1..n in order (map.hpp:78-95), and newBeginning() builds every
 * CLAUSE with mapLit(lit) = 2 vmap[var] + sign (map.hpp:52-76,195-205; sclause.cpp:34-35).  The mapping is computed on the host
 * from state[] / value[] AFTER the pass, so the order is: sigma_execute, sigma_download_state (state / value / trail / model
 * only), vmap.initiate on the host, sigma_set_var_map(map = *vmap, new_nvars = vmap.numVars()), then sigma_watches_sizes /
 * sigma_download_cnf (with rest->state == NULL: clause database only) / sigma_download_watches - all in the new numbering
 * (the `subsume` bytes stay indexed by the old variables, sclause.cpp:29-31).  map == NULL switches back. */
int  sigma_download_state(sigma_ctx *ctx, sigma_formula *out);
int  sigma_set_var_map(sigma_ctx *ctx, const uint32_t *map, uint32_t new_nvars);

/* ---- widening #4: rebuildWT() on the device -------------------------------------
 * The watch table rebuildWT(opts.simplify_priorbins) builds from the new clause database (simplify.cpp:218,
 * watch.cpp:135-157, attachBins / attachNonBins / attachClauses :22-133): for every clause in attach order - `priorbins`
 * bit 0: binaries of orgs and learnts first, then the longer clauses; bit 1: learnt binaries before original ones; 0: orgs
 * then learnts - sortClause (clause.cpp:116-173: best watches first, under the root assignment) on the CLAUSE record,
 * then ATTACH_TWO_WATCHES (watch.hpp:87-96): WATCH{ref, size, the other watched literal} pushed to wt[FLIP(c[0])] and
 * wt[FLIP(c[1])].  Handed back as one offsets array (nlit + 1 entries; list of literal l = records [offs[l], offs[l + 1]))
 * and the 16-byte WATCH records {C_REF ref; uint32 imp; int size} (watch.hpp:28-46) of every list in push order.
 * Call sigma_watches_sizes BEFORE sigma_download_cnf: sortClause reorders literals inside the CLAUSE records. */
typedef struct sigma_watches {
    uint32_t *offs;     uint64_t offs_cap;       /* entries */
    void     *records;  uint64_t records_cap;    /* 16-byte records */
    uint64_t  n_records; uint32_t nlit;          /* filled in */
} sigma_watches;
int  sigma_watches_sizes(sigma_ctx *ctx, int priorbins, sigma_watches *sizes);          /* after execute; builds them */
int  sigma_download_watches(sigma_ctx *ctx, sigma_watches *out);

/* pinned host memory for the hand-off buffers (the solver's SCNF arena allocator would call this
 * instead of malloc, malloc.hpp:60-69, so that H2D/D2H run at full PCIe rate) */
void *sigma_host_alloc(uint64_t bytes);
void  sigma_host_free(void *p);

/* pin / unpin memory the caller already owns (the solver's malloc'ed clause arena, malloc.hpp:60-69) for the duration
 * of an upload or download, so that the copy runs at PCIe rate without changing the solver's allocator.  Returns
 * SIGMA_OK or SIGMA_E_CUDA (the memory then simply stays pageable). */
int   sigma_host_register(void *p, uint64_t bytes);
int   sigma_host_unregister(void *p);

/* host-polled interrupt flag: the address of the reference's `bool interrupted` (solve.hpp:126, set by
 * Solver::interrupt() from the signal handlers, control.cpp:130-159).  Polled where the reference polls INTERRUPTED
 * (simplify.cpp:176,206; subsume.cpp:26; bounded.cpp:31): at the top of every phase, before SUB, before VE and after
 * the trailing prop(); sigma_execute then returns SIGMA_INTERRUPTED and leaves no result (the uploaded input stays
 * valid, so the call can be repeated).  NULL detaches. */
void sigma_set_interrupt_flag(sigma_ctx *ctx, const volatile unsigned char *flag);

/* stage-level parity hooks: stop after `stage` (SIGMA_T_* id of the stage) of `phase`, then copy
 * a named device array to the host.  Names: hdr lits ot_off ot_len ot_data occ order elected
 * state value trail model marks. phase < 0 disables. */
int  sigma_set_stop(sigma_ctx *ctx, int phase, int stage);
int  sigma_snapshot(sigma_ctx *ctx, const char *name, void *dst, uint64_t cap_bytes, uint64_t *nbytes);


#ifdef __cplusplus
}
#endif
#endif /* SIGMA_B200_H */
