/* sigma_types.h - device-side data layout of the B200 SIGmA pass.
 *
 * The reference keeps the simplifier formula as an AoS bump arena of SCLAUSE records
 * (reference src/sclause.hpp:45-49, src/simptypes.hpp:48-95) and the occurrence table as a
 * vector of growable vectors (simptypes.hpp:39-46).  On the device the same information lives
 * in flat arrays sized for HBM streaming:
 *
 *   clause store   hdr[i]  = {off, size, w0, sig}  one 16-byte record per clause id
 *                  lits[]  = literal pool; clause i owns lits[off .. off+size)
 *                  clause id i == position in the reference's `_refs` vector (allocation order)
 *   occurrence     ot_off[lit], ot_len[lit], ot_data[] (clause ids, ascending inside a list right
 *   table          after the build = the reference's push order, simplify.cpp:39-61)
 *   solver state   value[lit], state[var], marks[var], trail[], model[] (MODEL::resolved)
 *
 * w0 is SCLAUSE word0 verbatim: st:2 (0 ORIGINAL, 1 LEARNT, 2 DELETED) f:1 (molten) a:1 (added)
 * u:2 (usage) lbd:26, so the hand-off back to the solver needs no translation.
 */
#ifndef SIGMA_TYPES_H
#define SIGMA_TYPES_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SG_HD __host__ __device__ __forceinline__
#define SG_HDN __host__ __device__
#else
#define SG_HD inline
#define SG_HDN
#endif

#if !defined(__CUDACC__)
/* host-side stand-in for CUDA's uint4 (only used by the host emulation build under tests/) */
struct sg_u4 { uint32_t x, y, z, w; };
#else
typedef uint4 sg_u4;
#endif

/* ---- SCLAUSE word0 bits (constants.hpp:69-71, sclause.hpp:45) -------------------------------- */
#define SG_ST_MASK   0x3u
#define SG_LEARNT    0x1u
#define SG_DELETED   0x2u
#define SG_MOLTEN    0x4u
#define SG_ADDED     0x8u
#define SG_USAGE_SH  4
#define SG_LBD_SH    6
/* per-variable state bits (constants.hpp:59-61) */
#define SG_MELTED_M  1u
#define SG_FROZEN_M  2u

#define SG_ABS(l)   ((l) >> 1)
#define SG_FLIP(l)  ((l) ^ 1u)
#define SG_SIGN(l)  ((l) & 1u)
#define SG_HASH(l)  (1u << ((l) & 31u))

/* election status of a variable during one LCVE (lcve.cpp:52-83) */
enum {
    SG_E_UNDECIDED = 0,
    SG_E_SKIP      = 1,   /* never acts: assumed / inactive / no occurrences / behind the static break */
    SG_E_FROZEN    = 2,   /* frozen by an earlier acting candidate                                       */
    SG_E_NONE      = 3,   /* reached, depFreeze(P) failed: freezes nothing                               */
    SG_E_PONLY     = 4,   /* reached, depFreeze(P) ok, depFreeze(N) failed: P's freezes stay              */
    SG_E_BOTH      = 5,   /* elected                                                                     */
    SG_E_BREAK     = 6    /* reached a break condition (lcve.cpp:70,73)                                  */
};

/* what VE decided for an elected variable (bounded.cpp:41-182) */
enum {
    SG_VE_NONE = 0, SG_VE_PURE = 1, SG_VE_INVERTER = 2, SG_VE_RESOLVE = 3,
    SG_VE_SUBST = 4, SG_VE_CORE = 5
};

typedef struct SgOpts {
    int32_t phases, phase_lits_min, mu_pos, mu_neg, lcve_max_occurs, lcve_clause_max, lcve_min_vars;
    int32_t sub_max_occurs, sub_clause_max, ve_clause_max, xor_max_arity, bce_max_occurs;
    int32_t ere_max_occurs, ere_clause_max, collect_freq, lbd_tier1;
    int32_t sub_en, ve_en, ve_plus_en, ve_fun_en, ve_lbound_en, bce_en, ere_en, ere_extend, all_en, proof_en;
    int32_t firstcall;
} SgOpts;

/* scalars that kernels update and the host reads back at stage boundaries */
typedef struct SgScalars {
    uint32_t n_trail, propagated, unassigned, max_frozen;
    uint32_t unsat, n_units, wl_count[2];
    uint32_t brk_rank, n_acting, n_ponly, n_elected;
    uint32_t live_cls, max_score;
    unsigned long long live_lits;
    unsigned long long n_model;
    unsigned long long pures, inverters, andors, ites, xors, funs, resolutions, subsumed, strengthened;
    unsigned long long units_forced, sort_ties_gt24, elect_visits;
    unsigned long long ere_events, ere_dirty_vars, ere_full_pairs, ere_event_pairs;
    uint32_t core[12];        /* orgcore[] (lcve.cpp:24): variables holding marks 0..11 */
    uint32_t n_core, n_ot_add;
    uint32_t ere_npending, ere_overflow;
    uint32_t ere_first_dirty, ere_par_events, ere_list_start, ere_applied;
    uint32_t ere_pending[64];  /* elected positions that turned dirty during the replay and are not on its list */
} SgScalars;

/* one unit produced inside SUB/VE: applied in (elected index, sequence) order (SURVEY ledger 7) */
typedef struct SgUnit { uint32_t eidx, seq, lit, pad; } SgUnit;

typedef struct SgView {
    sg_u4    *hdr;
    uint32_t *lits;
    uint32_t *ot_off, *ot_len, *ot_data;
    int8_t   *value;
    uint8_t  *state;
    int8_t   *marks;
    const uint8_t  *ifrozen;
    const uint32_t *vorg;
    uint32_t *trail;
    uint32_t *model;
    uint32_t *occ;           /* occurrences per literal at election time (histogram.cpp:84-97) */
    uint32_t *scores, *order;
    uint32_t *rank;          /* election word per variable: rank | status << 29 (sigma_items.cuh) */
    uint32_t *elected;       /* elected variables in election order */
    uint32_t *acting;        /* acting candidates in election order, bit31 = P-only */
    /* per elected variable (index = position in `elected`) */
    uint32_t *ve_type, *ve_ncls, *ve_nlits, *ve_nmodel, *ve_aux;
    uint32_t *ve_cls_off, *ve_lit_off, *ve_mod_off;      /* exclusive scans of the three above */
    uint32_t *scr_off;       /* per elected variable scratch slice in `scratch` */
    uint32_t *scratch;
    SgUnit   *ubuf;
    uint32_t *ot_add;        /* (literal, clause id) pairs: occurrences created by substitution this phase */
    uint32_t  ot_add_cap;
    uint8_t  *ere_touched;   /* per clause: may be deleted/strengthened by ERE */
    uint8_t  *ere_flags;     /* per elected variable: 1 = has recorded pairs, 2 = dirty */
    uint32_t *ere_eidx;      /* variable -> position in `elected` */
    uint32_t *ere_owner;     /* per clause: earliest pending recorded pair that could write it (parallel rounds) */
    uint32_t *ere_done;      /* per recorded pair: 0 pending, 2 apply this round, 1 applied; NULL = no parallel rounds */
    const uint32_t *ere_filt; /* proposal pass only: {size (0 = deleted), sig} per clause, 8 bytes, a snapshot small enough for L2 */
    const uint8_t  *ere_size8; /* proposal pass only: min(size, 255) per clause, 0 = deleted: the first screen, one byte per clause */
    const uint8_t  *ere_maxsz; /* proposal pass only: largest ere_size8 in the occurrence list of each literal; NULL = not built */
    SgScalars *sc;
    uint32_t nvars, nlit;    /* nlit = 2*nvars + 2 */
    uint32_t n_cls;          /* clause ids in use */
    uint32_t pool_used;      /* literal pool bump pointer */
    uint32_t n_elected;
    uint32_t ubuf_cap;
    uint32_t pmax, nmax;     /* mu << multiplier (lcve.cpp:59-60) */
    uint32_t nbh_sort;       /* the fused per-variable kernel also sorts the two lists (sortOT) */
    uint32_t *sot_wl;        /* sortOT: literals whose lists have 25 .. SG_SOT_WIDE_MAX entries, collected by the thread-per-list kernel for the
                                warp-per-list kernel (count in sc->wl_count[1]); NULL = the thread-per-list kernel sorts them itself */
    uint32_t item_base;      /* be_items processes items [item_base, item_base + n) */
    uint32_t item_rev;       /* warp-per-variable kernels take their items last to first */
    uint32_t sortot_warp_min; /* sortOT: lists longer than this (and <= 24) are sorted by a warp, the others by one thread */
    uint64_t model_base;     /* model words present before this VE stage */
    SgOpts   o;
} SgView;

/* one frontier of prop (sigma_items.cuh, "prop by frontiers") */
typedef struct SgProp {
    uint32_t p0, p1;                      /* frontier */
    uint32_t *fpos;                       /* per literal: frontier position whose unit falsifies it, else SG_PP_NONE */
    unsigned long long *disc;             /* per literal: visit that makes it true in this frontier (the current guess) */
    unsigned long long *disc_new;         /* what this round reports */
    uint32_t *touched, *touched_new;      /* literals with an entry in disc / disc_new; counts in ctl[0], ctl[1] */
    uint32_t *cnt;                        /* per frontier position: units discovered by its visits */
    uint32_t *base;                       /* exclusive prefix sum of cnt; base[p1 - p0] = total */
    uint32_t *claim;                      /* one bit per clause id: rewritten in this frontier */
    uint32_t *ctl;                        /* 0,1: list counts; 2: guess changed; 3: conflict seen */
} SgProp;

#endif
