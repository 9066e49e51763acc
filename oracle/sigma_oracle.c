/* sigma_oracle.c - CPU ORACLE of the SIGmA pass.  TEST INFRASTRUCTURE ONLY (see sigma_oracle.h).
 *
 * Restates, in plain C over a structure-of-arrays clause store, what the reference
 * (muhos/SeqFROST v1.0.2, /root/reference/src) does between awaken() and the hand-back:
 * every function cites the reference file:line whose behaviour it follows.  Nothing here is
 * used by the product path.
 *
 * Store layout (identical to the device layout of seqfrost_b200/csrc, so arrays diff 1:1):
 *   clause id = position in the reference's scnf.refs() order
 *   c_w0[id]   word0 of SCLAUSE verbatim: st:2 (0 ORIGINAL,1 LEARNT,2 DELETED) f:1 a:1 u:2 lbd:26
 *   c_sig[id]  signature, c_size[id] size, c_off[id] first literal in the `lits` pool
 *   occurrence table: CSR ot_off[lit] / ot_len[lit] / ot_data[] of clause ids, literal-indexed
 */
#include "sigma_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ST_MASK 3u
#define ST_LEARNT 1u
#define ST_DELETED 2u
#define F_MOLTEN 4u
#define F_ADDED 8u
#define MELTED_M 1
#define FROZEN_M 2
#define MAXFUNVAR 12
#define FUNTABLEN 64
#define INSORT_THR 24
#define MAPHASH(l) (1u << ((l) & 31u))
#define UNASSIGNED(v) ((v) & -2)

typedef uint64_t Fun[FUNTABLEN];

struct sgo_ctx {
    sigma_opts o;
    uint32_t V;
    /* clause store */
    uint32_t n_cls, cls_cap;
    uint32_t *c_off, *c_w0, *c_sig;
    int32_t *c_size;
    uint32_t *lits;
    uint64_t lits_used, lits_cap;
    /* occurrence table */
    uint32_t *ot_off, *ot_len, *ot_data;
    uint64_t ot_cap;
    uint32_t *occ;            /* histogram per literal: occ[2v] = ps, occ[2v+1] = ns */
    uint32_t *scores, *order, *tmp_order;
    uint8_t *state, *frozen, *ifrozen;
    int8_t *value, *marks;
    uint32_t *trail, trail_size, propagated;
    uint32_t *stack, stack_n;
    uint32_t orgcore[MAXFUNVAR];
    int has_core;
    uint32_t *elected, n_elected;
    uint32_t *vorg;
    uint32_t *model;
    uint64_t model_size, model_cap;
    uint32_t unassigned, max_frozen, max_melted, simplify_calls;
    uint64_t n_clauses, n_literals;      /* inf.nClauses / inf.nLiterals */
    int phase, multiplier;
    int stop_phase, stop_stage, stopped;
    int unsat;
    uint32_t *tmp_lits;                  /* resolvent scratch */
    uint32_t tmp_cap;
    sigma_report rep;
    /* pristine copy for re-runs */
};

/* ------------------------------------------------------------------------------------------ */
static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "sigma_oracle: out of memory\n"); abort(); } return q; }
static void *xcalloc(size_t n, size_t s) { void *q = calloc(n ? n : 1, s); if (!q) { fprintf(stderr, "sigma_oracle: out of memory\n"); abort(); } return q; }

static inline int cl_original(const sgo_ctx *c, uint32_t id) { return (c->c_w0[id] & ST_MASK) == 0; }   /* sclause.hpp:138 */
static inline int cl_deleted(const sgo_ctx *c, uint32_t id) { return (c->c_w0[id] & ST_DELETED) != 0; }  /* sclause.hpp:139 */
static inline int cl_learnt(const sgo_ctx *c, uint32_t id) { return (c->c_w0[id] & ST_LEARNT) != 0; }    /* sclause.hpp:140 */
static inline int cl_molten(const sgo_ctx *c, uint32_t id) { return (c->c_w0[id] & F_MOLTEN) != 0; }
static inline void cl_melt(sgo_ctx *c, uint32_t id) { c->c_w0[id] |= F_MOLTEN; }                         /* sclause.hpp:128 */
static inline void cl_freeze(sgo_ctx *c, uint32_t id) { c->c_w0[id] &= ~F_MOLTEN; }                      /* sclause.hpp:127 */
static inline void cl_delete(sgo_ctx *c, uint32_t id) { c->c_w0[id] = (c->c_w0[id] & ~ST_MASK) | ST_DELETED; } /* sclause.hpp:130 */
static inline uint32_t *cl_lits(const sgo_ctx *c, uint32_t id) { return c->lits + c->c_off[id]; }
static inline uint32_t *ol(const sgo_ctx *c, uint32_t lit) { return c->ot_data + c->ot_off[lit]; }

static void model_push(sgo_ctx *c, uint32_t x)
{
    if (c->model_size == c->model_cap) {
        c->model_cap = c->model_cap ? c->model_cap * 2 : 1024;
        c->model = (uint32_t *)xrealloc(c->model, c->model_cap * 4);
    }
    c->model[c->model_size++] = x;
}

/* MODEL::saveLiteral, model.hpp:65-71: literals are stored in ORIGINAL variable ids */
static inline void save_literal(sgo_ctx *c, uint32_t lit) { model_push(c, (c->vorg[lit >> 1] << 1) | (lit & 1u)); }
/* MODEL::saveWitness, model.hpp:72-76 */
static inline void save_witness(sgo_ctx *c, uint32_t w) { save_literal(c, w); model_push(c, 1); }
/* MODEL::saveClause, model.hpp:86-101 */
static void save_clause(sgo_ctx *c, uint32_t id, uint32_t witlit)
{
    const uint64_t last = c->model_size;
    const uint32_t *l = cl_lits(c, id);
    const int size = c->c_size[id];
    int pos = -1;
    for (int i = 0; i < size; ++i) { save_literal(c, l[i]); if (l[i] == witlit) pos = i; }
    if (pos > 0) { uint32_t t = c->model[last + pos]; c->model[last + pos] = c->model[last]; c->model[last] = t; }
    model_push(c, (uint32_t)size);
}

/* Solver::enqueueUnit + learnUnit + markFrozen, solve.hpp:373-398,459-464,485-490 */
static void enqueue_unit(sgo_ctx *c, uint32_t lit)
{
    const uint32_t v = lit >> 1;
    c->state[v] = FROZEN_M;
    c->max_frozen++;
    c->value[lit] = 1;
    c->value[lit ^ 1u] = 0;
    c->trail[c->trail_size++] = lit;
    c->unassigned--;
    c->rep.units++;
}

/* "if UNASSIGNED enqueue; else if false -> contradiction" idiom of strengthen (subsume.cpp:61-69),
 * addResolvent (bounded.cpp:283-293) and substitute_single (equivalence.hpp:72-77) */
static void new_unit(sgo_ctx *c, uint32_t unit)
{
    const int8_t val = c->value[unit];
    if (UNASSIGNED(val)) enqueue_unit(c, unit);
    else if (!val) c->unsat = 1;
}

static void ensure_store(sgo_ctx *c, uint32_t more_cls, uint64_t more_lits)
{
    if ((uint64_t)c->n_cls + more_cls > c->cls_cap) {
        uint64_t cap = c->cls_cap ? c->cls_cap : 16;
        while (cap < (uint64_t)c->n_cls + more_cls) cap *= 2;
        c->cls_cap = (uint32_t)cap;
        c->c_off = (uint32_t *)xrealloc(c->c_off, cap * 4);
        c->c_w0 = (uint32_t *)xrealloc(c->c_w0, cap * 4);
        c->c_sig = (uint32_t *)xrealloc(c->c_sig, cap * 4);
        c->c_size = (int32_t *)xrealloc(c->c_size, cap * 4);
    }
    if (c->lits_used + more_lits > c->lits_cap) {
        uint64_t cap = c->lits_cap ? c->lits_cap : 64;
        while (cap < c->lits_used + more_lits) cap *= 2;
        c->lits_cap = cap;
        c->lits = (uint32_t *)xrealloc(c->lits, cap * 4);
    }
}

/* ---- a11: Solver::shrinkSimp, simplify.cpp:235-247 (SCLAUSE copy ctor sclause.hpp:76-89) ---- */
static void compact(sgo_ctx *c)
{
    uint32_t j = 0;
    uint64_t w = 0;
    for (uint32_t i = 0; i < c->n_cls; ++i) {
        if (cl_deleted(c, i)) continue;
        const int sz = c->c_size[i];
        memmove(c->lits + w, c->lits + c->c_off[i], (size_t)sz * 4);
        c->c_off[j] = (uint32_t)w;
        c->c_size[j] = sz;
        c->c_w0[j] = c->c_w0[i];
        c->c_sig[j] = c->c_sig[i];
        w += sz;
        ++j;
    }
    c->n_cls = j;
    c->lits_used = w;
}

/* ---- a1: Solver::createOT, simplify.cpp:39-61: lists ordered by clause position ----------- */
static void create_ot(sgo_ctx *c)
{
    const uint32_t nl = 2 * c->V + 2;
    memset(c->ot_len, 0, (size_t)nl * 4);
    uint64_t total = 0;
    for (uint32_t i = 0; i < c->n_cls; ++i) {
        if (cl_deleted(c, i)) continue;
        const uint32_t *l = cl_lits(c, i);
        for (int k = 0; k < c->c_size[i]; ++k) c->ot_len[l[k]]++;
        total += c->c_size[i];
    }
    if (total > c->ot_cap) { c->ot_cap = total; c->ot_data = (uint32_t *)xrealloc(c->ot_data, total * 4); }
    uint32_t run = 0;
    for (uint32_t l = 0; l < nl; ++l) { c->ot_off[l] = run; run += c->ot_len[l]; c->ot_len[l] = 0; }
    c->ot_off[nl] = run;
    for (uint32_t i = 0; i < c->n_cls; ++i) {
        if (cl_deleted(c, i)) continue;
        const uint32_t *l = cl_lits(c, i);
        for (int k = 0; k < c->c_size[i]; ++k) c->ot_data[c->ot_off[l[k]] + c->ot_len[l[k]]++] = i;
    }
}

/* Solver::reduceOL / reduceOT, simplify.cpp:63-83 */
static void reduce_ot(sgo_ctx *c)
{
    for (uint32_t l = 2; l < 2 * c->V + 2; ++l) {
        uint32_t *d = ol(c, l), j = 0;
        for (uint32_t i = 0; i < c->ot_len[l]; ++i) if (!cl_deleted(c, d[i])) d[j++] = d[i];
        c->ot_len[l] = j;
    }
}

/* ---- a6: Solver::prop / propClause, propelim.cpp:23-81 ------------------------------------ */
static int prop(sgo_ctx *c)
{
    uint32_t numforced = c->propagated;
    while (c->propagated < c->trail_size) {
        const uint32_t assign = c->trail[c->propagated++], flipped = assign ^ 1u;
        uint32_t *fot = ol(c, flipped);
        for (uint32_t i = 0; i < c->ot_len[flipped]; ++i) {
            const uint32_t id = fot[i];
            if (cl_deleted(c, id)) continue;
            /* propClause: compacts in place while scanning, returns early on a true literal */
            uint32_t *l = cl_lits(c, id), sig = 0;
            const int n = c->c_size[id];
            int j = 0, sat = 0;
            for (int k = 0; k < n; ++k) {
                const uint32_t other = l[k];
                if (other != flipped) {
                    if (c->value[other] > 0) { sat = 1; break; }
                    l[j++] = other;
                    sig |= MAPHASH(other);
                }
            }
            if (sat) { cl_delete(c, id); continue; }
            c->c_sig[id] = sig;
            c->c_size[id] = n - 1;
            if (n - 1 == 0) return 0;
            if (n - 1 == 1) {
                const uint32_t unit = l[0];
                if (UNASSIGNED(c->value[unit])) enqueue_unit(c, unit);
                else return 0;
            }
        }
        c->ot_len[flipped] = 0;
        uint32_t *aot = ol(c, assign);
        for (uint32_t i = 0; i < c->ot_len[assign]; ++i) cl_delete(c, aot[i]);   /* toblivion(scnf, ot[assign]) */
        c->ot_len[assign] = 0;
    }
    numforced = c->propagated - numforced;
    if (numforced && !c->o.sub_en) reduce_ot(c);
    return 1;
}

/* ---- a2/a3: histCNF + write_scores + stable ascending sort, histogram.cpp:84-97,
 *      histogram.hpp:23-50, lcve.cpp:36-39, key.hpp:138-144 ---------------------------------- */
static void order_variables(sgo_ctx *c)
{
    const uint32_t V = c->V;
    memset(c->occ, 0, (size_t)(2 * V + 2) * 4);
    for (uint32_t i = 0; i < c->n_cls; ++i) {
        if (cl_deleted(c, i)) continue;
        const uint32_t *l = cl_lits(c, i);
        for (int k = 0; k < c->c_size[i]; ++k) c->occ[l[k]]++;
    }
    for (uint32_t v = 1; v <= V; ++v) {
        c->scores[v] = c->occ[2 * v] * c->occ[2 * v + 1];    /* uint32 product, wraps like hist_score */
        c->order[v - 1] = v;
    }
    /* stable LSD radix sort by score: equal scores keep ascending variable id */
    uint32_t *a = c->order, *b = c->tmp_order;
    for (int shift = 0; shift < 32; shift += 8) {
        size_t cnt[257] = { 0 };
        for (uint32_t i = 0; i < V; ++i) cnt[((c->scores[a[i]] >> shift) & 255u) + 1]++;
        if (cnt[((c->scores[a[0]] >> shift) & 255u) + 1] == V) continue;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (uint32_t i = 0; i < V; ++i) b[cnt[(c->scores[a[i]] >> shift) & 255u]++] = a[i];
        uint32_t *t = a; a = b; b = t;
    }
    if (a != c->order) memcpy(c->order, a, (size_t)V * 4);
}

/* Solver::depFreeze, lcve.cpp:121-145 */
static int dep_freeze(sgo_ctx *c, uint32_t lit, uint32_t cand, uint32_t pmax, uint32_t nmax)
{
    const uint32_t first = c->stack_n;
    const uint32_t *d = ol(c, lit);
    for (uint32_t i = 0; i < c->ot_len[lit]; ++i) {
        const uint32_t id = d[i];
        if (cl_deleted(c, id)) continue;
        if (c->c_size[id] > c->o.lcve_clause_max) {
            for (uint32_t k = first; k < c->stack_n; ++k) c->frozen[c->stack[k]] = 0;
            c->stack_n = first;
            return 0;
        }
        const uint32_t *l = cl_lits(c, id);
        for (int k = 0; k < c->c_size[id]; ++k) {
            const uint32_t v = l[k] >> 1;
            if (!c->frozen[v] && v != cand && (c->occ[2 * v] < pmax || c->occ[2 * v + 1] < nmax)) {
                c->frozen[v] = 1;
                c->stack[c->stack_n++] = v;
            }
        }
    }
    return 1;
}

/* Solver::clearMapFrozen, solve.hpp:337-345 */
static void clear_map_frozen(sgo_ctx *c)
{
    if (c->has_core) {
        for (int i = 0; i < MAXFUNVAR; ++i) if (c->orgcore[i] <= c->V) c->marks[c->orgcore[i]] = -1;
        c->has_core = 0;
    }
}

/* ---- a4: Solver::LCVE, lcve.cpp:29-119 ----------------------------------------------------- */
static int lcve(sgo_ctx *c)
{
    order_variables(c);
    if (c->stop_phase == c->phase && c->stop_stage == SIGMA_T_VO) { c->stopped = 1; return 1; }
    const uint32_t maxoccurs = (uint32_t)c->o.lcve_max_occurs;
    const uint32_t pmax = (uint32_t)c->o.mu_pos << c->multiplier;
    const uint32_t nmax = (uint32_t)c->o.mu_neg << c->multiplier;
    c->stack_n = 0;
    c->n_elected = 0;
    for (uint32_t i = 0; i < c->V; ++i) {
        const uint32_t cand = c->order[i];
        if (c->ifrozen && c->ifrozen[cand]) continue;        /* iassumed, solve.hpp:816 */
        if (c->state[cand]) continue;
        if (c->frozen[cand]) continue;
        const uint32_t p = 2 * cand, n = p + 1;
        if (!c->occ[p] && !c->occ[n]) continue;
        if (c->occ[p] > maxoccurs || c->occ[n] > maxoccurs) break;
        if (c->ot_len[p] >= pmax && c->ot_len[n] >= nmax) break;
        if (dep_freeze(c, p, cand, pmax, nmax) && dep_freeze(c, n, cand, pmax, nmax))
            c->elected[c->n_elected++] = cand;
    }
    if (c->o.ve_fun_en) {                                    /* lcve.cpp:85-102 */
        clear_map_frozen(c);
        int coresize = 0;
        for (uint32_t k = 0; k < c->stack_n && coresize < MAXFUNVAR; ++k) {
            const uint32_t v = c->stack[k];
            c->marks[v] = (int8_t)coresize;
            c->orgcore[coresize++] = v;
            c->has_core = 1;
        }
        /* the reference leaves stale ids in orgcore[coresize..11]; they only ever get their mark
         * reset to -1, which they already have - normalise them to an out-of-range id here */
        for (int k = coresize; k < MAXFUNVAR; ++k) c->orgcore[k] = 0xFFFFFFFFu;
    }
    for (uint32_t k = 0; k < c->stack_n; ++k) c->frozen[c->stack[k]] = 0;     /* clearFrozen, solve.hpp:318 */
    return c->n_elected >= (uint32_t)c->o.lcve_min_vars;
}

/* ---- a5: CNF_CMP_KEY simplify.hpp:30-47; sortOT simplify.cpp:85-100; SORTCMP sort.hpp:64-72 - */
static inline int key_less(const sgo_ctx *c, uint32_t a, uint32_t b)
{
    const int xs = c->c_size[a], ys = c->c_size[b];
    if (xs < ys) return 1;
    if (xs > ys) return 0;
    const uint32_t *x = cl_lits(c, a), *y = cl_lits(c, b);
    if (x[0] < y[0]) return 1;
    if (x[0] > y[0]) return 0;
    if (x[xs - 1] < y[ys - 1]) return 1;
    if (x[xs - 1] > y[ys - 1]) return 0;
    return c->c_sig[a] < c->c_sig[b];
}

static void sort_list(sgo_ctx *c, uint32_t lit)
{
    uint32_t *d = ol(c, lit);
    const uint32_t n = c->ot_len[lit];
    if (n < 2) return;
    /* lists <= 24: the reference's insertionSort (stable).  Longer lists: the reference calls
     * std::sort, whose order among key ties is implementation-defined (SURVEY ledger 3); this
     * port (and the CUDA path) use the stable total order (key, OT position) and count such
     * ties so a mismatch against the real reference can be attributed. */
    for (uint32_t i = 1; i < n; ++i) {
        const uint32_t t = d[i];
        uint32_t j = i;
        while (j > 0 && key_less(c, t, d[j - 1])) { d[j] = d[j - 1]; --j; }
        d[j] = t;
    }
    if (n > INSORT_THR)
        for (uint32_t i = 1; i < n; ++i)
            if (!key_less(c, d[i - 1], d[i])) c->rep.sort_ties_gt24++;
}

static void sort_ot(sgo_ctx *c)
{
    for (uint32_t i = 0; i < c->n_elected; ++i) {
        sort_list(c, 2 * c->elected[i]);
        sort_list(c, 2 * c->elected[i] + 1);
    }
}

/* ---- a7: SUB ------------------------------------------------------------------------------- */
static inline int sig_sub(uint32_t A, uint32_t B) { return !(A & ~B); }                       /* simplify.hpp:119 */
static inline int sig_selfsub(uint32_t A, uint32_t B)                                          /* simplify.hpp:121-125 */
{
    const uint32_t Bt = B | ((B & 0xAAAAAAAAu) >> 1) | ((B & 0x55555555u) << 1);
    return !(A & ~Bt);
}

/* sub(SCLAUSE&, SCLAUSE&), subsume.hpp:38-63 */
static int lits_sub(const uint32_t *d1, int n1, const uint32_t *d2, int n2)
{
    int i = 0, j = 0, sub = 0;
    while (i < n1 && j < n2) {
        if (d1[i] < d2[j]) i++;
        else if (d2[j] < d1[i]) j++;
        else { sub++; i++; j++; }
    }
    return sub == n1;
}

/* selfsub(x, fx, subsuming, subsumed), subsume.hpp:93-137 */
static int lits_selfsub(uint32_t x, uint32_t fx, const uint32_t *d1, int n1, const uint32_t *d2, int n2)
{
    int i = 0, j = 0, sub = 0, self = 0;
    while (i < n1 && j < n2) {
        const uint32_t l1 = d1[i], l2 = d2[j];
        if (l1 == fx) i++;
        else if (l2 == x) { self = 1; j++; }
        else if (l1 < l2) i++;
        else if (l2 < l1) j++;
        else { sub++; i++; j++; }
    }
    if (sub + 1 == n1) {
        if (self) return 1;
        for (; j < n2; ++j) if (d2[j] == x) return 1;
    }
    return 0;
}

/* Solver::bumpShrunken, solve.hpp:725-737 */
static void bump_shrunken(sgo_ctx *c, uint32_t id)
{
    const int old_lbd = (int)(c->c_w0[id] >> 6);
    if (old_lbd <= c->o.lbd_tier1) return;
    const int size = c->c_size[id] - 1;
    const int new_lbd = size < old_lbd ? size : old_lbd;
    if (new_lbd >= old_lbd) return;
    c->c_w0[id] = (c->c_w0[id] & 0xFu) | (1u << 4) | ((uint32_t)new_lbd << 6);   /* lbd, usage = USAGET3 */
}

/* Solver::strengthen, subsume.cpp:45-76 */
static void strengthen(sgo_ctx *c, uint32_t id, uint32_t me)
{
    uint32_t *l = cl_lits(c, id), sig = 0;
    const int n = c->c_size[id];
    int j = 0;
    for (int k = 0; k < n; ++k) if (l[k] != me) { l[j++] = l[k]; sig |= MAPHASH(l[k]); }
    c->c_sig[id] = sig;
    c->c_size[id] = n - 1;                /* c.pop() */
    if (n - 1 == 1) new_unit(c, l[0]);
    else if (cl_learnt(c, id)) bump_shrunken(c, id);
}

/* selfsubsume, subsume.hpp:158-177 */
static int selfsubsume(sgo_ctx *c, uint32_t x, uint32_t fx, uint32_t list_lit, uint32_t cand)
{
    const int candsz = c->c_size[cand];
    const uint32_t candsig = c->c_sig[cand];
    const uint32_t *d = ol(c, list_lit);
    for (uint32_t j = 0; j < c->ot_len[list_lit]; ++j) {
        const uint32_t s = d[j];
        const int subsize = c->c_size[s];
        if (subsize > candsz) break;
        if (cl_deleted(c, s)) continue;
        if (subsize > 1 && sig_selfsub(c->c_sig[s], candsig) &&
            lits_selfsub(x, fx, cl_lits(c, s), subsize, cl_lits(c, cand), candsz)) {
            strengthen(c, cand, x);
            cl_melt(c, cand);
            return 1;
        }
    }
    return 0;
}

/* subsume, subsume.hpp:139-156 */
static int subsume(sgo_ctx *c, uint32_t list_lit, uint32_t end, uint32_t cand)
{
    const int candsz = c->c_size[cand];
    const uint32_t *d = ol(c, list_lit);
    for (uint32_t j = 0; j < end; ++j) {
        const uint32_t s = d[j];
        if (cl_deleted(c, s)) continue;
        if (cl_molten(c, cand) && c->c_size[s] > candsz) continue;
        if (c->c_size[s] > 1 && sig_sub(c->c_sig[s], c->c_sig[cand]) &&
            lits_sub(cl_lits(c, s), c->c_size[s], cl_lits(c, cand), candsz)) {
            if (cl_learnt(c, s) && cl_original(c, cand)) c->c_w0[s] &= ~ST_MASK;   /* set_status(ORIGINAL) */
            if (!cl_deleted(c, cand)) cl_delete(c, cand);                           /* removeClause, solve.hpp:717 */
            return 1;
        }
    }
    return 0;
}

/* updateOL, subsume.hpp:26-36 */
static void update_ol(sgo_ctx *c, uint32_t lit)
{
    uint32_t *d = ol(c, lit), j = 0;
    for (uint32_t i = 0; i < c->ot_len[lit]; ++i) {
        const uint32_t id = d[i];
        if (cl_molten(c, id)) cl_freeze(c, id);
        else if (!cl_deleted(c, id)) d[j++] = id;
    }
    c->ot_len[lit] = j;
}

/* self_sub_x, subsume.hpp:179-218 */
static void self_sub_x(sgo_ctx *c, uint32_t p)
{
    const uint32_t n = p | 1u;
    const int maxsz = c->o.sub_clause_max;
    for (int side = 0; side < 2; ++side) {
        const uint32_t me = side ? n : p, opp = side ? p : n;
        const uint32_t *d = ol(c, me);
        for (uint32_t i = 0; i < c->ot_len[me]; ++i) {
            const uint32_t id = d[i];
            if (c->c_size[id] > maxsz) break;
            if (cl_deleted(c, id)) continue;
            if (selfsubsume(c, me, opp, opp, id)) c->rep.strengthened++;
            if (c->unsat) return;
            if (subsume(c, me, i, id)) c->rep.subsumed++;
        }
        update_ol(c, me);
    }
}

/* Solver::SUB, subsume.cpp:23-43 */
static void sub_stage(sgo_ctx *c)
{
    if (!(c->o.sub_en || c->o.ve_plus_en)) return;
    const uint32_t maxoccurs = (uint32_t)c->o.sub_max_occurs;
    for (uint32_t i = 0; i < c->n_elected && !c->unsat; ++i) {
        const uint32_t p = 2 * c->elected[i], n = p + 1;
        if (c->ot_len[p] <= maxoccurs && c->ot_len[n] <= maxoccurs) self_sub_x(c, p);
    }
}

/* ---- a8/a9: BVE ---------------------------------------------------------------------------- */
/* merge (count), simplify.hpp:228-260 */
static int merge_count(uint32_t x, const uint32_t *d1, int n1, const uint32_t *d2, int n2)
{
    int i = 0, j = 0, len = n1 + n2 - 2;
    while (i < n1 && j < n2) {
        const uint32_t l1 = d1[i], l2 = d2[j], v1 = l1 >> 1, v2 = l2 >> 1;
        if (v1 == x) i++;
        else if (v2 == x) j++;
        else if ((l1 ^ l2) == 1u) return 0;
        else if (v1 < v2) i++;
        else if (v2 < v1) j++;
        else { i++; j++; len--; }
    }
    return len;
}

/* merge (emit), simplify.hpp:181-226; returns -1 for a tautology, else the resolvent size */
static int merge_emit(uint32_t x, const uint32_t *d1, int n1, const uint32_t *d2, int n2, uint32_t *out)
{
    int i = 0, j = 0, n = 0;
    while (i < n1 && j < n2) {
        const uint32_t l1 = d1[i], l2 = d2[j], v1 = l1 >> 1, v2 = l2 >> 1;
        if (v1 == x) i++;
        else if (v2 == x) j++;
        else if ((l1 ^ l2) == 1u) return -1;
        else if (v1 < v2) { i++; out[n++] = l1; }
        else if (v2 < v1) { j++; out[n++] = l2; }
        else { i++; j++; out[n++] = l1; }
    }
    for (; i < n1; ++i) if ((d1[i] >> 1) != x) out[n++] = d1[i];
    for (; j < n2; ++j) if ((d2[j] >> 1) != x) out[n++] = d2[j];
    return n;
}

enum { PAIR_ALL = 0, PAIR_SUBST = 1, PAIR_CORE = 2 };
static inline int pair_ok(int mode, int a, int b)
{
    if (mode == PAIR_SUBST) return a != b;          /* simplify.hpp:316, bounded.cpp:208 */
    if (mode == PAIR_CORE) return !a || !b;         /* function.hpp:108, bounded.cpp:240 */
    return 1;
}

static int count_lits_before(sgo_ctx *c, uint32_t lit)      /* countLitsBefore, simplify.hpp:289-296 */
{
    int n = 0;
    const uint32_t *d = ol(c, lit);
    for (uint32_t i = 0; i < c->ot_len[lit]; ++i) if (cl_original(c, d[i])) n += c->c_size[d[i]];
    return n;
}

/* countSubstituted simplify.hpp:298-328 / countResolvents(clsbefore) :355-383 /
 * countMinResolvents function.hpp:95-126.  Returns 1 when the bound is violated. */
static int count_bounded(sgo_ctx *c, uint32_t x, int clsbefore, uint32_t me, uint32_t other, int mode, int *ncls, int *nlits)
{
    const int rlimit = c->o.ve_clause_max;
    const uint32_t *dm = ol(c, me), *dt = ol(c, other);
    for (uint32_t i = 0; i < c->ot_len[me]; ++i) {
        const uint32_t ci = dm[i];
        if (!cl_original(c, ci)) continue;
        const int a = cl_molten(c, ci);
        for (uint32_t j = 0; j < c->ot_len[other]; ++j) {
            const uint32_t cj = dt[j];
            if (!cl_original(c, cj)) continue;
            const int b = cl_molten(c, cj);
            int rsize;
            if (pair_ok(mode, a, b) && (rsize = merge_count(x, cl_lits(c, ci), c->c_size[ci], cl_lits(c, cj), c->c_size[cj])) > 1) {
                if (++*ncls > clsbefore || (rlimit && rsize > rlimit)) return 1;
                *nlits += rsize;
            }
        }
    }
    if (c->o.ve_lbound_en) {
        const int before = count_lits_before(c, me) + count_lits_before(c, other);
        if (*nlits > before) return 1;
    }
    return 0;
}

/* countResolvents (1 x m, size bound only), simplify.hpp:330-353. Returns 1 on success. */
static int count_unbounded(sgo_ctx *c, uint32_t x, uint32_t me, uint32_t other, int *ncls, int *nlits)
{
    const int rlimit = c->o.ve_clause_max;
    const uint32_t *dm = ol(c, me), *dt = ol(c, other);
    for (uint32_t i = 0; i < c->ot_len[me]; ++i) {
        const uint32_t ci = dm[i];
        if (!cl_original(c, ci)) continue;
        for (uint32_t j = 0; j < c->ot_len[other]; ++j) {
            const uint32_t cj = dt[j];
            if (!cl_original(c, cj)) continue;
            const int rsize = merge_count(x, cl_lits(c, ci), c->c_size[ci], cl_lits(c, cj), c->c_size[cj]);
            if (rsize > 1) {
                if (rlimit && rsize > rlimit) return 0;
                ++*ncls;
                *nlits += rsize;
            }
        }
    }
    return 1;
}

static void freeze_binaries(sgo_ctx *c, uint32_t lit)       /* simplify.hpp:262-269 */
{
    const uint32_t *d = ol(c, lit);
    for (uint32_t i = 0; i < c->ot_len[lit]; ++i) if (cl_original(c, d[i]) && c->c_size[d[i]] == 2) cl_freeze(c, d[i]);
}
static void freeze_all(sgo_ctx *c, uint32_t lit)            /* simplify.hpp:271-278 */
{
    const uint32_t *d = ol(c, lit);
    for (uint32_t i = 0; i < c->ot_len[lit]; ++i) if (cl_original(c, d[i])) cl_freeze(c, d[i]);
}
static int count_orgs(sgo_ctx *c, uint32_t lit)             /* simplify.hpp:280-287 */
{
    int n = 0;
    const uint32_t *d = ol(c, lit);
    for (uint32_t i = 0; i < c->ot_len[lit]; ++i) if (cl_original(c, d[i])) n++;
    return n;
}

/* toblivion(p,poss,n,negs,..), simplify.hpp:385-423 */
static void toblivion(sgo_ctx *c, uint32_t p, uint32_t n, int pOrgs, int nOrgs)
{
    const uint32_t save = (pOrgs > nOrgs) ? n : p, oth = save ^ 1u;
    const uint32_t *d = ol(c, save);
    for (uint32_t i = 0; i < c->ot_len[save]; ++i) {
        if (cl_original(c, d[i])) save_clause(c, d[i], save);
        cl_delete(c, d[i]);
    }
    save_witness(c, oth);
    c->ot_len[save] = 0;
    d = ol(c, oth);
    for (uint32_t i = 0; i < c->ot_len[oth]; ++i) cl_delete(c, d[i]);
    c->ot_len[oth] = 0;
}

/* SCLAUSE::has, sclause.hpp:163-183 (membership in a sorted clause; linear for size 2) */
static int cl_has(const sgo_ctx *c, uint32_t id, uint32_t lit)
{
    const uint32_t *l = cl_lits(c, id);
    const int n = c->c_size[id];
    if (n == 2) return l[0] == lit || l[1] == lit;
    int low = 0, high = n - 1;
    if (n <= 0) return 0;
    uint32_t first = l[low], last = l[high];
    while (first <= lit && last >= lit) {
        const int mid = (low + high) >> 1;
        const uint32_t m = l[mid];
        if (m < lit) first = l[low = mid + 1];
        else if (m > lit) last = l[high = mid - 1];
        else return 1;
    }
    return l[low] == lit;
}

static void sort_u32(uint32_t *a, int n)
{
    for (int i = 1; i < n; ++i) { uint32_t t = a[i]; int j = i; while (j > 0 && t < a[j - 1]) { a[j] = a[j - 1]; --j; } a[j] = t; }
}

/* find_sfanin + find_BN_gate, equivalence.hpp:103-161 */
static uint32_t find_bn_gate(sgo_ctx *c, uint32_t p, uint32_t n)
{
    const uint32_t *dp = ol(c, p), *dn = ol(c, n);
    if (c->c_size[dp[0]] > 2 || c->c_size[dn[0]] > 2) return 0;
    uint32_t imp = 0;
    int nimps = 0, aborted = 0;
    for (uint32_t i = 0; i < c->ot_len[p]; ++i) {
        const uint32_t id = dp[i];
        if (cl_original(c, id) && c->c_size[id] == 2) {
            const uint32_t *l = cl_lits(c, id);
            imp = (l[0] ^ l[1] ^ p) ^ 1u;
            cl_melt(c, id);
            nimps++;
        }
        if (nimps > 1) { aborted = 1; break; }
    }
    uint32_t first = aborted ? 0 : imp;
    if (first) {
        uint32_t second = n;
        const uint32_t def = first;
        if (second < first) { first = second; second = def; }
        for (uint32_t i = 0; i < c->ot_len[n]; ++i) {
            const uint32_t id = dn[i];
            const uint32_t *l = cl_lits(c, id);
            if (cl_original(c, id) && c->c_size[id] == 2 && l[0] == first && l[1] == second) {
                cl_melt(c, id);
                return def;
            }
        }
    }
    freeze_binaries(c, p);
    return 0;
}

/* save_BN_gate, equivalence.hpp:163-183 */
static void save_bn_gate(sgo_ctx *c, uint32_t p, int pOrgs, int nOrgs)
{
    const uint32_t n = p | 1u;
    const uint32_t save = (pOrgs > nOrgs) ? n : p;
    const uint32_t *d = ol(c, save);
    for (uint32_t i = 0; i < c->ot_len[save]; ++i) if (cl_original(c, d[i])) save_clause(c, d[i], save);
    save_witness(c, save ^ 1u);
}

/* substitute_single (clause), equivalence.hpp:27-56: returns the unit literal or 0 */
static uint32_t subst_clause(sgo_ctx *c, uint32_t dx, uint32_t id, uint32_t def)
{
    uint32_t *l = cl_lits(c, id);
    const int n = c->c_size[id];
    int j = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t lit = l[k];
        if (lit == dx) l[j++] = def;
        else if (lit != def) l[j++] = lit;
    }
    c->c_size[id] = j;
    if (j == 1) return l[0];
    sort_u32(l, j);
    uint32_t sig = 0;
    for (int k = 0; k < j; ++k) sig |= MAPHASH(l[k]);
    c->c_sig[id] = sig;
    return 0;
}

/* substitute_single (lists), equivalence.hpp:58-101 */
static void subst_single(sgo_ctx *c, uint32_t p, uint32_t n, uint32_t def)
{
    const uint32_t def_f = def ^ 1u;
    for (int side = 0; side < 2 && !c->unsat; ++side) {
        const uint32_t me = side ? p : n;                /* negatives first */
        const uint32_t with = side ? def : def_f, taut = side ? def_f : def;
        const uint32_t *d = ol(c, me);
        for (uint32_t i = 0; i < c->ot_len[me]; ++i) {
            const uint32_t id = d[i];
            if (cl_learnt(c, id) || cl_molten(c, id) || cl_has(c, id, taut)) cl_delete(c, id);
            else if (cl_original(c, id)) {
                const uint32_t unit = subst_clause(c, me, id, with);
                if (unit) { new_unit(c, unit); if (c->unsat) return; }
            }
        }
    }
}

/* find_fanin + find_AO_gate, and.hpp:27-111 */
static int find_ao_gate(sgo_ctx *c, uint32_t dx, uint32_t fx, int orgCls, int *ncls, int *nlits)
{
    const uint32_t *dd = ol(c, dx), *df = ol(c, fx);
    const uint32_t nd = c->ot_len[dx], nf = c->ot_len[fx];
    if (c->c_size[dd[0]] > 2 || c->c_size[df[nf - 1]] < 3) return 0;
    uint32_t sig = 0;
    int n = 0;
    if (c->tmp_cap < nd + 2) { c->tmp_cap = 2 * (nd + 2); c->tmp_lits = (uint32_t *)xrealloc(c->tmp_lits, (size_t)c->tmp_cap * 4); }
    uint32_t *out = c->tmp_lits;
    for (uint32_t i = 0; i < nd; ++i) {
        const uint32_t id = dd[i];
        if (cl_original(c, id) && c->c_size[id] == 2) {
            const uint32_t *l = cl_lits(c, id);
            const uint32_t imp = (l[0] ^ l[1] ^ dx) ^ 1u;
            out[n++] = imp;
            sig |= MAPHASH(imp);
            cl_melt(c, id);
        }
    }
    if (n > 1) {
        out[n++] = fx;
        sig |= MAPHASH(fx);
        sort_u32(out, n);
        for (uint32_t i = 0; i < nf; ++i) {
            const uint32_t id = df[i];
            if (cl_original(c, id) && c->c_size[id] == n && sig_sub(c->c_sig[id], sig) &&
                memcmp(cl_lits(c, id), out, (size_t)n * 4) == 0) {
                cl_melt(c, id);
                *ncls = 0, *nlits = 0;
                if (count_bounded(c, dx >> 1, orgCls, dx, fx, PAIR_SUBST, ncls, nlits)) { cl_freeze(c, id); break; }
                return 1;
            }
        }
    }
    freeze_binaries(c, dx);
    return 0;
}

/* fast_equality_check, ifthenelse.hpp:26-46: UINT32_MAX when not found */
static uint32_t fast_equality_check(sgo_ctx *c, uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t t;
    if (c->ot_len[y] > c->ot_len[z]) { t = y; y = z; z = t; }
    if (c->ot_len[x] > c->ot_len[y]) { t = x; x = y; y = t; }
    const uint32_t list = x;
    /* sort3, simplify.hpp:150-156 */
    if (y > z) { t = y; y = z; z = t; }
    if (x > z) { t = x; x = z; z = t; }
    if (x > y) { t = x; x = y; y = t; }
    const uint32_t *d = ol(c, list);
    for (uint32_t i = 0; i < c->ot_len[list]; ++i) {
        const uint32_t id = d[i];
        if (cl_molten(c, id)) continue;
        const uint32_t *l = cl_lits(c, id);
        if (cl_original(c, id) && c->c_size[id] == 3 && l[0] == x && l[1] == y && l[2] == z) return id;
    }
    return 0xFFFFFFFFu;
}

/* find_ITE_gate, ifthenelse.hpp:48-125 */
static int find_ite_gate(sgo_ctx *c, uint32_t dx, uint32_t fx, int orgCls, int *ncls, int *nlits)
{
    const uint32_t *dd = ol(c, dx);
    const uint32_t nd = c->ot_len[dx];
    if (c->c_size[dd[nd - 1]] == 2) return 0;
    for (uint32_t i = 0; i < nd; ++i) {
        const uint32_t ci = dd[i];
        if (!(cl_original(c, ci) && c->c_size[ci] == 3)) continue;
        const uint32_t *li = cl_lits(c, ci);
        uint32_t xi = li[0], yi = li[1], zi = li[2], t;
        if (yi == dx) { t = xi; xi = yi; yi = t; }
        if (zi == dx) { t = xi; xi = zi; zi = t; }
        for (uint32_t j = i + 1; j < nd; ++j) {
            const uint32_t cj = dd[j];
            if (!(cl_original(c, cj) && c->c_size[cj] == 3)) continue;
            const uint32_t *lj = cl_lits(c, cj);
            uint32_t xj = lj[0], yj = lj[1], zj = lj[2];
            if (yj == dx) { t = xj; xj = yj; yj = t; }
            if (zj == dx) { t = xj; xj = zj; zj = t; }
            if ((yi >> 1) == (zj >> 1)) { t = yj; yj = zj; zj = t; }
            if ((zi >> 1) == (zj >> 1)) continue;
            if (yi != (yj ^ 1u)) continue;
            const uint32_t d1 = fast_equality_check(c, fx, yi, zi ^ 1u);
            if (d1 == 0xFFFFFFFFu) continue;
            const uint32_t d2 = fast_equality_check(c, fx, yj, zj ^ 1u);
            if (d2 == 0xFFFFFFFFu) continue;
            cl_melt(c, ci), cl_melt(c, cj), cl_melt(c, d1), cl_melt(c, d2);
            *ncls = 0, *nlits = 0;
            if (count_bounded(c, dx >> 1, orgCls, dx, fx, PAIR_SUBST, ncls, nlits)) {
                cl_freeze(c, ci), cl_freeze(c, cj), cl_freeze(c, d1), cl_freeze(c, d2);
                return 0;
            }
            return 1;
        }
    }
    return 0;
}

static void freeze_arities(sgo_ctx *c, uint32_t me, uint32_t other)     /* xor.hpp:32-44 */
{
    for (int s = 0; s < 2; ++s) {
        const uint32_t lit = s ? other : me;
        const uint32_t *d = ol(c, lit);
        for (uint32_t i = 0; i < c->ot_len[lit]; ++i)
            if (c->c_size[d[i]] > 2 && cl_molten(c, d[i])) cl_freeze(c, d[i]);
    }
}

/* makeArity + checkArity, xor.hpp:53-102 */
static int make_arity(sgo_ctx *c, uint32_t *parity, uint32_t *literals, int size)
{
    const uint32_t oldparity = *parity;
    do { ++*parity; } while (__builtin_parity(*parity));         /* COUNTFLIPS, xor.hpp:26 */
    for (int k = 0; k < size; ++k) {
        const uint32_t bit = 1u << k;
        if ((*parity & bit) != (oldparity & bit)) literals[k] ^= 1u;
    }
    uint32_t best = literals[0];
    uint32_t minsize = c->ot_len[best];
    for (int k = 1; k < size; ++k) {
        const uint32_t lsize = c->ot_len[literals[k]];
        if (lsize < minsize) { minsize = lsize; best = literals[k]; }
    }
    const uint32_t *d = ol(c, best);
    for (uint32_t i = 0; i < c->ot_len[best]; ++i) {
        const uint32_t id = d[i];
        if (cl_original(c, id) && c->c_size[id] == size) {
            const uint32_t *l = cl_lits(c, id);
            int all = 1;
            for (int a = 0; a < size && all; ++a) {
                int found = 0;
                for (int b = 0; b < size; ++b) if (l[a] == literals[b]) { found = 1; break; }
                all = found;
            }
            if (all) { cl_melt(c, id); return 1; }
        }
    }
    return 0;
}

/* find_XOR_gate, xor.hpp:104-181 */
static int find_xor_gate(sgo_ctx *c, uint32_t dx, uint32_t fx, int orgCls, int *ncls, int *nlits)
{
    const uint32_t *dd = ol(c, dx), *df = ol(c, fx);
    const uint32_t nd = c->ot_len[dx], nf = c->ot_len[fx];
    if (c->c_size[dd[nd - 1]] == 2 || c->c_size[df[nf - 1]] == 2) return 0;
    const int maxarity = c->o.xor_max_arity;
    if (c->c_size[dd[0]] - 1 > maxarity) return 0;
    uint32_t out[40];
    for (uint32_t i = 0; i < nd; ++i) {
        const uint32_t ci = dd[i];
        if (!cl_original(c, ci)) continue;
        const int size = c->c_size[ci], arity = size - 1;
        if (size < 3 || arity > maxarity || size > 31) continue;
        memcpy(out, cl_lits(c, ci), (size_t)size * 4);
        uint32_t parity = 0;
        int itargets = 1 << arity;
        while (--itargets && make_arity(c, &parity, out, size)) {}
        if (itargets) freeze_arities(c, dx, fx);
        else {
            cl_melt(c, ci);
            *ncls = 0, *nlits = 0;
            if (count_bounded(c, dx >> 1, orgCls, dx, fx, PAIR_SUBST, ncls, nlits)) { freeze_arities(c, dx, fx); break; }
            return 1;
        }
    }
    return 0;
}

/* ---- function tables, function.hpp:26-261 -------------------------------------------------- */
static const uint64_t magicnumbers[6] = { 0xaaaaaaaaaaaaaaaaull, 0xccccccccccccccccull, 0xf0f0f0f0f0f0f0f0ull,
                                          0xff00ff00ff00ff00ull, 0xffff0000ffff0000ull, 0xffffffff00000000ull };
static void clause2fun(int v, int sign, Fun f)               /* function.hpp:72-93 */
{
    if (v < 6) {
        uint64_t val = magicnumbers[v];
        if (sign) val = ~val;
        for (int i = 0; i < FUNTABLEN; ++i) f[i] |= val;
    } else {
        uint64_t val = sign ? ~0ull : 0ull;
        int j = 0;
        const int sv = 1 << (v - 6);
        for (int i = 0; i < FUNTABLEN; ++i) {
            f[i] |= val;
            if (++j >= sv) { val = ~val; j = 0; }
        }
    }
}
static int isfalsefun(const Fun f) { for (int i = 0; i < FUNTABLEN; ++i) if (f[i]) return 0; return 1; }

/* AND the clauses [0, upto) of ot[lit] (learnt skipped, deleted NOT skipped - the reference only
 * asserts it, function.hpp:137-139) into f. Returns 0 when a variable is outside the core. */
static int and_clauses(sgo_ctx *c, uint32_t lit, uint32_t upto, Fun f)
{
    const uint32_t *d = ol(c, lit);
    for (uint32_t i = 0; i < upto; ++i) {
        const uint32_t id = d[i];
        if (cl_learnt(c, id)) continue;
        Fun cls;
        memset(cls, 0, sizeof cls);
        const uint32_t *l = cl_lits(c, id);
        for (int k = 0; k < c->c_size[id]; ++k) {
            const uint32_t other = l[k];
            if (other == lit) continue;
            const int mvar = c->marks[other >> 1];
            if (mvar < 0) return 0;
            clause2fun(mvar, (int)(other & 1u), cls);
        }
        for (int k = 0; k < FUNTABLEN; ++k) f[k] &= cls[k];
    }
    return 1;
}

/* find_fun_gate, function.hpp:211-259.  Returns 1 (gate / or UNSAT flagged) or 0. */
static int find_fun_gate(sgo_ctx *c, uint32_t p, int orgCls, int *ncls, int *nlits)
{
    if (!c->o.ve_fun_en) return 0;
    const uint32_t n = p | 1u;
    Fun pos, neg, fun;
    memset(pos, 0xff, sizeof pos);
    if (!and_clauses(c, p, c->ot_len[p], pos)) return 0;
    memset(neg, 0xff, sizeof neg);
    if (!and_clauses(c, n, c->ot_len[n], neg)) return 0;
    uint64_t unsat = 0, allzero = 0;
    for (int i = 0; i < FUNTABLEN; ++i) { unsat |= pos[i] | neg[i]; allzero |= pos[i] & neg[i]; }   /* collapsefun */
    if (!unsat) { c->unsat = 1; return 1; }
    if (allzero) return 0;
    int core = 0;
    const uint32_t *dp = ol(c, p), *dn = ol(c, n);
    for (int i = (int)c->ot_len[p] - 1; i >= 0; --i) {
        memcpy(fun, neg, sizeof fun);
        if (cl_original(c, dp[i])) {
            and_clauses(c, p, (uint32_t)i, fun);
            if (isfalsefun(fun)) { cl_melt(c, dp[i]); core = 1; }
        }
    }
    for (int i = (int)c->ot_len[n] - 1; i >= 0; --i) {
        memset(fun, 0xff, sizeof fun);
        if (cl_original(c, dn[i])) {
            and_clauses(c, n, (uint32_t)i, fun);
            if (isfalsefun(fun)) { cl_melt(c, dn[i]); core = 1; }
        }
    }
    *ncls = 0, *nlits = 0;
    if (count_bounded(c, p >> 1, orgCls, p, n, PAIR_CORE, ncls, nlits)) {
        if (core) { freeze_all(c, p); freeze_all(c, n); }
        return 0;
    }
    return 1;
}

/* Solver::resolve / substitute / resolveCore + addResolvent, bounded.cpp:187-307.
 * Resolvents are appended to the store in pair order (outer loop = the SHORTER list, ties keep
 * the positive list outer, bounded.cpp:195-199). */
static void emit(sgo_ctx *c, uint32_t x, int mode, int ncls, int nlits)
{
    uint32_t dx = 2 * x, fx = dx | 1u;
    if (c->ot_len[dx] > c->ot_len[fx]) { uint32_t t = dx; dx = fx; fx = t; }
    ensure_store(c, (uint32_t)ncls + 1, (uint64_t)nlits + 1);
    const uint32_t *dm = ol(c, dx), *dt = ol(c, fx);
    for (uint32_t i = 0; i < c->ot_len[dx] && !c->unsat; ++i) {
        const uint32_t ci = dm[i];
        if (!cl_original(c, ci)) continue;
        const int a = cl_molten(c, ci);
        for (uint32_t j = 0; j < c->ot_len[fx]; ++j) {
            const uint32_t cj = dt[j];
            if (!cl_original(c, cj)) continue;
            const int b = cl_molten(c, cj);
            if (!pair_ok(mode, a, b)) continue;
            const uint32_t need = (uint32_t)(c->c_size[ci] + c->c_size[cj]);
            if (c->tmp_cap < need) { c->tmp_cap = 2 * need; c->tmp_lits = (uint32_t *)xrealloc(c->tmp_lits, (size_t)c->tmp_cap * 4); }
            const int rs = merge_emit(x, cl_lits(c, ci), c->c_size[ci], cl_lits(c, cj), c->c_size[cj], c->tmp_lits);
            if (rs < 0) continue;
            if (rs == 1) { new_unit(c, c->tmp_lits[0]); if (c->unsat) return; continue; }   /* bounded.cpp:283-293; see DESIGN.md "unit resolvent" */
            if (rs == 0) { c->unsat = 1; return; }
            ensure_store(c, 1, (uint64_t)rs);
            const uint32_t id = c->n_cls++;
            c->c_off[id] = (uint32_t)c->lits_used;
            c->c_size[id] = rs;
            c->c_w0[id] = F_ADDED;                                   /* ORIGINAL, markAdded */
            uint32_t sig = 0;
            for (int k = 0; k < rs; ++k) { c->lits[c->lits_used + k] = c->tmp_lits[k]; sig |= MAPHASH(c->tmp_lits[k]); }
            c->c_sig[id] = sig;
            c->lits_used += rs;
        }
    }
}

/* Solver::VE, bounded.cpp:27-185 */
static void ve_stage(sgo_ctx *c)
{
    if (!c->o.ve_en) return;
    const int firstcall = c->simplify_calls == 1;
    for (uint32_t e = 0; e < c->n_elected && !c->unsat; ++e) {
        const uint32_t v = c->elected[e], p = 2 * v, n = p + 1;
        int pOrgs, nOrgs, ncls = 0, nlits = 0, done = 0;
        if (firstcall) { pOrgs = (int)c->ot_len[p]; nOrgs = (int)c->ot_len[n]; }
        else { pOrgs = count_orgs(c, p); nOrgs = count_orgs(c, n); }
        uint32_t def;
        if (!pOrgs || !nOrgs) {
            toblivion(c, p, n, pOrgs, nOrgs);
            c->rep.pures++;
            done = 1;
        } else if ((def = find_bn_gate(c, p, n)) != 0) {
            c->rep.inverters++;
            save_bn_gate(c, p, pOrgs, nOrgs);
            subst_single(c, p, n, def);
            done = 1;
        } else if ((pOrgs == 1 || nOrgs == 1) && count_unbounded(c, v, p, n, &ncls, &nlits)) {
            if (ncls) emit(c, v, PAIR_ALL, ncls, nlits);
            toblivion(c, p, n, pOrgs, nOrgs);
            c->rep.resolutions++;
            done = 1;
        } else {
            ncls = 0, nlits = 0;
            const int orgCls = pOrgs + nOrgs;
            int type = 0;                                  /* 1 RESOLUTION 2 SUBSTITUTION 4 CORE */
            if (orgCls > 2) {
                if (find_ao_gate(c, n, p, orgCls, &ncls, &nlits)) { type = 2; c->rep.andors++; }
                else if (!ncls && find_ao_gate(c, p, n, orgCls, &ncls, &nlits)) { type = 2; c->rep.andors++; }
            }
            if (!type && orgCls > 3) {
                if (find_ite_gate(c, p, n, orgCls, &ncls, &nlits)) { type = 2; c->rep.ites++; }
                else if (!ncls && find_ite_gate(c, n, p, orgCls, &ncls, &nlits)) { type = 2; c->rep.ites++; }
                else if (find_xor_gate(c, p, n, orgCls, &ncls, &nlits)) { type = 2; c->rep.xors++; }
                else if (!ncls && find_xor_gate(c, n, p, orgCls, &ncls, &nlits)) { type = 2; c->rep.xors++; }
            }
            if (!type && orgCls > 2 && find_fun_gate(c, p, orgCls, &ncls, &nlits)) {
                if (c->unsat) return;
                type = 4;
                c->rep.funs++;
            } else if (!type && !ncls && !count_bounded(c, v, orgCls, p, n, PAIR_ALL, &ncls, &nlits)) {
                type = 1;
                c->rep.resolutions++;
            }
            if (type) {
                if (ncls) emit(c, v, type == 2 ? PAIR_SUBST : type == 4 ? PAIR_CORE : PAIR_ALL, ncls, nlits);
                toblivion(c, p, n, pOrgs, nOrgs);
                done = 1;
            }
        }
        if (done) {                                        /* markEliminated, solve.hpp:491-498 */
            c->state[v] = MELTED_M;
            c->unassigned--;
            c->elected[e] = 0;
        }
    }
}

/* ---- ERE: eager redundancy elimination (redundancy.cpp:24-123, redundancy.hpp:24-122) ------------------- */

/* merge_ere, redundancy.hpp:26-122: resolvent of c1, c2 on x into out (no tautology), its signature and the literal
 * with the shortest occurrence list; 0 = tautology or longer than maxlen (checked where the reference checks it) */
static int merge_ere(const sgo_ctx *c, uint32_t x, uint32_t id1, uint32_t id2, int maxlen, uint32_t *out, uint32_t *sig, uint32_t *best)
{
    const uint32_t *d1 = cl_lits(c, id1), *e1 = d1 + c->c_size[id1];
    const uint32_t *d2 = cl_lits(c, id2), *e2 = d2 + c->c_size[id2];
    int n = 0;
    uint32_t minsize = 0x7fffffffu;
    *sig = 0; *best = 0;
#define ERE_TAKE(L) do { const uint32_t l_ = (L); const uint32_t ls_ = c->ot_len[l_]; if (ls_ < minsize) { minsize = ls_; *best = l_; } \
                         *sig |= MAPHASH(l_); out[n++] = l_; } while (0)
    while (d1 != e1 && d2 != e2) {
        if (n > maxlen) return 0;                                            /* redundancy.hpp:49-50 */
        const uint32_t lit1 = *d1, lit2 = *d2, v1 = lit1 >> 1, v2 = lit2 >> 1;
        if (v1 == x) d1++;
        else if (v2 == x) d2++;
        else if ((lit1 ^ lit2) == 1u) return 0;
        else if (v1 < v2) { d1++; ERE_TAKE(lit1); }
        else if (v2 < v1) { d2++; ERE_TAKE(lit2); }
        else { d1++; d2++; ERE_TAKE(lit1); }
    }
    while (d1 != e1) { const uint32_t l = *d1++; if ((l >> 1) != x) ERE_TAKE(l); }
    if (n > maxlen) return 0;
    while (d2 != e2) { const uint32_t l = *d2++; if ((l >> 1) != x) ERE_TAKE(l); }
#undef ERE_TAKE
    return n > maxlen ? 0 : n;
}

/* Solver::ERE, redundancy.cpp:24-123 */
static void ere_stage(sgo_ctx *c)
{
    if (!c->o.ere_en) return;
    const int ereextend = c->o.ere_extend, maxoccurs = c->o.ere_max_occurs, maxsize = c->o.ere_clause_max;
    uint32_t *merged = (uint32_t *)xrealloc(NULL, sizeof(uint32_t) * (size_t)(2 * (maxsize > 0 ? maxsize : 0) + 4096));
    size_t merged_cap = (size_t)(2 * (maxsize > 0 ? maxsize : 0) + 4096);
    for (uint32_t e = 0; e < c->n_elected; ++e) {
        const uint32_t x = c->elected[e];
        uint32_t dx = 2 * x, fx = dx | 1u;
        if (c->ot_len[dx] > c->ot_len[fx]) { const uint32_t t = dx; dx = fx; fx = t; }
        const int ds = (int)c->ot_len[dx], fs = (int)c->ot_len[fx];
        if (!(ds && fs && ds <= maxoccurs && fs <= maxoccurs)) continue;
        if (c->c_size[ol(c, dx)[0]] > maxsize || c->c_size[ol(c, fx)[0]] > maxsize) continue;   /* redundancy.cpp:50-51 */
        for (int i = 0; i < ds; ++i) {
            const uint32_t iref = ol(c, dx)[i];
            if (cl_deleted(c, iref)) continue;
            const int ciorg = cl_original(c, iref);
            for (int j = 0; j < fs; ++j) {
                const uint32_t jref = ol(c, fx)[j];
                if (cl_deleted(c, jref)) continue;
                const int cjorg = cl_original(c, jref);
                const size_t need = (size_t)c->c_size[iref] + (size_t)c->c_size[jref] + 4;
                if (need > merged_cap) { merged_cap = 2 * need; merged = (uint32_t *)xrealloc(merged, sizeof(uint32_t) * merged_cap); }
                uint32_t sig, best;
                const int len = merge_ere(c, x, iref, jref, maxsize, merged, &sig, &best);
                if (len < 2) continue;
                const int nm = (int)c->ot_len[best];
                for (int m = 0; m < nm; ++m) {
                    const uint32_t mref = ol(c, best)[m];
                    if (mref == iref || mref == jref) continue;
                    if (cl_deleted(c, mref)) continue;
                    const int csize = c->c_size[mref];
                    if (len <= csize && sig_sub(sig, c->c_sig[mref]) && lits_sub(merged, len, cl_lits(c, mref), csize)) {
                        const int learnt = cl_learnt(c, mref), equal = len == csize;
                        if (equal && (learnt || (ciorg && cjorg))) { cl_delete(c, mref); break; }     /* removeClause, solve.hpp:717 */
                        else if (!equal && ereextend) {
                            if (learnt && ereextend == 1) continue;
                            memcpy(cl_lits(c, mref), merged, sizeof(uint32_t) * (size_t)len);            /* copyLitsFrom + resize: */
                            c->c_size[mref] = len;                                                       /* the signature stays stale */
                            if (learnt) bump_shrunken(c, mref);
                        }
                    }
                }
            }
        }
    }
    free(merged);
}

/* ---- BCE: blocked clause elimination (blocked.cpp:22-57) ---------------------------------------------------- */

/* isTautology, simplify.hpp:159-179 */
static int is_tautology(const sgo_ctx *c, uint32_t x, uint32_t id1, uint32_t id2)
{
    const uint32_t *a = cl_lits(c, id1), *b = cl_lits(c, id2);
    const int n1 = c->c_size[id1], n2 = c->c_size[id2];
    int i = 0, j = 0;
    while (i < n1 && j < n2) {
        const uint32_t l1 = a[i], l2 = b[j], v1 = l1 >> 1, v2 = l2 >> 1;
        if (v1 == x) i++;
        else if (v2 == x) j++;
        else if ((l1 ^ l2) == 1u) return 1;
        else if (v1 < v2) i++;
        else if (v2 < v1) j++;
        else { i++; j++; }
    }
    return 0;
}

/* Solver::BCE, blocked.cpp:22-57 */
static void bce_stage(sgo_ctx *c)
{
    if (!c->o.bce_en) return;
    for (uint32_t e = 0; e < c->n_elected; ++e) {
        const uint32_t v = c->elected[e];
        if (!v) continue;
        const uint32_t p = 2 * v, n = p | 1u;
        if ((int)c->ot_len[p] > c->o.bce_max_occurs || (int)c->ot_len[n] > c->o.bce_max_occurs) continue;
        for (uint32_t i = 0; i < c->ot_len[n]; ++i) {
            const uint32_t neg = ol(c, n)[i];
            if (!cl_original(c, neg)) continue;
            int all = 1;
            for (uint32_t j = 0; j < c->ot_len[p]; ++j) {
                const uint32_t pos = ol(c, p)[j];
                if (cl_original(c, pos) && !is_tautology(c, v, neg, pos)) { all = 0; break; }
            }
            if (all) { save_clause(c, neg, n); cl_delete(c, neg); }                   /* model.saveClause(neg, size, n) */
        }
    }
}

/* countAll solve.hpp:661-671 */
static void count_all(sgo_ctx *c, uint64_t *ncl, uint64_t *nli)
{
    uint64_t a = 0, b = 0;
    for (uint32_t i = 0; i < c->n_cls; ++i) if (!cl_deleted(c, i)) { a++; b += (uint64_t)c->c_size[i]; }
    *ncl = a; *nli = b;
}

#define STOP_AFTER(STAGE) do { if (c->stop_phase == c->phase && c->stop_stage == (STAGE)) { c->stopped = 1; return SIGMA_OK; } } while (0)

/* the replaced region: simplify.cpp:178-210 */
static int run_pass(sgo_ctx *c)
{
    c->phase = c->multiplier = 0;
    c->stopped = 0;
    int64_t litsbefore = (int64_t)c->n_literals, diff = INT64_MAX;
    while (litsbefore) {
        const int times = c->phase + 1;                                      /* resizeCNF, solve.hpp:626-630 */
        if (times > 1 && times != c->o.phases && (times % c->o.collect_freq) == 0) compact(c);
        STOP_AFTER(SIGMA_T_GC);
        create_ot(c);
        STOP_AFTER(SIGMA_T_OT_ORDER);
        if (!prop(c)) return SIGMA_UNSAT;
        STOP_AFTER(SIGMA_T_PROP);
        const int ok = lcve(c);
        if (c->stopped) return SIGMA_OK;
        if (c->phase < 8) c->rep.elected_per_phase[c->phase] = c->n_elected;
        STOP_AFTER(SIGMA_T_ELECT);
        if (!ok) break;
        sort_ot(c);
        STOP_AFTER(SIGMA_T_SOT);
        if (c->phase == c->o.phases || (diff <= c->o.phase_lits_min && c->phase > 2)) {
            ere_stage(c);                                                    /* simplify.cpp:189 */
            break;
        }
        sub_stage(c);
        if (c->unsat) return SIGMA_UNSAT;
        STOP_AFTER(SIGMA_T_SUB);
        ve_stage(c);
        if (c->unsat) return SIGMA_UNSAT;
        STOP_AFTER(SIGMA_T_VE);
        bce_stage(c);                                                        /* simplify.cpp:196 */
        uint64_t ncl, nli;
        count_all(c, &ncl, &nli);
        uint32_t j = 0;                                                      /* filterElected, solve.hpp:635-644 */
        for (uint32_t i = 0; i < c->n_elected; ++i) { const uint32_t x = c->elected[i]; if (x && !c->state[x]) c->elected[j++] = x; }
        c->n_elected = j;
        c->n_clauses = ncl, c->n_literals = nli;
        diff = litsbefore - (int64_t)nli, litsbefore = (int64_t)nli;
        STOP_AFTER(SIGMA_T_COUNT);
        c->phase++;
        c->multiplier++;
        c->multiplier += c->phase == c->o.phases;
    }
    c->rep.phases_run = c->phase;
    if (!prop(c)) return SIGMA_UNSAT;
    clear_map_frozen(c);
    uint32_t melted = 0;                                                     /* countMelted, solve.hpp:645-654 */
    for (uint32_t v = 1; v <= c->V; ++v) if (c->state[v] & MELTED_M) melted++;
    c->rep.max_melted_delta = melted - c->max_melted;
    c->max_melted = melted;
    count_all(c, &c->n_clauses, &c->n_literals);
    compact(c);
    return SIGMA_OK;
}

/* ---- hand-off stages ------------------------------------------------------------------------- */

static int cmp_u32(const void *a, const void *b) { const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : x > y; }

/* Solver::extract (simplify.cpp:118-133) over orgs then learnts (awaken(), simplify.cpp:146-148).  A CLAUSE record
 * (clause.hpp:47-59) is 4 header buckets - byte 0 the flag bits, byte 1 _l, byte 2 _v, byte 3 _used; then _sz, _pos,
 * _lbd - followed by the literals.  SCLAUSE(const CLAUSE&) (sclause.hpp:61-75): status LEARNT keeps lbd & MAX_LBD_M and
 * usage, ORIGINAL zeroes both; calcSig (sclause.hpp:158-162); SORT = ascending literals (sort.hpp:44-52). */
uint64_t sgo_extract(const sigma_cnf *cnf, uint32_t *arena, uint64_t *refs, uint64_t *n_buckets)
{
    uint64_t n = 0, pos = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const uint64_t *list = pass ? cnf->learnts : cnf->orgs;
        const uint64_t cnt = pass ? cnf->n_learnts : cnf->n_orgs;
        for (uint64_t i = 0; i < cnt; ++i) {
            const uint64_t r = list[i];
            if (r < cnf->stencil_size && cnf->stencil[r]) continue;          /* cm.deleted(ref), simplify.cpp:124 */
            const uint32_t f = cnf->cm[r], size = cnf->cm[r + 1], lbd = cnf->cm[r + 3];
            const int learnt = ((f >> 8) & 0xFFu) != 0;
            const uint32_t used = (f >> 24) & 3u;
            uint32_t *dst = arena + pos, sig = 0;
            for (uint32_t k = 0; k < size; ++k) { const uint32_t l = cnf->cm[r + 4 + k]; dst[3 + k] = l; sig |= MAPHASH(l); }
            qsort(dst + 3, size, sizeof(uint32_t), cmp_u32);
            dst[0] = learnt ? (ST_LEARNT | (used << 4) | ((lbd & 0x03FFFFFFu) << 6)) : 0u;
            dst[1] = sig;
            dst[2] = size;
            refs[n++] = pos;
            pos += 3 + size;
        }
    }
    *n_buckets = pos;
    return n;
}

/* Solver::newBeginning with mapped == false (simplify.cpp:249-268): addClause(SCLAUSE&) (sclause.cpp:22-62) per clause
 * in refs order - CLAUSE(size) constructor (clause.hpp:77-92: _k = 1, _b = size == 2, _pos = 2), learnt clauses take
 * lbd / usage and lose `keep` when size > 2 and lbd > lbd_tier1 -, orgs / learnts, stats.literals, mark_ssubsume
 * (sclause.hpp:209-215) for added clauses of later calls.  out->subsume (may be NULL) must be zeroed by the caller. */
void sgo_new_beginning(const uint32_t *arena, const uint64_t *refs, uint64_t n_clauses, int lbd_tier1, int later_call,
                       sigma_cnf_out *out)
{
    uint64_t pos = 0, no = 0, nl = 0, lo = 0, ll = 0;
    for (uint64_t i = 0; i < n_clauses; ++i) {
        const uint32_t *src = arena + refs[i];
        const uint32_t w0 = src[0], size = src[2];
        const int learnt = (w0 & ST_MASK) == ST_LEARNT;
        const uint32_t lbd = w0 >> 6, used = (w0 >> 4) & 3u;
        if (later_call && (w0 & F_ADDED) && out->subsume)
            for (uint32_t k = 0; k < size; ++k) out->subsume[src[3 + k] >> 1] = 1;
        uint32_t *dst = out->cm + pos;
        const int keep = !(learnt && size > 2 && (int)lbd > lbd_tier1);
        dst[0] = (keep ? 1u << 4 : 0u) | (size == 2 ? 1u << 5 : 0u) | (learnt ? 1u << 8 : 0u) | (learnt ? used << 24 : 0u);
        dst[1] = size; dst[2] = 2u; dst[3] = learnt ? lbd : 0u;
        memcpy(dst + 4, src + 3, sizeof(uint32_t) * size);
        if (learnt) { out->learnts[nl++] = pos; ll += size; }
        else { out->orgs[no++] = pos; lo += size; }
        out->last_ref = pos;
        pos += 4 + size;
    }
    out->cm_buckets = pos; out->n_orgs = no; out->n_learnts = nl; out->lits_original = lo; out->lits_learnt = ll;
}

/* ---- API ----------------------------------------------------------------------------------- */
int sgo_create(sgo_ctx **out) { *out = (sgo_ctx *)xcalloc(1, sizeof(sgo_ctx)); (*out)->stop_phase = -1; return SIGMA_OK; }

static void free_all(sgo_ctx *c)
{
    free(c->c_off); free(c->c_w0); free(c->c_sig); free(c->c_size); free(c->lits);
    free(c->ot_off); free(c->ot_len); free(c->ot_data); free(c->occ); free(c->scores); free(c->order); free(c->tmp_order);
    free(c->state); free(c->frozen); free(c->ifrozen); free(c->value); free(c->marks); free(c->trail); free(c->stack);
    free(c->elected); free(c->vorg); free(c->model); free(c->tmp_lits);
}
void sgo_destroy(sgo_ctx *c) { if (!c) return; free_all(c); free(c); }

/* AoS -> SoA ingest of the reference's SCNF (sclause.hpp:45-49, simptypes.hpp:48-95) */
int sgo_upload(sgo_ctx *c, const sigma_opts *o, const sigma_formula *in)
{
    const int sp = c->stop_phase, ss = c->stop_stage;
    free_all(c);
    memset(c, 0, sizeof *c);
    c->stop_phase = sp, c->stop_stage = ss;
    c->o = *o;
    const uint32_t V = c->V = in->nvars;
    uint64_t nl = 0;
    for (uint32_t i = 0; i < in->n_clauses; ++i) nl += in->arena[in->refs[i] + 2];
    ensure_store(c, in->n_clauses + 1, nl + 1);
    for (uint32_t i = 0; i < in->n_clauses; ++i) {
        const uint32_t *rec = in->arena + in->refs[i];
        c->c_w0[i] = rec[0];
        c->c_sig[i] = rec[1];
        c->c_size[i] = (int32_t)rec[2];
        c->c_off[i] = (uint32_t)c->lits_used;
        memcpy(c->lits + c->lits_used, rec + 3, (size_t)rec[2] * 4);
        c->lits_used += rec[2];
    }
    c->n_cls = in->n_clauses;
    c->ot_off = (uint32_t *)xcalloc(2 * (size_t)V + 3, 4);
    c->ot_len = (uint32_t *)xcalloc(2 * (size_t)V + 2, 4);
    c->occ = (uint32_t *)xcalloc(2 * (size_t)V + 2, 4);
    c->scores = (uint32_t *)xcalloc((size_t)V + 1, 4);
    c->order = (uint32_t *)xcalloc((size_t)V + 1, 4);
    c->tmp_order = (uint32_t *)xcalloc((size_t)V + 1, 4);
    c->state = (uint8_t *)xcalloc((size_t)V + 1, 1);
    c->frozen = (uint8_t *)xcalloc((size_t)V + 1, 1);
    c->value = (int8_t *)xcalloc(2 * (size_t)V + 2, 1);
    c->marks = (int8_t *)xcalloc((size_t)V + 1, 1);
    c->trail = (uint32_t *)xcalloc((size_t)V + 1, 4);
    c->stack = (uint32_t *)xcalloc((size_t)V + 1, 4);
    c->elected = (uint32_t *)xcalloc((size_t)V + 1, 4);
    c->vorg = (uint32_t *)xcalloc((size_t)V + 1, 4);
    memcpy(c->state, in->state, (size_t)V + 1);
    memcpy(c->value, in->value, 2 * (size_t)V + 2);
    memset(c->marks, -1, (size_t)V + 1);
    memcpy(c->trail, in->trail, (size_t)in->trail_size * 4);
    memcpy(c->vorg, in->vorg, ((size_t)V + 1) * 4);
    if (in->ifrozen) { c->ifrozen = (uint8_t *)xcalloc((size_t)V + 1, 1); memcpy(c->ifrozen, in->ifrozen, (size_t)V + 1); }
    c->trail_size = in->trail_size, c->propagated = in->propagated;
    for (uint64_t i = 0; i < in->model_size; ++i) model_push(c, in->model[i]);
    c->unassigned = in->unassigned, c->max_frozen = in->max_frozen, c->max_melted = in->max_melted;
    c->simplify_calls = in->simplify_calls;
    c->n_clauses = in->n_clauses, c->n_literals = in->n_literals ? in->n_literals : nl;
    c->rep.literals_in = c->n_literals;
    return SIGMA_OK;
}

int sgo_execute(sgo_ctx *c) { c->rep.status = run_pass(c); return c->rep.status; }

int sgo_result_sizes(sgo_ctx *c, sigma_formula *s)
{
    s->nvars = c->V;
    s->n_clauses = c->n_cls;
    s->n_buckets = 3ull * c->n_cls + c->lits_used;
    s->n_literals = c->n_literals;
    s->trail_size = c->trail_size;
    s->model_size = c->model_size;
    return SIGMA_OK;
}

/* SoA -> AoS export: the dense arena shrinkSimp() leaves behind (simplify.cpp:235-247) */
int sgo_download(sgo_ctx *c, sigma_formula *out)
{
    sigma_formula need;
    sgo_result_sizes(c, &need);
    if (out->arena_cap < need.n_buckets || out->refs_cap < need.n_clauses || out->trail_cap < need.trail_size || out->model_cap < need.model_size) {
        out->n_buckets = need.n_buckets, out->n_clauses = need.n_clauses, out->trail_size = need.trail_size, out->model_size = need.model_size;
        return SIGMA_E_NOSPACE;
    }
    uint64_t w = 0;
    for (uint32_t i = 0; i < c->n_cls; ++i) {
        out->refs[i] = w;
        out->arena[w] = c->c_w0[i];
        out->arena[w + 1] = c->c_sig[i];
        out->arena[w + 2] = (uint32_t)c->c_size[i];
        memcpy(out->arena + w + 3, cl_lits(c, i), (size_t)c->c_size[i] * 4);
        w += 3 + (uint64_t)c->c_size[i];
    }
    out->nvars = c->V, out->n_clauses = c->n_cls, out->n_buckets = w, out->n_literals = c->n_literals;
    memcpy(out->state, c->state, (size_t)c->V + 1);
    memcpy(out->value, c->value, 2 * (size_t)c->V + 2);
    memcpy(out->trail, c->trail, (size_t)c->trail_size * 4);
    memcpy(out->model, c->model, (size_t)c->model_size * 4);
    out->trail_size = c->trail_size, out->propagated = c->propagated, out->model_size = c->model_size;
    out->unassigned = c->unassigned, out->max_frozen = c->max_frozen, out->max_melted = c->max_melted;
    out->simplify_calls = c->simplify_calls;
    return SIGMA_OK;
}

int sgo_get_report(sgo_ctx *c, sigma_report *rep)
{
    c->rep.n_clauses = (uint32_t)c->n_clauses;
    c->rep.n_literals = c->n_literals;
    *rep = c->rep;
    return SIGMA_OK;
}

int sgo_set_stop(sgo_ctx *c, int phase, int stage) { c->stop_phase = phase, c->stop_stage = stage; return SIGMA_OK; }

int sgo_snapshot(sgo_ctx *c, const char *name, void *dst, uint64_t cap, uint64_t *nbytes)
{
    const void *src = NULL;
    uint64_t n = 0;
    const uint64_t nl = 2ull * c->V + 2;
    uint32_t scal[8] = { c->n_cls, (uint32_t)c->lits_used, c->n_elected, c->trail_size, c->propagated, (uint32_t)c->model_size, c->unassigned, c->max_frozen };
    if (!strcmp(name, "c_off")) src = c->c_off, n = 4ull * c->n_cls;
    else if (!strcmp(name, "c_size")) src = c->c_size, n = 4ull * c->n_cls;
    else if (!strcmp(name, "c_w0")) src = c->c_w0, n = 4ull * c->n_cls;
    else if (!strcmp(name, "c_sig")) src = c->c_sig, n = 4ull * c->n_cls;
    else if (!strcmp(name, "lits")) src = c->lits, n = 4ull * c->lits_used;
    else if (!strcmp(name, "ot_off")) src = c->ot_off, n = 4ull * (nl + 1);
    else if (!strcmp(name, "ot_len")) src = c->ot_len, n = 4ull * nl;
    else if (!strcmp(name, "ot_data")) src = c->ot_data, n = 4ull * c->ot_off[nl];
    else if (!strcmp(name, "occ")) src = c->occ, n = 4ull * nl;
    else if (!strcmp(name, "order")) src = c->order, n = 4ull * c->V;
    else if (!strcmp(name, "elected")) src = c->elected, n = 4ull * c->n_elected;
    else if (!strcmp(name, "state")) src = c->state, n = (uint64_t)c->V + 1;
    else if (!strcmp(name, "value")) src = c->value, n = nl;
    else if (!strcmp(name, "trail")) src = c->trail, n = 4ull * c->trail_size;
    else if (!strcmp(name, "model")) src = c->model, n = 4ull * c->model_size;
    else if (!strcmp(name, "marks")) src = c->marks, n = (uint64_t)c->V + 1;
    else if (!strcmp(name, "n")) src = scal, n = sizeof scal;
    else return SIGMA_E_ARG;
    *nbytes = n;
    if (n > cap) return SIGMA_E_NOSPACE;
    if (n) memcpy(dst, src, n);
    return SIGMA_OK;
}


/* ---- the DIMACS parser (dimacs.cpp:25-177, dimacs.hpp:57-74), sequential restatement ------------------------------------
 * One walk over the text, statement by statement, exactly as parser() does: eatWS; end marker; comment line; header;
 * otherwise makeClause (literal marks, tautologies, root values of the units enqueued so far, unit / clause / dropped).
 * Output into caller buffers as the reference's solver holds it when parser() returns (include/sigma_b200.h, sigma_dimacs).
 * Returns SIGMA_OK, SIGMA_UNSAT (empty clause), SIGMA_E_ARG (anything the reference stops on with LOGERR), SIGMA_E_NOSPACE. */
static int sgo_is_ws(char ch) { return (ch >= 9 && ch <= 13) || ch == 32; }
int sgo_parse_dimacs(const char *t, uint64_t n, sigma_dimacs *out)
{
    uint64_t i = 0;
    uint32_t nvars = 0, ncls = 0;
    int have_header = 0;
    int8_t *marks = NULL;                 /* l2marker: -1 unmarked, else the sign (dimacs.cpp:133-139) */
    int8_t *value = NULL;
    uint32_t *lits = NULL, *org = NULL;
    uint64_t cap = 0, pos = 0, norgs = 0, nunits = 0;
    int rc = SIGMA_OK;
    out->cm_buckets = out->n_orgs = out->n_units = out->binaries = out->ternaries = out->large = out->literals = 0;
    out->max_clause_size = 0; out->nvars = out->header_clauses = 0;
#define SGO_FAIL(code) do { rc = (code); goto done; } while (0)
    while (i < n) {
        while (i < n && sgo_is_ws(t[i])) ++i;
        if (i >= n || t[i] == 0 || t[i] == '%') break;                       /* dimacs.cpp:54 */
        if (t[i] == 'c') { while (i < n && t[i] && t[i++] != '\n') { } continue; }
        if (t[i] == 'p') {
            if (have_header) SGO_FAIL(SIGMA_E_ARG);
            if (n - i < 5 || memcmp(t + i, "p cnf", 5) != 0) SGO_FAIL(SIGMA_E_ARG);
            i += 5;
            for (int k = 0; k < 2; ++k) {
                while (i < n && sgo_is_ws(t[i])) ++i;
                int sign = 0;
                if (i < n && t[i] == '-') { sign = 1; ++i; } else if (i < n && t[i] == '+') ++i;
                if (i >= n || t[i] < '0' || t[i] > '9') SGO_FAIL(SIGMA_E_ARG);
                uint32_t x = 0;
                while (i < n && t[i] >= '0' && t[i] <= '9') x = x * 10u + (uint32_t)(t[i++] - '0');
                if (sign || x == 0) SGO_FAIL(SIGMA_E_ARG);
                if (k == 0) nvars = x; else ncls = x;
            }
            if (nvars >= 0x7FFFFFFEu) SGO_FAIL(SIGMA_E_ARG);
            have_header = 1;
            marks = (int8_t *)malloc(nvars + 1); value = (int8_t *)malloc(2 * (size_t)nvars + 2);
            memset(marks, -1, nvars + 1); memset(value, -1, 2 * (size_t)nvars + 2);
            if (out->state) memset(out->state, 0, nvars + 1);
            out->nvars = nvars; out->header_clauses = ncls;
            continue;
        }
        if (!have_header) SGO_FAIL(SIGMA_E_ARG);                               /* makeClause: "too many variables" with maxVar == 0 */
        /* makeClause (dimacs.cpp:121-177) */
        uint64_t nc = 0, no = 0;
        int satisfied = 0;
        while (1) {
            while (i < n && sgo_is_ws(t[i])) ++i;
            uint32_t sign = 0;
            if (i < n && t[i] == '-') { sign = 1; ++i; } else if (i < n && t[i] == '+') ++i;
            if (i >= n || t[i] < '0' || t[i] > '9') SGO_FAIL(SIGMA_E_ARG);       /* "expected a digit" */
            uint32_t v = 0;
            while (i < n && t[i] >= '0' && t[i] <= '9') v = v * 10u + (uint32_t)(t[i++] - '0');
            if (!v) break;
            if (v > nvars) SGO_FAIL(SIGMA_E_ARG);                                 /* "too many variables" */
            if (no + 1 > cap) { cap = cap ? 2 * cap : 64; lits = (uint32_t *)realloc(lits, cap * 4); org = (uint32_t *)realloc(org, cap * 4); }
            const uint32_t lit = 2u * v + sign;
            org[no++] = lit;
            if (marks[v] < 0) {
                marks[v] = (int8_t)sign;
                if (value[lit] < 0) lits[nc++] = lit;
                else if (value[lit]) satisfied = 1;
            } else if (marks[v] != (int8_t)sign) satisfied = 1;                   /* tautology */
        }
        for (uint64_t k = 0; k < no; ++k) marks[org[k] >> 1] = -1;
        if (satisfied) continue;
        if (!no) SGO_FAIL(SIGMA_UNSAT);                                           /* parser() returns false */
        if (nc == 1) {
            const uint32_t unit = lits[0];
            if (nunits + 1 > out->trail_cap) SGO_FAIL(SIGMA_E_NOSPACE);
            out->trail[nunits++] = unit;                                          /* enqueueUnit, solve.hpp:373-398 */
            value[unit] = 1; value[unit ^ 1u] = 0;
            if (out->state) out->state[unit >> 1] = 2;                            /* FROZEN_M */
        } else if (norgs + 1 > ncls) SGO_FAIL(SIGMA_E_ARG);                        /* "too many clauses" (also for a clause left empty) */
        else if (nc) {
            if (nc == 2) out->binaries++; else if (nc == 3) out->ternaries++; else out->large++;
            if (nc > out->max_clause_size) out->max_clause_size = (uint32_t)nc;
            if (pos + 4 + nc > out->cm_cap || norgs + 1 > out->orgs_cap) SGO_FAIL(SIGMA_E_NOSPACE);
            uint32_t *dst = out->cm + pos;                                        /* CLAUSE(const Lits_t&), clause.hpp:91-106 */
            dst[0] = (1u << 4) | (nc == 2 ? 1u << 5 : 0u); dst[1] = (uint32_t)nc; dst[2] = 2u; dst[3] = 0u;
            for (uint64_t k = 0; k < nc; ++k) dst[4 + k] = lits[k];
            out->orgs[norgs++] = pos;
            pos += 4 + nc;
            out->literals += nc;
        }
    }
    if (!have_header) rc = SIGMA_E_ARG;
done:
    if (rc == SIGMA_OK) {
        out->cm_buckets = pos; out->n_orgs = norgs; out->n_units = nunits;
        if (out->value && value) memcpy(out->value, value, 2 * (size_t)nvars + 2);
    }
    free(marks); free(value); free(lits); free(org);
    return rc;
#undef SGO_FAIL
}
