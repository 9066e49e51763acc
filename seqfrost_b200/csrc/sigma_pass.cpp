/* sigma_pass.cpp - host driver of the B200 SIGmA pass and its C ABI (include/sigma_b200.h).
 *
 * Replaces the phase loop of Solver::simplifying() (reference src/simplify.cpp:180-210): the
 * loop control below is the reference's (same stop rule, same compaction schedule, same
 * multiplier), every stage is a group of CUDA kernels launched through sigma_backend.h.
 * There is no CPU implementation of any stage in this library.
 */
#include "../../include/sigma_b200.h"
#include "sigma_backend.h"
#include "sigma_dimacs.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

namespace {

template <class T> struct DBuf {
    T *p = nullptr;
    size_t cap = 0;             /* elements */
    bool ensure(size_t n, void *stream = nullptr, size_t keep = 0) {
        if (n <= cap) return true;
        size_t ncap = n + n / 4 + 1024;
        T *q = (T *)be_malloc(ncap * sizeof(T));
        if (!q) return false;
        if (p && keep) be_d2d(stream, q, p, keep * sizeof(T));
        if (p) { if (stream) { char e[8]; be_sync(stream, e, sizeof e); } be_free(p); }
        p = q; cap = ncap;
        return true;
    }
    void release() { if (p) be_free(p); p = nullptr; cap = 0; }
};

struct StageEv { void *a, *b; int stage; };

} // namespace

struct sigma_ctx {
    int device = 0;
    void *stream = nullptr;
    char err[512] = { 0 };
    const volatile unsigned char *interrupt = nullptr;
    int stop_phase = -1, stop_stage = -1;
    bool uploaded = false, executed = false, stopped = false;

    sigma_opts opts;
    /* uploaded input (pristine, device + small host scalars) */
    uint32_t in_nvars = 0, in_ncls = 0, in_trail = 0, in_propagated = 0, in_unassigned = 0, in_max_frozen = 0,
             in_max_melted = 0, in_calls = 1;
    uint64_t in_buckets = 0, in_literals = 0, in_model = 0;
    bool in_has_ifrozen = false;
    DBuf<uint32_t> d_arena_in; DBuf<uint64_t> d_refs_in;
    DBuf<uint32_t> d_cm; DBuf<uint64_t> d_crefs; DBuf<uint8_t> d_stencil;   /* sigma_upload_cnf: the solver's clause database */
    uint64_t extract_us = 0, extract_bytes = 0;
    DBuf<uint8_t> d_state_in; DBuf<int8_t> d_value_in; DBuf<uint32_t> d_trail_in, d_model_in;
    DBuf<uint32_t> d_vorg; DBuf<uint8_t> d_ifrozen;

    /* working store */
    DBuf<sg_u4> hdr, hdr2; DBuf<uint32_t> lits, lits2;
    DBuf<uint32_t> ot_off, ot_len, ot_data, occ, cursor;
    DBuf<int8_t> value, marks; DBuf<uint8_t> state;
    DBuf<uint32_t> trail, model;
    DBuf<uint32_t> scores, order, rank, tmpk, tmpv, sorttmp, scantmp;
    DBuf<uint32_t> wl0, wl1, flags, pos, acting, elected;
    DBuf<uint32_t> ve_type, ve_ncls, ve_nlits, ve_nmodel, ve_aux, ve_cls_off, ve_lit_off, ve_mod_off, scr_len, scr_off, scratch;
    DBuf<SgUnit> ubuf;
    DBuf<uint8_t> ere_touched, ere_flags;
    DBuf<uint32_t> ere_filt, ere_claims; DBuf<uint8_t> ere_size8, ere_maxsz;
    DBuf<uint32_t> ot_off2, ot_len2, ot_data2, otutmp, ot_add, otbits, otb_tmp;
    uint64_t phase_add_lits = 0;     /* literals of the resolvents the latest VE stage added */
    uint32_t last_elected = 0;       /* size of `elected` after the latest election (hsc.n_elected does not survive a pull) */
    uint64_t ot_end = 0;             /* words of ot_data in use (the updated table is not packed) */
    uint32_t prev_max_score = 0; bool prev_max_valid = false;   /* largest score of the previous election of this pass */
    bool ot_end_stale = false;       /* the device-side end (totals[13]) has moved since ot_end was read */
    bool ot_valid = false, ot_incremental = true, ot_verify = false;
    int nbh_fused = 0;               /* 0: separate sortOT / SUB / BVE kernels; 1: fused per variable; 2: fused SUB + BVE only */
    uint32_t ot_first_new = 0;
    DBuf<uint32_t> totals;          /* device slots for scan totals */
    /* prop by frontiers */
    DBuf<uint32_t> pp_fpos, pp_lists, pp_cnt, pp_claim, pp_ctl, pp_tmp; DBuf<unsigned long long> pp_disc;
    size_t pp_nlit = 0;
    uint64_t pp_frontiers = 0, pp_rounds = 0, pp_units = 0;
    DBuf<SgScalars> sc;
    DBuf<uint32_t> c_flags, c_sizes, c_newid, c_newoff;
    DBuf<uint32_t> arena_out; DBuf<uint64_t> refs_out;
    DBuf<uint32_t> cm_out; DBuf<uint64_t> orgs_out, learnts_out; DBuf<uint8_t> subsume_out;   /* sigma_download_cnf */
    /* sigma_parse_dimacs: the clause part of the file, its tokens, and the parsed clause database (all on the device) */
    DBuf<uint8_t> dm_text, dm_state; DBuf<int8_t> dm_value;
    DBuf<uint32_t> dm_fn, dm_tilefn, dm_tstate, dm_cstate, dm_ntok, dm_nterm, dm_tokoff, dm_termoff, dm_tok, dm_term;
    DBuf<uint32_t> dm_utime, dm_utime2, dm_csize, dm_words, dm_kept, dm_isunit, dm_wordoff, dm_keptoff, dm_unitoff, dm_cm, dm_trail;
    DBuf<uint64_t> dm_orgs, dm_stats; DBuf<uint32_t> dm_vorg;
    bool dm_parsed = false; sigma_dimacs dm_out;
    DBuf<uint32_t> wt_flags, wt_pos; DBuf<sg_u4> wt_info, wt_rec;                             /* sigma_watches_* */
    int wt_code = -1; uint64_t wt_records = 0; uint32_t wt_nlit = 0;
    DBuf<uint32_t> d_vmap; bool vmap_on = false; uint32_t vmap_nvars = 0;     /* sigma_set_var_map */
    bool cm_built = false; uint32_t cm_learnts = 0; uint64_t cm_lits_org = 0, cm_lits_learnt = 0, cm_last_ref = 0;
    SgScalars hsc;                  /* host mirror */
    uint32_t *h_totals = nullptr;   /* pinned */
    SgScalars *h_sc = nullptr;      /* pinned */

    uint32_t nvars = 0, nlit = 0, n_cls = 0, pool_used = 0;
    uint64_t model_words = 0;
    /* result */
    uint32_t out_ncls = 0; uint64_t out_buckets = 0, out_literals = 0;
    uint32_t out_max_melted = 0;
    sigma_report rep;
    std::vector<StageEv> evs;
    std::vector<void *> evpool;
    size_t evused = 0;
    uint64_t launches0 = 0, n_syncs = 0;
    int cur_stage = -1; bool sync_trace = false;
    bool push_in_flight = false;    /* the pinned scalar mirror is still being read by a queued copy */
    bool count_fresh = false;       /* hsc.live_cls / live_lits already describe the store */
};

namespace {

const char *const kErr[] = { "ok", "formula is UNSATISFIABLE", "interrupted", "out of memory", "CUDA error",
                             "output buffer too small", "bad argument", "unsupported option combination" };

/* every host <-> device round trip of the pass goes through here (counted: sigma_report.reserved[4]) */
int ctx_sync(sigma_ctx *c)
{
    c->n_syncs++;
    c->push_in_flight = false;
    if (c->sync_trace) fprintf(stderr, "sync %d\n", c->cur_stage);
    return be_sync(c->stream, c->err, sizeof c->err);
}

bool interrupted(sigma_ctx *c) { return c->interrupt && *c->interrupt; }

int fail(sigma_ctx *c, int code, const char *msg)
{
    snprintf(c->err, sizeof c->err, "%s", msg);
    return code;
}

void *get_event(sigma_ctx *c)
{
    if (c->evused == c->evpool.size()) c->evpool.push_back(be_event_create());
    return c->evpool[c->evused++];
}

struct Stage {
    sigma_ctx *c; int stage; void *a; uint64_t l0;
    Stage(sigma_ctx *ctx, int st, uint64_t bytes = 0) : c(ctx), stage(st) {
        c->cur_stage = st;
        a = get_event(c);
        be_event_record(c->stream, a);
        l0 = be_launch_count(c->stream);
        c->rep.stage_bytes[st] += bytes;
    }
    ~Stage() {
        void *b = get_event(c);
        be_event_record(c->stream, b);
        c->evs.push_back({ a, b, stage });
        c->rep.stage_launches[stage] += (uint32_t)(be_launch_count(c->stream) - l0);
        c->cur_stage = -1;
    }
};

SgView make_view(sigma_ctx *c)
{
    SgView v;
    memset(&v, 0, sizeof v);
    { static const int rev = [] { const char *e = getenv("SIGMA_ITEMS_HEAVY_FIRST"); return e ? atoi(e) : 1; }(); v.item_rev = (uint32_t)rev; }
    v.hdr = c->hdr.p; v.lits = c->lits.p;
    v.ot_off = c->ot_off.p; v.ot_len = c->ot_len.p; v.ot_data = c->ot_data.p;
    v.value = c->value.p; v.state = c->state.p; v.marks = c->marks.p;
    v.ifrozen = c->in_has_ifrozen ? c->d_ifrozen.p : nullptr;
    v.vorg = c->d_vorg.p; v.trail = c->trail.p; v.model = c->model.p;
    v.occ = c->occ.p; v.scores = c->scores.p; v.order = c->order.p; v.rank = c->rank.p;
    v.elected = c->elected.p; v.acting = c->acting.p;
    v.ve_type = c->ve_type.p; v.ve_ncls = c->ve_ncls.p; v.ve_nlits = c->ve_nlits.p; v.ve_nmodel = c->ve_nmodel.p;
    v.ve_aux = c->ve_aux.p; v.ve_cls_off = c->ve_cls_off.p; v.ve_lit_off = c->ve_lit_off.p; v.ve_mod_off = c->ve_mod_off.p;
    v.scr_off = c->scr_off.p; v.scratch = c->scratch.p; v.ubuf = c->ubuf.p; v.ubuf_cap = (uint32_t)c->ubuf.cap;
    v.ot_add = c->ot_add.p; v.ot_add_cap = (uint32_t)(c->ot_add.cap / 2);
    v.ere_owner = nullptr; v.ere_done = nullptr;
    v.ere_touched = c->ere_touched.p; v.ere_flags = c->ere_flags.p; v.ere_eidx = c->scores.p;
    v.sc = c->sc.p;
    v.nvars = c->nvars; v.nlit = c->nlit; v.n_cls = c->n_cls; v.pool_used = c->pool_used;
    v.model_base = c->model_words;
    const sigma_opts &o = c->opts;
    SgOpts &d = v.o;
    d.phases = o.phases; d.phase_lits_min = o.phase_lits_min; d.mu_pos = o.mu_pos; d.mu_neg = o.mu_neg;
    d.lcve_max_occurs = o.lcve_max_occurs; d.lcve_clause_max = o.lcve_clause_max; d.lcve_min_vars = o.lcve_min_vars;
    d.sub_max_occurs = o.sub_max_occurs; d.sub_clause_max = o.sub_clause_max; d.ve_clause_max = o.ve_clause_max;
    d.xor_max_arity = o.xor_max_arity; d.bce_max_occurs = o.bce_max_occurs; d.ere_max_occurs = o.ere_max_occurs;
    d.ere_clause_max = o.ere_clause_max; d.collect_freq = o.collect_freq; d.lbd_tier1 = o.lbd_tier1;
    d.sub_en = o.sub_en; d.ve_en = o.ve_en; d.ve_plus_en = o.ve_plus_en; d.ve_fun_en = o.ve_fun_en;
    d.ve_lbound_en = o.ve_lbound_en; d.bce_en = o.bce_en; d.ere_en = o.ere_en; d.ere_extend = o.ere_extend;
    d.all_en = o.all_en; d.proof_en = o.proof_en;
    d.firstcall = c->in_calls == 1;
    return v;
}

/* device scalars -> host mirror (one small D2H + sync) */
int pull_scalars(sigma_ctx *c)
{
    be_copy_words(c->stream, c->h_sc, c->sc.p, sizeof(SgScalars));
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    c->hsc = *c->h_sc;
    return SIGMA_OK;
}
/* host mirror -> device scalars.  No round trip: the copy is queued, and the pinned staging copy stays untouched
 * until the next synchronisation (every pull synchronises before it rewrites the mirror) */
int push_scalars(sigma_ctx *c)
{
    if (c->push_in_flight && ctx_sync(c)) return SIGMA_E_CUDA;
    *c->h_sc = c->hsc;
    be_copy_words(c->stream, c->sc.p, c->h_sc, sizeof(SgScalars));
    c->push_in_flight = true;
    return SIGMA_OK;
}
int pull_totals(sigma_ctx *c, uint32_t n)
{
    be_copy_words(c->stream, c->h_totals, c->totals.p, n * sizeof(uint32_t));
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    return SIGMA_OK;
}
/* scan totals and scalars with one round trip */
int pull_both(sigma_ctx *c, uint32_t n)
{
    be_copy_words(c->stream, c->h_totals, c->totals.p, n * sizeof(uint32_t));
    return pull_scalars(c);
}

#define CK(x) do { int rc_ = (x); if (rc_ != SIGMA_OK) return rc_; } while (0)
#define NOMEM(x) do { if (!(x)) return fail(c, SIGMA_E_NOMEM, "device allocation failed"); } while (0)

/* ---- ingest: uploaded AoS arena -> device store --------------------------------------------- */
int stage_ingest(sigma_ctx *c)
{
    const uint32_t n = c->in_ncls;
    const uint64_t L = c->in_buckets - 3ull * n;
    Stage st(c, SIGMA_T_INGEST, 2 * (4 * L + 12ull * n) + 8ull * n);
    const size_t cap_cls = (size_t)n + n / 2 + (1u << 16);
    const size_t cap_lits = (size_t)L + L / 2 + (1u << 18);
    if (cap_lits >= 0xFFFFFFF0ull) return fail(c, SIGMA_E_UNSUPPORTED, "literal pool exceeds 32-bit offsets");
    NOMEM(c->hdr.ensure(cap_cls)); NOMEM(c->lits.ensure(cap_lits));
    NOMEM(c->c_sizes.ensure(n + 1)); NOMEM(c->c_newoff.ensure(n + 1));
    NOMEM(c->scantmp.ensure(be_scan_tmp_words((uint32_t)cap_cls) + be_scan_tmp_words(c->nlit + 1) + 64));
    be_ingest_fused(c->stream, c->d_arena_in.p, c->d_refs_in.p, n, c->in_buckets, c->hdr.p, c->lits.p, c->totals.p + 8);
    CK(pull_totals(c, 9));
    if (c->h_totals[8]) {                                  /* arena with gaps (or malformed): general path */
        be_ingest_sizes(c->stream, c->d_arena_in.p, c->d_refs_in.p, n, c->in_buckets, c->c_sizes.p, c->totals.p + 8);
        be_scan_u32(c->stream, c->c_sizes.p, c->c_newoff.p, n, c->totals.p + 9, c->scantmp.p);
        CK(pull_totals(c, 10));
        /* the sigma_formula contract (include/sigma_b200.h): every ref and every size word stays inside the arena, and the
         * clauses account for exactly n_buckets - 3 n_clauses literals (an arena with unused words is fine, overlapping or
         * truncated records are not) */
        if (c->h_totals[8]) return fail(c, SIGMA_E_ARG, "malformed arena: a clause ref or size points outside n_buckets");
        if ((uint64_t)c->h_totals[9] > L) return fail(c, SIGMA_E_ARG, "malformed arena: clause sizes exceed n_buckets - 3 n_clauses");
        be_ingest_copy(c->stream, c->d_arena_in.p, c->d_refs_in.p, n, c->c_newoff.p, c->hdr.p, c->lits.p);
        c->in_literals = c->h_totals[9];
        c->pool_used = c->h_totals[9];
        c->n_cls = n;
        return SIGMA_OK;
    }
    c->n_cls = n;
    c->pool_used = (uint32_t)L;
    c->in_literals = L;                                    /* gap-free: exactly the words that are not headers */
    return SIGMA_OK;
}

/* ---- createOT (simplify.cpp:39-61) ------------------------------------------------------------ */
int stage_create_ot(sigma_ctx *c, uint64_t live_lits, uint32_t live_cls, bool track = true)
{
    const uint32_t nlit = c->nlit;
    NOMEM(c->ot_data.ensure((size_t)live_lits + 16));
    SgView v = make_view(c);
    const bool want_bits = track && c->ot_incremental;
    /* default: the two-level partition build (sigma_backend.h).  The one-level path below remains for formulas outside its
     * geometry and as a second implementation the tests compare (SIGMA_OT_BUILD=0) */
    const char *ob = getenv("SIGMA_OT_BUILD");
    const size_t need = (ob && ob[0] == '0') ? 0 : be_ot_build_tmp_words(nlit, c->n_cls, live_lits);
    SgOtBuild plan;
    if (need && c->otb_tmp.ensure(need) && be_ot_build_plan(nlit, c->n_cls, live_lits, c->otb_tmp.p, &plan) == 0) {
        if (want_bits) NOMEM(c->otbits.ensure(be_ot_bits_words(c->n_cls) + 64));
        const uint64_t cw = 4ull * plan.ncnt;
        {
            Stage st(c, SIGMA_T_OT_HIST, 4 * live_lits + 8ull * c->n_cls + cw);
            be_ot_build_count(c->stream, v, plan, want_bits ? c->otbits.p : nullptr);
        }
        {
            Stage st(c, SIGMA_T_OT_SCAN, 2 * cw);
            be_scan_u32(c->stream, plan.cnt, plan.cnt, (uint32_t)plan.ncnt, c->totals.p, plan.scantmp);
        }
        {
            Stage st(c, SIGMA_T_OT_SCATTER, 4 * live_lits + 8ull * c->n_cls + cw + 4 * live_lits);      /* 4-byte records */
            be_ot_build_scatter(c->stream, v, plan);
        }
        {
            Stage st(c, SIGMA_T_OT_ORDER, 4 * live_lits + 4 * live_lits + 8ull * nlit);
            be_ot_build_lists(c->stream, v, plan, c->totals.p);
        }
        c->rep.ot_build_mode = 2;
        if (c->ot_verify && track) {
            /* test hook (SIGMA_OT_VERIFY=1): the one-level build into the spare buffers, compared list by list on the device */
            sigma_ctx &x = *c;
            DBuf<uint32_t> so = x.ot_off, sl = x.ot_len, sd = x.ot_data;
            NOMEM(x.ot_off2.ensure(nlit + 1)); NOMEM(x.ot_len2.ensure(nlit)); NOMEM(x.ot_data2.ensure((size_t)live_lits + 16));
            x.ot_off = x.ot_off2; x.ot_len = x.ot_len2; x.ot_data = x.ot_data2;
            SgView w = make_view(c);
            be_memset(c->stream, x.ot_len.p, 0, nlit * sizeof(uint32_t));
            be_hist(c->stream, w, x.ot_len.p);
            be_scan_u32(c->stream, x.ot_len.p, x.ot_off.p, nlit, c->totals.p + 11, c->scantmp.p);
            be_d2d(c->stream, c->cursor.p, x.ot_off.p, nlit * sizeof(uint32_t));
            be_scatter(c->stream, w, c->cursor.p);
            be_ot_order(c->stream, w);
            x.ot_off2 = x.ot_off; x.ot_len2 = x.ot_len; x.ot_data2 = x.ot_data;
            x.ot_off = so; x.ot_len = sl; x.ot_data = sd;
            w = make_view(c);
            be_ot_compare(c->stream, w, x.ot_off2.p, x.ot_len2.p, x.ot_data2.p, c->totals.p + 10);
            CK(pull_totals(c, 12));
            if (c->h_totals[10] || c->h_totals[11] != c->h_totals[0]) {
                snprintf(c->err, sizeof c->err, "two-level occurrence table differs from the one-level build in %u lists (entries %u vs %u)", c->h_totals[10], c->h_totals[0], c->h_totals[11]);
                return SIGMA_E_CUDA;
            }
        }
        if (want_bits) {
            c->ot_end = live_lits; c->ot_end_stale = false;
            c->h_totals[13] = (uint32_t)live_lits;             /* device-side end of the table (bump allocator of the updates) */
            be_copy_words(c->stream, c->totals.p + 13, c->h_totals + 13, 4);
        }
        return SIGMA_OK;
    }
    c->rep.ot_build_mode = 1;
    {
        Stage st(c, SIGMA_T_OT_HIST, 4 * live_lits + 8ull * c->n_cls + 4ull * nlit);
        be_memset(c->stream, c->ot_len.p, 0, nlit * sizeof(uint32_t));
        be_hist(c->stream, v, c->ot_len.p);
    }
    {
        Stage st(c, SIGMA_T_OT_SCAN, 8ull * nlit);
        be_scan_u32(c->stream, c->ot_len.p, c->ot_off.p, nlit, c->totals.p, c->scantmp.p);
    }
    {
        Stage st(c, SIGMA_T_OT_SCATTER, 8 * live_lits + 8ull * c->n_cls);
        be_d2d(c->stream, c->cursor.p, c->ot_off.p, nlit * sizeof(uint32_t));
        be_scatter(c->stream, v, c->cursor.p);
    }
    {
        Stage st(c, SIGMA_T_OT_ORDER, 8 * live_lits + 8ull * nlit);
        be_ot_order(c->stream, v);
    }
    (void)live_cls;
    if (track && c->ot_incremental) {
        /* remember which clauses the table represents: later phases update it instead of rebuilding */
        NOMEM(c->otbits.ensure(be_ot_bits_words(c->n_cls) + 64));
        Stage st(c, SIGMA_T_OT_UPDATE, 16ull * c->n_cls);
        be_ot_bitmap(c->stream, v, c->otbits.p);
        c->ot_end = live_lits; c->ot_end_stale = false;
        c->h_totals[13] = (uint32_t)live_lits;             /* device-side end of the table (bump allocator of the updates) */
        be_copy_words(c->stream, c->totals.p + 13, c->h_totals + 13, 4);
    }
    return SIGMA_OK;
}

/* ---- createOT for a later phase, as an update of the previous phase's table (sigma_backend.h) ---- */
int stage_update_ot(sigma_ctx *c, uint64_t live_lits)
{
    const uint32_t nlit = c->nlit;
    NOMEM(c->otutmp.ensure(be_ot_update_tmp_words(nlit, c->n_cls)));
    {
        const size_t have = c->otbits.cap;
        const size_t need = be_ot_bits_words(c->n_cls) + 64;
        if (need > have) NOMEM(c->otbits.ensure(need, c->stream, have));
    }
    SgOtUpdate u;
    u.bits = c->otbits.p; u.first_new = c->ot_first_new; u.n_add = c->hsc.n_ot_add; u.n_elected = c->last_elected;
    u.tmp = c->otutmp.p; u.total = c->totals.p + 12; u.bump = c->totals.p + 13;
    {
        Stage st(c, SIGMA_T_OT_UPDATE, 16ull * c->n_cls + 12ull * nlit);
        SgView v = make_view(c);
        if (c->ot_end_stale) {                             /* cannot happen in the phase loop (an election lies between two updates) */
            CK(pull_totals(c, 14));
            c->ot_end = c->h_totals[13]; c->ot_end_stale = false;
        }
        /* Room for the lists that grow (they move to the end of ot_data).  Their exact total is known on the device only;
         * a bound is known here: the moved lists are disjoint parts of the ot_end words in use, plus the entries added.
         * Sizing from the bound saves the round trip; the exact path remains for tables near the 32-bit limit. */
        const uint64_t bound = c->ot_end + c->phase_add_lits + c->hsc.n_ot_add;
        const char *ex = getenv("SIGMA_OT_EXACT_NEED");
        const bool exact = (ex && ex[0] == '1') || c->ot_end + bound >= 0xFFFFFFF0ull;
        if (!exact) u.total = nullptr;                     /* the plan skips the exact count */
        be_ot_update_plan(c->stream, v, u);
        if (exact) {
            CK(pull_totals(c, 13));
            const uint32_t need = c->h_totals[12];
            if (c->ot_end + need >= 0xFFFFFFF0ull) return fail(c, SIGMA_E_UNSUPPORTED, "occurrence table exceeds 32-bit offsets");
            NOMEM(c->ot_data.ensure((size_t)c->ot_end + need + 16, c->stream, (size_t)c->ot_end));
            v = make_view(c);
            be_ot_update(c->stream, v, u);
            c->rep.stage_bytes[SIGMA_T_OT_UPDATE] += 8ull * need;
            c->ot_end += need;                             /* the device-side bump advanced by the same amount */
        } else {
            NOMEM(c->ot_data.ensure((size_t)(c->ot_end + bound) + 16, c->stream, (size_t)c->ot_end));
            v = make_view(c);
            be_ot_update(c->stream, v, u);
            c->ot_end_stale = true;                        /* the new end comes back with the election's totals */
        }
    }
    if (c->ot_verify) {
        /* test hook: rebuild from scratch into the spare buffers and compare list by list */
        sigma_ctx &x = *c;
        DBuf<uint32_t> so = x.ot_off, sl = x.ot_len, sd = x.ot_data;
        NOMEM(x.ot_off2.ensure(nlit + 1)); NOMEM(x.ot_len2.ensure(nlit)); NOMEM(x.ot_data2.ensure((size_t)live_lits + 16));
        x.ot_off = x.ot_off2; x.ot_len = x.ot_len2; x.ot_data = x.ot_data2;
        int rc = stage_create_ot(c, live_lits, 0, false);
        x.ot_off2 = x.ot_off; x.ot_len2 = x.ot_len; x.ot_data2 = x.ot_data;
        x.ot_off = so; x.ot_len = sl; x.ot_data = sd;
        if (rc != SIGMA_OK) return rc;
        SgView w = make_view(c);
        be_ot_compare(c->stream, w, x.ot_off2.p, x.ot_len2.p, x.ot_data2.p, c->totals.p + 10);
        CK(pull_totals(c, 11));
        if (c->h_totals[10]) { snprintf(c->err, sizeof c->err, "incremental occurrence table differs from a rebuild in %u lists", c->h_totals[10]); return SIGMA_E_CUDA; }
    }
    return SIGMA_OK;
}

/* ---- prop (propelim.cpp:46-81) ---------------------------------------------------------------- */
/* side tables of the frontier kernels: allocated on first use, idle state = "no literal marked, nothing discovered" */
static int prop_tables(sigma_ctx *c)
{
    const size_t nlit = c->nlit;
    if (c->pp_nlit < nlit) {
        NOMEM(c->pp_fpos.ensure(nlit, c->stream) && c->pp_disc.ensure(2 * nlit, c->stream) && c->pp_lists.ensure(2 * nlit, c->stream) &&
              c->pp_cnt.ensure(2 * ((size_t)c->nvars + 4), c->stream) && c->pp_ctl.ensure(16, c->stream) &&
              c->pp_tmp.ensure(be_scan_tmp_words(c->nvars + 4), c->stream));
        be_memset(c->stream, c->pp_fpos.p, 0xFF, 4 * nlit);
        be_memset(c->stream, c->pp_disc.p, 0xFF, 16 * nlit);
        c->pp_nlit = nlit;
    }
    NOMEM(c->pp_claim.ensure((size_t)c->n_cls / 32 + 2, c->stream));
    return SIGMA_OK;
}

/* one frontier: trail[propagated, n_trail) all at once */
static int prop_frontier(sigma_ctx *c)
{
    CK(prop_tables(c));
    const size_t nlit = c->nlit;
    const uint32_t n = c->hsc.n_trail - c->hsc.propagated;
    SgView v = make_view(c);
    SgProp pp;
    pp.p0 = c->hsc.propagated; pp.p1 = c->hsc.n_trail;
    pp.fpos = c->pp_fpos.p;
    pp.disc = c->pp_disc.p; pp.disc_new = c->pp_disc.p + nlit;
    pp.touched = c->pp_lists.p; pp.touched_new = c->pp_lists.p + nlit;
    pp.cnt = c->pp_cnt.p; pp.base = c->pp_cnt.p + c->nvars + 4;
    pp.claim = c->pp_claim.p; pp.ctl = c->pp_ctl.p;
    uint32_t *h = c->h_totals + 32;
    /* One evaluation with the empty guess is enough.  A visit ends as UNIT(u) only when u is the one literal left, so the only
     * discovery that can change its outcome is an earlier one of u itself (the minimum already in the report) or of -u, which
     * is a conflict; hence the first report is the fixpoint unless some visit conflicts under it, and the earliest such visit
     * has the sequential past, so the reference conflicts as well.  SIGMA_PROP_CONFIRM=1 (tests) iterates to the fixpoint
     * anyway and checks that claim. */
    static const bool confirm = getenv("SIGMA_PROP_CONFIRM") && getenv("SIGMA_PROP_CONFIRM")[0] == '1';
    uint32_t rounds = 0;
    be_pp_mark(c->stream, v, pp);
    for (;;) {
        be_pp_round(c->stream, v, pp, confirm ? 1 : 0);
        c->pp_rounds++; rounds++;
        if (confirm) {
            be_copy_words(c->stream, h, pp.ctl, 8 * sizeof(uint32_t));
            if (ctx_sync(c)) return SIGMA_E_CUDA;
            if (!h[2]) break;
        }
        be_pp_next(c->stream, pp);
        std::swap(pp.disc, pp.disc_new);
        std::swap(pp.touched, pp.touched_new);
        if (!confirm) break;
    }
    be_pp_count(c->stream, v, pp, c->pp_tmp.p);
    be_copy_words(c->stream, h, pp.ctl, 8 * sizeof(uint32_t));
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    c->pp_frontiers++; c->pp_units += n;
    if (getenv("SIGMA_PROP_TRACE")) fprintf(stderr, "prop frontier %u units at %u: %u new, conflict %u, rounds so far %llu\n", n, pp.p0, h[4], h[3], (unsigned long long)c->pp_rounds);
    if (confirm && rounds > 2 && !h[3]) return fail(c, SIGMA_E_CUDA, "prop: the first report was not the fixpoint of a conflict-free frontier");
    if (h[3]) { be_pp_conflict(c->stream, v, pp); return SIGMA_OK; }        /* learnEmpty (propelim.cpp:62,67): unsat comes back with the scalars */
    be_memset(c->stream, pp.claim, 0, 4 * ((size_t)c->n_cls / 32 + 1));
    be_pp_commit(c->stream, v, pp, h[4]);
    return SIGMA_OK;
}

int stage_prop(sigma_ctx *c)
{
    if (c->hsc.propagated >= c->hsc.n_trail) return SIGMA_OK;
    /* frontiers of at least `wide` units go to the frontier kernels, shorter ones to the one-thread walk, which hands back as
     * soon as that many units wait (SIGMA_PROP_WIDE: 0 = one thread only, 1 = frontiers only) */
    uint32_t wide = 64;
    { const char *e = getenv("SIGMA_PROP_WIDE"); if (e) wide = (uint32_t)strtoul(e, nullptr, 10); }
    if (!be_has_prop_frontier()) wide = 0;
    bool fresh = false;                                    /* the scalars were pulled after the last launch */
    {
        Stage st(c, SIGMA_T_PROP);
        for (;;) {
            if (wide && c->hsc.n_trail - c->hsc.propagated >= wide) CK(prop_frontier(c));
            else be_serial(c->stream, SGS_PROP, make_view(c), wide);
            if (!wide) break;
            CK(pull_scalars(c));
            if (c->hsc.unsat || c->hsc.propagated >= c->hsc.n_trail) { fresh = c->opts.sub_en != 0; break; }
        }
        if (!c->opts.sub_en) be_items(c->stream, SGK_REDUCE, make_view(c), c->nlit);   /* propelim.cpp:79 */
    }
    if (!fresh) CK(pull_scalars(c));
    if (c->hsc.unsat) return SIGMA_UNSAT;
    return SIGMA_OK;
}

/* ---- LCVE (lcve.cpp:29-119) ------------------------------------------------------------------- */
int stage_lcve(sigma_ctx *c, int phase, uint32_t multiplier, bool ot_is_hist, bool *enough)
{
    const uint32_t V = c->nvars, nlit = c->nlit;
    SgView v = make_view(c);
    v.pmax = (uint32_t)c->opts.mu_pos << multiplier;
    v.nmax = (uint32_t)c->opts.mu_neg << multiplier;
    /* election scalars first (the host mirror is current here: nothing since the last pull wrote a scalar), so that the
     * largest score, which the ordering kernels leave in the scalars, survives until the election's pull */
    c->hsc.brk_rank = 0xFFFFFFFFu; c->hsc.wl_count[0] = c->hsc.wl_count[1] = 0;
    c->hsc.n_acting = c->hsc.n_ponly = c->hsc.n_elected = 0;
    c->hsc.n_ot_add = 0;
    CK(push_scalars(c));
    /* radix key width = top bit of the largest score.  From the second election of a pass on it is guessed from the
     * previous phase's largest score (x4) instead of fetched - one round trip less -, and checked when the election's
     * result comes back; a wrong guess (never seen so far) repeats ordering + election with the exact width */
    uint32_t key_bits = 8;
    bool guessed = false;
    bool force_narrow = false;                             /* SIGMA_VO_GUESS=8 (tests): always guess 8 bits, so the repeat path runs */
    { const char *e = getenv("SIGMA_VO_GUESS"); guessed = c->prev_max_valid && !(e && e[0] == '0'); force_narrow = e && e[0] == '8'; }
    if (force_narrow) guessed = true;
    else if (guessed) while (key_bits < 32 && ((4ull * c->prev_max_score + 1023) >> key_bits)) key_bits += 8;
    uint32_t rounds = 0;
again:
    {
        Stage st(c, SIGMA_T_VO, 8ull * V + 4 * 16ull * V);
        if (ot_is_hist) be_d2d(c->stream, c->occ.p, c->ot_len.p, nlit * sizeof(uint32_t));
        else {
            be_memset(c->stream, c->occ.p, 0, nlit * sizeof(uint32_t));
            be_hist(c->stream, v, c->occ.p);                 /* histCNF (histogram.cpp:84-97) */
        }
        /* default: one cooperative launch that takes the key width from the largest score on the device (no guess, no round
         * trip); SIGMA_VO_FUSED=0 or a forced narrow guess (tests) keep the separate launches and their guess / repeat logic */
        bool fused = false;
        { const char *e = getenv("SIGMA_VO_FUSED"); fused = !(e && e[0] == '0'); }
        if (fused && be_vo_sort(c->stream, v, c->tmpk.p, c->tmpv.p, c->sorttmp.p, force_narrow && guessed ? 8u : 0u) == 0) {
            if (!(force_narrow && guessed)) { guessed = false; key_bits = 32; }
        } else {
            be_scores(c->stream, v);
            if (!guessed) {
                CK(pull_scalars(c));
                key_bits = 8;                              /* digit passes that can differ: up to the top bit of the largest score */
                while (key_bits < 32 && (c->hsc.max_score >> key_bits)) key_bits += 8;
            }
            be_sort_pairs(c->stream, c->scores.p + 1, c->order.p, V, c->tmpk.p, c->tmpv.p, c->sorttmp.p, key_bits);
            be_rank(c->stream, v);
        }
    }
    rounds = 0;
    {
        Stage st(c, SIGMA_T_ELECT);
        be_items(c->stream, SGK_ELECT_PREP, v, V);
        /* rank-ordered batches of geometrically growing size: inside a batch, pull rounds until every
         * candidate is decided; acting candidates push FROZEN onto their undecided neighbours, so the
         * later (large) batches hold few candidates that are still undecided */
        long env_b0 = -1, env_grow = -1, env_rps = -1;
        {
            const char *e;
            env_b0 = (e = getenv("SIGMA_ELECT_B0")) ? atol(e) : 4096;
            env_grow = (e = getenv("SIGMA_ELECT_GROW")) ? atol(e) : 4;
            env_rps = (e = getenv("SIGMA_ELECT_RPS")) ? atol(e) : 3;
            if (env_b0 < 32) env_b0 = 32;
            if (env_grow < 2) env_grow = 2;
            if (env_rps < 1) env_rps = 1;
        }
        long env_pers = -1, env_pgrow = -1;
        {
            const char *e;
            env_pers = (e = getenv("SIGMA_ELECT_PERSISTENT")) ? atol(e) : 2;
            env_pgrow = (e = getenv("SIGMA_ELECT_PGROW")) ? atol(e) : 2;
            if (env_pgrow < 2) env_pgrow = 2;
        }
        bool done = false;
        /* default (2): no rounds at all - warps park blocked candidates and poll; 1: batches + rounds inside one cooperative
         * launch; 0: host-driven rounds */
        if (env_pers >= 2 && be_elect_async(c->stream, v) == 0) done = true;
        if (!done && env_pers && be_elect_all(c->stream, v, c->wl0.p, c->wl1.p, (uint32_t)(env_b0 < 16384 ? 16384 : env_b0), (uint32_t)env_pgrow) == 0)
            done = true;                                   /* the round count comes back with the totals below */
        const bool persistent = done;
        uint32_t r0 = 0, bsize = (uint32_t)env_b0;
        while (!done && r0 < V) {
            const uint32_t r1 = (V - r0 <= bsize) ? V : r0 + bsize;
            be_memset(c->stream, &c->sc.p->wl_count[0], 0, 2 * sizeof(uint32_t));
            be_elect_batch(c->stream, v, r0, r1, c->wl0.p);
            int parity = 0;
            while (true) {
                for (int k = 0; k < (int)env_rps; ++k) {
                    be_elect_round(c->stream, v, parity ? c->wl1.p : c->wl0.p, parity ? c->wl0.p : c->wl1.p, parity);
                    parity ^= 1;
                    rounds++;
                }
                CK(pull_scalars(c));
                if (!c->hsc.wl_count[parity]) break;
                if (rounds > 4u * V + 64) return fail(c, SIGMA_E_CUDA, "election did not converge");
            }
            if (c->hsc.brk_rank < r1) break;               /* the sequential loop ended inside this batch */
            r0 = r1;
            if (bsize < (1u << 28)) bsize *= (uint32_t)env_grow;
        }
        /* acting / elected candidates in election order, cut at the break (lcve.cpp:70,73) */
        be_acting_flags(c->stream, v, c->flags.p, c->flags.p + V);
        {
            SgScanSet ss = { { c->flags.p, c->flags.p + V, nullptr, nullptr }, { c->pos.p, c->pos.p + V, nullptr, nullptr },
                             { c->totals.p, c->totals.p + 1, nullptr, nullptr } };
            NOMEM(c->scantmp.ensure(2 * be_scan_tmp_words(V) + 64));
            be_scan_multi(c->stream, ss, 2, V, c->scantmp.p);
        }
        be_acting_scatter(c->stream, v, c->flags.p, c->pos.p, c->flags.p + V, c->pos.p + V);
        CK(pull_both(c, 14));
        if (c->ot_end_stale) {
            c->rep.stage_bytes[SIGMA_T_OT_UPDATE] += 8ull * (c->h_totals[13] - c->ot_end);
            c->ot_end = c->h_totals[13]; c->ot_end_stale = false;
        }
        if (guessed && key_bits < 32 && (c->hsc.max_score >> key_bits)) {
            /* the guess was too narrow: the order is wrong.  Everything above is recomputed from the store */
            guessed = false;
            c->hsc.brk_rank = 0xFFFFFFFFu; c->hsc.wl_count[0] = c->hsc.wl_count[1] = 0;
            c->hsc.n_acting = c->hsc.n_ponly = c->hsc.n_elected = 0;
            CK(push_scalars(c));
            goto again;
        }
        c->prev_max_score = c->hsc.max_score; c->prev_max_valid = true;
        if (persistent) rounds = c->hsc.wl_count[0];
        c->hsc.n_acting = c->h_totals[0]; c->hsc.n_elected = c->h_totals[1];
        c->last_elected = c->h_totals[1];
        be_serial(c->stream, SGS_MARKS, v, c->hsc.n_acting);
    }
    if (phase < 8) { c->rep.elected_per_phase[phase] = c->hsc.n_elected; c->rep.election_rounds[phase] = rounds; }
    *enough = c->hsc.n_elected >= (uint32_t)c->opts.lcve_min_vars;
    return SIGMA_OK;
}

static int prop_tables(sigma_ctx *c);
int apply_units(sigma_ctx *c, SgView &v)
{
    CK(pull_scalars(c));
    if (c->hsc.unsat) return SIGMA_UNSAT;
    const uint32_t n = c->hsc.n_units;
    if (!n) return SIGMA_OK;
    if (n > c->ubuf.cap) return fail(c, SIGMA_E_UNSUPPORTED, "unit buffer overflow");
    be_sort_units(c->stream, v, n);
    /* many units (hundreds of thousands after a SUB stage on a unit storm: 24 ms for the one-thread walk): all positions at once */
    uint32_t wide = 256;
    { const char *e = getenv("SIGMA_UNITS_WIDE"); if (e) wide = (uint32_t)strtoul(e, nullptr, 10); }
    if (wide && n >= wide && be_has_prop_frontier()) {
        CK(prop_tables(c));
        NOMEM(c->pp_cnt.ensure(2 * ((size_t)n + 4), c->stream) && c->pp_tmp.ensure(be_scan_tmp_words(n + 4), c->stream));
        v = make_view(c);
        be_apply_units(c->stream, v, n, c->pp_fpos.p, c->pp_cnt.p, c->pp_cnt.p + n + 4, c->pp_ctl.p + 8, c->pp_tmp.p);
    } else be_serial(c->stream, SGS_APPLY_UNITS, v, n);
    CK(pull_scalars(c));
    c->hsc.n_units = 0;
    CK(push_scalars(c));
    if (c->hsc.unsat) return SIGMA_UNSAT;
    return SIGMA_OK;
}

/* sortOT work split: a warp per list pays off when lists have around a dozen entries (k-SAT); with the two- to six-entry
 * lists of gate circuits, millions of warps would start only to find nothing to do, so everything goes to the
 * thread-per-list kernel there */
uint32_t sortot_warp_min(const sigma_ctx *c)
{
    const char *e = getenv("SIGMA_SORTOT_WARP_MIN");
    if (e) return (uint32_t)atoi(e);
    const uint64_t lits = c->hsc.live_lits ? c->hsc.live_lits : c->in_literals;
    return lits < 7ull * c->nlit ? 24u : 8u;               /* average list shorter than 7 entries: no warp kernel */
}

/* sortOT (simplify.cpp:85-100) */
/* the sortOT launches for lists [v.item_base, v.item_base + n): 9..24 entries one warp each, 25..512 entries collected by the
 * thread-per-list kernel and sorted one warp each (SIGMA_SOT_WIDE=0: by the thread-per-list kernel itself), the rest one
 * thread each */
void launch_sortot(sigma_ctx *c, SgView &v, uint32_t n)
{
    v.sortot_warp_min = sortot_warp_min(c);
    if (v.sortot_warp_min < 24) be_items(c->stream, SGK_SORTOT, v, n);
    bool wide = be_has_sortot_wide() && (uint64_t)n <= c->flags.cap;
    { const char *e = getenv("SIGMA_SOT_WIDE"); if (e && e[0] == '0') wide = false; }
    v.sot_wl = wide ? c->flags.p : nullptr;
    if (wide) be_memset(c->stream, &c->sc.p->wl_count[1], 0, sizeof(uint32_t));
    be_items(c->stream, SGK_SORTOT_LONG, v, n);
    if (wide) be_sortot_wide(c->stream, v);
    v.sot_wl = nullptr;
}

int stage_sortot(sigma_ctx *c)
{
    const uint32_t E = c->last_elected;
    Stage st(c, SIGMA_T_SOT);
    SgView v = make_view(c);
    v.n_elected = E;
    launch_sortot(c, v, 2 * E);
    return SIGMA_OK;
}

int stage_sub_ve(sigma_ctx *c, uint32_t multiplier)
{
    const uint32_t E = c->last_elected;
    NOMEM(c->ve_type.ensure(E + 1)); NOMEM(c->ve_ncls.ensure(E + 1)); NOMEM(c->ve_nlits.ensure(E + 1));
    NOMEM(c->ve_nmodel.ensure(E + 1)); NOMEM(c->ve_aux.ensure(E + 1)); NOMEM(c->ve_cls_off.ensure(E + 1));
    NOMEM(c->ve_lit_off.ensure(E + 1)); NOMEM(c->ve_mod_off.ensure(E + 1)); NOMEM(c->scr_len.ensure(E + 1));
    NOMEM(c->scr_off.ensure(E + 1));
    SgView v = make_view(c);
    v.n_elected = E;
    v.pmax = (uint32_t)c->opts.mu_pos << multiplier;
    v.nmax = (uint32_t)c->opts.mu_neg << multiplier;
    /* scratch slices + unit buffer.  A slice is max(|P|, |N|) + 2 words; elected variables share no clause (lcve.cpp:121-145),
     * so all their lists together hold at most n_cls entries: sized from that bound, no round trip for the exact total
     * (which is checked when it comes back with the BVE totals) */
    v.scr_off = c->scr_len.p;              /* SGK_SCRATCH_LEN writes lengths through scr_off */
    be_items(c->stream, SGK_SCRATCH_LEN, v, E);
    be_scan_u32(c->stream, c->scr_len.p, c->scr_off.p, E, c->totals.p + 4, c->scantmp.p);
    const size_t scr_bound = (size_t)c->n_cls + 2ull * E;
    NOMEM(c->scratch.ensure(scr_bound + 64));
    NOMEM(c->ubuf.ensure(scr_bound + 4096));
    NOMEM(c->ot_add.ensure(2 * scr_bound + 128));
    c->ot_first_new = c->n_cls;
    c->phase_add_lits = 0;
    v = make_view(c);
    v.n_elected = E;
    v.pmax = (uint32_t)c->opts.mu_pos << multiplier;
    v.nmax = (uint32_t)c->opts.mu_neg << multiplier;
    const bool fused = c->nbh_fused && c->opts.ve_en;
    if (fused) {
        /* sortOT + SUB + BVE decision per variable in one kernel (sigma_items.cuh: sg_neighbourhood_item) */
        if (c->nbh_fused == 2) CK(stage_sortot(c));
        Stage st(c, SIGMA_T_NBH);
        v.nbh_sort = c->nbh_fused == 1;
        be_items(c->stream, SGK_NBH, v, E);
    } else {
        /* sortOT, SUB and the BVE decision touch the same clauses three times.  Elected variables share no clause,
         * so the three kernels can run chunk after chunk of `elected` (SIGMA_NBH_CHUNK=n, or -1 for an L2-sized
         * estimate), keeping a chunk's clauses in L2 between them.  Measured slower than one launch each (19.1-20.6 ms
         * against 18.1 ms at 60k-20k variables per chunk: the smaller grids lose more to their tails than the reuse
         * gains), so the default is one piece.  (SUB's units wait in the buffer: nothing before the end of this phase
         * reads an assignment, and the (elected index, sequence) keys put them in front of the BVE units.) */
        long env_chunk = 0;
        { const char *e = getenv("SIGMA_NBH_CHUNK"); if (e) env_chunk = atol(e); }
        uint32_t chunk = E;
        if (env_chunk > 0) chunk = (uint32_t)env_chunk;
        else if (env_chunk < 0) {                          /* ~64 MB of clause sectors per chunk */
            const uint64_t occ = c->nvars ? (c->hsc.live_lits ? c->hsc.live_lits : c->in_literals) / c->nvars + 1 : 1;
            const uint64_t per_var = 96 * occ + 64;
            chunk = (uint32_t)std::min<uint64_t>(E, std::max<uint64_t>(4096, (64ull << 20) / per_var));
        }
        const bool do_sub = c->opts.sub_en || c->opts.ve_plus_en;      /* subsume.cpp:25 */
        for (uint32_t first = 0; first < E; first += chunk) {
            const uint32_t cnt = std::min(chunk, E - first);
            {
                Stage st(c, SIGMA_T_SOT);
                v.item_base = 2 * first;
                launch_sortot(c, v, 2 * cnt);
            }
            v.item_base = first;
            if (interrupted(c)) return SIGMA_INTERRUPTED;  /* subsume.cpp:26, bounded.cpp:31 */
            if (do_sub) {
                Stage st(c, SIGMA_T_SUB);
                be_items(c->stream, SGK_SUB, v, cnt);
            }
            if (c->opts.ve_en) {                           /* bounded.cpp:29 */
                Stage st(c, SIGMA_T_VE);
                be_items(c->stream, SGK_VE_COUNT, v, cnt);
            }
        }
        v.item_base = 0;
    }
    if (c->opts.ve_en) {                                   /* bounded.cpp:29 */
        Stage st(c, SIGMA_T_VE);
        {
            SgScanSet ss = { { c->ve_ncls.p, c->ve_nlits.p, c->ve_nmodel.p, nullptr }, { c->ve_cls_off.p, c->ve_lit_off.p, c->ve_mod_off.p, nullptr },
                             { c->totals.p + 0, c->totals.p + 1, c->totals.p + 2, nullptr } };
            NOMEM(c->scantmp.ensure(3 * be_scan_tmp_words(E) + 64));
            be_scan_multi(c->stream, ss, 3, E, c->scantmp.p);  /* clause ids, pool offsets, model offsets in election order */
        }
        CK(pull_both(c, 5));
        if (c->h_totals[4] > scr_bound) return fail(c, SIGMA_E_CUDA, "internal: elected lists overlap (scratch bound exceeded)");
        if (c->hsc.unsat) return SIGMA_UNSAT;              /* function-table UNSAT, bounded.cpp:141-145 */
        const uint32_t add_cls = c->h_totals[0], add_lits = c->h_totals[1], add_model = c->h_totals[2];
        if ((uint64_t)c->pool_used + add_lits >= 0xFFFFFFF0ull)
            return fail(c, SIGMA_E_UNSUPPORTED, "literal pool exceeds 32-bit offsets");
        NOMEM(c->hdr.ensure((size_t)c->n_cls + add_cls, c->stream, c->n_cls));
        NOMEM(c->lits.ensure((size_t)c->pool_used + add_lits, c->stream, c->pool_used));
        NOMEM(c->model.ensure((size_t)c->model_words + add_model, c->stream, (size_t)c->model_words));
        v = make_view(c);
        v.n_elected = E;
        be_items(c->stream, SGK_VE_EMIT, v, E);
        c->n_cls += add_cls;
        c->pool_used += add_lits;
        c->phase_add_lits = add_lits;
        c->model_words += add_model;
    }
    if (c->opts.bce_en && interrupted(c)) return SIGMA_INTERRUPTED;   /* blocked.cpp:26 */
    if (c->opts.bce_en) {                                  /* blocked.cpp:25 */
        Stage st(c, SIGMA_T_VE);
        NOMEM(c->ere_touched.ensure((size_t)c->n_cls + 16));
        be_memset(c->stream, c->ere_touched.p, 0, c->n_cls);
        v = make_view(c);
        v.n_elected = E;
        be_items(c->stream, SGK_BCE_COUNT, v, E);
        be_scan_u32(c->stream, c->ve_nmodel.p, c->ve_mod_off.p, E, c->totals.p, c->scantmp.p);
        CK(pull_totals(c, 1));
        const uint32_t add_model = c->h_totals[0];
        if (add_model) {
            NOMEM(c->model.ensure((size_t)c->model_words + add_model, c->stream, (size_t)c->model_words));
            v = make_view(c);
            v.n_elected = E;
            be_items(c->stream, SGK_BCE_EMIT, v, E);
            c->model_words += add_model;
        }
    }
    /* end of the phase: the live counts (countAll, solve.hpp:661-671) come back with the round trip that fetches the
     * units of SUB and BVE; applying units assigns values and never changes a clause, so the counts stay valid */
    {
        Stage st(c, SIGMA_T_COUNT, 8ull * c->n_cls);
        v = make_view(c);
        v.n_elected = E;
        be_count_live(c->stream, v);
        CK(apply_units(c, v));
        c->count_fresh = true;
    }
    return SIGMA_OK;
}

/* ERE (redundancy.cpp:26-123): propose in parallel, replay the few recorded pairs in sequential order */
int stage_ere(sigma_ctx *c)
{
    if (!c->opts.ere_en) return SIGMA_OK;
    if (interrupted(c)) return SIGMA_INTERRUPTED;          /* redundancy.cpp:29 */
    if (c->opts.ere_clause_max > 256) return fail(c, SIGMA_E_UNSUPPORTED, "--redundancyclausemax above 256 is not supported");
    if (c->opts.ere_max_occurs > 65535) return fail(c, SIGMA_E_UNSUPPORTED, "--redundancymaxoccurs above 65535 is not supported");
    const uint32_t E = c->last_elected;
    if (!E) return SIGMA_OK;
    NOMEM(c->ere_touched.ensure((size_t)c->n_cls + 16)); NOMEM(c->ere_flags.ensure((size_t)E + 16));
    NOMEM(c->ubuf.ensure((size_t)(1u << 20) + 4ull * E));
    Stage st(c, SIGMA_T_ERE);
    be_memset(c->stream, c->ere_touched.p, 0, c->n_cls);
    be_memset(c->stream, c->ere_flags.p, 0, E);
    c->hsc.n_units = 0;
    CK(push_scalars(c));
    NOMEM(c->ve_ncls.ensure(E + 1)); NOMEM(c->ve_cls_off.ensure(E + 1));
    SgView v = make_view(c);
    v.n_elected = E;
    be_items(c->stream, SGK_ERE_CHUNKS, v, E);
    be_scan_u32(c->stream, c->ve_ncls.p, c->ve_cls_off.p, E, c->totals.p, c->scantmp.p);
    CK(pull_totals(c, 1));
    const uint32_t n_items = c->h_totals[0];
    if (n_items) {
        int use_filt = -1;
        if (use_filt < 0) { const char *e = getenv("SIGMA_ERE_FILT"); use_filt = e ? atoi(e) : 1; }
        if (use_filt) {
            NOMEM(c->ere_filt.ensure(2 * (size_t)c->n_cls + 16)); NOMEM(c->ere_size8.ensure((size_t)c->n_cls + 16));
            be_ere_filt(c->stream, v, c->ere_filt.p, c->ere_size8.p);
            v.ere_filt = c->ere_filt.p; v.ere_size8 = c->ere_size8.p;
            if (use_filt != 2) {                           /* SIGMA_ERE_FILT=2 (tests): without the per-list bound */
                NOMEM(c->ere_maxsz.ensure((size_t)c->nlit + 16));
                be_ere_maxsz(c->stream, v, c->ere_size8.p, c->ere_maxsz.p);
                v.ere_maxsz = c->ere_maxsz.p;
            }
        }
        be_items(c->stream, SGK_ERE_PROPOSE, v, n_items);
        v.ere_filt = nullptr; v.ere_size8 = nullptr; v.ere_maxsz = nullptr;
    }
    CK(pull_scalars(c));
    const uint32_t n = c->hsc.n_units;
    if (n > c->ubuf.cap) return fail(c, SIGMA_E_UNSUPPORTED, "ERE event buffer overflow");
    c->hsc.ere_events += n;
    if (n) {
        be_sort_units(c->stream, v, n);
        c->hsc.ere_first_dirty = 0xFFFFFFFFu; c->hsc.ere_applied = 0; c->hsc.ere_par_events = 0; c->hsc.ere_list_start = 0;
        CK(push_scalars(c));
        be_items(c->stream, SGK_ERE_DIRTY, v, E);
        /* ordered list of the positions to visit: flags -> 0/1 -> exclusive scan -> scatter into v.acting */
        be_memset(c->stream, c->ve_ncls.p, 0, (size_t)E * 4);
        be_items(c->stream, SGK_ERE_FLAG01, v, E);
        be_scan_u32(c->stream, c->ve_ncls.p, c->ve_cls_off.p, E, c->totals.p, c->scantmp.p);
        CK(pull_totals(c, 1));
        const uint32_t n_list = c->h_totals[0];
        be_items(c->stream, SGK_ERE_LIST, v, E);
        /* default: the whole replay by one block - parallel claim / apply rounds between consecutive dirty variables
         * (k_ere_replay_block); SIGMA_ERE_BLOCK=0: prefix rounds driven from the host + the one-warp replay */
        bool by_block = false;
        { const char *e = getenv("SIGMA_ERE_BLOCK"); by_block = !(e && e[0] == '0') && 2ull * n <= c->flags.cap; }
        if (by_block && c->c_newid.ensure((size_t)c->n_cls + 16) && c->ere_claims.ensure(9 * (size_t)n + 64)) {
            v.ere_owner = c->c_newid.p; v.ere_done = c->flags.p;
            be_memset(c->stream, v.ere_done, 0, (size_t)n * 4);
            be_memset(c->stream, v.ere_owner, 0xFF, (size_t)c->n_cls * 4);
            /* (no scalar push here: the device copy holds what the kernels above left; the kernel clears its own queue words) */
            if (be_ere_replay_block(c->stream, v, n, n_list, c->ere_claims.p, c->ere_claims.p + 8 * (size_t)n) != 0) by_block = false;
        } else by_block = false;
        if (!by_block) {
        v.ere_owner = nullptr; v.ere_done = nullptr;
        /* recorded pairs before the first dirty variable: parallel claim/apply rounds */
        be_serial(c->stream, SGS_ERE_SPLIT, v, n, n_list);
        CK(pull_scalars(c));
        const uint32_t P = c->hsc.ere_par_events;
        if (P && 2ull * n <= c->flags.cap) {
            NOMEM(c->c_newid.ensure((size_t)c->n_cls + 16));
            v.ere_owner = c->c_newid.p; v.ere_done = c->flags.p;
            be_memset(c->stream, v.ere_done, 0, (size_t)n * 4);
            for (int round = 0; round < 256 && c->hsc.ere_applied < P; ++round) {
                be_memset(c->stream, v.ere_owner, 0xFF, (size_t)c->n_cls * 4);
                be_items(c->stream, SGK_ERE_CLAIM, v, P);
                be_items(c->stream, SGK_ERE_SAFE, v, P);
                be_items(c->stream, SGK_ERE_APPLY, v, P);
                CK(pull_scalars(c));
            }
        }
        c->hsc.ere_npending = 0; c->hsc.ere_overflow = 0;
        CK(push_scalars(c));
        be_serial(c->stream, SGS_ERE_REPLAY, v, n, n_list);
        }
    }
    CK(pull_scalars(c));
    c->hsc.n_units = 0;
    CK(push_scalars(c));
    return SIGMA_OK;
}

int stage_count(sigma_ctx *c)
{
    if (c->count_fresh) { c->count_fresh = false; return SIGMA_OK; }
    {
        Stage st(c, SIGMA_T_COUNT, 8ull * c->n_cls);
        SgView v = make_view(c);
        be_count_live(c->stream, v);
    }
    return pull_scalars(c);
}

/* shrinkSimp (simplify.cpp:235-247) between phases: device store -> compacted device store */
int stage_compact(sigma_ctx *c)
{
    const uint32_t n = c->n_cls;
    NOMEM(c->scantmp.ensure(be_live_tmp_words(n) + 64));
    uint32_t ncls, nlits;
    {
        Stage st(c, SIGMA_T_GC);
        SgView v = make_view(c);
        be_live_scan(c->stream, v, c->scantmp.p, c->totals.p);
        CK(pull_totals(c, 2));
        ncls = c->h_totals[0]; nlits = c->h_totals[1];
        const size_t cap_cls = (size_t)ncls + ncls / 2 + (1u << 16);
        const size_t cap_lits = (size_t)nlits + nlits / 2 + (1u << 18);
        NOMEM(c->hdr2.ensure(cap_cls)); NOMEM(c->lits2.ensure(cap_lits));
        be_compact_apply(c->stream, v, c->scantmp.p, c->hdr2.p, c->lits2.p);
        c->rep.stage_bytes[SIGMA_T_GC] += 4ull * c->pool_used + 16ull * n + 4ull * nlits + 16ull * ncls;
    }
    { DBuf<sg_u4> t = c->hdr; c->hdr = c->hdr2; c->hdr2 = t; }
    { DBuf<uint32_t> t = c->lits; c->lits = c->lits2; c->lits2 = t; }
    c->n_cls = ncls;
    c->pool_used = nlits;
    return SIGMA_OK;
}

/* final shrinkSimp fused with the export to the reference's AoS layout */
int stage_export(sigma_ctx *c)
{
    const uint32_t n = c->n_cls;
    NOMEM(c->scantmp.ensure(be_live_tmp_words(n) + 64));
    Stage st(c, SIGMA_T_EXPORT);
    SgView v = make_view(c);
    be_live_scan(c->stream, v, c->scantmp.p, c->totals.p);
    CK(pull_totals(c, 2));
    const uint32_t ncls = c->h_totals[0], nlits = c->h_totals[1];
    NOMEM(c->arena_out.ensure(3ull * ncls + nlits + 16)); NOMEM(c->refs_out.ensure((size_t)ncls + 16));
    be_export_apply(c->stream, v, c->scantmp.p, c->arena_out.p, c->refs_out.p);
    c->out_ncls = ncls; c->out_literals = nlits; c->out_buckets = 3ull * ncls + nlits;
    c->rep.stage_bytes[SIGMA_T_EXPORT] += 4ull * c->pool_used + 16ull * n + 4ull * nlits + 20ull * ncls;
    return SIGMA_OK;
}

bool stop_here(sigma_ctx *c, int phase, int stage)
{
    if (c->stop_phase != phase || c->stop_stage != stage) return false;
    c->stopped = true;
    return true;
}

int run_phases(sigma_ctx *c)
{
    const sigma_opts &o = c->opts;
    if (o.proof_en) return fail(c, SIGMA_E_UNSUPPORTED, "DRAT proof emission from the device pass is not built");
    if (o.collect_freq <= 0) return fail(c, SIGMA_E_ARG, "collect_freq must be positive");
    CK(stage_ingest(c));
    int phase = 0;
    uint32_t multiplier = 0;
    int64_t litsbefore = (int64_t)c->in_literals, diff = INT64_MAX;
    uint64_t live_lits = c->in_literals;
    int64_t prev_live_lits = (int64_t)c->in_literals;
    uint32_t live_cls = c->in_ncls;
    while (litsbefore) {
        if (interrupted(c)) return SIGMA_INTERRUPTED;
        const int times = phase + 1;                       /* resizeCNF, solve.hpp:626-630 */
        /* update the table in place when the last phase changed little of it (literals added + literals
         * removed under 10 % of the live ones) and the store is still mostly live; otherwise rebuild */
        const uint64_t removed = (uint64_t)(prev_live_lits + (int64_t)c->phase_add_lits - (int64_t)live_lits);
        const bool incr = c->ot_incremental && c->ot_valid && c->hsc.n_ot_add <= c->ot_add.cap / 2 &&
                          2ull * live_cls >= c->n_cls && 10ull * (c->phase_add_lits + removed) < live_lits;
        /* shrinkSimp between phases only renumbers the clauses (order, flags and literals are preserved), so it
         * is skipped while the table is being updated in place; it still runs when half of the store is dead */
        const bool compact = !incr && times > 1 && times != o.phases && (times % o.collect_freq) == 0;
        if (compact) CK(stage_compact(c));
        if (incr) CK(stage_update_ot(c, live_lits));
        else CK(stage_create_ot(c, live_lits, live_cls));
        c->ot_valid = false;
        if (stop_here(c, phase, SIGMA_T_OT_ORDER)) return SIGMA_OK;
        const bool had_units = c->hsc.propagated < c->hsc.n_trail;
        CK(stage_prop(c));
        bool enough = false;
        CK(stage_lcve(c, phase, multiplier, !had_units, &enough));
        if (stop_here(c, phase, SIGMA_T_ELECT)) return SIGMA_OK;
        if (!enough) break;
        if (phase == o.phases || (diff <= o.phase_lits_min && phase > 2)) {
            CK(stage_sortot(c));                            /* simplify.cpp:187 */
            CK(stage_ere(c));                               /* simplify.cpp:189 */
            break;
        }
        CK(stage_sub_ve(c, multiplier));
        if (stop_here(c, phase, SIGMA_T_VE)) return SIGMA_OK;
        CK(stage_count(c));
        c->ot_valid = true;                                /* a complete phase: the table can be updated in place of a rebuild */
        prev_live_lits = (int64_t)live_lits;
        live_cls = c->hsc.live_cls; live_lits = c->hsc.live_lits;
        diff = litsbefore - (int64_t)live_lits; litsbefore = (int64_t)live_lits;
        phase++;
        multiplier++;
        multiplier += phase == o.phases;
    }
    c->rep.phases_run = phase;
    CK(stage_prop(c));                                     /* simplify.cpp:204 */
    if (interrupted(c)) return SIGMA_INTERRUPTED;
    /* countFinal (solve.hpp:655-660) */
    {
        Stage st(c, SIGMA_T_COUNT);
        SgView v = make_view(c);
        be_count_melted(c->stream, v, c->totals.p + 3);
    }
    CK(stage_export(c));
    CK(pull_both(c, 4));
    c->out_max_melted = c->h_totals[3];
    return SIGMA_OK;
}

} // namespace

/* =============================================================================================== */
extern "C" {

int sigma_create(int device, sigma_ctx **out)
{
    if (!out) return SIGMA_E_ARG;
    sigma_ctx *c = new (std::nothrow) sigma_ctx();
    if (!c) return SIGMA_E_NOMEM;
    c->device = device;
    memset(&c->rep, 0, sizeof c->rep);
    memset(&c->hsc, 0, sizeof c->hsc);
    sigma_default_opts(&c->opts);
    if (be_init(device, &c->stream, c->err, sizeof c->err)) { *out = c; return SIGMA_E_CUDA; }
    c->h_totals = (uint32_t *)be_host_alloc(64 * sizeof(uint32_t));
    c->h_sc = (SgScalars *)be_host_alloc(sizeof(SgScalars));
    if (!c->h_totals || !c->h_sc || !c->totals.ensure(64) || !c->sc.ensure(1)) { *out = c; return fail(c, SIGMA_E_NOMEM, "allocation failed"); }
    *out = c;
    return SIGMA_OK;
}

void sigma_destroy(sigma_ctx *c)
{
    if (!c) return;
    if (c->stream) {
        be_bind(c->device);
        char e[64]; be_sync(c->stream, e, sizeof e);
        c->d_arena_in.release(); c->d_refs_in.release(); c->d_cm.release(); c->d_crefs.release(); c->d_stencil.release(); c->d_state_in.release(); c->d_value_in.release();
        c->d_trail_in.release(); c->d_model_in.release(); c->d_vorg.release(); c->d_ifrozen.release();
        c->hdr.release(); c->hdr2.release(); c->lits.release(); c->lits2.release();
        c->ot_off.release(); c->ot_len.release(); c->ot_data.release(); c->occ.release(); c->cursor.release();
        c->value.release(); c->marks.release(); c->state.release(); c->trail.release(); c->model.release();
        c->scores.release(); c->order.release(); c->rank.release(); c->tmpk.release(); c->tmpv.release();
        c->sorttmp.release(); c->scantmp.release(); c->wl0.release(); c->wl1.release(); c->flags.release(); c->pos.release();
        c->acting.release(); c->elected.release(); c->ve_type.release(); c->ve_ncls.release(); c->ve_nlits.release();
        c->ve_nmodel.release(); c->ve_aux.release(); c->ve_cls_off.release(); c->ve_lit_off.release(); c->ve_mod_off.release();
        c->scr_len.release(); c->scr_off.release(); c->scratch.release(); c->ubuf.release(); c->ere_touched.release(); c->ere_flags.release(); c->ere_filt.release(); c->ere_size8.release(); c->ere_maxsz.release();
        c->ot_off2.release(); c->ot_len2.release(); c->ot_data2.release(); c->otutmp.release(); c->ot_add.release(); c->otbits.release(); c->otb_tmp.release(); c->totals.release(); c->sc.release();
        c->pp_fpos.release(); c->pp_lists.release(); c->pp_cnt.release(); c->pp_claim.release(); c->pp_ctl.release(); c->pp_tmp.release(); c->pp_disc.release();
        c->c_flags.release(); c->c_sizes.release(); c->c_newid.release(); c->c_newoff.release();
        c->arena_out.release(); c->refs_out.release(); c->cm_out.release(); c->orgs_out.release(); c->learnts_out.release(); c->subsume_out.release();
        for (void *e2 : c->evpool) be_event_destroy(e2);
        if (c->h_totals) be_host_free(c->h_totals);
        if (c->h_sc) be_host_free(c->h_sc);
        be_shutdown(c->stream);
    }
    delete c;
}

void sigma_default_opts(sigma_opts *o)
{   /* options.cpp:24-51,146 */
    memset(o, 0, sizeof *o);
    o->phases = 5; o->phase_lits_min = 500; o->mu_pos = 32; o->mu_neg = 32;
    o->lcve_max_occurs = 3000; o->lcve_clause_max = 30000; o->lcve_min_vars = 1;
    o->sub_max_occurs = 1000; o->sub_clause_max = 100; o->ve_clause_max = 100; o->xor_max_arity = 10;
    o->bce_max_occurs = 1000; o->ere_max_occurs = 1000; o->ere_clause_max = 200; o->collect_freq = 2; o->lbd_tier1 = 2;
    o->sub_en = 1; o->ve_en = 1; o->ve_plus_en = 1; o->ve_fun_en = 1; o->ve_lbound_en = 0;
    o->bce_en = 0; o->ere_en = 1; o->ere_extend = 1; o->all_en = 0; o->proof_en = 0;
}

const char *sigma_strerror(int code) { return (code >= 0 && code <= 7) ? kErr[code] : "unknown"; }
const char *sigma_last_error(sigma_ctx *c) { return c ? c->err : "null context"; }
void sigma_set_interrupt_flag(sigma_ctx *c, const volatile unsigned char *flag) { if (c) c->interrupt = flag; }
int sigma_host_register(void *p, uint64_t bytes) { return (p && bytes && be_host_register(p, (size_t)bytes) == 0) ? SIGMA_OK : SIGMA_E_CUDA; }
int sigma_host_unregister(void *p) { return (p && be_host_unregister(p) == 0) ? SIGMA_OK : SIGMA_E_CUDA; }
void *sigma_host_alloc(uint64_t bytes) { return be_host_alloc((size_t)bytes); }
void sigma_host_free(void *p) { be_host_free(p); }

int sigma_set_stop(sigma_ctx *c, int phase, int stage)
{
    if (!c) return SIGMA_E_ARG;
    c->stop_phase = phase; c->stop_stage = stage;
    return SIGMA_OK;
}

} /* extern "C" */

namespace {

/* sigma_upload (cnf == nullptr: the SCNF arena comes from the host) and sigma_upload_cnf (the solver's clause
 * database comes from the host, extract() runs here) */
int upload_impl(sigma_ctx *c, const sigma_opts *o, const sigma_formula *in, const sigma_cnf *cnf, bool from_device = false)
{
    /* from_device (sigma_upload_parsed): every pointer of `cnf` and `in` is device memory of this context */
    auto be_in = [from_device](void *st, void *dst, const void *src, size_t bytes) { return from_device ? be_d2d(st, dst, src, bytes) : be_h2d(st, dst, src, bytes); };
    if (!c || !in) return SIGMA_E_ARG;
    if (!c->stream) return fail(c, SIGMA_E_CUDA, c->err[0] ? c->err : "no CUDA device");
    be_bind(c->device);                                    /* callers may use one host thread per context */
    if (o) c->opts = *o;
    if (!in->state || !in->value || !in->vorg) return fail(c, SIGMA_E_ARG, "null input array");
    uint64_t n_src = 0;
    if (cnf) {
        n_src = cnf->n_orgs + cnf->n_learnts;
        if (!cnf->cm || (cnf->n_orgs && !cnf->orgs) || (cnf->n_learnts && !cnf->learnts) || (cnf->stencil_size && !cnf->stencil))
            return fail(c, SIGMA_E_ARG, "null clause database array");
        if (cnf->cm_buckets >= 0xFFFFFFF0ull || n_src >= 0xFFFFFFF0ull) return fail(c, SIGMA_E_UNSUPPORTED, "clause database exceeds 32-bit bucket offsets");
    } else {
        if (!in->arena || !in->refs) return fail(c, SIGMA_E_ARG, "null input array");
        if (in->n_buckets >= 0xFFFFFFF0ull) return fail(c, SIGMA_E_UNSUPPORTED, "arena exceeds 32-bit bucket offsets");
        if (in->n_buckets < 3ull * in->n_clauses) return fail(c, SIGMA_E_ARG, "n_buckets smaller than the clause headers");
    }
    if (in->trail_size > in->nvars) return fail(c, SIGMA_E_ARG, "trail_size exceeds nvars");
    if (in->propagated > in->trail_size) return fail(c, SIGMA_E_ARG, "propagated exceeds trail_size");
    if (in->trail_size && !in->trail) return fail(c, SIGMA_E_ARG, "null trail");
    if (in->model_size && !in->model) return fail(c, SIGMA_E_ARG, "null model");
    if (in->nvars >= 0x1FFFFFF0u) return fail(c, SIGMA_E_UNSUPPORTED, "more than 2^29 variables");   /* election word: rank | status << 29 */
    if (!cnf && in->n_literals > in->n_buckets - 3ull * in->n_clauses)
        return fail(c, SIGMA_E_ARG, "n_literals exceeds n_buckets - 3 n_clauses");    /* the count used is derived from the arena itself */
    c->uploaded = c->executed = false;
    c->evused = 0; c->evs.clear();
    memset(&c->rep, 0, sizeof c->rep);
    c->extract_us = 0; c->extract_bytes = 0;
    const uint32_t V = in->nvars, nlit = 2 * V + 2, n = cnf ? 0 : in->n_clauses;
    c->in_nvars = V; c->in_ncls = n; c->in_buckets = cnf ? 0 : in->n_buckets; c->in_literals = cnf ? 0 : in->n_literals;
    c->in_trail = in->trail_size; c->in_propagated = in->propagated; c->in_unassigned = in->unassigned;
    c->in_max_frozen = in->max_frozen; c->in_max_melted = in->max_melted; c->in_calls = in->simplify_calls ? in->simplify_calls : 1;
    c->in_model = in->model_size; c->in_has_ifrozen = in->ifrozen != nullptr;
    c->nvars = V; c->nlit = nlit;
    if (cnf) {
        NOMEM(c->d_cm.ensure(cnf->cm_buckets + 16)); NOMEM(c->d_crefs.ensure(n_src + 16)); NOMEM(c->d_stencil.ensure(cnf->stencil_size + 16));
        NOMEM(c->c_sizes.ensure(n_src + 16)); NOMEM(c->c_flags.ensure(n_src + 16));
        NOMEM(c->c_newoff.ensure(n_src + 16)); NOMEM(c->c_newid.ensure(n_src + 16));
        NOMEM(c->scantmp.ensure(2 * be_scan_tmp_words((uint32_t)n_src) + 64));
    } else {
        NOMEM(c->d_arena_in.ensure(in->n_buckets + 16)); NOMEM(c->d_refs_in.ensure((size_t)n + 16));
    }
    NOMEM(c->d_state_in.ensure(V + 1)); NOMEM(c->d_value_in.ensure(nlit)); NOMEM(c->d_trail_in.ensure((size_t)V + 16));
    NOMEM(c->d_model_in.ensure((size_t)in->model_size + 16)); NOMEM(c->d_vorg.ensure(V + 1));
    if (in->ifrozen) NOMEM(c->d_ifrozen.ensure(V + 1));
    /* working arrays sized by the variable count */
    NOMEM(c->ot_off.ensure(nlit + 1)); NOMEM(c->ot_len.ensure(nlit)); NOMEM(c->occ.ensure(nlit)); NOMEM(c->cursor.ensure(nlit));
    NOMEM(c->value.ensure(nlit)); NOMEM(c->marks.ensure(V + 1)); NOMEM(c->state.ensure(V + 1));
    NOMEM(c->trail.ensure((size_t)V + 16));
    NOMEM(c->scores.ensure(V + 2)); NOMEM(c->order.ensure(V + 1)); NOMEM(c->rank.ensure(V + 1));
    NOMEM(c->tmpk.ensure(V + 1)); NOMEM(c->tmpv.ensure(V + 1)); NOMEM(c->sorttmp.ensure(be_sort_tmp_words(V) + 64));
    NOMEM(c->wl0.ensure(V + 1)); NOMEM(c->wl1.ensure(V + 1)); NOMEM(c->flags.ensure(2 * (size_t)V + 2)); NOMEM(c->pos.ensure(2 * (size_t)V + 2));
    NOMEM(c->acting.ensure(V + 1)); NOMEM(c->elected.ensure(V + 1));
    void *e0 = get_event(c), *e1 = get_event(c);
    be_event_record(c->stream, e0);
    if (cnf) {
        be_in(c->stream, c->d_cm.p, cnf->cm, cnf->cm_buckets * 4);
        if (cnf->n_orgs) be_in(c->stream, c->d_crefs.p, cnf->orgs, cnf->n_orgs * 8);
        if (cnf->n_learnts) be_in(c->stream, c->d_crefs.p + cnf->n_orgs, cnf->learnts, cnf->n_learnts * 8);
        if (cnf->stencil_size) be_in(c->stream, c->d_stencil.p, cnf->stencil, cnf->stencil_size);
    } else {
        be_in(c->stream, c->d_arena_in.p, in->arena, in->n_buckets * 4);
        be_in(c->stream, c->d_refs_in.p, in->refs, (size_t)n * 8);
    }
    be_in(c->stream, c->d_state_in.p, in->state, V + 1);
    be_in(c->stream, c->d_value_in.p, in->value, nlit);
    if (in->trail_size) be_in(c->stream, c->d_trail_in.p, in->trail, (size_t)in->trail_size * 4);
    if (in->model_size) be_in(c->stream, c->d_model_in.p, in->model, (size_t)in->model_size * 4);
    be_in(c->stream, c->d_vorg.p, in->vorg, (size_t)(V + 1) * 4);
    if (in->ifrozen) be_in(c->stream, c->d_ifrozen.p, in->ifrozen, V + 1);
    be_event_record(c->stream, e1);
    void *e2 = nullptr;
    if (cnf) {
        /* extract(orgs); extract(learnts) (simplify.cpp:118-133,146-148): sizes -> two scans -> write */
        const uint32_t ns = (uint32_t)n_src;
        be_extract_sizes(c->stream, c->d_cm.p, cnf->cm_buckets, c->d_crefs.p, ns, c->d_stencil.p, cnf->stencil_size, c->c_sizes.p, c->c_flags.p, c->totals.p + 8);
        {
            SgScanSet ss = { { c->c_sizes.p, c->c_flags.p, nullptr, nullptr }, { c->c_newoff.p, c->c_newid.p, nullptr, nullptr },
                             { c->totals.p + 0, c->totals.p + 1, nullptr, nullptr } };
            be_scan_multi(c->stream, ss, 2, ns, c->scantmp.p);
        }
        CK(pull_totals(c, 9));
        if (c->h_totals[8]) return fail(c, SIGMA_E_ARG, "malformed clause database: a ref or a record points outside cm_buckets");
        const uint32_t buckets = ns ? c->h_totals[0] : 0, ncls = ns ? c->h_totals[1] : 0;
        NOMEM(c->d_arena_in.ensure((size_t)buckets + 16)); NOMEM(c->d_refs_in.ensure((size_t)ncls + 16));
        be_extract_write(c->stream, c->d_cm.p, c->d_crefs.p, ns, c->c_sizes.p, c->c_newoff.p, c->c_newid.p, c->d_arena_in.p, c->d_refs_in.p);
        e2 = get_event(c);
        be_event_record(c->stream, e2);
        c->in_ncls = ncls; c->in_buckets = buckets; c->in_literals = (uint64_t)buckets - 3ull * ncls;
        /* read 16-byte header + literals + ref + stencil byte, write 12-byte header + literals + ref */
        c->extract_bytes = 2 * 4ull * c->in_literals + 28ull * ncls + 9ull * ns + 8ull * ncls;
    }
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    c->rep.stage_ms[SIGMA_T_H2D] = be_event_ms(e0, e1);
    if (e2) c->extract_us = (uint64_t)(1000.0f * be_event_ms(e1, e2));
    c->rep.reserved[5] = c->extract_us; c->rep.reserved[6] = c->extract_bytes;
    c->rep.h2d_bytes = from_device ? 0 : (cnf ? cnf->cm_buckets * 4 + n_src * 8 + cnf->stencil_size : in->n_buckets * 4 + (uint64_t)n * 8) +
                       (V + 1) + nlit + (uint64_t)in->trail_size * 4 +
                       (uint64_t)in->model_size * 4 + (uint64_t)(V + 1) * 4 + (in->ifrozen ? V + 1 : 0);
    c->uploaded = true;
    return SIGMA_OK;
}

} // namespace

extern "C" {

int sigma_upload(sigma_ctx *c, const sigma_opts *o, const sigma_formula *in) { return upload_impl(c, o, in, nullptr); }

int sigma_upload_cnf(sigma_ctx *c, const sigma_opts *o, const sigma_cnf *cnf, const sigma_formula *in)
{
    if (!cnf) return SIGMA_E_ARG;
    return upload_impl(c, o, in, cnf);
}

int sigma_upload_parsed(sigma_ctx *c, const sigma_opts *o)
{
    if (!c) return SIGMA_E_ARG;
    if (!c->dm_parsed) return fail(c, SIGMA_E_ARG, "sigma_upload_parsed without a successful sigma_parse_dimacs");
    const sigma_dimacs &r = c->dm_out;
    if (r.n_units) return fail(c, SIGMA_E_UNSUPPORTED, "the file has unit clauses: the solver's BCP() and shrinkTop() come between parser() and simplify() "
                                                        "(solve.cpp:52, simplify.cpp:161-162); download the parse and let the solver run them");
    if (!r.n_orgs) return fail(c, SIGMA_E_ARG, "no clauses");
    be_bind(c->device);
    NOMEM(c->dm_vorg.ensure((size_t)r.nvars + 1));
    be_iota_u32(c->stream, c->dm_vorg.p, r.nvars + 1);
    sigma_cnf cnf;
    memset(&cnf, 0, sizeof cnf);
    cnf.cm = c->dm_cm.p; cnf.cm_buckets = r.cm_buckets; cnf.orgs = c->dm_orgs.p; cnf.n_orgs = r.n_orgs;
    sigma_formula f;
    memset(&f, 0, sizeof f);
    f.nvars = r.nvars; f.state = c->dm_state.p; f.value = c->dm_value.p; f.vorg = c->dm_vorg.p;
    f.unassigned = r.nvars; f.simplify_calls = 1;
    return upload_impl(c, o, &f, &cnf, true);
}

int sigma_extracted_sizes(sigma_ctx *c, uint32_t *n_clauses, uint64_t *n_buckets, uint64_t *n_literals)
{
    if (!c || !c->uploaded) return SIGMA_E_ARG;
    if (n_clauses) *n_clauses = c->in_ncls;
    if (n_buckets) *n_buckets = c->in_buckets;
    if (n_literals) *n_literals = c->in_literals;
    return SIGMA_OK;
}

int sigma_execute(sigma_ctx *c)
{
    if (!c) return SIGMA_E_ARG;
    if (!c->uploaded) return fail(c, SIGMA_E_ARG, "sigma_execute before sigma_upload");
    be_bind(c->device);
    const float h2d_ms = c->rep.stage_ms[SIGMA_T_H2D];
    const uint64_t h2d_bytes = c->rep.h2d_bytes;
    memset(&c->rep, 0, sizeof c->rep);
    c->rep.stage_ms[SIGMA_T_H2D] = h2d_ms; c->rep.h2d_bytes = h2d_bytes;
    c->rep.reserved[5] = c->extract_us; c->rep.reserved[6] = c->extract_bytes;
    c->evused = 0; c->evs.clear();
    c->executed = false; c->cm_built = false; c->stopped = false; c->wt_code = -1; c->vmap_on = false;
    c->ot_valid = false;
    { const char *e = getenv("SIGMA_OT_INCREMENTAL"); c->ot_incremental = !(e && e[0] == '0'); }
    { const char *e = getenv("SIGMA_OT_VERIFY"); c->ot_verify = e && e[0] == '1'; }
    { const char *e = getenv("SIGMA_NBH_FUSED"); c->nbh_fused = e ? atoi(e) : 0; }
    c->launches0 = be_launch_count(c->stream);
    c->n_syncs = 0; c->count_fresh = false; c->push_in_flight = false; c->prev_max_valid = false;
    c->sync_trace = getenv("SIGMA_SYNC_TRACE") != nullptr;
    const uint32_t V = c->nvars, nlit = c->nlit;
    /* restore the pristine solver state */
    NOMEM(c->model.ensure((size_t)c->in_model + (1u << 16)));
    void *e0 = get_event(c), *e1 = get_event(c);
    be_event_record(c->stream, e0);
    be_d2d(c->stream, c->state.p, c->d_state_in.p, V + 1);
    be_d2d(c->stream, c->value.p, c->d_value_in.p, nlit);
    if (c->in_trail) be_d2d(c->stream, c->trail.p, c->d_trail_in.p, (size_t)c->in_trail * 4);
    if (c->in_model) be_d2d(c->stream, c->model.p, c->d_model_in.p, (size_t)c->in_model * 4);
    be_memset(c->stream, c->marks.p, 0xFF, V + 1);
    be_memset(c->stream, c->ot_len.p, 0, (size_t)nlit * 4);
    be_memset(c->stream, c->ot_off.p, 0, (size_t)(nlit + 1) * 4);
    memset(&c->hsc, 0, sizeof c->hsc);
    c->hsc.n_trail = c->in_trail; c->hsc.propagated = c->in_propagated; c->hsc.unassigned = c->in_unassigned;
    c->hsc.max_frozen = c->in_max_frozen; c->hsc.brk_rank = 0xFFFFFFFFu;
    c->model_words = c->in_model;
    int rc = push_scalars(c);
    if (rc == SIGMA_OK) rc = run_phases(c);
    be_event_record(c->stream, e1);
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    /* resolve timers */
    for (const StageEv &e : c->evs) c->rep.stage_ms[e.stage] += be_event_ms(e.a, e.b);
    c->rep.stage_ms[SIGMA_T_TOTAL_DEVICE] = be_event_ms(e0, e1);
    c->rep.kernel_launches = be_launch_count(c->stream) - c->launches0;
    c->rep.reserved[4] = c->n_syncs;
    c->rep.status = rc;
    c->rep.literals_in = c->in_literals;
    c->rep.n_clauses = c->out_ncls; c->rep.n_literals = c->out_literals;
    c->rep.max_melted_delta = c->out_max_melted - c->in_max_melted;
    const SgScalars &s = c->hsc;
    c->rep.pures = (int64_t)s.pures; c->rep.inverters = (int64_t)s.inverters; c->rep.andors = (int64_t)s.andors;
    c->rep.ites = (int64_t)s.ites; c->rep.xors = (int64_t)s.xors; c->rep.funs = (int64_t)s.funs;
    c->rep.resolutions = (int64_t)s.resolutions; c->rep.subsumed = (int64_t)s.subsumed;
    c->rep.strengthened = (int64_t)s.strengthened; c->rep.units = (int64_t)s.units_forced;
    c->rep.sort_ties_gt24 = (int64_t)s.sort_ties_gt24;
    c->rep.reserved[0] = s.ere_events; c->rep.reserved[1] = s.ere_dirty_vars; c->rep.reserved[2] = s.ere_full_pairs; c->rep.reserved[3] = s.ere_event_pairs;
    c->executed = (rc == SIGMA_OK) && !c->stopped;            /* a run cut short by sigma_set_stop has no result to download */
    return rc;
}

int sigma_result_sizes(sigma_ctx *c, sigma_formula *sz)
{
    if (!c || !sz) return SIGMA_E_ARG;
    if (!c->executed) return fail(c, SIGMA_E_ARG, "no result: sigma_execute did not complete");
    sz->nvars = c->nvars; sz->n_clauses = c->out_ncls; sz->n_buckets = c->out_buckets; sz->n_literals = c->out_literals;
    sz->trail_size = c->hsc.n_trail; sz->propagated = c->hsc.propagated; sz->model_size = c->model_words;
    sz->unassigned = c->hsc.unassigned; sz->max_frozen = c->hsc.max_frozen; sz->max_melted = c->out_max_melted;
    sz->simplify_calls = c->in_calls;
    return SIGMA_OK;
}

} /* extern "C" */

namespace {

/* newBeginning() on the device: the exported SCNF -> CLAUSE records + orgs / learnts (built once per execute) */
int build_cm(sigma_ctx *c)
{
    if (c->cm_built) return SIGMA_OK;
    const uint32_t n = c->out_ncls;
    NOMEM(c->cm_out.ensure(c->out_buckets + n + 16)); NOMEM(c->orgs_out.ensure((size_t)n + 16)); NOMEM(c->learnts_out.ensure((size_t)n + 16));
    NOMEM(c->c_flags.ensure((size_t)n + 16)); NOMEM(c->c_newid.ensure((size_t)n + 16)); NOMEM(c->subsume_out.ensure(c->nvars + 16));
    NOMEM(c->scantmp.ensure(be_scan_tmp_words(n) + 64));
    void *e0 = get_event(c), *e1 = get_event(c);
    be_event_record(c->stream, e0);
    be_memset(c->stream, c->subsume_out.p, 0, c->nvars + 1);
    be_cm_flags(c->stream, c->arena_out.p, c->refs_out.p, n, c->c_flags.p);
    be_scan_u32(c->stream, c->c_flags.p, c->c_newid.p, n, c->totals.p + 0, c->scantmp.p);
    unsigned long long *counts = (unsigned long long *)(c->totals.p + 2);
    be_cm_write(c->stream, c->arena_out.p, c->refs_out.p, n, c->out_buckets, c->c_newid.p, c->opts.lbd_tier1, c->cm_out.p, c->orgs_out.p,
                c->learnts_out.p, counts, c->in_calls > 1 ? c->subsume_out.p : nullptr,          /* sclause.cpp:29-31 */
                c->vmap_on ? c->d_vmap.p : nullptr);
    be_event_record(c->stream, e1);
    if (n) be_copy_words(c->stream, c->h_totals + 8, c->refs_out.p + (n - 1), 8);
    CK(pull_totals(c, 6));
    c->cm_last_ref = n ? ((uint64_t)c->h_totals[8] | ((uint64_t)c->h_totals[9] << 32)) + (n - 1) : 0;
    c->cm_learnts = n ? c->h_totals[0] : 0;
    c->cm_lits_org = (uint64_t)c->h_totals[2] | ((uint64_t)c->h_totals[3] << 32);
    c->cm_lits_learnt = (uint64_t)c->h_totals[4] | ((uint64_t)c->h_totals[5] << 32);
    c->rep.reserved[7] = (uint64_t)(1000.0f * be_event_ms(e0, e1));
    c->cm_built = true;
    return SIGMA_OK;
}

/* rebuildWT(code) on the device (watch.cpp:135-157), from the CLAUSE records of build_cm: sortClause on every non-binary
 * record, then the watch lists as a stable counting sort of (watched literal, position in the attach order) - the same
 * two-level partition that builds the occurrence table, run on a scratch store of one two-literal entry per clause */
int build_watches(sigma_ctx *c, uint32_t code)
{
    CK(build_cm(c));
    if (c->wt_code == (int)code) return SIGMA_OK;
    if (c->wt_code >= 0) { c->cm_built = false; CK(build_cm(c)); }      /* another order was built before: sortClause starts from the records as written */
    const uint32_t n = c->out_ncls, nlit = c->vmap_on ? 2 * c->vmap_nvars + 2 : c->nlit;      /* inf.nDualVars after map(true) */
    NOMEM(c->wt_flags.ensure(4 * (size_t)n + 64)); NOMEM(c->wt_pos.ensure(4 * (size_t)n + 64)); NOMEM(c->wt_info.ensure((size_t)n + 16));
    NOMEM(c->wt_rec.ensure(2 * (size_t)n + 16)); NOMEM(c->hdr2.ensure((size_t)n + 16)); NOMEM(c->lits2.ensure(2 * (size_t)n + 16));
    NOMEM(c->scantmp.ensure(4 * be_scan_tmp_words(n) + 64));
    const sigma_report keep = c->rep;
    void *e0 = get_event(c), *e1 = get_event(c);
    be_event_record(c->stream, e0);
    if (n) {
        uint32_t *f = c->wt_flags.p, *p = c->wt_pos.p;
        be_wt_groups(c->stream, c->arena_out.p, c->refs_out.p, n, code, f, f + n, f + 2 * (size_t)n, f + 3 * (size_t)n);
        SgScanSet ss = { { f, f + n, f + 2 * (size_t)n, f + 3 * (size_t)n }, { p, p + n, p + 2 * (size_t)n, p + 3 * (size_t)n },
                         { c->totals.p + 16, c->totals.p + 17, c->totals.p + 18, c->totals.p + 19 } };
        be_scan_multi(c->stream, ss, 4, n, c->scantmp.p);
        be_wt_fill(c->stream, c->cm_out.p, c->arena_out.p, c->refs_out.p, n, code, p, p + n, p + 2 * (size_t)n, p + 3 * (size_t)n, c->totals.p + 16,
                   c->vmap_on ? nullptr : c->value.p,          /* after map(true) every literal of the database is unassigned: sortClause moves nothing */
                   c->hdr2.p, c->lits2.p, c->wt_info.p);
        /* the occurrence table of the scratch store = the watch lists (entries ascending = push order) */
        { DBuf<sg_u4> t = c->hdr; c->hdr = c->hdr2; c->hdr2 = t; }
        { DBuf<uint32_t> t = c->lits; c->lits = c->lits2; c->lits2 = t; }
        const uint32_t s_ncls = c->n_cls, s_pool = c->pool_used, s_nlit = c->nlit;
        const bool s_verify = c->ot_verify;
        c->n_cls = n; c->pool_used = 2 * n; c->ot_verify = false; c->nlit = nlit;
        const int rc = stage_create_ot(c, 2ull * n, n, false);
        c->n_cls = s_ncls; c->pool_used = s_pool; c->ot_verify = s_verify; c->ot_valid = false; c->nlit = s_nlit;
        { DBuf<sg_u4> t = c->hdr; c->hdr = c->hdr2; c->hdr2 = t; }
        { DBuf<uint32_t> t = c->lits; c->lits = c->lits2; c->lits2 = t; }
        if (rc != SIGMA_OK) { c->rep = keep; return rc; }
        be_wt_records(c->stream, nlit, c->ot_off.p, c->ot_len.p, c->ot_data.p, c->wt_info.p, c->wt_rec.p);
    } else be_memset(c->stream, c->ot_off.p, 0, ((size_t)nlit + 1) * 4);
    be_event_record(c->stream, e1);
    if (ctx_sync(c)) { c->rep = keep; return SIGMA_E_CUDA; }
    c->rep = keep;
    c->rep.reserved32[0] = (uint32_t)(1000.0f * be_event_ms(e0, e1));
    c->wt_code = (int)code; c->wt_records = 2ull * n; c->wt_nlit = nlit;
    return SIGMA_OK;
}

int download_impl(sigma_ctx *c, sigma_formula *out, sigma_cnf_out *cnf)
{
    if (!c || !out) return SIGMA_E_ARG;
    if (!c->executed) return fail(c, SIGMA_E_ARG, "no result: sigma_execute did not complete");
    be_bind(c->device);
    const uint32_t V = c->nvars, nlit = c->nlit;
    /* sigma_download_cnf with rest->state == NULL: only the clause database (the solver state was fetched before, with
     * sigma_download_state - the order map(true) needs) */
    const bool with_state = !(cnf && !out->state);
    bool small = with_state && (out->trail_cap < c->hsc.n_trail || out->model_cap < c->model_words);
    if (cnf) {
        CK(build_cm(c));
        const uint64_t n_orgs = c->out_ncls - c->cm_learnts;
        small = small || cnf->cm_cap < c->out_buckets + c->out_ncls || cnf->orgs_cap < n_orgs || cnf->learnts_cap < c->cm_learnts;
        cnf->cm_buckets = c->out_buckets + c->out_ncls; cnf->n_orgs = n_orgs; cnf->n_learnts = c->cm_learnts;
        cnf->lits_original = c->cm_lits_org; cnf->lits_learnt = c->cm_lits_learnt; cnf->last_ref = c->cm_last_ref;
    } else small = small || out->arena_cap < c->out_buckets || out->refs_cap < c->out_ncls;
    const uint64_t acap = out->arena_cap, mcap = out->model_cap; const uint32_t rcap = out->refs_cap, tcap = out->trail_cap;
    sigma_result_sizes(c, out);
    out->arena_cap = acap; out->model_cap = mcap; out->refs_cap = rcap; out->trail_cap = tcap;
    if (small) return fail(c, SIGMA_E_NOSPACE, "output buffers too small (required sizes filled in)");
    if (with_state && (!out->state || !out->value || !out->trail || !out->model)) return fail(c, SIGMA_E_ARG, "null output array");
    if (cnf ? (!cnf->cm || !cnf->orgs || !cnf->learnts) : (!out->arena || !out->refs)) return fail(c, SIGMA_E_ARG, "null output array");
    void *e0 = get_event(c), *e1 = get_event(c);
    be_event_record(c->stream, e0);
    uint64_t clause_bytes;
    if (cnf) {
        if (cnf->cm_buckets) be_d2h(c->stream, cnf->cm, c->cm_out.p, cnf->cm_buckets * 4);
        if (cnf->n_orgs) be_d2h(c->stream, cnf->orgs, c->orgs_out.p, cnf->n_orgs * 8);
        if (cnf->n_learnts) be_d2h(c->stream, cnf->learnts, c->learnts_out.p, cnf->n_learnts * 8);
        if (cnf->subsume) be_d2h(c->stream, cnf->subsume, c->subsume_out.p, V + 1);
        clause_bytes = cnf->cm_buckets * 4 + (cnf->n_orgs + cnf->n_learnts) * 8 + (cnf->subsume ? V + 1 : 0);
    } else {
        if (c->out_buckets) be_d2h(c->stream, out->arena, c->arena_out.p, c->out_buckets * 4);
        if (c->out_ncls) be_d2h(c->stream, out->refs, c->refs_out.p, (size_t)c->out_ncls * 8);
        clause_bytes = c->out_buckets * 4 + (uint64_t)c->out_ncls * 8;
    }
    if (with_state) {
        be_d2h(c->stream, out->state, c->state.p, V + 1);
        be_d2h(c->stream, out->value, c->value.p, nlit);
        if (c->hsc.n_trail) be_d2h(c->stream, out->trail, c->trail.p, (size_t)c->hsc.n_trail * 4);
        if (c->model_words) be_d2h(c->stream, out->model, c->model.p, (size_t)c->model_words * 4);
    }
    be_event_record(c->stream, e1);
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    c->rep.stage_ms[SIGMA_T_D2H] = be_event_ms(e0, e1);
    c->rep.d2h_bytes = clause_bytes + (V + 1) + nlit + (uint64_t)c->hsc.n_trail * 4 + c->model_words * 4;
    return SIGMA_OK;
}

} // namespace

extern "C" {

int sigma_download(sigma_ctx *c, sigma_formula *out) { return download_impl(c, out, nullptr); }

int sigma_download_cnf(sigma_ctx *c, sigma_cnf_out *out, sigma_formula *rest)
{
    if (!out) return SIGMA_E_ARG;
    return download_impl(c, rest, out);
}

int sigma_cnf_out_sizes(sigma_ctx *c, sigma_cnf_out *sz)
{
    if (!c || !sz) return SIGMA_E_ARG;
    if (!c->executed) return fail(c, SIGMA_E_ARG, "no result: sigma_execute did not complete");
    be_bind(c->device);
    CK(build_cm(c));
    sz->cm_buckets = c->out_buckets + c->out_ncls; sz->n_learnts = c->cm_learnts; sz->n_orgs = c->out_ncls - c->cm_learnts;
    sz->lits_original = c->cm_lits_org; sz->lits_learnt = c->cm_lits_learnt; sz->last_ref = c->cm_last_ref;
    return SIGMA_OK;
}

int sigma_set_var_map(sigma_ctx *c, const uint32_t *map, uint32_t new_nvars)
{
    if (!c) return SIGMA_E_ARG;
    if (!c->executed) return fail(c, SIGMA_E_ARG, "no result: sigma_execute did not complete");
    be_bind(c->device);
    c->cm_built = false; c->wt_code = -1;                   /* whatever was built used the other numbering */
    if (!map) { c->vmap_on = false; return SIGMA_OK; }
    if (new_nvars > c->nvars) return fail(c, SIGMA_E_ARG, "variable map: more mapped variables than variables");
    NOMEM(c->d_vmap.ensure((size_t)c->nvars + 1));
    be_h2d(c->stream, c->d_vmap.p, map, ((size_t)c->nvars + 1) * 4);
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    c->vmap_on = true; c->vmap_nvars = new_nvars;
    return SIGMA_OK;
}

int sigma_download_state(sigma_ctx *c, sigma_formula *out)
{
    if (!c || !out) return SIGMA_E_ARG;
    if (!c->executed) return fail(c, SIGMA_E_ARG, "no result: sigma_execute did not complete");
    be_bind(c->device);
    const bool small = out->trail_cap < c->hsc.n_trail || out->model_cap < c->model_words;
    const uint64_t mcap = out->model_cap; const uint32_t tcap = out->trail_cap;
    void *st = out->state, *va = out->value, *tr = out->trail, *mo = out->model;
    sigma_result_sizes(c, out);
    out->state = (uint8_t *)st; out->value = (int8_t *)va; out->trail = (uint32_t *)tr; out->model = (uint32_t *)mo;
    out->model_cap = mcap; out->trail_cap = tcap;
    if (small) return fail(c, SIGMA_E_NOSPACE, "output buffers too small (required sizes filled in)");
    if (!out->state || !out->value || (c->hsc.n_trail && !out->trail) || (c->model_words && !out->model)) return fail(c, SIGMA_E_ARG, "null output array");
    be_d2h(c->stream, out->state, c->state.p, (size_t)c->nvars + 1);
    be_d2h(c->stream, out->value, c->value.p, c->nlit);
    if (c->hsc.n_trail) be_d2h(c->stream, out->trail, c->trail.p, (size_t)c->hsc.n_trail * 4);
    if (c->model_words) be_d2h(c->stream, out->model, c->model.p, (size_t)c->model_words * 4);
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    return SIGMA_OK;
}

int sigma_watches_sizes(sigma_ctx *c, int priorbins, sigma_watches *sz)
{
    if (!c || !sz || priorbins < 0 || priorbins > 3) return SIGMA_E_ARG;
    if (!c->executed) return fail(c, SIGMA_E_ARG, "no result: sigma_execute did not complete");
    be_bind(c->device);
    CK(build_watches(c, (uint32_t)priorbins));
    sz->nlit = c->wt_nlit; sz->n_records = c->wt_records;
    return SIGMA_OK;
}

int sigma_download_watches(sigma_ctx *c, sigma_watches *out)
{
    if (!c || !out) return SIGMA_E_ARG;
    if (!c->executed || c->wt_code < 0) return fail(c, SIGMA_E_ARG, "sigma_download_watches without sigma_watches_sizes");
    be_bind(c->device);
    const bool small = out->offs_cap < (uint64_t)c->wt_nlit + 1 || out->records_cap < c->wt_records;
    out->nlit = c->wt_nlit; out->n_records = c->wt_records;
    if (small) return fail(c, SIGMA_E_NOSPACE, "watch buffers too small (required sizes filled in)");
    if (!out->offs || (c->wt_records && !out->records)) return fail(c, SIGMA_E_ARG, "null output array");
    be_d2h(c->stream, out->offs, c->ot_off.p, (size_t)c->wt_nlit * 4);
    if (c->wt_records) be_d2h(c->stream, out->records, c->wt_rec.p, c->wt_records * 16);
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    out->offs[c->wt_nlit] = (uint32_t)c->wt_records;
    return SIGMA_OK;
}

int sigma_get_report(sigma_ctx *c, sigma_report *rep)
{
    if (!c || !rep) return SIGMA_E_ARG;
    *rep = c->rep;
    return SIGMA_OK;
}

int sigma_run(sigma_ctx *c, const sigma_opts *o, const sigma_formula *in, sigma_formula *out, sigma_report *rep)
{
    int rc = sigma_upload(c, o, in);
    if (rc == SIGMA_OK) rc = sigma_execute(c);
    if (rc == SIGMA_OK && out) rc = sigma_download(c, out);
    if (rep && c) { *rep = c->rep; rep->status = rc; }
    return rc;
}

int sigma_snapshot(sigma_ctx *c, const char *name, void *dst, uint64_t cap_bytes, uint64_t *nbytes)
{
    if (!c || !name || !nbytes) return SIGMA_E_ARG;
    be_bind(c->device);
    const void *src = nullptr;
    uint64_t n = 0;
    const uint32_t V = c->nvars, nlit = c->nlit;
    if (!strcmp(name, "arena_in")) { src = c->d_arena_in.p; n = 4ull * c->in_buckets; }
    else if (!strcmp(name, "refs_in")) { src = c->d_refs_in.p; n = 8ull * c->in_ncls; }
    else if (!strcmp(name, "hdr")) { src = c->hdr.p; n = 16ull * c->n_cls; }
    else if (!strcmp(name, "lits")) { src = c->lits.p; n = 4ull * c->pool_used; }
    else if (!strcmp(name, "ot_off")) { src = c->ot_off.p; n = 4ull * (nlit + 1); }
    else if (!strcmp(name, "ot_len")) { src = c->ot_len.p; n = 4ull * nlit; }
    else if (!strcmp(name, "ot_data")) { src = c->ot_data.p; n = 4ull * c->ot_data.cap; }
    else if (!strcmp(name, "occ")) { src = c->occ.p; n = 4ull * nlit; }
    else if (!strcmp(name, "order")) { src = c->order.p; n = 4ull * V; }
    else if (!strcmp(name, "elected")) { src = c->elected.p; n = 4ull * c->hsc.n_elected; }
    else if (!strcmp(name, "state")) { src = c->state.p; n = V + 1; }
    else if (!strcmp(name, "value")) { src = c->value.p; n = nlit; }
    else if (!strcmp(name, "trail")) { src = c->trail.p; n = 4ull * c->hsc.n_trail; }
    else if (!strcmp(name, "model")) { src = c->model.p; n = 4ull * c->model_words; }
    else if (!strcmp(name, "marks")) { src = c->marks.p; n = V + 1; }
    else return fail(c, SIGMA_E_ARG, "unknown snapshot name");
    *nbytes = n;
    if (!dst || cap_bytes < n) return SIGMA_E_NOSPACE;
    if (n) be_d2h(c->stream, dst, src, n);
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    return SIGMA_OK;
}

} /* extern "C" */

/* ---- DIMACS parser on the device (dimacs.cpp:25-177; method in sigma_dimacs.cuh) ------------------- */
namespace {

/* the part of parser() in front of the clauses, on the host: comments, then "p cnf <vars> <clauses>" (dimacs.cpp:52-86,
 * dimacs.hpp:57-74).  Returns the offset of the first byte after the header, or 0 with *why set. */
uint64_t dimacs_header(const char *t, uint64_t n, uint32_t *nvars, uint32_t *ncls, const char **why)
{
    uint64_t i = 0;
    auto ws = [&]() { while (i < n && ((t[i] >= 9 && t[i] <= 13) || t[i] == 32)) ++i; };
    auto integer = [&](uint32_t *v, uint32_t *sign) -> bool {
        ws();
        *sign = 0;
        if (i < n && t[i] == '-') { *sign = 1; ++i; }
        else if (i < n && t[i] == '+') ++i;
        if (i >= n || t[i] < '0' || t[i] > '9') return false;
        uint32_t x = 0;
        while (i < n && t[i] >= '0' && t[i] <= '9') x = x * 10u + (uint32_t)(t[i++] - '0');
        *v = x;
        return true;
    };
    while (true) {
        ws();
        if (i >= n || t[i] == 0 || t[i] == '%') { *why = "no header before the end of the file"; return 0; }
        if (t[i] == 'c') { while (i < n && t[i] && t[i++] != '\n') { } continue; }
        if (t[i] != 'p') { *why = "a clause before the header (the reference stops with 'too many variables')"; return 0; }
        if (n - i < 5 || memcmp(t + i, "p cnf", 5) != 0) { *why = "header has wrong format"; return 0; }
        i += 5;
        uint32_t sign = 0;
        if (!integer(nvars, &sign)) { *why = "expected a digit in the header"; return 0; }
        if (sign) { *why = "number of variables in header is negative"; return 0; }
        if (*nvars == 0) { *why = "zero number of variables in header"; return 0; }
        if (*nvars >= 0x7FFFFFFEu) { *why = "number of variables not supported"; return 0; }
        if (!integer(ncls, &sign)) { *why = "expected a digit in the header"; return 0; }
        if (sign) { *why = "number of clauses in header is negative"; return 0; }
        if (*ncls == 0) { *why = "zero number of clauses in header"; return 0; }
        return i;
    }
}

int parse_dimacs_impl(sigma_ctx *c, const char *text, uint64_t bytes, sigma_dimacs *out)
{
    if (!c || !text || !out) return SIGMA_E_ARG;
    if (!c->stream) return fail(c, SIGMA_E_CUDA, c->err[0] ? c->err : "no CUDA device");
    be_bind(c->device);
    c->dm_parsed = false;
    sigma_dimacs &r = c->dm_out;
    memset(&r, 0, sizeof r);
    const char *why = nullptr;
    const uint64_t body = dimacs_header(text, bytes, &r.nvars, &r.header_clauses, &why);
    if (!body) return fail(c, SIGMA_E_ARG, why);
    const uint64_t n = bytes - body;
    if (n >= 0xFFFFFFFFull * SGD_CHUNK / 2) return fail(c, SIGMA_E_UNSUPPORTED, "file too large for 32-bit token indices");
    const uint32_t V = r.nvars, nlit = 2 * V + 2;
    const uint32_t nchunks = (uint32_t)((n + SGD_CHUNK - 1) / SGD_CHUNK), ntiles = (nchunks + 255) / 256;
    NOMEM(c->dm_text.ensure(n + 64)); NOMEM(c->dm_fn.ensure(nchunks + 16)); NOMEM(c->dm_tilefn.ensure(ntiles + 16)); NOMEM(c->dm_tstate.ensure(ntiles + 16));
    NOMEM(c->dm_cstate.ensure(nchunks + 16)); NOMEM(c->dm_ntok.ensure(nchunks + 16)); NOMEM(c->dm_nterm.ensure(nchunks + 16));
    NOMEM(c->dm_tokoff.ensure(nchunks + 16)); NOMEM(c->dm_termoff.ensure(nchunks + 16));
    NOMEM(c->scantmp.ensure(2 * be_scan_tmp_words(nchunks > 64 ? nchunks : 64) + 64));
    NOMEM(c->dm_utime.ensure(nlit)); NOMEM(c->dm_utime2.ensure(nlit)); NOMEM(c->dm_value.ensure(nlit)); NOMEM(c->dm_state.ensure(V + 1));
    NOMEM(c->dm_stats.ensure(8));
    void *e0 = get_event(c), *e1 = get_event(c), *e2 = get_event(c);
    be_event_record(c->stream, e0);
    if (n) be_h2d(c->stream, c->dm_text.p, text + body, n);
    be_event_record(c->stream, e1);
    for (int k = 0; k < 8; ++k) c->h_totals[k] = 0;
    c->h_totals[4] = 0xFFFFFFFFu;                          /* first empty clause */
    be_copy_words(c->stream, c->totals.p, c->h_totals, 8 * sizeof(uint32_t));
    /* totals: 0 tokens, 1 terminators, 2 final state, 3 bad, 4 first empty clause, 5 changed */
    uint32_t ntok = 0, ncl = 0;
    if (nchunks) {
        be_dm_chunk_states(c->stream, c->dm_text.p, n, nchunks, c->dm_fn.p, c->dm_tilefn.p, c->dm_tstate.p, c->dm_cstate.p, c->totals.p + 2);
        be_dm_tokens(c->stream, c->dm_text.p, n, nchunks, c->dm_cstate.p, V, 0, c->dm_ntok.p, c->dm_nterm.p, nullptr, nullptr, nullptr, nullptr, c->totals.p + 3);
        SgScanSet ss = { { c->dm_ntok.p, c->dm_nterm.p, nullptr, nullptr }, { c->dm_tokoff.p, c->dm_termoff.p, nullptr, nullptr },
                         { c->totals.p, c->totals.p + 1, nullptr, nullptr } };
        be_scan_multi(c->stream, ss, 2, nchunks, c->scantmp.p);
        CK(pull_totals(c, 8));
        ntok = c->h_totals[0]; ncl = c->h_totals[1];
        const uint32_t fin = c->h_totals[2];
        if (c->h_totals[3] == 2 || fin == SGD_ERR) return fail(c, SIGMA_E_ARG, "expected a digit (a character the reference's parser rejects, or a comment inside a clause)");
        if (fin == SGD_K || fin == SGD_TN || fin == SGD_SGN) return fail(c, SIGMA_E_ARG, "the last clause is not terminated by 0");
    }
    r.device_ms = 0;
    if (ncl) {
        NOMEM(c->dm_tok.ensure((size_t)ntok + 16)); NOMEM(c->dm_term.ensure((size_t)ncl + 16));
        NOMEM(c->dm_csize.ensure((size_t)ncl + 16)); NOMEM(c->dm_words.ensure((size_t)ncl + 16)); NOMEM(c->dm_kept.ensure((size_t)ncl + 16));
        NOMEM(c->dm_isunit.ensure((size_t)ncl + 16)); NOMEM(c->dm_wordoff.ensure((size_t)ncl + 16)); NOMEM(c->dm_keptoff.ensure((size_t)ncl + 16));
        NOMEM(c->dm_unitoff.ensure((size_t)ncl + 16));
        NOMEM(c->scantmp.ensure(3 * be_scan_tmp_words(ncl) + 64));
        be_dm_tokens(c->stream, c->dm_text.p, n, nchunks, c->dm_cstate.p, V, 1, nullptr, nullptr, c->dm_tokoff.p, c->dm_termoff.p, c->dm_tok.p, c->dm_term.p,
                     c->totals.p + 3);
        /* makeClause rounds: units found while parsing act on the clauses after them (sigma_dimacs.cuh) */
        be_memset(c->stream, c->dm_utime.p, 0xFF, (size_t)nlit * 4);
        uint32_t rounds = 0;
        while (true) {
            be_memset(c->stream, c->dm_utime2.p, 0xFF, (size_t)nlit * 4);
            be_dm_clauses(c->stream, c->dm_tok.p, c->dm_term.p, ncl, c->dm_utime.p, c->dm_utime2.p, c->dm_csize.p, c->totals.p + 4);
            c->h_totals[5] = 0;
            be_copy_words(c->stream, c->totals.p + 5, c->h_totals + 5, 4);
            be_dm_diff(c->stream, c->dm_utime.p, c->dm_utime2.p, nlit, c->totals.p + 5);
            CK(pull_totals(c, 8));
            ++rounds;
            if (c->h_totals[3] == 3) return fail(c, SIGMA_E_ARG, "too many variables (a literal beyond the header's variable count)");
            if (!c->h_totals[5]) break;
            DBuf<uint32_t> t = c->dm_utime; c->dm_utime = c->dm_utime2; c->dm_utime2 = t;
            if (rounds > ncl + 2) return fail(c, SIGMA_E_CUDA, "unit rounds of the parser did not converge");
        }
        r.unit_rounds = rounds;
        if (c->h_totals[4] != 0xFFFFFFFFu) {
            /* an empty clause: parser() returns false at that clause; nothing after it matters (solve.cpp:52: learnEmpty) */
            be_event_record(c->stream, e2);
            if (ctx_sync(c)) return SIGMA_E_CUDA;
            r.device_ms = be_event_ms(e1, e2);
            return SIGMA_UNSAT;
        }
        be_dm_sizes(c->stream, c->dm_tok.p, c->dm_term.p, ncl, c->dm_utime.p, c->dm_csize.p, c->dm_words.p, c->dm_kept.p, c->dm_isunit.p);
        SgScanSet ss = { { c->dm_words.p, c->dm_kept.p, c->dm_isunit.p, nullptr }, { c->dm_wordoff.p, c->dm_keptoff.p, c->dm_unitoff.p, nullptr },
                         { c->totals.p, c->totals.p + 1, c->totals.p + 6, nullptr } };
        be_scan_multi(c->stream, ss, 3, ncl, c->scantmp.p);
        CK(pull_totals(c, 8));
        r.cm_buckets = c->h_totals[0]; r.n_orgs = c->h_totals[1]; r.n_units = c->h_totals[6];
        if (r.n_orgs > r.header_clauses) return fail(c, SIGMA_E_ARG, "too many clauses (more than the header announces)");
    }
    NOMEM(c->dm_cm.ensure((size_t)r.cm_buckets + 16)); NOMEM(c->dm_orgs.ensure((size_t)r.n_orgs + 16)); NOMEM(c->dm_trail.ensure((size_t)r.n_units + 16));
    be_memset(c->stream, c->dm_value.p, 0xFF, nlit);                 /* UNDEF_VAL = -1 */
    be_memset(c->stream, c->dm_state.p, 0, V + 1);
    be_memset(c->stream, c->dm_stats.p, 0, 8 * sizeof(uint64_t));
    if (ncl)
        be_dm_write(c->stream, c->dm_tok.p, c->dm_term.p, ncl, c->dm_utime.p, c->dm_csize.p, c->dm_wordoff.p, c->dm_keptoff.p, c->dm_unitoff.p, c->dm_isunit.p,
                    c->dm_cm.p, (unsigned long long *)c->dm_orgs.p, c->dm_trail.p, c->dm_value.p, c->dm_state.p, (unsigned long long *)c->dm_stats.p,
                    r.header_clauses, c->totals.p + 3);
    be_event_record(c->stream, e2);
    uint64_t st[8];
    be_d2h(c->stream, st, c->dm_stats.p, sizeof st);
    CK(pull_totals(c, 8));
    if (c->h_totals[3] == 4) return fail(c, SIGMA_E_ARG, "too many clauses (more than the header announces)");
    r.binaries = st[0]; r.ternaries = st[1]; r.large = st[2]; r.literals = st[3]; r.max_clause_size = (uint32_t)st[4];
    r.h2d_ms = be_event_ms(e0, e1);
    r.device_ms = be_event_ms(e1, e2);
    r.body_offset = body;
    c->dm_parsed = true;
    return SIGMA_OK;
}

} // namespace

extern "C" {

int sigma_parse_dimacs(sigma_ctx *c, const char *text, uint64_t bytes, sigma_dimacs *out)
{
    const int rc = parse_dimacs_impl(c, text, bytes, out);
    if (c && out) {
        sigma_dimacs keep = *out;
        *out = c->dm_out;
        out->cm = keep.cm; out->cm_cap = keep.cm_cap; out->orgs = keep.orgs; out->orgs_cap = keep.orgs_cap;
        out->trail = keep.trail; out->trail_cap = keep.trail_cap; out->value = keep.value; out->state = keep.state;
    }
    return rc;
}

int sigma_parse_download(sigma_ctx *c, sigma_dimacs *out)
{
    if (!c || !out) return SIGMA_E_ARG;
    if (!c->dm_parsed) return fail(c, SIGMA_E_ARG, "sigma_parse_download without a successful sigma_parse_dimacs");
    be_bind(c->device);
    const sigma_dimacs &r = c->dm_out;
    if (out->cm_cap < r.cm_buckets || out->orgs_cap < r.n_orgs || out->trail_cap < r.n_units || (r.cm_buckets && !out->cm) || (r.n_orgs && !out->orgs) ||
        (r.n_units && !out->trail)) {
        sigma_dimacs keep = *out;
        *out = r;
        out->cm = keep.cm; out->cm_cap = keep.cm_cap; out->orgs = keep.orgs; out->orgs_cap = keep.orgs_cap;
        out->trail = keep.trail; out->trail_cap = keep.trail_cap; out->value = keep.value; out->state = keep.state;
        return SIGMA_E_NOSPACE;
    }
    if (r.cm_buckets) be_d2h(c->stream, out->cm, c->dm_cm.p, r.cm_buckets * 4);
    if (r.n_orgs) be_d2h(c->stream, out->orgs, c->dm_orgs.p, r.n_orgs * 8);
    if (r.n_units) be_d2h(c->stream, out->trail, c->dm_trail.p, r.n_units * 4);
    if (out->value) be_d2h(c->stream, out->value, c->dm_value.p, 2ull * r.nvars + 2);
    if (out->state) be_d2h(c->stream, out->state, c->dm_state.p, (size_t)r.nvars + 1);
    if (ctx_sync(c)) return SIGMA_E_CUDA;
    sigma_dimacs keep = *out;
    *out = r;
    out->cm = keep.cm; out->cm_cap = keep.cm_cap; out->orgs = keep.orgs; out->orgs_cap = keep.orgs_cap;
    out->trail = keep.trail; out->trail_cap = keep.trail_cap; out->value = keep.value; out->state = keep.state;
    return SIGMA_OK;
}

} /* extern "C" */
