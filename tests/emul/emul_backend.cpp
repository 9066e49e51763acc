/* emul_backend.cpp - TEST INFRASTRUCTURE ONLY (never part of libsigma_b200.so).
 *
 * A serial host stand-in for seqfrost_b200/csrc/sigma_backend.h so that the pass driver
 * (sigma_pass.cpp) and the per-item bodies (sigma_items.cuh) can be exercised against the
 * reference in a container without a GPU.  "Device memory" is the host heap; every "kernel" is a
 * plain loop over its work items, optionally in a shuffled order (SIGMA_EMUL_SHUFFLE=seed) to
 * shake out order dependence that the real, concurrent kernels would expose.
 */
#include "../../seqfrost_b200/csrc/sigma_backend.h"
#include "../../seqfrost_b200/csrc/sigma_items.cuh"
#include "../../seqfrost_b200/csrc/sigma_dimacs.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

static uint64_t g_launches = 0;
static long g_shuffle = -1;

static std::vector<uint32_t> item_order(uint32_t n)
{
    std::vector<uint32_t> o(n);
    std::iota(o.begin(), o.end(), 0u);
    if (g_shuffle >= 0) { std::mt19937 rng((unsigned)g_shuffle + n); std::shuffle(o.begin(), o.end(), rng); }
    return o;
}

int be_init(int, void **stream_out, char *, size_t)
{
    const char *e = getenv("SIGMA_EMUL_SHUFFLE");
    g_shuffle = e ? atol(e) : -1;
    *stream_out = (void *)1;
    return 0;
}
void be_shutdown(void *) {}
void be_bind(int) {}
void *be_malloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void be_free(void *p) { free(p); }
void *be_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void be_host_free(void *p) { free(p); }
int be_host_register(void *, size_t) { return 0; }
int be_host_unregister(void *) { return 0; }
int be_h2d(void *, void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int be_d2h(void *, void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int be_d2d(void *, void *d, const void *s, size_t n) { memmove(d, s, n); return 0; }
int be_memset(void *, void *d, int b, size_t n) { memset(d, b, n); return 0; }
int be_sync(void *, char *, size_t) { return 0; }
struct Ev { double t; };
static double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
void *be_event_create(void) { return new Ev{ 0 }; }
void be_event_destroy(void *e) { delete (Ev *)e; }
void be_event_record(void *, void *e) { ((Ev *)e)->t = now_ms(); }
float be_event_ms(void *a, void *b) { return (float)(((Ev *)b)->t - ((Ev *)a)->t); }
uint64_t be_launch_count(void *) { return g_launches; }
void be_copy_words(void *, void *d, const void *s, size_t n) { memmove(d, s, n); }

void be_extract_sizes(void *, const uint32_t *cm, uint64_t cm_buckets, const uint64_t *crefs, uint32_t n, const uint8_t *stencil, uint64_t stencil_size,
                      uint32_t *sizes, uint32_t *live, uint32_t *bad)
{
    g_launches++;
    *bad = 0;
    for (uint32_t i : item_order(n)) {
        sizes[i] = sg_extract_size(cm, cm_buckets, stencil, stencil_size, crefs[i]);
        if (sizes[i] == SG_EXTRACT_BAD) { *bad = 1; sizes[i] = 0; }
        live[i] = sizes[i] != 0;
    }
}
void be_extract_write(void *, const uint32_t *cm, const uint64_t *crefs, uint32_t n, const uint32_t *sizes, const uint32_t *newoff,
                      const uint32_t *newid, uint32_t *arena, uint64_t *refs)
{
    g_launches++;
    for (uint32_t i : item_order(n)) {
        if (!sizes[i]) continue;
        refs[newid[i]] = newoff[i];
        sg_extract_write(cm, crefs[i], arena + newoff[i]);
    }
}
void be_cm_flags(void *, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t *learnt)
{
    g_launches++;
    for (uint32_t i : item_order(n)) learnt[i] = (arena[refs[i]] & SG_ST_MASK) == SG_LEARNT;
}
void be_cm_write(void *, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t, const uint32_t *learnt_before,
                 int lbd_tier1, uint32_t *cm, uint64_t *orgs, uint64_t *learnts, unsigned long long *counts, uint8_t *subsume, const uint32_t *vmap)
{
    g_launches++;
    counts[0] = counts[1] = 0;
    for (uint32_t i : item_order(n)) {
        const uint64_t r = refs[i];
        const uint32_t *src = arena + r;
        sg_cm_write(src, cm + r + i, lbd_tier1, vmap);
        if ((src[0] & SG_ST_MASK) == SG_LEARNT) { learnts[learnt_before[i]] = r + i; counts[1] += src[2]; }
        else { orgs[i - learnt_before[i]] = r + i; counts[0] += src[2]; }
        if (subsume && (src[0] & SG_ADDED))
            for (uint32_t k = 0; k < src[2]; ++k) subsume[SG_ABS(src[3 + k])] = 1;
    }
}
void be_ere_maxsz(void *, const SgView &v, const uint8_t *size8, uint8_t *maxsz)
{
    g_launches++;
    for (uint32_t lit = 0; lit < v.nlit; ++lit) {
        uint8_t m = 0;
        for (uint32_t i = 0; i < v.ot_len[lit]; ++i) { const uint8_t s8 = size8[v.ot_data[v.ot_off[lit] + i]]; if (s8 > m) m = s8; }
        maxsz[lit] = m;
    }
}
void be_ere_filt(void *, const SgView &v, uint32_t *filt, uint8_t *size8)
{
    g_launches++;
    for (uint32_t c = 0; c < v.n_cls; ++c) {
        const uint32_t size = (v.hdr[c].z & SG_DELETED) ? 0u : v.hdr[c].y;
        filt[2 * c] = size; filt[2 * c + 1] = v.hdr[c].w;
        size8[c] = (uint8_t)(size < 255u ? size : 255u);
    }
}
void be_ot_order(void *, const SgView &v)
{
    g_launches++;
    for (uint32_t lit : item_order(v.nlit)) sg_idsort_item(v, lit);
}
void be_scan_multi(void *s, const SgScanSet &ss, uint32_t count, uint32_t n, uint32_t *tmp)
{
    for (uint32_t k = 0; k < count; ++k) be_scan_u32(s, ss.in[k], ss.out[k], n, ss.total[k], tmp);
}
void be_ingest_sizes(void *, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, uint32_t *sizes, uint32_t *bad)
{
    g_launches++;
    *bad = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t r = refs[i];
        sizes[i] = 0;
        if (r + 3 > n_buckets) { *bad = 2; continue; }
        if (r + 3 + arena[r + 2] > n_buckets) { *bad = 2; continue; }
        sizes[i] = arena[r + 2];
    }
}
void be_ingest_fused(void *, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, sg_u4 *hdr, uint32_t *lits, uint32_t *gap)
{
    g_launches++;
    *gap = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t r = refs[i];
        if (r + 3 > n_buckets || r + 3 + arena[r + 2] > n_buckets) { *gap = 1; return; }
        const uint32_t size = arena[r + 2];
        if ((i + 1 < n && refs[i + 1] != r + 3 + size) || (i + 1 == n && r + 3 + size != n_buckets) || (i == 0 && r != 0) || r < 3ull * i) { *gap = 1; return; }
        sg_u4 h; h.x = (uint32_t)(r - 3ull * i); h.y = size; h.z = arena[r]; h.w = arena[r + 1];
        hdr[i] = h;
        for (uint32_t k = 0; k < size; ++k) lits[h.x + k] = arena[r + 3 + k];
    }
}
void be_ingest_copy(void *, const uint32_t *arena, const uint64_t *refs, uint32_t n, const uint32_t *offs, sg_u4 *hdr, uint32_t *lits)
{
    g_launches++;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t r = refs[i];
        sg_u4 h; h.x = offs[i]; h.y = arena[r + 2]; h.z = arena[r]; h.w = arena[r + 1];
        hdr[i] = h;
        for (uint32_t k = 0; k < h.y; ++k) lits[h.x + k] = arena[r + 3 + k];
    }
}
size_t be_scan_tmp_words(uint32_t) { return 1; }
void be_scan_u32(void *, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total, uint32_t *)
{
    g_launches++;
    uint32_t s = 0;
    for (uint32_t i = 0; i < n; ++i) { const uint32_t x = in[i]; out[i] = s; s += x; }
    if (total) *total = s;
}
void be_hist(void *, const SgView &v, uint32_t *counts)
{
    g_launches++;
    for (uint32_t c = 0; c < v.n_cls; ++c) {
        if (v.hdr[c].z & SG_DELETED) continue;
        const uint32_t *l = v.lits + v.hdr[c].x;
        for (uint32_t k = 0; k < v.hdr[c].y; ++k) counts[l[k]]++;
    }
}
/* two-level createOT: the stand-in keeps the call structure (count / caller's scan / scatter / lists) with one bucket
 * per 256 literals and tiles of 64 clauses, so that the driver's scan and buffer layout are exercised */
size_t be_ot_build_tmp_words(uint32_t nlit, uint32_t n_cls, uint64_t lits)
{
    if (!nlit || !n_cls) return 0;
    const size_t NB = (nlit + 255) / 256, T = (n_cls + 63) / 64;
    return NB * T + 8 + be_scan_tmp_words((uint32_t)(NB * T + 8)) + 8 + 2 * (size_t)lits + 64;
}
int be_ot_build_plan(uint32_t nlit, uint32_t n_cls, uint64_t lits, uint32_t *tmp, SgOtBuild *p)
{
    if (!nlit || !n_cls) return 1;
    (void)lits;
    p->shift = 8; p->NB = (nlit + 255) / 256; p->per_tile = 64; p->T = (n_cls + 63) / 64;
    p->ncnt = (size_t)p->NB * p->T;
    p->cnt = tmp; p->scantmp = tmp + p->ncnt + 8;
    p->pairs = (void *)(((uintptr_t)(p->scantmp + be_scan_tmp_words((uint32_t)(p->ncnt + 8))) + 15) & ~(uintptr_t)15);
    return 0;
}
void be_ot_build_count(void *, const SgView &v, const SgOtBuild &p, uint32_t *bits)
{
    g_launches++;
    memset(p.cnt, 0, p.ncnt * 4);
    if (bits) memset(bits, 0, ((size_t)v.n_cls / 32 + 1) * 4);
    for (uint32_t c : item_order(v.n_cls)) {
        if (v.hdr[c].z & SG_DELETED) continue;
        if (bits) bits[c >> 5] |= 1u << (c & 31);
        for (uint32_t k = 0; k < v.hdr[c].y; ++k) p.cnt[(size_t)(v.lits[v.hdr[c].x + k] >> p.shift) * p.T + c / p.per_tile]++;
    }
}
void be_ot_build_scatter(void *, const SgView &v, const SgOtBuild &p)
{
    g_launches++;
    uint32_t *pairs = (uint32_t *)p.pairs;
    std::vector<uint32_t> cur(p.cnt, p.cnt + p.ncnt);
    for (uint32_t c : item_order(v.n_cls)) {           /* any order inside a tile, as on the device */
        if (v.hdr[c].z & SG_DELETED) continue;
        for (uint32_t k = 0; k < v.hdr[c].y; ++k) {
            const uint32_t lit = v.lits[v.hdr[c].x + k];
            const uint32_t pos = cur[(size_t)(lit >> p.shift) * p.T + c / p.per_tile]++;
            pairs[2 * (size_t)pos] = lit; pairs[2 * (size_t)pos + 1] = c;
        }
    }
}
void be_ot_build_lists(void *, const SgView &v, const SgOtBuild &p, const uint32_t *total)
{
    g_launches++;
    const uint32_t *pairs = (const uint32_t *)p.pairs;
    for (uint32_t b : item_order(p.NB)) {
        const uint32_t g0 = p.cnt[(size_t)b * p.T], g1 = b + 1 < p.NB ? p.cnt[(size_t)(b + 1) * p.T] : *total;
        const uint32_t lit0 = b << p.shift, W = 1u << p.shift;
        std::vector<uint32_t> n(W + 1, 0), cur(W, 0);
        for (uint32_t q = g0; q < g1; ++q) n[pairs[2 * (size_t)q] - lit0]++;
        uint32_t e = 0;
        for (uint32_t i = 0; i < W; ++i) {
            if (lit0 + i < v.nlit) { v.ot_off[lit0 + i] = g0 + e; v.ot_len[lit0 + i] = n[i]; }
            cur[i] = e; e += n[i];
        }
        for (uint32_t q = g0; q < g1; ++q) v.ot_data[g0 + cur[pairs[2 * (size_t)q] - lit0]++] = pairs[2 * (size_t)q + 1];
        for (uint32_t i = 0; i < W && lit0 + i < v.nlit; ++i) std::sort(v.ot_data + v.ot_off[lit0 + i], v.ot_data + v.ot_off[lit0 + i] + v.ot_len[lit0 + i]);
    }
}
void be_scatter(void *, const SgView &v, uint32_t *cursor)
{
    g_launches++;
    for (uint32_t c : item_order(v.n_cls)) {
        if (v.hdr[c].z & SG_DELETED) continue;
        const uint32_t *l = v.lits + v.hdr[c].x;
        for (uint32_t k = 0; k < v.hdr[c].y; ++k) v.ot_data[cursor[l[k]]++] = c;
    }
}
void be_scores(void *, const SgView &v)
{
    g_launches++;
    v.sc->max_score = 0;
    for (uint32_t x = 1; x <= v.nvars; ++x) {
        v.scores[x] = v.occ[2 * x] * v.occ[2 * x + 1]; v.order[x - 1] = x;
        if (v.scores[x] > v.sc->max_score) v.sc->max_score = v.scores[x];
    }
}
size_t be_sort_tmp_words(uint32_t) { return 1; }
void be_sort_pairs(void *, uint32_t *keys, uint32_t *vals, uint32_t n, uint32_t *, uint32_t *, uint32_t *, uint32_t)
{
    g_launches++;
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> k2(n), v2(n);
    for (uint32_t i = 0; i < n; ++i) { k2[i] = keys[idx[i]]; v2[i] = vals[idx[i]]; }
    memcpy(keys, k2.data(), n * 4ull); memcpy(vals, v2.data(), n * 4ull);
}
void be_rank(void *, const SgView &v)
{
    g_launches++;
    for (uint32_t r = 0; r < v.nvars; ++r) v.rank[v.order[r]] = r;
}
void be_elect_batch(void *, const SgView &v, uint32_t r0, uint32_t r1, uint32_t *wl)
{
    g_launches++;
    uint32_t n = 0;
    for (uint32_t i : item_order(r1 - r0)) {
        const uint32_t var = v.order[r0 + i];
        if (SG_ER_STAT(v.rank[var]) == SG_E_UNDECIDED) wl[n++] = var;
    }
    v.sc->wl_count[0] = n;
}
void be_elect_round(void *, const SgView &v, const uint32_t *wl_in, uint32_t *wl_out, int parity)
{
    g_launches++;
    const uint32_t n = v.sc->wl_count[parity];
    uint32_t m = 0;
    for (uint32_t i : item_order(n))
        if (sg_elect_round_item(v, wl_in[i])) wl_out[m++] = wl_in[i];
    v.sc->wl_count[parity ^ 1] = m;
}
int be_elect_all(void *, const SgView &, uint32_t *, uint32_t *, uint32_t, uint32_t) { return 1; }
int be_elect_async(void *, const SgView &) { return 1; }
int be_vo_sort(void *, const SgView &, uint32_t *, uint32_t *, uint32_t *, uint32_t) { return 1; }
int be_has_sortot_wide(void) { return 0; }
int be_sortot_wide(void *, const SgView &) { return 1; }
int be_ere_replay_block(void *, const SgView &, uint32_t, uint32_t, uint32_t *, uint32_t *) { return 1; }   /* host-driven rounds in the stand-in */
void be_acting_flags(void *, const SgView &v, uint32_t *fa, uint32_t *fe)
{
    g_launches++;
    for (uint32_t r = 0; r < v.nvars; ++r) {
        const uint32_t s = SG_ER_STAT(v.rank[v.order[r]]);
        const bool in = r < v.sc->brk_rank;
        fa[r] = in && (s == SG_E_BOTH || s == SG_E_PONLY);
        fe[r] = in && s == SG_E_BOTH;
    }
}
void be_acting_scatter(void *, const SgView &v, const uint32_t *fa, const uint32_t *pa, const uint32_t *fe, const uint32_t *pe)
{
    g_launches++;
    for (uint32_t r = 0; r < v.nvars; ++r) {
        const uint32_t var = v.order[r];
        if (fa[r]) v.acting[pa[r]] = var | (SG_ER_STAT(v.rank[var]) == SG_E_PONLY ? 0x80000000u : 0u);
        if (fe[r]) v.elected[pe[r]] = var;
    }
}
void be_items(void *, int kind, const SgView &v, uint32_t n)
{
    g_launches++;
    for (uint32_t i0 : item_order(n)) {
        const uint32_t i = v.item_base + i0;
        switch (kind) {
        case SGK_IDSORT: sg_idsort_item(v, i); break;
        case SGK_REDUCE: sg_reduce_item(v, i); break;
        case SGK_ELECT_PREP: sg_elect_prep_item(v, i + 1); break;
        case SGK_SORTOT: sg_sortot_warp_item(v, 2 * v.elected[i >> 1] + (i & 1)); break;
        case SGK_SORTOT_LONG: sg_sortot_long_item(v, 2 * v.elected[i >> 1] + (i & 1)); break;
        case SGK_SCRATCH_LEN: {
            const uint32_t var = v.elected[i];
            const uint32_t a = v.ot_len[2 * var], b = v.ot_len[2 * var + 1];
            v.scr_off[i] = (a > b ? a : b) + 2;
        } break;
        case SGK_SUB: sg_sub_item(v, i); break;
        case SGK_VE_COUNT: sg_ve_count_item(v, i); break;
        case SGK_VE_EMIT: sg_ve_emit_item(v, i); break;
        case SGK_NBH: sg_neighbourhood_item(v, i); break;
        case SGK_BCE_COUNT: sg_bce_count_item(v, i); break;
        case SGK_BCE_EMIT: sg_bce_emit_item(v, i); break;
        case SGK_ERE_CHUNKS: sg_ere_chunks_item(v, i); break;
        case SGK_ERE_PROPOSE: sg_ere_propose_item(v, i); break;
        case SGK_ERE_DIRTY: sg_ere_dirty_item(v, i); break;
        case SGK_ERE_CLAIM: sg_ere_claim_item(v, i); break;
        case SGK_ERE_SAFE: sg_ere_safe_item(v, i); break;
        case SGK_ERE_APPLY: sg_ere_apply_item(v, i); break;
        case SGK_ERE_FLAG01: v.ve_ncls[i] = v.ere_flags[i] != 0; break;
        case SGK_ERE_LIST: if (v.ere_flags[i]) v.acting[v.ve_cls_off[i]] = i; break;
        default: abort();
        }
    }
}
void be_serial(void *, int kind, const SgView &v, uint32_t arg, uint32_t arg2)
{
    g_launches++;
    switch (kind) {
    case SGS_PROP: sg_prop_serial(v, arg); break;
    case SGS_MARKS: sg_marks_serial(v, arg); break;
    case SGS_APPLY_UNITS: sg_apply_units_serial(v, arg); break;
    case SGS_ERE_REPLAY: sg_ere_replay_serial(v, arg, arg2); break;
    case SGS_ERE_SPLIT: sg_ere_split_serial(v, arg, arg2); break;
    default: abort();
    }
}
void be_sort_units(void *, const SgView &v, uint32_t n)
{
    g_launches++;
    std::sort(v.ubuf, v.ubuf + n, [](const SgUnit &a, const SgUnit &b) { return a.eidx != b.eidx ? a.eidx < b.eidx : a.seq < b.seq; });
}
void be_count_live(void *, const SgView &v)
{
    g_launches++;
    uint32_t nc = 0; unsigned long long nl = 0;
    for (uint32_t c = 0; c < v.n_cls; ++c) if (!(v.hdr[c].z & SG_DELETED)) { nc++; nl += v.hdr[c].y; }
    v.sc->live_cls = nc; v.sc->live_lits = nl;
}
void be_count_melted(void *, const SgView &v, uint32_t *out)
{
    g_launches++;
    uint32_t n = 0;
    for (uint32_t x = 1; x <= v.nvars; ++x) if (v.state[x] & SG_MELTED_M) n++;
    *out = n;
}
size_t be_live_tmp_words(uint32_t n_cls) { return 2ull * n_cls + 4; }
void be_live_scan(void *, const SgView &v, uint32_t *tmp, uint32_t *totals)
{   /* tmp[2c], tmp[2c+1] = new id / new offset of clause c */
    g_launches++;
    uint32_t id = 0, off = 0;
    for (uint32_t c = 0; c < v.n_cls; ++c) {
        tmp[2 * c] = id; tmp[2 * c + 1] = off;
        if (!(v.hdr[c].z & SG_DELETED)) { id++; off += v.hdr[c].y; }
    }
    totals[0] = id; totals[1] = off;
}
void be_compact_apply(void *, const SgView &v, const uint32_t *tmp, sg_u4 *hdr2, uint32_t *lits2)
{
    g_launches++;
    for (uint32_t c = 0; c < v.n_cls; ++c) {
        if (v.hdr[c].z & SG_DELETED) continue;
        sg_u4 h = v.hdr[c];
        const uint32_t *l = v.lits + h.x;
        h.x = tmp[2 * c + 1];
        hdr2[tmp[2 * c]] = h;
        for (uint32_t k = 0; k < h.y; ++k) lits2[h.x + k] = l[k];
    }
}
void be_export_apply(void *, const SgView &v, const uint32_t *tmp, uint32_t *arena, uint64_t *refs)
{
    g_launches++;
    for (uint32_t c = 0; c < v.n_cls; ++c) {
        if (v.hdr[c].z & SG_DELETED) continue;
        const sg_u4 h = v.hdr[c];
        const uint64_t r = 3ull * tmp[2 * c] + tmp[2 * c + 1];
        refs[tmp[2 * c]] = r;
        arena[r] = h.z; arena[r + 1] = h.w; arena[r + 2] = h.y;
        for (uint32_t k = 0; k < h.y; ++k) arena[r + 3 + k] = v.lits[h.x + k];
    }
}

/* serial restatement of the incremental occurrence-table update (k_otu_* in sigma_kernels.cu) */
size_t be_ot_update_tmp_words(uint32_t nlit, uint32_t) { return 3 * (size_t)nlit + 64; }
size_t be_ot_bits_words(uint32_t n_cls) { return (size_t)n_cls / 32 + 2; }
void be_ot_bitmap(void *, const SgView &v, uint32_t *bits)
{
    g_launches++;
    for (uint32_t w = 0; w <= v.n_cls / 32; ++w) bits[w] = 0;
    for (uint32_t c = 0; c < v.n_cls; ++c) if (!(v.hdr[c].z & SG_DELETED)) bits[c >> 5] |= 1u << (c & 31);
}
void be_ot_update_plan(void *, const SgView &v, const SgOtUpdate &u)
{
    g_launches++;
    uint32_t *addc = u.tmp, *flag = u.tmp + 2 * (size_t)v.nlit;
    memset(addc, 0, (size_t)v.nlit * 4); memset(flag, 0, (size_t)v.nlit * 4);
    for (uint32_t c = 0; c < v.n_cls; ++c) {
        const sg_u4 h = v.hdr[c];
        const bool live = !(h.z & SG_DELETED);
        const uint32_t *l = v.lits + h.x;
        if (c >= u.first_new) { if (live) for (uint32_t k = 0; k < h.y; ++k) { addc[l[k]]++; flag[l[k]] = 1; } }
        else if (!live && ((u.bits[c >> 5] >> (c & 31)) & 1u)) for (uint32_t k = 0; k < h.y; ++k) flag[l[k]] = 1;
        if (live) u.bits[c >> 5] |= 1u << (c & 31); else u.bits[c >> 5] &= ~(1u << (c & 31));
    }
    for (uint32_t i = 0; i < u.n_add; ++i) {
        const uint32_t lit = v.ot_add[2 * i], c = v.ot_add[2 * i + 1];
        if (v.hdr[c].z & SG_DELETED) continue;
        addc[lit]++; flag[lit] = 1;
    }
    for (uint32_t i = 0; i < u.n_elected; ++i) { flag[2 * v.elected[i]] = 1; flag[2 * v.elected[i] + 1] = 1; }
    uint32_t need = 0;
    for (uint32_t lit = 0; lit < v.nlit; ++lit) if (addc[lit] && !(v.state[lit >> 1] & SG_MELTED_M)) need += v.ot_len[lit] + addc[lit];
    if (u.total) *u.total = need;
}
void be_ot_update(void *, const SgView &v, const SgOtUpdate &u)
{
    g_launches++;
    uint32_t *addc = u.tmp, *cursor = u.tmp + (size_t)v.nlit, *flag = u.tmp + 2 * (size_t)v.nlit;
    for (uint32_t lit = 0; lit < v.nlit; ++lit) {
        if (v.state[lit >> 1] & SG_MELTED_M) { v.ot_len[lit] = 0; flag[lit] = 0; continue; }
        if (!flag[lit]) continue;
        const uint32_t len = v.ot_len[lit], a = addc[lit];
        const uint32_t *src = v.ot_data + v.ot_off[lit];
        uint32_t *dst = v.ot_data + v.ot_off[lit];
        if (a) { const uint32_t base = *u.bump; *u.bump += len + a; dst = v.ot_data + base; v.ot_off[lit] = base; }
        uint32_t n = 0;
        for (uint32_t i = 0; i < len; ++i) { const uint32_t c = src[i]; if (v.hdr[c].z & SG_DELETED) continue; dst[n++] = c; }
        std::sort(dst, dst + n);
        v.ot_len[lit] = n + a;
        cursor[lit] = v.ot_off[lit] + n;
        flag[lit] = a ? 2u : 0u;
    }
    for (uint32_t c = u.first_new; c < v.n_cls; ++c) {
        const sg_u4 h = v.hdr[c];
        if (h.z & SG_DELETED) continue;
        for (uint32_t k = 0; k < h.y; ++k) v.ot_data[cursor[v.lits[h.x + k]]++] = c;
    }
    for (uint32_t i = 0; i < u.n_add; ++i) {
        const uint32_t lit = v.ot_add[2 * i], c = v.ot_add[2 * i + 1];
        if (v.hdr[c].z & SG_DELETED) continue;
        v.ot_data[cursor[lit]++] = c;
    }
    for (uint32_t lit = 0; lit < v.nlit; ++lit) if (flag[lit] == 2u) std::sort(v.ot_data + v.ot_off[lit], v.ot_data + v.ot_off[lit] + v.ot_len[lit]);
}
void be_ot_compare(void *, const SgView &v, const uint32_t *off2, const uint32_t *len2, const uint32_t *data2, uint32_t *mismatch)
{
    g_launches++;
    uint32_t bad = 0;
    for (uint32_t l = 0; l < v.nlit; ++l) {
        if (v.ot_len[l] != len2[l]) {
            if (getenv("SIGMA_EMUL_DEBUG") && bad < 4) {
                fprintf(stderr, "lit %u: len %u vs rebuild %u state %u\n  upd:", l, v.ot_len[l], len2[l], v.state[l >> 1]);
                for (uint32_t i = 0; i < v.ot_len[l]; ++i) fprintf(stderr, " %u%s", v.ot_data[v.ot_off[l] + i], (v.hdr[v.ot_data[v.ot_off[l] + i]].z & SG_DELETED) ? "x" : "");
                fprintf(stderr, "\n  new:");
                for (uint32_t i = 0; i < len2[l]; ++i) fprintf(stderr, " %u", data2[off2[l] + i]);
                fprintf(stderr, "\n");
            }
            bad++; continue; }
        for (uint32_t i = 0; i < len2[l]; ++i) if (v.ot_data[v.ot_off[l] + i] != data2[off2[l] + i]) {
            if (getenv("SIGMA_EMUL_DEBUG") && bad < 4) {
                fprintf(stderr, "lit %u: entries differ (len %u) state %u off %u\n  upd:", l, len2[l], v.state[l >> 1], v.ot_off[l]);
                for (uint32_t q = 0; q < v.ot_len[l]; ++q) fprintf(stderr, " %u", v.ot_data[v.ot_off[l] + q]);
                fprintf(stderr, "\n  new:");
                for (uint32_t q = 0; q < len2[l]; ++q) fprintf(stderr, " %u", data2[off2[l] + q]);
                fprintf(stderr, "\n");
            }
            bad++; break; }
    }
    *mismatch = bad;
}

void be_iota_u32(void *, uint32_t *p, uint32_t n) { g_launches++; for (uint32_t i = 0; i < n; ++i) p[i] = i; }
/* ---- DIMACS parser: the bodies of sigma_dimacs.cuh, one work item after the other ---------------------------------- */
void be_dm_chunk_states(void *, const uint8_t *text, uint64_t n, uint32_t nchunks, uint32_t *fn, uint32_t *, uint32_t *, uint32_t *cstate, uint32_t *final_state)
{
    g_launches++;
    for (uint32_t i = 0; i < nchunks; ++i) {
        const uint64_t lo = (uint64_t)i * SGD_CHUNK, hi = lo + SGD_CHUNK < n ? lo + SGD_CHUNK : n;
        fn[i] = sgd_chunk_fn(text, lo, hi);
    }
    uint32_t pre = SGD_IDENT;                              /* composition, not state chasing: the scan the device runs */
    for (uint32_t i = 0; i < nchunks; ++i) { cstate[i] = sgd_apply(pre, SGD_B); pre = sgd_compose(pre, fn[i]); }
    *final_state = sgd_apply(pre, SGD_B);
}
void be_dm_tokens(void *, const uint8_t *text, uint64_t n, uint32_t nchunks, const uint32_t *cstate, uint32_t nvars, int emit, uint32_t *ntok,
                  uint32_t *nterm, const uint32_t *tokoff, const uint32_t *termoff, uint32_t *tok, uint32_t *term, uint32_t *bad)
{
    g_launches++;
    for (uint32_t i : item_order(nchunks)) {
        const uint64_t lo = (uint64_t)i * SGD_CHUNK, hi = lo + SGD_CHUNK < n ? lo + SGD_CHUNK : n;
        uint32_t a, b;
        const uint32_t end = sgd_chunk_tokens(text, lo, hi, n, cstate[i], nvars, emit != 0, &a, &b, emit ? tokoff[i] : 0u, emit ? termoff[i] : 0u, tok, term, bad);
        if (!emit) { ntok[i] = a; nterm[i] = b; if (end == SGD_ERR) *bad = 2; }
    }
}
void be_dm_clauses(void *, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, uint32_t *utime_new, uint32_t *csize, uint32_t *unsat)
{
    g_launches++;
    for (uint32_t c : item_order(ncl)) {
        const uint32_t first = c ? term[c - 1] + 1 : 0u, last = term[c];
        uint32_t kind;
        const uint32_t n = sgd_make_clause(tok, first, last, c, utime, nullptr, &kind);
        csize[c] = kind == 3 ? 0u : (kind == 0 && n == 0 && !sgd_satisfied_or_empty(tok, first, last, c, utime)) ? 0xFFFFFFFFu : n;
        if (kind == 3 && c < *unsat) *unsat = c;
        if (kind == 1) { uint32_t &u = utime_new[sgd_unit_literal(tok, first, last, c, utime)]; if (c < u) u = c; }
    }
}
void be_dm_diff(void *, const uint32_t *a, const uint32_t *b, uint32_t n, uint32_t *changed)
{
    g_launches++;
    for (uint32_t i = 0; i < n; ++i) if (a[i] != b[i]) *changed = 1;
}
void be_dm_sizes(void *, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, const uint32_t *csize, uint32_t *words,
                 uint32_t *kept, uint32_t *isunit)
{
    g_launches++;
    for (uint32_t c : item_order(ncl)) {
        const uint32_t n = csize[c] == 0xFFFFFFFFu ? 0u : csize[c];
        words[c] = n >= 2 ? n + 4u : 0u;
        kept[c] = n >= 2;
        uint32_t u = 0;
        if (n == 1) u = utime[sgd_unit_literal(tok, c ? term[c - 1] + 1 : 0u, term[c], c, utime)] == c;
        isunit[c] = u;
    }
}
void be_dm_write(void *, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, const uint32_t *csize, const uint32_t *wordoff,
                 const uint32_t *keptoff, const uint32_t *unitoff, const uint32_t *isunit, uint32_t *cm, unsigned long long *orgs, uint32_t *trail,
                 int8_t *value, uint8_t *state, unsigned long long *stats, uint32_t header_clauses, uint32_t *bad)
{
    g_launches++;
    for (uint32_t c : item_order(ncl)) {
        if (csize[c] == 0xFFFFFFFFu) { if (keptoff[c] + 1 > header_clauses) *bad = 4; continue; }
        const uint32_t n = csize[c];
        const uint32_t first = c ? term[c - 1] + 1 : 0u, last = term[c];
        if (n >= 2) {
            uint32_t *dst = cm + wordoff[c];
            uint32_t kind;
            dst[0] = (1u << 4) | (n == 2 ? 1u << 5 : 0u); dst[1] = n; dst[2] = 2u; dst[3] = 0u;
            (void)sgd_make_clause(tok, first, last, c, utime, dst + 4, &kind);
            orgs[keptoff[c]] = wordoff[c];
            stats[n == 2 ? 0 : n == 3 ? 1 : 2]++;
            stats[3] += n;
            if (n > stats[4]) stats[4] = n;
        } else if (n == 1 && isunit[c]) {
            const uint32_t lit = sgd_unit_literal(tok, first, last, c, utime);
            trail[unitoff[c]] = lit;
            value[lit] = 1; value[lit ^ 1u] = 0;
            state[lit >> 1] = (uint8_t)SG_FROZEN_M;
        }
    }
}

/* ---- rebuildWT --------------------------------------------------------------------------------------------------- */
void be_wt_groups(void *, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t code, uint32_t *f0, uint32_t *f1, uint32_t *f2, uint32_t *f3)
{
    g_launches++;
    for (uint32_t i : item_order(n)) {
        const uint32_t *src = arena + refs[i];
        const uint32_t g = sg_wt_group(src[2] == 2u, (src[0] & SG_ST_MASK) == SG_LEARNT, code);
        f0[i] = g == 0; f1[i] = g == 1; f2[i] = g == 2; f3[i] = g == 3;
    }
}
void be_wt_fill(void *, uint32_t *cm, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t code, const uint32_t *p0, const uint32_t *p1,
                const uint32_t *p2, const uint32_t *p3, const uint32_t *totals, const int8_t *value, sg_u4 *hdr, uint32_t *lits, sg_u4 *winfo)
{
    g_launches++;
    for (uint32_t i : item_order(n)) {
        const uint64_t r = refs[i];
        const uint32_t *src = arena + r;
        const uint32_t size = src[2];
        const uint32_t g = sg_wt_group(size == 2u, (src[0] & SG_ST_MASK) == SG_LEARNT, code);
        uint32_t seq = g == 0 ? p0[i] : g == 1 ? p1[i] : g == 2 ? p2[i] : p3[i];
        if (g > 0) seq += totals[0];
        if (g > 1) seq += totals[1];
        if (g > 2) seq += totals[2];
        const uint64_t cref = r + i;
        uint32_t *l = cm + cref + SG_CM_HDR;
        if (size != 2u && value) sg_sort_clause(l, (int)size, value);
        const uint32_t a = l[0], b = l[1];
        hdr[seq].x = 2u * seq; hdr[seq].y = 2u; hdr[seq].z = 0u; hdr[seq].w = 0u;
        lits[2 * seq] = SG_FLIP(a); lits[2 * seq + 1] = SG_FLIP(b);
        winfo[seq].x = (uint32_t)cref; winfo[seq].y = (uint32_t)(cref >> 32); winfo[seq].z = size; winfo[seq].w = a ^ b;
    }
}
void be_wt_records(void *, uint32_t nlit, const uint32_t *ot_off, const uint32_t *ot_len, const uint32_t *ot_data, const sg_u4 *winfo, sg_u4 *rec)
{
    g_launches++;
    for (uint32_t lit : item_order(nlit))
        for (uint32_t k = 0; k < ot_len[lit]; ++k) {
            const sg_u4 w = winfo[ot_data[ot_off[lit] + k]];
            sg_u4 &o = rec[ot_off[lit] + k];
            o.x = w.x; o.y = w.y; o.z = w.w ^ SG_FLIP(lit); o.w = w.z;
        }
}

/* prop by frontiers: the kernels' loops, one visit at a time (same per-visit functions) */
int  be_has_prop_frontier(void) { return 1; }
void be_apply_units(void *, const SgView &v, uint32_t n, uint32_t *first, uint32_t *flags, uint32_t *pos, uint32_t *total, uint32_t *)
{   /* the kernels' steps, one position at a time */
    g_launches += 7;
    for (uint32_t k = 0; k < n; ++k) if (k < first[v.ubuf[k].lit]) first[v.ubuf[k].lit] = k;
    uint32_t run = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t lit = v.ubuf[k].lit;
        const int8_t val = v.value[lit];
        uint32_t f = 0;
        if (val & (int8_t)-2) { if (first[lit] == k) { if (first[SG_FLIP(lit)] < k) v.sc->unsat = 1; else f = 1; } }
        else if (!val) v.sc->unsat = 1;
        flags[k] = f; pos[k] = run; run += f;
    }
    *total = run;
    if (!v.sc->unsat) {
        for (uint32_t k = 0; k < n; ++k) if (flags[k]) {
            const uint32_t lit = v.ubuf[k].lit;
            v.trail[v.sc->n_trail + pos[k]] = lit; v.state[SG_ABS(lit)] = (uint8_t)SG_FROZEN_M; v.value[lit] = 1; v.value[SG_FLIP(lit)] = 0;
        }
        v.sc->n_trail += run; v.sc->max_frozen += run; v.sc->unassigned -= run; v.sc->units_forced += run;
    }
    for (uint32_t k = 0; k < n; ++k) first[v.ubuf[k].lit] = 0xFFFFFFFFu;
}
void be_pp_mark(void *, const SgView &v, const SgProp &pp)
{
    g_launches++;
    for (uint32_t k = 0; k < pp.p1 - pp.p0; ++k) { pp.fpos[SG_FLIP(v.trail[pp.p0 + k])] = pp.p0 + k; pp.cnt[k] = 0; }
    for (int q = 0; q < 5; ++q) pp.ctl[q] = 0;
}
template <class F> static void pp_for_visits(const SgView &v, const SgProp &pp, F f)
{
    for (uint32_t k = 0; k < pp.p1 - pp.p0; ++k) {
        const uint32_t p = pp.p0 + k, flipped = SG_FLIP(v.trail[p]);
        const uint32_t len = v.ot_len[flipped];
        const uint32_t *d = sg_list(v, flipped);
        for (uint32_t i = 0; i < len; ++i) f(k, p, i, flipped, d[i]);
    }
}
void be_pp_round(void *, const SgView &v, const SgProp &pp, int diff)
{
    g_launches += 2;
    pp_for_visits(v, pp, [&](uint32_t, uint32_t p, uint32_t i, uint32_t flipped, uint32_t c) {
        uint32_t unit = 0;
        if (sg_pp_visit(v, pp, p, i, flipped, c, &unit) != SG_PP_UNIT) return;
        const unsigned long long t = ((unsigned long long)p << 32) | i;
        if (pp.disc_new[unit] == SG_PP_NEVER) pp.touched_new[pp.ctl[1]++] = unit;
        if (t < pp.disc_new[unit]) pp.disc_new[unit] = t;
    });
    if (!diff) return;
    for (uint32_t k = 0; k < pp.ctl[0]; ++k) if (pp.disc[pp.touched[k]] != pp.disc_new[pp.touched[k]]) pp.ctl[2] = 1;
    for (uint32_t k = 0; k < pp.ctl[1]; ++k) if (pp.disc[pp.touched_new[k]] != pp.disc_new[pp.touched_new[k]]) pp.ctl[2] = 1;
}
void be_pp_next(void *, const SgProp &pp)
{
    g_launches += 2;
    for (uint32_t k = 0; k < pp.ctl[0]; ++k) pp.disc[pp.touched[k]] = SG_PP_NEVER;
    pp.ctl[0] = pp.ctl[1]; pp.ctl[1] = 0; pp.ctl[2] = 0;
}
void be_pp_count(void *, const SgView &v, const SgProp &pp, uint32_t *)
{
    g_launches += 4;
    pp_for_visits(v, pp, [&](uint32_t k, uint32_t p, uint32_t i, uint32_t flipped, uint32_t c) {
        uint32_t unit = 0;
        const int out = sg_pp_visit(v, pp, p, i, flipped, c, &unit);
        if (out == SG_PP_CONFLICT) pp.ctl[3] = 1;
        if (out == SG_PP_UNIT) pp.cnt[k]++;
    });
    uint32_t run = 0;
    for (uint32_t k = 0; k < pp.p1 - pp.p0; ++k) { pp.base[k] = run; run += pp.cnt[k]; }
    pp.ctl[4] = run;
}
static void pp_idle(const SgView &v, const SgProp &pp)
{
    for (uint32_t k = 0; k < pp.p1 - pp.p0; ++k) pp.fpos[SG_FLIP(v.trail[pp.p0 + k])] = SG_PP_NONE;
    for (uint32_t k = 0; k < pp.ctl[0]; ++k) pp.disc[pp.touched[k]] = SG_PP_NEVER;
    for (uint32_t k = 0; k < pp.ctl[1]; ++k) pp.disc_new[pp.touched_new[k]] = SG_PP_NEVER;
}
void be_pp_commit(void *, const SgView &v, const SgProp &pp, uint32_t n_new)
{
    g_launches += 5;
    std::vector<uint32_t> run(pp.p1 - pp.p0, 0);
    std::vector<uint32_t> units(n_new);
    pp_for_visits(v, pp, [&](uint32_t k, uint32_t p, uint32_t i, uint32_t flipped, uint32_t c) {
        uint32_t unit = 0;
        if (sg_pp_visit(v, pp, p, i, flipped, c, &unit) == SG_PP_UNIT) units[pp.base[k] + run[k]++] = unit;
    });
    for (uint32_t j = 0; j < n_new; ++j) v.trail[pp.p1 + j] = units[j];
    pp_for_visits(v, pp, [&](uint32_t, uint32_t, uint32_t, uint32_t, uint32_t c) {
        const uint32_t bit = 1u << (c & 31u);
        if (pp.claim[c >> 5] & bit) return;
        pp.claim[c >> 5] |= bit;
        sg_pp_rewrite(v, pp, c);
    });
    for (uint32_t k = 0; k < pp.p1 - pp.p0; ++k) {
        const uint32_t assign = v.trail[pp.p0 + k];
        sg_toblivion_list(v, assign);
        v.ot_len[SG_FLIP(assign)] = 0;
    }
    pp_idle(v, pp);
    for (uint32_t j = 0; j < n_new; ++j) {
        const uint32_t u = v.trail[pp.p1 + j];
        v.state[SG_ABS(u)] = (uint8_t)SG_FROZEN_M; v.value[u] = 1; v.value[SG_FLIP(u)] = 0;
    }
    v.sc->propagated = pp.p1; v.sc->n_trail = pp.p1 + n_new; v.sc->max_frozen += n_new; v.sc->unassigned -= n_new;
}
void be_pp_conflict(void *, const SgView &v, const SgProp &pp) { g_launches += 3; pp_idle(v, pp); v.sc->unsat = 1; }
