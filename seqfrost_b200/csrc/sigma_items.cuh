/* sigma_items.cuh - the per-work-item bodies of the irregular SIGmA stages.
 *
 * Every function here processes ONE independent work item (one elected variable, one
 * occurrence list, one election candidate) and touches only that item's clause neighbourhood,
 * which LCVE guarantees to be disjoint from every other elected variable's (reference
 * src/lcve.cpp:64-83,121-145).  The CUDA kernels in sigma_kernels.cu run one item per thread;
 * the behaviour of each item follows the cited reference lines so that the result is
 * bit-identical to the reference's sequential loop (ledger 4-7).
 */
#ifndef SIGMA_ITEMS_CUH
#define SIGMA_ITEMS_CUH

#include "sigma_types.h"

/* ------------------------------------------------------------------------------------------ */
/* atomics (plain memory operations when compiled for the host test harness)                    */
/* ------------------------------------------------------------------------------------------ */
SG_HD uint32_t sg_atomic_add(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    uint32_t o = *p; *p = o + v; return o;
#endif
}
SG_HD void sg_atomic_sub(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicSub(p, v);
#else
    *p -= v;
#endif
}
SG_HD void sg_atomic_min(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
SG_HD void sg_atomic_add64(unsigned long long *p, unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
SG_HD void sg_set_flag(uint32_t *p)
{
#if defined(__CUDA_ARCH__)
    atomicExch(p, 1u);
#else
    *p = 1u;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* clause accessors                                                                             */
/* ------------------------------------------------------------------------------------------ */
#define SG_OFF(v, c)   ((v).hdr[c].x)
#define SG_SIZE(v, c)  ((int)(v).hdr[c].y)
#define SG_W0(v, c)    ((v).hdr[c].z)
#define SG_SIG(v, c)   ((v).hdr[c].w)
#define SG_LITS(v, c)  ((v).lits + (v).hdr[c].x)

SG_HD bool sg_original(const SgView &v, uint32_t c) { return (SG_W0(v, c) & SG_ST_MASK) == 0; }
SG_HD bool sg_deleted(const SgView &v, uint32_t c)  { return (SG_W0(v, c) & SG_DELETED) != 0; }
SG_HD bool sg_learnt(const SgView &v, uint32_t c)   { return (SG_W0(v, c) & SG_LEARNT) != 0; }
SG_HD bool sg_molten(const SgView &v, uint32_t c)   { return (SG_W0(v, c) & SG_MOLTEN) != 0; }
SG_HD void sg_melt(const SgView &v, uint32_t c)     { v.hdr[c].z |= SG_MOLTEN; }
SG_HD void sg_freeze(const SgView &v, uint32_t c)   { v.hdr[c].z &= ~SG_MOLTEN; }
SG_HD void sg_mark_deleted(const SgView &v, uint32_t c) { v.hdr[c].z = (v.hdr[c].z & ~SG_ST_MASK) | SG_DELETED; }

SG_HD uint32_t *sg_list(const SgView &v, uint32_t lit) { return v.ot_data + v.ot_off[lit]; }

SG_HD void sg_push_unit(const SgView &v, uint32_t eidx, uint32_t &seq, uint32_t lit)
{
    const uint32_t k = sg_atomic_add(&v.sc->n_units, 1u);
    if (k < v.ubuf_cap) {
        SgUnit u; u.eidx = eidx; u.seq = seq; u.lit = lit; u.pad = 0;
        v.ubuf[k] = u;
    }
    seq++;
}

/* ascending sort of a short literal array (sort.hpp:36-44 SORT: any correct sort gives the
 * same result on plain integers) */
SG_HD void sg_sort_lits(uint32_t *d, int n)
{
    for (int i = 1; i < n; ++i) {
        const uint32_t t = d[i];
        int j = i;
        while (j > 0 && d[j - 1] > t) { d[j] = d[j - 1]; --j; }
        d[j] = t;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* sortOT (simplify.cpp:85-100): CNF_CMP_KEY (simplify.hpp:30-47) through SORTCMP (sort.hpp:64-72) */
/* <= 24 entries: the reference's insertion sort; larger: libstdc++ std::sort (introsort),       */
/* restated step for step so that even the order of tied keys is the reference's.                */
/* ------------------------------------------------------------------------------------------ */
struct SgKey { uint32_t size, first, last, sig; };
SG_HD bool sg_key_lt(const SgKey &x, const SgKey &y)
{   /* CNF_CMP_KEY (simplify.hpp:30-47) */
    if ((int)x.size < (int)y.size) return true;
    if ((int)x.size > (int)y.size) return false;
    if (x.first < y.first) return true;
    if (x.first > y.first) return false;
    if (x.last < y.last) return true;
    if (x.last > y.last) return false;
    return x.sig < y.sig;
}
/* the same order over keys fetched once: elements are positions in `key` */
struct SgKeyCache { const SgKey *key; };
SG_HD bool sg_key_less(const SgKeyCache &kc, uint32_t a, uint32_t b) { return sg_key_lt(kc.key[a], kc.key[b]); }

SG_HD bool sg_key_less(const SgView &v, uint32_t a, uint32_t b)
{
    const sg_u4 x = v.hdr[a], y = v.hdr[b];
    const int xs = (int)x.y, ys = (int)y.y;
    if (xs < ys) return true;
    if (xs > ys) return false;
    uint32_t xk = v.lits[x.x], yk = v.lits[y.x];
    if (xk < yk) return true;
    if (xk > yk) return false;
    xk = v.lits[x.x + xs - 1], yk = v.lits[y.x + ys - 1];
    if (xk < yk) return true;
    if (xk > yk) return false;
    return x.w < y.w;
}

SG_HD void sg_ins_sort_ref(const SgView &v, uint32_t *d, int size)
{   /* sort.hpp:92-108 */
    if (size == 2) { if (sg_key_less(v, d[1], d[0])) { uint32_t t = d[0]; d[0] = d[1]; d[1] = t; } }
    else if (size > 2) {
        for (int i = 1; i < size; ++i) {
            if (sg_key_less(v, d[i], d[i - 1])) {
                const uint32_t tmp = d[i];
                int j = i, j1 = i - 1;
                do { d[j--] = d[j1]; } while (j != 0 && sg_key_less(v, tmp, d[--j1]));
                d[j] = tmp;
            }
        }
    }
}

/* libstdc++ (GCC 13) bits/stl_algo.h + bits/stl_heap.h, restated over an index range */
template <class K> SG_HD void sg_std_unguarded_linear_insert(const K &v, uint32_t *d, int last)
{
    const uint32_t val = d[last];
    int next = last - 1;
    while (sg_key_less(v, val, d[next])) { d[last] = d[next]; last = next; --next; }
    d[last] = val;
}
template <class K> SG_HD void sg_std_insertion_sort(const K &v, uint32_t *d, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (sg_key_less(v, d[i], d[first])) {
            const uint32_t val = d[i];
            for (int k = i; k > first; --k) d[k] = d[k - 1];
            d[first] = val;
        } else sg_std_unguarded_linear_insert(v, d, i);
    }
}
template <class K> SG_HD void sg_std_adjust_heap(const K &v, uint32_t *d, int first, int hole, int len, uint32_t value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (sg_key_less(v, d[first + child], d[first + child - 1])) child--;
        d[first + hole] = d[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        d[first + hole] = d[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;       /* __push_heap */
    while (hole > top && sg_key_less(v, d[first + parent], value)) {
        d[first + hole] = d[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    d[first + hole] = value;
}
template <class K> SG_HD void sg_std_heapsort(const K &v, uint32_t *d, int first, int last)
{   /* __partial_sort(first,last,last): __heap_select = make_heap, then __sort_heap */
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            const uint32_t value = d[first + parent];
            sg_std_adjust_heap(v, d, first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    int l = last;
    while (l - first > 1) {
        --l;
        const uint32_t value = d[l];
        d[l] = d[first];
        sg_std_adjust_heap(v, d, first, 0, l - first, value);
    }
}
template <class K> SG_HDN inline void sg_std_sort(const K &v, uint32_t *d, int n)
{
    if (n < 2) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    int sf[64], sl[64], sd[64], sp = 0;
    sf[0] = 0; sl[0] = n; sd[0] = 2 * lg; sp = 1;
    while (sp) {
        --sp;
        int first = sf[sp], last = sl[sp], depth = sd[sp];
        while (last - first > 16) {
            if (depth == 0) { sg_std_heapsort(v, d, first, last); break; }
            --depth;
            /* __unguarded_partition_pivot */
            const int mid = first + (last - first) / 2;
            {   /* __move_median_to_first(first, first+1, mid, last-1) */
                const int a = first + 1, b = mid, c = last - 1;
                int m;
                if (sg_key_less(v, d[a], d[b])) {
                    if (sg_key_less(v, d[b], d[c])) m = b;
                    else if (sg_key_less(v, d[a], d[c])) m = c;
                    else m = a;
                } else if (sg_key_less(v, d[a], d[c])) m = a;
                else if (sg_key_less(v, d[b], d[c])) m = c;
                else m = b;
                const uint32_t t = d[first]; d[first] = d[m]; d[m] = t;
            }
            int lo = first + 1, hi = last;
            while (true) {
                while (sg_key_less(v, d[lo], d[first])) ++lo;
                --hi;
                while (sg_key_less(v, d[first], d[hi])) --hi;
                if (!(lo < hi)) break;
                const uint32_t t = d[lo]; d[lo] = d[hi]; d[hi] = t;
                ++lo;
            }
            const int cut = lo;
            /* recurse on [cut,last), iterate on [first,cut) */
            if (sp < 64) { sf[sp] = cut; sl[sp] = last; sd[sp] = depth; sp++; }
            last = cut;
        }
    }
    /* __final_insertion_sort */
    if (n > 16) {
        sg_std_insertion_sort(v, d, 0, 16);
        for (int i = 16; i < n; ++i) sg_std_unguarded_linear_insert(v, d, i);
    } else sg_std_insertion_sort(v, d, 0, n);
}

/* short lists (the common case): fetch every key once, then run the reference's insertion sort
 * (sort.hpp:92-108) on the local copies - the same comparisons in the same order, so the same result */
SG_HDN inline void sg_sortot_small(const SgView &v, uint32_t *d, int size)
{
    SgKey key[24];
    uint32_t id[24];
    for (int b = 0; b < size; b += 8) {
        /* ids, then headers, then first/last literals of 8 entries: independent loads issued together */
        sg_u4 h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (b + k < size) id[b + k] = d[b + k];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (b + k < size) h[k] = v.hdr[id[b + k]];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (b + k < size) {
            key[b + k].size = h[k].y; key[b + k].sig = h[k].w;
            key[b + k].first = v.lits[h[k].x]; key[b + k].last = v.lits[h[k].x + h[k].y - 1];
        }
    }
    bool moved = false;
    if (size == 2) {
        if (sg_key_lt(key[1], key[0])) { const uint32_t t = id[0]; id[0] = id[1]; id[1] = t; moved = true; }
    } else {
        for (int i = 1; i < size; ++i) {
            if (sg_key_lt(key[i], key[i - 1])) {
                const SgKey tk = key[i];
                const uint32_t ti = id[i];
                int j = i, j1 = i - 1;
                do { key[j] = key[j1]; id[j] = id[j1]; j--; } while (j != 0 && sg_key_lt(tk, key[--j1]));
                key[j] = tk; id[j] = ti;
                moved = true;
            }
        }
    }
    if (moved) for (int i = 0; i < size; ++i) d[i] = id[i];
}

/* one occurrence list of an elected variable */
#define SG_SORTOT_CACHED_MAX 64
SG_HDN inline void sg_sortot_item(const SgView &v, uint32_t lit)
{
    const int n = (int)v.ot_len[lit];
    if (n < 2) return;
    uint32_t *d = sg_list(v, lit);
    if (n <= 24) sg_sortot_small(v, d, n);
    else if (n <= SG_SORTOT_CACHED_MAX) {
        /* every key fetched once (batched), then std::sort over positions with the keys in local memory: the same
         * comparisons and moves as sorting the ids themselves, without a header + literal fetch per comparison */
        SgKey key[SG_SORTOT_CACHED_MAX];
        uint32_t id[SG_SORTOT_CACHED_MAX], q[SG_SORTOT_CACHED_MAX];
        for (int b = 0; b < n; b += 8) {
            sg_u4 h[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) if (b + k < n) id[b + k] = d[b + k];
#pragma unroll
            for (int k = 0; k < 8; ++k) if (b + k < n) h[k] = v.hdr[id[b + k]];
#pragma unroll
            for (int k = 0; k < 8; ++k) if (b + k < n) {
                key[b + k].size = h[k].y; key[b + k].sig = h[k].w;
                key[b + k].first = v.lits[h[k].x]; key[b + k].last = v.lits[h[k].x + h[k].y - 1];
                q[b + k] = (uint32_t)(b + k);
            }
        }
        SgKeyCache kc; kc.key = key;
        sg_std_sort(kc, q, n);
        uint32_t ties = 0;
        for (int i = 1; i < n; ++i)
            if (!sg_key_less(kc, q[i - 1], q[i])) ties++;
        for (int i = 0; i < n; ++i) d[i] = id[q[i]];
        if (ties) sg_atomic_add64(&v.sc->sort_ties_gt24, ties);
    } else {
        sg_std_sort(v, d, n);
        uint32_t ties = 0;
        for (int i = 1; i < n; ++i)
            if (!sg_key_less(v, d[i - 1], d[i])) ties++;
        if (ties) sg_atomic_add64(&v.sc->sort_ties_gt24, ties);
    }
}

/* in-place heap sort (ascending): the O(n log n) fallback wherever an insertion sort could meet a long, badly ordered run */
SG_HD void sg_sift_down_u32(uint32_t *a, int root, int end)
{
    const uint32_t val = a[root];
    int hole = root;
    for (;;) {
        int child = 2 * hole + 1;
        if (child >= end) break;
        if (child + 1 < end && a[child + 1] > a[child]) child++;
        if (a[child] <= val) break;
        a[hole] = a[child];
        hole = child;
    }
    a[hole] = val;
}
SG_HD void sg_heapsort_u32(uint32_t *a, int n)
{
    for (int s = n / 2 - 1; s >= 0; --s) sg_sift_down_u32(a, s, n);
    for (int end = n - 1; end > 0; --end) {
        const uint32_t t = a[0]; a[0] = a[end]; a[end] = t;
        sg_sift_down_u32(a, 0, end);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* occurrence-list order fix after the atomic scatter: ascending clause id == the reference's   */
/* push order (simplify.cpp:50-58)                                                              */
/* ------------------------------------------------------------------------------------------ */
SG_HD void sg_idsort_item(const SgView &v, uint32_t lit)
{
    const int n = (int)v.ot_len[lit];
    uint32_t *d = sg_list(v, lit);
    /* the scatter leaves a list nearly sorted (an entry is displaced by at most the occurrences of its literal among the
     * clauses in flight), which an insertion sort fixes in about one pass - except for a hub literal that sits in a large
     * share of all clauses: there the displacement grows with the list, so long lists take the n log n route */
    if (n > 1024) { sg_heapsort_u32(d, n); return; }
    for (int i = 1; i < n; ++i) {
        const uint32_t t = d[i];
        int j = i;
        while (j > 0 && d[j - 1] > t) { d[j] = d[j - 1]; --j; }
        d[j] = t;
    }
}

/* reduceOL (simplify.cpp:63-72) */
SG_HD void sg_reduce_item(const SgView &v, uint32_t lit)
{
    const int n = (int)v.ot_len[lit];
    if (!n) return;
    uint32_t *d = sg_list(v, lit);
    int j = 0;
    for (int i = 0; i < n; ++i) {
        if (sg_deleted(v, d[i])) continue;
        d[j++] = d[i];
    }
    v.ot_len[lit] = (uint32_t)j;
}

/* ------------------------------------------------------------------------------------------ */
/* units: enqueueUnit (solve.hpp:373-398, markFrozen :485-490)                                   */
/* ------------------------------------------------------------------------------------------ */
SG_HD bool sg_unassigned(const SgView &v, uint32_t lit) { return (v.value[lit] & (int8_t)-2) != 0; }

SG_HD void sg_enqueue_unit(const SgView &v, uint32_t lit)
{
    const uint32_t var = SG_ABS(lit);
    v.state[var] = (uint8_t)SG_FROZEN_M;
    v.sc->max_frozen++;
    v.value[lit] = 1;
    v.value[SG_FLIP(lit)] = 0;
    v.trail[v.sc->n_trail++] = lit;
    v.sc->unassigned--;
}

/* applies the units collected by one SUB or VE stage in the reference's sequential order.
 * Run by ONE thread; `n` units already sorted by (eidx, seq). */
SG_HDN inline void sg_apply_units_serial(const SgView &v, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t lit = v.ubuf[k].lit;
        if (sg_unassigned(v, lit)) { sg_enqueue_unit(v, lit); v.sc->units_forced++; }
        else if (!v.value[lit]) { v.sc->unsat = 1; return; }   /* subsume.cpp:64-68, bounded.cpp:290-294 */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* prop (propelim.cpp:23-81): BCP over the occurrence table; sequential, ONE thread.            */
/* ------------------------------------------------------------------------------------------ */
SG_HD void sg_toblivion_list(const SgView &v, uint32_t lit)
{   /* simplify.hpp:385-391 */
    const int n = (int)v.ot_len[lit];
    const uint32_t *d = sg_list(v, lit);
    for (int i = 0; i < n; ++i) sg_mark_deleted(v, d[i]);
    v.ot_len[lit] = 0;
}

/* `wide` != 0: return as soon as that many units wait on the trail (the frontier kernels below take over) */
SG_HDN inline void sg_prop_serial(const SgView &v, uint32_t wide)
{
    SgScalars *sc = v.sc;
    while (sc->propagated < sc->n_trail) {
        if (wide && sc->n_trail - sc->propagated >= wide) return;
        const uint32_t assign = v.trail[sc->propagated++], flipped = SG_FLIP(assign);
        const int n = (int)v.ot_len[flipped];
        const uint32_t *d = sg_list(v, flipped);
        for (int i = 0; i < n; ++i) {
            const uint32_t c = d[i];
            if (sg_deleted(v, c)) continue;
            /* propClause (propelim.cpp:23-44), including its partial rewrite of a satisfied clause */
            uint32_t *l = SG_LITS(v, c);
            const int size = SG_SIZE(v, c);
            uint32_t sig = 0;
            int j = 0;
            bool sat = false;
            for (int k = 0; k < size; ++k) {
                const uint32_t other = l[k];
                if (other != flipped) {
                    if (v.value[other] > 0) { sat = true; break; }
                    l[j++] = other;
                    sig |= SG_HASH(other);
                }
            }
            if (sat) { sg_mark_deleted(v, c); continue; }
            v.hdr[c].w = sig;
            v.hdr[c].y = (uint32_t)(size - 1);
            const int nsize = size - 1;
            if (!nsize) { sc->unsat = 1; return; }
            if (nsize == 1) {
                const uint32_t unit = l[0];
                if (sg_unassigned(v, unit)) sg_enqueue_unit(v, unit);
                else { sc->unsat = 1; return; }
            }
        }
        v.ot_len[flipped] = 0;
        sg_toblivion_list(v, assign);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* prop by frontiers.                                                                           */
/*                                                                                              */
/* The frontier is trail[p0, p1): every unit that is assigned and not yet propagated.  The      */
/* reference takes them in trail order and walks the list of the falsified literal in list      */
/* order, so every clause visit has a time t = (trail position p, list index i).  What a visit  */
/* does depends on the past only through (a) how many literals of the clause earlier visits     */
/* removed and (b) which of its other literals are true at t.  Literals assigned before the     */
/* frontier started are true or false for every visit; the only moving part is the set of units */
/* the frontier itself discovers, each true from the time of the visit that found it.  With     */
/* disc[lit] = that time, a visit is a pure function of the clause, fpos[] (literal -> trail    */
/* position that removes it) and the entries of disc[] that are EARLIER than t.  So the whole   */
/* frontier is evaluated at once with a guess of disc[] (initially: nothing discovered), the    */
/* discoveries it reports (earliest visit per literal) become the next guess, and the guess     */
/* that reproduces itself is the sequential result: by induction over t, after k rounds all     */
/* entries earlier than the k-th discovery are final.  Two rounds when the discovered units do  */
/* not meet each other's clauses inside the frontier, which is the common case.                 */
/* New units go on the trail in order of t: per frontier position a count, a prefix sum, and    */
/* the list order inside the position.                                                          */
/* ------------------------------------------------------------------------------------------ */
#define SG_PP_NONE 0xFFFFFFFFu
#define SG_PP_NEVER 0xFFFFFFFFFFFFFFFFull
enum { SG_PP_SKIP = 0, SG_PP_SAT = 1, SG_PP_SHRINK = 2, SG_PP_UNIT = 3, SG_PP_CONFLICT = 4 };

SG_HD uint32_t sg_pp_index(const SgView &v, uint32_t lit, uint32_t c)
{
    const uint32_t n = v.ot_len[lit];
    const uint32_t *d = sg_list(v, lit);
    for (uint32_t i = 0; i < n; ++i) if (d[i] == c) return i;
    return SG_PP_NONE;
}

/* the visit (p, i) of clause c, reached through `flipped` */
SG_HDN inline int sg_pp_visit(const SgView &v, const SgProp &pp, uint32_t p, uint32_t i, uint32_t flipped, uint32_t c, uint32_t *unit)
{
    const sg_u4 h = v.hdr[c];
    if (h.z & SG_DELETED) return SG_PP_SKIP;
    const uint32_t *l = v.lits + h.x;
    unsigned long long m = SG_PP_NEVER;            /* first time one of the other literals turns true */
    uint32_t earlier = 0, last_p = 0, last_lit = 0, left = 0, left_lit = 0;
    for (uint32_t k = 0; k < h.y; ++k) {
        const uint32_t other = l[k];
        if (other == flipped) continue;
        const int8_t val = v.value[other];
        if (val > 0) return SG_PP_SKIP;            /* true since before the frontier: the clause goes with that literal's list */
        const uint32_t fp = pp.fpos[other];
        if (fp != SG_PP_NONE && fp < p) {          /* an earlier visit of this clause */
            if (!earlier || fp > last_p) { last_p = fp; last_lit = other; }
            earlier++;
            continue;
        }
        left++; left_lit = other;
        if (val & (int8_t)-2) { const unsigned long long d = pp.disc[other]; if (d < m) m = d; }
    }
    const unsigned long long t = ((unsigned long long)p << 32) | i;
    if (earlier && m != SG_PP_NEVER) {
        /* an earlier visit that came after m found the clause satisfied; the latest one decides */
        const uint32_t mp = (uint32_t)(m >> 32);
        bool gone = last_p > mp;
        if (!gone && last_p == mp) gone = sg_pp_index(v, last_lit, c) > (uint32_t)m;
        if (gone) return SG_PP_SKIP;
    }
    if (m < t) return SG_PP_SAT;
    if (!left) return SG_PP_CONFLICT;
    if (left == 1) {
        if (!(v.value[left_lit] & (int8_t)-2)) return SG_PP_CONFLICT;       /* false: assigned, not propagated yet */
        if (pp.disc[SG_FLIP(left_lit)] < t) return SG_PP_CONFLICT;
        *unit = left_lit;
        return SG_PP_UNIT;
    }
    return SG_PP_SHRINK;
}

/* the clause after the frontier (disc[] final): every visit before the first true literal removed its literal, a visit after
 * it deleted the clause.  Run once per clause. */
SG_HDN inline void sg_pp_rewrite(const SgView &v, const SgProp &pp, uint32_t c)
{
    const sg_u4 h = v.hdr[c];
    if (h.z & SG_DELETED) return;
    uint32_t *l = v.lits + h.x;
    unsigned long long m = SG_PP_NEVER;
    uint32_t visits = 0, last_p = 0, last_lit = 0;
    for (uint32_t k = 0; k < h.y; ++k) {
        const uint32_t other = l[k];
        const int8_t val = v.value[other];
        if (val > 0) return;                       /* deleted with the list of that literal */
        const uint32_t fp = pp.fpos[other];
        if (fp != SG_PP_NONE) {
            if (!visits || fp > last_p) { last_p = fp; last_lit = other; }
            visits++;
        } else if (val & (int8_t)-2) { const unsigned long long d = pp.disc[other]; if (d < m) m = d; }
    }
    if (!visits) return;
    if (m != SG_PP_NEVER) {
        const uint32_t mp = (uint32_t)(m >> 32);
        bool gone = last_p > mp;
        if (!gone && last_p == mp) gone = sg_pp_index(v, last_lit, c) > (uint32_t)m;
        if (gone) { v.hdr[c].z = (h.z & ~SG_ST_MASK) | SG_DELETED; return; }
    }
    uint32_t j = 0, sig = 0;
    for (uint32_t k = 0; k < h.y; ++k) {
        const uint32_t other = l[k];
        if (pp.fpos[other] != SG_PP_NONE) continue;
        l[j++] = other;
        sig |= SG_HASH(other);
    }
    v.hdr[c].y = j;
    v.hdr[c].w = sig;
}

/* ------------------------------------------------------------------------------------------ */
/* warp-cooperative execution of one elected variable                                           */
/*                                                                                              */
/* SUB and the BVE decision run one WARP per elected variable.  Control flow is warp-uniform    */
/* (every lane follows the reference's sequential decision tree with the same data); the heavy  */
/* inner loops - sorted-clause merge tests over the |P| x |N| pair space, and "first clause in  */
/* list order that ..." scans - are split over the lanes and recombined with ballots/reductions */
/* so that the outcome is the sequential one.  Clause writes are made by lane 0.                */
/* (The host test harness instantiates the same code with a warp of width 1.)                   */
/* ------------------------------------------------------------------------------------------ */
struct SgLane { uint32_t lane, width; };

SG_HD SgLane sg_lane_ctx()
{
    SgLane L;
#if defined(__CUDA_ARCH__)
    L.lane = threadIdx.x & 31u; L.width = 32u;
#else
    L.lane = 0u; L.width = 1u;
#endif
    return L;
}
SG_HD uint32_t sg_ballot(bool p)
{
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xffffffffu, p);
#else
    return p ? 1u : 0u;
#endif
}
SG_HD int sg_wsum(int x)
{
#if defined(__CUDA_ARCH__)
    return __reduce_add_sync(0xffffffffu, x);
#else
    return x;
#endif
}
SG_HD uint32_t sg_wor(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __reduce_or_sync(0xffffffffu, x);
#else
    return x;
#endif
}
SG_HD void sg_wsync()
{
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}
/* value of lane 0, for data another warp may be changing: every lane must act on ONE reading */
SG_HD uint32_t sg_bcast(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __shfl_sync(0xffffffffu, x, 0);
#else
    return x;
#endif
}
SG_HD int sg_ffs(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)m);
#else
    return __builtin_ffs((int)m);
#endif
}
SG_HD int sg_popc(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

/* index of the first entry in [0,n) satisfying pred, or -1; pred must be side-effect free */
template <class P> SG_HD int sg_find_first(const SgLane &L, int n, P pred)
{
    for (int base = 0; base < n; base += (int)L.width) {
        const int idx = base + (int)L.lane;
        const bool p = idx < n && pred(idx);
        const uint32_t m = sg_ballot(p);
        if (m) return base + sg_ffs(m) - 1;
    }
    return -1;
}

/* sortOT for the common short list (simplify.cpp:85-100, <= 24 entries: the reference's insertion sort, i.e. a STABLE sort
 * by CNF_CMP_KEY): one warp per list.  Lane i fetches the key of entry i - all fetches of the list in flight at once, no
 * per-thread key arrays - and its final position is its rank: the number of entries with a smaller key, or an equal key
 * and a smaller index.  That is exactly the stable order. */
SG_HDN inline void sg_sortot_warp_item(const SgView &v, uint32_t lit)
{
    const int n = (int)v.ot_len[lit];
    if (n <= (int)v.sortot_warp_min || n > 24) return;
    uint32_t *d = sg_list(v, lit);
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31u);
    uint32_t id = 0;
    SgKey k; k.size = 0; k.first = 0; k.last = 0; k.sig = 0;
    if (lane < n) {
        id = d[lane];
        const sg_u4 h = v.hdr[id];
        k.size = h.y; k.sig = h.w; k.first = v.lits[h.x]; k.last = v.lits[h.x + h.y - 1];
    }
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        SgKey o;
        o.size = __shfl_sync(0xffffffffu, k.size, j); o.first = __shfl_sync(0xffffffffu, k.first, j);
        o.last = __shfl_sync(0xffffffffu, k.last, j); o.sig = __shfl_sync(0xffffffffu, k.sig, j);
        if (sg_key_lt(o, k) || (j < lane && !sg_key_lt(k, o))) rank++;
    }
    const bool moved = lane < n && rank != lane;
    if (__any_sync(0xffffffffu, moved)) { __syncwarp(); if (lane < n) d[rank] = id; }
#else
    sg_sortot_small(v, d, n);
#endif
}
/* the rest, one thread per list: very short lists (a warp would idle; gate circuits have millions of them) and lists of
 * more than 24 entries (libstdc++'s introsort restated step for step) */
#define SG_SOT_WIDE_MAX 512
SG_HDN inline void sg_sortot_long_item(const SgView &v, uint32_t lit)
{
    const int n = (int)v.ot_len[lit];
    if (n > 24 && n <= SG_SOT_WIDE_MAX && v.sot_wl) { v.sot_wl[sg_atomic_add(&v.sc->wl_count[1], 1u)] = lit; return; }
    if (n <= (int)v.sortot_warp_min || n > 24) sg_sortot_item(v, lit);
}

/* ------------------------------------------------------------------------------------------ */
/* LCVE election (lcve.cpp:52-83, depFreeze :121-145) as a deterministic parallel greedy:       */
/* a candidate is decided once every earlier-ranked clause neighbour is decided.                */
/* ------------------------------------------------------------------------------------------ */
SG_HD bool sg_freezable(const SgView &v, uint32_t var)
{   /* lcve.cpp:137 */
    return v.occ[2 * var] < v.pmax || v.occ[2 * var + 1] < v.nmax;
}
SG_HD bool sg_break_cond(const SgView &v, uint32_t var)
{   /* lcve.cpp:70,73 */
    const uint32_t maxoccurs = (uint32_t)v.o.lcve_max_occurs;
    const uint32_t ps = v.occ[2 * var], ns = v.occ[2 * var + 1];
    if (ps > maxoccurs || ns > maxoccurs) return true;
    return v.ot_len[2 * var] >= v.pmax && v.ot_len[2 * var + 1] >= v.nmax;
}

/* election word of a variable: rank in the sorted candidate order (low 29 bits) | status << 29 */
#define SG_ER_RANK(w)  ((w) & 0x1FFFFFFFu)
#define SG_ER_STAT(w)  ((w) >> 29)
#define SG_ER_MAKE(st, r) (((uint32_t)(st) << 29) | (r))

SG_HD uint32_t sg_er_load(const SgView &v, uint32_t var) { return *(volatile uint32_t *)&v.rank[var]; }
SG_HD void sg_er_set(const SgView &v, uint32_t var, uint32_t st)
{
    const uint32_t w = v.rank[var];
#if defined(__CUDA_ARCH__)
    __threadfence();
#endif
    *(volatile uint32_t *)&v.rank[var] = SG_ER_MAKE(st, SG_ER_RANK(w));
}
/* UNDECIDED -> FROZEN, and only that transition (an acting candidate freezing a neighbour) */
SG_HD void sg_er_freeze(const SgView &v, uint32_t var, uint32_t w_undecided)
{
#if defined(__CUDA_ARCH__)
    atomicCAS(&v.rank[var], w_undecided, SG_ER_MAKE(SG_E_FROZEN, SG_ER_RANK(w_undecided)));
#else
    if (v.rank[var] == w_undecided) v.rank[var] = SG_ER_MAKE(SG_E_FROZEN, SG_ER_RANK(w_undecided));
#endif
}

/* static classification, one thread per variable (lcve.cpp:66-70) */
SG_HD void sg_elect_prep_item(const SgView &v, uint32_t var)
{
    const uint32_t ps = v.occ[2 * var], ns = v.occ[2 * var + 1];
    const uint32_t r = SG_ER_RANK(v.rank[var]);
    uint32_t st;
    if ((v.ifrozen && v.ifrozen[var]) || v.state[var] || (!ps && !ns)) st = SG_E_SKIP;
    else if (!sg_freezable(v, var)) {
        /* can never be frozen, and ot sizes >= occurrences >= mu: the loop ends here if reached */
        st = SG_E_BREAK;
        sg_atomic_min(&v.sc->brk_rank, r);
    } else st = SG_E_UNDECIDED;
    v.rank[var] = SG_ER_MAKE(st, r);
}

/* One pull step for an undecided candidate of the current rank batch (every candidate of an
 * earlier batch is decided).  Returns true when it stays undecided.  A candidate that turns
 * out to act (depFreeze succeeded, lcve.cpp:76-78) freezes its still-undecided neighbours at
 * once, so later batches never have to look at them.  One warp per candidate: the lanes take
 * the clauses of both occurrence lists in turn. */
#ifndef SG_ELECT_BATCH
#define SG_ELECT_BATCH 8        /* literals (and their election words) in flight per lane */
#endif
/* blocker (may be null): when the candidate stays undecided, *blocker receives the undecided earlier-ranked neighbour of the
 * highest rank - the one most likely to be decided last; until its election word changes, looking again is pointless */
/* GW = lanes that share one candidate: 32 (one candidate per warp), or 8 in the barrier-free election, where four
 * candidates are looked at side by side - the function waits on a chain of dependent fetches, so four chains per warp in
 * flight instead of one is what counts; every lane of a group must call it (the groups of a warp may diverge). */
template <int GW> SG_HDN inline bool sg_elect_eval_g(const SgView &v, uint32_t cand, uint32_t *blocker)
{
#if defined(__CUDA_ARCH__)
    SgLane L;
    L.lane = threadIdx.x & (GW - 1); L.width = GW;
    const uint32_t gmask = GW == 32 ? 0xffffffffu : (((1u << (GW & 31)) - 1u) << ((threadIdx.x & 31u) & ~(uint32_t)(GW - 1)));
    const int gleader = (int)((threadIdx.x & 31u) & ~(uint32_t)(GW - 1));
#define SG_G_BCAST(x) __shfl_sync(gmask, (x), gleader)
#define SG_G_WOR(x) __reduce_or_sync(gmask, (x))
#else
    const SgLane L = sg_lane_ctx();
#define SG_G_BCAST(x) (x)
#define SG_G_WOR(x) (x)
#endif
    uint32_t bl_key = 0, bl_var = 0;                          /* rank + 1 of this lane's best blocker */
    /* the list headers are fetched together with the election word, not after it: one memory round trip less on the
     * chain word -> lists -> headers -> literals -> neighbours' words that bounds this function */
    const int np = (int)v.ot_len[2 * cand], nn = (int)v.ot_len[2 * cand + 1];
    const uint32_t *dp = sg_list(v, 2 * cand), *dn = sg_list(v, 2 * cand + 1);
    const uint32_t brk = *(volatile uint32_t *)&v.sc->brk_rank;
    const uint32_t wc = SG_G_BCAST(sg_er_load(v, cand));
    if (SG_ER_STAT(wc) != SG_E_UNDECIDED) return false;       /* frozen by a push since it was queued */
    const uint32_t rc = SG_ER_RANK(wc);
    if (rc > SG_G_BCAST(brk)) { if (L.lane == 0) sg_er_set(v, cand, SG_E_SKIP); return false; }
    const int maxcsize = v.o.lcve_clause_max;
    bool blocked = false, frozen = false, failp = false, failn = false;
    for (int q = (int)L.lane; q < np + nn; q += (int)L.width) {
        const bool neg = q >= np;
        const sg_u4 h = v.hdr[neg ? dn[q - np] : dp[q]];
        if (h.z & SG_DELETED) continue;
        const int size = (int)h.y;
        if (size > maxcsize) { if (neg) failn = true; else failp = true; }
        const uint32_t *l = v.lits + h.x;
        for (int base = 0; base < size; base += SG_ELECT_BATCH) {
            /* literals first, then their election words: independent loads in flight together */
            uint32_t other[SG_ELECT_BATCH], w[SG_ELECT_BATCH];
#pragma unroll
            for (int k = 0; k < SG_ELECT_BATCH; ++k) other[k] = base + k < size ? l[base + k] : 2 * cand;
#pragma unroll
            for (int k = 0; k < SG_ELECT_BATCH; ++k) w[k] = SG_ABS(other[k]) == cand ? 0xFFFFFFFFu : sg_er_load(v, SG_ABS(other[k]));
#pragma unroll
            for (int k = 0; k < SG_ELECT_BATCH; ++k) {
                if (w[k] == 0xFFFFFFFFu || SG_ER_RANK(w[k]) > rc) continue;
                const uint32_t s = SG_ER_STAT(w[k]);
                if (s == SG_E_UNDECIDED) {
                    blocked = true;
                    if (SG_ER_RANK(w[k]) + 1 > bl_key) { bl_key = SG_ER_RANK(w[k]) + 1; bl_var = SG_ABS(other[k]); }
                }
                else if (s == SG_E_BOTH || (s == SG_E_PONLY && !SG_SIGN(other[k]))) frozen = true;
            }
        }
    }
    const uint32_t flags = SG_G_WOR((blocked ? 1u : 0u) | (frozen ? 2u : 0u) | (failp ? 4u : 0u) | (failn ? 8u : 0u));
    if (flags & 2u) { if (L.lane == 0) sg_er_set(v, cand, SG_E_FROZEN); return false; }
    if (flags & 1u) {
        if (blocker) {
#if defined(__CUDA_ARCH__)
            const uint32_t best = __reduce_max_sync(gmask, bl_key);
            *blocker = __shfl_sync(gmask, bl_var, __ffs((int)(__ballot_sync(gmask, bl_key == best) & gmask)) - 1);
#else
            *blocker = bl_var;
#endif
        }
        return true;
    }
    uint32_t st;
    if (sg_break_cond(v, cand)) { st = SG_E_BREAK; if (L.lane == 0) sg_atomic_min(&v.sc->brk_rank, rc); }
    else if (flags & 4u) st = SG_E_NONE;
    else if (flags & 8u) st = SG_E_PONLY;
    else st = SG_E_BOTH;
    /* a concurrent push can only come from an earlier-ranked acting neighbour; had it still been
     * undecided when we read it we would be blocked, had it acted we would be frozen: the status
     * word is ours to write */
    if (L.lane == 0) sg_er_set(v, cand, st);
    if (st == SG_E_BOTH || st == SG_E_PONLY) {
        const int total = st == SG_E_BOTH ? np + nn : np;
        for (int q = (int)L.lane; q < total; q += (int)L.width) {
            const sg_u4 h = v.hdr[q >= np ? dn[q - np] : dp[q]];
            if (h.z & SG_DELETED) continue;
            const uint32_t *l = v.lits + h.x;
            for (int base = 0; base < (int)h.y; base += SG_ELECT_BATCH) {
                uint32_t u[SG_ELECT_BATCH], w[SG_ELECT_BATCH];
#pragma unroll
                for (int k = 0; k < SG_ELECT_BATCH; ++k) u[k] = base + k < (int)h.y ? SG_ABS(l[base + k]) : cand;
#pragma unroll
                for (int k = 0; k < SG_ELECT_BATCH; ++k) w[k] = u[k] == cand ? 0xFFFFFFFFu : sg_er_load(v, u[k]);
#pragma unroll
                for (int k = 0; k < SG_ELECT_BATCH; ++k)
                    if (w[k] != 0xFFFFFFFFu && SG_ER_STAT(w[k]) == SG_E_UNDECIDED && SG_ER_RANK(w[k]) > rc) sg_er_freeze(v, u[k], w[k]);
            }
        }
    }
    return false;
}
#undef SG_G_BCAST
#undef SG_G_WOR
SG_HDN inline bool sg_elect_eval(const SgView &v, uint32_t cand, uint32_t *blocker) { return sg_elect_eval_g<32>(v, cand, blocker); }

SG_HDN inline bool sg_elect_round_item(const SgView &v, uint32_t cand) { return sg_elect_eval(v, cand, nullptr); }

/* function-table core marks (lcve.cpp:85-102): the first 12 variables of the freeze stack.
 * ONE thread; `acting` holds the acting candidates in election order (bit31 = P only). */
SG_HDN inline void sg_marks_serial(const SgView &v, uint32_t n_acting)
{
    SgScalars *sc = v.sc;
    for (uint32_t i = 0; i < sc->n_core; ++i) v.marks[sc->core[i]] = -1;   /* clearMapFrozen, solve.hpp:337-345 */
    sc->n_core = 0;
    if (!v.o.ve_fun_en) return;
    uint32_t core[12];
    uint32_t nc = 0;
    for (uint32_t a = 0; a < n_acting && nc < 12; ++a) {
        const uint32_t cand = v.acting[a] & 0x7fffffffu;
        const uint32_t sides = (v.acting[a] >> 31) ? 1u : 2u;
        for (uint32_t side = 0; side < sides && nc < 12; ++side) {
            const uint32_t lit = 2 * cand + side;
            const int n = (int)v.ot_len[lit];
            const uint32_t *d = sg_list(v, lit);
            for (int i = 0; i < n && nc < 12; ++i) {
                const uint32_t c = d[i];
                if (sg_deleted(v, c)) continue;
                const uint32_t *l = SG_LITS(v, c);
                const int size = SG_SIZE(v, c);
                for (int k = 0; k < size && nc < 12; ++k) {
                    const uint32_t u = SG_ABS(l[k]);
                    if (u == cand || !sg_freezable(v, u)) continue;
                    bool seen = false;
                    for (uint32_t q = 0; q < nc; ++q) if (core[q] == u) { seen = true; break; }
                    if (!seen) core[nc++] = u;
                }
            }
        }
    }
    for (uint32_t q = 0; q < nc; ++q) { v.marks[core[q]] = (int8_t)q; sc->core[q] = core[q]; }
    sc->n_core = nc;
}

/* ------------------------------------------------------------------------------------------ */
/* extract (simplify.cpp:118-133): CLAUSE records of the solver's clause arena -> SCLAUSE records */
/* ------------------------------------------------------------------------------------------ */
#define SG_CM_HDR 4u          /* header buckets of a CLAUSE (clause.hpp:47-59: sizeof == 24, two of them literals) */

/* buckets the clause will take in the SCNF arena: 0 when cm.deleted(ref) (simplify.cpp:124) */
#define SG_EXTRACT_BAD 0xFFFFFFFFu   /* the ref (or the record's size word) points outside the cm_buckets words of the database */
SG_HD uint32_t sg_extract_size(const uint32_t *cm, uint64_t cm_buckets, const uint8_t *stencil, uint64_t stencil_size, uint64_t ref)
{
    if (ref + SG_CM_HDR > cm_buckets) return SG_EXTRACT_BAD;
    if (ref < stencil_size && stencil[ref]) return 0;
    if (ref + SG_CM_HDR + cm[ref + 1] > cm_buckets) return SG_EXTRACT_BAD;
    return 3u + cm[ref + 1];
}

/* SCLAUSE(const CLAUSE&) + calcSig + SORT (sclause.hpp:61-75,158-162; sort.hpp:44-52): dst = {w0, sig, size, lits} */
SG_HD void sg_extract_write(const uint32_t *cm, uint64_t ref, uint32_t *dst)
{
    const uint32_t f = cm[ref], size = cm[ref + 1], lbd = cm[ref + 3];
    const bool learnt = ((f >> 8) & 0xFFu) != 0;                      /* bool _l: byte 1 of the first bucket */
    const uint32_t used = (f >> 24) & 3u;                             /* CL_ST _used: byte 3 */
    const uint32_t *src = cm + ref + SG_CM_HDR;
    uint32_t sig = 0;
    uint32_t *out = dst + 3;
    if (size <= 8) {
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = k < (int)size ? src[k] : 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < (int)size) sig |= SG_HASH(a[k]);
        /* odd-even transposition network: data-independent, stays in registers */
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = r & 1; k + 1 < 8; k += 2) {
                const uint32_t lo = a[k] < a[k + 1] ? a[k] : a[k + 1], hi = a[k] < a[k + 1] ? a[k + 1] : a[k];
                a[k] = lo; a[k + 1] = hi;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < (int)size) out[k] = a[k];
    } else {
        for (uint32_t k = 0; k < size; ++k) { const uint32_t l = src[k]; sig |= SG_HASH(l); out[k] = l; }
        if (size <= 32) {
            for (uint32_t i = 1; i < size; ++i) {
                const uint32_t x = out[i];
                uint32_t j = i;
                while (j > 0 && out[j - 1] > x) { out[j] = out[j - 1]; --j; }
                out[j] = x;
            }
        } else sg_heapsort_u32(out, (int)size);
    }
    dst[0] = learnt ? (SG_LEARNT | (used << SG_USAGE_SH) | ((lbd & 0x03FFFFFFu) << SG_LBD_SH)) : 0u;
    dst[1] = sig;
    dst[2] = size;
}

/* addClause(SCLAUSE&) (sclause.cpp:22-62): one CLAUSE record (clause.hpp:47-59,77-92) from an exported SCLAUSE */
/* vmap (may be null): variable mapping of map(true) (map.cpp:24-46, map.hpp:52-76,195-205): the clause is built with
 * mapLit(lit) = 2 vmap[var] + sign instead of the literal itself (every literal of the simplified formula is unassigned and
 * active, so mapLit's branch for root-assigned variables is never taken) */
SG_HD void sg_cm_write(const uint32_t *src, uint32_t *dst, int lbd_tier1, const uint32_t *vmap = nullptr)
{
    const uint32_t w0 = src[0], size = src[2];
    const bool learnt = (w0 & SG_ST_MASK) == SG_LEARNT;
    const uint32_t lbd = w0 >> SG_LBD_SH, used = (w0 >> SG_USAGE_SH) & 3u;
    const bool keep = !(learnt && size > 2 && (int)lbd > lbd_tier1);
    /* byte 0: _r _m _s _h _k _b; byte 1: _l; byte 2: _v; byte 3: _used */
    dst[0] = (keep ? 1u << 4 : 0u) | (size == 2 ? 1u << 5 : 0u) | (learnt ? 1u << 8 : 0u) | (learnt ? used << 24 : 0u);
    dst[1] = size;
    dst[2] = 2u;                                  /* _pos */
    dst[3] = learnt ? lbd : 0u;
    if (vmap) for (uint32_t k = 0; k < size; ++k) dst[SG_CM_HDR + k] = 2u * vmap[SG_ABS(src[3 + k])] + SG_SIGN(src[3 + k]);
    else for (uint32_t k = 0; k < size; ++k) dst[SG_CM_HDR + k] = src[3 + k];
}

/* ------------------------------------------------------------------------------------------ */
/* rebuildWT (watch.cpp:135-157): sortClause + ATTACH_TWO_WATCHES for every clause of the new database  */
/* ------------------------------------------------------------------------------------------ */
/* sortClause(c, start, size, satonly) (clause.cpp:116-166) on the literals of a CLAUSE record; every assigned variable is
 * assigned at the root here (the pass runs at decision level 0), so l2dl() is 0 for all of them */
SG_HD int8_t sg_sort_clause_from(uint32_t *lits, int start, int size, bool satonly, const int8_t *value)
{
    const uint32_t x = lits[start];
    int8_t xval = value[x];
    if (xval < 0 || (xval && satonly)) return xval;
    uint32_t best = x;
    int pos = 0;
    for (int i = start + 1; i < size; ++i) {
        const uint32_t y = lits[i];
        const int8_t yval = value[y];
        if (yval < 0 || (yval && satonly)) { best = y; pos = i; xval = yval; break; }
        bool better;
        if (!xval && yval > 0) better = true;
        else if (xval > 0 && !yval) better = false;
        else better = false;                                  /* equal values: the levels decide, and they are all 0 */
        if (better) { best = y; pos = i; xval = yval; }
    }
    if (!pos) return xval;
    lits[start] = best;
    lits[pos] = x;
    return xval;
}
SG_HD void sg_sort_clause(uint32_t *lits, int size, const int8_t *value)
{   /* clause.cpp:168-173 */
    const int8_t val = sg_sort_clause_from(lits, 0, size, false, value);
    if (size > 2) (void)sg_sort_clause_from(lits, 1, size, val != 0, value);       /* `const bool& satonly` bound to the LIT_ST */
}
/* the group of a clause in the order rebuildWT attaches them (watch.cpp:139-156): code bit 0 = binaries first,
 * bit 1 = learnt binaries before original ones */
SG_HD uint32_t sg_wt_group(bool binary, bool learnt, uint32_t code)
{
    if (!(code & 1u)) return learnt ? 1u : 0u;
    if (binary) return (code & 2u) ? (learnt ? 0u : 1u) : (learnt ? 1u : 0u);
    return learnt ? 3u : 2u;
}

/* ------------------------------------------------------------------------------------------ */
/* SUB: self-subsuming strengthening + backward subsumption (subsume.hpp:26-218, subsume.cpp:45-76) */
/* ------------------------------------------------------------------------------------------ */
SG_HD bool sg_sig_sub(uint32_t A, uint32_t B) { return !(A & ~B); }                 /* simplify.hpp:119 */
SG_HD bool sg_sig_selfsub(uint32_t A, uint32_t B)
{   /* simplify.hpp:121-125 */
    const uint32_t Bt = B | ((B & 0xAAAAAAAAu) >> 1) | ((B & 0x55555555u) << 1);
    return !(A & ~Bt);
}

SG_HD bool sg_sub_lits(const uint32_t *d1, int n1, const uint32_t *d2, int n2)
{   /* subsume.hpp:38-64 */
    int i = 0, j = 0, sub = 0;
    while (i < n1 && j < n2) {
        const uint32_t a = d1[i], b = d2[j];
        if (a < b) i++;
        else if (b < a) j++;
        else { sub++; i++; j++; }
    }
    return sub == n1;
}

SG_HD bool sg_selfsub_lits(uint32_t x, uint32_t fx, const uint32_t *d1, int n1, const uint32_t *d2, int n2)
{   /* subsume.hpp:93-137 */
    int i = 0, j = 0, sub = 0;
    bool self = false;
    while (i < n1 && j < n2) {
        const uint32_t a = d1[i], b = d2[j];
        if (a == fx) i++;
        else if (b == x) { self = true; j++; }
        else if (a < b) i++;
        else if (b < a) j++;
        else { sub++; i++; j++; }
    }
    if (sub + 1 == n1) {
        if (self) return true;
        while (j < n2) { if (d2[j] == x) return true; j++; }
    }
    return false;
}

/* strengthen (subsume.cpp:45-76); lane 0 rewrites the clause, every lane tracks `seq` */
SG_HD void sg_strengthen(const SgView &v, const SgLane &L, uint32_t c, uint32_t me, uint32_t eidx, uint32_t &seq)
{
    const int size = SG_SIZE(v, c);
    sg_wsync();
    if (L.lane == 0) {
        uint32_t *l = SG_LITS(v, c);
        uint32_t sig = 0;
        int j = 0;
        for (int k = 0; k < size; ++k) {
            const uint32_t lit = l[k];
            if (lit != me) { l[j++] = lit; sig |= SG_HASH(lit); }
        }
        uint32_t w0 = SG_W0(v, c);
        if (size - 1 == 1) { uint32_t s2 = seq; sg_push_unit(v, eidx, s2, l[0]); }
        else if (w0 & SG_LEARNT) {
            /* bumpShrunken (solve.hpp:725-737) */
            const int old_lbd = (int)(w0 >> SG_LBD_SH);
            if (old_lbd > v.o.lbd_tier1) {
                const int sz = size - 2;
                const int new_lbd = sz < old_lbd ? sz : old_lbd;
                if (new_lbd < old_lbd) w0 = (w0 & 0xFu) | (1u << SG_USAGE_SH) | ((uint32_t)new_lbd << SG_LBD_SH);
            }
        }
        sg_u4 h = v.hdr[c];
        h.y = (uint32_t)(size - 1); h.z = w0 | SG_MOLTEN; h.w = sig;       /* cand.melt(), subsume.hpp:170 */
        v.hdr[c] = h;
    }
    if (size - 1 == 1) seq++;
    sg_wsync();
}

#define SG_SUB_PRECHECK_MAX 192          /* pairs up to which the side-by-side test runs first (six steps of the warp; measured: 96 / 192 / all) */
SG_HDN inline void sg_self_sub_half(const SgView &v, const SgLane &L, uint32_t x, uint32_t fx, uint32_t eidx, uint32_t &seq,
                                    uint32_t &n_str, uint32_t &n_sub)
{   /* one loop of self_sub_x (subsume.hpp:186-199): clauses of `x` against the list of `fx` */
    const int maxsz = v.o.sub_clause_max;
    uint32_t *me = sg_list(v, x);
    const int nme = (int)v.ot_len[x];
    const uint32_t *ot = sg_list(v, fx);
    const int not_ = (int)v.ot_len[fx];
#if defined(__CUDA_ARCH__)
    if (nme <= 32 && not_ <= 32) {
        /* Both lists fit a warp (the common case): lane j keeps the header of ot[j] and of me[j] in registers, so one
         * step of the loop is a broadcast of clause i's header, one predicate per lane and a ballot - the same tests in
         * the same order as the general loop below, without re-fetching a dozen headers per clause. */
        const int lane = (int)L.lane;
        const uint32_t oc = lane < not_ ? ot[lane] : 0u;
        const sg_u4 oh = lane < not_ ? v.hdr[oc] : make_uint4(0, 0, SG_DELETED, 0);
        const uint32_t mc = lane < nme ? me[lane] : 0u;
        sg_u4 mh = lane < nme ? v.hdr[mc] : make_uint4(0, 0, SG_DELETED, 0);
        {
            /* Would the loop below change anything?  It strengthens clause i when a clause of the other list hits it and deletes
             * it when an earlier clause of its own list subsumes it; as long as neither happens the state stays what it was at
             * the start, so "no pair hits in the initial state" means the loop does nothing.  The pairs are tested side by
             * side, one per lane, read-only (the loop itself keeps 32 - |list| lanes idle in every step, and on gate circuits,
             * with two to six clauses per list, hits are rare): the usual case ends here. */
            const uint32_t big = __ballot_sync(0xffffffffu, lane < nme && (int)mh.y > maxsz);
            const int ilim = big ? __ffs((int)big) - 1 : nme;                 /* the loop stops at the first clause above the size limit */
            const int W2 = not_ + ilim, tot = ilim * W2;
            bool any = tot > SG_SUB_PRECHECK_MAX;                             /* long lists (k-SAT with planted subsumption: hits are the rule) go straight to the loop */
            const float rcp = 1.0f / (float)(W2 > 0 ? W2 : 1);               /* q / W2 for q < 256, W2 <= 64: (q + 0.5) / W2 is at least 1/128 away from an integer */
            for (int q0 = 0; q0 < tot && !any; q0 += 32) {
                const int q = q0 + lane;
                const bool valid = q < tot;
                const int i = valid ? (int)(((float)q + 0.5f) * rcp) : 0, r = valid ? q - i * W2 : 0;
                const bool cross = r < not_;                                  /* (clause i, clause r of the other list) or (clause i, earlier clause of its list) */
                const int src = cross ? r : r - not_;
                sg_u4 ch, so, sm;
                ch.x = __shfl_sync(0xffffffffu, mh.x, i); ch.y = __shfl_sync(0xffffffffu, mh.y, i);
                ch.z = __shfl_sync(0xffffffffu, mh.z, i); ch.w = __shfl_sync(0xffffffffu, mh.w, i);
                so.x = __shfl_sync(0xffffffffu, oh.x, src); so.y = __shfl_sync(0xffffffffu, oh.y, src);
                so.z = __shfl_sync(0xffffffffu, oh.z, src); so.w = __shfl_sync(0xffffffffu, oh.w, src);
                sm.x = __shfl_sync(0xffffffffu, mh.x, src); sm.y = __shfl_sync(0xffffffffu, mh.y, src);
                sm.z = __shfl_sync(0xffffffffu, mh.z, src); sm.w = __shfl_sync(0xffffffffu, mh.w, src);
                bool hit = false;
                if (valid && !(ch.z & SG_DELETED)) {
                    const int candsz = (int)ch.y;
                    const uint32_t *cl = v.lits + ch.x;
                    if (cross) {
                        const int subsize = (int)so.y;
                        hit = !(so.z & SG_DELETED) && subsize <= candsz && subsize > 1 && sg_sig_selfsub(so.w, ch.w) &&
                              sg_selfsub_lits(x, fx, v.lits + so.x, subsize, cl, candsz);
                    } else if (src < i) {
                        const int ssz = (int)sm.y;
                        hit = !(sm.z & SG_DELETED) && !((ch.z & SG_MOLTEN) && ssz > candsz) && ssz > 1 && sg_sig_sub(sm.w, ch.w) &&
                              sg_sub_lits(v.lits + sm.x, ssz, cl, candsz);
                    }
                }
                any = __ballot_sync(0xffffffffu, hit) != 0;
            }
            if (!any) goto update_ol;
        }
        for (int i = 0; i < nme; ++i) {
            const uint32_t c = __shfl_sync(0xffffffffu, mc, i);
            sg_u4 ch;
            ch.x = __shfl_sync(0xffffffffu, mh.x, i); ch.y = __shfl_sync(0xffffffffu, mh.y, i);
            ch.z = __shfl_sync(0xffffffffu, mh.z, i); ch.w = __shfl_sync(0xffffffffu, mh.w, i);
            if ((int)ch.y > maxsz) break;
            if (ch.z & SG_DELETED) continue;
            const uint32_t *cl = v.lits + ch.x;
            {   /* selfsubsume (subsume.hpp:158-177) */
                const int candsz = (int)ch.y;
                bool hit = false;
                if (lane < not_) {
                    const int subsize = (int)oh.y;
                    if (subsize > candsz) hit = true;
                    else if (!(oh.z & SG_DELETED))
                        hit = subsize > 1 && sg_sig_selfsub(oh.w, ch.w) && sg_selfsub_lits(x, fx, v.lits + oh.x, subsize, cl, candsz);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    const int k = __ffs((int)m) - 1;
                    if ((int)__shfl_sync(0xffffffffu, oh.y, k) <= candsz) {
                        sg_strengthen(v, L, c, x, eidx, seq);
                        n_str++;
                        if (lane == i) mh = v.hdr[c];                     /* size, signature and molten flag changed */
                        ch.y = __shfl_sync(0xffffffffu, mh.y, i); ch.z = __shfl_sync(0xffffffffu, mh.z, i);
                        ch.w = __shfl_sync(0xffffffffu, mh.w, i);
                    }
                }
            }
            {   /* subsume (subsume.hpp:139-156): the first earlier clause of the list that subsumes c */
                const int candsz = (int)ch.y;
                const bool cmolten = (ch.z & SG_MOLTEN) != 0;
                bool hit = false;
                if (lane < i && !(mh.z & SG_DELETED)) {
                    const int ssz = (int)mh.y;
                    if (!(cmolten && ssz > candsz))
                        hit = ssz > 1 && sg_sig_sub(mh.w, ch.w) && sg_sub_lits(v.lits + mh.x, ssz, cl, candsz);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    const int k = __ffs((int)m) - 1;
                    const bool corg = !(ch.z & SG_ST_MASK);
                    if (lane == k && (mh.z & SG_LEARNT) && corg) { mh.z &= ~SG_ST_MASK; v.hdr[mc].z = mh.z; }   /* subsume.hpp:147-148 */
                    if (lane == i) { mh.z = (mh.z & ~SG_ST_MASK) | SG_DELETED; v.hdr[c].z = mh.z; }
                    __syncwarp();
                    n_sub++;
                }
            }
        }
        goto update_ol;
    }
#endif
    for (int i = 0; i < nme; ++i) {
        const uint32_t c = me[i];
        if (SG_SIZE(v, c) > maxsz) break;
        if (sg_deleted(v, c)) continue;
        {   /* selfsubsume (subsume.hpp:158-177): first entry that either ends the scan or strengthens */
            const int candsz = SG_SIZE(v, c);
            const uint32_t candsig = SG_SIG(v, c);
            const uint32_t *cl = SG_LITS(v, c);
            const int k = sg_find_first(L, not_, [&](int j) {
                const sg_u4 hs = v.hdr[ot[j]];
                const int subsize = (int)hs.y;
                if (subsize > candsz) return true;
                if (hs.z & SG_DELETED) return false;
                return subsize > 1 && sg_sig_selfsub(hs.w, candsig) && sg_selfsub_lits(x, fx, v.lits + hs.x, subsize, cl, candsz);
            });
            if (k >= 0 && (int)v.hdr[ot[k]].y <= candsz) {
                sg_strengthen(v, L, c, x, eidx, seq);
                n_str++;
            }
        }
        {   /* subsume (subsume.hpp:139-156) */
            const int candsz = SG_SIZE(v, c);
            const bool cmolten = sg_molten(v, c);
            const uint32_t candsig = SG_SIG(v, c);
            const uint32_t *cl = SG_LITS(v, c);
            const int k = sg_find_first(L, i, [&](int j) {
                const sg_u4 hs = v.hdr[me[j]];
                if (hs.z & SG_DELETED) return false;
                const int ssz = (int)hs.y;
                if (cmolten && ssz > candsz) return false;
                return ssz > 1 && sg_sig_sub(hs.w, candsig) && sg_sub_lits(v.lits + hs.x, ssz, cl, candsz);
            });
            if (k >= 0) {
                if (L.lane == 0) {
                    const uint32_t s = me[k];
                    const uint32_t sz0 = v.hdr[s].z;
                    if ((sz0 & SG_LEARNT) && sg_original(v, c)) v.hdr[s].z = sz0 & ~SG_ST_MASK;
                    sg_mark_deleted(v, c);
                }
                sg_wsync();
                n_sub++;
            }
        }
    }
#if defined(__CUDA_ARCH__)
update_ol:
#endif
    /* updateOL (subsume.hpp:26-36): in-place, order-preserving */
    int j = 0;
    for (int base = 0; base < nme; base += (int)L.width) {
        const int idx = base + (int)L.lane;
        uint32_t c = 0;
        bool keep = false;
        if (idx < nme) {
            c = me[idx];
            const uint32_t w0 = SG_W0(v, c);
            if (w0 & SG_MOLTEN) v.hdr[c].z = w0 & ~SG_MOLTEN;
            else keep = !(w0 & SG_DELETED);
        }
        const uint32_t m = sg_ballot(keep);
        sg_wsync();
        const int dst = j + sg_popc(m & ((1u << L.lane) - 1u));
        if (keep && dst != idx) me[dst] = c;                   /* nothing dropped so far: the entry is where it belongs */
        j += sg_popc(m);
        sg_wsync();
    }
    if (L.lane == 0 && j != nme) v.ot_len[x] = (uint32_t)j;
    sg_wsync();
}

SG_HDN inline void sg_sub_item(const SgView &v, uint32_t eidx)
{   /* subsume.cpp:32-39 */
    const SgLane L = sg_lane_ctx();
    const uint32_t var = v.elected[eidx];
    const uint32_t p = 2 * var, n = p | 1u;
    const int maxoccurs = v.o.sub_max_occurs;
    if ((int)v.ot_len[p] > maxoccurs || (int)v.ot_len[n] > maxoccurs) return;
    uint32_t seq = 0, n_str = 0, n_sub = 0;
    sg_self_sub_half(v, L, p, n, eidx, seq, n_str, n_sub);
    sg_self_sub_half(v, L, n, p, eidx, seq, n_str, n_sub);
    if (L.lane == 0) {
        if (n_str) sg_atomic_add64(&v.sc->strengthened, n_str);
        if (n_sub) sg_atomic_add64(&v.sc->subsumed, n_sub);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* BVE: decision + resolvent counting (bounded.cpp:41-160 and the gate headers)                 */
/* ------------------------------------------------------------------------------------------ */
/* Clauses keep their literals in ascending code order (code = 2*var + sign), so two clauses can be walked like two sorted
 * key lists.  At every step the smaller code is consumed, or both when they are equal (a literal the clauses share); codes
 * that differ in the sign bit only make the resolvent a tautology; the pivot's own two codes are stepped over.  The walk
 * ends when either clause is exhausted, as in the reference (simplify.hpp:181-260): what is left of the other one cannot
 * hold a shared or a clashing variable. */
SG_HD int sg_merge_count(uint32_t pivot, const uint32_t *a, int na, const uint32_t *b, int nb)
{   /* size of the resolvent on VARIABLE `pivot`, 0 for a tautology */
    int ia = 0, ib = 0, shared = 0;
#pragma unroll 1
    while (ia < na && ib < nb) {
        const uint32_t ka = a[ia], kb = b[ib];
        const bool pa = (ka >> 1) == pivot, pb = (kb >> 1) == pivot;
        if (pa | pb) { ia += pa; ib += pb; continue; }
        const uint32_t diff = ka ^ kb;
        if (diff == 1u) return 0;
        shared += diff == 0u;
        ia += ka <= kb;
        ib += kb <= ka;
    }
    return na + nb - 2 - shared;
}

SG_HD int sg_merge_emit(uint32_t pivot, const uint32_t *a, int na, const uint32_t *b, int nb, uint32_t *out, uint32_t *sig_out)
{   /* the same walk writing the resolvent and its signature; -1 for a tautology, else the resolvent size */
    int ia = 0, ib = 0, n = 0;
    uint32_t sig = 0;
    while (ia < na && ib < nb) {
        const uint32_t ka = a[ia], kb = b[ib];
        const bool pa = (ka >> 1) == pivot, pb = (kb >> 1) == pivot;
        if (pa | pb) { ia += pa; ib += pb; continue; }
        if ((ka ^ kb) == 1u) return -1;
        const uint32_t k = ka < kb ? ka : kb;
        out[n++] = k, sig |= SG_HASH(k);
        ia += ka <= kb;
        ib += kb <= ka;
    }
    for (; ia < na; ++ia) if ((a[ia] >> 1) != pivot) { out[n++] = a[ia]; sig |= SG_HASH(a[ia]); }
    for (; ib < nb; ++ib) if ((b[ib] >> 1) != pivot) { out[n++] = b[ib]; sig |= SG_HASH(b[ib]); }
    *sig_out = sig;
    return n;
}

/* pair filter shared by the count loops and the emit loops:
 * mode RESOLVE: all original pairs; SUBST: a != b; CORE: !a || !b (bounded.cpp:205,235; function.hpp:109) */
SG_HD bool sg_pair_ok(int mode, bool a, bool b)
{
    return mode == SG_VE_RESOLVE ? true : (mode == SG_VE_SUBST ? (a != b) : (!a || !b));
}

SG_HD int sg_count_lits_before(const SgView &v, const SgLane &L, uint32_t lit)
{   /* simplify.hpp:288-296 */
    const int len = (int)v.ot_len[lit];
    const uint32_t *d = sg_list(v, lit);
    int n = 0;
#pragma unroll 1
    for (int i = (int)L.lane; i < len; i += (int)L.width) if (sg_original(v, d[i])) n += SG_SIZE(v, d[i]);
    return sg_wsum(n);
}

/* countSubstituted (simplify.hpp:298-328) / countMinResolvents (function.hpp:95-126) /
 * countResolvents bounded (simplify.hpp:355-383) / countResolvents 1xm (simplify.hpp:330-353,
 * clsbefore < 0 = no clause bound).  Returns TRUE when a bound is EXCEEDED.  The pair space is
 * split over the lanes: the sums are order free, and "some prefix exceeds the clause bound" is
 * equivalent to "the total exceeds it", so the verdict is the sequential one; on TRUE only
 * nAddedCls != 0 is relied upon by the caller (bounded.cpp:105,120,133,153), which holds. */
SG_HDN inline bool sg_count_pairs(const SgView &v, const SgLane &L, int mode, uint32_t x, int clsbefore, uint32_t me,
                                  uint32_t other, int &nAddedCls, int &nAddedLits)
{
    const int rlimit = v.o.ve_clause_max;
    const int nme = (int)v.ot_len[me], noth = (int)v.ot_len[other];
    const uint32_t *dm = sg_list(v, me), *dt = sg_list(v, other);
    const long long total = (long long)nme * noth;
    const long long chunk = (long long)L.width * 2;       /* bound check every 64 pairs: a variable that blows up is dropped early */
    int cnt = 0, lits = 0;
    for (long long base = 0; base < total; base += chunk) {
        int c1 = 0, l1 = 0;
        bool over = false;
        for (int k = 0; k < 2; ++k) {
            const long long q = base + (long long)k * L.width + L.lane;
            if (q >= total) break;
            const int i = (int)(q / noth), j = (int)(q - (long long)i * noth);
            const sg_u4 hi = v.hdr[dm[i]];
            if (hi.z & SG_ST_MASK) continue;
            const sg_u4 hj = v.hdr[dt[j]];
            if (hj.z & SG_ST_MASK) continue;
            if (!sg_pair_ok(mode, (hi.z & SG_MOLTEN) != 0, (hj.z & SG_MOLTEN) != 0)) continue;
            const int rsize = sg_merge_count(x, v.lits + hi.x, (int)hi.y, v.lits + hj.x, (int)hj.y);
            if (rsize > 1) {
                c1++; l1 += rsize;
                if (rlimit && rsize > rlimit) over = true;
            }
        }
        cnt += sg_wsum(c1);
        lits += sg_wsum(l1);
        const bool anyover = sg_ballot(over) != 0;
        if (anyover || (clsbefore >= 0 && cnt > clsbefore)) {
            nAddedCls = cnt > 0 ? cnt : 1;
            nAddedLits = lits;
            return true;
        }
    }
    nAddedCls = cnt;
    nAddedLits = lits;
    if (clsbefore >= 0 && v.o.ve_lbound_en) {
        const int before = sg_count_lits_before(v, L, me) + sg_count_lits_before(v, L, other);
        if (nAddedLits > before) return true;
    }
    return false;
}

/* clears the molten flag of the list entries selected by pred; one entry per lane */
template <class P> SG_HD void sg_freeze_if(const SgView &v, const SgLane &L, uint32_t lit, P pred)
{
    const int n = (int)v.ot_len[lit];
    const uint32_t *d = sg_list(v, lit);
    sg_wsync();
#pragma unroll 1
    for (int i = (int)L.lane; i < n; i += (int)L.width) {
        const sg_u4 h = v.hdr[d[i]];
        if ((h.z & SG_MOLTEN) && pred(h)) v.hdr[d[i]].z = h.z & ~SG_MOLTEN;
    }
    sg_wsync();
}
SG_HD void sg_freeze_binaries(const SgView &v, const SgLane &L, uint32_t lit)
{   /* simplify.hpp:262-269 */
    sg_freeze_if(v, L, lit, [](const sg_u4 &h) { return !(h.z & SG_ST_MASK) && h.y == 2; });
}
SG_HD void sg_freeze_all(const SgView &v, const SgLane &L, uint32_t lit)
{   /* simplify.hpp:271-278 */
    sg_freeze_if(v, L, lit, [](const sg_u4 &h) { return !(h.z & SG_ST_MASK); });
}
SG_HD void sg_melt1(const SgView &v, const SgLane &L, uint32_t c)
{
    sg_wsync();
    if (L.lane == 0) sg_melt(v, c);
    sg_wsync();
}
SG_HD void sg_freeze1(const SgView &v, const SgLane &L, uint32_t c)
{
    sg_wsync();
    if (L.lane == 0) sg_freeze(v, c);
    sg_wsync();
}

/* equivalence / inverter gate (equivalence.hpp:103-161).  find_sfanin melts original binaries
 * of P until it has seen two; with two or more the gate is rejected and freezeBinaries(P) clears
 * them again, so only the one-binary case leaves a mark. */
SG_HDN inline uint32_t sg_find_bn_gate(const SgView &v, const SgLane &L, uint32_t p, uint32_t n)
{
    const int np = (int)v.ot_len[p], nn = (int)v.ot_len[n];
    const uint32_t *dp = sg_list(v, p), *dn = sg_list(v, n);
    if (SG_SIZE(v, dp[0]) > 2 || SG_SIZE(v, dn[0]) > 2) return 0;
    int nbin = 0, first_bin = -1;
    for (int base = 0; base < np; base += (int)L.width) {
        const int idx = base + (int)L.lane;
        bool b = false;
        if (idx < np) { const sg_u4 h = v.hdr[dp[idx]]; b = !(h.z & SG_ST_MASK) && h.y == 2; }
        const uint32_t m = sg_ballot(b);
        if (m && first_bin < 0) first_bin = base + sg_ffs(m) - 1;
        nbin += sg_popc(m);
        if (nbin > 1) break;
    }
    if (nbin != 1) return 0;      /* 0: nothing melted; >1: melted then frozen again (net nothing) */
    const uint32_t cb = dp[first_bin];
    const uint32_t *lb = SG_LITS(v, cb);
    const uint32_t def = SG_FLIP(lb[0] ^ lb[1] ^ p);
    uint32_t first = def, second = n;
    if (second < first) { first = second; second = def; }
    const int k = sg_find_first(L, nn, [&](int i) {
        const sg_u4 h = v.hdr[dn[i]];
        if ((h.z & SG_ST_MASK) || h.y != 2) return false;
        const uint32_t *l = v.lits + h.x;
        return l[0] == first && l[1] == second;
    });
    if (k < 0) return 0;          /* the binary of P was melted and frozen again: net nothing */
    sg_wsync();
    if (L.lane == 0) { sg_melt(v, cb); sg_melt(v, dn[k]); }
    sg_wsync();
    return def;
}

/* AND/OR gate (and.hpp:27-111); `gate_in` is this variable's scratch slice */
SG_HDN inline bool sg_find_ao_gate(const SgView &v, const SgLane &L, uint32_t dx, uint32_t fx, uint32_t *gate_in, int orgCls,
                                   int &nAddedCls, int &nAddedLits)
{
    const int nd = (int)v.ot_len[dx], nf = (int)v.ot_len[fx];
    const uint32_t *dd = sg_list(v, dx), *df = sg_list(v, fx);
    if (SG_SIZE(v, dd[0]) > 2 || SG_SIZE(v, df[nf - 1]) < 3) return false;
    uint32_t sig = 0;
    int nout = 0;
    sg_wsync();
    for (int base = 0; base < nd; base += (int)L.width) {        /* find_fanin (and.hpp:27-44) */
        const int idx = base + (int)L.lane;
        bool b = false;
        uint32_t imp = 0, c = 0;
        if (idx < nd) {
            c = dd[idx];
            const sg_u4 h = v.hdr[c];
            b = !(h.z & SG_ST_MASK) && h.y == 2;
            if (b) { const uint32_t *l = v.lits + h.x; imp = SG_FLIP(l[0] ^ l[1] ^ dx); }
        }
        const uint32_t m = sg_ballot(b);
        if (b) {
            gate_in[nout + sg_popc(m & ((1u << L.lane) - 1u))] = imp;
            v.hdr[c].z |= SG_MOLTEN;
        }
        sig |= sg_wor(b ? SG_HASH(imp) : 0u);
        nout += sg_popc(m);
    }
    sg_wsync();
    if (nout > 1) {
        if (L.lane == 0) { gate_in[nout] = fx; sg_sort_lits(gate_in, nout + 1); }
        nout++;
        sig |= SG_HASH(fx);
        sg_wsync();
        const uint32_t x = SG_ABS(dx);
        const int k = sg_find_first(L, nf, [&](int i) {
            const sg_u4 h = v.hdr[df[i]];
            if ((h.z & SG_ST_MASK) || (int)h.y != nout || !sg_sig_sub(h.w, sig)) return false;
            const uint32_t *l = v.lits + h.x;
            for (int q = 0; q < nout; ++q) if (l[q] != gate_in[q]) return false;
            return true;
        });
        if (k >= 0) {
            const uint32_t c = df[k];
            sg_melt1(v, L, c);
            nAddedCls = 0, nAddedLits = 0;
            if (!sg_count_pairs(v, L, SG_VE_SUBST, x, orgCls, dx, fx, nAddedCls, nAddedLits)) return true;
            sg_freeze1(v, L, c);
        }
    }
    sg_freeze_binaries(v, L, dx);
    return false;
}

/* ITE gate (ifthenelse.hpp:26-125) */
SG_HD void sg_swap(uint32_t &a, uint32_t &b) { const uint32_t t = a; a = b; b = t; }

/* the original ternary clause {a, b, c}, looked for in the shortest of the three occurrence lists (the first of the shortest
 * ones in the order a, b, c - that is what the reference's two conditional swaps select, and with duplicate clauses in the
 * formula the list decides which copy is found; ifthenelse.hpp:26-45).  Returns the clause id or 0xFFFFFFFF. */
SG_HDN inline uint32_t sg_find_ternary(const SgView &v, const SgLane &L, uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t via = a;
    if (v.ot_len[b] < v.ot_len[via]) via = b;
    if (v.ot_len[c] < v.ot_len[via]) via = c;
    /* stored clauses are ascending: compare against (lo, mid, hi) of the three codes */
    const uint32_t ab_lo = a < b ? a : b, ab_hi = a < b ? b : a;
    const uint32_t lo = c < ab_lo ? c : ab_lo, hi = c > ab_hi ? c : ab_hi, mid = a ^ b ^ c ^ lo ^ hi;
    const int n = (int)v.ot_len[via];
    const uint32_t *d = sg_list(v, via);
    const int k = sg_find_first(L, n, [&](int i) {
        const sg_u4 h = v.hdr[d[i]];
        if (h.y != 3u || (h.z & (SG_MOLTEN | SG_ST_MASK))) return false;
        const uint32_t *l = v.lits + h.x;
        return l[0] == lo && l[1] == mid && l[2] == hi;
    });
    return k < 0 ? 0xFFFFFFFFu : d[k];
}

SG_HDN inline bool sg_find_ite_gate(const SgView &v, const SgLane &L, uint32_t dx, uint32_t fx, int orgCls, int &nAddedCls,
                                    int &nAddedLits)
{
    const int nd = (int)v.ot_len[dx];
    const uint32_t *dd = sg_list(v, dx);
    if (SG_SIZE(v, dd[nd - 1]) == 2) return false;
    const uint32_t var = SG_ABS(dx);
#if defined(__CUDA_ARCH__)
    if (nd <= 32) {
        /* the loops below only ever stop at original ternary clauses: with fewer than two of them in the list (one
         * look per lane) there is nothing to find - the usual case outside gate circuits */
        bool tern = false;
        if ((int)L.lane < nd) { const sg_u4 h = v.hdr[dd[L.lane]]; tern = !(h.z & SG_ST_MASK) && h.y == 3u; }
        if (sg_popc(sg_ballot(tern)) < 2) return false;
    }
#endif
    for (int i = 0; i < nd; ++i) {
        const uint32_t ci = dd[i];
        if (!(sg_original(v, ci) && SG_SIZE(v, ci) == 3)) continue;
        const uint32_t *li = SG_LITS(v, ci);
        uint32_t xi = li[0], yi = li[1], zi = li[2];
        if (yi == dx) sg_swap(xi, yi);
        if (zi == dx) sg_swap(xi, zi);
        for (int j = i + 1; j < nd; ++j) {
            const uint32_t cj = dd[j];
            if (!(sg_original(v, cj) && SG_SIZE(v, cj) == 3)) continue;
            const uint32_t *lj = SG_LITS(v, cj);
            uint32_t xj = lj[0], yj = lj[1], zj = lj[2];
            if (yj == dx) sg_swap(xj, yj);
            if (zj == dx) sg_swap(xj, zj);
            if (SG_ABS(yi) == SG_ABS(zj)) sg_swap(yj, zj);
            if (SG_ABS(zi) == SG_ABS(zj)) continue;
            if (yi != SG_FLIP(yj)) continue;
            const uint32_t d1 = sg_find_ternary(v, L, fx, yi, SG_FLIP(zi));
            if (d1 == 0xFFFFFFFFu) continue;
            const uint32_t d2 = sg_find_ternary(v, L, fx, yj, SG_FLIP(zj));
            if (d2 == 0xFFFFFFFFu) continue;
            sg_wsync();
            if (L.lane == 0) { sg_melt(v, ci), sg_melt(v, cj), sg_melt(v, d1), sg_melt(v, d2); }
            sg_wsync();
            nAddedCls = 0, nAddedLits = 0;
            if (sg_count_pairs(v, L, SG_VE_SUBST, var, orgCls, dx, fx, nAddedCls, nAddedLits)) {
                sg_wsync();
                if (L.lane == 0) { sg_freeze(v, ci), sg_freeze(v, cj), sg_freeze(v, d1), sg_freeze(v, d2); }
                sg_wsync();
                return false;
            }
            return true;
        }
    }
    return false;
}

/* XOR gate (xor.hpp:26-181) */
SG_HD void sg_freeze_arities(const SgView &v, const SgLane &L, uint32_t a, uint32_t b)
{   /* xor.hpp:32-44 */
    sg_freeze_if(v, L, a, [](const sg_u4 &h) { return (int)h.y > 2; });
    sg_freeze_if(v, L, b, [](const sg_u4 &h) { return (int)h.y > 2; });
}

/* An XOR over n variables is 2^(n-1) clauses: one base clause and every copy of it with an even number of literals
 * negated.  The even-weight masks in ascending order are e(t) = 2t | (popcount(t) & 1), t = 0 .. 2^(n-1)-1 - the order in
 * which the reference enumerates them (xor.hpp:26,72-78) - so partner t of a base clause is written down directly. */
/* The loops over a clause's literals in these helpers stay rolled: the BVE decision kernel waits for instructions more than
 * for data (ncu: 9.9 warps per issue slot stalled on "no instruction" at 15 280 instructions), the helpers are inlined into
 * every XOR attempt, and unrolled four times each they were a sixth of the kernel. */
SG_HD void sg_xor_partner(const uint32_t *base, int size, uint32_t t, uint32_t *out)
{
    const uint32_t flips = (t << 1) | (uint32_t)(sg_popc(t) & 1);
#pragma unroll 1
    for (int k = 0; k < size; ++k) out[k] = base[k] ^ ((flips >> k) & 1u);
}
/* the list a partner is searched in: the first of the shortest occurrence lists of its literals (xor.hpp:81-92) */
SG_HD uint32_t sg_arity_best(const SgView &v, const uint32_t *literals, int size, uint32_t *sig_out)
{
    uint32_t via = literals[0], sig = 0;
#pragma unroll 1
    for (int k = 0; k < size; ++k) {
        sig |= SG_HASH(literals[k]);
        if (v.ot_len[literals[k]] < v.ot_len[via]) via = literals[k];
    }
    *sig_out = sig;
    return via;
}
/* original clause of `size` literals made of exactly the pattern's literals (checkArity, xor.hpp:53-68);
 * equal literal sets have equal signatures, so the signature is compared first */
SG_HD bool sg_arity_match(const SgView &v, const sg_u4 &h, const uint32_t *literals, int size, uint32_t sig)
{
    if ((h.z & SG_ST_MASK) || (int)h.y != size || h.w != sig) return false;
    /* a stored clause is in ascending literal order, and so is a partner pattern: negating literals of the (ascending) base
     * clause changes sign bits only, never the order of the variables - equal sets are equal sequences */
    const uint32_t *l = v.lits + h.x;
#pragma unroll 1
    for (int a = 0; a < size; ++a) if (l[a] != literals[a]) return false;
    return true;
}

/* finds partner clause `pattern` and marks it molten (xor.hpp:70-102) */
SG_HDN inline bool sg_melt_partner(const SgView &v, const SgLane &L, const uint32_t *pattern, int size)
{
    uint32_t sig;
    const uint32_t via = sg_arity_best(v, pattern, size, &sig);
    const int n = (int)v.ot_len[via];
    const uint32_t *d = sg_list(v, via);
    const int k = sg_find_first(L, n, [&](int i) { return sg_arity_match(v, v.hdr[d[i]], pattern, size, sig); });
    if (k < 0) return false;
    sg_melt1(v, L, d[k]);
    return true;
}

/* does clause `ci` have all its 2^arity - 1 partner clauses?  Read-only, one lane: whether a
 * partner exists does not depend on any molten flag, and a failed search leaves no mark behind
 * (freezeArities, xor.hpp:141-142), so the lanes can test different clauses at the same time. */
SG_HDN inline bool sg_xor_complete(const SgView &v, uint32_t ci, int size)
{
    uint32_t pattern[24];
    const uint32_t *base = SG_LITS(v, ci);
    for (uint32_t t = 1; t < (1u << (size - 1)); ++t) {
        sg_xor_partner(base, size, t, pattern);
        uint32_t sig;
        const uint32_t via = sg_arity_best(v, pattern, size, &sig);
        const int n = (int)v.ot_len[via];
        const uint32_t *d = sg_list(v, via);
        bool found = false;
        for (int i = 0; i < n && !found; ++i) found = sg_arity_match(v, v.hdr[d[i]], pattern, size, sig);
        if (!found) return false;
    }
    return true;
}

/* the same question answered by the whole warp, lanes over the partners: for wide XORs (2^(arity-1) - 1 partners, each a
 * search through a list of about 2^(arity-1) clauses) the first clause tried is nearly always the complete one, and testing
 * 32 clauses at once, one per lane, does 32 times the work the reference does */
#define SG_XOR_COOP_MIN 4                 /* clause size from which the partners are spread over the lanes */
SG_HDN inline bool sg_xor_complete_warp(const SgView &v, const SgLane &L, uint32_t ci, int size)
{
    uint32_t pattern[24];
    const uint32_t *base = SG_LITS(v, ci);
    const uint32_t npat = 1u << (size - 1);
    for (uint32_t t0 = 1; t0 < npat; t0 += L.width) {
        const uint32_t t = t0 + L.lane;
        bool found = true;
        if (t < npat) {
            sg_xor_partner(base, size, t, pattern);
            uint32_t sig;
            const uint32_t via = sg_arity_best(v, pattern, size, &sig);
            const int n = (int)v.ot_len[via];
            const uint32_t *d = sg_list(v, via);
            found = false;
            for (int i = 0; i < n && !found; ++i) found = sg_arity_match(v, v.hdr[d[i]], pattern, size, sig);
        }
        if (sg_ballot(!found)) return false;
    }
    return true;
}

SG_HDN inline bool sg_find_xor_gate(const SgView &v, const SgLane &L, uint32_t dx, uint32_t fx, int orgCls, int &nAddedCls,
                                    int &nAddedLits)
{
    const int nd = (int)v.ot_len[dx], nf = (int)v.ot_len[fx];
    const uint32_t *dd = sg_list(v, dx), *df = sg_list(v, fx);
    if (SG_SIZE(v, dd[nd - 1]) == 2 || SG_SIZE(v, df[nf - 1]) == 2) return false;
    const int maxarity = v.o.xor_max_arity;
    if (SG_SIZE(v, dd[0]) - 1 > maxarity) return false;
    const uint32_t var = SG_ABS(dx);
    uint32_t pattern[24];
#if defined(__CUDA_ARCH__)
    /* when both lists fit one warp, the count of the screen below - original clauses of the two lists with the same size
     * and the same sign-blind signature - is two warp matches and a ballot: lane q looks at entry q of (dx list, fx list) */
    const bool fits = nd + nf <= 32;
    int same_fast = 0;
    if (fits) {
        const int q = (int)L.lane;
        sg_u4 g = make_uint4(0, 0, SG_DELETED, 0);
        if (q < nd + nf) g = v.hdr[q < nd ? dd[q] : df[q - nd]];
        const uint32_t gw = g.w | ((g.w & 0xAAAAAAAAu) >> 1) | ((g.w & 0x55555555u) << 1);
        const uint32_t m_sig = __match_any_sync(0xffffffffu, gw), m_size = __match_any_sync(0xffffffffu, g.y);
        const uint32_t m_org = __ballot_sync(0xffffffffu, q < nd + nf && !(g.z & SG_ST_MASK));
        same_fast = __popc(m_sig & m_size & m_org);
    }
#endif
    for (int base = 0; base < nd; base += (int)L.width) {
        /* the reference tries the clauses of dx in list order and stops at the first complete XOR */
        const int idx = base + (int)L.lane;
        bool complete = false, wide_cand = false;
        if (idx < nd) {
            const sg_u4 h = v.hdr[dd[idx]];
            const int size = (int)h.y;
            if (!(h.z & SG_ST_MASK) && size >= 3 && size - 1 <= maxarity && size <= 24) {
                /* necessary condition, from the headers of the variable's own two lists only: the 2^arity clauses
                 * of an XOR are over ONE variable set and all contain x or -x, so at least 2^arity original
                 * clauses of this size with the same sign-blind signature must be in the two lists */
                const uint32_t wide = h.w | ((h.w & 0xAAAAAAAAu) >> 1) | ((h.w & 0x55555555u) << 1);
                int same = 0;
#if defined(__CUDA_ARCH__)
                if (fits) same = same_fast;
                else
#endif
                for (int q = 0; q < nd + nf; ++q) {
                    const sg_u4 g = v.hdr[q < nd ? dd[q] : df[q - nd]];
                    if ((g.z & SG_ST_MASK) || (int)g.y != size) continue;
                    const uint32_t gw = g.w | ((g.w & 0xAAAAAAAAu) >> 1) | ((g.w & 0x55555555u) << 1);
                    if (gw == wide) same++;
                }
                if (same >= (1 << (size - 1))) {
                    if (size >= SG_XOR_COOP_MIN) wide_cand = true;
                    else complete = sg_xor_complete(v, dd[idx], size);
                }
            }
        }
        /* first complete clause in list order: the narrow ones are answered already, a wide one is tested now, by the warp */
        const uint32_t mc = sg_ballot(complete);
        uint32_t m = mc | sg_ballot(wide_cand), ci = 0xFFFFFFFFu;
        while (m) {
            const int q = sg_ffs(m) - 1;
            m &= m - 1u;
            const uint32_t cq = dd[base + q];
            if (((mc >> q) & 1u) || sg_xor_complete_warp(v, L, cq, SG_SIZE(v, cq))) { ci = cq; break; }
        }
        if (ci == 0xFFFFFFFFu) continue;
        const int size = SG_SIZE(v, ci);
        const uint32_t *first = SG_LITS(v, ci);
        bool all = true;
        for (uint32_t t = 1; all && t < (1u << (size - 1)); ++t) {
            sg_xor_partner(first, size, t, pattern);
            all = sg_melt_partner(v, L, pattern, size);
        }
        if (!all) { sg_freeze_arities(v, L, dx, fx); continue; }           /* cannot happen: ci is complete */
        sg_melt1(v, L, ci);
        nAddedCls = 0, nAddedLits = 0;
        if (sg_count_pairs(v, L, SG_VE_SUBST, var, orgCls, dx, fx, nAddedCls, nAddedLits)) {
            sg_freeze_arities(v, L, dx, fx);
            break;
        }
        return true;
    }
    return false;
}

/* function-table gate (function.hpp:128-261), evaluated one 64-bit table word at a time */
SG_HD unsigned long long sg_fun_col(int mvar, bool sign, uint32_t w)
{   /* clause2fun (function.hpp:71-93) */
    if (mvar < 6) {
        const unsigned long long magic[6] = {
            0xaaaaaaaaaaaaaaaaull, 0xccccccccccccccccull, 0xf0f0f0f0f0f0f0f0ull,
            0xff00ff00ff00ff00ull, 0xffff0000ffff0000ull, 0xffffffff00000000ull };
        const unsigned long long val = magic[mvar];
        return sign ? ~val : val;
    }
    const unsigned long long base = sign ? ~0ull : 0ull;
    return ((w >> (mvar - 6)) & 1u) ? ~base : base;
}
SG_HD unsigned long long sg_fun_clause_word(const SgView &v, uint32_t c, uint32_t lit, uint32_t w)
{
    const uint32_t *l = SG_LITS(v, c);
    const int size = SG_SIZE(v, c);
    unsigned long long f = 0;
#pragma unroll 1
    for (int k = 0; k < size; ++k) {
        const uint32_t other = l[k];
        if (other == lit) continue;
        f |= sg_fun_col((int)v.marks[SG_ABS(other)], SG_SIGN(other) != 0, w);
    }
    return f;
}
/* word w of buildfuntab(lit) restricted to the first `upto` list entries (function.hpp:128-196) */
SG_HD unsigned long long sg_fun_list_word(const SgView &v, uint32_t lit, int upto, uint32_t w, unsigned long long init)
{
    const uint32_t *d = sg_list(v, lit);
    unsigned long long f = init;
    for (int i = 0; i < upto; ++i) {
        if (sg_learnt(v, d[i])) continue;
        f &= sg_fun_clause_word(v, d[i], lit, w);
    }
    return f;
}
SG_HD bool sg_fun_in_core(const SgView &v, const SgLane &L, uint32_t lit)
{   /* buildfuntab bails out on the first literal outside the 12 marked variables (function.hpp:150) */
    const int n = (int)v.ot_len[lit];
    const uint32_t *d = sg_list(v, lit);
    const int k = sg_find_first(L, n, [&](int i) {
        const sg_u4 h = v.hdr[d[i]];
        if (h.z & SG_LEARNT) return false;
        const uint32_t *l = v.lits + h.x;
        for (int q = 0; q < (int)h.y; ++q) {
            if (l[q] == lit) continue;
            if (v.marks[SG_ABS(l[q])] < 0) return true;
        }
        return false;
    });
    return k < 0;
}

SG_HDN inline bool sg_find_fun_gate(const SgView &v, const SgLane &L, uint32_t p, int orgCls, int &nAddedCls, int &nAddedLits)
{
    if (!v.o.ve_fun_en) return false;
    const uint32_t n = p | 1u;
    if (!sg_fun_in_core(v, L, p) || !sg_fun_in_core(v, L, n)) return false;
    const int np = (int)v.ot_len[p], nn = (int)v.ot_len[n];
    unsigned long long unsat = 0, allzero = 0;
    for (uint32_t w = 0; w < 64; ++w) {       /* collapsefun (function.hpp:198-209) */
        const unsigned long long b = sg_fun_list_word(v, p, np, w, ~0ull);
        const unsigned long long c = sg_fun_list_word(v, n, nn, w, ~0ull);
        unsat |= (b | c);
        allzero |= (b & c);
    }
    if (!unsat) { if (L.lane == 0) sg_set_flag(&v.sc->unsat); return true; }
    if (allzero) return false;
    bool core = false;
    const uint32_t *dp = sg_list(v, p), *dn = sg_list(v, n);
    for (int i = np - 1; i >= 0; --i) {
        if (!sg_original(v, dp[i])) continue;
        bool isfalse = true;
        for (uint32_t w = 0; w < 64 && isfalse; ++w) {
            const unsigned long long negw = sg_fun_list_word(v, n, nn, w, ~0ull);
            if (sg_fun_list_word(v, p, i, w, negw)) isfalse = false;
        }
        if (isfalse) { sg_melt1(v, L, dp[i]); core = true; }
    }
    for (int i = nn - 1; i >= 0; --i) {
        if (!sg_original(v, dn[i])) continue;
        bool isfalse = true;
        for (uint32_t w = 0; w < 64 && isfalse; ++w)
            if (sg_fun_list_word(v, n, i, w, ~0ull)) isfalse = false;
        if (isfalse) { sg_melt1(v, L, dn[i]); core = true; }
    }
    nAddedCls = 0, nAddedLits = 0;
    if (sg_count_pairs(v, L, SG_VE_CORE, SG_ABS(p), orgCls, p, n, nAddedCls, nAddedLits)) {
        if (core) { sg_freeze_all(v, L, p); sg_freeze_all(v, L, n); }
        return false;
    }
    return true;
}

/* model words written by toblivion / save_BN_gate for this variable (simplify.hpp:393-423) */
SG_HD uint32_t sg_model_words(const SgView &v, const SgLane &L, uint32_t p, int pOrgs, int nOrgs)
{
    const uint32_t side = (pOrgs > nOrgs) ? (p | 1u) : p;
    const int n = (int)v.ot_len[side];
    const uint32_t *d = sg_list(v, side);
    int words = 0;
#pragma unroll 1
    for (int i = (int)L.lane; i < n; i += (int)L.width) if (sg_original(v, d[i])) words += SG_SIZE(v, d[i]) + 1;
    return (uint32_t)sg_wsum(words) + 2u;
}

SG_HD int sg_count_orgs(const SgView &v, const SgLane &L, uint32_t lit)
{   /* simplify.hpp:280-286 */
    const int n = (int)v.ot_len[lit];
    const uint32_t *d = sg_list(v, lit);
    int c = 0;
#pragma unroll 1
    for (int i = (int)L.lane; i < n; i += (int)L.width) if (sg_original(v, d[i])) c++;
    return sg_wsum(c);
}

SG_HDN inline void sg_ve_count_item(const SgView &v, uint32_t eidx)
{
    const SgLane L = sg_lane_ctx();
    const uint32_t var = v.elected[eidx];
    const uint32_t p = 2 * var, n = p | 1u;
    const int np = (int)v.ot_len[p], nn = (int)v.ot_len[n];
    int pOrgs = 0, nOrgs = 0, nAddedCls = 0, nAddedLits = 0;
    if (v.o.firstcall) { pOrgs = np; nOrgs = nn; }
    else { pOrgs = sg_count_orgs(v, L, p); nOrgs = sg_count_orgs(v, L, n); }
    uint32_t type = SG_VE_NONE, aux = 0;
    unsigned long long *stat = nullptr;
    uint32_t def;
    if (!pOrgs || !nOrgs) {
        type = SG_VE_PURE;
        stat = &v.sc->pures;
    } else if ((def = sg_find_bn_gate(v, L, p, n)) != 0) {
        type = SG_VE_INVERTER;
        aux = def;
        stat = &v.sc->inverters;
    } else if ((pOrgs == 1 || nOrgs == 1) && !sg_count_pairs(v, L, SG_VE_RESOLVE, var, -1, p, n, nAddedCls, nAddedLits)) {
        type = SG_VE_RESOLVE;
        stat = &v.sc->resolutions;
    } else {
        nAddedCls = 0, nAddedLits = 0;
        const int orgCls = pOrgs + nOrgs;
        uint32_t *gate_in = v.scratch + v.scr_off[eidx];
        /* each finder is tried on (definition side, other side) and then the other way round unless the first attempt already
         * counted resolvents; written as two-trip loops that stay rolled, so that the kernel holds one copy of each finder */
        if (orgCls > 2) {
#pragma unroll 1
            for (int side = 0; side < 2 && !type; ++side) {
                if (side && nAddedCls) break;
                if (sg_find_ao_gate(v, L, side ? p : n, side ? n : p, gate_in, orgCls, nAddedCls, nAddedLits)) type = SG_VE_SUBST;
            }
            if (type) stat = &v.sc->andors;
        }
        if (!type && orgCls > 3) {
#pragma unroll 1
            for (int side = 0; side < 2 && !type; ++side) {
                if (side && nAddedCls) break;
                if (sg_find_ite_gate(v, L, side ? n : p, side ? p : n, orgCls, nAddedCls, nAddedLits)) { type = SG_VE_SUBST; stat = &v.sc->ites; }
            }
            if (!type) {
#pragma unroll 1
                for (int side = 0; side < 2 && !type; ++side) {
                    if (side && nAddedCls) break;
                    if (sg_find_xor_gate(v, L, side ? n : p, side ? p : n, orgCls, nAddedCls, nAddedLits)) { type = SG_VE_SUBST; stat = &v.sc->xors; }
                }
            }
        }
        if (!type && orgCls > 2 && sg_find_fun_gate(v, L, p, orgCls, nAddedCls, nAddedLits)) {
            if (sg_bcast(*(volatile uint32_t *)&v.sc->unsat)) { type = SG_VE_NONE; nAddedCls = 0; nAddedLits = 0; }
            else { type = SG_VE_CORE; stat = &v.sc->funs; }
        } else if (!type && !nAddedCls) {
            if (!sg_count_pairs(v, L, SG_VE_RESOLVE, var, orgCls, p, n, nAddedCls, nAddedLits)) {
                type = SG_VE_RESOLVE;
                stat = &v.sc->resolutions;
            }
        }
    }
    uint32_t ncls = 0, nlits = 0, nmodel = 0;
    if (type != SG_VE_NONE) {
        nmodel = sg_model_words(v, L, p, pOrgs, nOrgs);
        if (type == SG_VE_RESOLVE || type == SG_VE_SUBST || type == SG_VE_CORE) { ncls = (uint32_t)nAddedCls; nlits = (uint32_t)nAddedLits; }
    }
    if (L.lane == 0) {
        if (stat) sg_atomic_add64(stat, 1);
        v.ve_type[eidx] = type | ((uint32_t)(pOrgs > nOrgs) << 8);
        v.ve_aux[eidx] = aux;
        v.ve_ncls[eidx] = ncls;
        v.ve_nlits[eidx] = nlits;
        v.ve_nmodel[eidx] = nmodel;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* ERE: eager redundancy elimination (redundancy.cpp:26-123, redundancy.hpp:26-122)             */
/*                                                                                              */
/* The reference walks the elected variables in order and, for every pair of live clauses,      */
/* builds the resolvent m, picks the shortest occurrence list among m's literals and deletes or */
/* strengthens clauses there that m subsumes.  Unlike SUB/VE the targets lie OUTSIDE the        */
/* variable's neighbourhood, so the sequential order matters.  Exact parallel scheme:           */
/*  1. propose (one warp per elected variable, lanes over pairs, read-only on the pre-ERE       */
/*     state): every pair with at least one clause c that m subsumes is recorded as an event    */
/*     and every such c is marked `touched`.  A target only ever shrinks and its signature is   */
/*     never refreshed, so whatever a pair can hit later it could already hit now: the events   */
/*     are a superset of the sequential hits as long as the pair's parents are unchanged.       */
/*  2. an elected variable with a touched clause in its lists is `dirty` (its parents can       */
/*     change under it).                                                                         */
/*  3. replay (one warp, elected order): a dirty variable is processed in full exactly like the */
/*     reference, a clean one only at its recorded pairs; whenever a clause that belongs to a   */
/*     later elected variable is modified, that variable turns dirty before it is reached.      */
/* ------------------------------------------------------------------------------------------ */
#define SG_ERE_MAXLEN 256

/* merge_ere (redundancy.hpp:26-122): resolvent literals, signature and the literal with the shortest
 * occurrence list (first minimum in emission order).  Returns 0 for tautologies and oversize results. */
SG_HD int sg_merge_ere(const SgView &v, uint32_t x, const uint32_t *d1, int n1, const uint32_t *d2, int n2, int maxlen,
                       uint32_t *out, uint32_t *sig_out, uint32_t *best_out)
{
    int i = 0, j = 0, n = 0;
    uint32_t sig = 0, best = 0;
    int minsize = 0x7fffffff;
#define SG_ERE_PUSH(l) do { const int ls_ = (int)v.ot_len[l]; if (ls_ < minsize) { minsize = ls_; best = (l); } \
                            sig |= SG_HASH(l); if (n < SG_ERE_MAXLEN) out[n] = (l); n++; } while (0)
    while (i < n1 && j < n2) {
        if (n > maxlen) return 0;
        const uint32_t l1 = d1[i], l2 = d2[j];
        const uint32_t v1 = SG_ABS(l1), v2 = SG_ABS(l2);
        if (v1 == x) i++;
        else if (v2 == x) j++;
        else if ((l1 ^ l2) == 1u) return 0;
        else if (v1 < v2) { i++; SG_ERE_PUSH(l1); }
        else if (v2 < v1) { j++; SG_ERE_PUSH(l2); }
        else { i++; j++; SG_ERE_PUSH(l1); }
    }
    while (i < n1) { const uint32_t l1 = d1[i++]; if (SG_ABS(l1) != x) SG_ERE_PUSH(l1); }
    if (n > maxlen) return 0;
    while (j < n2) { const uint32_t l2 = d2[j++]; if (SG_ABS(l2) != x) SG_ERE_PUSH(l2); }
#undef SG_ERE_PUSH
    if (n > maxlen) return 0;
    *sig_out = sig; *best_out = best;
    return n;
}

/* does the resolvent subsume c (redundancy.cpp:84)?  c is read as it is NOW */
SG_HD bool sg_ere_hits(const SgView &v, const uint32_t *merged, int len, uint32_t sig, uint32_t c)
{
    const sg_u4 h = v.hdr[c];
    if (h.z & SG_DELETED) return false;
    return len <= (int)h.y && sg_sig_sub(sig, h.w) && sg_sub_lits(merged, len, v.lits + h.x, (int)h.y);
}

/* list setup shared by propose and replay (redundancy.cpp:44-58); false = the variable is skipped */
SG_HD bool sg_ere_lists(const SgView &v, uint32_t x, uint32_t &dx, uint32_t &fx)
{
    dx = 2 * x; fx = dx | 1u;
    if (v.ot_len[dx] > v.ot_len[fx]) { const uint32_t t = dx; dx = fx; fx = t; }
    const int ds = (int)v.ot_len[dx], fs = (int)v.ot_len[fx];
    const int maxoccurs = v.o.ere_max_occurs;
    if (!(ds && fs && ds <= maxoccurs && fs <= maxoccurs)) return false;
    const int maxsize = v.o.ere_clause_max;
    if (SG_SIZE(v, sg_list(v, dx)[0]) > maxsize || SG_SIZE(v, sg_list(v, fx)[0]) > maxsize) return false;
    return true;
}

/* work decomposition of the proposal pass: 256 clause pairs per warp-sized item, so that a variable with
 * long lists (up to 1000 x 1000 pairs) is spread over many warps */
#define SG_ERE_CHUNK 256
SG_HD void sg_ere_chunks_item(const SgView &v, uint32_t eidx)
{
    const uint32_t x = v.elected[eidx];
    v.ere_eidx[x] = eidx;
    uint32_t dx, fx, n = 0;
    if (sg_ere_lists(v, x, dx, fx)) {
        const unsigned long long pairs = (unsigned long long)v.ot_len[dx] * v.ot_len[fx];
        n = (uint32_t)((pairs + SG_ERE_CHUNK - 1) / SG_ERE_CHUNK);
    }
    v.ve_ncls[eidx] = n;                      /* scanned into ve_cls_off by the driver */
}

SG_HDN inline void sg_ere_propose_item(const SgView &v, uint32_t chunk)
{
    const SgLane L = sg_lane_ctx();
    /* which elected variable owns this chunk: last e with ve_cls_off[e] <= chunk */
    uint32_t lo = 0, hi = v.n_elected;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (v.ve_cls_off[mid] <= chunk) lo = mid; else hi = mid;
    }
    const uint32_t eidx = lo;
    const uint32_t x = v.elected[eidx];
    uint32_t dx, fx;
    if (!sg_ere_lists(v, x, dx, fx)) return;
    const int ds = (int)v.ot_len[dx], fs = (int)v.ot_len[fx];
    const uint32_t *dm = sg_list(v, dx), *dt = sg_list(v, fx);
    const int maxsize = v.o.ere_clause_max;
    uint32_t merged[SG_ERE_MAXLEN];
    bool any = false;
    const long long total = (long long)ds * fs;
    const long long first = (long long)(chunk - v.ve_cls_off[eidx]) * SG_ERE_CHUNK;
    const long long last = first + SG_ERE_CHUNK < total ? first + SG_ERE_CHUNK : total;
    for (long long q = first + L.lane; q < last; q += L.width) {
        const int i = (int)(q / fs), j = (int)(q - (long long)i * fs);
        const uint32_t iref = dm[i], jref = dt[j];
        const sg_u4 hi2 = v.hdr[iref];
        if (hi2.z & SG_DELETED) continue;
        const sg_u4 hj = v.hdr[jref];
        if (hj.z & SG_DELETED) continue;
        uint32_t sig, best;
        const int len = sg_merge_ere(v, x, v.lits + hi2.x, (int)hi2.y, v.lits + hj.x, (int)hj.y, maxsize, merged, &sig, &best);
        if (len < 2) continue;
        /* no clause of the list is as long as the resolvent (one byte per literal, cache resident): nothing to look at - on
         * k-SAT nine resolvents in ten are longer than everything in their shortest list */
        if (v.ere_maxsz) { const uint32_t mx = v.ere_maxsz[best]; if (mx != 255u && mx < (uint32_t)len) continue; }
        const int nm = (int)v.ot_len[best];
        const uint32_t *ml = sg_list(v, best);
        bool hit = false;
        for (int m0 = 0; m0 < nm; m0 += 8) {
            /* 8 entries of the shortest list at a time: ids, then headers (independent loads), then the cheap
             * size/signature screen; the literal comparison runs only for the survivors */
            uint32_t cc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) cc[k] = m0 + k < nm ? ml[m0 + k] : iref;
            if (v.ere_filt) {
                /* the screen reads the 8-byte {size, sig} snapshot (nothing is modified during the proposal pass) */
                uint32_t fs2[8], fg[8];
                /* first the one-byte sizes (10 MB for 10M clauses: cache resident, where the 8-byte records of a
                 * random dozen clauses per pair cost a DRAM sector each - ncu: 4.1 GB read per launch): a resolvent
                 * longer than the clause cannot subsume it, which settles almost every entry on k-SAT */
                uint32_t s8[8];
                bool any8 = false;
#pragma unroll
                for (int k = 0; k < 8; ++k) s8[k] = v.ere_size8[cc[k]];
#pragma unroll
                for (int k = 0; k < 8; ++k) any8 = any8 || (uint32_t)len <= s8[k] || s8[k] == 255u;
                if (!any8) continue;
#pragma unroll
                for (int k = 0; k < 8; ++k) {                 /* one 8-byte load per surviving entry */
                    fs2[k] = 0; fg[k] = 0;
                    if ((uint32_t)len <= s8[k] || s8[k] == 255u) {
                        const unsigned long long f = ((const unsigned long long *)v.ere_filt)[cc[k]];
                        fs2[k] = (uint32_t)f; fg[k] = (uint32_t)(f >> 32);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (cc[k] == iref || cc[k] == jref) continue;
                    if (len <= (int)fs2[k] && sg_sig_sub(sig, fg[k]) && sg_sub_lits(merged, len, v.lits + v.hdr[cc[k]].x, (int)fs2[k])) {
                        hit = true; v.ere_touched[cc[k]] = 1;
                    }
                }
                continue;
            }
            sg_u4 hh[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) hh[k] = v.hdr[cc[k]];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (cc[k] == iref || cc[k] == jref || (hh[k].z & SG_DELETED)) continue;
                if (len <= (int)hh[k].y && sg_sig_sub(sig, hh[k].w) && sg_sub_lits(merged, len, v.lits + hh[k].x, (int)hh[k].y)) {
                    hit = true; v.ere_touched[cc[k]] = 1;
                }
            }
        }
        if (hit) { uint32_t s2 = ((uint32_t)i << 16) | (uint32_t)j; sg_push_unit(v, eidx, s2, 0); any = true; }
    }
    if (sg_ballot(any) && L.lane == 0) v.ere_flags[eidx] |= 1u;
}

/* dirty = some clause of the variable's own lists may be modified by an event */
SG_HD void sg_ere_dirty_item(const SgView &v, uint32_t eidx)
{
    const uint32_t x = v.elected[eidx];
    for (uint32_t side = 0; side < 2; ++side) {
        const uint32_t lit = 2 * x + side;
        const int n = (int)v.ot_len[lit];
        const uint32_t *d = sg_list(v, lit);
        for (int i = 0; i < n; ++i) if (v.ere_touched[d[i]]) { v.ere_flags[eidx] |= 2u; sg_atomic_min(&v.sc->ere_first_dirty, eidx); return; }
    }
}

/* a clause was deleted or rewritten: the elected variable it belongs to (if any, and if still ahead in the
 * sequential order) must be processed in full */
SG_HD void sg_ere_mark_owner(const SgView &v, uint32_t c, const uint32_t *oldlits, int oldsize, uint32_t cur_eidx)
{
    for (int k = 0; k < oldsize; ++k) {
        const uint32_t u = SG_ABS(oldlits[k]);
        const uint32_t w = v.rank[u];
        if (SG_ER_STAT(w) == SG_E_BOTH && SG_ER_RANK(w) < v.sc->brk_rank) {
            const uint32_t e = v.ere_eidx[u];
            if (e <= cur_eidx) continue;
            if (!v.ere_flags[e]) {
                /* not on the replay list yet: queue it (ascending) so that the replay reaches it in order */
                SgScalars *sc = v.sc;
                if (sc->ere_npending < 64) {
                    uint32_t j = sc->ere_npending++;
                    while (j > 0 && sc->ere_pending[j - 1] > e) { sc->ere_pending[j] = sc->ere_pending[j - 1]; --j; }
                    sc->ere_pending[j] = e;
                } else sc->ere_overflow = 1;
            }
            v.ere_flags[e] |= 2u;
        }
    }
    (void)c;
}

/* the body of the reference's pair loop for one (ci, cj) (redundancy.cpp:59-117), current state */
SG_HDN inline void sg_ere_pair(const SgView &v, const SgLane &L, uint32_t x, uint32_t eidx, uint32_t iref, uint32_t jref, uint32_t *merged)
{
    const sg_u4 hi = v.hdr[iref];
    if (hi.z & SG_DELETED) return;
    const sg_u4 hj = v.hdr[jref];
    if (hj.z & SG_DELETED) return;
    const bool ciorg = !(hi.z & SG_ST_MASK), cjorg = !(hj.z & SG_ST_MASK);
    uint32_t sig, best;
    const int len = sg_merge_ere(v, x, v.lits + hi.x, (int)hi.y, v.lits + hj.x, (int)hj.y, v.o.ere_clause_max, merged, &sig, &best);
    if (len < 2) return;
    const int ereextend = v.o.ere_extend;
    const int nm = (int)v.ot_len[best];
    const uint32_t *ml = sg_list(v, best);
    for (int base = 0; base < nm; base += (int)L.width) {
        const int m = base + (int)L.lane;
        bool hit = false;
        if (m < nm) { const uint32_t c = ml[m]; hit = c != iref && c != jref && sg_ere_hits(v, merged, len, sig, c); }
        uint32_t mask = sg_ballot(hit);
        while (mask) {
            const int k = sg_ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t c = ml[base + k];
            const sg_u4 hc = v.hdr[c];
            const bool learnt = (hc.z & SG_LEARNT) != 0, equal = len == (int)hc.y;
            if (equal && (learnt || (ciorg && cjorg))) {
                sg_wsync();
                if (L.lane == 0) { sg_mark_deleted(v, c); sg_ere_mark_owner(v, c, v.lits + hc.x, (int)hc.y, eidx); }
                sg_wsync();
                return;                                   /* break: next pair */
            } else if (!equal && ereextend) {
                if (learnt && ereextend == 1) continue;
                sg_wsync();
                if (L.lane == 0) {
                    sg_ere_mark_owner(v, c, v.lits + hc.x, (int)hc.y, eidx);
                    uint32_t *l = v.lits + hc.x;
                    for (int q = 0; q < len; ++q) l[q] = merged[q];
                    uint32_t w0 = hc.z;
                    if (learnt) {                         /* bumpShrunken (solve.hpp:725-737) */
                        const int old_lbd = (int)(w0 >> SG_LBD_SH);
                        if (old_lbd > v.o.lbd_tier1) {
                            const int sz = len - 1, new_lbd = sz < old_lbd ? sz : old_lbd;
                            if (new_lbd < old_lbd) w0 = (w0 & 0xFu) | (1u << SG_USAGE_SH) | ((uint32_t)new_lbd << SG_LBD_SH);
                        }
                    }
                    v.hdr[c].y = (uint32_t)len;          /* signature deliberately left stale (redundancy.cpp:110-111) */
                    v.hdr[c].z = w0;
                }
                sg_wsync();
            }
        }
    }
}

/* would sg_ere_pair change anything for (iref, jref) in the current state?  Read-only, the same tests in the same order */
SG_HDN inline bool sg_ere_pair_acts(const SgView &v, const SgLane &L, uint32_t x, uint32_t iref, uint32_t jref, uint32_t *merged)
{
    const sg_u4 hi = v.hdr[iref];
    if (hi.z & SG_DELETED) return false;
    const sg_u4 hj = v.hdr[jref];
    if (hj.z & SG_DELETED) return false;
    const bool ciorg = !(hi.z & SG_ST_MASK), cjorg = !(hj.z & SG_ST_MASK);
    uint32_t sig, best;
    const int len = sg_merge_ere(v, x, v.lits + hi.x, (int)hi.y, v.lits + hj.x, (int)hj.y, v.o.ere_clause_max, merged, &sig, &best);
    if (len < 2) return false;
    const int ereextend = v.o.ere_extend;
    const int nm = (int)v.ot_len[best];
    const uint32_t *ml = sg_list(v, best);
    bool acts = false;
    for (int m = (int)L.lane; m < nm && !acts; m += (int)L.width) {
        const uint32_t c = ml[m];
        if (c == iref || c == jref || !sg_ere_hits(v, merged, len, sig, c)) continue;
        const sg_u4 hc = v.hdr[c];
        const bool learnt = (hc.z & SG_LEARNT) != 0, equal = len == (int)hc.y;
        acts = equal ? (learnt || (ciorg && cjorg)) : (ereextend && !(learnt && ereextend == 1));
    }
    return sg_ballot(acts) != 0;
}

/* ---- recorded pairs of the variables BEFORE the first dirty one: parallel rounds ------------------------------
 * Such a pair reads only its (untouched) parents and the clauses of its shortest list, and writes only clauses
 * its resolvent subsumes.  Per round every pending pair claims the clauses it could write (atomicMin of its
 * position in the sequential order); a pair that holds all its claims is the earliest pending writer of each of
 * them, so the state it sees is the state it would see in the sequential run, and the pairs applied in one round
 * write disjoint clauses.  The earliest pending pair always qualifies, so the rounds terminate. */
template <class F> SG_HD void sg_ere_pair_hits(const SgView &v, const SgLane &L, uint32_t eidx, uint32_t key, uint32_t *merged, F on_hit);
template <class F> SG_HD void sg_ere_event_hits(const SgView &v, const SgLane &L, uint32_t k, uint32_t *merged, F on_hit)
{
    sg_ere_pair_hits(v, L, v.ubuf[k].eidx, v.ubuf[k].seq, merged, on_hit);
}
/* pair (key >> 16, key & 0xffff) of elected variable eidx: every clause of the resolvent's shortest list it subsumes NOW */
template <class F> SG_HD void sg_ere_pair_hits(const SgView &v, const SgLane &L, uint32_t eidx, uint32_t key, uint32_t *merged, F on_hit)
{
    const uint32_t x = v.elected[eidx];
    uint32_t dx, fx;
    if (!sg_ere_lists(v, x, dx, fx)) return;
    const uint32_t iref = sg_list(v, dx)[key >> 16], jref = sg_list(v, fx)[key & 0xffffu];
    const sg_u4 hi = v.hdr[iref], hj = v.hdr[jref];
    if ((hi.z | hj.z) & SG_DELETED) return;
    uint32_t sig, best;
    const int len = sg_merge_ere(v, x, v.lits + hi.x, (int)hi.y, v.lits + hj.x, (int)hj.y, v.o.ere_clause_max, merged, &sig, &best);
    if (len < 2) return;
    const int nm = (int)v.ot_len[best];
    const uint32_t *ml = sg_list(v, best);
    for (int m = (int)L.lane; m < nm; m += (int)L.width) {
        const uint32_t c = ml[m];
        if (c != iref && c != jref && sg_ere_hits(v, merged, len, sig, c)) on_hit(c);
    }
}
SG_HDN inline void sg_ere_claim_item(const SgView &v, uint32_t k)
{
    if (v.ere_done[k]) return;
    const SgLane L = sg_lane_ctx();
    uint32_t merged[SG_ERE_MAXLEN];
    sg_ere_event_hits(v, L, k, merged, [&](uint32_t c) { sg_atomic_min(&v.ere_owner[c], k); });
}
SG_HDN inline void sg_ere_safe_item(const SgView &v, uint32_t k)
{
    if (v.ere_done[k]) return;
    const SgLane L = sg_lane_ctx();
    uint32_t merged[SG_ERE_MAXLEN];
    bool mine = true;
    sg_ere_event_hits(v, L, k, merged, [&](uint32_t c) { if (v.ere_owner[c] != k) mine = false; });
    const bool all = sg_ballot(!mine) == 0;
    if (L.lane == 0) v.ere_done[k] = all ? 2u : 0u;          /* 2 = apply in this round */
}
SG_HDN inline void sg_ere_apply_item(const SgView &v, uint32_t k)
{
    if (v.ere_done[k] != 2u) return;
    const SgLane L = sg_lane_ctx();
    uint32_t merged[SG_ERE_MAXLEN];
    const uint32_t eidx = v.ubuf[k].eidx, key = v.ubuf[k].seq;
    const uint32_t x = v.elected[eidx];
    uint32_t dx, fx;
    if (sg_ere_lists(v, x, dx, fx)) sg_ere_pair(v, L, x, eidx, sg_list(v, dx)[key >> 16], sg_list(v, fx)[key & 0xffffu], merged);
    if (L.lane == 0) { v.ere_done[k] = 1u; sg_atomic_add(&v.sc->ere_applied, 1u); }
}
/* ONE thread: how many recorded pairs / list entries lie before the first dirty variable */
SG_HD void sg_ere_split_serial(const SgView &v, uint32_t n_events, uint32_t n_list)
{
    const uint32_t fd = v.sc->ere_first_dirty;
    uint32_t lo = 0, hi = n_events;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (v.ubuf[mid].eidx < fd) lo = mid + 1; else hi = mid; }
    v.sc->ere_par_events = lo;
    lo = 0; hi = n_list;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (v.acting[mid] < fd) lo = mid + 1; else hi = mid; }
    v.sc->ere_list_start = lo;
}

/* processes elected variable e during the replay: in full when dirty, else only at its recorded pairs */
SG_HDN inline void sg_ere_replay_var(const SgView &v, const SgLane &L, uint32_t e, uint32_t fl, uint32_t &ep, uint32_t n_events, uint32_t *merged)
{
    const uint32_t x = v.elected[e];
    while (ep < n_events && v.ubuf[ep].eidx < e) ep++;
    uint32_t dx, fx;
    if (fl & 2u) {
        if (L.lane == 0) v.sc->ere_dirty_vars++;
        if (sg_ere_lists(v, x, dx, fx)) {
            const int ds = (int)v.ot_len[dx], fs = (int)v.ot_len[fx];
            const uint32_t *dm = sg_list(v, dx), *dt = sg_list(v, fx);
            if (L.lane == 0) v.sc->ere_full_pairs += (unsigned long long)ds * fs;
            for (int i = 0; i < ds; ++i) {
                if (sg_deleted(v, dm[i])) continue;
                for (int j = 0; j < fs; ++j) sg_ere_pair(v, L, x, e, dm[i], dt[j], merged);
            }
        }
    } else {
        if (!sg_ere_lists(v, x, dx, fx)) return;
        const uint32_t *dm = sg_list(v, dx), *dt = sg_list(v, fx);
        while (ep < n_events && v.ubuf[ep].eidx == e) {
            const uint32_t key = v.ubuf[ep].seq;
            if (!(v.ere_done && v.ere_done[ep])) {             /* not already applied by a parallel round */
                sg_ere_pair(v, L, x, e, dm[key >> 16], dt[key & 0xffffu], merged);
                if (L.lane == 0) v.sc->ere_event_pairs++;
            }
            ep++;
        }
    }
}

/* ONE warp: sequential replay in elected order over the variables that have recorded pairs or are dirty
 * (`list`, ascending), merged with the variables that turn dirty on the way (sc->ere_pending) */
SG_HDN inline void sg_ere_replay_serial(const SgView &v, uint32_t n_events, uint32_t n_list)
{
    const SgLane L = sg_lane_ctx();
    uint32_t merged[SG_ERE_MAXLEN];
    const uint32_t *list = v.acting;
    /* when the parallel rounds finished every pair before the first dirty variable, start there */
    const bool skip = v.ere_done && sg_bcast(*(volatile uint32_t *)&v.sc->ere_applied) == sg_bcast(*(volatile uint32_t *)&v.sc->ere_par_events);
    uint32_t ep = skip ? sg_bcast(*(volatile uint32_t *)&v.sc->ere_par_events) : 0u;
    uint32_t li = skip ? sg_bcast(*(volatile uint32_t *)&v.sc->ere_list_start) : 0u, cur = 0;
    bool started = false;
    while (true) {
        sg_wsync();
        if (sg_bcast(*(volatile uint32_t *)&v.sc->ere_overflow)) break;
        const uint32_t npend = sg_bcast(*(volatile uint32_t *)&v.sc->ere_npending);
        const uint32_t pend0 = npend ? sg_bcast(*(volatile uint32_t *)&v.sc->ere_pending[0]) : 0xFFFFFFFFu;
        const uint32_t next = li < n_list ? list[li] : 0xFFFFFFFFu;
        if (next == 0xFFFFFFFFu && pend0 == 0xFFFFFFFFu) return;
        uint32_t e;
        if (pend0 < next) {
            e = pend0;
            if (L.lane == 0) {                                  /* pop the head of the pending queue */
                SgScalars *sc = v.sc;
                for (uint32_t k = 1; k < sc->ere_npending; ++k) sc->ere_pending[k - 1] = sc->ere_pending[k];
                sc->ere_npending--;
            }
        } else { e = next; li++; }
        sg_wsync();
        const uint32_t fl = sg_bcast(*(volatile uint8_t *)&v.ere_flags[e]);
        sg_ere_replay_var(v, L, e, fl, ep, n_events, merged);
        cur = e; started = true;
    }
    /* more than 64 variables were waiting at once: finish with a plain scan over the flags of the
     * variables after the last one processed (everything queued or listed has a non-zero flag) */
    for (uint32_t e = started ? cur + 1 : 0; e < v.n_elected; ++e) {
        sg_wsync();
        const uint32_t fl = sg_bcast(*(volatile uint8_t *)&v.ere_flags[e]);
        if (fl) sg_ere_replay_var(v, L, e, fl, ep, n_events, merged);
    }
}

/* Helper warps of the replay block: they walk the recorded pairs ahead of the replaying warp and evaluate them
 * read-only (results discarded), which pulls the parents, the resolvent's shortest list and its candidates into
 * L1/L2 - the replay itself is one long chain of dependent fetches.  No effect on the result. */
SG_HDN inline void sg_ere_prefetch(const SgView &v, uint32_t n_events, uint32_t n_list, uint32_t helper, uint32_t n_helpers)
{
    const SgLane L = sg_lane_ctx();
    uint32_t merged[SG_ERE_MAXLEN];
    uint32_t hits = 0;
    /* variables that are dirty from the start are replayed in full: all their pairs */
    for (uint32_t li = helper; li < n_list; li += n_helpers) {
        const uint32_t e = v.acting[li];
        if (!(v.ere_flags[e] & 2u)) continue;
        uint32_t dx, fx;
        if (!sg_ere_lists(v, v.elected[e], dx, fx)) continue;
        const uint32_t ds = v.ot_len[dx], fs = v.ot_len[fx];
        if (ds > 0xffffu || fs > 0xffffu) continue;
        for (uint32_t i = 0; i < ds; ++i)
            for (uint32_t j = 0; j < fs; ++j) sg_ere_pair_hits(v, L, e, (i << 16) | j, merged, [&](uint32_t) { hits++; });
    }
    for (uint32_t k = helper; k < n_events; k += n_helpers) {
        if (v.ere_done && v.ere_done[k] == 1u) continue;
        sg_ere_event_hits(v, L, k, merged, [&](uint32_t) { hits++; });
    }
    if (hits == 0xFFFFFFFFu) v.ere_flags[0] |= 0x80u;      /* never true: keeps the loads alive */
}

/* ------------------------------------------------------------------------------------------ */
/* BVE: emission (bounded.cpp:187-307), model records (simplify.hpp:393-423, model.hpp:65-101)  */
/* ------------------------------------------------------------------------------------------ */
SG_HD void sg_save_clause(const SgView &v, uint32_t c, uint32_t witlit, unsigned long long &mpos)
{   /* model.hpp:86-101 */
    const uint32_t *l = SG_LITS(v, c);
    const int size = SG_SIZE(v, c);
    const unsigned long long last = mpos;
    int pos = -1;
    for (int i = 0; i < size; ++i) {
        const uint32_t lit = l[i];
        v.model[mpos++] = (v.vorg[SG_ABS(lit)] << 1) | SG_SIGN(lit);
        if (lit == witlit) pos = i;
    }
    if (pos > 0) { const uint32_t t = v.model[last + pos]; v.model[last + pos] = v.model[last]; v.model[last] = t; }
    v.model[mpos++] = (uint32_t)size;
}
SG_HD void sg_save_witness(const SgView &v, uint32_t lit, unsigned long long &mpos)
{   /* model.hpp:72-76 */
    v.model[mpos++] = (v.vorg[SG_ABS(lit)] << 1) | SG_SIGN(lit);
    v.model[mpos++] = 1;
}

/* substitute_single on one clause (equivalence.hpp:27-56): returns the unit or 0 */
SG_HD uint32_t sg_substitute_clause(const SgView &v, uint32_t c, uint32_t dx, uint32_t def)
{
    uint32_t *l = SG_LITS(v, c);
    const int size = SG_SIZE(v, c);
    int j = 0;
    bool had_def = false;
    for (int i = 0; i < size; ++i) {
        const uint32_t lit = l[i];
        if (lit == dx) l[j++] = def;
        else if (lit != def) l[j++] = lit;
        else had_def = true;
    }
    v.hdr[c].y = (uint32_t)j;
    /* the clause now occurs in the list of `def` (it did already if a copy of `def` was dropped): remembered
     * for the incremental occurrence-table update of the next phase */
    if (!had_def && v.ot_add) {
        const uint32_t k = sg_atomic_add(&v.sc->n_ot_add, 1u);
        if (k < v.ot_add_cap) { v.ot_add[2 * k] = def; v.ot_add[2 * k + 1] = c; }
    }
    if (j == 1) return l[0];
    sg_sort_lits(l, j);
    uint32_t sig = 0;
    for (int i = 0; i < j; ++i) sig |= SG_HASH(l[i]);
    v.hdr[c].w = sig;
    return 0;
}

SG_HD bool sg_clause_has(const SgView &v, uint32_t c, uint32_t lit)
{   /* SCLAUSE::has (sclause.hpp:165-184): membership in a sorted clause */
    const uint32_t *l = SG_LITS(v, c);
    const int size = SG_SIZE(v, c);
    for (int i = 0; i < size; ++i) if (l[i] == lit) return true;
    return false;
}

SG_HDN inline void sg_ve_emit_item(const SgView &v, uint32_t eidx)
{
    const uint32_t tw = v.ve_type[eidx];
    const uint32_t type = tw & 0xffu;
    if (type == SG_VE_NONE) return;
    const bool save_neg = (tw >> 8) & 1u;               /* pOrgs > nOrgs */
    const uint32_t var = v.elected[eidx];
    const uint32_t p = 2 * var, n = p | 1u;
    const int np = (int)v.ot_len[p], nn = (int)v.ot_len[n];
    const uint32_t *dp = sg_list(v, p), *dn = sg_list(v, n);
    unsigned long long mpos = v.model_base + v.ve_mod_off[eidx];
    uint32_t seq = 0;

    if (type == SG_VE_INVERTER) {
        /* save_BN_gate (equivalence.hpp:163-183) */
        if (save_neg) {
            for (int i = 0; i < nn; ++i) if (sg_original(v, dn[i])) sg_save_clause(v, dn[i], n, mpos);
            sg_save_witness(v, p, mpos);
        } else {
            for (int i = 0; i < np; ++i) if (sg_original(v, dp[i])) sg_save_clause(v, dp[i], p, mpos);
            sg_save_witness(v, n, mpos);
        }
        /* substitute_single (equivalence.hpp:58-101) */
        const uint32_t def = v.ve_aux[eidx], def_f = SG_FLIP(def);
        for (int i = 0; i < nn; ++i) {
            const uint32_t c = dn[i];
            const uint32_t w0 = SG_W0(v, c);
            if ((w0 & SG_LEARNT) || (w0 & SG_MOLTEN) || sg_clause_has(v, c, def)) sg_mark_deleted(v, c);
            else if (!(w0 & SG_ST_MASK)) {
                const uint32_t unit = sg_substitute_clause(v, c, n, def_f);
                if (unit) sg_push_unit(v, eidx | 0x80000000u, seq, unit);
            }
        }
        for (int i = 0; i < np; ++i) {
            const uint32_t c = dp[i];
            const uint32_t w0 = SG_W0(v, c);
            if ((w0 & SG_LEARNT) || (w0 & SG_MOLTEN) || sg_clause_has(v, c, def_f)) sg_mark_deleted(v, c);
            else if (!(w0 & SG_ST_MASK)) {
                const uint32_t unit = sg_substitute_clause(v, c, p, def);
                if (unit) sg_push_unit(v, eidx | 0x80000000u, seq, unit);
            }
        }
    } else {
        if (type != SG_VE_PURE && v.ve_ncls[eidx]) {
            /* resolve / substitute / resolveCore (bounded.cpp:187-277): outer loop = shorter list */
            uint32_t dx = p, fx = n;
            if (v.ot_len[dx] > v.ot_len[fx]) { dx = n; fx = p; }
            const int nme = (int)v.ot_len[dx], noth = (int)v.ot_len[fx];
            const uint32_t *dm = sg_list(v, dx), *dt = sg_list(v, fx);
            uint32_t cid = v.n_cls + v.ve_cls_off[eidx];
            uint32_t off = v.pool_used + v.ve_lit_off[eidx];
            const uint32_t cid_end = cid + v.ve_ncls[eidx];
            for (int i = 0; i < nme; ++i) {
                const sg_u4 hi = v.hdr[dm[i]];
                if (hi.z & SG_ST_MASK) continue;
                const bool a = (hi.z & SG_MOLTEN) != 0;
                for (int j = 0; j < noth; ++j) {
                    const sg_u4 hj = v.hdr[dt[j]];
                    if (hj.z & SG_ST_MASK) continue;
                    const bool b = (hj.z & SG_MOLTEN) != 0;
                    if (!sg_pair_ok((int)type, a, b)) continue;
                    /* a unit resolvent is never counted (simplify.hpp:341,366), so size it first */
                    const int rsize = sg_merge_count(var, v.lits + hi.x, (int)hi.y, v.lits + hj.x, (int)hj.y);
                    if (rsize > 1 && cid < cid_end) {
                        uint32_t sig;
                        const int sz = sg_merge_emit(var, v.lits + hi.x, (int)hi.y, v.lits + hj.x, (int)hj.y, v.lits + off, &sig);
                        sg_u4 h; h.x = off; h.y = (uint32_t)sz; h.z = SG_ADDED; h.w = sig;   /* addResolvent, bounded.cpp:296-306 */
                        v.hdr[cid] = h;
                        cid++; off += (uint32_t)sz;
                    } else if (rsize == 1) {
                        uint32_t tmp[2], sig;
                        const uint32_t *a1 = v.lits + hi.x, *a2 = v.lits + hj.x;
                        /* at most one literal survives: find it without touching the pool */
                        (void)sig; tmp[0] = 0;
                        for (int k = 0; k < (int)hi.y; ++k) if (SG_ABS(a1[k]) != var) tmp[0] = a1[k];
                        for (int k = 0; k < (int)hj.y; ++k) if (SG_ABS(a2[k]) != var) tmp[0] = a2[k];
                        if (tmp[0]) sg_push_unit(v, eidx | 0x80000000u, seq, tmp[0]);
                    }
                }
            }
        }
        /* toblivion (simplify.hpp:393-423) */
        if (save_neg) {
            for (int i = 0; i < nn; ++i) { if (sg_original(v, dn[i])) sg_save_clause(v, dn[i], n, mpos); sg_mark_deleted(v, dn[i]); }
            sg_save_witness(v, p, mpos);
            for (int i = 0; i < np; ++i) sg_mark_deleted(v, dp[i]);
        } else {
            for (int i = 0; i < np; ++i) { if (sg_original(v, dp[i])) sg_save_clause(v, dp[i], p, mpos); sg_mark_deleted(v, dp[i]); }
            sg_save_witness(v, n, mpos);
            for (int i = 0; i < nn; ++i) sg_mark_deleted(v, dn[i]);
        }
        v.ot_len[p] = 0;
        v.ot_len[n] = 0;
    }
    /* markEliminated (solve.hpp:491-498) */
    v.state[var] = (uint8_t)SG_MELTED_M;
    sg_atomic_sub(&v.sc->unassigned, 1u);
}

/* ------------------------------------------------------------------------------------------ */
/* one elected variable, start to finish of the per-variable stages: sortOT of its two lists,   */
/* SUB, then the BVE decision, by ONE warp, so that the clause neighbourhood is fetched from HBM */
/* once and stays in L1 for all three (the reference runs the three stages as separate loops    */
/* over the elected variables; their neighbourhoods are disjoint, so the order is immaterial).  */
/* Units found by SUB keep their place before those found by BVE (bit 31 of the unit key).      */
/* ------------------------------------------------------------------------------------------ */
SG_HDN inline void sg_neighbourhood_item(const SgView &v, uint32_t eidx)
{
    const SgLane L = sg_lane_ctx();
    const uint32_t var = v.elected[eidx];
    if (v.nbh_sort && (L.lane < 2 || L.width == 1)) {
        sg_sortot_item(v, 2 * var + (L.width == 1 ? 0u : L.lane));
        if (L.width == 1) sg_sortot_item(v, 2 * var + 1);
    }
    sg_wsync();
    if (v.o.sub_en || v.o.ve_plus_en) sg_sub_item(v, eidx);
    sg_wsync();
    sg_ve_count_item(v, eidx);
}

/* ------------------------------------------------------------------------------------------ */
/* BCE: blocked clause elimination (blocked.cpp:23-58), after VE, on the elected variables that  */
/* were not eliminated.  Two passes like VE: mark + size the model records, scan, then write.    */
/* ------------------------------------------------------------------------------------------ */
SG_HD bool sg_is_tautology(uint32_t x, const uint32_t *d1, int n1, const uint32_t *d2, int n2)
{   /* simplify.hpp:159-179 */
    int i = 0, j = 0;
    while (i < n1 && j < n2) {
        const uint32_t l1 = d1[i], l2 = d2[j];
        const uint32_t v1 = SG_ABS(l1), v2 = SG_ABS(l2);
        if (v1 == x) i++;
        else if (v2 == x) j++;
        else if ((l1 ^ l2) == 1u) return true;
        else if (v1 < v2) i++;
        else if (v2 < v1) j++;
        else { i++; j++; }
    }
    return false;
}
SG_HDN inline void sg_bce_count_item(const SgView &v, uint32_t eidx)
{
    const uint32_t var = v.elected[eidx];
    uint32_t words = 0;
    if (!v.state[var]) {                                   /* VE zeroes the slot of an eliminated variable (bounded.cpp:178-181) */
        const uint32_t p = 2 * var, n = p | 1u;
        const int np = (int)v.ot_len[p], nn = (int)v.ot_len[n];
        if (np <= v.o.bce_max_occurs && nn <= v.o.bce_max_occurs) {
            const uint32_t *dp = sg_list(v, p), *dn = sg_list(v, n);
            for (int i = 0; i < nn; ++i) {
                const sg_u4 hn = v.hdr[dn[i]];
                if (hn.z & SG_ST_MASK) continue;
                bool all = true;
                for (int j = 0; j < np && all; ++j) {
                    const sg_u4 hp = v.hdr[dp[j]];
                    if (hp.z & SG_ST_MASK) continue;
                    if (!sg_is_tautology(var, v.lits + hn.x, (int)hn.y, v.lits + hp.x, (int)hp.y)) all = false;
                }
                if (all) { v.ere_touched[dn[i]] = 1; words += hn.y + 1; }
            }
        }
    }
    v.ve_nmodel[eidx] = words;
}
SG_HDN inline void sg_bce_emit_item(const SgView &v, uint32_t eidx)
{
    if (!v.ve_nmodel[eidx]) return;
    const uint32_t var = v.elected[eidx];
    const uint32_t n = 2 * var + 1;
    const int nn = (int)v.ot_len[n];
    const uint32_t *dn = sg_list(v, n);
    unsigned long long mpos = v.model_base + v.ve_mod_off[eidx];
    for (int i = 0; i < nn; ++i) {
        const uint32_t c = dn[i];
        if (!v.ere_touched[c]) continue;
        sg_save_clause(v, c, n, mpos);                     /* model.saveClause(neg, neg.size(), n) */
        sg_mark_deleted(v, c);
        v.ere_touched[c] = 0;
    }
}

#endif /* SIGMA_ITEMS_CUH */
