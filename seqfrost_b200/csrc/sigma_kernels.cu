/* sigma_kernels.cu - the CUDA (sm_100a) implementation of sigma_backend.h.
 *
 * Streaming kernels (ingest, histogram, scan, scatter, radix sort, count, compaction, export)
 * are HBM-bound integer work: coalesced 128-bit header loads, grid sized from the element count,
 * no tensor cores (nothing here is a contraction).  The irregular stages run one work item per
 * thread with the bodies of sigma_items.cuh.  Everything is launched on one stream per context.
 */
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "sigma_backend.h"
#include "sigma_items.cuh"
#include "sigma_dimacs.cuh"

namespace {

/* Per-context launch state: the handle sigma_pass.cpp passes around as its "stream".  The launch counter and the first
 * CUDA error belong to the context that saw them (contexts are independent and may be driven by different host threads). */
struct BeStream {
    cudaStream_t st;
    unsigned long long launches;
    char err[256];
    void *scratch;                  /* be_sort_units (large n): kept between calls - cudaMalloc / cudaFree inside a stage cost up */
    size_t scratch_bytes;           /* to a second now and then (seen as 34 -> 1500 ms ERE stages on the 5M-gate circuit) */
};

inline void note_err(void *s, const char *what)
{
    cudaError_t e = cudaGetLastError();
    BeStream *b = (BeStream *)s;
    if (e != cudaSuccess && !b->err[0]) snprintf(b->err, sizeof b->err, "%s: %s", what, cudaGetErrorString(e));
}

#define ST(s) (((BeStream *)(s))->st)
#define LAUNCH(kernel, grid, block, stream, ...)                      \
    do {                                                              \
        kernel<<<(grid), (block), 0, ST(stream)>>>(__VA_ARGS__);      \
        ((BeStream *)(stream))->launches++;                           \
        note_err(stream, #kernel);                                    \
    } while (0)

constexpr int TPB = 256;
inline unsigned blocks_for(uint64_t n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

/* ---------------------------------------------------------------------------------------------- */
/* ingest                                                                                           */
/* ---------------------------------------------------------------------------------------------- */
/* Bulk (TMA) store of a staged tile: shared memory -> global, both 16-byte aligned, `bytes` a multiple of 16.
 * Issued by one thread; the staged words were written through the generic proxy, hence the proxy fence (executed by
 * every writer before the barrier that precedes this call). */
__device__ __forceinline__ void stage_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(sa), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

/* A contiguous run of W output words starting at global word `g0` of `out`, staged in `stage` at index (g0 & 3) so that
 * shared and global addresses agree modulo 16 bytes: the aligned body goes out as one bulk store, the up-to-3-word head
 * and tail as plain stores.  Called by the whole block after the barrier that follows the staging writes. */
__device__ __forceinline__ void flush_staged(uint32_t *out, unsigned long long g0, const uint32_t *stage, uint32_t W)
{
    const uint32_t shift = (uint32_t)(g0 & 3ull);
    uint32_t head = (4u - shift) & 3u;
    if (head > W) head = W;
    const uint32_t body = (W - head) & ~3u, tail = W - head - body;
    if (threadIdx.x == 0 && body) bulk_store(out + g0 + head, stage + shift + head, body * 4u);
    if (threadIdx.x >= 32 && threadIdx.x - 32 < head) out[g0 + (threadIdx.x - 32)] = stage[shift + (threadIdx.x - 32)];
    if (threadIdx.x >= 64 && threadIdx.x - 64 < tail) out[g0 + head + body + (threadIdx.x - 64)] = stage[shift + head + body + (threadIdx.x - 64)];
    if (threadIdx.x == 0 && body) bulk_store_wait();
}

constexpr uint32_t EX_CAP = 6144;      /* staging words per 256-clause tile (24 KB); larger tiles write directly */

__global__ void k_extract_sizes(const uint32_t *__restrict__ cm, uint64_t cm_buckets, const uint64_t *__restrict__ crefs, uint32_t n,
                                const uint8_t *__restrict__ stencil, uint64_t stencil_size, uint32_t *__restrict__ sizes,
                                uint32_t *__restrict__ live, uint32_t *bad)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t sz = sg_extract_size(cm, cm_buckets, stencil, stencil_size, crefs[i]);
    if (sz == SG_EXTRACT_BAD) { *bad = 1; sz = 0; }
    sizes[i] = sz;
    live[i] = sz != 0;
}

__global__ void __launch_bounds__(TPB) k_extract_write(const uint32_t *__restrict__ cm, const uint64_t *__restrict__ crefs, uint32_t n,
                                const uint32_t *__restrict__ sizes, const uint32_t *__restrict__ newoff,
                                const uint32_t *__restrict__ newid, uint32_t *__restrict__ arena,
                                unsigned long long *__restrict__ refs)
{
    /* the SCLAUSE records of a block's 256 source clauses are one contiguous run of the arena: assembled (and their
     * literals sorted) in shared memory, written out with one bulk store */
    __shared__ __align__(128) uint32_t stage[EX_CAP + 4];
    const uint32_t i0 = blockIdx.x * blockDim.x, i = i0 + threadIdx.x;
    const uint32_t next = i0 + blockDim.x < n ? i0 + blockDim.x : n;
    const unsigned long long g0 = newoff[i0];
    const uint32_t end = next < n ? newoff[next] : newoff[n - 1] + sizes[n - 1];
    const uint32_t W = end - (uint32_t)g0;
    const bool staged = W <= EX_CAP;
    if (i < n && sizes[i]) {
        const uint32_t off = newoff[i];
        refs[newid[i]] = off;
        sg_extract_write(cm, crefs[i], staged ? stage + (uint32_t)(g0 & 3ull) + (off - (uint32_t)g0) : arena + off);
    }
    if (staged) {
        stage_fence();
        __syncthreads();
        if (W) flush_staged(arena, g0, stage, W);
    }
}

__global__ void k_cm_flags(const uint32_t *__restrict__ arena, const unsigned long long *__restrict__ refs, uint32_t n,
                           uint32_t *__restrict__ learnt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) learnt[i] = (arena[refs[i]] & SG_ST_MASK) == SG_LEARNT;
}

__global__ void __launch_bounds__(TPB) k_cm_write(const uint32_t *__restrict__ arena, const unsigned long long *__restrict__ refs, uint32_t n,
                           const uint32_t *__restrict__ learnt_before, int lbd_tier1, uint32_t *__restrict__ cm,
                           unsigned long long *__restrict__ orgs, unsigned long long *__restrict__ learnts,
                           unsigned long long *__restrict__ counts, uint8_t *__restrict__ subsume, unsigned long long cm_total,
                           const uint32_t *__restrict__ vmap)
{
    /* the CLAUSE records of 256 consecutive clauses are one contiguous run of `cm`: staged, one bulk store */
    __shared__ __align__(128) uint32_t stage[EX_CAP + 4];
    const uint32_t i0 = blockIdx.x * blockDim.x, i = i0 + threadIdx.x;
    const uint32_t next = i0 + blockDim.x < n ? i0 + blockDim.x : n;
    const unsigned long long g0 = refs[i0] + i0;
    const unsigned long long end = next < n ? refs[next] + next : cm_total;
    const uint32_t W = (uint32_t)(end - g0);
    const bool staged = end - g0 <= EX_CAP;
    uint32_t lo = 0, ll = 0;
    if (i < n) {
        const unsigned long long r = refs[i];
        const uint32_t *src = arena + r;
        const uint32_t w0 = src[0], size = src[2];
        sg_cm_write(src, staged ? stage + (uint32_t)(g0 & 3ull) + (uint32_t)(r + i - g0) : cm + r + i, lbd_tier1, vmap);
        const uint32_t lb = learnt_before[i];
        if ((w0 & SG_ST_MASK) == SG_LEARNT) { learnts[lb] = r + i; ll = size; }
        else { orgs[i - lb] = r + i; lo = size; }
        if (subsume && (w0 & SG_ADDED))
            for (uint32_t k = 0; k < size; ++k) subsume[SG_ABS(src[3 + k])] = 1;
    }
    if (staged) {
        stage_fence();
        __syncthreads();
        if (W) flush_staged(cm, g0, stage, W);
    }
    /* literal totals (stats.literals.original / learnt): one atomic per warp */
    for (int d = 16; d; d >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, d); ll += __shfl_xor_sync(0xffffffffu, ll, d); }
    if ((threadIdx.x & 31) == 0) {
        if (lo) atomicAdd(&counts[0], (unsigned long long)lo);
        if (ll) atomicAdd(&counts[1], (unsigned long long)ll);
    }
}

__global__ void k_ere_filt(SgView v, uint2 *__restrict__ filt, uint8_t *__restrict__ size8)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.n_cls) return;
    const uint4 h = v.hdr[c];
    const uint32_t size = (h.z & SG_DELETED) ? 0u : h.y;
    filt[c] = make_uint2(size, h.w);
    size8[c] = (uint8_t)(size < 255u ? size : 255u);
}

/* largest clause (as in size8) in every occurrence list: a resolvent longer than that subsumes nothing in the list */
__global__ void k_ere_maxsz(SgView v, const uint8_t *__restrict__ size8, uint8_t *__restrict__ maxsz)
{
    const uint32_t lit = blockIdx.x * blockDim.x + threadIdx.x;
    if (lit >= v.nlit) return;
    const uint32_t n = v.ot_len[lit];
    const uint32_t *d = v.ot_data + v.ot_off[lit];
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; ++i) { const uint32_t s8 = size8[d[i]]; m = s8 > m ? s8 : m; }
    maxsz[lit] = (uint8_t)m;
}

/* a ref or size word that points outside the uploaded arena raises *bad (the driver returns SIGMA_E_ARG) instead of
 * being followed */
__global__ void k_ingest_sizes(const uint32_t *__restrict__ arena, const uint64_t *__restrict__ refs, uint32_t n,
                               uint64_t n_buckets, uint32_t *__restrict__ sizes, uint32_t *bad)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = refs[i];
    uint32_t size = 0;
    if (r + 3 > n_buckets) *bad = 2;
    else {
        size = arena[r + 2];
        if (r + 3 + size > n_buckets) { *bad = 2; size = 0; }
    }
    sizes[i] = size;
}

__global__ void k_ingest_copy(const uint32_t *__restrict__ arena, const uint64_t *__restrict__ refs, uint32_t n,
                              const uint32_t *__restrict__ offs, uint4 *__restrict__ hdr, uint32_t *__restrict__ lits)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = refs[i];
    const uint32_t w0 = arena[r], sig = arena[r + 1], size = arena[r + 2], off = offs[i];
    hdr[i] = make_uint4(off, size, w0, sig);
    const uint32_t *src = arena + r + 3;
    uint32_t *dst = lits + off;
    for (uint32_t k = 0; k < size; ++k) dst[k] = src[k];
}

/* fast path: awaken() allocates the clauses back to back in refs order (simplify.cpp:118-133), so the
 * pool offset is refs[i] - 3*i and no scan is needed; any gap raises *gap and the driver re-runs the
 * general path above */
__global__ void k_ingest_fused(const uint32_t *__restrict__ arena, const uint64_t *__restrict__ refs, uint32_t n,
                               uint64_t n_buckets, uint4 *__restrict__ hdr, uint32_t *__restrict__ lits, uint32_t *gap)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = refs[i];
    if (r + 3 > n_buckets) { *gap = 1; return; }            /* the general path reports malformed input */
    const uint32_t w0 = arena[r], sig = arena[r + 1], size = arena[r + 2];
    if (r + 3 + size > n_buckets) { *gap = 1; return; }
    const uint64_t expect = i ? r : 0;
    if ((i + 1 < n && refs[i + 1] != r + 3 + size) || (i + 1 == n && r + 3 + size != n_buckets) || (i == 0 && r != expect) || r < 3ull * i) { *gap = 1; return; }
    const uint32_t off = (uint32_t)(r - 3ull * i);
    hdr[i] = make_uint4(off, size, w0, sig);
    const uint32_t *src = arena + r + 3;
    uint32_t *dst = lits + off;
    for (uint32_t k = 0; k < size; ++k) dst[k] = src[k];
}

/* ---------------------------------------------------------------------------------------------- */
/* exclusive scan (reduce-then-scan, 2048 elements per block)                                       */
/* ---------------------------------------------------------------------------------------------- */
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = TPB * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t *total)
{   /* exclusive scan of one value per thread over a 256-thread block */
    __shared__ uint32_t warp_sums[TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < TPB / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < TPB / 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += y;
        }
        if (lane < TPB / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t base = warp ? warp_sums[warp - 1] : 0;
    if (total) *total = warp_sums[TPB / 32 - 1];
    __syncthreads();
    return base + incl - x;
}

__global__ void k_scan_reduce(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ partial)
{
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * TPB + threadIdx.x;
        if (i < n) s += in[i];
    }
    __shared__ uint32_t tot;
    (void)block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void k_scan_partials(uint32_t *__restrict__ partial, uint32_t nb, uint32_t *__restrict__ total)
{   /* one block: exclusive scan of the block sums, in chunks of 256 */
    __shared__ uint32_t carry, tot;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += TPB) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = i < nb ? partial[i] : 0;
        const uint32_t e = block_exclusive_scan(x, &tot);
        if (i < nb) partial[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void k_scan_down(const uint32_t *in, uint32_t *out, uint32_t n, const uint32_t *__restrict__ partial)
{
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t x[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + k;
        x[k] = i < n ? in[i] : 0;
        s += x[k];
    }
    uint32_t e = block_exclusive_scan(s, nullptr) + partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + k;
        if (i < n) out[i] = e;
        e += x[k];
    }
}

/* up to four independent scans of the same length in the same three launches: blockIdx.y selects the array */
__global__ void k_scan_reduce_multi(SgScanSet ss, uint32_t n, uint32_t nb, uint32_t *__restrict__ partial)
{
    const uint32_t *in = ss.in[blockIdx.y];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * TPB + threadIdx.x;
        if (i < n) s += in[i];
    }
    __shared__ uint32_t tot;
    (void)block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * nb + blockIdx.x] = tot;
}
__global__ void k_scan_partials_multi(SgScanSet ss, uint32_t nb, uint32_t *__restrict__ partial_all)
{
    uint32_t *partial = partial_all + (size_t)blockIdx.x * nb;
    uint32_t *total = ss.total[blockIdx.x];
    __shared__ uint32_t carry, tot;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += TPB) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = i < nb ? partial[i] : 0;
        const uint32_t e = block_exclusive_scan(x, &tot);
        if (i < nb) partial[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void k_scan_down_multi(SgScanSet ss, uint32_t n, uint32_t nb, const uint32_t *__restrict__ partial)
{
    const uint32_t *in = ss.in[blockIdx.y];
    uint32_t *out = ss.out[blockIdx.y];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t x[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + k;
        x[k] = i < n ? in[i] : 0;
        s += x[k];
    }
    uint32_t e = block_exclusive_scan(s, nullptr) + partial[(size_t)blockIdx.y * nb + blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const uint64_t i = base + k;
        if (i < n) out[i] = e;
        e += x[k];
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* occurrence table build                                                                           */
/* ---------------------------------------------------------------------------------------------- */
__global__ void k_hist(const uint4 *__restrict__ hdr, const uint32_t *__restrict__ lits, uint32_t n_cls,
                       uint32_t *__restrict__ counts)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cls) return;
    const uint4 h = hdr[c];
    if (h.z & SG_DELETED) return;
    const uint32_t *l = lits + h.x;
    for (uint32_t k = 0; k < h.y; ++k) atomicAdd(&counts[l[k]], 1u);
}

/* One pass writes the lists of the literals in [lo, hi) only: the slice of ot_data it touches
 * stays resident in L2, so the 4-byte scattered stores merge there instead of each costing a
 * 32-byte sector read + write in HBM (ncu: 3.9 GB of DRAM traffic for 0.48 GB of payload when the
 * whole 204 MB table was scattered in one pass). */
template <bool STREAM_IN>
__global__ void k_scatter(const uint4 *__restrict__ hdr, const uint32_t *__restrict__ lits, uint32_t n_cls,
                          uint32_t *__restrict__ cursor, uint32_t *__restrict__ ot_data, uint32_t lo, uint32_t hi)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cls) return;
    /* STREAM_IN: the store is read once per pass - evict-first loads keep the slice of the table resident in L2 */
    const uint4 h = STREAM_IN ? __ldcs(hdr + c) : hdr[c];
    if (h.z & SG_DELETED) return;
    const uint32_t *l = lits + h.x;
    for (uint32_t base = 0; base < h.y; base += 8) {
        /* issue the (independent) cursor atomics of up to 8 literals before any dependent store */
        uint32_t pos[8];
        bool take[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            take[k] = false;
            if (base + k < h.y) {
                const uint32_t lit = STREAM_IN ? __ldcs(l + base + k) : l[base + k];
                take[k] = lit >= lo && lit < hi;
                if (take[k]) pos[k] = atomicAdd(&cursor[lit], 1u);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) if (take[k]) ot_data[pos[k]] = c;
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* occurrence table build as a two-level partition (the default createOT)                           */
/* ---------------------------------------------------------------------------------------------- */
/* createOT (simplify.cpp:39-61) is a stable counting sort of (literal, clause) pairs by literal.  With millions of
 * literal bins a one-level scatter costs one global atomic per literal plus a 4-byte store at a random place (round 1:
 * k_hist + 4 x k_scatter + order fix, 2.0 GB of DRAM traffic for 0.49 GB of payload).  Here the literal space is cut
 * into NB buckets of W = 2^shift consecutive literals, chosen so that one bucket's share of the table fits in shared
 * memory, and the clause ids are cut into T tiles (a tile's literals fit in shared memory as well):
 *   k_otb_count    tile t counts its literals per bucket in shared memory            -> cnt[b][t]        (reads the store once)
 *   scan           exclusive scan of cnt in (bucket, tile) order = where tile t's records of bucket b go
 *   k_otb_scatter  tile t sorts its literals by bucket in shared memory (4-byte records: literal within the bucket |
 *                  clause within the tile << shift) and writes every bucket's run with consecutive lanes: whole sectors
 *                  instead of one partial sector per literal (ncu, 8-byte records stored one by one: 0.8 GB read + 0.8 GB
 *                  written in DRAM for 0.36 + 0.41 GB, every sector filled before and written back after its last store)
 *   k_otb_lists    one block per bucket: the bucket's records are T runs, one warp per run (the run number is the tile,
 *                  which gives the clause id back); per-literal counts (= ot_len), their scan (= ot_off: the bucket's
 *                  records and its slice of ot_data share one index range), clause ids placed in shared memory, every
 *                  list put in ascending clause id order there (= the reference's push order; the warps walk the runs
 *                  in tile order, so the lists are nearly sorted already), one bulk (TMA) store of the bucket's slice
 * No global atomics, the store is read twice instead of five times, every byte of ot_data is written once, coalesced.
 * A tile or a bucket whose share exceeds the staging buffer (a literal with tens of thousands of occurrences) is
 * written / placed and sorted in global memory instead. */
constexpr int OTB_TPB = 1024;
constexpr uint32_t OTB_SMEM_WORDS = 55000; /* dynamic shared memory of the scatter and the lists kernel (220 000 bytes) */
constexpr uint32_t OTB_WMAX = 4096;        /* literals per bucket: two shared arrays of that many words */
constexpr uint32_t OTB_NBMAX = 4096;       /* buckets: three shared words each in the scatter kernel */
constexpr uint32_t OTB_TMAX = 148 * 32;    /* tiles: the run bounds of one bucket sit in shared memory in the lists kernel */
constexpr uint32_t OTB_PAD = 0xFFFFFFFFu;

/* block-wide exclusive scan of one value per thread (1024 threads); s_part: 32 words */
__device__ __forceinline__ uint32_t otb_block_scan(uint32_t x, uint32_t *s_part, uint32_t *total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += y; }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_part[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (uint32_t)d) w += y; }
        s_part[lane] = w;
    }
    __syncthreads();
    const uint32_t e = (warp ? s_part[warp - 1] : 0) + incl - x;
    if (total) *total = s_part[31];
    __syncthreads();
    return e;
}

/* The records of a tile (scatter) or of a bucket (lists) are a sequence of runs, given by n + 1 ascending bounds; runs may
 * be empty.  otb_compact_runs keeps the non-empty ones: starts[0 .. NR] strictly ascending (starts[NR] = bounds[n]) and their
 * numbers ids[0 .. NR).  Whole block; `starts` / `ids` must not alias `bounds`. */
__device__ __forceinline__ uint32_t otb_compact_runs(const uint32_t *bounds, uint32_t n, uint32_t *starts, uint32_t *ids, uint32_t *s_part)
{
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x, i0 = threadIdx.x * per;
    uint32_t cnt = 0;
    for (uint32_t k = 0; k < per; ++k) { const uint32_t i = i0 + k; if (i < n && bounds[i + 1] > bounds[i]) ++cnt; }
    uint32_t NR;
    uint32_t e = otb_block_scan(cnt, s_part, &NR);
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t i = i0 + k;
        if (i < n && bounds[i + 1] > bounds[i]) { starts[e] = bounds[i]; ids[e] = i; ++e; }
    }
    if (threadIdx.x == 0) starts[NR] = bounds[n];
    __syncthreads();
    return NR;
}
/* the run r with starts[r] <= p < starts[r + 1], searched by the whole warp: 32 probes per round */
__device__ __forceinline__ uint32_t otb_locate_warp(const uint32_t *starts, uint32_t n, uint32_t p)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t step = (hi - lo + 31) >> 5;
        const uint32_t idx = lo + (lane + 1) * step;
        const bool le = idx < hi && starts[idx] <= p;
        const uint32_t k = __popc(__ballot_sync(0xffffffffu, le));
        lo += k * step;
        if (lo + step < hi) hi = lo + step;
    }
    return lo;
}
/* A warp at the 32 consecutive positions p0 .. p0 + 31, p0 inside run rc (starts[rc] <= p0 <= starts[rc + 1]): returns the run
 * of this lane's position and moves rc on to the run of p0 + 32.  At most 32 runs can start inside 32 positions, so one load
 * per lane and one warp-wide OR see them all. */
__device__ __forceinline__ uint32_t otb_group_runs(const uint32_t *starts, uint32_t NR, uint32_t &rc, uint32_t p0)
{
    const uint32_t lane = threadIdx.x & 31, k = rc + 1 + lane;
    uint32_t bit = 0;
    if (k <= NR) { const uint32_t d = starts[k] - p0; if (d < 32u) bit = 1u << d; }
    const uint32_t mask = __reduce_or_sync(0xffffffffu, bit);
    const uint32_t mine = rc + __popc(mask & (0xFFFFFFFFu >> (31u - lane)));
    rc += __popc(mask);
    return mine;
}

/* The tile kernels run as a persistent grid: block g takes tiles g, g + G, g + 2G, ... in order, so that at any time the
 * tiles in flight are neighbours, and so are their runs inside every bucket: the sectors two neighbouring runs share are
 * completed in L2. */
__global__ void __launch_bounds__(OTB_TPB, 2) k_otb_count(const uint4 *__restrict__ hdr, const uint32_t *__restrict__ lits, uint32_t n_cls,
                                                           uint32_t shift, uint32_t NB, uint32_t T, uint32_t per_tile,
                                                           uint32_t *__restrict__ cnt, uint32_t *__restrict__ bits)
{
    extern __shared__ uint32_t s_cnt[];
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[(size_t)NB * T] = 0;          /* the scan's last word: the total */
    for (uint32_t t = blockIdx.x; t < T; t += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < NB; b += blockDim.x) s_cnt[b] = 0;
        __syncthreads();
        const uint32_t c0 = t * per_tile, c1 = min(n_cls, c0 + per_tile);
        /* per_tile is a multiple of 32, so a warp's 32 clauses share one word of the liveness bitmap */
        for (uint32_t base = c0; base < c1; base += 2 * blockDim.x) {
            uint4 h[2];
            bool live[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t c = base + j * blockDim.x + threadIdx.x;
                live[j] = false;
                if (c < c1) { h[j] = __ldcs(hdr + c); live[j] = !(h[j].z & SG_DELETED); }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t c = base + j * blockDim.x + threadIdx.x;
                if (live[j]) {
                    const uint32_t *l = lits + h[j].x;
                    for (uint32_t k0 = 0; k0 < h[j].y; k0 += 8) {
                        uint32_t lit[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) lit[k] = k0 + k < h[j].y ? __ldcs(l + k0 + k) : 0u;
#pragma unroll
                        for (int k = 0; k < 8; ++k) if (k0 + k < h[j].y) atomicAdd(&s_cnt[lit[k] >> shift], 1u);
                    }
                }
                if (bits) {
                    const uint32_t m = __ballot_sync(0xffffffffu, live[j]);
                    if ((threadIdx.x & 31) == 0 && (c & ~31u) < c1) bits[c >> 5] = m;
                }
            }
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < NB; b += blockDim.x) cnt[(size_t)b * T + t] = s_cnt[b];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(OTB_TPB, 1) k_otb_scatter(const uint4 *__restrict__ hdr, const uint32_t *__restrict__ lits, uint32_t n_cls,
                                                             uint32_t shift, uint32_t NB, uint32_t T, uint32_t per_tile, uint32_t rcap,
                                                             const uint32_t *__restrict__ base_of, uint32_t *__restrict__ recs)
{
    extern __shared__ __align__(128) uint32_t s_mem[];
    __shared__ uint32_t s_part[32];
    __shared__ uint32_t s_total;
    uint32_t *s_off = s_mem;                       /* NB + 1: start of bucket b's run in the staged tile */
    uint32_t *s_cur = s_mem + NB + 4;              /* NB: fill cursors */
    uint32_t *s_gb = s_mem + 2 * NB + 8;           /* NB: where the run goes in recs */
    uint32_t *s_ids = s_mem + 3 * NB + 12;         /* NB: non-empty buckets of the tile (their starts reuse s_cur) */
    uint32_t *stage = s_mem + 4 * NB + 16;         /* rcap records */
    const uint32_t per = (NB + blockDim.x - 1) / blockDim.x;                    /* <= 4 */
    const uint32_t lmask = (1u << shift) - 1u;
    for (uint32_t t = blockIdx.x; t < T; t += gridDim.x) {
        uint32_t g[4], n[4], sum = 0;
        const uint32_t b0 = threadIdx.x * per;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = b0 + k;
            n[k] = 0; g[k] = 0;
            if ((uint32_t)k < per && b < NB) {
                const size_t i = (size_t)b * T + t;
                g[k] = base_of[i];
                n[k] = base_of[i + 1] - g[k];
                sum += n[k];
            }
        }
        uint32_t e = otb_block_scan(sum, s_part, &s_total);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = b0 + k;
            if ((uint32_t)k < per && b < NB) { s_off[b] = e; s_cur[b] = e; s_gb[b] = g[k]; e += n[k]; }
        }
        if (threadIdx.x == 0) s_off[NB] = s_total;
        __syncthreads();
        const uint32_t R = s_total;
        const bool staged = R <= rcap;
        const uint32_t c0 = t * per_tile, c1 = min(n_cls, c0 + per_tile);
        for (uint32_t base = c0; base < c1; base += 4 * blockDim.x) {
            /* four clauses per thread: their headers, then the first eight literals of each, in flight together */
            uint4 h[4];
            uint32_t lit[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t c = base + j * blockDim.x + threadIdx.x;
                h[j] = c < c1 ? __ldcs(hdr + c) : make_uint4(0u, 0u, SG_DELETED, 0u);
                if (h[j].z & SG_DELETED) h[j].y = 0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 8; ++k) lit[j][k] = (uint32_t)k < h[j].y ? __ldcs(lits + h[j].x + k) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t cl = (base + j * blockDim.x + threadIdx.x - c0) << shift;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((uint32_t)k < h[j].y) {
                        const uint32_t b = lit[j][k] >> shift;
                        const uint32_t slot = atomicAdd(&s_cur[b], 1u);
                        const uint32_t rec = (lit[j][k] & lmask) | cl;
                        if (staged) stage[slot] = rec; else recs[s_gb[b] + (slot - s_off[b])] = rec;
                    }
                for (uint32_t k = 8; k < h[j].y; ++k) {
                    const uint32_t x = __ldcs(lits + h[j].x + k);
                    const uint32_t b = x >> shift;
                    const uint32_t slot = atomicAdd(&s_cur[b], 1u);
                    const uint32_t rec = (x & lmask) | cl;
                    if (staged) stage[slot] = rec; else recs[s_gb[b] + (slot - s_off[b])] = rec;
                }
            }
        }
        __syncthreads();
        if (staged) {
            /* consecutive lanes, consecutive staged records -> consecutive words of a run in recs; which run: the non-empty
             * buckets compacted, then a mask of the run starts inside each group of 32 records */
            const uint32_t NR = otb_compact_runs(s_off, NB, s_cur, s_ids, s_part);
            const uint32_t lane = threadIdx.x & 31, nw = blockDim.x >> 5;
            for (uint32_t i0 = (threadIdx.x >> 5) * 128; i0 < R; i0 += nw * 128) {
                uint32_t rc = otb_locate_warp(s_cur, NR, i0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t i = i0 + q * 32 + lane;
                    const uint32_t r = otb_group_runs(s_cur, NR, rc, i0 + q * 32);
                    if (i < R) recs[s_gb[s_ids[r]] + (i - s_cur[r])] = stage[i];
                }
            }
        }
        __syncthreads();
    }
}

/* ascending order of one list, by one thread: nearly sorted input (see above), so insertion; a long list that really is
 * out of order goes through the heap sort instead of a quadratic walk */
__device__ __forceinline__ void otb_sort_list(uint32_t *d, uint32_t n)
{
    if (n < 2) return;
    if (n > 64) {
        uint32_t inv = 0;
        for (uint32_t i = 1; i < n && inv <= n; ++i) inv += d[i - 1] > d[i];
        if (inv == 0) return;
        if (inv > 8) { sg_heapsort_u32(d, (int)n); return; }
    }
    for (uint32_t i = 1; i < n; ++i) {
        const uint32_t x = d[i];
        uint32_t j = i;
        while (j > 0 && d[j - 1] > x) { d[j] = d[j - 1]; --j; }
        d[j] = x;
    }
}

/* ascending order of one list by a whole warp: bitonic network in its "flip / disperse" form, where every compare-exchange
 * moves the smaller key to the lower index - so the keys a power-of-two length would need beyond n can be imagined as
 * +infinity and never touched.  n log^2 n / 64 steps per lane whatever the input order: the clauses of one gate sit in
 * one tile, and the scatter leaves the records of a tile in no particular order, which makes a thread's insertion sort
 * quadratic on exactly the long lists (XOR gates: 64-96 entries per literal, 70 us per block on the 1M-clause XOR/PHP mix). */
constexpr uint32_t OTB_WARP_SORT = 32;      /* lists longer than this are sorted by a warp */
__device__ __forceinline__ void otb_sort_list_warp(uint32_t *d, uint32_t n)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lm = 1;
    while ((1u << lm) < n) ++lm;
    const uint32_t half = 1u << (lm - 1);
    for (uint32_t lk = 1; lk <= lm; ++lk) {
        const uint32_t k = 1u << lk;
        for (uint32_t t = lane; t < half; t += 32) {       /* flip: i against its mirror inside the block of k */
            const uint32_t o = t & ((k >> 1) - 1u), b0 = (t >> (lk - 1)) << lk;
            const uint32_t i = b0 + o, l = b0 + (k - 1u - o);
            if (l < n) { const uint32_t a = d[i], b = d[l]; if (a > b) { d[i] = b; d[l] = a; } }
        }
        __syncwarp();
        for (int lj = (int)lk - 2; lj >= 0; --lj) {        /* disperse: i against i + j inside blocks of 2j */
            const uint32_t j = 1u << lj;
            for (uint32_t t = lane; t < half; t += 32) {
                const uint32_t i = ((t >> lj) << (lj + 1)) + (t & (j - 1u)), l = i + j;
                if (l < n) { const uint32_t a = d[i], b = d[l]; if (a > b) { d[i] = b; d[l] = a; } }
            }
            __syncwarp();
        }
    }
}

/* the records of one bucket, 128 consecutive ones per warp and step (the warps walk the bucket side by side, so the
 * records of a list are met nearly in tile order); f(record, tile).  The tile of a record is the run it lies in. */
template <bool LAST_USE, class F>
__device__ __forceinline__ void otb_for_records(const uint32_t *__restrict__ recs, const uint32_t *s_st, const uint32_t *s_tile, uint32_t NR, uint32_t P, F f)
{
    const uint32_t lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const uint32_t g0 = s_st[0];
    for (uint32_t i0 = (threadIdx.x >> 5) * 128; i0 < P; i0 += nw * 128) {
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = i0 + q * 32 + lane;
            r[q] = i < P ? (LAST_USE ? __ldcs(recs + g0 + i) : __ldg(recs + g0 + i)) : OTB_PAD;
        }
        uint32_t rc = otb_locate_warp(s_st, NR, g0 + i0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t run = otb_group_runs(s_st, NR, rc, g0 + i0 + q * 32);
            if (i0 + q * 32 + lane < P) f(r[q], s_tile[run]);
        }
    }
}

__global__ void __launch_bounds__(OTB_TPB, 1) k_otb_lists(const uint32_t *__restrict__ recs, const uint32_t *__restrict__ base_of, uint32_t NB,
                                                          uint32_t T, uint32_t shift, uint32_t per_tile, uint32_t pcap, uint32_t nlit,
                                                          uint32_t *__restrict__ ot_off, uint32_t *__restrict__ ot_len, uint32_t *__restrict__ ot_data)
{
    extern __shared__ __align__(128) uint32_t s_mem[];
    __shared__ uint32_t s_part[32];
    const uint32_t W = 1u << shift, lmask = W - 1u;
    uint32_t *s_loc = s_mem;                 /* W + 1: exclusive offsets of the bucket's lists */
    uint32_t *s_cur = s_mem + W + 4;         /* W: fill cursors */
    const uint32_t Tq = (T + 1 + 3) & ~3u;
    uint32_t *s_tb = s_mem + 2 * W + 8;      /* T + 1: run bounds */
    uint32_t *s_st = s_tb + Tq;              /* T + 1: starts of the non-empty runs */
    uint32_t *s_tile = s_st + Tq;            /* T: their tiles */
    uint32_t *stage = s_tile + Tq;           /* pcap + 4 staged clause ids (16-byte aligned) */
    const uint32_t b = blockIdx.x;
    for (uint32_t i = threadIdx.x; i <= T; i += blockDim.x) s_tb[i] = base_of[(size_t)b * T + i];
    for (uint32_t i = threadIdx.x; i < W; i += blockDim.x) s_cur[i] = 0;
    __syncthreads();
    const uint32_t g0 = s_tb[0], P = s_tb[T] - g0, lit0 = b << shift;
    const uint32_t NR = otb_compact_runs(s_tb, T, s_st, s_tile, s_part);
    if (P) otb_for_records<false>(recs, s_st, s_tile, NR, P, [&](uint32_t r, uint32_t) { atomicAdd(&s_cur[r & lmask], 1u); });
    __syncthreads();
    /* exclusive scan of the W counts: each thread owns W / blockDim consecutive literals */
    const uint32_t per = (W + blockDim.x - 1) / blockDim.x;
    const uint32_t i0 = threadIdx.x * per;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < per; ++k) if (i0 + k < W) sum += s_cur[i0 + k];
    uint32_t e = otb_block_scan(sum, s_part, nullptr);
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t i = i0 + k;
        if (i < W) {
            const uint32_t n = s_cur[i];
            s_loc[i] = e;
            if (lit0 + i < nlit) { ot_off[lit0 + i] = g0 + e; ot_len[lit0 + i] = n; }
            e += n;
        }
    }
    if (threadIdx.x == blockDim.x - 1) s_loc[W] = P;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < W; i += blockDim.x) s_cur[i] = s_loc[i];
    __syncthreads();
    /* The bucket's slice of ot_data is assembled in shared memory and leaves with one bulk store.  A bucket that does not fit
     * (the literals of one bucket can hold far more than the average: XOR gates next to pigeon-hole clauses) is taken as
     * consecutive ranges of its lists that do fit, each with its own walk over the bucket's records (L2 hits); only a single
     * list longer than the staging room is placed and sorted in global memory. */
    __shared__ uint32_t s_hi;
    uint32_t lo = 0;
    while (lo < W) {
        if (threadIdx.x == 0) {
            uint32_t a = lo + 1, z = W;                    /* largest hi with s_loc[hi] - s_loc[lo] <= pcap, at least lo + 1 */
            const uint32_t lim = s_loc[lo] + pcap;
            while (a < z) { const uint32_t mid = (a + z + 1) >> 1; if (s_loc[mid] <= lim) a = mid; else z = mid - 1; }
            s_hi = a;
        }
        __syncthreads();
        const uint32_t hi = s_hi, base = s_loc[lo], n = s_loc[hi] - base;
        const bool whole = lo == 0 && hi == W;
        if (n) {
            const bool staged = n <= pcap;
            const uint32_t shift4 = (g0 + base) & 3u;
            uint32_t *dst = staged ? stage + shift4 - base : ot_data + g0;       /* indexed by the offset inside the bucket */
            /* second read of the bucket's records: L2 hits */
            if (whole) otb_for_records<true>(recs, s_st, s_tile, NR, P, [&](uint32_t r, uint32_t t) { dst[atomicAdd(&s_cur[r & lmask], 1u)] = t * per_tile + (r >> shift); });
            else otb_for_records<false>(recs, s_st, s_tile, NR, P, [&](uint32_t r, uint32_t t) {
                const uint32_t l = r & lmask;
                if (l >= lo && l < hi) dst[atomicAdd(&s_cur[l], 1u)] = t * per_tile + (r >> shift);
            });
            __syncthreads();
            if (!staged) __threadfence_block();
            /* every list in ascending clause id order: short ones one thread each, the others one warp each */
            bool any_long = false;
            for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                const uint32_t len = s_loc[i + 1] - s_loc[i];
                uint32_t *d = dst + s_loc[i];
                bool by_warp = false;
                if (staged && len > OTB_WARP_SORT) {       /* nearly sorted (the usual case: one entry per tile) stays with the thread */
                    uint32_t inv = 0;
                    for (uint32_t q = 1; q < len && inv <= 4; ++q) inv += d[q - 1] > d[q];
                    by_warp = inv > 4;
                }
                s_cur[i] = by_warp;                        /* the fill cursors are not needed any more */
                if (by_warp) any_long = true; else otb_sort_list(d, len);
            }
            if (staged && __syncthreads_or(any_long))
                for (uint32_t i = lo + (threadIdx.x >> 5); i < hi; i += blockDim.x >> 5)
                    if (s_cur[i]) otb_sort_list_warp(dst + s_loc[i], s_loc[i + 1] - s_loc[i]);
            if (staged) {
                stage_fence();
                __syncthreads();
                flush_staged(ot_data, (unsigned long long)g0 + base, stage, n);
            }
        }
        __syncthreads();                                   /* the staging room and s_hi are free again */
        lo = hi;
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* variable ordering: scores + stable LSD radix sort (8-bit digits)                                  */
/* ---------------------------------------------------------------------------------------------- */
__global__ void k_scores(SgView v)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x + 1;
    uint32_t sc = 0;
    if (x <= v.nvars) {
        sc = v.occ[2 * x] * v.occ[2 * x + 1];
        v.scores[x] = sc;
        v.order[x - 1] = x;
    }
    /* largest score: the radix sort only needs the digit passes that can differ */
    /* one atomic per block (tens of thousands of same-address atomics cost more than the rest of the kernel) */
    __shared__ uint32_t s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const uint32_t m = __reduce_max_sync(0xffffffffu, sc);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_max, m);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(&v.sc->max_score, s_max);
}

constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = TPB * RS_ITEMS;

__global__ void k_rs_hist(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t nblocks, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * TPB + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void k_rs_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n, int shift,
                             uint32_t nblocks, const uint32_t *__restrict__ hist, uint32_t *__restrict__ okeys,
                             uint32_t *__restrict__ ovals)
{
    __shared__ uint32_t base[256];
    __shared__ uint32_t wcnt[TPB / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
    for (int w = 0; w < TPB / 32; ++w) wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint64_t tile = (uint64_t)blockIdx.x * RS_TILE;
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint64_t i = tile + (uint64_t)k * TPB + threadIdx.x;
        const bool valid = i < n;
        const uint32_t key = valid ? keys[i] : 0xFFFFFFFFu;
        const uint32_t val = valid ? vals[i] : 0;
        const uint32_t d = valid ? ((key >> shift) & 255u) : 256u;   /* 256 = no digit */
        const uint32_t mask = __match_any_sync(0xffffffffu, d);
        const uint32_t before = __popc(mask & ((1u << lane) - 1u));
        if (valid && before == 0) wcnt[warp][d] = __popc(mask);
        __syncthreads();
        if (valid) {
            uint32_t off = base[d] + before;
            for (int w = 0; w < warp; ++w) off += wcnt[w][d];
            okeys[off] = key;
            ovals[off] = val;
        }
        __syncthreads();
        {
            uint32_t s = 0;
            for (int w = 0; w < TPB / 32; ++w) { s += wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = 0; }
            base[threadIdx.x] += s;
        }
        __syncthreads();
    }
}

__global__ void k_rank(SgView v)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < v.nvars) v.rank[v.order[r]] = r;
}

/* The whole variable ordering in ONE cooperative launch: scores (histogram.hpp:23-41), stable LSD radix sort of
 * (score, variable) with 8-bit digits, rank[] (lcve.cpp:39 = wolfsort(..., STABLE_LCV): score ascending, ties by variable).
 * The separate-launch version above costs 12 launches per election for 16 MB of keys - launch latency, not bandwidth.
 * Block b owns the same contiguous chunk of positions in every pass and warp w of the block a contiguous slice of it, so
 * "stable" = (block, warp, position in the slice).  Per pass: every warp counts the digits of its slice in its own row of
 * a shared [32][256] table | block totals to global memory | grid barrier | every block sums the totals of the blocks before
 * it (37 888 words in L2) and turns its table into start offsets per (warp, digit) | every warp walks its slice 32 keys at a
 * time: __match_any_sync ranks the lanes of one digit, the row entry moves on - no block barrier inside the walk | grid
 * barrier.  The number of passes comes from the largest score, which the first phase leaves in the scalars
 * (forced_bits != 0: the caller's width, a test hook for the too-narrow-guess path of the separate-launch version). */
constexpr int VO_TPB = 1024;
__global__ void __launch_bounds__(VO_TPB, 1) k_vo_sort(SgView v, uint32_t *tmpk, uint32_t *tmpv, uint32_t *hist, uint32_t forced_bits)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ uint32_t s_dyn[];
    uint32_t (*s_cnt)[256] = (uint32_t (*)[256])s_dyn;                 /* [32][256] digit counts, then start offsets, per warp */
    uint32_t (*s_tmp)[256] = (uint32_t (*)[256])(s_dyn + 32 * 256);    /* [8][256] partial sums of the offsets phase */
    __shared__ uint32_t s_base[256], s_wsum[8];
    const uint32_t n = v.nvars, G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t chunk = (((n + G - 1) / G) + VO_TPB - 1) / VO_TPB * VO_TPB, slice = chunk / 32;     /* slice: a multiple of 32 */
    const uint64_t lo64 = (uint64_t)b * chunk + (uint64_t)warp * slice;
    const uint32_t w_lo = lo64 < n ? (uint32_t)lo64 : n, w_hi = (lo64 + slice < n) ? (uint32_t)(lo64 + slice) : n;
    uint32_t *keys = v.scores + 1, *vals = v.order;
    const uint2 *occ2 = (const uint2 *)v.occ;                          /* occ[2x], occ[2x + 1] */
    for (uint32_t i = tid; i < 32 * 256; i += VO_TPB) s_dyn[i] = 0;
    __syncthreads();
    /* scores, identity order, digit-0 counts, largest score (four steps of the slice in flight) */
    uint32_t m = 0;
    for (uint32_t i0 = w_lo + lane; i0 < w_hi; i0 += 128) {
        uint2 o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const uint32_t i = i0 + 32 * q; o[q] = i < w_hi ? __ldg(occ2 + i + 1) : make_uint2(0u, 0u); }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = i0 + 32 * q;
            if (i < w_hi) {
                const uint32_t sc = o[q].x * o[q].y;
                keys[i] = sc; vals[i] = i + 1;
                m = sc > m ? sc : m;
                atomicAdd(&s_cnt[warp][sc & 255u], 1u);
            }
        }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0 && m) atomicMax(&v.sc->max_score, m);
    uint32_t bits = forced_bits;
    uint32_t *ik = keys, *iv = vals, *ok = tmpk, *ov = tmpv;
    for (uint32_t shift = 0; shift == 0 || shift < bits; shift += 8) {
        if (shift) {
            for (uint32_t i = tid; i < 32 * 256; i += VO_TPB) s_dyn[i] = 0;
            __syncthreads();
            for (uint32_t i0 = w_lo + lane; i0 < w_hi; i0 += 128) {
                uint32_t k[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const uint32_t i = i0 + 32 * q; k[q] = i < w_hi ? __ldcg(ik + i) : 0u; }
#pragma unroll
                for (int q = 0; q < 4; ++q) if (i0 + 32 * q < w_hi) atomicAdd(&s_cnt[warp][(k[q] >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t t = 0;
#pragma unroll 8
            for (int w = 0; w < 32; ++w) t += s_cnt[w][tid];
            hist[tid * G + b] = t;
        }
        grid.sync();
        if (!shift && !bits) { const uint32_t mx = *(volatile uint32_t *)&v.sc->max_score; bits = 8; while (bits < 32 && (mx >> bits)) bits += 8; }
        /* where this block's keys of digit d start: all keys of smaller digits + the keys of digit d in earlier blocks.
         * Four threads per digit, each a quarter of the blocks, eight loads in flight */
        {
            const uint32_t d = tid & 255u, part = tid >> 8, per = (G + 3) / 4;
            const uint32_t k0 = part * per, k1 = k0 + per < G ? k0 + per : G;
            const uint32_t *h = hist + d * G;
            uint32_t tot = 0, pre = 0;
            for (uint32_t k = k0; k < k1; k += 8) {
                uint32_t x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = k + q < k1 ? __ldcg(h + k + q) : 0u;
#pragma unroll
                for (int q = 0; q < 8; ++q) { tot += x[q]; pre += k + q < b ? x[q] : 0u; }
            }
            s_tmp[part][d] = tot; s_tmp[4 + part][d] = pre;
        }
        __syncthreads();
        uint32_t tot = 0, pre = 0;
        if (tid < 256) {
            tot = s_tmp[0][tid] + s_tmp[1][tid] + s_tmp[2][tid] + s_tmp[3][tid];
            pre = s_tmp[4][tid] + s_tmp[5][tid] + s_tmp[6][tid] + s_tmp[7][tid];
        }
        uint32_t incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += y; }
        if (tid < 256 && lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (tid < 256) {
            uint32_t run = incl - tot + pre;
            for (uint32_t w = 0; w < warp; ++w) run += s_wsum[w];
            /* counts -> start offsets per (warp of the block, digit) */
#pragma unroll 8
            for (int w = 0; w < 32; ++w) { const uint32_t c = s_cnt[w][tid]; s_cnt[w][tid] = run; run += c; }
        }
        __syncthreads();
        /* the warp's slice in order, 32 keys per step, the next step's keys fetched while these are ranked */
        uint32_t nkey = 0, nval = 0;
        if (w_lo + lane < w_hi) { nkey = __ldcg(ik + w_lo + lane); nval = __ldcg(iv + w_lo + lane); }
        for (uint32_t base = w_lo; base < w_hi; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < w_hi;
            const uint32_t key = nkey, val = nval;
            if (i + 32 < w_hi) { nkey = __ldcg(ik + i + 32); nval = __ldcg(iv + i + 32); }
            const uint32_t d = valid ? ((key >> shift) & 255u) : 256u;
            const uint32_t mask = __match_any_sync(0xffffffffu, d);
            const uint32_t before = __popc(mask & ((1u << lane) - 1u));
            uint32_t o = 0;
            if (valid) o = s_cnt[warp][d] + before;
            __syncwarp();
            if (valid && before == 0) s_cnt[warp][d] += __popc(mask);
            __syncwarp();
            if (valid) { ok[o] = key; ov[o] = val; }
        }
        grid.sync();
        uint32_t *t = ik; ik = ok; ok = t;
        t = iv; iv = ov; ov = t;
    }
    for (uint32_t i0 = w_lo + lane; i0 < w_hi; i0 += 128) {
        uint32_t x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const uint32_t i = i0 + 32 * q; x[q] = i < w_hi ? __ldcg(iv + i) : 0u; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = i0 + 32 * q;
            if (i < w_hi) { if (iv != vals) vals[i] = x[q]; v.rank[x[q]] = i; }
        }
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* election                                                                                         */
/* ---------------------------------------------------------------------------------------------- */
__global__ void k_elect_batch(SgView v, uint32_t r0, uint32_t r1, uint32_t *__restrict__ wl)
{
    const uint32_t r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r1) return;
    const uint32_t var = v.order[r];
    if (SG_ER_STAT(v.rank[var]) != SG_E_UNDECIDED) return;
    wl[atomicAdd(&v.sc->wl_count[0], 1u)] = var;
}

__global__ void k_elect_round(SgView v, const uint32_t *__restrict__ wl_in, uint32_t *__restrict__ wl_out, int parity)
{   /* one warp per queued candidate; still-undecided candidates are re-queued through a per-warp buffer so
     * that the single queue counter sees one atomic per 32 candidates (same-address atomics serialise in L2) */
    __shared__ uint32_t s_buf[4][32];
    const uint32_t n = v.sc->wl_count[parity];
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t held = 0;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
        __syncwarp();
        const uint32_t cand = wl_in[i];
        const bool stay = sg_elect_round_item(v, cand);
        if (stay) {
            if (lane == 0) s_buf[wib][held] = cand;
            held++;
            if (held == 32) {
                __syncwarp();
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&v.sc->wl_count[parity ^ 1], 32u);
                base = __shfl_sync(0xffffffffu, base, 0);
                wl_out[base + lane] = s_buf[wib][lane];
                held = 0;
            }
        }
    }
    __syncwarp();
    if (held) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&v.sc->wl_count[parity ^ 1], held);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < held) wl_out[base + lane] = s_buf[wib][lane];
    }
}

/* The whole election in ONE cooperative launch: the rank batches and their pull rounds run inside the kernel,
 * separated by grid-wide barriers (a few microseconds) instead of kernel launches and host round trips, which
 * makes small batch growth factors - fewer wasted scans - and deep dependency chains (hundreds of rounds on
 * XOR/pigeonhole instances) cheap.  Same per-candidate body as k_elect_round. */
#ifndef SG_LB_ELECT
#define SG_LB_ELECT 1          /* minimum resident blocks per SM asked of the compiler (register budget) */
#endif
#ifndef SG_ELECT_TPB
#define SG_ELECT_TPB 1024      /* threads per block of the cooperative election: fewer, larger blocks make grid.sync cheaper */
#endif
__global__ void __launch_bounds__(SG_ELECT_TPB, SG_LB_ELECT) k_elect_persistent(SgView v, uint32_t *wl0, uint32_t *wl1, uint32_t b0, uint32_t grow)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_buf[SG_ELECT_TPB / 32][32];
    const uint32_t V = v.nvars;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    const uint32_t gwarp = gtid >> 5, nwarps = gsize >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    volatile uint32_t *cnt = v.sc->ere_pending;            /* three rotating queue counters (scratch words of SgScalars) */
    uint32_t r0 = 0, bsize = b0, rounds = 0;
    while (r0 < V) {
        const uint32_t r1 = (V - r0 <= bsize) ? V : r0 + bsize;
        if (gtid == 0) { cnt[0] = 0; cnt[1] = 0; cnt[2] = 0; }
        grid.sync();
        for (uint32_t r = r0 + gtid; r < r1; r += gsize) {
            const uint32_t var = v.order[r];
            if (SG_ER_STAT(sg_er_load(v, var)) == SG_E_UNDECIDED) wl0[atomicAdd((uint32_t *)&cnt[0], 1u)] = var;
        }
        grid.sync();
        uint32_t k = 0;                                    /* queue k%3 is read, (k+1)%3 written, (k+2)%3 cleared */
        uint32_t *bufs[2] = { wl0, wl1 };
        while (true) {
            const uint32_t n = cnt[k % 3];
            if (n == 0) break;
            const uint32_t *wl_in = bufs[k & 1];
            uint32_t *wl_out = bufs[(k + 1) & 1];
            volatile uint32_t *out_cnt = &cnt[(k + 1) % 3];
            if (gtid == 0) cnt[(k + 2) % 3] = 0;
            uint32_t held = 0;
            for (uint32_t i = gwarp; i < n; i += nwarps) {
                __syncwarp();
                const uint32_t cand = __ldcg(&wl_in[i]);
                const bool stay = sg_elect_round_item(v, cand);
                if (stay) {
                    if (lane == 0) s_buf[wib][held] = cand;
                    held++;
                    if (held == 32) {
                        __syncwarp();
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd((uint32_t *)out_cnt, 32u);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        wl_out[base + lane] = s_buf[wib][lane];
                        held = 0;
                    }
                }
            }
            __syncwarp();
            if (held) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd((uint32_t *)out_cnt, held);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (lane < held) wl_out[base + lane] = s_buf[wib][lane];
            }
            grid.sync();
            k++; rounds++;
        }
        if (*(volatile uint32_t *)&v.sc->brk_rank < r1) break;      /* the sequential loop ended inside this batch */
        r0 = r1;
        if (bsize < (1u << 28)) bsize *= grow;
    }
    if (gtid == 0) v.sc->wl_count[0] = rounds;
}

/* The election without rounds and without grid-wide barriers.  Warp w takes the candidates of rank w, w + W, w + 2W, ...
 * in that order.  A candidate that is blocked (an earlier-ranked neighbour is still undecided) is parked in one of the
 * warp's 32 slots - lane j keeps slot j in registers - together with its blocker; the warp goes on with its next
 * candidates, polls the blockers' election words (one load per lane), and looks at a parked candidate again only when its
 * blocker's word has changed.  Progress: the warps are co-resident (cooperative launch) and take their ranks in ascending
 * order, so the lowest undecided rank of the whole election is either parked - then nothing blocks it and its next look
 * decides it - or next in a warp whose parked candidates all have lower ranks, i.e. none.  The outcome is the sequential
 * one for the same reason as in the round version: a candidate is decided on final election words of all its
 * earlier-ranked neighbours.  sc->wl_count[0] receives the largest number of looks any warp needed (a report figure). */
#ifndef SG_ELECT_GW
#define SG_ELECT_GW 8          /* lanes per candidate in k_elect_async: 32 / SG_ELECT_GW candidates are looked at side by side */
#endif
#ifndef SG_ELECT_SMALL
#define SG_ELECT_SMALL (2u * SG_ELECT_GW)   /* list entries (both lists) up to which a candidate is looked at by a lane group (4x: C4 election 6.2 -> 7.6 ms) */
#endif
__global__ void __launch_bounds__(SG_ELECT_TPB, SG_LB_ELECT) k_elect_async(SgView v)
{
    constexpr int NG = 32 / SG_ELECT_GW;       /* lane groups per warp */
    const uint32_t V = v.nvars;
    const uint32_t lane = threadIdx.x & 31, grp = lane / SG_ELECT_GW;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t p_cand = 0, p_blk = 0;            /* slot `lane`: parked candidate (0 = free) and its blocker */
    uint32_t f_var = 0, f_mask = 0;            /* the fetched block: lane j holds the candidate of rank f_base + j * nwarps */
    bool p_small = false, f_small = false;     /* short lists: looked at by SG_ELECT_GW lanes, NG candidates side by side */
    uint64_t f_base = gwarp;                   /* rank of lane 0's entry of the NEXT block to fetch */
    bool more = gwarp < V;
    uint32_t looks = 0, ready = 0;
    while (true) {
        const uint32_t parked = __ballot_sync(0xffffffffu, p_cand != 0);
        if (!ready && parked) {
            /* parked candidates worth another look: the blocker has been decided, a push froze the candidate itself, or the
             * break moved below its rank (its blocker may then never be looked at) */
            bool rdy = false;
            if (p_cand) {
                const uint32_t w = sg_er_load(v, p_cand);
                rdy = SG_ER_STAT(w) != SG_E_UNDECIDED || SG_ER_STAT(sg_er_load(v, p_blk)) != SG_E_UNDECIDED ||
                      SG_ER_RANK(w) > *(volatile uint32_t *)&v.sc->brk_rank;
            }
            ready = __ballot_sync(0xffffffffu, rdy);
        }
        /* work items, chosen the same way by every lane: parked candidates that are ready first, then new candidates in rank
         * order, each with a free slot reserved in case it has to wait.  A candidate with long lists is looked at by the whole
         * warp, alone; candidates with short lists NG at a time, SG_ELECT_GW lanes each (the look waits on a chain of dependent
         * fetches: several chains in flight per warp) */
        const uint32_t psmall = __ballot_sync(0xffffffffu, p_small), fsmall = __ballot_sync(0xffffffffu, f_small);
        uint32_t freeslots = ~parked;
        int src = -1, slot = -1;               /* this lane's group: source lane (parked slot or fetch-block entry) */
        bool from_parked = false, wide = false;
        int slots[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) slots[g] = -1;
        uint32_t have = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            int s_ = -1, d_ = -1;
            bool fp = false;
            if (!wide) {
                if (ready) {
                    const int j = __ffs((int)ready) - 1;
                    const bool small = (psmall >> j) & 1u;
                    if (small || g == 0) { s_ = j; ready &= ready - 1; d_ = j; fp = true; wide = !small; }
                } else if (f_mask && freeslots) {
                    const int k = __ffs((int)f_mask) - 1;
                    const bool small = (fsmall >> k) & 1u;
                    if (small || g == 0) { s_ = k; f_mask &= f_mask - 1; d_ = __ffs((int)freeslots) - 1; freeslots &= freeslots - 1; wide = !small; }
                }
            }
            slots[g] = d_;
            if (s_ >= 0) have |= 1u << g;
            if ((uint32_t)g == grp || (wide && g == 0)) { src = s_; slot = d_; from_parked = fp; }
            if (wide) break;                                  /* the whole warp looks at this one */
        }
        if (!have) {
            if (parked != 0xFFFFFFFFu && more) {
                const uint32_t brk = __shfl_sync(0xffffffffu, *(volatile uint32_t *)&v.sc->brk_rank, 0);
                const uint64_t r = f_base + (uint64_t)lane * nwarps;
                f_var = 0; f_small = false;
                if (r < V && r <= brk) { const uint32_t x = v.order[(uint32_t)r]; if (SG_ER_STAT(sg_er_load(v, x)) == SG_E_UNDECIDED) f_var = x; }
                f_mask = __ballot_sync(0xffffffffu, f_var != 0);
                /* the two occurrence lists of a candidate are neighbours in the table: ask for their lines now (one round trip
                 * for the whole block of candidates), so that the look finds them in L2 */
                if (f_var) {
                    const uint32_t *dp = v.ot_data + v.ot_off[2 * f_var], *dn = v.ot_data + v.ot_off[2 * f_var + 1];
                    const uint32_t np = v.ot_len[2 * f_var], nn = v.ot_len[2 * f_var + 1];
                    if (np) asm volatile("prefetch.global.L2 [%0];" :: "l"(dp));
                    if (nn) asm volatile("prefetch.global.L2 [%0];" :: "l"(dn + nn - 1));
                    f_small = np + nn <= SG_ELECT_SMALL;
                }
                f_base += 32ull * nwarps;
                more = f_base < V && f_base <= brk;
                continue;
            }
            if (!parked && !f_mask && !more) break;
            __nanosleep(200);
            continue;
        }
        const int srcl = src < 0 ? 0 : src;
        const uint32_t c_p = __shfl_sync(0xffffffffu, p_cand, srcl), c_f = __shfl_sync(0xffffffffu, f_var, srcl);
        const bool s_p = __shfl_sync(0xffffffffu, (int)p_small, srcl) != 0, s_f = __shfl_sync(0xffffffffu, (int)f_small, srcl) != 0;
        const uint32_t cand = src < 0 ? 0u : (from_parked ? c_p : c_f);
        const bool csmall = from_parked ? s_p : s_f;
        uint32_t blk = 0;
        bool stay = false;
        if (wide) stay = sg_elect_eval_g<32>(v, cand, &blk);                   /* every lane: one candidate */
        else if (src >= 0) stay = sg_elect_eval_g<SG_ELECT_GW>(v, cand, &blk);  /* returns at once when a push froze it meanwhile */
        __syncwarp();
        ++looks;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const uint32_t cg_ = __shfl_sync(0xffffffffu, cand, g * SG_ELECT_GW), bg = __shfl_sync(0xffffffffu, blk, g * SG_ELECT_GW);
            const bool sg_ = __shfl_sync(0xffffffffu, (int)stay, g * SG_ELECT_GW) != 0;
            const bool sm_ = __shfl_sync(0xffffffffu, (int)csmall, g * SG_ELECT_GW) != 0;
            if (((have >> g) & 1u) && (int)lane == slots[g]) { p_cand = sg_ ? cg_ : 0u; p_blk = bg; p_small = sg_ && sm_; }
        }
    }
    if (lane == 0) atomicMax(&v.sc->wl_count[0], looks);
}

__global__ void k_acting_flags(SgView v, uint32_t *__restrict__ fa, uint32_t *__restrict__ fe)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= v.nvars) return;
    const uint32_t s = SG_ER_STAT(v.rank[v.order[r]]);
    const bool in = r < v.sc->brk_rank;
    fa[r] = in && (s == SG_E_BOTH || s == SG_E_PONLY);
    fe[r] = in && s == SG_E_BOTH;
}

__global__ void k_acting_scatter(SgView v, const uint32_t *__restrict__ fa, const uint32_t *__restrict__ pa,
                                 const uint32_t *__restrict__ fe, const uint32_t *__restrict__ pe)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= v.nvars) return;
    const uint32_t var = v.order[r];
    if (fa[r]) v.acting[pa[r]] = var | (SG_ER_STAT(v.rank[var]) == SG_E_PONLY ? 0x80000000u : 0u);
    if (fe[r]) v.elected[pe[r]] = var;
}

/* ---------------------------------------------------------------------------------------------- */
/* one item per thread                                                                              */
/* ---------------------------------------------------------------------------------------------- */
/* resident blocks per SM asked of the compiler, per kind (measured on the 10M-clause instance: these kernels wait on
 * random sector fetches, so the BVE decision, SUB and the ERE proposal gain from more warps in flight even at the price
 * of a few spilled registers; sortOT and the election lose) */
constexpr int items_min_blocks(int kind)
{
    return kind == SGK_VE_COUNT ? 8 : kind == SGK_SUB ? 8 : kind == SGK_ERE_PROPOSE ? 5 : 1;
}
template <int KIND> __global__ void __launch_bounds__(TPB, items_min_blocks(KIND)) k_items(SgView v, uint32_t n)
{
    /* SUB and the BVE decision run one warp per elected variable, everything else one thread */
    const bool warp_item = (KIND == SGK_SORTOT || KIND == SGK_SUB || KIND == SGK_VE_COUNT || KIND == SGK_NBH || KIND == SGK_ERE_PROPOSE || KIND == SGK_ERE_CLAIM ||
                            KIND == SGK_ERE_SAFE || KIND == SGK_ERE_APPLY);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = warp_item ? (t >> 5) : t;
    if (i0 >= n) return;
    /* the elected variables come in ascending score order, the heaviest neighbourhoods last: the warp-per-variable kernels
     * start with those, so that the long items do not end up alone at the tail of the launch (items are independent) */
    const bool heavy_first = (KIND == SGK_SUB || KIND == SGK_VE_COUNT || KIND == SGK_NBH || KIND == SGK_ERE_PROPOSE) && v.item_rev;
    const uint32_t i = v.item_base + (heavy_first ? n - 1u - i0 : i0);
    if (KIND == SGK_IDSORT) sg_idsort_item(v, i);
    else if (KIND == SGK_IDSORT_LONG) { if (v.ot_len[i] > 64u) sg_idsort_item(v, i); }
    else if (KIND == SGK_REDUCE) sg_reduce_item(v, i);
    else if (KIND == SGK_ELECT_PREP) sg_elect_prep_item(v, i + 1);
    else if (KIND == SGK_SORTOT) sg_sortot_warp_item(v, 2 * v.elected[i >> 1] + (i & 1));
    else if (KIND == SGK_SORTOT_LONG) sg_sortot_long_item(v, 2 * v.elected[i >> 1] + (i & 1));
    else if (KIND == SGK_SCRATCH_LEN) {
        const uint32_t var = v.elected[i];
        const uint32_t a = v.ot_len[2 * var], b = v.ot_len[2 * var + 1];
        v.scr_off[i] = (a > b ? a : b) + 2;
    }
    else if (KIND == SGK_SUB) sg_sub_item(v, i);
    else if (KIND == SGK_VE_COUNT) sg_ve_count_item(v, i);
    else if (KIND == SGK_VE_EMIT) sg_ve_emit_item(v, i);
    else if (KIND == SGK_NBH) sg_neighbourhood_item(v, i);
    else if (KIND == SGK_BCE_COUNT) sg_bce_count_item(v, i);
    else if (KIND == SGK_BCE_EMIT) sg_bce_emit_item(v, i);
    else if (KIND == SGK_ERE_CHUNKS) sg_ere_chunks_item(v, i);
    else if (KIND == SGK_ERE_PROPOSE) sg_ere_propose_item(v, i);
    else if (KIND == SGK_ERE_DIRTY) sg_ere_dirty_item(v, i);
    else if (KIND == SGK_ERE_CLAIM) sg_ere_claim_item(v, i);
    else if (KIND == SGK_ERE_SAFE) sg_ere_safe_item(v, i);
    else if (KIND == SGK_ERE_APPLY) sg_ere_apply_item(v, i);
    else if (KIND == SGK_ERE_FLAG01) v.ve_ncls[i] = v.ere_flags[i] != 0;
    else if (KIND == SGK_ERE_LIST) { if (v.ere_flags[i]) v.acting[v.ve_cls_off[i]] = i; }
}

/* sortOT for lists of 25 .. SG_SOT_WIDE_MAX entries (simplify.cpp:85-100: std::sort by CNF_CMP_KEY), one warp per list of the
 * work list the thread-per-list kernel left.  std::sort is not stable, but when no two keys of the list are equal there is
 * only one sorted order, whatever the algorithm: the keys of the list are fetched once into shared memory (all fetches in
 * flight together), every entry's final position is its rank = the number of smaller keys (each lane ranks its share of the
 * entries against all keys, broadcast reads), and equal keys are counted on the way.  A list WITH equal keys (duplicate
 * clauses) goes through the step-for-step restatement of libstdc++'s introsort on one lane, as before. */
__global__ void __launch_bounds__(128) k_sortot_wide(SgView v)
{
    __shared__ SgKey s_key[4][SG_SOT_WIDE_MAX];
    __shared__ uint32_t s_id[4][SG_SOT_WIDE_MAX];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t count = *(volatile uint32_t *)&v.sc->wl_count[1];
    for (uint32_t item = gwarp; item < count; item += nwarps) {
        const uint32_t lit = v.sot_wl[item];
        const int n = (int)v.ot_len[lit];
        uint32_t *d = sg_list(v, lit);
        __syncwarp();
        for (int i = (int)lane; i < n; i += 32) {
            const uint32_t id = d[i];
            const sg_u4 h = v.hdr[id];
            SgKey k; k.size = h.y; k.sig = h.w; k.first = v.lits[h.x]; k.last = v.lits[h.x + h.y - 1];
            s_key[w][i] = k; s_id[w][i] = id;
        }
        __syncwarp();
        bool tie = false;
        uint32_t rank[SG_SOT_WIDE_MAX / 32];
#pragma unroll
        for (int q = 0; q < SG_SOT_WIDE_MAX / 32; ++q) {
            const int i = q * 32 + (int)lane;
            uint32_t r = 0;
            if (i < n) {
                const SgKey k = s_key[w][i];
                for (int j = 0; j < n; ++j) {
                    const SgKey o = s_key[w][j];
                    if (sg_key_lt(o, k)) ++r;
                    else if (j != i && !sg_key_lt(k, o)) tie = true;
                }
            }
            rank[q] = r;
            if (q * 32 + 32 >= n) break;
        }
        if (__any_sync(0xffffffffu, tie)) {
            if (lane == 0) sg_sortot_item(v, lit);            /* equal keys: the order is std::sort's own */
            continue;
        }
#pragma unroll
        for (int q = 0; q < SG_SOT_WIDE_MAX / 32; ++q) {
            const int i = q * 32 + (int)lane;
            if (i < n) d[rank[q]] = s_id[w][i];
            if (q * 32 + 32 >= n) break;
        }
    }
}

/* ---- prop by frontiers (method in sigma_items.cuh) ---------------------------------------------------------------------- */
__global__ void k_pp_mark(SgView v, SgProp pp)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x, n = pp.p1 - pp.p0;
    if (k < n) { pp.fpos[SG_FLIP(v.trail[pp.p0 + k])] = pp.p0 + k; pp.cnt[k] = 0; }
    if (k == 0) { pp.ctl[0] = pp.ctl[1] = pp.ctl[2] = pp.ctl[3] = pp.ctl[4] = 0; }
}
/* one warp per frontier position, lanes over the list of the falsified literal in list order.
 * MODE 0: a round - discoveries into disc_new (earliest visit per literal).  MODE 1: units per position, conflicts.
 * MODE 2: units onto the trail: position base + rank inside the list. */
template <int MODE>
__global__ void k_pp_visits(SgView v, SgProp pp)
{
    const uint32_t lane = threadIdx.x & 31u, nw = (gridDim.x * blockDim.x) >> 5, n = pp.p1 - pp.p0;
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += nw) {
        const uint32_t p = pp.p0 + k, flipped = SG_FLIP(v.trail[p]);
        const uint32_t len = v.ot_len[flipped];
        const uint32_t *d = sg_list(v, flipped);
        uint32_t run = 0;
        for (uint32_t b = 0; b < len; b += 32) {
            const uint32_t i = b + lane;
            int out = SG_PP_SKIP;
            uint32_t unit = 0;
            if (i < len) out = sg_pp_visit(v, pp, p, i, flipped, d[i], &unit);
            if (MODE == 0) {
                if (out == SG_PP_UNIT) {
                    const unsigned long long t = ((unsigned long long)p << 32) | i;
                    if (atomicMin(&pp.disc_new[unit], t) == SG_PP_NEVER) pp.touched_new[atomicAdd(&pp.ctl[1], 1u)] = unit;
                }
            } else {
                if (MODE == 1 && out == SG_PP_CONFLICT) pp.ctl[3] = 1;
                const uint32_t mk = __ballot_sync(0xffffffffu, out == SG_PP_UNIT);
                if (MODE == 2 && out == SG_PP_UNIT) v.trail[pp.p1 + pp.base[k] + run + __popc(mk & ((1u << lane) - 1u))] = unit;
                run += __popc(mk);
            }
        }
        if (MODE == 1 && lane == 0) pp.cnt[k] = run;
    }
}
__global__ void k_pp_diff(SgProp pp)
{
    const uint32_t n0 = pp.ctl[0], n1 = pp.ctl[1];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n0 + n1; k += gridDim.x * blockDim.x) {
        const uint32_t lit = k < n0 ? pp.touched[k] : pp.touched_new[k - n0];
        if (pp.disc[lit] != pp.disc_new[lit]) pp.ctl[2] = 1;
    }
}
/* forget the entries of `arr` named by list[0, *n) */
__global__ void k_pp_forget(unsigned long long *arr, const uint32_t *list, const uint32_t *n)
{
    const uint32_t cnt = *n;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) arr[list[k]] = SG_PP_NEVER;
}
__global__ void k_pp_shift(SgProp pp) { pp.ctl[0] = pp.ctl[1]; pp.ctl[1] = 0; pp.ctl[2] = 0; }
/* every clause of the frontier's lists once: the first visitor to set its bit rewrites it */
__global__ void k_pp_rewrite(SgView v, SgProp pp)
{
    const uint32_t lane = threadIdx.x & 31u, nw = (gridDim.x * blockDim.x) >> 5, n = pp.p1 - pp.p0;
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += nw) {
        const uint32_t flipped = SG_FLIP(v.trail[pp.p0 + k]);
        const uint32_t len = v.ot_len[flipped];
        const uint32_t *d = sg_list(v, flipped);
        for (uint32_t i = lane; i < len; i += 32) {
            const uint32_t c = d[i], bit = 1u << (c & 31u);
            if (atomicOr(&pp.claim[c >> 5], bit) & bit) continue;
            sg_pp_rewrite(v, pp, c);
        }
    }
}
/* fot.clear, toblivion(ot[assign]) for the frontier; enqueueUnit for the new units; scalars */
__global__ void k_pp_finish(SgView v, SgProp pp, uint32_t n_new, int conflict)
{
    const uint32_t lane = threadIdx.x & 31u, nw = (gridDim.x * blockDim.x) >> 5, n = pp.p1 - pp.p0;
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += nw) {
        const uint32_t assign = v.trail[pp.p0 + k], flipped = SG_FLIP(assign);
        if (!conflict) {
            const uint32_t len = v.ot_len[assign];
            const uint32_t *d = sg_list(v, assign);
            for (uint32_t i = lane; i < len; i += 32) sg_mark_deleted(v, d[i]);
            __syncwarp();
            if (lane == 0) { v.ot_len[assign] = 0; v.ot_len[flipped] = 0; }
        }
        if (lane == 0) pp.fpos[flipped] = SG_PP_NONE;
    }
    if (conflict) { if (blockIdx.x == 0 && threadIdx.x == 0) v.sc->unsat = 1; return; }
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_new; j += gridDim.x * blockDim.x) {
        const uint32_t u = v.trail[pp.p1 + j];
        v.state[SG_ABS(u)] = (uint8_t)SG_FROZEN_M;
        v.value[u] = 1;
        v.value[SG_FLIP(u)] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        SgScalars *sc = v.sc;
        sc->propagated = pp.p1;
        sc->n_trail = pp.p1 + n_new;
        sc->max_frozen += n_new;
        sc->unassigned -= n_new;
    }
}

/* ---- units of one SUB / VE stage, sorted by (elected index, sequence), applied at once ------------------------------------- */
/* The one-thread walk (sg_apply_units_serial) enqueues a unit when its literal is unassigned, skips it when the literal is
 * true and stops with the empty clause when it is false.  With first[lit] = the earliest position of lit in the sorted list:
 * position k enqueues iff lit was unassigned before the stage and first[lit] == k and -lit does not come earlier; a literal
 * that was false before the stage, or whose negation comes earlier, is the conflict (any conflict ends the pass, whichever
 * position finds it).  Trail order = list order: a prefix sum over the enqueue flags. */
__global__ void k_au_first(SgView v, uint32_t n, uint32_t *first)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) atomicMin(&first[v.ubuf[k].lit], k);
}
__global__ void k_au_flags(SgView v, uint32_t n, const uint32_t *first, uint32_t *flags)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t lit = v.ubuf[k].lit;
    const int8_t val = v.value[lit];
    uint32_t f = 0;
    if (val & (int8_t)-2) {
        if (first[lit] == k) { if (first[SG_FLIP(lit)] < k) v.sc->unsat = 1; else f = 1; }
    } else if (!val) v.sc->unsat = 1;
    flags[k] = f;
}
__global__ void k_au_write(SgView v, uint32_t n, uint32_t *first, const uint32_t *flags, const uint32_t *pos, const uint32_t *total)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    SgScalars *sc = v.sc;
    const bool unsat = sc->unsat != 0;                     /* written by the launch before this one */
    if (k < n) {
        const uint32_t lit = v.ubuf[k].lit;
        if (flags[k] && !unsat) {
            v.trail[sc->n_trail + pos[k]] = lit;
            v.state[SG_ABS(lit)] = (uint8_t)SG_FROZEN_M;
            v.value[lit] = 1;
            v.value[SG_FLIP(lit)] = 0;
        }
    }
}
__global__ void k_au_finish(SgView v, uint32_t n, uint32_t *first, const uint32_t *total)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) first[v.ubuf[k].lit] = 0xFFFFFFFFu;
    if (k == 0 && !v.sc->unsat) {
        SgScalars *sc = v.sc;
        const uint32_t t = *total;
        sc->n_trail += t; sc->max_frozen += t; sc->unassigned -= t; sc->units_forced += t;
    }
}

__global__ void k_serial(SgView v, int kind, uint32_t arg, uint32_t arg2)
{   /* launched with one warp (only the ERE replay uses more than lane 0); the replay gets helper warps that prefetch */
    if (kind == SGS_ERE_REPLAY) {
        const uint32_t warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if (warp == 0) sg_ere_replay_serial(v, arg, arg2);
        else sg_ere_prefetch(v, arg, arg2, warp - 1, nw - 1);
        return;
    }
    if (threadIdx.x) return;
    if (kind == SGS_PROP) sg_prop_serial(v, arg);
    else if (kind == SGS_MARKS) sg_marks_serial(v, arg);
    else if (kind == SGS_APPLY_UNITS) sg_apply_units_serial(v, arg);
    else if (kind == SGS_ERE_SPLIT) sg_ere_split_serial(v, arg, arg2);
}

/* The ERE replay (redundancy.cpp:26-123 in elected order) by ONE block of 32 warps instead of one warp.
 * Units of work in sequential order: the recorded pairs of clean variables, and the dirty variables (processed in full).
 * A recorded pair only ever writes clauses the proposal pass marked `touched`, and the owner of a touched clause is dirty
 * from the start; so between two consecutive dirty variables the pairs can be applied by the claim / safe / apply rounds
 * of the prefix version (every pending pair claims the clauses it hits now - a superset of what it can hit later -, a pair
 * that holds all its claims is the earliest pending writer of each of them and sees exactly the sequential state), here
 * inside the block: warps over pairs, block barriers between the steps, the claimed clause ids kept per pair so that the
 * check and the reset of the claim words need no second enumeration.  A dirty variable is a barrier: warp 0 processes it
 * like the reference (it may turn LATER variables dirty, which the search for the next barrier then sees).  No host round
 * trips, no per-round launches; the one-warp replay remains as fall-back (claim buffer missing, > 64 variables queued). */
constexpr uint32_t ERE_CLAIMS = 8;
__global__ void __launch_bounds__(1024, 1) k_ere_replay_block(SgView v, uint32_t n_events, uint32_t n_list, uint32_t *__restrict__ claims,
                                                             uint32_t *__restrict__ ccnt)
{
    __shared__ uint32_t s_ep, s_cur, s_li, s_nd, s_P, s_left, s_stop, s_min;
    const SgLane L = sg_lane_ctx();
    const uint32_t tid = threadIdx.x, warp = tid >> 5, nw = blockDim.x >> 5;
    uint32_t merged[SG_ERE_MAXLEN];
    SgScalars *sc = v.sc;
    if (tid == 0) { s_ep = 0; s_cur = 0; s_li = 0; s_stop = 0; sc->ere_npending = 0; sc->ere_overflow = 0; }
    __syncthreads();
    while (true) {
        /* the next dirty variable at or after s_cur: on the list (flags can have changed), or queued by sg_ere_mark_owner */
        if (tid == 0) {
            uint32_t li = s_li;
            while (li < n_list && v.acting[li] < s_cur) ++li;
            s_li = li; s_min = 0xFFFFFFFFu;
        }
        __syncthreads();
        for (uint32_t base = s_li; base < n_list; base += blockDim.x) {
            const uint32_t q = base + tid;
            if (q < n_list) { const uint32_t e = v.acting[q]; if (*(volatile uint8_t *)&v.ere_flags[e] & 2u) atomicMin(&s_min, e); }
            __syncthreads();
            const uint32_t found = s_min;                  /* every thread reads it before anybody moves on to the next chunk */
            __syncthreads();
            if (found != 0xFFFFFFFFu) break;
        }
        if (tid == 0) {
            uint32_t nd = s_min;
            const uint32_t npend = *(volatile uint32_t *)&sc->ere_npending;
            if (npend && *(volatile uint32_t *)&sc->ere_pending[0] < nd) nd = *(volatile uint32_t *)&sc->ere_pending[0];
            s_nd = nd;
            uint32_t lo = s_ep, hi = n_events;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (v.ubuf[mid].eidx < nd) lo = mid + 1; else hi = mid; }
            s_P = lo;
        }
        __syncthreads();
        const uint32_t ep = s_ep, P = s_P, nd = s_nd;
        /* recorded pairs [ep, P) of the clean variables in front of it */
        while (ep < P) {
            if (tid == 0) s_left = 0;
            for (uint32_t k = ep + warp; k < P; k += nw) {                       /* claim */
                if (v.ere_done[k]) continue;
                if (L.lane == 0) ccnt[k] = 0;
                __syncwarp();
                sg_ere_event_hits(v, L, k, merged, [&](uint32_t c) {
                    atomicMin(&v.ere_owner[c], k);
                    const uint32_t i = atomicAdd(&ccnt[k], 1u);
                    if (i < ERE_CLAIMS) claims[(size_t)k * ERE_CLAIMS + i] = c;
                });
            }
            __syncthreads();
            for (uint32_t k = ep + warp; k < P; k += nw) {                       /* holds every claim? */
                if (v.ere_done[k]) continue;
                const uint32_t n = ccnt[k];
                bool mine = true;
                if (n <= ERE_CLAIMS) { if (L.lane < n) mine = v.ere_owner[claims[(size_t)k * ERE_CLAIMS + L.lane]] == k; }
                else sg_ere_event_hits(v, L, k, merged, [&](uint32_t c) { if (v.ere_owner[c] != k) mine = false; });
                const bool all = __ballot_sync(0xffffffffu, !mine) == 0;
                __syncwarp();
                if (L.lane == 0) v.ere_done[k] = all ? 2u : 0u;
            }
            __syncthreads();
            for (uint32_t k = ep + warp; k < P; k += nw) {                       /* claim words back to "nobody" */
                if (v.ere_done[k] == 1u) continue;
                const uint32_t n = ccnt[k];
                if (n <= ERE_CLAIMS) { if (L.lane < n) v.ere_owner[claims[(size_t)k * ERE_CLAIMS + L.lane]] = 0xFFFFFFFFu; }
                else sg_ere_event_hits(v, L, k, merged, [&](uint32_t c) { v.ere_owner[c] = 0xFFFFFFFFu; });
            }
            __syncthreads();
            for (uint32_t k = ep + warp; k < P; k += nw) {                       /* apply; the applied pairs write disjoint clauses */
                const uint32_t d = v.ere_done[k];
                if (d == 2u) {
                    const uint32_t eidx = v.ubuf[k].eidx, key = v.ubuf[k].seq;
                    const uint32_t x = v.elected[eidx];
                    uint32_t dx, fx;
                    if (sg_ere_lists(v, x, dx, fx)) sg_ere_pair(v, L, x, eidx, sg_list(v, dx)[key >> 16], sg_list(v, fx)[key & 0xffffu], merged);
                    __syncwarp();
                    if (L.lane == 0) { v.ere_done[k] = 1u; atomicAdd(&sc->ere_applied, 1u); atomicAdd((unsigned long long *)&sc->ere_event_pairs, 1ull); }
                } else if (d == 0u && L.lane == 0) atomicAdd(&s_left, 1u);
            }
            __syncthreads();
            const uint32_t left = s_left;                  /* read by everybody before thread 0 clears it for the next round */
            __syncthreads();
            if (left == 0) break;
        }
        if (nd == 0xFFFFFFFFu) break;
        /* the dirty variable in full (redundancy.cpp:59-117 for every pair of its two lists, in order).  Nearly all pairs change
         * nothing, and a pair that changes nothing can be skipped: the 32 warps test the pairs from q0 on side by side,
         * read-only, for the first one that acts; warp 0 applies that pair; the search resumes behind it on the new state.
         * (One warp walking all pairs - two dependent fetch chains per pair - was half of the replay's time on the 5M-gate
         * circuit: 2 904 pairs of 66 variables, 4 ms.) */
        {
            if (tid == 0) {
                if (sc->ere_npending && sc->ere_pending[0] == nd) {                  /* pop the head of the queue */
                    for (uint32_t k = 1; k < sc->ere_npending; ++k) sc->ere_pending[k - 1] = sc->ere_pending[k];
                    sc->ere_npending--;
                }
                sc->ere_dirty_vars++;
            }
            const uint32_t x = v.elected[nd];
            uint32_t dx, fx;
            if (sg_ere_lists(v, x, dx, fx)) {
                const uint32_t ds = v.ot_len[dx], fs = v.ot_len[fx], total = ds * fs;
                const uint32_t *dm = sg_list(v, dx), *dt = sg_list(v, fx);
                if (tid == 0) sc->ere_full_pairs += (unsigned long long)ds * fs;
                uint32_t q0 = 0;
                while (q0 < total) {
                    if (tid == 0) s_min = 0xFFFFFFFFu;
                    __syncthreads();
                    for (uint32_t q = q0 + warp; q < total; q += nw) {
                        if (q > *(volatile uint32_t *)&s_min) break;
                        const uint32_t i = q / fs, j = q - i * fs;
                        if (sg_ere_pair_acts(v, L, x, dm[i], dt[j], merged)) { if (L.lane == 0) atomicMin(&s_min, q); break; }
                    }
                    __syncthreads();
                    const uint32_t qm = s_min;
                    __syncthreads();
                    if (qm == 0xFFFFFFFFu) break;
                    if (warp == 0) sg_ere_pair(v, L, x, nd, dm[qm / fs], dt[qm % fs], merged);
                    __syncthreads();
                    q0 = qm + 1;
                }
            }
            if (tid == 0) {
                uint32_t q = P;
                while (q < n_events && v.ubuf[q].eidx <= nd) ++q;                    /* the recorded pairs of a dirty variable are not used */
                s_ep = q; s_cur = nd + 1;
                if (*(volatile uint32_t *)&sc->ere_overflow) s_stop = 1;
            }
        }
        __syncthreads();
        if (s_stop) break;
    }
    /* more than 64 variables were queued at once: finish like the one-warp replay, scanning the flags */
    if (s_stop && warp == 0) {
        uint32_t ep = s_ep;
        for (uint32_t e = s_cur; e < v.n_elected; ++e) {
            __syncwarp();
            const uint32_t fl = sg_bcast(*(volatile uint8_t *)&v.ere_flags[e]);
            if (fl) sg_ere_replay_var(v, L, e, fl, ep, n_events, merged);
        }
    }
}

__global__ void k_sort_units_small(SgView v, uint32_t n)
{   /* one thread: insertion sort by (eidx, seq); n is tiny in practice */
    for (uint32_t i = 1; i < n; ++i) {
        const SgUnit t = v.ubuf[i];
        uint32_t j = i;
        while (j > 0 && (v.ubuf[j - 1].eidx > t.eidx || (v.ubuf[j - 1].eidx == t.eidx && v.ubuf[j - 1].seq > t.seq))) {
            v.ubuf[j] = v.ubuf[j - 1];
            --j;
        }
        v.ubuf[j] = t;
    }
}
/* up to 2048 units: bitonic sort of (eidx, seq) keys in shared memory by one block */
__global__ void k_sort_units_block(SgView v, uint32_t n)
{
    __shared__ unsigned long long key[2048];
    __shared__ uint32_t lit[2048];
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) {
        if (i < n) { const SgUnit u = v.ubuf[i]; key[i] = ((unsigned long long)u.eidx << 32) | u.seq; lit[i] = u.lit; }
        else { key[i] = ~0ull; lit[i] = 0; }
    }
    __syncthreads();
    for (uint32_t k = 2; k <= 2048; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) {
                const uint32_t p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = key[i], b = key[p];
                    if ((a > b) == up) { key[i] = b; key[p] = a; const uint32_t t = lit[i]; lit[i] = lit[p]; lit[p] = t; }
                }
            }
            __syncthreads();
        }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        SgUnit u; u.eidx = (uint32_t)(key[i] >> 32); u.seq = (uint32_t)key[i]; u.lit = lit[i]; u.pad = 0;
        v.ubuf[i] = u;
    }
}
__global__ void k_units_keys(SgView v, uint32_t n, uint32_t *keys, uint32_t *idx, int which)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = which ? idx[i] : i;
    keys[i] = which ? v.ubuf[src].eidx : v.ubuf[src].seq;
    if (!which) idx[i] = i;
}
__global__ void k_units_gather(SgView v, uint32_t n, const uint32_t *idx, SgUnit *tmp)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tmp[i] = v.ubuf[idx[i]];
}

/* ---------------------------------------------------------------------------------------------- */
/* counts, compaction, export                                                                       */
/* ---------------------------------------------------------------------------------------------- */
__global__ void k_zero_live(SgView v) { v.sc->live_cls = 0; v.sc->live_lits = 0; }

__global__ void k_count_live(SgView v)
{   /* persistent grid (a few blocks per SM), four independent 16-byte header loads in flight per thread,
     * one pair of atomics per block */
    uint32_t nc = 0;
    unsigned long long nl = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    for (; c + 3 * stride < v.n_cls; c += 4 * stride) {
        const uint4 h0 = v.hdr[c], h1 = v.hdr[c + stride], h2 = v.hdr[c + 2 * stride], h3 = v.hdr[c + 3 * stride];
        if (!(h0.z & SG_DELETED)) { nc++; nl += h0.y; }
        if (!(h1.z & SG_DELETED)) { nc++; nl += h1.y; }
        if (!(h2.z & SG_DELETED)) { nc++; nl += h2.y; }
        if (!(h3.z & SG_DELETED)) { nc++; nl += h3.y; }
    }
    for (; c < v.n_cls; c += stride) {
        const uint4 h = v.hdr[c];
        if (!(h.z & SG_DELETED)) { nc++; nl += h.y; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        nc += __shfl_down_sync(0xffffffffu, nc, d);
        nl += __shfl_down_sync(0xffffffffu, nl, d);
    }
    __shared__ uint32_t sc_[TPB / 32];
    __shared__ unsigned long long sl_[TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sc_[warp] = nc; sl_[warp] = nl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a2 = 0; unsigned long long b2 = 0;
        for (int w = 0; w < TPB / 32; ++w) { a2 += sc_[w]; b2 += sl_[w]; }
        if (a2) { atomicAdd(&v.sc->live_cls, a2); atomicAdd(&v.sc->live_lits, b2); }
    }
}

__global__ void k_zero_u32(uint32_t *p) { *p = 0; }
/* control words between device memory and MAPPED pinned host memory, by a kernel instead of the copy engine: the
 * DMA queues then carry only the bulk formula transfers (a second context's 400 MB copy would otherwise sit in
 * front of every little read-back of this one) */
__global__ void k_copy_words(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, uint32_t n)
{
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__global__ void k_count_melted(SgView v, uint32_t *out)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const bool m = x <= v.nvars && (v.state[x] & SG_MELTED_M);
    const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, m));
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}

/* compaction = exclusive scan of (live, size) over the clause ids + move, without materialising flags:
 * tile sums straight from the headers, a one-block scan of the tile sums, then a down-sweep that
 * re-derives each clause's (new id, new offset) in registers and moves the clause at once */
constexpr int CP_ITEMS = 4;
constexpr int CP_TILE = TPB * CP_ITEMS;

__device__ __forceinline__ uint2 cp_block_scan(uint32_t a, uint32_t b, uint2 *total)
{   /* exclusive scan of a pair per thread over the block */
    __shared__ uint2 ws[TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ya = __shfl_up_sync(0xffffffffu, ia, d), yb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ia += ya; ib += yb; }
    }
    if (lane == 31) ws[warp] = make_uint2(ia, ib);
    __syncthreads();
    if (warp == 0) {
        uint2 w = lane < TPB / 32 ? ws[lane] : make_uint2(0, 0);
#pragma unroll
        for (int d = 1; d < TPB / 32; d <<= 1) {
            const uint32_t ya = __shfl_up_sync(0xffffffffu, w.x, d), yb = __shfl_up_sync(0xffffffffu, w.y, d);
            if (lane >= d) { w.x += ya; w.y += yb; }
        }
        if (lane < TPB / 32) ws[lane] = w;
    }
    __syncthreads();
    const uint2 base = warp ? ws[warp - 1] : make_uint2(0, 0);
    if (total) *total = ws[TPB / 32 - 1];
    __syncthreads();
    return make_uint2(base.x + ia - a, base.y + ib - b);
}

__global__ void k_live_reduce(SgView v, uint2 *__restrict__ part)
{
    const uint64_t base = (uint64_t)blockIdx.x * CP_TILE;
    uint32_t cnt = 0, lits = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const uint64_t c = base + (uint64_t)k * TPB + threadIdx.x;
        if (c < v.n_cls) {
            const uint4 h = v.hdr[c];
            if (!(h.z & SG_DELETED)) { cnt++; lits += h.y; }
        }
    }
    __shared__ uint2 tot;
    (void)cp_block_scan(cnt, lits, &tot);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

__global__ void k_live_partials(uint2 *__restrict__ part, uint32_t nb, uint32_t *__restrict__ totals)
{
    __shared__ uint2 carry, tot;
    if (threadIdx.x == 0) carry = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += TPB) {
        const uint32_t i = base + threadIdx.x;
        const uint2 x = i < nb ? part[i] : make_uint2(0, 0);
        const uint2 e = cp_block_scan(x.x, x.y, &tot);
        if (i < nb) part[i] = make_uint2(carry.x + e.x, carry.y + e.y);
        __syncthreads();
        if (threadIdx.x == 0) { carry.x += tot.x; carry.y += tot.y; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry.x; totals[1] = carry.y; }
}

template <bool EXPORT>
__global__ void k_live_apply(SgView v, const uint2 *__restrict__ part, uint4 *__restrict__ hdr2, uint32_t *__restrict__ lits2,
                             uint32_t *__restrict__ arena, unsigned long long *__restrict__ refs)
{
    /* four sub-tiles of 256 consecutive clauses, one clause per thread: neighbouring lanes copy neighbouring
     * pool ranges, which the L1/L2 sectors coalesce (ncu: 66 % of DRAM peak for the plain per-clause copy) */
    const uint2 b = part[blockIdx.x];                       /* (clauses, literals) before this tile */
    uint32_t id0 = b.x, off0 = b.y;
    __shared__ uint2 tot;
    /* all headers first (independent loads), then the scans, then the copies */
    uint4 hs[CP_ITEMS];
    uint2 es[CP_ITEMS];
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const uint64_t c = (uint64_t)blockIdx.x * CP_TILE + (uint64_t)k * TPB + threadIdx.x;
        hs[k] = c < v.n_cls ? v.hdr[c] : make_uint4(0, 0, SG_DELETED, 0);
    }
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const bool live = !(hs[k].z & SG_DELETED);
        const uint2 e = cp_block_scan(live ? 1u : 0u, live ? hs[k].y : 0u, &tot);
        es[k] = make_uint2(id0 + e.x, off0 + e.y);
        id0 += tot.x; off0 += tot.y;
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const uint4 h = hs[k];
        const bool live = !(h.z & SG_DELETED);
        if (live) {
            const uint32_t id = es[k].x, off = es[k].y;
            const uint32_t *src = v.lits + h.x;
            if (EXPORT) {
                const unsigned long long r = 3ull * id + off;
                refs[id] = r;
                arena[r] = h.z; arena[r + 1] = h.w; arena[r + 2] = h.y;
                uint32_t *dst = arena + r + 3;
                for (uint32_t q = 0; q < h.y; ++q) dst[q] = src[q];
            } else {
                hdr2[id] = make_uint4(off, h.y, h.z, h.w);
                uint32_t *dst = lits2 + off;
                for (uint32_t q = 0; q < h.y; ++q) dst[q] = src[q];
            }
        }
    }
}

/* final shrinkSimp + export: like k_live_apply<true>, but every 256-clause sub-tile is assembled in shared memory (its
 * output is one contiguous run of the arena) and leaves as one bulk store instead of eight 4-byte stores per thread
 * at a 32-byte stride.  Sub-tiles whose output exceeds the staging buffer (long clauses) write directly. */
__global__ void __launch_bounds__(TPB) k_export_staged(SgView v, const uint2 *__restrict__ part, uint32_t *__restrict__ arena,
                                                       unsigned long long *__restrict__ refs)
{
    __shared__ __align__(128) uint32_t stage[EX_CAP + 4];
    __shared__ uint2 tot;
    const uint2 b = part[blockIdx.x];
    uint32_t id0 = b.x, off0 = b.y;
    uint4 hs[CP_ITEMS];
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const uint64_t c = (uint64_t)blockIdx.x * CP_TILE + (uint64_t)k * TPB + threadIdx.x;
        hs[k] = c < v.n_cls ? v.hdr[c] : make_uint4(0, 0, SG_DELETED, 0);
    }
#pragma unroll
    for (int k = 0; k < CP_ITEMS; ++k) {
        const uint4 h = hs[k];
        const bool live = !(h.z & SG_DELETED);
        const uint2 e = cp_block_scan(live ? 1u : 0u, live ? h.y : 0u, &tot);
        const unsigned long long g0 = 3ull * id0 + off0;             /* first arena word of this sub-tile */
        const uint32_t W = 3u * tot.x + tot.y;
        const bool staged = W <= EX_CAP;
        if (live) {
            const uint32_t id = id0 + e.x;
            const unsigned long long r = 3ull * id + off0 + e.y;
            refs[id] = r;
            uint32_t *d = staged ? stage + (uint32_t)(g0 & 3ull) + (uint32_t)(r - g0) : arena + r;
            d[0] = h.z; d[1] = h.w; d[2] = h.y;
            const uint32_t *src = v.lits + h.x;
            for (uint32_t q = 0; q < h.y; ++q) d[3 + q] = src[q];
        }
        if (staged) {
            stage_fence();
            __syncthreads();
            if (W) flush_staged(arena, g0, stage, W);
        }
        id0 += tot.x; off0 += tot.y;
        __syncthreads();
    }
}

/* createOT order fix, staged: right after the build the lists are packed in literal order, so the lists of 256 consecutive
 * literals are one contiguous run of ot_data.  The run is loaded into shared memory with coalesced loads, every thread
 * sorts its own list there (ascending clause id = the reference's push order), and the run goes back with one bulk store.
 * Runs larger than the staging buffer fall back to sorting in global memory. */
__global__ void __launch_bounds__(TPB) k_ot_order_staged(SgView v)
{
    __shared__ __align__(128) uint32_t stage[EX_CAP + 4];
    const uint32_t l0 = blockIdx.x * blockDim.x, lit = l0 + threadIdx.x;
    const uint32_t lend = l0 + blockDim.x < v.nlit ? l0 + blockDim.x : v.nlit;
    const uint32_t g0 = v.ot_off[l0], g1 = lend < v.nlit ? v.ot_off[lend] : v.ot_off[v.nlit - 1] + v.ot_len[v.nlit - 1];
    const uint32_t W = g1 - g0, shift = g0 & 3u;
    /* lists of more than 64 entries are left to k_items<SGK_IDSORT_LONG> (which runs with the whole L1: this kernel's
     * staging buffers take most of it, and a long list sorted here would keep 255 threads waiting at the barrier) */
    const bool mine = lit < v.nlit && v.ot_len[lit] <= 64u;
    if (W > EX_CAP) { if (mine) sg_idsort_item(v, lit); return; }
    for (uint32_t k = threadIdx.x; k < W; k += blockDim.x) stage[shift + k] = v.ot_data[g0 + k];
    __syncthreads();
    bool changed = false;
    if (mine) {
        const int n = (int)v.ot_len[lit];
        uint32_t *d = stage + shift + (v.ot_off[lit] - g0);
        for (int i = 1; i < n; ++i) {
            const uint32_t t = d[i];
            int j = i;
            while (j > 0 && d[j - 1] > t) { d[j] = d[j - 1]; --j; }
            if (j != i) { d[j] = t; changed = true; }
        }
    }
    stage_fence();
    if (!__syncthreads_or(changed)) return;                             /* already in order: nothing to write */
    flush_staged(v.ot_data, g0, stage, W);
}

/* ---------------------------------------------------------------------------------------------- */
/* incremental occurrence-table update (see sigma_backend.h): only the lists that changed are touched  */
/* ---------------------------------------------------------------------------------------------- */
struct OtuTmp { uint32_t *addc, *cursor, *flag; };
__host__ __device__ inline OtuTmp otu_layout(uint32_t *tmp, uint32_t nlit)
{
    OtuTmp t;
    t.addc = tmp; t.cursor = t.addc + nlit; t.flag = t.cursor + nlit;
    return t;
}
enum { OTU_DIRTY = 1u, OTU_GROWN = 2u };

/* bitmap of the clauses represented in the table (= live when it was last built or updated) */
__global__ void k_otu_bitmap(SgView v, uint32_t *__restrict__ bits)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < v.n_cls && !(v.hdr[c].z & SG_DELETED);
    const uint32_t m = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0) bits[c >> 5] = m;
}

/* clauses that died since: their lists are dirty; clauses that are new: their literals are additions */
__global__ void k_otu_mark(SgView v, uint32_t *__restrict__ bits, uint32_t first_new, uint32_t *__restrict__ addc, uint32_t *__restrict__ flag)
{   /* one warp per 128 consecutive clause ids: four independent header loads per lane, four bitmap words per warp */
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 128u;
    uint4 h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = base + k * 32u + lane;
        h[k] = c < v.n_cls ? v.hdr[c] : make_uint4(0, 0, SG_DELETED, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = base + k * 32u + lane;
        const bool live = c < v.n_cls && !(h[k].z & SG_DELETED);
        if (c < v.n_cls) {
            const uint32_t *l = v.lits + h[k].x;
            if (c >= first_new) {
                if (live) for (uint32_t q = 0; q < h[k].y; ++q) { atomicAdd(&addc[l[q]], 1u); flag[l[q]] = OTU_DIRTY; }
            } else if (!live && ((bits[c >> 5] >> (c & 31)) & 1u)) {
                for (uint32_t q = 0; q < h[k].y; ++q) flag[l[q]] = OTU_DIRTY;
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, live);
        if (lane == 0 && (c >> 5) <= v.n_cls / 32) bits[c >> 5] = m;
    }
}
__global__ void k_otu_pairs_count(SgView v, uint32_t n, uint32_t *__restrict__ addc, uint32_t *__restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t lit = v.ot_add[2 * i], c = v.ot_add[2 * i + 1];
    if (v.hdr[c].z & SG_DELETED) return;
    atomicAdd(&addc[lit], 1u);
    flag[lit] = OTU_DIRTY;
}
/* the lists of last phase's elected variables are in sortOT order (or belong to an eliminated variable) */
__global__ void k_otu_elected(SgView v, uint32_t n, uint32_t *__restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = v.elected[i];
    flag[2 * x] = OTU_DIRTY; flag[2 * x + 1] = OTU_DIRTY;
}
/* words the grown lists need at the end of ot_data */
__global__ void k_otu_need(SgView v, const uint32_t *__restrict__ addc, uint32_t *need)
{
    const uint32_t lit = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w = 0;
    if (lit < v.nlit && addc[lit] && !(v.state[lit >> 1] & SG_MELTED_M)) w = v.ot_len[lit] + addc[lit];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w += __shfl_down_sync(0xffffffffu, w, d);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(need, w);
}

__global__ void k_otu_apply(SgView v, const uint32_t *__restrict__ bits, const uint32_t *__restrict__ addc,
                            uint32_t *__restrict__ cursor, uint32_t *__restrict__ flag, uint32_t *bump)
{
    const uint32_t lit = blockIdx.x * blockDim.x + threadIdx.x;
    if (lit >= v.nlit) return;
    if (v.state[lit >> 1] & SG_MELTED_M) { v.ot_len[lit] = 0; flag[lit] = 0; return; }   /* eliminated: no live clause holds it */
    const uint32_t f = flag[lit];
    if (!f) return;                                       /* untouched list */
    const uint32_t len = v.ot_len[lit], a = addc[lit];
    const uint32_t *src = v.ot_data + v.ot_off[lit];
    uint32_t *dst = v.ot_data + v.ot_off[lit];
    if (a) {                                              /* grows: move to fresh room at the end of the table */
        const uint32_t base = atomicAdd(bump, len + a);
        dst = v.ot_data + base;
        v.ot_off[lit] = base;
    }
    uint32_t n = 0, prev = 0, unsorted = 0;
    for (uint32_t i0 = 0; i0 < len; i0 += 8) {
        /* liveness from the bitmap the plan just refreshed (1 bit per clause, cache resident); 8 entries per step */
        uint32_t c[8], w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = i0 + k < len ? src[i0 + k] : 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = c[k] != 0xFFFFFFFFu ? bits[c[k] >> 5] : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (c[k] == 0xFFFFFFFFu || !((w[k] >> (c[k] & 31)) & 1u)) continue;
            if (n && c[k] < prev) unsorted = 1;
            prev = c[k];
            dst[n++] = c[k];
        }
    }
    if (unsorted && n > 256) sg_heapsort_u32(dst, (int)n);   /* a long list in sortOT order: no quadratic walk */
    else if (unsorted) {                                  /* sortOT order -> clause-id order */
        for (uint32_t i = 1; i < n; ++i) {
            const uint32_t t = dst[i];
            uint32_t j = i;
            while (j > 0 && dst[j - 1] > t) { dst[j] = dst[j - 1]; --j; }
            dst[j] = t;
        }
    }
    v.ot_len[lit] = n + a;
    cursor[lit] = v.ot_off[lit] + n;
    flag[lit] = a ? OTU_GROWN : 0u;
}

__global__ void k_otu_append_clauses(SgView v, uint32_t first_new, uint32_t *__restrict__ cursor)
{
    const uint32_t c = first_new + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.n_cls) return;
    const uint4 h = v.hdr[c];
    if (h.z & SG_DELETED) return;
    const uint32_t *l = v.lits + h.x;
    for (uint32_t k = 0; k < h.y; ++k) v.ot_data[atomicAdd(&cursor[l[k]], 1u)] = c;
}
__global__ void k_otu_append_pairs(SgView v, uint32_t n, uint32_t *__restrict__ cursor)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t lit = v.ot_add[2 * i], c = v.ot_add[2 * i + 1];
    if (v.hdr[c].z & SG_DELETED) return;
    v.ot_data[atomicAdd(&cursor[lit], 1u)] = c;
}
__global__ void k_otu_sort(SgView v, const uint32_t *__restrict__ flag)
{
    const uint32_t lit = blockIdx.x * blockDim.x + threadIdx.x;
    if (lit >= v.nlit || flag[lit] != OTU_GROWN) return;
    uint32_t *d = v.ot_data + v.ot_off[lit];
    const int n = (int)v.ot_len[lit];
    for (int i = 1; i < n; ++i) {
        const uint32_t t = d[i];
        int j = i;
        while (j > 0 && d[j - 1] > t) { d[j] = d[j - 1]; --j; }
        d[j] = t;
    }
}

__global__ void k_ot_compare(SgView v, const uint32_t *__restrict__ off2, const uint32_t *__restrict__ len2,
                             const uint32_t *__restrict__ data2, uint32_t *mismatch)
{
    const uint32_t lit = blockIdx.x * blockDim.x + threadIdx.x;
    if (lit >= v.nlit) return;
    uint32_t bad = 0;
    if (v.ot_len[lit] != len2[lit]) bad = 1;
    else for (uint32_t i = 0; i < len2[lit]; ++i) if (v.ot_data[v.ot_off[lit] + i] != data2[off2[lit] + i]) { bad = 1; break; }
    if (bad) atomicAdd(mismatch, 1u);
}

/* ---------------------------------------------------------------------------------------------- */
/* DIMACS parser (dimacs.cpp:25-177) - bodies and the method in sigma_dimacs.cuh                     */
/* ---------------------------------------------------------------------------------------------- */
/* inclusive scan over the block (256 threads) of transition functions under composition, in thread order */
__device__ __forceinline__ uint32_t dm_block_scan_fn(uint32_t f, uint32_t *s_w /* 8 */, uint32_t *total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t g = __shfl_up_sync(0xffffffffu, f, d); if (lane >= (uint32_t)d) f = sgd_compose(g, f); }
    if (lane == 31) s_w[warp] = f;
    __syncthreads();
    uint32_t pre = SGD_IDENT;
    for (uint32_t w = 0; w < warp; ++w) pre = sgd_compose(pre, s_w[w]);
    f = sgd_compose(pre, f);
    if (total) { uint32_t t = SGD_IDENT; for (uint32_t w = 0; w < TPB / 32; ++w) t = sgd_compose(t, s_w[w]); *total = t; }
    __syncthreads();
    return f;
}
/* the 64 bytes of chunk i as four 16-byte words (the text buffer is 16-byte aligned and padded) */
__device__ __forceinline__ void dm_load_chunk(const uint8_t *__restrict__ text, uint32_t i, uint32_t w[16])
{
    const uint4 *p = (const uint4 *)text + (size_t)i * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const uint4 x = __ldg(p + q); w[4 * q] = x.x; w[4 * q + 1] = x.y; w[4 * q + 2] = x.z; w[4 * q + 3] = x.w; }
}
__global__ void __launch_bounds__(TPB) k_dm_fns(const uint8_t *__restrict__ text, uint64_t n, uint32_t nchunks, uint32_t *__restrict__ fn,
                                                uint32_t *__restrict__ tilefn)
{
    __shared__ uint32_t s_w[TPB / 32];
    __shared__ uint8_t s_cls[256];
    __shared__ uint32_t s_pair[SGD_NCLASS * SGD_NCLASS];          /* transition function of two bytes, by their classes */
    s_cls[threadIdx.x] = (uint8_t)sgd_class((uint8_t)threadIdx.x);
    if (threadIdx.x < SGD_NCLASS * SGD_NCLASS)
        s_pair[threadIdx.x] = sgd_compose(sgd_class_fn(threadIdx.x / SGD_NCLASS), sgd_class_fn(threadIdx.x % SGD_NCLASS));
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t f = SGD_IDENT;
    if (i < nchunks) {
        const uint64_t lo = (uint64_t)i * SGD_CHUNK;
        if (lo + SGD_CHUNK <= n) {
            uint32_t w[16];
            dm_load_chunk(text, i, w);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t x = w[k];
                f = sgd_compose(f, s_pair[s_cls[x & 255u] * SGD_NCLASS + s_cls[(x >> 8) & 255u]]);
                f = sgd_compose(f, s_pair[s_cls[(x >> 16) & 255u] * SGD_NCLASS + s_cls[x >> 24]]);
            }
        } else f = sgd_chunk_fn(text, lo, n);
        fn[i] = f;
    }
    uint32_t tot;
    (void)dm_block_scan_fn(f, s_w, &tot);
    if (threadIdx.x == 0) tilefn[blockIdx.x] = tot;
}
/* one block: start state of every tile (exclusive scan of the tile functions applied to B); tiles in chunks of 256 */
__global__ void __launch_bounds__(TPB) k_dm_tile_states(const uint32_t *__restrict__ tilefn, uint32_t ntiles, uint32_t *__restrict__ tstate,
                                                        uint32_t *__restrict__ final_state)
{
    __shared__ uint32_t s_w[TPB / 32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = SGD_IDENT;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += TPB) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t f = i < ntiles ? tilefn[i] : SGD_IDENT;
        uint32_t tot;
        const uint32_t incl = dm_block_scan_fn(f, s_w, &tot);
        const uint32_t carry = s_carry;
        /* exclusive prefix = carry . (inclusive of the previous thread) */
        const uint32_t prev = __shfl_up_sync(0xffffffffu, incl, 1);
        __shared__ uint32_t s_last[TPB / 32];
        if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t excl = (threadIdx.x & 31) ? prev : (threadIdx.x ? s_last[(threadIdx.x >> 5) - 1] : SGD_IDENT);
        if (i < ntiles) tstate[i] = sgd_apply(sgd_compose(carry, excl), SGD_B);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = sgd_compose(carry, tot);
        __syncthreads();
    }
    if (threadIdx.x == 0) *final_state = sgd_apply(s_carry, SGD_B);
}
__global__ void __launch_bounds__(TPB) k_dm_chunk_states(const uint32_t *__restrict__ fn, uint32_t nchunks, const uint32_t *__restrict__ tstate,
                                                         uint32_t *__restrict__ cstate)
{
    __shared__ uint32_t s_w[TPB / 32];
    __shared__ uint32_t s_last[TPB / 32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t f = i < nchunks ? fn[i] : SGD_IDENT;
    const uint32_t incl = dm_block_scan_fn(f, s_w, nullptr);
    const uint32_t prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = incl;
    __syncthreads();
    const uint32_t excl = (threadIdx.x & 31) ? prev : (threadIdx.x ? s_last[(threadIdx.x >> 5) - 1] : SGD_IDENT);
    if (i < nchunks) cstate[i] = sgd_apply(excl, tstate[blockIdx.x]);
}
/* The integers that start in chunk i, in one walk over its 64 bytes held in registers: the value is built digit by digit
 * as the state machine moves (no second loop per integer, so the lanes of a warp stay together); an integer that runs past
 * the end of the chunk is finished from global memory by the chunk it started in.  Same result as sgd_chunk_tokens. */
template <bool EMIT>
__global__ void __launch_bounds__(TPB) k_dm_tokens(const uint8_t *__restrict__ text, uint64_t n, uint32_t nchunks, const uint32_t *__restrict__ cstate,
                                                   uint32_t nvars, uint32_t *ntok, uint32_t *nterm, const uint32_t *__restrict__ tokoff,
                                                   const uint32_t *__restrict__ termoff, uint32_t *__restrict__ tok, uint32_t *__restrict__ term, uint32_t *bad)
{
    __shared__ uint8_t s_cls[256];
    __shared__ uint32_t s_fn[16];
    s_cls[threadIdx.x] = (uint8_t)sgd_class((uint8_t)threadIdx.x);
    if (threadIdx.x < SGD_NCLASS) s_fn[threadIdx.x] = sgd_class_fn(threadIdx.x);
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks) return;
    const uint64_t lo = (uint64_t)i * SGD_CHUNK;
    if (lo + SGD_CHUNK > n) {                                                     /* the last, partial chunk: the plain walk */
        uint32_t a, b;
        const uint32_t end = sgd_chunk_tokens(text, lo, n, n, cstate[i], nvars, EMIT, &a, &b, EMIT ? tokoff[i] : 0u, EMIT ? termoff[i] : 0u, tok, term, bad);
        if (!EMIT) { ntok[i] = a; nterm[i] = b; if (end == SGD_ERR) *bad = 2; }
        return;
    }
    uint32_t w[16];
    dm_load_chunk(text, i, w);
    uint32_t state = cstate[i], nt = 0, nz = 0, cur = 0, sign = 0;
    bool owned = false;
    const uint32_t tb = EMIT ? tokoff[i] : 0u, zb = EMIT ? termoff[i] : 0u;
    auto emit = [&]() {
        if (EMIT) {
            uint32_t v = cur;
            if (v > nvars) { *bad = 3; v = 1; }
            tok[tb + nt] = v ? 2u * v + sign : 0u;
            if (!v) term[zb + nz] = tb + nt;
        }
        ++nt;
        nz += cur == 0;
        owned = false;
    };
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t ch = (w[k] >> (8 * q)) & 255u;
            const uint32_t cls = s_cls[ch];
            if (sgd_token_start(state, cls)) {
                if (owned) emit();                           /* "1-2": a sign ends the integer before it */
                owned = true; cur = 0; sign = ch == '-';
            }
            if (owned && (cls == SGD_Z || cls == SGD_D)) cur = cur * 10u + (ch - '0');
            state = sgd_apply(s_fn[cls], state);
            if (owned && !sgd_in_token(state)) emit();
        }
    }
    if (owned) {                                             /* runs on into the next chunk(s) */
        uint64_t p = lo + SGD_CHUNK;
        while (p < n && text[p] >= '0' && text[p] <= '9') cur = cur * 10u + (uint32_t)(text[p++] - '0');
        emit();
    }
    if (!EMIT) { ntok[i] = nt; nterm[i] = nz; if (state == SGD_ERR) *bad = 2; }
}
/* one round of makeClause over all clauses from the units of the previous round; csize = literals left (0: dropped) */
__global__ void __launch_bounds__(TPB) k_dm_clauses(const uint32_t *__restrict__ tok, const uint32_t *__restrict__ term, uint32_t ncl,
                                                    const uint32_t *__restrict__ utime, uint32_t *__restrict__ utime_new, uint32_t *__restrict__ csize,
                                                    uint32_t *unsat)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncl) return;
    const uint32_t first = c ? term[c - 1] + 1 : 0u, last = term[c];
    uint32_t kind;
    const uint32_t n = sgd_make_clause(tok, first, last, c, utime, nullptr, &kind);
    /* 0xFFFFFFFF: every literal false, nothing satisfied - dropped, but it still runs into the "too many clauses" check */
    csize[c] = kind == 3 ? 0u : (kind == 0 && n == 0 && !sgd_satisfied_or_empty(tok, first, last, c, utime)) ? 0xFFFFFFFFu : n;
    if (kind == 3) atomicMin(unsat, c);
    if (kind == 1) atomicMin(&utime_new[sgd_unit_literal(tok, first, last, c, utime)], c);
}
__global__ void k_dm_diff(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t n, uint32_t *changed)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) *changed = 1;
}
/* per clause: words of its CLAUSE record (0: no record), 1 if it is kept, 1 if it is the unit that enqueues its literal */
__global__ void k_dm_sizes(const uint32_t *__restrict__ tok, const uint32_t *__restrict__ term, uint32_t ncl, const uint32_t *__restrict__ utime,
                           const uint32_t *__restrict__ csize, uint32_t *__restrict__ words, uint32_t *__restrict__ kept, uint32_t *__restrict__ isunit)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncl) return;
    const uint32_t n = csize[c] == 0xFFFFFFFFu ? 0u : csize[c];
    words[c] = n >= 2 ? n + 4u : 0u;
    kept[c] = n >= 2;
    uint32_t u = 0;
    if (n == 1) {
        const uint32_t first = c ? term[c - 1] + 1 : 0u;
        u = utime[sgd_unit_literal(tok, first, term[c], c, utime)] == c;      /* the first clause that made it true enqueues it */
    }
    isunit[c] = u;
}
__global__ void k_dm_write(const uint32_t *__restrict__ tok, const uint32_t *__restrict__ term, uint32_t ncl, const uint32_t *__restrict__ utime,
                           const uint32_t *__restrict__ csize, const uint32_t *__restrict__ wordoff, const uint32_t *__restrict__ keptoff,
                           const uint32_t *__restrict__ unitoff, const uint32_t *__restrict__ isunit, uint32_t *__restrict__ cm,
                           unsigned long long *__restrict__ orgs, uint32_t *__restrict__ trail, int8_t *__restrict__ value, uint8_t *__restrict__ state,
                           unsigned long long *__restrict__ stats /* binaries, ternaries, large, literals, max size */, uint32_t header_clauses,
                           uint32_t *bad)
{
    /* counters: one set of global atomics per block */
    __shared__ uint32_t s_st[5];
    if (threadIdx.x < 5) s_st[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < ncl;
    if (live && csize[c] == 0xFFFFFFFFu && keptoff[c] + 1 > header_clauses) *bad = 4;            /* dimacs.cpp:160 */
    const uint32_t n = live && csize[c] != 0xFFFFFFFFu ? csize[c] : 0u;
    const uint32_t first = live && c ? term[c - 1] + 1 : 0u, last = live ? term[c] : 0u;
    if (n >= 2) {
        uint32_t *dst = cm + wordoff[c];
        uint32_t kind;
        dst[0] = (1u << 4) | (n == 2 ? 1u << 5 : 0u); dst[1] = n; dst[2] = 2u; dst[3] = 0u;
        (void)sgd_make_clause(tok, first, last, c, utime, dst + 4, &kind);
        orgs[keptoff[c]] = wordoff[c];
        atomicAdd(&s_st[n == 2 ? 0 : n == 3 ? 1 : 2], 1u);
        atomicAdd(&s_st[3], n);
        atomicMax(&s_st[4], n);
    } else if (n == 1 && isunit[c]) {
        const uint32_t lit = sgd_unit_literal(tok, first, last, c, utime);
        trail[unitoff[c]] = lit;
        value[lit] = 1; value[lit ^ 1u] = 0;                  /* enqueueUnit, solve.hpp:373-398 */
        state[lit >> 1] = (uint8_t)SG_FROZEN_M;
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_st[threadIdx.x]) atomicAdd(&stats[threadIdx.x], (unsigned long long)s_st[threadIdx.x]);
    if (threadIdx.x == 4 && s_st[4]) atomicMax(&stats[4], (unsigned long long)s_st[4]);
}

/* ---------------------------------------------------------------------------------------------- */
/* rebuildWT (watch.cpp:135-157) from the CLAUSE records built by k_cm_write                          */
/* ---------------------------------------------------------------------------------------------- */
__global__ void k_wt_groups(const uint32_t *__restrict__ arena, const unsigned long long *__restrict__ refs, uint32_t n, uint32_t code,
                            uint32_t *__restrict__ f0, uint32_t *__restrict__ f1, uint32_t *__restrict__ f2, uint32_t *__restrict__ f3)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t *src = arena + refs[i];
    const uint32_t g = sg_wt_group(src[2] == 2u, (src[0] & SG_ST_MASK) == SG_LEARNT, code);
    f0[i] = g == 0; f1[i] = g == 1; f2[i] = g == 2; f3[i] = g == 3;
}
/* clause i becomes entry `seq` of the attach order: sortClause on its record, then its two watched literals as a
 * two-literal "clause" of a scratch store, so that the occurrence-table build (stable by entry number) yields the watch
 * lists in push order */
__global__ void k_wt_fill(uint32_t *__restrict__ cm, const uint32_t *__restrict__ arena, const unsigned long long *__restrict__ refs, uint32_t n,
                          uint32_t code, const uint32_t *__restrict__ p0, const uint32_t *__restrict__ p1, const uint32_t *__restrict__ p2,
                          const uint32_t *__restrict__ p3, const uint32_t *__restrict__ totals, const int8_t *__restrict__ value,
                          uint4 *__restrict__ hdr, uint32_t *__restrict__ lits, uint4 *__restrict__ winfo)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long r = refs[i];
    const uint32_t *src = arena + r;
    const uint32_t size = src[2];
    const uint32_t g = sg_wt_group(size == 2u, (src[0] & SG_ST_MASK) == SG_LEARNT, code);
    uint32_t seq = g == 0 ? p0[i] : g == 1 ? p1[i] : g == 2 ? p2[i] : p3[i];
    if (g > 0) seq += totals[0];
    if (g > 1) seq += totals[1];
    if (g > 2) seq += totals[2];
    const unsigned long long cref = r + i;
    uint32_t *l = cm + cref + SG_CM_HDR;
    if (size != 2u && value) sg_sort_clause(l, (int)size, value);               /* attachNonBins / attachClauses: if (!c.binary()) sortClause(c) */
    const uint32_t a = l[0], b = l[1];
    hdr[seq] = make_uint4(2u * seq, 2u, 0u, 0u);
    lits[2 * seq] = SG_FLIP(a); lits[2 * seq + 1] = SG_FLIP(b);        /* ATTACH_TWO_WATCHES: wt[FLIP(first)], wt[FLIP(second)] */
    winfo[seq] = make_uint4((uint32_t)cref, (uint32_t)(cref >> 32), size, a ^ b);
}
/* WATCH {C_REF ref; uint32 imp; int size} (watch.hpp:28-46) for every entry of every list; imp = the other watched literal */
__global__ void k_wt_records(uint32_t nlit, const uint32_t *__restrict__ ot_off, const uint32_t *__restrict__ ot_len, const uint32_t *__restrict__ ot_data,
                             const uint4 *__restrict__ winfo, uint4 *__restrict__ rec)
{
    const uint32_t lit = blockIdx.x * blockDim.x + threadIdx.x;
    if (lit >= nlit) return;
    const uint32_t o = ot_off[lit], m = ot_len[lit];
    for (uint32_t k = 0; k < m; ++k) {
        const uint4 w = winfo[ot_data[o + k]];
        rec[o + k] = make_uint4(w.x, w.y, w.w ^ SG_FLIP(lit), w.z);      /* a ^ b ^ watched = the other one */
    }
}

} // namespace

/* ================================================================================================ */
/* host side of the launch layer                                                                     */
/* ================================================================================================ */
int be_init(int device, void **stream_out, char *err, size_t errlen)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        snprintf(err, errlen, "no CUDA device: %s (libsigma_b200 has no CPU fallback)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return 1;
    }
    if (device < 0 || device >= n) { snprintf(err, errlen, "device %d out of range (%d devices)", device, n); return 1; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { snprintf(err, errlen, "cudaSetDevice: %s", cudaGetErrorString(e)); return 1; }
    BeStream *b = new (std::nothrow) BeStream();
    if (!b) { snprintf(err, errlen, "out of host memory"); return 1; }
    b->launches = 0; b->err[0] = 0;
    if ((e = cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking)) != cudaSuccess) { snprintf(err, errlen, "cudaStreamCreate: %s", cudaGetErrorString(e)); delete b; return 1; }
    *stream_out = (void *)b;
    return 0;
}
void be_shutdown(void *s) { if (s) { if (((BeStream *)s)->scratch) cudaFree(((BeStream *)s)->scratch); cudaStreamDestroy(ST(s)); delete (BeStream *)s; } }
void be_bind(int device) { cudaSetDevice(device); }
void *be_malloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
void be_free(void *p) { if (p) cudaFree(p); }
void *be_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
void be_host_free(void *p) { if (p) cudaFreeHost(p); }
int be_host_register(void *p, size_t bytes)
{
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
    return 0;
}
int be_host_unregister(void *p)
{
    if (cudaHostUnregister(p) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
    return 0;
}
int be_h2d(void *s, void *d, const void *h, size_t n) { if (n) cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, ST(s)); note_err(s, "h2d"); return 0; }
int be_d2h(void *s, void *h, const void *d, size_t n) { if (n) cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, ST(s)); note_err(s, "d2h"); return 0; }
int be_d2d(void *s, void *d, const void *src, size_t n) { if (n) cudaMemcpyAsync(d, src, n, cudaMemcpyDeviceToDevice, ST(s)); note_err(s, "d2d"); return 0; }
int be_memset(void *s, void *d, int b, size_t n) { if (n) cudaMemsetAsync(d, b, n, ST(s)); note_err(s, "memset"); return 0; }
int be_sync(void *s, char *err, size_t errlen)
{
    cudaError_t e = cudaStreamSynchronize(ST(s));
    BeStream *b = (BeStream *)s;
    if (e != cudaSuccess && !b->err[0]) snprintf(b->err, sizeof b->err, "stream sync: %s", cudaGetErrorString(e));
    if (b->err[0]) { snprintf(err, errlen, "%s", b->err); return 1; }
    return 0;
}
void *be_event_create(void) { cudaEvent_t e; cudaEventCreate(&e); return (void *)e; }
void be_event_destroy(void *e) { cudaEventDestroy((cudaEvent_t)e); }
void be_event_record(void *s, void *e) { cudaEventRecord((cudaEvent_t)e, ST(s)); }
float be_event_ms(void *a, void *b) { float ms = 0; if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) { (void)cudaGetLastError(); ms = 0; } return ms; }
uint64_t be_launch_count(void *s) { return ((BeStream *)s)->launches; }
void be_copy_words(void *s, void *dst, const void *src, size_t bytes)
{
    k_copy_words<<<1, 128, 0, ST(s)>>>((uint32_t *)dst, (const uint32_t *)src, (uint32_t)(bytes / 4));
    note_err(s, "k_copy_words");
}

void be_extract_sizes(void *s, const uint32_t *cm, uint64_t cm_buckets, const uint64_t *crefs, uint32_t n, const uint8_t *stencil, uint64_t stencil_size,
                      uint32_t *sizes, uint32_t *live, uint32_t *bad)
{
    cudaMemsetAsync(bad, 0, sizeof(uint32_t), ST(s));
    if (n) LAUNCH(k_extract_sizes, blocks_for(n), TPB, s, cm, cm_buckets, crefs, n, stencil, stencil_size, sizes, live, bad);
}
void be_extract_write(void *s, const uint32_t *cm, const uint64_t *crefs, uint32_t n, const uint32_t *sizes, const uint32_t *newoff,
                      const uint32_t *newid, uint32_t *arena, uint64_t *refs)
{
    if (n) LAUNCH(k_extract_write, blocks_for(n), TPB, s, cm, crefs, n, sizes, newoff, newid, arena, (unsigned long long *)refs);
}
void be_cm_flags(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t *learnt)
{
    if (n) LAUNCH(k_cm_flags, blocks_for(n), TPB, s, arena, (const unsigned long long *)refs, n, learnt);
}
void be_cm_write(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, const uint32_t *learnt_before,
                 int lbd_tier1, uint32_t *cm, uint64_t *orgs, uint64_t *learnts, unsigned long long *counts, uint8_t *subsume, const uint32_t *vmap)
{
    cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), ST(s));
    if (n) LAUNCH(k_cm_write, blocks_for(n), TPB, s, arena, (const unsigned long long *)refs, n, learnt_before, lbd_tier1, cm,
                  (unsigned long long *)orgs, (unsigned long long *)learnts, counts, subsume, (unsigned long long)(n_buckets + n), vmap);
}
void be_ere_maxsz(void *s, const SgView &v, const uint8_t *size8, uint8_t *maxsz) { if (v.nlit) LAUNCH(k_ere_maxsz, blocks_for(v.nlit), TPB, s, v, size8, maxsz); }
void be_ere_filt(void *s, const SgView &v, uint32_t *filt, uint8_t *size8)
{
    if (v.n_cls) LAUNCH(k_ere_filt, blocks_for(v.n_cls), TPB, s, v, (uint2 *)filt, size8);
}
void be_ot_order(void *s, const SgView &v)
{
    if (!v.nlit) return;
    const char *e = getenv("SIGMA_OT_ORDER_STAGED");
    if (e && e[0] == '0') LAUNCH(k_items<SGK_IDSORT>, blocks_for(v.nlit), TPB, s, v, v.nlit);
    else {
        LAUNCH(k_ot_order_staged, blocks_for(v.nlit), TPB, s, v);
        LAUNCH(k_items<SGK_IDSORT_LONG>, blocks_for(v.nlit), TPB, s, v, v.nlit);
    }
}
void be_ingest_sizes(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, uint32_t *sizes, uint32_t *bad)
{
    cudaMemsetAsync(bad, 0, sizeof(uint32_t), ST(s));
    if (n) LAUNCH(k_ingest_sizes, blocks_for(n), TPB, s, arena, refs, n, n_buckets, sizes, bad);
}
void be_ingest_fused(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint64_t n_buckets, sg_u4 *hdr, uint32_t *lits, uint32_t *gap)
{
    cudaMemsetAsync(gap, 0, sizeof(uint32_t), ST(s));
    if (n) LAUNCH(k_ingest_fused, blocks_for(n), TPB, s, arena, refs, n, n_buckets, hdr, lits, gap);
}
void be_ingest_copy(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, const uint32_t *offs, sg_u4 *hdr, uint32_t *lits)
{
    if (n) LAUNCH(k_ingest_copy, blocks_for(n), TPB, s, arena, refs, n, offs, hdr, lits);
}

size_t be_scan_tmp_words(uint32_t n) { return (size_t)(n / SCAN_TILE) + 2; }
void be_scan_u32(void *s, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total, uint32_t *tmp)
{
    if (!n) { if (total) LAUNCH(k_zero_u32, 1, 1, s, total); return; }
    const uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    LAUNCH(k_scan_reduce, nb, TPB, s, in, n, tmp);
    LAUNCH(k_scan_partials, 1, TPB, s, tmp, nb, total);
    LAUNCH(k_scan_down, nb, TPB, s, in, out, n, tmp);
}

void be_scan_multi(void *s, const SgScanSet &ss, uint32_t count, uint32_t n, uint32_t *tmp)
{
    if (!count) return;
    if (!n) { for (uint32_t k = 0; k < count; ++k) if (ss.total[k]) LAUNCH(k_zero_u32, 1, 1, s, ss.total[k]); return; }
    const uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    LAUNCH(k_scan_reduce_multi, dim3(nb, count), TPB, s, ss, n, nb, tmp);
    LAUNCH(k_scan_partials_multi, count, TPB, s, ss, nb, tmp);
    LAUNCH(k_scan_down_multi, dim3(nb, count), TPB, s, ss, n, nb, tmp);
}

void be_hist(void *s, const SgView &v, uint32_t *counts)
{
    if (v.n_cls) LAUNCH(k_hist, blocks_for(v.n_cls), TPB, s, (const uint4 *)v.hdr, v.lits, v.n_cls, counts);
}
/* geometry of the two-level build; returns false when the formula does not fit it (caller: one-level path) */
static bool otb_geometry(uint32_t nlit, uint32_t n_cls, uint64_t lits, uint32_t *shift, uint32_t *NB, uint32_t *T, uint32_t *per_tile)
{
    if (!nlit || !n_cls) return false;
    if (!lits) lits = 1;
    /* buckets: as wide as possible (long runs per tile and bucket = whole sectors in the scatter) while an average bucket
     * takes at most 60 % of the lists kernel's staging room and every SM has two buckets to work on */
    uint32_t sh = 12, best = 0;
    for (; sh >= 6; --sh) {
        const uint64_t nb = ((uint64_t)nlit + (1u << sh) - 1) >> sh;
        if (nb > OTB_NBMAX) break;
        const uint64_t avg = lits / nb + 1;
        const uint64_t pcap = OTB_SMEM_WORDS - 2ull * (1u << sh) - 16 - 3 * (OTB_TMAX / 4 + 4);
        if (avg * 10 <= pcap * 6) { best = sh; if (nb >= 296) break; }
    }
    if (!best) return false;
    sh = best;
    { const char *e = getenv("SIGMA_OTB_SHIFT"); if (e) { const uint32_t x = (uint32_t)atoi(e); if (x >= 6 && x <= 12 && (((uint64_t)nlit + (1u << x) - 1) >> x) <= OTB_NBMAX) sh = x; } }   /* experiments */
    const uint64_t nb = ((uint64_t)nlit + (1u << sh) - 1) >> sh;
    /* tiles: a multiple of the SM count, an average tile fills 80 % of the scatter's staging room; a multiple of 32 clauses
     * (bitmap words) */
    const uint64_t rcap = OTB_SMEM_WORDS - 4 * nb - 24;
    uint64_t tiles = (lits * 10 + rcap * 8 - 1) / (rcap * 8);
    tiles = (tiles + 147) / 148 * 148;
    if (tiles > OTB_TMAX) return false;
    uint32_t pt = (uint32_t)((n_cls + tiles - 1) / tiles);
    pt = (pt + 31u) & ~31u;
    if ((uint64_t)pt >= (1ull << (32 - sh)) - 1) return false;          /* clause within the tile: 32 - shift bits of a record */
    tiles = ((uint64_t)n_cls + pt - 1) / pt;
    if (nb * tiles >= 0xFFFFFFF0ull) return false;
    *shift = sh; *NB = (uint32_t)nb; *T = (uint32_t)tiles; *per_tile = pt;
    return true;
}
size_t be_ot_build_tmp_words(uint32_t nlit, uint32_t n_cls, uint64_t lits)
{
    uint32_t sh, NB, T, pt;
    if (!otb_geometry(nlit, n_cls, lits, &sh, &NB, &T, &pt)) return 0;
    const size_t cnt = (size_t)NB * T + 8;
    return cnt + be_scan_tmp_words((uint32_t)cnt) + 8 + (size_t)lits + 64;
}
int be_ot_build_plan(uint32_t nlit, uint32_t n_cls, uint64_t lits, uint32_t *tmp, SgOtBuild *p)
{
    if (!otb_geometry(nlit, n_cls, lits, &p->shift, &p->NB, &p->T, &p->per_tile)) return 1;
    static int smem_ok = -1;
    if (smem_ok < 0) {
        smem_ok = cudaFuncSetAttribute(k_otb_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(OTB_SMEM_WORDS * 4)) == cudaSuccess &&
                  cudaFuncSetAttribute(k_otb_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(OTB_SMEM_WORDS * 4)) == cudaSuccess;
        (void)cudaGetLastError();
    }
    if (!smem_ok) return 1;
    p->ncnt = (size_t)p->NB * p->T + 1;                 /* + the word that receives the total */
    p->cnt = tmp;
    p->scantmp = tmp + p->ncnt + 7;
    uint32_t *after = p->scantmp + be_scan_tmp_words((uint32_t)(p->ncnt + 7));
    p->pairs = (void *)(((uintptr_t)after + 15) & ~(uintptr_t)15);
    return 0;
}
static unsigned otb_grid(uint32_t T, unsigned per_sm) { return T < 148u * per_sm ? T : 148u * per_sm; }
void be_ot_build_count(void *s, const SgView &v, const SgOtBuild &p, uint32_t *bits)
{
    k_otb_count<<<otb_grid(p.T, 2), OTB_TPB, p.NB * 4, ST(s)>>>((const uint4 *)v.hdr, v.lits, v.n_cls, p.shift, p.NB, p.T, p.per_tile, p.cnt, bits);
    ((BeStream *)s)->launches++; note_err(s, "k_otb_count");
}
void be_ot_build_scatter(void *s, const SgView &v, const SgOtBuild &p)
{
    uint32_t rcap = OTB_SMEM_WORDS - 4 * p.NB - 24;
    { const char *e = getenv("SIGMA_OTB_CAPS"); if (e && (uint32_t)atol(e) < rcap) rcap = (uint32_t)atol(e); }   /* test hook: unstaged tiles */
    k_otb_scatter<<<otb_grid(p.T, 1), OTB_TPB, OTB_SMEM_WORDS * 4, ST(s)>>>((const uint4 *)v.hdr, v.lits, v.n_cls, p.shift, p.NB, p.T, p.per_tile, rcap,
                                                                           p.cnt, (uint32_t *)p.pairs);
    ((BeStream *)s)->launches++; note_err(s, "k_otb_scatter");
}
void be_ot_build_lists(void *s, const SgView &v, const SgOtBuild &p, const uint32_t *total)
{
    (void)total;                                        /* the scan left it behind the last count */
    uint32_t pcap = OTB_SMEM_WORDS - 2 * (1u << p.shift) - 16 - 3 * ((p.T + 4) & ~3u) - 4;
    { const char *e = getenv("SIGMA_OTB_CAPS"); if (e && (uint32_t)atol(e) < pcap) pcap = (uint32_t)atol(e); }   /* test hook: unstaged buckets */
    k_otb_lists<<<p.NB, OTB_TPB, OTB_SMEM_WORDS * 4, ST(s)>>>((const uint32_t *)p.pairs, p.cnt, p.NB, p.T, p.shift, p.per_tile, pcap, v.nlit,
                                                             v.ot_off, v.ot_len, v.ot_data);
    ((BeStream *)s)->launches++; note_err(s, "k_otb_lists");
}
void be_scatter(void *s, const SgView &v, uint32_t *cursor)
{
    if (!v.n_cls) return;
    long l2_mb = -1;               /* knobs are read per call: tests flip them inside one process */
    int stream_in = -1;
    if (l2_mb < 0) { const char *e = getenv("SIGMA_SCATTER_L2_MB"); l2_mb = e ? atol(e) : 64; if (l2_mb < 1) l2_mb = 1; }
    if (stream_in < 0) { const char *e = getenv("SIGMA_SCATTER_CS"); stream_in = e ? atoi(e) : 0; }
    const uint64_t bytes = 4ull * v.pool_used;
    uint32_t passes = (uint32_t)((bytes + ((uint64_t)l2_mb << 20) - 1) / ((uint64_t)l2_mb << 20));
    if (passes < 1) passes = 1;
    if (passes > 64) passes = 64;
    const uint32_t step = (v.nlit + passes - 1) / passes;
    for (uint32_t p = 0; p < passes; ++p) {
        const uint32_t lo = p * step, hi = (p + 1 == passes) ? 0xFFFFFFFFu : (p + 1) * step;
        if (stream_in) LAUNCH(k_scatter<true>, blocks_for(v.n_cls), TPB, s, (const uint4 *)v.hdr, v.lits, v.n_cls, cursor, v.ot_data, lo, hi);
        else LAUNCH(k_scatter<false>, blocks_for(v.n_cls), TPB, s, (const uint4 *)v.hdr, v.lits, v.n_cls, cursor, v.ot_data, lo, hi);
    }
}
void be_scores(void *s, const SgView &v)
{
    cudaMemsetAsync(&v.sc->max_score, 0, sizeof(uint32_t), ST(s));
    if (v.nvars) LAUNCH(k_scores, blocks_for(v.nvars), TPB, s, v);
}

size_t be_sort_tmp_words(uint32_t n)
{
    const size_t nb = ((size_t)n + RS_TILE - 1) / RS_TILE;
    const size_t sep = 256 * nb + be_scan_tmp_words((uint32_t)(256 * nb)) + 16, fused = 256 * 2048;     /* k_vo_sort: 256 x grid */
    return sep > fused ? sep : fused;
}
void be_sort_pairs(void *s, uint32_t *keys, uint32_t *vals, uint32_t n, uint32_t *tmpk, uint32_t *tmpv, uint32_t *tmp, uint32_t key_bits)
{
    if (n < 2) return;
    const uint32_t nb = (n + RS_TILE - 1) / RS_TILE;
    uint32_t *hist = tmp, *scantmp = tmp + 256ull * nb;
    uint32_t *ik = keys, *iv = vals, *ok = tmpk, *ov = tmpv;
    int passes = 0;
    for (uint32_t shift = 0; shift < key_bits; shift += 8) {
        LAUNCH(k_rs_hist, nb, TPB, s, ik, n, (int)shift, nb, hist);
        be_scan_u32(s, hist, hist, 256 * nb, nullptr, scantmp);
        LAUNCH(k_rs_scatter, nb, TPB, s, ik, iv, n, (int)shift, nb, hist, ok, ov);
        uint32_t *t = ik; ik = ok; ok = t;
        t = iv; iv = ov; ov = t;
        passes++;
    }
    if (passes & 1) {
        be_d2d(s, keys, ik, (size_t)n * 4);
        be_d2d(s, vals, iv, (size_t)n * 4);
    }
}
/* scores + sort + rank in one cooperative launch (k_vo_sort); returns nonzero when that is not possible (caller: separate launches).
 * tmp: 256 x grid words (be_sort_tmp_words covers it). */
int be_vo_sort(void *s, const SgView &v, uint32_t *tmpk, uint32_t *tmpv, uint32_t *tmp, uint32_t forced_bits)
{
    static int ok = -1, sms = 0;
    const int smem = (32 + 8) * 256 * 4;
    if (ok < 0) {
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        ok = coop && cudaFuncSetAttribute(k_vo_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
             cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vo_sort, VO_TPB, smem) == cudaSuccess && per_sm >= 1;
        (void)cudaGetLastError();
    }
    if (!ok || !v.nvars || sms > 2048) return 1;
    cudaMemsetAsync(&v.sc->max_score, 0, sizeof(uint32_t), ST(s));
    SgView vv = v;
    void *args[] = { (void *)&vv, (void *)&tmpk, (void *)&tmpv, (void *)&tmp, (void *)&forced_bits };
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_vo_sort, dim3((unsigned)sms), dim3(VO_TPB), args, smem, ST(s));
    if (e != cudaSuccess) { (void)cudaGetLastError(); ok = 0; return 1; }
    ((BeStream *)s)->launches++;
    return 0;
}
void be_rank(void *s, const SgView &v) { if (v.nvars) LAUNCH(k_rank, blocks_for(v.nvars), TPB, s, v); }

void be_elect_batch(void *s, const SgView &v, uint32_t r0, uint32_t r1, uint32_t *wl)
{
    if (r1 > r0) LAUNCH(k_elect_batch, blocks_for(r1 - r0), TPB, s, v, r0, r1, wl);
}
void be_elect_round(void *s, const SgView &v, const uint32_t *wl_in, uint32_t *wl_out, int parity)
{
    cudaMemsetAsync(&v.sc->wl_count[parity ^ 1], 0, sizeof(uint32_t), ST(s));
    unsigned grid = blocks_for(32ull * v.nvars, 128);
    if (grid > 148u * 16u) grid = 148u * 16u;
    if (!grid) grid = 1;
    LAUNCH(k_elect_round, grid, 128, s, v, wl_in, wl_out, parity);
}
int be_elect_all(void *s, const SgView &v, uint32_t *wl0, uint32_t *wl1, uint32_t b0, uint32_t grow)
{
    static int blocks_per_sm = -1, sms = 0;
    if (blocks_per_sm < 0) {
        int dev = 0, coop = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_elect_persistent, SG_ELECT_TPB, 0) != cudaSuccess) blocks_per_sm = 0;
        (void)cudaGetLastError();
    }
    if (blocks_per_sm <= 0 || !v.nvars) return 1;
    SgView vv = v;
    void *args[] = { (void *)&vv, (void *)&wl0, (void *)&wl1, (void *)&b0, (void *)&grow };
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_elect_persistent, dim3((unsigned)(blocks_per_sm * sms)), dim3(SG_ELECT_TPB), args, 0, ST(s));
    if (e != cudaSuccess) { (void)cudaGetLastError(); blocks_per_sm = 0; return 1; }
    ((BeStream *)s)->launches++;
    return 0;
}
int be_elect_async(void *s, const SgView &v)
{
    static int blocks_per_sm = -1, sms = 0;
    if (blocks_per_sm < 0) {
        int dev = 0, coop = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_elect_async, SG_ELECT_TPB, 0) != cudaSuccess) blocks_per_sm = 0;
        (void)cudaGetLastError();
    }
    if (blocks_per_sm <= 0 || !v.nvars) return 1;
    SgView vv = v;
    void *args[] = { (void *)&vv };
    /* cooperative launch: the kernel has no grid-wide barrier, but its warps wait for each other, so all of them must be resident */
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_elect_async, dim3((unsigned)(blocks_per_sm * sms)), dim3(SG_ELECT_TPB), args, 0, ST(s));
    if (e != cudaSuccess) { (void)cudaGetLastError(); blocks_per_sm = 0; return 1; }
    ((BeStream *)s)->launches++;
    return 0;
}
void be_acting_flags(void *s, const SgView &v, uint32_t *fa, uint32_t *fe) { if (v.nvars) LAUNCH(k_acting_flags, blocks_for(v.nvars), TPB, s, v, fa, fe); }
void be_acting_scatter(void *s, const SgView &v, const uint32_t *fa, const uint32_t *pa, const uint32_t *fe, const uint32_t *pe)
{
    if (v.nvars) LAUNCH(k_acting_scatter, blocks_for(v.nvars), TPB, s, v, fa, pa, fe, pe);
}

void be_items(void *s, int kind, const SgView &v, uint32_t n)
{
    if (!n) return;
    switch (kind) {
    case SGK_IDSORT: LAUNCH(k_items<SGK_IDSORT>, blocks_for(n), TPB, s, v, n); break;
    case SGK_REDUCE: LAUNCH(k_items<SGK_REDUCE>, blocks_for(n), TPB, s, v, n); break;
    case SGK_ELECT_PREP: LAUNCH(k_items<SGK_ELECT_PREP>, blocks_for(n), TPB, s, v, n); break;
    case SGK_SORTOT: LAUNCH(k_items<SGK_SORTOT>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_SORTOT_LONG: LAUNCH(k_items<SGK_SORTOT_LONG>, blocks_for(n, 64), 64, s, v, n); break;
    case SGK_SCRATCH_LEN: LAUNCH(k_items<SGK_SCRATCH_LEN>, blocks_for(n), TPB, s, v, n); break;
    case SGK_SUB: LAUNCH(k_items<SGK_SUB>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_VE_COUNT: LAUNCH(k_items<SGK_VE_COUNT>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_VE_EMIT: LAUNCH(k_items<SGK_VE_EMIT>, blocks_for(n, 64), 64, s, v, n); break;
    case SGK_NBH: LAUNCH(k_items<SGK_NBH>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_BCE_COUNT: LAUNCH(k_items<SGK_BCE_COUNT>, blocks_for(n, 64), 64, s, v, n); break;
    case SGK_BCE_EMIT: LAUNCH(k_items<SGK_BCE_EMIT>, blocks_for(n, 64), 64, s, v, n); break;
    case SGK_ERE_CHUNKS: LAUNCH(k_items<SGK_ERE_CHUNKS>, blocks_for(n), TPB, s, v, n); break;
    case SGK_ERE_PROPOSE: LAUNCH(k_items<SGK_ERE_PROPOSE>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_ERE_DIRTY: LAUNCH(k_items<SGK_ERE_DIRTY>, blocks_for(n, 128), 128, s, v, n); break;
    case SGK_ERE_CLAIM: LAUNCH(k_items<SGK_ERE_CLAIM>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_ERE_SAFE: LAUNCH(k_items<SGK_ERE_SAFE>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_ERE_APPLY: LAUNCH(k_items<SGK_ERE_APPLY>, blocks_for(32ull * n, 128), 128, s, v, n); break;
    case SGK_ERE_FLAG01: LAUNCH(k_items<SGK_ERE_FLAG01>, blocks_for(n), TPB, s, v, n); break;
    case SGK_ERE_LIST: LAUNCH(k_items<SGK_ERE_LIST>, blocks_for(n), TPB, s, v, n); break;
    default: { BeStream *b = (BeStream *)s; snprintf(b->err, sizeof b->err, "be_items: bad kind %d", kind); }
    }
}
void be_serial(void *s, int kind, const SgView &v, uint32_t arg, uint32_t arg2)
{
    unsigned threads = 32;
    if (kind == SGS_ERE_REPLAY) {
        int helpers = -1;
        if (helpers < 0) { const char *e = getenv("SIGMA_ERE_HELPERS"); helpers = e ? atoi(e) : 15; if (helpers < 0) helpers = 0; if (helpers > 31) helpers = 31; }
        threads = 32u * (1u + (unsigned)helpers);
    }
    LAUNCH(k_serial, 1, threads, s, v, kind, arg, arg2);
}

void be_apply_units(void *s, const SgView &v, uint32_t n, uint32_t *first, uint32_t *flags, uint32_t *pos, uint32_t *total, uint32_t *scan_tmp)
{
    LAUNCH(k_au_first, blocks_for(n), TPB, s, v, n, first);
    LAUNCH(k_au_flags, blocks_for(n), TPB, s, v, n, (const uint32_t *)first, flags);
    be_scan_u32(s, flags, pos, n, total, scan_tmp);
    LAUNCH(k_au_write, blocks_for(n), TPB, s, v, n, first, (const uint32_t *)flags, (const uint32_t *)pos, (const uint32_t *)total);
    LAUNCH(k_au_finish, blocks_for(n), TPB, s, v, n, first, (const uint32_t *)total);
}
int be_has_prop_frontier(void) { return 1; }
static unsigned pp_grid(uint32_t n_warps) { unsigned g = blocks_for((uint64_t)n_warps * 32u); return g > 148u * 8u ? 148u * 8u : (g ? g : 1u); }
void be_pp_mark(void *s, const SgView &v, const SgProp &pp) { LAUNCH(k_pp_mark, blocks_for(pp.p1 - pp.p0), TPB, s, v, pp); }
void be_pp_round(void *s, const SgView &v, const SgProp &pp, int diff)
{
    LAUNCH(k_pp_visits<0>, pp_grid(pp.p1 - pp.p0), TPB, s, v, pp);
    if (diff) LAUNCH(k_pp_diff, 148, TPB, s, pp);
}
void be_pp_next(void *s, const SgProp &pp)
{
    LAUNCH(k_pp_forget, 148, TPB, s, pp.disc, pp.touched, pp.ctl);
    LAUNCH(k_pp_shift, 1, 1, s, pp);
}
void be_pp_count(void *s, const SgView &v, const SgProp &pp, uint32_t *scan_tmp)
{
    LAUNCH(k_pp_visits<1>, pp_grid(pp.p1 - pp.p0), TPB, s, v, pp);
    be_scan_u32(s, pp.cnt, pp.base, pp.p1 - pp.p0, pp.ctl + 4, scan_tmp);
}
void be_pp_commit(void *s, const SgView &v, const SgProp &pp, uint32_t n_new)
{
    if (n_new) LAUNCH(k_pp_visits<2>, pp_grid(pp.p1 - pp.p0), TPB, s, v, pp);
    LAUNCH(k_pp_rewrite, pp_grid(pp.p1 - pp.p0), TPB, s, v, pp);
    LAUNCH(k_pp_finish, pp_grid(pp.p1 - pp.p0), TPB, s, v, pp, n_new, 0);
    /* guess and report are equal here: both go back to "nothing discovered" */
    LAUNCH(k_pp_forget, 148, TPB, s, pp.disc, pp.touched, pp.ctl);
    LAUNCH(k_pp_forget, 148, TPB, s, pp.disc_new, pp.touched_new, pp.ctl + 1);
}
void be_pp_conflict(void *s, const SgView &v, const SgProp &pp)
{
    LAUNCH(k_pp_finish, pp_grid(pp.p1 - pp.p0), TPB, s, v, pp, 0u, 1);
    LAUNCH(k_pp_forget, 148, TPB, s, pp.disc, pp.touched, pp.ctl);
    LAUNCH(k_pp_forget, 148, TPB, s, pp.disc_new, pp.touched_new, pp.ctl + 1);
}

void be_sort_units(void *s, const SgView &v, uint32_t n)
{
    if (n < 2) return;
    const char *force = getenv("SIGMA_UNITS_RADIX");          /* test hook: take the large-n path for any n */
    const bool radix = force && force[0] == '1';
    if (n <= 32 && !radix) { LAUNCH(k_sort_units_small, 1, 1, s, v, n); return; }
    if (n <= 2048 && !radix) { LAUNCH(k_sort_units_block, 1, 512, s, v, n); return; }
    /* rare: LSD by (seq, eidx) through an index permutation; the scratch memory belongs to the context and only ever grows */
    BeStream *b = (BeStream *)s;
    const size_t words = 4 * (size_t)n + be_sort_tmp_words(n) + 64, bytes = 4 * words + sizeof(SgUnit) * (size_t)n + 256;
    if (b->scratch_bytes < bytes) {
        cudaStreamSynchronize(ST(s));
        if (b->scratch) cudaFree(b->scratch);
        b->scratch = be_malloc(bytes + bytes / 2);
        b->scratch_bytes = b->scratch ? bytes + bytes / 2 : 0;
    }
    if (!b->scratch) { snprintf(b->err, sizeof b->err, "be_sort_units: out of memory"); return; }
    uint32_t *keys = (uint32_t *)b->scratch, *idx = keys + n, *tk = idx + n, *tv = tk + n, *tmp = tv + n;
    SgUnit *ut = (SgUnit *)(((uintptr_t)(tmp + be_sort_tmp_words(n) + 64) + 15) & ~(uintptr_t)15);
    for (int which = 0; which < 2; ++which) {
        LAUNCH(k_units_keys, blocks_for(n), TPB, s, v, n, keys, idx, which);
        be_sort_pairs(s, keys, idx, n, tk, tv, tmp, 32);
    }
    LAUNCH(k_units_gather, blocks_for(n), TPB, s, v, n, idx, ut);
    be_d2d(s, v.ubuf, ut, sizeof(SgUnit) * (size_t)n);
}

void be_count_live(void *s, const SgView &v)
{
    cudaMemsetAsync(&v.sc->live_cls, 0, sizeof(uint32_t), ST(s));
    cudaMemsetAsync(&v.sc->live_lits, 0, sizeof(unsigned long long), ST(s));
    if (v.n_cls) {
        unsigned grid = blocks_for(v.n_cls);
        if (grid > 148u * 8u) grid = 148u * 8u;
        LAUNCH(k_count_live, grid, TPB, s, v);
    }
}
void be_count_melted(void *s, const SgView &v, uint32_t *out)
{
    LAUNCH(k_zero_u32, 1, 1, s, out);
    LAUNCH(k_count_melted, blocks_for((uint64_t)v.nvars + 1), TPB, s, v, out);
}
size_t be_live_tmp_words(uint32_t n_cls) { return 2 * ((size_t)n_cls / CP_TILE + 2); }
void be_live_scan(void *s, const SgView &v, uint32_t *tmp, uint32_t *totals)
{
    const uint32_t nb = (v.n_cls + CP_TILE - 1) / CP_TILE;
    if (!nb) { LAUNCH(k_zero_u32, 1, 1, s, totals); LAUNCH(k_zero_u32, 1, 1, s, totals + 1); return; }
    LAUNCH(k_live_reduce, nb, TPB, s, v, (uint2 *)tmp);
    LAUNCH(k_live_partials, 1, TPB, s, (uint2 *)tmp, nb, totals);
}
void be_compact_apply(void *s, const SgView &v, const uint32_t *tmp, sg_u4 *hdr2, uint32_t *lits2)
{
    const uint32_t nb = (v.n_cls + CP_TILE - 1) / CP_TILE;
    if (nb) LAUNCH(k_live_apply<false>, nb, TPB, s, v, (const uint2 *)tmp, hdr2, lits2, (uint32_t *)nullptr, (unsigned long long *)nullptr);
}
void be_export_apply(void *s, const SgView &v, const uint32_t *tmp, uint32_t *arena, uint64_t *refs)
{
    const uint32_t nb = (v.n_cls + CP_TILE - 1) / CP_TILE;
    int staged = -1;
    if (staged < 0) { const char *e = getenv("SIGMA_EXPORT_STAGED"); staged = e ? atoi(e) : 1; }
    if (!nb) return;
    if (staged) LAUNCH(k_export_staged, nb, TPB, s, v, (const uint2 *)tmp, arena, (unsigned long long *)refs);
    else LAUNCH(k_live_apply<true>, nb, TPB, s, v, (const uint2 *)tmp, (uint4 *)nullptr, (uint32_t *)nullptr, arena, (unsigned long long *)refs);
}

size_t be_ot_update_tmp_words(uint32_t nlit, uint32_t n_cls) { (void)n_cls; return 3 * (size_t)nlit + 64; }
size_t be_ot_bits_words(uint32_t n_cls) { return (size_t)n_cls / 32 + 2; }
void be_ot_bitmap(void *s, const SgView &v, uint32_t *bits)
{
    LAUNCH(k_otu_bitmap, blocks_for(((uint64_t)v.n_cls / 32 + 1) * 32), TPB, s, v, bits);
}
void be_ot_update_plan(void *s, const SgView &v, const SgOtUpdate &u)
{
    const OtuTmp t = otu_layout(u.tmp, v.nlit);
    cudaMemsetAsync(t.addc, 0, (size_t)v.nlit * 4, ST(s));
    cudaMemsetAsync(t.flag, 0, (size_t)v.nlit * 4, ST(s));
    if (u.total) cudaMemsetAsync(u.total, 0, 4, ST(s));
    LAUNCH(k_otu_mark, blocks_for(((uint64_t)v.n_cls / 128 + 1) * 32), TPB, s, v, u.bits, u.first_new, t.addc, t.flag);
    if (u.n_add) LAUNCH(k_otu_pairs_count, blocks_for(u.n_add), TPB, s, v, u.n_add, t.addc, t.flag);
    if (u.n_elected) LAUNCH(k_otu_elected, blocks_for(u.n_elected), TPB, s, v, u.n_elected, t.flag);
    if (u.total) LAUNCH(k_otu_need, blocks_for(v.nlit), TPB, s, v, t.addc, u.total);
}
void be_ot_update(void *s, const SgView &v, const SgOtUpdate &u)
{
    const OtuTmp t = otu_layout(u.tmp, v.nlit);
    LAUNCH(k_otu_apply, blocks_for(v.nlit), TPB, s, v, u.bits, t.addc, t.cursor, t.flag, u.bump);
    if (v.n_cls > u.first_new) LAUNCH(k_otu_append_clauses, blocks_for(v.n_cls - u.first_new), TPB, s, v, u.first_new, t.cursor);
    if (u.n_add) LAUNCH(k_otu_append_pairs, blocks_for(u.n_add), TPB, s, v, u.n_add, t.cursor);
    LAUNCH(k_otu_sort, blocks_for(v.nlit), TPB, s, v, t.flag);
}
void be_ot_compare(void *s, const SgView &v, const uint32_t *off2, const uint32_t *len2, const uint32_t *data2, uint32_t *mismatch)
{
    LAUNCH(k_zero_u32, 1, 1, s, mismatch);
    LAUNCH(k_ot_compare, blocks_for(v.nlit), TPB, s, v, off2, len2, data2, mismatch);
}

namespace { __global__ void k_iota_u32(uint32_t *p, uint32_t n) { const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; } }
void be_iota_u32(void *s, uint32_t *p, uint32_t n) { if (n) LAUNCH(k_iota_u32, blocks_for(n), TPB, s, p, n); }
/* ---- DIMACS parser launches (sigma_backend.h) ------------------------------------------------------- */
void be_dm_chunk_states(void *s, const uint8_t *text, uint64_t n, uint32_t nchunks, uint32_t *fn, uint32_t *tilefn, uint32_t *tstate, uint32_t *cstate,
                        uint32_t *final_state)
{
    const uint32_t ntiles = blocks_for(nchunks);
    LAUNCH(k_dm_fns, ntiles, TPB, s, text, n, nchunks, fn, tilefn);
    LAUNCH(k_dm_tile_states, 1, TPB, s, tilefn, ntiles, tstate, final_state);
    LAUNCH(k_dm_chunk_states, ntiles, TPB, s, fn, nchunks, tstate, cstate);
}
void be_dm_tokens(void *s, const uint8_t *text, uint64_t n, uint32_t nchunks, const uint32_t *cstate, uint32_t nvars, int emit, uint32_t *ntok,
                  uint32_t *nterm, const uint32_t *tokoff, const uint32_t *termoff, uint32_t *tok, uint32_t *term, uint32_t *bad)
{
    if (emit) LAUNCH(k_dm_tokens<true>, blocks_for(nchunks), TPB, s, text, n, nchunks, cstate, nvars, ntok, nterm, tokoff, termoff, tok, term, bad);
    else LAUNCH(k_dm_tokens<false>, blocks_for(nchunks), TPB, s, text, n, nchunks, cstate, nvars, ntok, nterm, tokoff, termoff, tok, term, bad);
}
void be_dm_clauses(void *s, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, uint32_t *utime_new, uint32_t *csize, uint32_t *unsat)
{
    if (ncl) LAUNCH(k_dm_clauses, blocks_for(ncl), TPB, s, tok, term, ncl, utime, utime_new, csize, unsat);
}
void be_dm_diff(void *s, const uint32_t *a, const uint32_t *b, uint32_t n, uint32_t *changed) { if (n) LAUNCH(k_dm_diff, blocks_for(n), TPB, s, a, b, n, changed); }
void be_dm_sizes(void *s, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, const uint32_t *csize, uint32_t *words,
                 uint32_t *kept, uint32_t *isunit)
{
    if (ncl) LAUNCH(k_dm_sizes, blocks_for(ncl), TPB, s, tok, term, ncl, utime, csize, words, kept, isunit);
}
void be_dm_write(void *s, const uint32_t *tok, const uint32_t *term, uint32_t ncl, const uint32_t *utime, const uint32_t *csize, const uint32_t *wordoff,
                 const uint32_t *keptoff, const uint32_t *unitoff, const uint32_t *isunit, uint32_t *cm, unsigned long long *orgs, uint32_t *trail,
                 int8_t *value, uint8_t *state, unsigned long long *stats, uint32_t header_clauses, uint32_t *bad)
{
    if (ncl) LAUNCH(k_dm_write, blocks_for(ncl), TPB, s, tok, term, ncl, utime, csize, wordoff, keptoff, unitoff, isunit, cm, orgs, trail, value, state, stats,
                    header_clauses, bad);
}

/* ---- rebuildWT launches ------------------------------------------------------------------------------ */
void be_wt_groups(void *s, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t code, uint32_t *f0, uint32_t *f1, uint32_t *f2, uint32_t *f3)
{
    if (n) LAUNCH(k_wt_groups, blocks_for(n), TPB, s, arena, (const unsigned long long *)refs, n, code, f0, f1, f2, f3);
}
void be_wt_fill(void *s, uint32_t *cm, const uint32_t *arena, const uint64_t *refs, uint32_t n, uint32_t code, const uint32_t *p0, const uint32_t *p1,
                const uint32_t *p2, const uint32_t *p3, const uint32_t *totals, const int8_t *value, sg_u4 *hdr, uint32_t *lits, sg_u4 *winfo)
{
    if (n) LAUNCH(k_wt_fill, blocks_for(n), TPB, s, cm, arena, (const unsigned long long *)refs, n, code, p0, p1, p2, p3, totals, value, (uint4 *)hdr, lits, (uint4 *)winfo);
}
void be_wt_records(void *s, uint32_t nlit, const uint32_t *ot_off, const uint32_t *ot_len, const uint32_t *ot_data, const sg_u4 *winfo, sg_u4 *rec)
{
    if (nlit) LAUNCH(k_wt_records, blocks_for(nlit), TPB, s, nlit, ot_off, ot_len, ot_data, (const uint4 *)winfo, (uint4 *)rec);
}

int be_ere_replay_block(void *s, const SgView &v, uint32_t n_events, uint32_t n_list, uint32_t *claims, uint32_t *ccnt)
{
    LAUNCH(k_ere_replay_block, 1, 1024, s, v, n_events, n_list, claims, ccnt);
    return 0;
}

/* sortOT: the lists the thread-per-list kernel put on the work list (v.sot_wl, count in sc->wl_count[1]), one warp each */
int be_sortot_wide(void *s, const SgView &v)
{
    if (!v.sot_wl) return 1;
    LAUNCH(k_sortot_wide, 148 * 4, 128, s, v);
    return 0;
}
int be_has_sortot_wide(void) { return 1; }
