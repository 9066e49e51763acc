/* sigma_dimacs.cuh - the reference's DIMACS parser (dimacs.cpp:25-177, dimacs.hpp:57-74) as data-parallel steps.
 *
 * The reference walks the file once: eatWS; 'c' at a statement start skips the line; "p cnf V C" is the header; anything
 * else is a clause = integers up to a 0 (toInteger accepts [+-]?digits and stops at the first non-digit, so "1-2" is two
 * integers), handed to makeClause: repeated literals dropped, a tautology or a literal already true at the root drops the
 * clause, literals already false are removed, one literal left = a root unit (enqueueUnit), the rest becomes a CLAUSE in
 * the arena, in file order.  Two things make that sequential: (1) what a byte means depends on where the statement
 * started (inside a comment? inside a clause?), (2) units found while parsing change the clauses that FOLLOW them.
 *
 * (1) is a finite-state machine over the bytes.  Every 64-byte chunk is first reduced to its transition function
 * (8 states -> 8 states, 3 bits each, one 24-bit word); function composition is associative, so an exclusive scan of
 * the chunk functions gives every chunk its start state, after which the chunks are independent:
 *      states  B  between statements          K  inside a clause (a non-zero literal since the last 0)
 *              C  inside a comment            E  after the end marker ('%' or NUL at a statement start, dimacs.cpp:54)
 *              SGN just read a sign           TZ inside an integer, all digits 0 so far      TN inside an integer, non-zero
 *              ERR the reference stops with "expected a digit" / "header has wrong format"
 * (2): `utime[lit]` = index of the clause that made `lit` true (a unit), and clause c sees exactly the units of earlier
 * clauses.  Evaluated for all clauses at once from the previous round's utime and repeated until utime is stable: a
 * round fixes at least the first clause that was still wrong (it only looks at earlier, already right ones), and a file
 * without units needs one round.
 *
 * Everything here is host + device code: the CUDA kernels (sigma_kernels.cu) and the serial test stand-in
 * (tests/emul) run the same bodies. */
#ifndef SIGMA_DIMACS_CUH
#define SIGMA_DIMACS_CUH

#include "sigma_types.h"

enum { SGD_B = 0, SGD_K, SGD_C, SGD_E, SGD_SGN, SGD_TZ, SGD_TN, SGD_ERR };
enum { SGD_WS = 0, SGD_NL, SGD_Z, SGD_D, SGD_SG, SGD_CC, SGD_PP, SGD_PCT, SGD_NUL, SGD_OTH, SGD_NCLASS };

#define SGD_CHUNK 64u                      /* bytes per work item */
#define SGD_IDENT 0xFAC688u                /* the identity function: state s -> s, 3 bits per state */
#define SGD_NEVER 0xFFFFFFFFu              /* utime: the literal is never made true by the file */

SG_HD uint32_t sgd_class(uint8_t ch)
{
    if (ch == '\n') return SGD_NL;
    if ((ch >= 9 && ch <= 13) || ch == 32) return SGD_WS;          /* eatWS, dimacs.hpp:59 */
    if (ch == '0') return SGD_Z;
    if (ch >= '1' && ch <= '9') return SGD_D;
    if (ch == '-' || ch == '+') return SGD_SG;                     /* toInteger, dimacs.hpp:67-68 */
    if (ch == 'c') return SGD_CC;
    if (ch == 'p') return SGD_PP;
    if (ch == '%') return SGD_PCT;
    if (ch == 0) return SGD_NUL;
    return SGD_OTH;
}

/* transition function of one byte class, packed: bits [3s, 3s+3) = next state from state s */
#define SGD_PACK8(b, k, c, e, sg, tz, tn, er) \
    ((uint32_t)(b) | ((uint32_t)(k) << 3) | ((uint32_t)(c) << 6) | ((uint32_t)(e) << 9) | ((uint32_t)(sg) << 12) | ((uint32_t)(tz) << 15) | \
     ((uint32_t)(tn) << 18) | ((uint32_t)(er) << 21))
SG_HD uint32_t sgd_class_fn(uint32_t cls)
{
    switch (cls) {                     /*        from: B        K        C      E      SGN      TZ       TN       ERR */
    case SGD_WS:  return SGD_PACK8(SGD_B,   SGD_K,   SGD_C, SGD_E, SGD_ERR, SGD_B,   SGD_K,   SGD_ERR);
    case SGD_NL:  return SGD_PACK8(SGD_B,   SGD_K,   SGD_B, SGD_E, SGD_ERR, SGD_B,   SGD_K,   SGD_ERR);
    case SGD_Z:   return SGD_PACK8(SGD_TZ,  SGD_TZ,  SGD_C, SGD_E, SGD_TZ,  SGD_TZ,  SGD_TN,  SGD_ERR);
    case SGD_D:   return SGD_PACK8(SGD_TN,  SGD_TN,  SGD_C, SGD_E, SGD_TN,  SGD_TN,  SGD_TN,  SGD_ERR);
    case SGD_SG:  return SGD_PACK8(SGD_SGN, SGD_SGN, SGD_C, SGD_E, SGD_ERR, SGD_SGN, SGD_SGN, SGD_ERR);
    case SGD_CC:  return SGD_PACK8(SGD_C,   SGD_ERR, SGD_C, SGD_E, SGD_ERR, SGD_C,   SGD_ERR, SGD_ERR);
    case SGD_PP:  return SGD_PACK8(SGD_ERR, SGD_ERR, SGD_C, SGD_E, SGD_ERR, SGD_ERR, SGD_ERR, SGD_ERR);   /* a second header: not supported */
    case SGD_PCT: return SGD_PACK8(SGD_E,   SGD_ERR, SGD_C, SGD_E, SGD_ERR, SGD_E,   SGD_ERR, SGD_ERR);
    case SGD_NUL: return SGD_PACK8(SGD_E,   SGD_ERR, SGD_E, SGD_E, SGD_ERR, SGD_E,   SGD_ERR, SGD_ERR);
    default:      return SGD_PACK8(SGD_ERR, SGD_ERR, SGD_C, SGD_E, SGD_ERR, SGD_ERR, SGD_ERR, SGD_ERR);
    }
}
SG_HD uint32_t sgd_apply(uint32_t fn, uint32_t state) { return (fn >> (3u * state)) & 7u; }
/* first f, then g */
SG_HD uint32_t sgd_compose(uint32_t f, uint32_t g)
{
    uint32_t r = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t s = 0; s < 8; ++s) r |= sgd_apply(g, sgd_apply(f, s)) << (3u * s);
    return r;
}
SG_HD uint32_t sgd_chunk_fn(const uint8_t *text, uint64_t lo, uint64_t hi)
{
    uint32_t f = SGD_IDENT;
    for (uint64_t i = lo; i < hi; ++i) f = sgd_compose(f, sgd_class_fn(sgd_class(text[i])));
    return f;
}
SG_HD bool sgd_in_token(uint32_t state) { return state == SGD_SGN || state == SGD_TZ || state == SGD_TN; }
/* does the byte of class `cls`, read in state `prev`, start an integer? */
SG_HD bool sgd_token_start(uint32_t prev, uint32_t cls)
{
    if (prev == SGD_C || prev == SGD_E || prev == SGD_ERR) return false;
    if (cls == SGD_SG) return prev != SGD_SGN;
    if (cls == SGD_Z || cls == SGD_D) return !sgd_in_token(prev);
    return false;
}
/* the integer that starts at text[i] (toInteger: sign, digits, uint32 arithmetic that wraps); may run past the chunk */
SG_HD uint32_t sgd_read_int(const uint8_t *text, uint64_t i, uint64_t n, uint32_t *sign)
{
    *sign = 0;
    if (text[i] == '-') { *sign = 1; ++i; }
    else if (text[i] == '+') ++i;
    uint32_t v = 0;
    while (i < n && text[i] >= '0' && text[i] <= '9') v = v * 10u + (uint32_t)(text[i++] - '0');
    return v;
}
/* one pass over a chunk from its start state: the integers that START in it.  emit == false: counts only.
 * tok[k] = 0 for a terminator, 2 v + sign otherwise (V2DEC); term[j] = index of the j-th terminator among all tokens.
 * Returns the state after the chunk; *bad is raised for a variable beyond nvars ("too many variables"). */
SG_HD uint32_t sgd_chunk_tokens(const uint8_t *text, uint64_t lo, uint64_t hi, uint64_t n, uint32_t state, uint32_t nvars, bool emit,
                                uint32_t *ntok, uint32_t *nterm, uint32_t tok_base, uint32_t term_base, uint32_t *tok, uint32_t *term, uint32_t *bad)
{
    uint32_t nt = 0, nz = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t cls = sgd_class(text[i]);
        if (sgd_token_start(state, cls)) {
            uint32_t sign;
            uint32_t v = sgd_read_int(text, i, n, &sign);
            if (emit) {
                if (v > nvars) { *bad = 3; v = 1; }                 /* the parse fails; keep the literal inside the tables */
                tok[tok_base + nt] = v ? 2u * v + sign : 0u;
                if (!v) term[term_base + nz] = tok_base + nt;
            }
            ++nt;
            nz += v == 0;
        }
        state = sgd_apply(sgd_class_fn(cls), state);
    }
    *ntok = nt; *nterm = nz;
    return state;
}

/* makeClause (dimacs.cpp:121-177) for clause `c` = tok[first, last) under the units of EARLIER clauses.
 * Returns the number of literals left (0 also for a dropped clause); *kind: 0 dropped (tautology, satisfied, or all
 * literals false), 1 unit, 2 clause, 3 the empty clause (parser() returns false).  out (may be null): the literals left. */
SG_HD uint32_t sgd_make_clause(const uint32_t *tok, uint32_t first, uint32_t last, uint32_t c, const uint32_t *utime, uint32_t *out, uint32_t *kind)
{
    if (first == last) { *kind = 3; return 0; }
    uint32_t n = 0;
    bool satisfied = false;
    for (uint32_t i = first; i < last; ++i) {
        const uint32_t lit = tok[i];
        bool seen = false;
        for (uint32_t j = first; j < i; ++j) {
            if (tok[j] == lit) { seen = true; break; }                    /* marker == sign: ignored */
            if (tok[j] == (lit ^ 1u)) { seen = true; satisfied = true; break; }   /* tautology */
        }
        if (seen) continue;
        if (utime[lit] < c) satisfied = true;                             /* true at the root by now */
        else if (utime[lit ^ 1u] < c) { }                                  /* false by now: removed */
        else { if (out) out[n] = lit; ++n; }
    }
    if (satisfied) { *kind = 0; return 0; }
    *kind = n == 0 ? 0u : n == 1 ? 1u : 2u;
    return n;
}
/* is the clause dropped because it is satisfied / a tautology (true) or because every literal is false (false)? */
SG_HD bool sgd_satisfied_or_empty(const uint32_t *tok, uint32_t first, uint32_t last, uint32_t c, const uint32_t *utime)
{
    if (first == last) return true;
    for (uint32_t i = first; i < last; ++i) {
        const uint32_t lit = tok[i];
        bool seen = false;
        for (uint32_t j = first; j < i; ++j) {
            if (tok[j] == lit) { seen = true; break; }
            if (tok[j] == (lit ^ 1u)) return true;
        }
        if (!seen && utime[lit] < c) return true;
    }
    return false;
}
/* the only literal left of a unit clause */
SG_HD uint32_t sgd_unit_literal(const uint32_t *tok, uint32_t first, uint32_t last, uint32_t c, const uint32_t *utime)
{
    for (uint32_t i = first; i < last; ++i) {
        const uint32_t lit = tok[i];
        bool seen = false;
        for (uint32_t j = first; j < i; ++j) if ((tok[j] >> 1) == (lit >> 1)) { seen = true; break; }
        if (!seen && !(utime[lit] < c) && !(utime[lit ^ 1u] < c)) return lit;
    }
    return 0;
}
/* CLAUSE record of a parsed original (clause.hpp:47-59, CLAUSE(const Lits_t&) :91-106): 4 header words + literals */
SG_HD void sgd_write_record(uint32_t *dst, const uint32_t *lits, uint32_t n)
{
    dst[0] = (1u << 4) | (n == 2 ? 1u << 5 : 0u);          /* _k, _b */
    dst[1] = n;
    dst[2] = 2u;                                            /* _pos */
    dst[3] = 0u;                                            /* _lbd */
    for (uint32_t k = 0; k < n; ++k) dst[4 + k] = lits[k];
}

#endif
