/*
 * x3_oracle.h -- CPU restatement of the x3 forward-window match search.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA library, the
 * backend shim, the x3 host program) may include, link or execute this file.
 * It is the checker used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status: PINNED.  The reference holds no golden vectors for this path
 * (SURVEY.md section 4), so the oracle is pinned against the reference itself:
 * oracle/Makefile compiles the unmodified /root/reference sources into
 * oracle/_ref/ and tests/test_oracle.py checks every function below
 * against the compiled reference find_best_match() (reference backend.c:56-100),
 * with an empty and with a live dictionary.
 *
 * Every function cites the reference lines it restates.
 */
#ifndef X3_ORACLE_H
#define X3_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X3O_MAX_MATCH_LEN 32 /* reference backend.h:7-10 */

/* Dictionary queries used by the filter (reference dict.c:105-130, dict.c:159-162).
 * find returns (size_t)-1 when no dictionary string is a prefix of q. */
typedef size_t (*x3o_dict_find_fn)(const char *q);
typedef size_t (*x3o_dict_len_fn)(size_t index);

/* Histogram loop, reference backend.c:58-74, literal restatement.
 * count[i] = #{ s in [p+1, p+W-33] : x[p..p+i] == x[s..s+i] }, exact. */
void x3o_histogram(const uint8_t *p, size_t W, size_t count[X3O_MAX_MATCH_LEN]);

/* Selection + dictionary filter, reference backend.c:76-99, literal restatement
 * (same loop nest, same casts, same double evaluation of the dictionary query). */
size_t x3o_select(const size_t count[X3O_MAX_MATCH_LEN], const uint8_t *p, int t,
                  size_t f1, size_t f2, x3o_dict_find_fn find, x3o_dict_len_fn len);

/* Whole function, reference backend.c:56-100. */
size_t x3o_find_best_match(const uint8_t *p, size_t W, int t, size_t f1, size_t f2,
                           x3o_dict_find_fn find, x3o_dict_len_fn len);

/* The collapsed selection the GPU epilogue implements:
 * 0 when t <= 0 or count[0] < 2, else #{ i : count[i] > min(t, count[0]-1) }.
 * `count` may be saturated at any cap >= t+1. */
uint8_t x3o_lstar_from_count(const size_t count[X3O_MAX_MATCH_LEN], int t);

/* Host-side remainder of the split (SURVEY.md 8(a) rows a2+a3): given L* for p,
 * scan i = L*-1 .. 0 applying the filter of reference backend.c:79-90. */
size_t x3o_filter_from_lstar(uint8_t lstar, const uint8_t *p, size_t f1, size_t f2,
                             x3o_dict_find_fn find, x3o_dict_len_fn len);

/* Table builders over positions [p0, p1) of the padded buffer x (N data bytes
 * followed by >= W zero bytes, reference x3.c:579,590).  Row-major 32 cells per
 * position.  "plain" walks the reference loop per position; "fast" uses the
 * diagonal run-length identity LCP(p,p+d) = 1 + LCP(p+1,p+1+d) and OpenMP; tests
 * check they agree.  H8 saturates at 255, H16 at 65535.  Any pointer may be NULL. */
void x3o_table_plain(const uint8_t *x, size_t p0, size_t p1, size_t W, int t,
                     uint8_t *H8, uint16_t *H16, uint8_t *lstar);
void x3o_table_fast(const uint8_t *x, size_t xlen, size_t p0, size_t p1, size_t W, int t,
                    uint8_t *H8, uint16_t *H16, uint8_t *lstar);

/* Times fn(base + i) for i in [i0, i1) step `stride`; returns the sum of results
 * (so the calls cannot be elided).  Used to time the compiled reference
 * find_best_match from several threads in bench.py. */
typedef size_t (*x3o_fbm_fn)(char *p);
uint64_t x3o_call_range(x3o_fbm_fn fn, char *base, size_t i0, size_t i1, size_t stride);

#ifdef __cplusplus
}
#endif
#endif
