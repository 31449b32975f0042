/*
 * x3_oracle.c -- CPU restatement of the x3 forward-window match search.
 *
 * TEST INFRASTRUCTURE ONLY (see x3_oracle.h).  Parity status: PINNED against the
 * compiled, unmodified reference (oracle/Makefile -> oracle/_ref/libx3ref.so,
 * checked by tests/test_oracle.py).
 */
#include "x3_oracle.h"

#include <stdlib.h>
#include <string.h>

/* reference backend.c:58-74 */
void x3o_histogram(const uint8_t *p, size_t W, size_t count[X3O_MAX_MATCH_LEN])
{
	for (int i = 0; i < X3O_MAX_MATCH_LEN; ++i) {
		count[i] = 0;
	}

	/* reference: for (s = p + 1; s < end - MAX_MATCH_LEN; ++s), end = p + W.
	 * Written with offsets so that W < 33 cannot wrap. */
	for (size_t d = 1; d + X3O_MAX_MATCH_LEN < W; ++d) {
		const uint8_t *s = p + d;
		for (int i = 0; i < X3O_MAX_MATCH_LEN; ++i) {
			if (p[i] == s[i]) {
				count[i]++;
			} else {
				break;
			}
		}
	}
}

/* reference backend.c:76-99 */
size_t x3o_select(const size_t count[X3O_MAX_MATCH_LEN], const uint8_t *p, int t,
                  size_t f1, size_t f2, x3o_dict_find_fn find, x3o_dict_len_fn len)
{
	for (int tc = t; tc > 0; --tc) {
		for (int i = X3O_MAX_MATCH_LEN - 1; i >= 0; --i) {
			if (count[i] > (size_t)tc) {
				int skip = 0;
				/* backend.c:79-83 */
				if (i >= 2 && f1 > 0) {
					if (find((const char *)p + i) != (size_t)-1 &&
					    len(find((const char *)p + i)) * f1 > (size_t)(i + 1)) {
						skip = 1;
					}
				}
				/* backend.c:84-90, int arithmetic as in the reference */
				if (!skip && i >= 1 && f2 > 0) {
					for (int o = 1; o <= i; ++o) {
						if (find((const char *)p + o) != (size_t)-1 &&
						    ((int)len(find((const char *)p + o)) - o) * (int)f2 > i + 1) {
							skip = 1;
							break;
						}
					}
				}
				if (!skip) {
					return (size_t)i + 1; /* backend.c:92 */
				}
			}
		}
	}
	return 1; /* backend.c:99 */
}

/* reference backend.c:56-100 */
size_t x3o_find_best_match(const uint8_t *p, size_t W, int t, size_t f1, size_t f2,
                           x3o_dict_find_fn find, x3o_dict_len_fn len)
{
	size_t count[X3O_MAX_MATCH_LEN];
	x3o_histogram(p, W, count);
	return x3o_select(count, p, t, f1, f2, find, len);
}

/* what the loop nest of backend.c:76-78 reduces to when the
 * filter is taken out.  count is non-increasing in i and i == 0 is never
 * filtered, so the first tc with any count[i] > tc is min(t, count[0]-1) and the
 * loop returns inside that tc. */
uint8_t x3o_lstar_from_count(const size_t count[X3O_MAX_MATCH_LEN], int t)
{
	if (t <= 0 || count[0] < 2) {
		return 0;
	}
	size_t tcs = count[0] - 1;
	if (tcs > (size_t)t) {
		tcs = (size_t)t;
	}
	uint8_t n = 0;
	for (int i = 0; i < X3O_MAX_MATCH_LEN; ++i) {
		if (count[i] > tcs) {
			n++;
		}
	}
	return n;
}

/* backend.c:79-92 applied to i = L*-1 .. 0 */
size_t x3o_filter_from_lstar(uint8_t lstar, const uint8_t *p, size_t f1, size_t f2,
                             x3o_dict_find_fn find, x3o_dict_len_fn len)
{
	for (int i = (int)lstar - 1; i >= 0; --i) {
		int skip = 0;
		if (i >= 2 && f1 > 0) {
			size_t m = find((const char *)p + i);
			if (m != (size_t)-1 && len(m) * f1 > (size_t)(i + 1)) {
				skip = 1;
			}
		}
		if (!skip && i >= 1 && f2 > 0) {
			for (int o = 1; o <= i; ++o) {
				size_t m = find((const char *)p + o);
				if (m != (size_t)-1 && ((int)len(m) - o) * (int)f2 > i + 1) {
					skip = 1;
					break;
				}
			}
		}
		if (!skip) {
			return (size_t)i + 1;
		}
	}
	return 1;
}

static void store_row(const size_t count[X3O_MAX_MATCH_LEN], size_t row, int t,
                      uint8_t *H8, uint16_t *H16, uint8_t *lstar)
{
	if (H8 != NULL) {
		for (int i = 0; i < X3O_MAX_MATCH_LEN; ++i) {
			H8[row * X3O_MAX_MATCH_LEN + i] = count[i] > 255 ? 255 : (uint8_t)count[i];
		}
	}
	if (H16 != NULL) {
		for (int i = 0; i < X3O_MAX_MATCH_LEN; ++i) {
			H16[row * X3O_MAX_MATCH_LEN + i] = count[i] > 65535 ? 65535 : (uint16_t)count[i];
		}
	}
	if (lstar != NULL) {
		lstar[row] = x3o_lstar_from_count(count, t);
	}
}

void x3o_table_plain(const uint8_t *x, size_t p0, size_t p1, size_t W, int t,
                     uint8_t *H8, uint16_t *H16, uint8_t *lstar)
{
	for (size_t p = p0; p < p1; ++p) {
		size_t count[X3O_MAX_MATCH_LEN];
		x3o_histogram(x + p, W, count);
		store_row(count, p - p0, t, H8, H16, lstar);
	}
}

/*
 * Same table through the diagonal identity: for a fixed distance d,
 * LCP32(p, p+d) = x[p]==x[p+d] ? min(32, 1 + LCP32(p+1, p+1+d)) : 0, so one
 * backward sweep per d yields every position's prefix length.  Per block of
 * positions a histogram over prefix lengths is kept and suffix-summed into
 * count[i] = #{d : LCP >= i+1} (the quantity backend.c:66-74 accumulates).
 */
void x3o_table_fast(const uint8_t *x, size_t xlen, size_t p0, size_t p1, size_t W, int t,
                    uint8_t *H8, uint16_t *H16, uint8_t *lstar)
{
	const size_t D = W > X3O_MAX_MATCH_LEN + 1 ? W - X3O_MAX_MATCH_LEN - 1 : 0;
	const size_t BLK = 2048;
	const long nblk = (long)((p1 - p0 + BLK - 1) / BLK);
	(void)xlen;

#pragma omp parallel
	{
		uint32_t *hist = malloc(BLK * (X3O_MAX_MATCH_LEN + 1) * sizeof(uint32_t));
#pragma omp for schedule(dynamic, 1)
		for (long b = 0; b < nblk; ++b) {
			size_t lo = p0 + (size_t)b * BLK;
			size_t hi = lo + BLK < p1 ? lo + BLK : p1; /* positions [lo, hi) */
			memset(hist, 0, BLK * (X3O_MAX_MATCH_LEN + 1) * sizeof(uint32_t));
			for (size_t d = 1; d <= D; ++d) {
				uint32_t rl = 0;
				/* warm-up over the 31 positions past the block */
				for (size_t p = hi + X3O_MAX_MATCH_LEN - 2; p >= hi; --p) {
					rl = x[p] == x[p + d] ? rl + 1 : 0;
				}
				for (size_t p = hi; p-- > lo;) {
					rl = x[p] == x[p + d] ? rl + 1 : 0;
					uint32_t c = rl > X3O_MAX_MATCH_LEN ? X3O_MAX_MATCH_LEN : rl;
					hist[(p - lo) * (X3O_MAX_MATCH_LEN + 1) + c]++;
				}
			}
			for (size_t p = lo; p < hi; ++p) {
				size_t count[X3O_MAX_MATCH_LEN];
				size_t acc = 0;
				const uint32_t *h = hist + (p - lo) * (X3O_MAX_MATCH_LEN + 1);
				for (int i = X3O_MAX_MATCH_LEN - 1; i >= 0; --i) {
					acc += h[i + 1];
					count[i] = acc;
				}
				store_row(count, p - p0, t, H8, H16, lstar);
			}
		}
		free(hist);
	}
}

uint64_t x3o_call_range(x3o_fbm_fn fn, char *base, size_t i0, size_t i1, size_t stride)
{
	uint64_t acc = 0;
	if (stride == 0) {
		stride = 1;
	}
	for (size_t i = i0; i < i1; i += stride) {
		acc += (uint64_t)fn(base + i);
	}
	return acc;
}
