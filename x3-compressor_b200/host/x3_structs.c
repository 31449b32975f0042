/*
 * x3_structs.c -- bit I/O, arithmetic coder, adaptive models, contexts, tag-pair
 * map and dictionary of the x3 host pass (see x3_host.h).
 *
 * Observable behaviour is the reference's (files cited per function); the data
 * structures are not: Fenwick trees instead of full cumulative-frequency
 * recomputation (ac.c:5-18,215-228), hash maps instead of linear scans and an
 * unbalanced BST (context.c:20-40, tag_pair.c:67-130), a trie instead of a
 * memcmp over every dictionary element (dict.c:105-157), use stamps + an
 * order-statistic tree instead of a qsort per step (dict.c:132-146).
 */
#define _POSIX_C_SOURCE 200112L
#include "x3_host.h"

#include <stdlib.h>
#include <string.h>

static void *xmalloc(size_t n)
{
	void *p = malloc(n ? n : 1);
	if (p == NULL) {
		abort(); /* reference convention: dict.c:33-35, x3.c:582-588 */
	}
	return p;
}

static void *xcalloc(size_t n, size_t m)
{
	void *p = calloc(n ? n : 1, m ? m : 1);
	if (p == NULL) {
		abort();
	}
	return p;
}

static void *xrealloc(void *q, size_t n)
{
	void *p = realloc(q, n ? n : 1);
	if (p == NULL) {
		abort();
	}
	return p;
}

static uint64_t mix64(uint64_t x)
{
	x ^= x >> 30;
	x *= 0xbf58476d1ce4e5b9ULL;
	x ^= x >> 27;
	x *= 0x94d049bb133111ebULL;
	x ^= x >> 31;
	return x;
}

/* ===================================================================================== */
/* bit I/O (reference bio.c)                                                              */
/* ===================================================================================== */

void x3_bitw_open(struct x3_bitw *w, size_t reserve_bytes)
{
	w->cap = reserve_bytes / 4 + 16;
	w->buf = xmalloc(w->cap * sizeof(uint32_t));
	w->n = 0;
	w->acc = 0;
	w->nacc = 0;
}

static void bitw_flush_word(struct x3_bitw *w)
{
	if (w->n == w->cap) {
		w->cap *= 2;
		w->buf = xrealloc(w->buf, w->cap * sizeof(uint32_t));
	}
	w->buf[w->n++] = (uint32_t)w->acc; /* native-endian word store, bio.c:25 */
	w->acc >>= 32;
	w->nacc -= 32;
}

void x3_bitw_put(struct x3_bitw *w, unsigned bit)
{
	w->acc |= (uint64_t)(bit & 1u) << w->nacc; /* LSB first, bio.c:58 */
	if (++w->nacc >= 32) {
		bitw_flush_word(w);
	}
}

void x3_bitw_run(struct x3_bitw *w, unsigned bit, uint64_t count)
{
	while (count > 0) {
		unsigned k = 32 - w->nacc;
		if ((uint64_t)k > count) {
			k = (unsigned)count;
		}
		if (bit & 1u) {
			w->acc |= ((((uint64_t)1 << k) - 1u)) << w->nacc;
		}
		w->nacc += k;
		count -= k;
		if (w->nacc >= 32) {
			bitw_flush_word(w);
		}
	}
}

size_t x3_bitw_close(struct x3_bitw *w)
{
	if (w->nacc > 0) { /* bio_close, bio.c:105-112 */
		w->nacc = 32;
		bitw_flush_word(w);
		w->acc = 0;
		w->nacc = 0;
	}
	return w->n * 4;
}

void x3_bitr_open(struct x3_bitr *r, const void *buf, size_t bytes)
{
	/* bio_open: end = buf + bytes - 3; a word is read while ptr < end (bio.c:10,35) */
	r->ptr = (const uint32_t *)buf;
	r->end = r->ptr + (bytes >= 3 ? bytes / 4 : 0);
	r->b = 0;
	r->c = 32;
}

unsigned x3_bitr_get(struct x3_bitr *r)
{
	if (r->c == 32) {
		if (r->ptr < r->end) {
			uint32_t v;
			memcpy(&v, r->ptr, 4);
			r->ptr++;
			r->b = v;
		} else {
			r->b = 0x80000000u; /* bio.c:38 */
		}
		r->c = 0;
	}
	unsigned bit = r->b & 1u;
	r->b >>= 1;
	r->c++;
	return bit;
}

/* ===================================================================================== */
/* arithmetic coder (reference ac.c:31-198)                                               */
/* ===================================================================================== */

#define AC_Q1 0x20000000ULL
#define AC_HALF 0x40000000ULL
#define AC_Q3 0x60000000ULL

void x3_ac_init(struct x3_ac *ac)
{
	ac->low = 0;
	ac->high = 0x7FFFFFFFULL;
	ac->scale = 0;
	ac->buffer = 0;
}

void x3_ac_encode(struct x3_ac *ac, struct x3_bitw *w, uint64_t low_freq, uint64_t high_freq, uint64_t total)
{
	/* ac_encode, ac.c:77-85 */
	const uint64_t step = (ac->high - ac->low + 1) / total;
	ac->high = ac->low + step * high_freq - 1;
	ac->low = ac->low + step * low_freq;
	/* ac_encode_scale, ac.c:46-75 */
	while (ac->high < AC_HALF || ac->low >= AC_HALF) {
		if (ac->high < AC_HALF) {
			x3_bitw_put(w, 0);
			ac->low = 2 * ac->low;
			ac->high = 2 * ac->high + 1;
			x3_bitw_run(w, 1, ac->scale);
		} else {
			x3_bitw_put(w, 1);
			ac->low = 2 * (ac->low - AC_HALF);
			ac->high = 2 * (ac->high - AC_HALF) + 1;
			x3_bitw_run(w, 0, ac->scale);
		}
		ac->scale = 0;
	}
	while (AC_Q1 <= ac->low && ac->high < AC_Q3) {
		ac->scale++;
		ac->low = 2 * (ac->low - AC_Q1);
		ac->high = 2 * (ac->high - AC_Q1) + 1;
	}
}

void x3_ac_encode_flush(struct x3_ac *ac, struct x3_bitw *w)
{
	/* ac.c:115-126 */
	if (ac->low < AC_Q1) {
		x3_bitw_put(w, 0);
		x3_bitw_run(w, 1, ac->scale + 1);
	} else {
		x3_bitw_put(w, 1);
	}
}

void x3_ac_decode_init(struct x3_ac *ac, struct x3_bitr *r)
{
	ac->buffer = 0; /* ac.c:133-140 */
	for (int i = 0; i < 31; i++) {
		ac->buffer = (ac->buffer << 1) | x3_bitr_get(r);
	}
}

uint64_t x3_ac_decode_target(struct x3_ac *ac, uint64_t total, uint64_t *step)
{
	*step = (ac->high - ac->low + 1) / total; /* ac.c:174-176 */
	return (ac->buffer - ac->low) / *step;
}

void x3_ac_decode_update(struct x3_ac *ac, struct x3_bitr *r, uint64_t step, uint64_t low_freq, uint64_t high_freq)
{
	ac->high = ac->low + step * high_freq - 1; /* ac.c:183-186 */
	ac->low = ac->low + step * low_freq;
	/* ac_decode_scale, ac.c:142-165 */
	while (ac->high < AC_HALF || ac->low >= AC_HALF) {
		if (ac->high < AC_HALF) {
			ac->low = 2 * ac->low;
			ac->high = 2 * ac->high + 1;
			ac->buffer = 2 * ac->buffer + x3_bitr_get(r);
		} else {
			ac->low = 2 * (ac->low - AC_HALF);
			ac->high = 2 * (ac->high - AC_HALF) + 1;
			ac->buffer = 2 * (ac->buffer - AC_HALF) + x3_bitr_get(r);
		}
		ac->scale = 0;
	}
	while (AC_Q1 <= ac->low && ac->high < AC_Q3) {
		ac->scale++;
		ac->low = 2 * (ac->low - AC_Q1);
		ac->high = 2 * (ac->high - AC_Q1) + 1;
		ac->buffer = 2 * (ac->buffer - AC_Q1) + x3_bitr_get(r);
	}
}

/* ===================================================================================== */
/* Fenwick helpers (1-based tree over a 0-based value array)                              */
/* ===================================================================================== */

static void fen_build(uint64_t *tree, uint32_t cap, const uint32_t *val, uint32_t n)
{
	memset(tree, 0, ((size_t)cap + 1) * sizeof(uint64_t));
	for (uint32_t i = 1; i <= cap; ++i) {
		if (i <= n) {
			tree[i] += val[i - 1];
		}
		const uint32_t j = i + (i & (0u - i));
		if (j <= cap) {
			tree[j] += tree[i];
		}
	}
}

static inline void fen_add(uint64_t *tree, uint32_t cap, uint32_t i0, uint64_t delta)
{
	for (uint32_t i = i0 + 1; i <= cap; i += i & (0u - i)) {
		tree[i] += delta;
	}
}

/* sum of values [0, i) */
static inline uint64_t fen_prefix(const uint64_t *tree, uint32_t i)
{
	uint64_t s = 0;
	for (; i > 0; i &= i - 1) {
		s += tree[i];
	}
	return s;
}

/* largest pos with prefix(pos) <= value; *cum = prefix(pos).  cap is a power of two. */
static inline uint32_t fen_find(const uint64_t *tree, uint32_t cap, uint64_t value, uint64_t *cum)
{
	uint32_t pos = 0;
	uint64_t acc = 0;
	for (uint32_t step = cap; step > 0; step >>= 1) {
		const uint32_t nx = pos + step;
		if (nx <= cap && acc + tree[nx] <= value) {
			pos = nx;
			acc += tree[nx];
		}
	}
	*cum = acc;
	return pos;
}

static uint32_t pow2_at_least(uint32_t n)
{
	uint32_t c = 1;
	while (c < n) {
		c <<= 1;
	}
	return c;
}

/* ===================================================================================== */
/* adaptive model (reference ac.c:200-273)                                                */
/* ===================================================================================== */

void x3_model_create(struct x3_model *m, uint32_t n)
{
	m->n = n;
	m->cap = pow2_at_least(n < 4 ? 4 : n);
	m->freq = xmalloc((size_t)m->cap * sizeof(uint32_t));
	m->tree = xmalloc(((size_t)m->cap + 1) * sizeof(uint64_t));
	for (uint32_t i = 0; i < n; ++i) {
		m->freq[i] = 1; /* ac.c:241-244 */
	}
	fen_build(m->tree, m->cap, m->freq, n);
	m->total = n;
}

void x3_model_set(struct x3_model *m, uint32_t i, uint32_t freq)
{
	m->total += (uint64_t)freq - m->freq[i];
	m->freq[i] = freq;
	fen_build(m->tree, m->cap, m->freq, m->n);
}

void x3_model_append(struct x3_model *m)
{
	if (m->n == m->cap) {
		m->cap *= 2;
		m->freq = xrealloc(m->freq, (size_t)m->cap * sizeof(uint32_t));
		m->tree = xrealloc(m->tree, ((size_t)m->cap + 1) * sizeof(uint64_t));
		fen_build(m->tree, m->cap, m->freq, m->n);
	}
	m->freq[m->n] = 1; /* ac.c:259-260 */
	fen_add(m->tree, m->cap, m->n, 1);
	m->n++;
	m->total++;
}

void x3_model_inc(struct x3_model *m, uint32_t i)
{
	m->freq[i]++; /* ac.c:225-227; the reference indexes by symbol == position */
	fen_add(m->tree, m->cap, i, 1);
	m->total++;
}

uint64_t x3_model_cum(const struct x3_model *m, uint32_t i)
{
	return fen_prefix(m->tree, i);
}

uint32_t x3_model_find(const struct x3_model *m, uint64_t value)
{
	uint64_t cum;
	const uint32_t i = fen_find(m->tree, m->cap, value, &cum);
	if (i >= m->n) {
		abort(); /* index_of_value, ac.c:169 */
	}
	return i;
}

void x3_model_destroy(struct x3_model *m)
{
	free(m->freq);
	free(m->tree);
	m->freq = NULL;
	m->tree = NULL;
}

/* ===================================================================================== */
/* open-addressing hash map: uint64 key -> uint32 value                                   */
/* ===================================================================================== */

struct hmap {
	uint64_t *key; /* key + 1 stored; 0 = empty */
	uint32_t *val;
	uint64_t mask;
	uint64_t used;
};

static void hmap_init(struct hmap *h, uint64_t cap_pow2)
{
	h->key = xcalloc(cap_pow2, sizeof(uint64_t));
	h->val = xmalloc(cap_pow2 * sizeof(uint32_t));
	h->mask = cap_pow2 - 1;
	h->used = 0;
}

static void hmap_free(struct hmap *h)
{
	free(h->key);
	free(h->val);
}

static inline int64_t hmap_get(const struct hmap *h, uint64_t key)
{
	const uint64_t k1 = key + 1;
	for (uint64_t i = mix64(key) & h->mask;; i = (i + 1) & h->mask) {
		const uint64_t s = h->key[i];
		if (s == k1) {
			return (int64_t)h->val[i];
		}
		if (s == 0) {
			return -1;
		}
	}
}

/* cache hints for a lookup that will happen a few steps from now */
static inline void hmap_prefetch(const struct hmap *h, uint64_t key)
{
	const uint64_t i = mix64(key) & h->mask;
	__builtin_prefetch(&h->key[i], 0, 1);
	__builtin_prefetch(&h->val[i], 0, 1);
}

static void hmap_put_nogrow(struct hmap *h, uint64_t key, uint32_t val)
{
	const uint64_t k1 = key + 1;
	for (uint64_t i = mix64(key) & h->mask;; i = (i + 1) & h->mask) {
		if (h->key[i] == 0) {
			h->key[i] = k1;
			h->val[i] = val;
			h->used++;
			return;
		}
		if (h->key[i] == k1) {
			h->val[i] = val;
			return;
		}
	}
}

static void hmap_put(struct hmap *h, uint64_t key, uint32_t val)
{
	if ((h->used + 1) * 2 > h->mask + 1) {
		struct hmap n;
		hmap_init(&n, (h->mask + 1) * 2);
		for (uint64_t i = 0; i <= h->mask; ++i) {
			if (h->key[i] != 0) {
				hmap_put_nogrow(&n, h->key[i] - 1, h->val[i]);
			}
		}
		hmap_free(h);
		*h = n;
	}
	hmap_put_nogrow(h, key, val);
}

/* ===================================================================================== */
/* contexts (reference context.c)                                                         */
/* ===================================================================================== */

#define CTX_LINEAR 16  /* contexts up to this many items are searched linearly */
#define CTX_FENWICK 64 /* contexts beyond this many items keep a Fenwick tree */

struct x3_ctxset {
	struct x3_ctx *arr;
	uint32_t size;
	struct hmap map; /* (context id, tag) -> item index, for contexts with > CTX_LINEAR items */
};

static struct x3_ctx *ctx_array_alloc(size_t n)
{
	void *p = NULL;
	if (posix_memalign(&p, 64, n * sizeof(struct x3_ctx)) != 0 || p == NULL) {
		abort();
	}
	return p;
}

struct x3_ctxset *x3_ctxset_create(void)
{
	struct x3_ctxset *s = xmalloc(sizeof(*s));
	s->size = 2;
	s->arr = ctx_array_alloc(s->size);
	memset(s->arr, 0, (size_t)s->size * sizeof(struct x3_ctx));
	hmap_init(&s->map, 1024);
	return s;
}

void x3_ctxset_destroy(struct x3_ctxset *s)
{
	for (uint32_t i = 0; i < s->size; ++i) {
		if (s->arr[i].cap) {
			free(s->arr[i].u.big.tag);
			free(s->arr[i].u.big.freq);
			free(s->arr[i].u.big.tree);
		}
	}
	free(s->arr);
	hmap_free(&s->map);
	free(s);
}

struct x3_ctx *x3_ctxset_get(struct x3_ctxset *s, uint32_t id)
{
	if (id >= s->size) {
		uint32_t ns = s->size;
		while (id >= ns) {
			ns *= 2;
		}
		struct x3_ctx *na = ctx_array_alloc(ns);
		memcpy(na, s->arr, (size_t)s->size * sizeof(struct x3_ctx));
		memset(na + s->size, 0, (size_t)(ns - s->size) * sizeof(struct x3_ctx)); /* ctx_enlarge, context.c:7-18 */
		free(s->arr);
		s->arr = na;
		s->size = ns;
	}
	return &s->arr[id];
}

void x3_ctx_prefetch(const struct x3_ctxset *s, uint32_t id, uint32_t tag, int with_items)
{
	if (id >= s->size) {
		return;
	}
	const struct x3_ctx *c = &s->arr[id];
	if (!with_items) {
		__builtin_prefetch(c, 1, 1);
		return;
	}
	/* the context's line is (being) fetched: reach for what find / cum / inc will touch behind it */
	if (c->items > CTX_LINEAR) {
		hmap_prefetch(&s->map, ((uint64_t)id << 32) | tag);
	}
	if (c->cap) {
		__builtin_prefetch(c->u.big.tag, 0, 1);
		__builtin_prefetch(c->u.big.freq, 1, 1);
	}
}

int64_t x3_ctx_find(struct x3_ctxset *s, uint32_t id, uint32_t tag)
{
	const struct x3_ctx *c = x3_ctxset_get(s, id);
	if (c->items <= CTX_LINEAR) {
		const uint32_t *tags = x3_ctx_tags(c);
		for (uint32_t i = 0; i < c->items; ++i) {
			if (tags[i] == tag) {
				return i;
			}
		}
		return -1;
	}
	return hmap_get(&s->map, ((uint64_t)id << 32) | tag);
}

void x3_ctx_add(struct x3_ctxset *s, uint32_t id, uint32_t tag)
{
	struct x3_ctx *c = x3_ctxset_get(s, id);
	if (c->cap == 0 && c->items == X3_CTX_INLINE) {
		/* leave the inline representation */
		const uint32_t ncap = 4 * X3_CTX_INLINE;
		uint32_t *tg = xmalloc((size_t)ncap * sizeof(uint32_t));
		uint32_t *fr = xmalloc((size_t)ncap * sizeof(uint32_t));
		memcpy(tg, c->u.small.tag, sizeof(c->u.small.tag));
		memcpy(fr, c->u.small.freq, sizeof(c->u.small.freq));
		c->u.big.tag = tg;
		c->u.big.freq = fr;
		c->u.big.tree = NULL;
		c->u.big.tree_cap = 0;
		c->cap = ncap;
	} else if (c->cap != 0 && c->items == c->cap) {
		c->cap *= 2;
		c->u.big.tag = xrealloc(c->u.big.tag, (size_t)c->cap * sizeof(uint32_t));
		c->u.big.freq = xrealloc(c->u.big.freq, (size_t)c->cap * sizeof(uint32_t));
	}
	if (c->cap == 0) {
		c->u.small.tag[c->items] = tag;
		c->u.small.freq[c->items] = 1;
	} else {
		c->u.big.tag[c->items] = tag;
		c->u.big.freq[c->items] = 1; /* context.c:54-55 */
	}
	c->items++;
	c->total++;
	if (c->items == CTX_LINEAR + 1) {
		for (uint32_t i = 0; i < c->items; ++i) {
			hmap_put(&s->map, ((uint64_t)id << 32) | c->u.big.tag[i], i);
		}
	} else if (c->items > CTX_LINEAR + 1) {
		hmap_put(&s->map, ((uint64_t)id << 32) | tag, c->items - 1);
	}
	if (c->items > CTX_FENWICK) {
		if (c->u.big.tree == NULL || c->items > c->u.big.tree_cap) {
			c->u.big.tree_cap = pow2_at_least(c->items * 2);
			c->u.big.tree = xrealloc(c->u.big.tree, ((size_t)c->u.big.tree_cap + 1) * sizeof(uint64_t));
			fen_build(c->u.big.tree, c->u.big.tree_cap, c->u.big.freq, c->items);
		} else {
			fen_add(c->u.big.tree, c->u.big.tree_cap, c->items - 1, 1);
		}
	}
}

void x3_ctx_inc(struct x3_ctxset *s, uint32_t id, uint32_t item)
{
	struct x3_ctx *c = x3_ctxset_get(s, id);
	c->total++;
	if (c->cap == 0) {
		c->u.small.freq[item]++;
		return;
	}
	c->u.big.freq[item]++;
	if (c->u.big.tree != NULL) {
		fen_add(c->u.big.tree, c->u.big.tree_cap, item, 1);
	}
}

uint64_t x3_ctx_cum(const struct x3_ctx *c, uint32_t item)
{
	if (c->cap != 0 && c->u.big.tree != NULL) {
		return fen_prefix(c->u.big.tree, item);
	}
	const uint32_t *fr = x3_ctx_freqs(c);
	uint64_t s = 0;
	for (uint32_t i = 0; i < item; ++i) {
		s += fr[i];
	}
	return s;
}

uint32_t x3_ctx_find_value(const struct x3_ctx *c, uint64_t value, uint64_t *cum)
{
	if (c->cap != 0 && c->u.big.tree != NULL) {
		const uint32_t i = fen_find(c->u.big.tree, c->u.big.tree_cap, value, cum);
		if (i >= c->items) {
			abort();
		}
		return i;
	}
	const uint32_t *fr = x3_ctx_freqs(c);
	uint64_t s = 0;
	for (uint32_t i = 0; i < c->items; ++i) {
		if (value < s + fr[i]) {
			*cum = s;
			return i;
		}
		s += fr[i];
	}
	abort(); /* index_of_value, ac.c:169 */
}

/* ===================================================================================== */
/* tag-pair map (reference tag_pair.c)                                                    */
/* ===================================================================================== */

struct x3_pairmap {
	struct hmap map;
	uint32_t elems;
};

struct x3_pairmap *x3_pairmap_create(void)
{
	struct x3_pairmap *m = xmalloc(sizeof(*m));
	hmap_init(&m->map, 1024);
	m->elems = 0;
	return m;
}

void x3_pairmap_destroy(struct x3_pairmap *m)
{
	hmap_free(&m->map);
	free(m);
}

int64_t x3_pairmap_query(const struct x3_pairmap *m, uint32_t t0, uint32_t t1)
{
	return hmap_get(&m->map, ((uint64_t)t0 << 32) | t1);
}

void x3_pairmap_prefetch(const struct x3_pairmap *m, uint32_t t0, uint32_t t1)
{
	hmap_prefetch(&m->map, ((uint64_t)t0 << 32) | t1);
}

uint32_t x3_pairmap_add(struct x3_pairmap *m, uint32_t t0, uint32_t t1)
{
	hmap_put(&m->map, ((uint64_t)t0 << 32) | t1, m->elems); /* id = insertion ordinal, tag_pair.c:122 */
	return m->elems++;
}

uint32_t x3_pairmap_elems(const struct x3_pairmap *m)
{
	return m->elems;
}

/* ===================================================================================== */
/* dictionary (reference dict.c)                                                          */
/* ===================================================================================== */

struct x3_dict {
	/* elements, indexed by tag */
	uint8_t (*s)[X3_MAX_MATCH_LEN];
	uint8_t *len;
	uint32_t *stamp; /* use stamp of the element: larger = more recently used */
	uint32_t elems, cap;
	/* trie over the strings */
	int32_t *node_elem; /* tag of the string ending in this node, or -1 */
	uint32_t nodes, node_cap;
	struct hmap child; /* (node << 8 | byte) -> node */
	/* order statistics over stamps: one bit per stamp (live or not) and a Fenwick tree over the
	 * population counts of the 64-bit words -- six levels less to walk than a tree over the stamps */
	uint64_t *bits;      /* stamp_cap / 64 words */
	uint64_t *tree;      /* Fenwick over popcount(bits[w]) */
	int32_t *stamp_elem; /* element that holds a stamp, or -1 */
	uint32_t stamp_cap, next_stamp; /* stamp_cap: a power of two, at least 64 */
};

static inline void stamp_set(struct x3_dict *d, uint32_t st)
{
	d->bits[st >> 6] |= 1ull << (st & 63);
	fen_add(d->tree, d->stamp_cap >> 6, st >> 6, 1);
}

static inline void stamp_clear(struct x3_dict *d, uint32_t st)
{
	d->bits[st >> 6] &= ~(1ull << (st & 63));
	fen_add(d->tree, d->stamp_cap >> 6, st >> 6, (uint64_t)-1);
}

/* live stamps in [0, i) */
static inline uint64_t stamp_prefix(const struct x3_dict *d, uint32_t i)
{
	uint64_t s = fen_prefix(d->tree, i >> 6);
	if (i & 63) {
		s += (uint64_t)__builtin_popcountll(d->bits[i >> 6] & ((1ull << (i & 63)) - 1ull));
	}
	return s;
}

/* the live stamp with exactly k live stamps below it (k < number of live stamps) */
static inline uint32_t stamp_select(const struct x3_dict *d, uint64_t k)
{
	uint64_t cum;
	const uint32_t w = fen_find(d->tree, d->stamp_cap >> 6, k, &cum);
	if (w >= (d->stamp_cap >> 6)) {
		abort();
	}
	uint64_t word = d->bits[w];
	for (uint64_t r = k - cum; r > 0; --r) {
		word &= word - 1; /* drop the lowest live stamp of the word */
	}
	if (word == 0) {
		abort();
	}
	return (w << 6) + (uint32_t)__builtin_ctzll(word);
}

struct x3_dict *x3_dict_create(void)
{
	struct x3_dict *d = xcalloc(1, sizeof(*d));
	d->cap = 256;
	d->s = xmalloc((size_t)d->cap * X3_MAX_MATCH_LEN);
	d->len = xmalloc(d->cap);
	d->stamp = xmalloc((size_t)d->cap * sizeof(uint32_t));
	d->node_cap = 1024;
	d->node_elem = xmalloc((size_t)d->node_cap * sizeof(int32_t));
	d->node_elem[0] = -1;
	d->nodes = 1;
	hmap_init(&d->child, 4096);
	d->stamp_cap = 1024;
	d->bits = xcalloc((size_t)d->stamp_cap / 64, sizeof(uint64_t));
	d->tree = xcalloc((size_t)d->stamp_cap / 64 + 1, sizeof(uint64_t));
	d->stamp_elem = xmalloc((size_t)d->stamp_cap * sizeof(int32_t));
	for (uint32_t i = 0; i < d->stamp_cap; ++i) {
		d->stamp_elem[i] = -1;
	}
	return d;
}

void x3_dict_destroy(struct x3_dict *d)
{
	free(d->s);
	free(d->len);
	free(d->stamp);
	free(d->node_elem);
	hmap_free(&d->child);
	free(d->bits);
	free(d->tree);
	free(d->stamp_elem);
	free(d);
}

uint32_t x3_dict_elems(const struct x3_dict *d)
{
	return d->elems;
}

int64_t x3_dict_find_match(const struct x3_dict *d, const uint8_t *p)
{
	/* longest dictionary string that is a prefix of p (dict.c:105-130); strings are
	 * unique, so "longest" is unambiguous */
	int64_t best = -1;
	uint32_t node = 0;
	for (int i = 0; i < X3_MAX_MATCH_LEN; ++i) {
		const int64_t nx = hmap_get(&d->child, ((uint64_t)node << 8) | p[i]);
		if (nx < 0) {
			break;
		}
		node = (uint32_t)nx;
		if (d->node_elem[node] >= 0) {
			best = d->node_elem[node];
		}
	}
	return best;
}

int x3_dict_query(const struct x3_dict *d, const uint8_t *s, uint32_t len)
{
	uint32_t node = 0;
	for (uint32_t i = 0; i < len; ++i) {
		const int64_t nx = hmap_get(&d->child, ((uint64_t)node << 8) | s[i]);
		if (nx < 0) {
			return 0;
		}
		node = (uint32_t)nx;
	}
	return d->node_elem[node] >= 0;
}

/* gives every live element a stamp 0 .. elems-1 in the same order, doubling the
 * stamp space when it is more than half full */
static void dict_compact(struct x3_dict *d)
{
	uint32_t ncap = d->stamp_cap;
	while ((uint64_t)d->elems * 2 + 2 > ncap) {
		ncap *= 2;
	}
	int32_t *order = xmalloc((size_t)(d->elems ? d->elems : 1) * sizeof(int32_t));
	uint32_t k = 0;
	for (uint32_t st = 0; st < d->next_stamp; ++st) {
		if (d->stamp_elem[st] >= 0) {
			order[k++] = d->stamp_elem[st];
		}
	}
	if (ncap != d->stamp_cap) {
		d->stamp_elem = xrealloc(d->stamp_elem, (size_t)ncap * sizeof(int32_t));
		d->bits = xrealloc(d->bits, ((size_t)ncap / 64) * sizeof(uint64_t));
		d->tree = xrealloc(d->tree, ((size_t)ncap / 64 + 1) * sizeof(uint64_t));
		d->stamp_cap = ncap;
	}
	for (uint32_t i = 0; i < d->stamp_cap; ++i) {
		d->stamp_elem[i] = -1;
	}
	for (uint32_t i = 0; i < k; ++i) {
		d->stamp_elem[i] = order[i];
		d->stamp[order[i]] = i;
	}
	/* k leading live stamps */
	const uint32_t words = d->stamp_cap / 64;
	memset(d->tree, 0, ((size_t)words + 1) * sizeof(uint64_t));
	for (uint32_t w = 0; w < words; ++w) {
		const uint32_t lo = w * 64;
		d->bits[w] = k >= lo + 64 ? ~0ull : (k > lo ? (1ull << (k - lo)) - 1ull : 0ull);
	}
	for (uint32_t i = 1; i <= words; ++i) {
		d->tree[i] += (uint64_t)__builtin_popcountll(d->bits[i - 1]);
		const uint32_t j = i + (i & (0u - i));
		if (j <= words) {
			d->tree[j] += d->tree[i];
		}
	}
	d->next_stamp = k;
	free(order);
}

static void dict_stamp_front(struct x3_dict *d, uint32_t tag, int had_stamp)
{
	if (had_stamp) {
		const uint32_t old = d->stamp[tag];
		d->stamp_elem[old] = -1;
		stamp_clear(d, old);
	}
	if (d->next_stamp == d->stamp_cap) {
		if (!had_stamp) {
			/* the new element is not live yet: compact the others first */
			d->elems--;
			dict_compact(d);
			d->elems++;
		} else {
			d->elems--;
			dict_compact(d);
			d->elems++;
		}
	}
	const uint32_t st = d->next_stamp++;
	d->stamp[tag] = st;
	d->stamp_elem[st] = (int32_t)tag;
	stamp_set(d, st);
}

uint32_t x3_dict_insert(struct x3_dict *d, const uint8_t *s, uint32_t len)
{
	if (d->elems == d->cap) {
		d->cap *= 2; /* dict_enlarge, dict.c:26-36 */
		d->s = xrealloc(d->s, (size_t)d->cap * X3_MAX_MATCH_LEN);
		d->len = xrealloc(d->len, d->cap);
		d->stamp = xrealloc(d->stamp, (size_t)d->cap * sizeof(uint32_t));
	}
	const uint32_t tag = d->elems; /* tag = insertion ordinal, dict.c:100 */
	memset(d->s[tag], 0, X3_MAX_MATCH_LEN);
	memcpy(d->s[tag], s, len);
	d->len[tag] = (uint8_t)len;
	uint32_t node = 0;
	for (uint32_t i = 0; i < len; ++i) {
		const uint64_t key = ((uint64_t)node << 8) | s[i];
		int64_t nx = hmap_get(&d->child, key);
		if (nx < 0) {
			if (d->nodes == d->node_cap) {
				d->node_cap *= 2;
				d->node_elem = xrealloc(d->node_elem, (size_t)d->node_cap * sizeof(int32_t));
			}
			nx = d->nodes++;
			d->node_elem[nx] = -1;
			hmap_put(&d->child, key, (uint32_t)nx);
		}
		node = (uint32_t)nx;
	}
	d->node_elem[node] = (int32_t)tag;
	d->elems++;
	/* the new element's cost after dict_update_costs is its own length: the smallest
	 * of all (dict.c:132-146), i.e. it moves to the front */
	dict_stamp_front(d, tag, 0);
	return tag;
}

void x3_dict_touch(struct x3_dict *d, uint32_t tag)
{
	dict_stamp_front(d, tag, 1);
}

uint32_t x3_dict_len(const struct x3_dict *d, uint32_t tag)
{
	return d->len[tag];
}

const uint8_t *x3_dict_str(const struct x3_dict *d, uint32_t tag)
{
	return d->s[tag];
}

uint32_t x3_dict_index_of(const struct x3_dict *d, uint32_t tag)
{
	/* number of elements used more recently = position in the cost-sorted array */
	const uint64_t upto = stamp_prefix(d, d->stamp[tag] + 1);
	return (uint32_t)(d->elems - upto);
}

uint32_t x3_dict_tag_at(const struct x3_dict *d, uint32_t index)
{
	if (index >= d->elems) {
		abort();
	}
	/* the element with exactly `index` more recent ones is the (elems - index)-th live
	 * stamp in ascending order: find the largest pos with prefix(pos) <= k - 1 */
	const uint64_t k = (uint64_t)d->elems - index;
	const uint32_t pos = stamp_select(d, k - 1);
	/* stamps [0,pos) hold k-1 live ones and stamp `pos` is live */
	if (pos >= d->stamp_cap || d->stamp_elem[pos] < 0) {
		abort();
	}
	return (uint32_t)d->stamp_elem[pos];
}
