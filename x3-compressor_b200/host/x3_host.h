/*
 * x3_host.h -- the sequential host pass of x3:
 * dictionary with move-to-front order, context models, tag-pair map, adaptive
 * models, arithmetic coder and bit I/O.
 *
 * Every structure here is a re-design with sub-linear operations of a structure
 * the reference implements with linear scans, full recomputation or qsort per
 * step (reference dict.c, context.c, tag_pair.c, ac.c, bio.c).  The observable
 * behaviour is frozen: every probability, every coded interval and therefore
 * every emitted bit equals the reference's.  Each function cites what it replaces.
 *
 * C99; no CUDA types.  The GPU search is reached only through x3_backend.h.
 */
#ifndef X3_HOST_H
#define X3_HOST_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#define X3_MAX_MATCH_LEN 32 /* reference backend.h:7-10 */

/* ---- bit I/O: reference bio.c (32-bit words, bits packed LSB first) ---------------- */
struct x3_bitw {
	uint32_t *buf;
	size_t cap, n; /* words */
	uint64_t acc;
	unsigned nacc;
};
void x3_bitw_open(struct x3_bitw *w, size_t reserve_bytes);
void x3_bitw_put(struct x3_bitw *w, unsigned bit);
void x3_bitw_run(struct x3_bitw *w, unsigned bit, uint64_t count);
size_t x3_bitw_close(struct x3_bitw *w); /* flushes the partial word (bio.c:105-112); returns bytes */

struct x3_bitr {
	const uint32_t *ptr;
	const uint32_t *end; /* first word that may no longer be read (bio.c:10: end - 3 bytes) */
	uint32_t b;
	unsigned c;
};
void x3_bitr_open(struct x3_bitr *r, const void *buf, size_t bytes);
unsigned x3_bitr_get(struct x3_bitr *r);

/* ---- arithmetic coder: reference ac.c:31-198 --------------------------------------- */
struct x3_ac {
	uint64_t low, high, buffer, scale;
};
void x3_ac_init(struct x3_ac *ac);
void x3_ac_encode(struct x3_ac *ac, struct x3_bitw *w, uint64_t low_freq, uint64_t high_freq, uint64_t total);
void x3_ac_encode_flush(struct x3_ac *ac, struct x3_bitw *w);
void x3_ac_decode_init(struct x3_ac *ac, struct x3_bitr *r);
uint64_t x3_ac_decode_target(struct x3_ac *ac, uint64_t total, uint64_t *step);
void x3_ac_decode_update(struct x3_ac *ac, struct x3_bitr *r, uint64_t step, uint64_t low_freq, uint64_t high_freq);

/* ---- adaptive model over symbols 0..n-1: reference ac.c:200-273 --------------------
 * freq[] plus a Fenwick tree for the cumulative frequencies the reference recomputes
 * in full on every inc_model(). */
struct x3_model {
	uint32_t n, cap; /* cap: power of two */
	uint64_t total;
	uint32_t *freq;
	uint64_t *tree; /* 1-based Fenwick over freq */
};
void x3_model_create(struct x3_model *m, uint32_t n);               /* all freq 1 (ac.c:230-247) */
void x3_model_set(struct x3_model *m, uint32_t i, uint32_t freq);   /* x3.c:239-244 */
void x3_model_append(struct x3_model *m);                           /* model_enlarge, ac.c:249-265 */
void x3_model_inc(struct x3_model *m, uint32_t i);                  /* inc_model, ac.c:215-228 */
uint64_t x3_model_cum(const struct x3_model *m, uint32_t i);        /* sum of freq[0..i) */
uint32_t x3_model_find(const struct x3_model *m, uint64_t value);   /* index_of_value, ac.c:157-170 */
void x3_model_destroy(struct x3_model *m);
static inline float x3_model_prob(const struct x3_model *m, uint32_t i)
{
	return (float)m->freq[i] / (float)m->total; /* ac.c:108-113 */
}

/* ---- context: reference context.c (items in insertion order, ctx_sort is a no-op) --- */
#define X3_CTX_INLINE 6
struct x3_ctx { /* exactly one 64-byte cache line: most contexts hold a handful of tags */
	uint32_t items;
	uint32_t cap; /* 0: the items live in u.small */
	uint64_t total;
	union {
		struct {
			uint32_t tag[X3_CTX_INLINE], freq[X3_CTX_INLINE];
		} small;
		struct {
			uint32_t *tag, *freq;
			uint64_t *tree; /* Fenwick over freq once the context is large, else NULL */
			uint32_t tree_cap;
		} big;
	} u;
};
static inline const uint32_t *x3_ctx_tags(const struct x3_ctx *c)
{
	return c->cap ? c->u.big.tag : c->u.small.tag;
}
static inline const uint32_t *x3_ctx_freqs(const struct x3_ctx *c)
{
	return c->cap ? c->u.big.freq : c->u.small.freq;
}
struct x3_ctxset; /* growable array of contexts + (context, tag) -> item index */
struct x3_ctxset *x3_ctxset_create(void);
void x3_ctxset_destroy(struct x3_ctxset *s);
struct x3_ctx *x3_ctxset_get(struct x3_ctxset *s, uint32_t id); /* grows on demand (ctx_enlarge) */
int64_t x3_ctx_find(struct x3_ctxset *s, uint32_t id, uint32_t tag); /* item index or -1 (context.c:20-40) */
/* cache hints for a lookup a few steps ahead: the context's line (with_items = 0), then what a
 * lookup of `tag` in it will touch (with_items = 1, once the line has arrived) */
void x3_ctx_prefetch(const struct x3_ctxset *s, uint32_t id, uint32_t tag, int with_items);
void x3_ctx_add(struct x3_ctxset *s, uint32_t id, uint32_t tag);     /* ctx_add_tag, context.c:42-56 */
void x3_ctx_inc(struct x3_ctxset *s, uint32_t id, uint32_t item);    /* ctx_item_inc_freq, context.c:88-93 */
uint64_t x3_ctx_cum(const struct x3_ctx *c, uint32_t item);
uint32_t x3_ctx_find_value(const struct x3_ctx *c, uint64_t value, uint64_t *cum);

/* ---- tag-pair map: reference tag_pair.c ((tag0, tag1) -> insertion ordinal) ---------- */
struct x3_pairmap;
struct x3_pairmap *x3_pairmap_create(void);
void x3_pairmap_destroy(struct x3_pairmap *m);
int64_t x3_pairmap_query(const struct x3_pairmap *m, uint32_t t0, uint32_t t1); /* -1 if absent */
uint32_t x3_pairmap_add(struct x3_pairmap *m, uint32_t t0, uint32_t t1);
void x3_pairmap_prefetch(const struct x3_pairmap *m, uint32_t t0, uint32_t t1);
uint32_t x3_pairmap_elems(const struct x3_pairmap *m);

/* ---- dictionary: reference dict.c --------------------------------------------------
 * Elements are identified by their tag (= insertion ordinal, dict.c:101).  The
 * reference keeps the array sorted by cost = p - last_pos with a qsort per step
 * (dict.c:132-146); all last_pos are distinct, so that order is exactly
 * most-recently-used first.  Here: a trie for longest-prefix / exact queries and a
 * Fenwick tree over use stamps for the MTF rank ("index") in both directions. */
struct x3_dict;
struct x3_dict *x3_dict_create(void);
void x3_dict_destroy(struct x3_dict *d);
uint32_t x3_dict_elems(const struct x3_dict *d);
int64_t x3_dict_find_match(const struct x3_dict *d, const uint8_t *p);        /* tag of the longest prefix or -1 (dict.c:105-130) */
int x3_dict_query(const struct x3_dict *d, const uint8_t *s, uint32_t len);   /* dict_query_elem, dict.c:148-157 */
uint32_t x3_dict_insert(struct x3_dict *d, const uint8_t *s, uint32_t len);   /* dict_insert_elem + MTF to front; returns tag */
void x3_dict_touch(struct x3_dict *d, uint32_t tag);                          /* dict_set_last_pos + dict_update_costs */
uint32_t x3_dict_len(const struct x3_dict *d, uint32_t tag);
const uint8_t *x3_dict_str(const struct x3_dict *d, uint32_t tag);
uint32_t x3_dict_index_of(const struct x3_dict *d, uint32_t tag);             /* MTF rank of an element */
uint32_t x3_dict_tag_at(const struct x3_dict *d, uint32_t index);             /* element at an MTF rank */

/* ---- codec: reference x3.c:19-434 --------------------------------------------------- */
enum { X3_E_CTX0 = 0, X3_E_CTX1, X3_E_IDX1, X3_E_NEW, X3_E_EOF, X3_E_LAST }; /* x3.c:33-40 */

struct x3_stats {
	size_t events[X3_E_LAST];
	float sizes[X3_E_LAST];
	size_t ctx0_entries, ctx1_entries;
};

struct x3_codec;
struct x3_codec *x3_codec_create(void); /* create(), x3.c:225-249 */
void x3_codec_destroy(struct x3_codec *c);
void x3_codec_set_nl(struct x3_codec *c, int nl); /* -x, x3.c:355-370 */
const struct x3_stats *x3_codec_stats(struct x3_codec *c);
/* dictionary queries in the shape backend.c:80,86 expects (find returns a handle or (size_t)-1) */
size_t x3_codec_dict_find(const char *p);
size_t x3_codec_dict_len(size_t handle);

/* search callback: find_best_match(p) of backend.h */
typedef size_t (*x3_fbm_fn)(char *p);

/* compress(), x3.c:372-434 + flush (x3.c:603-604).  base: isize bytes followed by
 * padding; returns the stream (malloc'd) and its size in bytes. */
void *x3_compress(struct x3_codec *c, char *base, size_t isize, x3_fbm_fn fbm, size_t *out_bytes);
/* decompress(), x3.c:285-353.  Returns the data (malloc'd, grown on demand). */
void *x3_decompress(struct x3_codec *c, const void *stream, size_t bytes, size_t *out_bytes);

#endif
