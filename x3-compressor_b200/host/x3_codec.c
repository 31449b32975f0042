/*
 * x3_codec.c -- the parse loop and event model of x3 (reference x3.c:19-434) over
 * the structures of x3_structs.c.  Same events, same probabilities (float32, same
 * operand order, strict '>' tie order), same coded intervals: the stream is the
 * reference's, bit for bit.
 */
#define _POSIX_C_SOURCE 200809L
#include "x3_host.h"

#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CTX0_SHARDS 2
#define C0SET(c, id) ((c)->ctx0[(id) % CTX0_SHARDS])
#define C0ID(id) ((id) / CTX0_SHARDS)

struct x3_codec {
	struct x3_dict *dict;
	/* previous two tags (via the tag-pair map), x3.c:19; context id i lives in shard i % CTX0_SHARDS
	 * as its context i / CTX0_SHARDS, so that the pipeline can give every shard a thread of its own */
	struct x3_ctxset *ctx0[CTX0_SHARDS];
	struct x3_ctxset *ctx1; /* previous tag, x3.c:20 */
	struct x3_pairmap *pairs;
	struct x3_model events, match_size, chars, index1; /* x3.c:47-50 */
	struct x3_ac ac;
	struct x3_stats st;
	int nl;
	/* id of the pair (carry_t0, carry_t1) registered by the previous tag event: it is exactly the
	 * (prev_context1, context1) the next tag event asks for (x3.c:138-145 after x3.c:213-222) */
	int carry_valid;
	uint32_t carry_t0, carry_t1, carry_id;
};

static struct x3_codec *g_codec = NULL; /* for the dictionary callbacks of the search backend */

static float prob_to_bits(float prob)
{
	return -log2f(prob); /* x3.c:52-55 */
}

struct x3_codec *x3_codec_create(void)
{
	struct x3_codec *c = calloc(1, sizeof(*c));
	if (c == NULL) {
		abort();
	}
	c->dict = x3_dict_create();
	for (int i = 0; i < CTX0_SHARDS; ++i) {
		c->ctx0[i] = x3_ctxset_create();
	}
	c->ctx1 = x3_ctxset_create();
	c->pairs = x3_pairmap_create();
	x3_model_create(&c->events, X3_E_LAST);
	/* initial frequencies, x3.c:239-244 */
	x3_model_set(&c->events, X3_E_CTX0, 1024);
	x3_model_set(&c->events, X3_E_CTX1, 1024);
	x3_model_set(&c->events, X3_E_IDX1, 1);
	x3_model_set(&c->events, X3_E_NEW, 1);
	x3_model_create(&c->match_size, X3_MAX_MATCH_LEN);
	x3_model_create(&c->chars, 256);
	x3_model_create(&c->index1, 0);
	g_codec = c;
	return c;
}

void x3_codec_destroy(struct x3_codec *c)
{
	if (g_codec == c) {
		g_codec = NULL;
	}
	c->st.ctx0_entries = x3_pairmap_elems(c->pairs);
	c->st.ctx1_entries = x3_dict_elems(c->dict);
	x3_dict_destroy(c->dict);
	for (int i = 0; i < CTX0_SHARDS; ++i) {
		x3_ctxset_destroy(c->ctx0[i]);
	}
	x3_ctxset_destroy(c->ctx1);
	x3_pairmap_destroy(c->pairs);
	x3_model_destroy(&c->events);
	x3_model_destroy(&c->match_size);
	x3_model_destroy(&c->chars);
	x3_model_destroy(&c->index1);
	free(c);
}

void x3_codec_set_nl(struct x3_codec *c, int nl)
{
	c->nl = nl;
}

const struct x3_stats *x3_codec_stats(struct x3_codec *c)
{
	c->st.ctx0_entries = x3_pairmap_elems(c->pairs);
	c->st.ctx1_entries = x3_dict_elems(c->dict);
	return &c->st;
}

size_t x3_codec_dict_find(const char *p)
{
	const int64_t tag = x3_dict_find_match(g_codec->dict, (const uint8_t *)p);
	return tag < 0 ? (size_t)-1 : (size_t)tag;
}

size_t x3_codec_dict_len(size_t handle)
{
	return x3_dict_len(g_codec->dict, (uint32_t)handle);
}

static size_t nl_len(const struct x3_codec *c, size_t len)
{
	if (c->nl != 0) { /* x3.c:357-370 */
		switch (len - 1) {
			case 0: return 1;
			case 1: return 4;
			case 2: return 6;
			case 3: return 8;
			default: return 9999;
		}
	}
	return len;
}

static void enc_symbol(struct x3_codec *c, struct x3_bitw *w, struct x3_model *m, uint32_t sym)
{
	const uint64_t lo = x3_model_cum(m, sym);
	x3_ac_encode(&c->ac, w, lo, lo + m->freq[sym], m->total);
}

static uint32_t dec_symbol(struct x3_codec *c, struct x3_bitr *r, struct x3_model *m)
{
	uint64_t step;
	const uint64_t value = x3_ac_decode_target(&c->ac, m->total, &step);
	const uint32_t sym = x3_model_find(m, value);
	const uint64_t lo = x3_model_cum(m, sym);
	x3_ac_decode_update(&c->ac, r, step, lo, lo + m->freq[sym]);
	return sym;
}

/* (prev_context1, context1) -> linear context id, 0 when the pair is unknown (x3.c:138-145) */
static uint32_t ctx0_lookup(struct x3_codec *c, uint32_t prev_context1, uint32_t context1)
{
	if (c->carry_valid && c->carry_t0 == prev_context1 && c->carry_t1 == context1) {
		return c->carry_id;
	}
	const int64_t id = x3_pairmap_query(c->pairs, prev_context1, context1);
	return id < 0 ? 0u : (uint32_t)id;
}

/* (context1, tag) constitutes a new pair of tags (x3.c:211-222).  The pair map is independent of
 * the contexts, so the encoder registers the pair before it touches them: the id is the next tag
 * event's ctx0 id, whose cache line can be fetched while this event is being coded. */
static void register_pair(struct x3_codec *c, uint32_t context1, uint32_t tag)
{
	int64_t id = x3_pairmap_query(c->pairs, context1, tag);
	if (id < 0) {
		id = x3_pairmap_add(c->pairs, context1, tag);
	}
	const struct x3_ctx *nx = x3_ctxset_get(C0SET(c, (uint32_t)id), C0ID((uint32_t)id)); /* enlarge_ctx0 */
	__builtin_prefetch(nx, 1, 1);
	c->carry_valid = 1;
	c->carry_t0 = context1;
	c->carry_t1 = tag;
	c->carry_id = (uint32_t)id;
}

/* context update shared by encode_tag / decode_tag (x3.c:95-110,197-209) */
static void update_contexts(struct x3_codec *c, uint32_t ctx0_id, uint32_t context1, uint32_t tag, int64_t item0,
                            int64_t item1)
{
	if (item0 < 0) {
		x3_ctx_add(C0SET(c, ctx0_id), C0ID(ctx0_id), tag);
	} else {
		x3_ctx_inc(C0SET(c, ctx0_id), C0ID(ctx0_id), (uint32_t)item0);
	}
	if (item1 < 0) {
		x3_ctx_add(c->ctx1, context1, tag);
	} else {
		x3_ctx_inc(c->ctx1, context1, (uint32_t)item1);
	}
}

/* encode_tag, x3.c:132-223.  `tag` is the element, `index` its position in the
 * cost-sorted dictionary at this moment. */
static void encode_tag(struct x3_codec *c, struct x3_bitw *w, uint32_t prev_context1, uint32_t context1,
                       uint32_t tag, uint32_t index)
{
	const uint32_t ctx0_id = ctx0_lookup(c, prev_context1, context1);
	register_pair(c, context1, tag);

	const int64_t item0 = x3_ctx_find(C0SET(c, ctx0_id), C0ID(ctx0_id), tag);
	const int64_t item1 = x3_ctx_find(c->ctx1, context1, tag);
	const struct x3_ctx *c0 = x3_ctxset_get(C0SET(c, ctx0_id), C0ID(ctx0_id));
	const struct x3_ctx *c1 = x3_ctxset_get(c->ctx1, context1);

	/* x3.c:152-160: float products, in this order */
	float prob_ctx0 = 0;
	if (item0 >= 0) {
		prob_ctx0 = x3_model_prob(&c->events, X3_E_CTX0) * ((float)x3_ctx_freqs(c0)[item0] / (float)c0->total);
	}
	float prob_ctx1 = 0;
	if (item1 >= 0) {
		prob_ctx1 = x3_model_prob(&c->events, X3_E_CTX1) * ((float)x3_ctx_freqs(c1)[item1] / (float)c1->total);
	}
	const float prob_idx1 = x3_model_prob(&c->events, X3_E_IDX1) * x3_model_prob(&c->index1, index);

	int mode = X3_E_IDX1;
	float prob = prob_idx1;
	if (prob_ctx0 > prob) {
		mode = X3_E_CTX0;
		prob = prob_ctx0;
	}
	if (prob_ctx1 > prob) {
		mode = X3_E_CTX1;
		prob = prob_ctx1;
	}

	enc_symbol(c, w, &c->events, (uint32_t)mode);
	x3_model_inc(&c->events, (uint32_t)mode);

	switch (mode) {
		case X3_E_CTX0: {
			const uint64_t lo = x3_ctx_cum(c0, (uint32_t)item0);
			x3_ac_encode(&c->ac, w, lo, lo + x3_ctx_freqs(c0)[item0], c0->total);
			break;
		}
		case X3_E_CTX1: {
			const uint64_t lo = x3_ctx_cum(c1, (uint32_t)item1);
			x3_ac_encode(&c->ac, w, lo, lo + x3_ctx_freqs(c1)[item1], c1->total);
			break;
		}
		default:
			enc_symbol(c, w, &c->index1, index);
			x3_model_inc(&c->index1, index);
			break;
	}

	c->st.events[mode]++;
	c->st.sizes[mode] += prob_to_bits(prob);

	update_contexts(c, ctx0_id, context1, tag, item0, item1);
}

/* encode_match, x3.c:251-270 */
static void encode_match(struct x3_codec *c, struct x3_bitw *w, const uint8_t *p, size_t len)
{
	c->st.sizes[X3_E_NEW] += prob_to_bits(x3_model_prob(&c->events, X3_E_NEW));
	enc_symbol(c, w, &c->events, X3_E_NEW);
	x3_model_inc(&c->events, X3_E_NEW);

	c->st.sizes[X3_E_NEW] += prob_to_bits(x3_model_prob(&c->match_size, (uint32_t)(len - 1)));
	enc_symbol(c, w, &c->match_size, (uint32_t)(len - 1));
	x3_model_inc(&c->match_size, (uint32_t)(len - 1));

	for (size_t k = 0; k < len; ++k) {
		c->st.sizes[X3_E_NEW] += prob_to_bits(x3_model_prob(&c->chars, p[k]));
		enc_symbol(c, w, &c->chars, p[k]);
		x3_model_inc(&c->chars, p[k]);
	}
	c->st.events[X3_E_NEW]++;
}

/*
 * compress() is split into two stages that only communicate forwards:
 *
 *   parse  (x3.c:379-383,392-427 minus the coding): dictionary lookup, find_best_match(),
 *          hit/miss decision, move-to-front bookkeeping.  Depends on the dictionary only.
 *   code   (encode_tag / encode_match, x3.c:132-270): contexts, tag pairs, adaptive models,
 *          arithmetic coder, statistics.  Never feeds back into the parse.
 *
 * So the stages run as a two-thread pipeline over a single-producer/single-consumer ring of
 * step records; with X3_THREADS=1 the same two functions run in one thread.  The code stage is
 * split further below (X3_THREADS=4, the default).  The stream does not depend on the shape
 * (tests/test_host_x3.py runs every shape against the reference's streams).
 */
struct step_rec {
	uint32_t a;   /* hit: the element's tag;            miss: fragment length */
	uint32_t b;   /* hit: REC_HIT | MTF index;          miss: 1 if the fragment entered the dictionary */
	uint64_t off; /* miss: offset of the fragment in the input */
};
#define REC_HIT 0x80000000u
#define RING_LOG 15
#define RING_SIZE (1u << RING_LOG)
#define RING_BATCH 256u

struct coder_state {
	struct x3_codec *c;
	struct x3_bitw *w;
	const uint8_t *base;
	uint32_t prev_context1, context1; /* x3.c:376-377 */
	uint32_t dict_elems;
};

static void code_step(struct coder_state *cs, const struct step_rec *r)
{
	struct x3_codec *c = cs->c;
	if (r->b & REC_HIT) {
		encode_tag(c, cs->w, cs->prev_context1, cs->context1, r->a, r->b & ~REC_HIT);
		cs->prev_context1 = cs->context1;
		cs->context1 = r->a; /* dict_get_tag_by_index, x3.c:389-390 */
	} else {
		encode_match(c, cs->w, cs->base + r->off, r->a);
		if (r->b) {
			(void)x3_ctxset_get(c->ctx1, cs->dict_elems); /* enlarge_ctx1, x3.c:414-415 */
			cs->dict_elems++;
			x3_model_append(&c->index1); /* x3.c:419 */
		}
		cs->prev_context1 = 0; /* x3.c:423-424 */
		cs->context1 = 0;
	}
}

/* one parse step at p: fills r, returns the number of bytes consumed */
static size_t parse_step(struct x3_codec *c, uint8_t *p, uint8_t *ptr, uint8_t *end, x3_fbm_fn fbm, struct step_rec *r)
{
	/* (1) look into the dictionary, x3.c:381-383.  find_best_match() is a pure function of
	 * (p, dictionary state); the reference evaluates it lazily and possibly twice. */
	const int64_t tag = x3_dict_find_match(c->dict, p);
	size_t best = 0;
	int hit = 0;
	if (tag >= 0) {
		const size_t dl = x3_dict_len(c->dict, (uint32_t)tag);
		best = fbm((char *)p);
		hit = nl_len(c, dl) >= best && p + dl <= end;
	}
	if (hit) {
		const uint32_t len = x3_dict_len(c->dict, (uint32_t)tag);
		r->a = (uint32_t)tag;
		r->b = REC_HIT | x3_dict_index_of(c->dict, (uint32_t)tag);
		r->off = 0;
		x3_dict_touch(c->dict, (uint32_t)tag); /* dict_set_last_pos + dict_update_costs */
		return len;
	}
	/* (2) new fragment, x3.c:399-428 */
	size_t len = tag >= 0 ? best : fbm((char *)p);
	if (p + len > end) {
		len = (size_t)(end - p);
	}
	r->a = (uint32_t)len;
	r->b = 0;
	r->off = (uint64_t)(p - ptr);
	if (!x3_dict_query(c->dict, p, (uint32_t)len)) {
		x3_dict_insert(c->dict, p, (uint32_t)len);
		r->b = 1;
	}
	return len;
}

struct ring {
	struct step_rec *rec;
	uint64_t mask; /* RING_SIZE - 1, or all ones when rec holds every record (stage timing aid) */
	_Atomic uint64_t head; /* records published by the parser */
	_Atomic uint64_t tail; /* records consumed by the coder */
	_Atomic int done;
	struct coder_state *cs;
};

/*
 * Lookahead of the coder: the records behind the one being coded are already in the ring, and what
 * a tag event will look up is a function of the records alone (context1 = the previous hit's tag).
 * Two stages of cache hints run ahead of code_step(); they never change what is coded.
 *   far  (2 LOOK steps ahead): the pair-map slot of (context1, tag) and the line of ctx1[context1]
 *   near (LOOK steps ahead):   the pair's id if it is registered already -> the line of ctx0[id];
 *                              the item tables behind ctx1[context1]
 */
#define LOOK 6u

static inline uint32_t rec_context1(const struct ring *rg, uint64_t j)
{
	if (j == 0) {
		return 0;
	}
	const struct step_rec *pr = &rg->rec[(j - 1) & rg->mask];
	return (pr->b & REC_HIT) ? pr->a : 0u; /* x3.c:389-390,423-424 */
}

static inline void look_far(const struct ring *rg, uint64_t j)
{
	const struct step_rec *r = &rg->rec[j & rg->mask];
	if (r->b & REC_HIT) {
		const struct x3_codec *c = rg->cs->c;
		const uint32_t context1 = rec_context1(rg, j);
		x3_pairmap_prefetch(c->pairs, context1, r->a);
		x3_ctx_prefetch(c->ctx1, context1, r->a, 0);
	}
}

static inline void look_near(const struct ring *rg, uint64_t j)
{
	const struct step_rec *r = &rg->rec[j & rg->mask];
	if (r->b & REC_HIT) {
		const struct x3_codec *c = rg->cs->c;
		const uint32_t context1 = rec_context1(rg, j);
		x3_ctx_prefetch(c->ctx1, context1, r->a, 1);
		/* the pair this event registers is the NEXT tag event's ctx0 id */
		const int64_t id = x3_pairmap_query(c->pairs, context1, r->a);
		if (id >= 0) {
			x3_ctx_prefetch(C0SET(c, (uint32_t)id), C0ID((uint32_t)id), 0, 0);
		}
	}
}

static void *coder_thread(void *arg)
{
	struct ring *rg = arg;
	uint64_t tail = 0;
	for (;;) {
		uint64_t head = atomic_load_explicit(&rg->head, memory_order_acquire);
		if (head == tail) {
			if (atomic_load_explicit(&rg->done, memory_order_acquire)) {
				head = atomic_load_explicit(&rg->head, memory_order_acquire);
				if (head == tail) {
					break;
				}
			} else {
				sched_yield();
				continue;
			}
		}
		for (; tail < head; ++tail) {
			if (tail + 2 * LOOK < head) {
				look_far(rg, tail + 2 * LOOK);
			}
			if (tail + LOOK < head) {
				look_near(rg, tail + LOOK);
			}
			code_step(rg->cs, &rg->rec[tail & (RING_SIZE - 1)]);
		}
		atomic_store_explicit(&rg->tail, tail, memory_order_release);
	}
	return NULL;
}

/*
 * Five stages.  The coding of a tag event (x3.c:132-223) reads two context sets that never see
 * each other: ctx0 (addressed through the tag-pair map) and ctx1.  Both are updated by every tag
 * event whatever mode is chosen (x3.c:197-209), so each can be run ahead by a thread of its own
 * that only needs the step records:
 *
 *   parse  ->  pair stage: the tag-pair map: the ctx0 id of every tag event
 *          ->  ctx0 stage: ctx0: (found, freq, total, cum) of the tag, then the update
 *          ->  ctx1 stage: the same for ctx1 (and ctx1's growth with the dictionary, x3.c:414-415)
 *          ->  coder: event / index / match models, the mode decision on the float products of
 *              x3.c:152-172 (same operands, same order), arithmetic coder, statistics
 *
 * The cumulative frequency is taken eagerly (the reference takes it only for the chosen mode; its
 * value is the same).  Every structure has exactly one owner thread.
 */
struct ctx_out {
	uint64_t total, cum;
	uint32_t freq, found;
};

struct ring4 {
	struct step_rec *rec;
	struct ctx_out *o0, *o1;
	uint32_t *pid;               /* ctx0 id of every tag event, from the pair stage */
	uint64_t mask;               /* RING_SIZE - 1, or all ones when the arrays hold every record (stage timing aid) */
	_Atomic uint64_t head;       /* records published by the parser */
	_Atomic uint64_t tp;         /* records consumed by the pair stage */
	_Atomic uint64_t t0[CTX0_SHARDS], t1, t2; /* records consumed by the ctx0 shards, the ctx1 stage, the coder */
	_Atomic int done;
	struct coder_state *cs;
};

static inline uint32_t r4_context1(const struct ring4 *rg, uint64_t j)
{
	if (j == 0) {
		return 0;
	}
	const struct step_rec *pr = &rg->rec[(j - 1) & rg->mask];
	return (pr->b & REC_HIT) ? pr->a : 0u;
}

/* waits until the stage in front (`up`: the parser's head or another stage's tail) is past `tail`;
 * returns 0 when everything has been consumed */
static inline int r4_wait(struct ring4 *rg, _Atomic uint64_t *up, uint64_t tail, uint64_t *lim)
{
	for (;;) {
		*lim = atomic_load_explicit(up, memory_order_acquire);
		if (*lim != tail) {
			return 1;
		}
		if (atomic_load_explicit(&rg->done, memory_order_acquire) &&
		    atomic_load_explicit(&rg->head, memory_order_acquire) == tail) {
			return 0;
		}
		sched_yield();
	}
}

/* pair stage: the tag-pair map (x3.c:138-145, 211-222): which ctx0 context every tag event uses */
static void *pair_thread(void *arg)
{
	struct ring4 *rg = arg;
	struct x3_codec *c = rg->cs->c;
	uint32_t prev_context1 = 0, context1 = 0;
	uint64_t tail = 0, head;
	while (r4_wait(rg, &rg->head, tail, &head)) {
		for (; tail < head; ++tail) {
			if (tail + 2 * LOOK < head) {
				const struct step_rec *f = &rg->rec[(tail + 2 * LOOK) & rg->mask];
				if (f->b & REC_HIT) {
					x3_pairmap_prefetch(c->pairs, r4_context1(rg, tail + 2 * LOOK), f->a);
				}
			}
			const struct step_rec *r = &rg->rec[tail & rg->mask];
			if (r->b & REC_HIT) {
				const uint32_t tag = r->a;
				rg->pid[tail & rg->mask] = ctx0_lookup(c, prev_context1, context1);
				/* register (context1, tag): its id is the next tag event's ctx0 id */
				int64_t id = x3_pairmap_query(c->pairs, context1, tag);
				if (id < 0) {
					id = x3_pairmap_add(c->pairs, context1, tag);
				}
				c->carry_valid = 1;
				c->carry_t0 = context1;
				c->carry_t1 = tag;
				c->carry_id = (uint32_t)id;
				prev_context1 = context1;
				context1 = tag;
			} else {
				prev_context1 = 0;
				context1 = 0;
			}
		}
		atomic_store_explicit(&rg->tp, tail, memory_order_release);
	}
	return NULL;
}

struct ctx0_arg {
	struct ring4 *rg;
	uint32_t shard;
};

/* ctx0 stage, one thread per shard: the tag events whose ctx0 id falls into the shard */
static void *ctx0_thread(void *arg)
{
	const struct ctx0_arg *ca = arg;
	struct ring4 *rg = ca->rg;
	struct x3_codec *c = rg->cs->c;
	struct x3_ctxset *set = c->ctx0[ca->shard];
	uint64_t tail = 0, lim;
	while (r4_wait(rg, &rg->tp, tail, &lim)) {
		for (; tail < lim; ++tail) {
			/* the pair stage is ahead: the contexts of the events to come are known exactly */
			if (tail + 2 * LOOK < lim) {
				const uint64_t f = (tail + 2 * LOOK) & rg->mask;
				if ((rg->rec[f].b & REC_HIT) && rg->pid[f] % CTX0_SHARDS == ca->shard) {
					x3_ctx_prefetch(set, C0ID(rg->pid[f]), 0, 0);
				}
			}
			if (tail + LOOK < lim) {
				const uint64_t f = (tail + LOOK) & rg->mask;
				if ((rg->rec[f].b & REC_HIT) && rg->pid[f] % CTX0_SHARDS == ca->shard) {
					x3_ctx_prefetch(set, C0ID(rg->pid[f]), rg->rec[f].a, 1);
				}
			}
			const struct step_rec *r = &rg->rec[tail & rg->mask];
			if ((r->b & REC_HIT) && rg->pid[tail & rg->mask] % CTX0_SHARDS == ca->shard) {
				const uint32_t tag = r->a;
				const uint32_t id = C0ID(rg->pid[tail & rg->mask]);
				const int64_t item = x3_ctx_find(set, id, tag); /* grows the shard on demand (enlarge_ctx0) */
				struct ctx_out *o = &rg->o0[tail & rg->mask];
				o->found = item >= 0;
				if (item >= 0) {
					const struct x3_ctx *cx = x3_ctxset_get(set, id);
					o->freq = x3_ctx_freqs(cx)[item];
					o->total = cx->total;
					o->cum = x3_ctx_cum(cx, (uint32_t)item);
					x3_ctx_inc(set, id, (uint32_t)item);
				} else {
					x3_ctx_add(set, id, tag);
				}
			}
		}
		atomic_store_explicit(&rg->t0[ca->shard], tail, memory_order_release);
	}
	return NULL;
}

static void *ctx1_thread(void *arg)
{
	struct ring4 *rg = arg;
	struct x3_codec *c = rg->cs->c;
	uint32_t context1 = 0;
	uint32_t dict_elems = rg->cs->dict_elems;
	uint64_t tail = 0, head;
	while (r4_wait(rg, &rg->head, tail, &head)) {
		for (; tail < head; ++tail) {
			if (tail + 2 * LOOK < head) {
				const struct step_rec *f = &rg->rec[(tail + 2 * LOOK) & rg->mask];
				if (f->b & REC_HIT) {
					x3_ctx_prefetch(c->ctx1, r4_context1(rg, tail + 2 * LOOK), f->a, 0);
				}
			}
			if (tail + LOOK < head) {
				const struct step_rec *f = &rg->rec[(tail + LOOK) & rg->mask];
				if (f->b & REC_HIT) {
					x3_ctx_prefetch(c->ctx1, r4_context1(rg, tail + LOOK), f->a, 1);
				}
			}
			const struct step_rec *r = &rg->rec[tail & rg->mask];
			if (r->b & REC_HIT) {
				const uint32_t tag = r->a;
				const int64_t item = x3_ctx_find(c->ctx1, context1, tag);
				struct ctx_out *o = &rg->o1[tail & rg->mask];
				o->found = item >= 0;
				if (item >= 0) {
					const struct x3_ctx *cx = x3_ctxset_get(c->ctx1, context1);
					o->freq = x3_ctx_freqs(cx)[item];
					o->total = cx->total;
					o->cum = x3_ctx_cum(cx, (uint32_t)item);
					x3_ctx_inc(c->ctx1, context1, (uint32_t)item);
				} else {
					x3_ctx_add(c->ctx1, context1, tag);
				}
				context1 = tag;
			} else {
				if (r->b) {
					(void)x3_ctxset_get(c->ctx1, dict_elems); /* enlarge_ctx1, x3.c:414-415 */
					dict_elems++;
				}
				context1 = 0;
			}
		}
		atomic_store_explicit(&rg->t1, tail, memory_order_release);
	}
	return NULL;
}

/* the rest of encode_tag (x3.c:152-195) on the two stages' answers */
static void code_tag4(struct x3_codec *c, struct x3_bitw *w, const struct ctx_out *a0, const struct ctx_out *a1,
                      uint32_t index)
{
	float prob_ctx0 = 0;
	if (a0->found) {
		prob_ctx0 = x3_model_prob(&c->events, X3_E_CTX0) * ((float)a0->freq / (float)a0->total);
	}
	float prob_ctx1 = 0;
	if (a1->found) {
		prob_ctx1 = x3_model_prob(&c->events, X3_E_CTX1) * ((float)a1->freq / (float)a1->total);
	}
	const float prob_idx1 = x3_model_prob(&c->events, X3_E_IDX1) * x3_model_prob(&c->index1, index);

	int mode = X3_E_IDX1;
	float prob = prob_idx1;
	if (prob_ctx0 > prob) {
		mode = X3_E_CTX0;
		prob = prob_ctx0;
	}
	if (prob_ctx1 > prob) {
		mode = X3_E_CTX1;
		prob = prob_ctx1;
	}

	enc_symbol(c, w, &c->events, (uint32_t)mode);
	x3_model_inc(&c->events, (uint32_t)mode);

	switch (mode) {
		case X3_E_CTX0:
			x3_ac_encode(&c->ac, w, a0->cum, a0->cum + a0->freq, a0->total);
			break;
		case X3_E_CTX1:
			x3_ac_encode(&c->ac, w, a1->cum, a1->cum + a1->freq, a1->total);
			break;
		default:
			enc_symbol(c, w, &c->index1, index);
			x3_model_inc(&c->index1, index);
			break;
	}
	c->st.events[mode]++;
	c->st.sizes[mode] += prob_to_bits(prob);
}

static void *coder4_thread(void *arg)
{
	struct ring4 *rg = arg;
	struct coder_state *cs = rg->cs;
	struct x3_codec *c = cs->c;
	uint64_t tail = 0;
	for (;;) {
		/* a record can be coded once both context stages are through with it */
		uint64_t lim = atomic_load_explicit(&rg->t1, memory_order_acquire);
		for (int i = 0; i < CTX0_SHARDS; ++i) {
			const uint64_t l0 = atomic_load_explicit(&rg->t0[i], memory_order_acquire);
			if (l0 < lim) {
				lim = l0;
			}
		}
		if (lim == tail) {
			if (atomic_load_explicit(&rg->done, memory_order_acquire) &&
			    atomic_load_explicit(&rg->head, memory_order_acquire) == tail) {
				break;
			}
			sched_yield();
			continue;
		}
		for (; tail < lim; ++tail) {
			const struct step_rec *r = &rg->rec[tail & rg->mask];
			if (r->b & REC_HIT) {
				code_tag4(c, cs->w, &rg->o0[tail & rg->mask], &rg->o1[tail & rg->mask], r->b & ~REC_HIT);
			} else {
				encode_match(c, cs->w, cs->base + r->off, r->a);
				if (r->b) {
					x3_model_append(&c->index1); /* x3.c:419 */
				}
			}
		}
		atomic_store_explicit(&rg->t2, tail, memory_order_release);
	}
	return NULL;
}

void *x3_compress(struct x3_codec *c, char *base, size_t isize, x3_fbm_fn fbm, size_t *out_bytes)
{
	struct x3_bitw w;
	x3_bitw_open(&w, isize / 2 + 64);
	x3_ac_init(&c->ac);
	g_codec = c;

	uint8_t *ptr = (uint8_t *)base;
	uint8_t *end = ptr + isize;
	struct coder_state cs = {c, &w, ptr, 0, 0, x3_dict_elems(c->dict)};

	const char *env = getenv("X3_THREADS");
	/* 1: one thread; 2: parse | code; 4 (default where 4 cores are online): parse | ctx0 | ctx1 | code */
	const int threads = env != NULL ? atoi(env) : (sysconf(_SC_NPROCESSORS_ONLN) >= 4 ? 4 : 2);
	if (threads == -4) {
		/* stage timing aid for the four-stage shape: every stage alone, one after the other */
		struct ring4 lin;
		lin.rec = malloc(sizeof(struct step_rec) * (isize + 1));
		lin.o0 = malloc(sizeof(struct ctx_out) * (isize + 1));
		lin.o1 = malloc(sizeof(struct ctx_out) * (isize + 1));
		lin.pid = malloc(sizeof(uint32_t) * (isize + 1));
		lin.mask = ~(uint64_t)0;
		lin.cs = &cs;
		atomic_init(&lin.head, 0);
		atomic_init(&lin.tp, 0);
		for (int i = 0; i < CTX0_SHARDS; ++i) {
			atomic_init(&lin.t0[i], 0);
		}
		atomic_init(&lin.t1, 0);
		atomic_init(&lin.t2, 0);
		atomic_init(&lin.done, 0);
		size_t nrec = 0;
		struct timespec t[6];
		clock_gettime(CLOCK_MONOTONIC, &t[0]);
		for (uint8_t *p = ptr; p < end;) {
			p += parse_step(c, p, ptr, end, fbm, &lin.rec[nrec++]);
		}
		atomic_store(&lin.head, nrec);
		atomic_store(&lin.done, 1);
		clock_gettime(CLOCK_MONOTONIC, &t[1]);
		pair_thread(&lin);
		clock_gettime(CLOCK_MONOTONIC, &t[2]);
		for (uint32_t i = 0; i < CTX0_SHARDS; ++i) {
			struct ctx0_arg one = {&lin, i};
			ctx0_thread(&one);
		}
		clock_gettime(CLOCK_MONOTONIC, &t[3]);
		ctx1_thread(&lin);
		clock_gettime(CLOCK_MONOTONIC, &t[4]);
		coder4_thread(&lin);
		clock_gettime(CLOCK_MONOTONIC, &t[5]);
		double d[5];
		for (int i = 0; i < 5; ++i) {
			d[i] = (t[i + 1].tv_sec - t[i].tv_sec) + (t[i + 1].tv_nsec - t[i].tv_nsec) * 1e-9;
		}
		fprintf(stderr, "x3_compress stages: parse %.3f s, pairs %.3f s, ctx0 (all shards, one after the other) %.3f s, ctx1 %.3f s, code %.3f s, %zu steps\n",
		        d[0], d[1], d[2], d[3], d[4], nrec);
		free(lin.rec);
		free(lin.o0);
		free(lin.o1);
		free(lin.pid);
	} else if (threads == -1) {
		/* stage timing aid: parse everything, then code everything */
		struct step_rec *all = malloc(sizeof(struct step_rec) * (isize + 1));
		size_t nrec = 0;
		struct timespec t0, t1, t2;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for (uint8_t *p = ptr; p < end;) {
			p += parse_step(c, p, ptr, end, fbm, &all[nrec++]);
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		struct ring lin;
		lin.rec = all;
		lin.mask = ~(uint64_t)0;
		lin.cs = &cs;
		const int look = getenv("X3_NO_LOOKAHEAD") == NULL;
		for (size_t i = 0; i < nrec; ++i) {
			if (look && i + 2 * LOOK < nrec) {
				look_far(&lin, i + 2 * LOOK);
			}
			if (look && i + LOOK < nrec) {
				look_near(&lin, i + LOOK);
			}
			code_step(&cs, &all[i]);
		}
		clock_gettime(CLOCK_MONOTONIC, &t2);
		fprintf(stderr, "x3_compress stages: parse %.3f s, code %.3f s, %zu steps\n",
		        (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9,
		        (t2.tv_sec - t1.tv_sec) + (t2.tv_nsec - t1.tv_nsec) * 1e-9, nrec);
		free(all);
	} else if (threads <= 1 || (env == NULL && isize < 65536)) { /* small inputs: not worth the threads, unless asked for */
		struct step_rec r;
		for (uint8_t *p = ptr; p < end;) {
			p += parse_step(c, p, ptr, end, fbm, &r);
			code_step(&cs, &r);
		}
	} else if (threads >= 4) {
		struct ring4 rg;
		rg.rec = malloc(sizeof(struct step_rec) * RING_SIZE);
		rg.o0 = malloc(sizeof(struct ctx_out) * RING_SIZE);
		rg.o1 = malloc(sizeof(struct ctx_out) * RING_SIZE);
		rg.pid = malloc(sizeof(uint32_t) * RING_SIZE);
		if (rg.rec == NULL || rg.o0 == NULL || rg.o1 == NULL || rg.pid == NULL) {
			abort();
		}
		rg.mask = RING_SIZE - 1;
		atomic_init(&rg.head, 0);
		atomic_init(&rg.tp, 0);
		for (int i = 0; i < CTX0_SHARDS; ++i) {
			atomic_init(&rg.t0[i], 0);
		}
		atomic_init(&rg.t1, 0);
		atomic_init(&rg.t2, 0);
		atomic_init(&rg.done, 0);
		rg.cs = &cs;
		pthread_t th[3 + CTX0_SHARDS];
		struct ctx0_arg ca[CTX0_SHARDS];
		int nth = 0;
		int bad = pthread_create(&th[nth++], NULL, pair_thread, &rg) != 0;
		for (int i = 0; i < CTX0_SHARDS; ++i) {
			ca[i].rg = &rg;
			ca[i].shard = (uint32_t)i;
			bad |= pthread_create(&th[nth++], NULL, ctx0_thread, &ca[i]) != 0;
		}
		bad |= pthread_create(&th[nth++], NULL, ctx1_thread, &rg) != 0;
		bad |= pthread_create(&th[nth++], NULL, coder4_thread, &rg) != 0;
		if (bad) {
			abort();
		}
		uint64_t head = 0, published = 0, tail_seen = 0;
		for (uint8_t *p = ptr; p < end;) {
			while (head - tail_seen >= RING_SIZE) { /* ring full: wait for the coder (the last stage) */
				tail_seen = atomic_load_explicit(&rg.t2, memory_order_acquire);
				if (head - tail_seen >= RING_SIZE) {
					sched_yield();
				}
			}
			p += parse_step(c, p, ptr, end, fbm, &rg.rec[head & (RING_SIZE - 1)]);
			++head;
			if (head - published >= RING_BATCH) {
				atomic_store_explicit(&rg.head, head, memory_order_release);
				published = head;
			}
		}
		atomic_store_explicit(&rg.head, head, memory_order_release);
		atomic_store_explicit(&rg.done, 1, memory_order_release);
		for (int i = 0; i < nth; ++i) {
			pthread_join(th[i], NULL);
		}
		free(rg.rec);
		free(rg.o0);
		free(rg.o1);
		free(rg.pid);
	} else {
		struct ring rg;
		rg.rec = malloc(sizeof(struct step_rec) * RING_SIZE);
		if (rg.rec == NULL) {
			abort();
		}
		rg.mask = RING_SIZE - 1;
		atomic_init(&rg.head, 0);
		atomic_init(&rg.tail, 0);
		atomic_init(&rg.done, 0);
		rg.cs = &cs;
		pthread_t th;
		if (pthread_create(&th, NULL, coder_thread, &rg) != 0) {
			abort();
		}
		uint64_t head = 0, published = 0, tail_seen = 0;
		for (uint8_t *p = ptr; p < end;) {
			while (head - tail_seen >= RING_SIZE) { /* ring full: wait for the coder */
				tail_seen = atomic_load_explicit(&rg.tail, memory_order_acquire);
				if (head - tail_seen >= RING_SIZE) {
					sched_yield();
				}
			}
			p += parse_step(c, p, ptr, end, fbm, &rg.rec[head & (RING_SIZE - 1)]);
			++head;
			if (head - published >= RING_BATCH) {
				atomic_store_explicit(&rg.head, head, memory_order_release);
				published = head;
			}
		}
		atomic_store_explicit(&rg.head, head, memory_order_release);
		atomic_store_explicit(&rg.done, 1, memory_order_release);
		pthread_join(th, NULL);
		free(rg.rec);
	}

	/* signal end of input, x3.c:431-433 */
	enc_symbol(c, &w, &c->events, X3_E_EOF);
	x3_model_inc(&c->events, X3_E_EOF);

	x3_ac_encode_flush(&c->ac, &w); /* x3.c:603 */
	*out_bytes = x3_bitw_close(&w); /* x3.c:604 */
	return w.buf;
}

/* decode_tag, x3.c:58-129: returns the element (tag) */
static uint32_t decode_tag(struct x3_codec *c, struct x3_bitr *r, uint32_t decision, uint32_t prev_context1,
                           uint32_t context1)
{
	const uint32_t ctx0_id = ctx0_lookup(c, prev_context1, context1);
	const struct x3_ctx *c0 = x3_ctxset_get(C0SET(c, ctx0_id), C0ID(ctx0_id));
	const struct x3_ctx *c1 = x3_ctxset_get(c->ctx1, context1);

	uint32_t tag;
	float size;
	switch (decision) {
		case X3_E_CTX0:
		case X3_E_CTX1: {
			const struct x3_ctx *cx = decision == X3_E_CTX0 ? c0 : c1;
			if (cx->items == 0 || cx->total == 0) {
				abort();
			}
			uint64_t step, lo;
			const uint64_t value = x3_ac_decode_target(&c->ac, cx->total, &step);
			const uint32_t item = x3_ctx_find_value(cx, value, &lo);
			x3_ac_decode_update(&c->ac, r, step, lo, lo + x3_ctx_freqs(cx)[item]);
			tag = x3_ctx_tags(cx)[item];
			size = prob_to_bits((float)x3_ctx_freqs(cx)[item] / (float)cx->total);
			break;
		}
		case X3_E_IDX1: {
			if (c->index1.n == 0) {
				abort();
			}
			const uint32_t index = dec_symbol(c, r, &c->index1);
			size = prob_to_bits(x3_model_prob(&c->index1, index));
			x3_model_inc(&c->index1, index);
			tag = x3_dict_tag_at(c->dict, index);
			break;
		}
		default:
			abort();
	}
	c->st.events[decision]++;
	c->st.sizes[decision] += size;

	const int64_t item0 = x3_ctx_find(C0SET(c, ctx0_id), C0ID(ctx0_id), tag);
	const int64_t item1 = x3_ctx_find(c->ctx1, context1, tag);
	update_contexts(c, ctx0_id, context1, tag, item0, item1);
	register_pair(c, context1, tag);
	return tag;
}

void *x3_decompress(struct x3_codec *c, const void *stream, size_t bytes, size_t *out_bytes)
{
	struct x3_bitr r;
	x3_bitr_open(&r, stream, bytes);
	x3_ac_init(&c->ac);
	x3_ac_decode_init(&c->ac, &r);
	g_codec = c;

	size_t cap = bytes * 4 + 4096, n = 0;
	uint8_t *out = malloc(cap);
	if (out == NULL) {
		abort();
	}
	uint32_t prev_context1 = 0, context1 = 0;

	for (;;) {
		const uint32_t decision = dec_symbol(c, &r, &c->events);
		c->st.sizes[decision] += prob_to_bits(x3_model_prob(&c->events, decision));
		x3_model_inc(&c->events, decision);

		if (n + X3_MAX_MATCH_LEN > cap) {
			cap *= 2;
			out = realloc(out, cap);
			if (out == NULL) {
				abort();
			}
		}
		if (decision == X3_E_EOF) {
			break;
		} else if (decision == X3_E_NEW) {
			/* decode_match, x3.c:272-283 */
			const uint32_t len = dec_symbol(c, &r, &c->match_size) + 1;
			c->st.sizes[X3_E_NEW] += prob_to_bits(x3_model_prob(&c->match_size, len - 1));
			x3_model_inc(&c->match_size, len - 1);
			for (uint32_t k = 0; k < len; ++k) {
				const uint32_t ch = dec_symbol(c, &r, &c->chars);
				out[n + k] = (uint8_t)ch;
				c->st.sizes[X3_E_NEW] += prob_to_bits(x3_model_prob(&c->chars, ch));
				x3_model_inc(&c->chars, ch);
			}
			if (!x3_dict_query(c->dict, out + n, len)) {
				x3_dict_insert(c->dict, out + n, len);
				(void)x3_ctxset_get(c->ctx1, x3_dict_elems(c->dict) - 1);
				x3_model_append(&c->index1);
			}
			n += len;
			prev_context1 = 0;
			context1 = 0;
			c->st.events[X3_E_NEW]++;
		} else {
			const uint32_t tag = decode_tag(c, &r, decision, prev_context1, context1);
			const uint32_t len = x3_dict_len(c->dict, tag);
			prev_context1 = context1;
			context1 = tag;
			memcpy(out + n, x3_dict_str(c->dict, tag), len);
			x3_dict_touch(c->dict, tag);
			n += len;
		}
	}
	*out_bytes = n;
	return out;
}
