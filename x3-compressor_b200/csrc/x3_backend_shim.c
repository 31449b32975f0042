/*
 * x3_backend_shim.c -- the nine symbols of the reference's backend.h
 * (reference backend.h:20-31, backend.c:8-100) on top of the GPU search.
 *
 * Split:
 *   a1 histogram loop   backend.c:58-74   -> GPU, all positions at once
 *   a2 selection        backend.c:76-78   -> GPU epilogue, Lstar[p]
 *   a3 dictionary filter backend.c:79-90  -> here, per call, live dictionary
 *
 * C99, no CUDA types.  Talks to the device layer only through x3_search.h.
 */
#define _GNU_SOURCE
#include "x3_backend.h"
#include "x3_search.h"

#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>

#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif

/* reference backend.c:8,21,33,34 */
static size_t g_forward_window = 8 * 1024;
static int g_max_match_count = 15;
static size_t g_factor1 = 4;
static size_t g_factor2 = 0;

void set_forward_window(size_t n) { g_forward_window = n; }
size_t get_forward_window(void) { return g_forward_window; }
void set_max_match_count(int n) { g_max_match_count = n; }
int get_max_match_count(void) { return g_max_match_count; }
size_t get_magic_factor1(void) { return g_factor1; }
void set_magic_factor1(size_t factor) { g_factor1 = factor; }
size_t get_magic_factor2(void) { return g_factor2; }
void set_magic_factor2(size_t factor) { g_factor2 = factor; }

/* Dictionary queries (reference dict.c:105-130, dict.c:159-162).  A C host that
 * links its dict.c next to this shim provides them as ordinary symbols; the
 * weak declarations let the shared library load in hosts that register
 * callbacks instead. */
extern size_t dict_find_match(const char *p) __attribute__((weak));
extern size_t dict_get_len_by_index(size_t index) __attribute__((weak));

static x3_dict_find_fn g_dict_find = NULL;
static x3_dict_len_fn g_dict_len = NULL;

void x3_backend_set_dict(x3_dict_find_fn find, x3_dict_len_fn len)
{
	g_dict_find = find;
	g_dict_len = len;
}

static const char *g_base = NULL;
static size_t g_isize = 0;
static uint8_t *g_lstar = NULL;
static uint8_t *g_spare = NULL;       /* the last table's memory, kept for the next prepare (X3_TABLE_KEEP=1) */
static size_t g_spare_bytes = 0;
static size_t g_lstar_bytes = 0;
static uint8_t *g_table = NULL;
static int g_lstar_pinned = 0;       /* g_lstar came from x3s_host_alloc */
static void *g_registered = NULL;    /* the caller's buffer, page-locked in place for the duration of the tables */
static double g_prepare_ms = 0.0;
static double g_startup_ms = 0.0; /* one-off CUDA start-up (driver + first context) inside the last prepare */
static double g_register_ms = 0.0; /* page-locking the caller's buffer inside the last prepare */

/* The search runs on a thread of its own (x3s_search_host_stream) and the table lands piece by piece,
 * from the left: find_best_match(p) answers as soon as position p has landed -- the reference's
 * compress() reads positions in increasing order (x3.c:379), so it starts on the first piece while
 * the rest of the input is still being uploaded and searched. */
static volatile size_t g_ready = 0;     /* leading positions of g_lstar that are final */
static size_t g_ready_seen = 0;         /* the consumer's last look at it */
static pthread_t g_worker, g_toucher;
static int g_worker_on = 0, g_toucher_on = 0;
static volatile int g_worker_done = 0;
static double g_landed_ms = 0.0;        /* prepare entered -> whole table landed */
static struct timespec g_t_enter;
static struct {
	int t, ngpus, variant;
} g_job;

static double ms_since(const struct timespec *t0)
{
	struct timespec t1;
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (t1.tv_sec - t0->tv_sec) * 1e3 + (t1.tv_nsec - t0->tv_nsec) * 1e-6;
}

static void die(const char *what)
{
	fprintf(stderr, "x3 search backend: %s\n", what);
	abort();
}

static void die(const char *what);

static void *worker_main(void *arg)
{
	(void)arg;
	int rc = x3s_search_host_stream(g_base, g_isize, g_forward_window, g_job.t, g_job.ngpus, g_job.variant, g_lstar, NULL,
	                                &g_ready);
	if (rc != X3S_OK) {
		die(x3s_last_error());
	}
	g_landed_ms = ms_since(&g_t_enter);
	__atomic_store_n(&g_worker_done, 1, __ATOMIC_RELEASE);
	return NULL;
}

/* the table is fresh memory: have the kernel map it (a copy from the device into unmapped pages spends
 * most of its time in page faults) while the input is uploaded and searched; nothing is written */
static void *toucher_main(void *arg)
{
	(void)arg;
	const size_t step = (size_t)8 << 20;
	for (size_t o = 0; o < g_isize; o += step) {
		const size_t len = g_isize - o < step ? g_isize - o : step;
		if (madvise(g_lstar + o, len, MADV_POPULATE_WRITE) != 0) {
			break; /* an older kernel: the copies fault the pages in themselves */
		}
	}
	return NULL;
}

void x3_search_wait(void)
{
	if (g_worker_on) {
		pthread_join(g_worker, NULL);
		g_worker_on = 0;
	}
	if (g_toucher_on) {
		pthread_join(g_toucher, NULL);
		g_toucher_on = 0;
	}
}

size_t x3_search_ready(void)
{
	return g_base == NULL ? 0 : __atomic_load_n(&g_ready, __ATOMIC_ACQUIRE);
}

double x3_search_landed_ms(void)
{
	return g_landed_ms;
}

void x3_search_release(void)
{
	x3_search_wait();
	if (g_registered != NULL) {
		x3s_host_unregister(g_registered);
		g_registered = NULL;
	}
	if (g_lstar_pinned) {
		x3s_host_free(g_lstar);
	} else if (g_lstar != NULL && getenv("X3_TABLE_KEEP") != NULL && g_spare == NULL) {
		g_spare = g_lstar; /* a host that prepares again and again skips mapping the table anew */
		g_spare_bytes = g_lstar_bytes;
	} else {
		free(g_lstar);
	}
	g_lstar_pinned = 0;
	g_lstar_bytes = 0;
	g_ready = 0;
	g_ready_seen = 0;
	g_worker_done = 0;
	free(g_table);
	g_lstar = NULL;
	g_table = NULL;
	g_base = NULL;
	g_isize = 0;
}

void x3_search_prepare(const char *base, size_t isize)
{
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);

	x3_search_release();
	g_t_enter = t0;
	g_landed_ms = 0.0;

	int ngpus = 0;
	const char *env = getenv("X3_GPUS");
	if (env != NULL) {
		ngpus = atoi(env);
	}
	int variant = X3S_KERNEL_DEFAULT;
	env = getenv("X3_SEARCH_KERNEL");
	if (env != NULL) {
		variant = atoi(env);
	}
	env = getenv("X3_SEARCH_TABLE");
	const int want_table = env != NULL && atoi(env) != 0;

	if (want_table) {
		g_table = malloc(isize > 0 ? isize * MAX_MATCH_LEN : 1);
		if (g_table == NULL) {
			die("out of memory");
		}
	}

	g_startup_ms = 0.0;
	if (isize > 0) {
		/* The first CUDA call of a process loads the driver and creates the context: between 0.2 s and
		 * several seconds depending on the machine, and nothing to do with the search.  Take it here,
		 * separately timed, so that a host can report it apart from the search proper. */
		struct timespec s0, s1;
		clock_gettime(CLOCK_MONOTONIC, &s0);
		if (x3s_device_count() > 0) {
			x3s_host_free(x3s_host_alloc(64));
		}
		clock_gettime(CLOCK_MONOTONIC, &s1);
		g_startup_ms = (s1.tv_sec - s0.tv_sec) * 1e3 + (s1.tv_nsec - s0.tv_nsec) * 1e-6;
	}
	/* X3_PREPARE_REGISTER=1 (measurement knob; never changes the table): the table lives in page-locked
	 * memory and the caller's buffer (the reference's malloc'ed iptr, x3.c:579) is page-locked where it
	 * lies while the tables exist.  Measured on the B200 box: cudaHostRegister of 212 MB costs 80-130 ms
	 * and cudaMallocHost of as much 140 ms -- more than the whole search -- so the default leaves both
	 * pageable and lets the device layer stage them through its own small page-locked ring. */
	g_register_ms = 0.0;
	if (isize > 0 && getenv("X3_PREPARE_REGISTER") != NULL && x3s_device_count() > 0) {
		struct timespec s0, s1;
		g_lstar = x3s_host_alloc(isize);
		g_lstar_pinned = g_lstar != NULL;
		clock_gettime(CLOCK_MONOTONIC, &s0);
		if (g_lstar_pinned && x3s_host_register((void *)base, isize + g_forward_window) == X3S_OK) {
			g_registered = (void *)base;
		}
		clock_gettime(CLOCK_MONOTONIC, &s1);
		g_register_ms = (s1.tv_sec - s0.tv_sec) * 1e3 + (s1.tv_nsec - s0.tv_nsec) * 1e-6;
	}
	int fresh = 0;
	if (g_lstar == NULL) {
		/* 2 MB aligned, so that the kernel may back it with huge pages */
		void *mem = NULL;
		const size_t bytes = ((isize > 0 ? isize : 1) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
		if (g_spare != NULL && g_spare_bytes >= bytes) {
			mem = g_spare;
			g_lstar_bytes = g_spare_bytes;
			g_spare = NULL;
		} else {
			free(g_spare);
			g_spare = NULL;
			if (posix_memalign(&mem, (size_t)2 << 20, bytes) != 0 || mem == NULL) {
				die("out of memory");
			}
			if (getenv("X3_TABLE_NO_THP") == NULL) {
				(void)madvise(mem, bytes, MADV_HUGEPAGE);
			}
			g_lstar_bytes = bytes;
			fresh = 1;
		}
		g_lstar = mem;
	}
	g_base = base;
	g_isize = isize;
	/* backend.c:76: the selection loop never runs for t <= 0 */
	g_job.t = g_max_match_count < 0 ? 0 : g_max_match_count;
	g_job.ngpus = ngpus;
	g_job.variant = variant;
	if (isize > 0 && !want_table && getenv("X3_PREPARE_SYNC") == NULL) {
		/* the search on its own thread; the caller goes on and find_best_match() waits where it must */
		if (fresh && isize >= ((size_t)4 << 20)) {
			g_toucher_on = pthread_create(&g_toucher, NULL, toucher_main, NULL) == 0;
		}
		if (pthread_create(&g_worker, NULL, worker_main, NULL) != 0) {
			die("cannot start the search thread");
		}
		g_worker_on = 1;
	} else if (isize > 0) {
		int rc = x3s_search_host(base, isize, g_forward_window, g_job.t, ngpus, variant, g_lstar, g_table, NULL);
		if (rc != X3S_OK) {
			die(x3s_last_error());
		}
		g_ready = isize;
		g_landed_ms = ms_since(&t0);
	}

	clock_gettime(CLOCK_MONOTONIC, &t1);
	g_prepare_ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
}

double x3_search_prepare_ms(void)
{
	return g_prepare_ms;
}

double x3_search_startup_ms(void)
{
	return g_startup_ms;
}

double x3_search_register_ms(void)
{
	return g_register_ms;
}

void x3_search_table(const uint8_t **H, const uint8_t **Lstar, size_t *n)
{
	x3_search_wait(); /* the whole table */
	if (H != NULL) {
		*H = g_table;
	}
	if (Lstar != NULL) {
		*Lstar = g_lstar;
	}
	if (n != NULL) {
		*n = g_isize;
	}
}

/* replaces reference backend.c:56-100 */
size_t find_best_match(char *p)
{
	if (g_base == NULL) {
		die("find_best_match() before x3_search_prepare()");
	}
	if (p < g_base || (size_t)(p - g_base) >= g_isize) {
		die("find_best_match(): pointer outside the prepared buffer");
	}

	x3_dict_find_fn find = g_dict_find != NULL ? g_dict_find : dict_find_match;
	x3_dict_len_fn len = g_dict_len != NULL ? g_dict_len : dict_get_len_by_index;

	/* Lstar = number of i with count[i] > tc*, the only tc the reference's loop
	 * nest returns from (count is non-increasing in i and i == 0 is never
	 * filtered).  0 stands for "return 1 without looking at the dictionary". */
	const size_t pos = (size_t)(p - g_base);
	if (pos >= g_ready_seen) {
		/* not landed the last time we looked: look again, and wait for the piece if it is still on its way */
		size_t r = __atomic_load_n(&g_ready, __ATOMIC_ACQUIRE);
		while (pos >= r) {
			sched_yield();
			r = __atomic_load_n(&g_ready, __ATOMIC_ACQUIRE);
		}
		g_ready_seen = r;
	}
	const int lstar = g_lstar[pos];

	for (int i = lstar - 1; i >= 0; --i) {
		/* backend.c:79-83 */
		if (i >= 2 && g_factor1 > 0) {
			if (find == NULL || len == NULL) {
				die("no dictionary queries available (link dict.c or call x3_backend_set_dict)");
			}
			const size_t m = find(p + i);
			if (m != (size_t)-1 && len(m) * g_factor1 > (size_t)(i + 1)) {
				continue;
			}
		}
		/* backend.c:84-90; int arithmetic as in the reference */
		if (i >= 1 && g_factor2 > 0) {
			if (find == NULL || len == NULL) {
				die("no dictionary queries available (link dict.c or call x3_backend_set_dict)");
			}
			int skip = 0;
			for (int o = 1; o <= i; ++o) {
				const size_t m = find(p + o);
				if (m != (size_t)-1 && ((int)len(m) - o) * (int)g_factor2 > i + 1) {
					skip = 1;
					break;
				}
			}
			if (skip) {
				continue;
			}
		}
		return (size_t)i + 1; /* backend.c:92 */
	}

	return 1; /* backend.c:99 */
}
