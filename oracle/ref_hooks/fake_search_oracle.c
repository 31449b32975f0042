/*
 * fake_search_oracle.c -- TEST INFRASTRUCTURE.  An oracle-backed stand-in for
 * the device layer (include/x3_search.h), used ONLY to prove on a CPU-only
 * machine that the histogram/filter split of the backend shim reproduces the
 * reference stream.  It is linked into oracle/_ref/x3_ref_*_cpuoracle and never
 * into the product library or the product x3 binary.
 */
#include "x3_search.h"
#include "x3_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int x3s_device_count(void) { return 0; }
const char *x3s_last_error(void) { return "fake_search_oracle"; }
const char *x3s_version(void) { return "oracle-backed fake (tests only)"; }
size_t x3s_required_bytes(size_t n, size_t W) { return n + W; }
void *x3s_host_alloc(size_t bytes) { return malloc(bytes); }
void x3s_host_free(void *p) { free(p); }
int x3s_host_register(void *p, size_t bytes) { (void)p; (void)bytes; return X3S_ERR_CUDA; }
int x3s_host_unregister(void *p) { (void)p; return X3S_ERR_CUDA; }
void x3s_release(void) {}
int x3s_set_devices(const int *ids, int count) { (void)ids; (void)count; return X3S_OK; }

int x3s_search_device(int device, const void *d_x, size_t n, size_t W, int t, void *d_lstar, void *d_H,
                      void *stream, int variant)
{
	(void)device; (void)d_x; (void)n; (void)W; (void)t; (void)d_lstar; (void)d_H; (void)stream; (void)variant;
	return X3S_ERR_CUDA;
}

int x3s_search_host(const void *x, size_t n, size_t W, int t, int ngpus, int variant, void *lstar, void *H,
                    x3s_timing *timing)
{
	(void)ngpus; (void)variant;
	if (t > X3S_MAX_T && H != NULL) {
		return X3S_ERR_UNSUPP;
	}
	x3o_table_fast((const uint8_t *)x, n + W, 0, n, W, t, (uint8_t *)H, NULL, (uint8_t *)lstar);
	if (timing != NULL) {
		memset(timing, 0, sizeof(*timing));
	}
	return X3S_OK;
}

/* the streamed form: the fake computes the table in pieces of 64 K positions and moves *ready on behind each,
 * so that the shim's waiting find_best_match() is exercised on a CPU-only machine too */
int x3s_search_host_stream(const void *x, size_t n, size_t W, int t, int ngpus, int variant, void *lstar,
                           x3s_timing *timing, volatile size_t *ready)
{
	(void)ngpus; (void)variant;
	if (ready == NULL) {
		return X3S_ERR_ARG;
	}
	__atomic_store_n(ready, (size_t)0, __ATOMIC_RELEASE);
	/* X3_FAKE_TABLE_FILE (host-pass profiling on a CPU-only machine): the table is kept in / taken from a file,
	 * so that repeated runs over the same input measure the sequential pass and not this oracle */
	const char *tf = getenv("X3_FAKE_TABLE_FILE");
	if (tf != NULL) {
		FILE *f = fopen(tf, "rb");
		if (f != NULL) {
			const size_t got = fread(lstar, 1, n, f);
			fclose(f);
			if (got == n) {
				__atomic_store_n(ready, n, __ATOMIC_RELEASE);
				if (timing != NULL) {
					memset(timing, 0, sizeof(*timing));
				}
				return X3S_OK;
			}
		}
	}
	const size_t step = (size_t)1 << 16;
	for (size_t p0 = 0; p0 < n; p0 += step) {
		const size_t p1 = n - p0 < step ? n : p0 + step;
		x3o_table_fast((const uint8_t *)x, n + W, p0, p1, W, t, NULL, NULL, (uint8_t *)lstar + p0);
		__atomic_store_n(ready, p1, __ATOMIC_RELEASE);
	}
	__atomic_store_n(ready, n, __ATOMIC_RELEASE);
	if (tf != NULL) {
		FILE *f = fopen(tf, "wb");
		if (f != NULL) {
			fwrite(lstar, 1, n, f);
			fclose(f);
		}
	}
	if (timing != NULL) {
		memset(timing, 0, sizeof(*timing));
	}
	return X3S_OK;
}

size_t x3s_part_positions(size_t W) { (void)W; return 0; }
int x3s_default_kernel(size_t W, int t, int want_table) { (void)W; (void)t; (void)want_table; return X3S_KERNEL_DEFAULT; }
