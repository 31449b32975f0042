/*
 * x3_check.c -- TEST INFRASTRUCTURE.  Interposition harness:
 * the reference compress() drives BOTH backends against the same live
 * dictionary; every find_best_match() call is compared.  The stream is encoded
 * from the reference's value, so the output equals the reference's output; the
 * verdict goes to stderr and to the exit status via atexit().
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

/* reference backend under ref_* names (ref_backend_renamed.c) */
size_t ref_find_best_match(char *p);
void ref_set_forward_window(size_t n);
size_t ref_get_forward_window(void);
void ref_set_max_match_count(int n);
int ref_get_max_match_count(void);
size_t ref_get_magic_factor1(void);
void ref_set_magic_factor1(size_t f);
size_t ref_get_magic_factor2(void);
void ref_set_magic_factor2(size_t f);

/* new backend (libx3b200.so, include/x3_backend.h) */
size_t find_best_match(char *p);
void set_forward_window(size_t n);
void set_max_match_count(int n);
void set_magic_factor1(size_t f);
void set_magic_factor2(size_t f);

static unsigned long long g_calls = 0, g_mismatch = 0;
static int g_registered = 0;

static void report(void)
{
	fprintf(stderr, "x3_check: calls %llu mismatches %llu\n", g_calls, g_mismatch);
	if (g_mismatch != 0) {
		_Exit(3);
	}
}

size_t chk_find_best_match(char *p)
{
	if (!g_registered) {
		atexit(report);
		g_registered = 1;
	}
	size_t r = ref_find_best_match(p);
	size_t n = find_best_match(p);
	g_calls++;
	if (r != n) {
		if (g_mismatch < 10) {
			fprintf(stderr, "x3_check: MISMATCH at call %llu: reference %zu new %zu\n", g_calls, r, n);
		}
		g_mismatch++;
	}
	return r;
}

void chk_set_forward_window(size_t n) { ref_set_forward_window(n); set_forward_window(n); }
size_t chk_get_forward_window(void) { return ref_get_forward_window(); }
void chk_set_max_match_count(int n) { ref_set_max_match_count(n); set_max_match_count(n); }
int chk_get_max_match_count(void) { return ref_get_max_match_count(); }
size_t chk_get_magic_factor1(void) { return ref_get_magic_factor1(); }
void chk_set_magic_factor1(size_t f) { ref_set_magic_factor1(f); set_magic_factor1(f); }
size_t chk_get_magic_factor2(void) { return ref_get_magic_factor2(); }
void chk_set_magic_factor2(size_t f) { ref_set_magic_factor2(f); set_magic_factor2(f); }
