/* Where does the first x3s_search_host() of a process spend its time?  (test utility) */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "x3_search.h"
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }
int main(int argc, char **argv)
{
	size_t n = argc > 1 ? (size_t)atol(argv[1]) : 1000000, W = 8192;
	unsigned char *x = calloc(n + W + 64, 1), *l = malloc(n);
	for (size_t i = 0; i < n; ++i) x[i] = (unsigned char)((i * 2654435761u) >> 24) & 31;
	double t0 = now();
	int nd = x3s_device_count();
	double t1 = now();
	printf("device_count=%d: %.3f s\n", nd, t1 - t0);
	int reps = argc > 2 ? atoi(argv[2]) : 3;
	for (int rep = 0; rep < reps; ++rep) {
		x3s_timing tm;
		double a = now();
		int rc = x3s_search_host(x, n, W, 15, 1, 0, l, NULL, &tm);
		double b = now();
		printf("search_host rep %d rc=%d: wall %.3f s (h2d %.3f ms kernel %.3f ms d2h %.3f ms total %.3f ms)\n", rep, rc,
		       b - a, tm.h2d_ms, tm.kernel_ms, tm.d2h_ms, tm.total_ms);
	}
	return 0;
}
