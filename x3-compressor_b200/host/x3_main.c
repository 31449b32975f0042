/*
 * x3_main.c -- the x3 command line over the B200 match search.
 *
 * Same interface as the reference's main() (reference x3.c:460-702): options
 * "zdfkht:w:m:n:x", 0/1/2 file arguments, ".x3" suffix handling, refusal to
 * overwrite without -f, banner and report on stderr.  The help text, the switch over
 * the file arguments and the report's printf lines follow x3.c:465-548,662-693 closely
 * ON PURPOSE: those strings and that format are the drop-in contract of the CLI
 * (tests/test_host_x3.py: test_report_matches_reference, test_cli_behaviour).  Compression calls the GPU
 * search once (x3_search_prepare, the hook of INTEGRATION.md section 1) and then
 * runs the sequential pass of x3_codec.c; the stream is byte-identical to the
 * reference's for the same input and flags.  Decompression needs no GPU.
 * Errors: message on stderr + abort(), as in the reference.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "x3_backend.h"
#include "x3_host.h"

enum { COMPRESS, DECOMPRESS };

static void print_help(char *path)
{
	fprintf(stderr, "Usage :\n\t%s [arguments] [input-file] [output-file]\n\n", path);
	fprintf(stderr, "Arguments :\n");
	fprintf(stderr, " -d     : force decompression\n");
	fprintf(stderr, " -z     : force compression\n");
	fprintf(stderr, " -f     : overwrite existing output file\n");
	fprintf(stderr, " -k     : keep (don't delete) input file (default)\n");
	fprintf(stderr, " -h     : print this message\n");
	fprintf(stderr, " -t NUM : maximum number of matches (affects compression ratio and speed)\n");
	fprintf(stderr, " -w NUM : window size (in kilobytes, affects compression ratio and speed)\n");
	fprintf(stderr, " -m NUM : magic factor (affects compression ratio and speed)\n");
}

/* reference file.c:21-55 */
static size_t stream_size(FILE *stream)
{
	long begin = ftell(stream);
	if (begin == (long)-1) {
		fprintf(stderr, "Stream is not seekable\n");
		abort();
	}
	if (fseek(stream, 0, SEEK_END)) {
		abort();
	}
	long end = ftell(stream);
	if (end == (long)-1) {
		abort();
	}
	if (fseek(stream, begin, SEEK_SET)) {
		abort();
	}
	return (size_t)end - (size_t)begin;
}

static FILE *force_fopen(const char *pathname, const char *mode, int force)
{
	if (force == 0 && access(pathname, F_OK) != -1) {
		fprintf(stderr, "File already exists\n");
		abort();
	}
	return fopen(pathname, mode);
}

static double now_s(void)
{
	struct timespec t;
	clock_gettime(CLOCK_REALTIME, &t);
	return (double)t.tv_sec + (double)t.tv_nsec * 1e-9;
}

int main(int argc, char *argv[])
{
	int mode = COMPRESS;
	int force = 0;
	int nl = 0;
	int c;

	while ((c = getopt(argc, argv, "zdfkht:w:m:n:x")) != -1) {
		switch (c) {
			case 'z': mode = COMPRESS; break;
			case 'd': mode = DECOMPRESS; break;
			case 'f': force = 1; break;
			case 'k': break;
			case 'h': print_help(argv[0]); return 0;
			case 't': set_max_match_count(atoi(optarg)); break;
			case 'w': set_forward_window(atoi(optarg) * 1024); break; /* int arithmetic, x3.c:503 */
			case 'm': set_magic_factor1(atoi(optarg)); break;
			case 'n': set_magic_factor2(atoi(optarg)); break;
			case 'x': nl = 1; break;
			default: abort();
		}
	}

	FILE *istream = NULL, *ostream = NULL;
	switch (argc - optind) {
		case 0:
			istream = stdin;
			ostream = stdout;
			break;
		case 1:
			istream = fopen(argv[optind], "r");
			if (mode == COMPRESS) {
				char path[4096];
				snprintf(path, sizeof(path), "%s.x3", argv[optind]);
				ostream = force_fopen(path, "w", force);
			} else {
				if (strrchr(argv[optind], '.') != NULL) {
					*strrchr(argv[optind], '.') = 0; /* remove suffix */
				}
				ostream = force_fopen(argv[optind], "w", force);
			}
			break;
		case 2:
			istream = fopen(argv[optind + 0], "r");
			ostream = force_fopen(argv[optind + 1], "w", force);
			break;
		default:
			fprintf(stderr, "Unexpected argument\n");
			abort();
	}

	fprintf(stderr, "%s\n", mode == COMPRESS ? "Compressing..." : "Decompressing...");
	if (istream == NULL) {
		fprintf(stderr, "Cannot open input file\n");
		abort();
	}
	if (ostream == NULL) {
		fprintf(stderr, "Cannot open output file\n");
		abort();
	}

	struct x3_codec *codec = x3_codec_create();
	x3_codec_set_nl(codec, nl);
	size_t size, asize;

	if (mode == COMPRESS) {
		fprintf(stderr, "max match count: %i\n", get_max_match_count());
		fprintf(stderr, "forward window: %zu\n", get_forward_window());
		fprintf(stderr, "magic factor 1: %zu\n", get_magic_factor1());
		fprintf(stderr, "magic factor 2: %zu\n", get_magic_factor2());

		const size_t isize = stream_size(istream);
		const size_t pad = get_forward_window() + 64;
		char *iptr = malloc(isize + pad);
		if (iptr == NULL) {
			abort();
		}
		memset(iptr + isize, 0, pad); /* x3.c:590 */
		if (fread(iptr, 1, isize, istream) < isize) {
			abort();
		}

		x3_backend_set_dict(x3_codec_dict_find, x3_codec_dict_len);
		const double t0 = now_s();
		x3_search_prepare(iptr, isize); /* GPU: every position at once; the table lands piece by piece from the left */
		const double t1 = now_s();
		void *optr = x3_compress(codec, iptr, isize, find_best_match, &asize); /* starts on the first piece */
		const double t2 = now_s();
		x3_search_wait();
		fprintf(stderr, "elapsed time: %f\n", (float)(t2 - t0));
		fprintf(stderr, "of which match search (GPU, incl. transfers): %f\n", (float)(x3_search_landed_ms() * 1e-3));
		fprintf(stderr, "of which before the sequential pass could start: %f\n", (float)(t1 - t0));
		fprintf(stderr, "of which CUDA start-up (driver + first context, once per process): %f\n",
		        (float)(x3_search_startup_ms() * 1e-3));
		x3_search_release();

		size = isize;
		if (fwrite(optr, 1, asize, ostream) < asize) {
			abort();
		}
		free(iptr);
		free(optr);
	} else {
		const size_t isize = stream_size(istream);
		asize = isize;
		char *iptr = malloc(isize + 8);
		if (iptr == NULL) {
			abort();
		}
		if (fread(iptr, 1, isize, istream) < isize) {
			abort();
		}
		const double t0 = now_s();
		void *optr = x3_decompress(codec, iptr, isize, &size);
		fprintf(stderr, "elapsed time: %f\n", (float)(now_s() - t0));
		if (fwrite(optr, 1, size, ostream) < size) {
			abort();
		}
		free(iptr);
		free(optr);
	}

	const struct x3_stats st = *x3_codec_stats(codec);
	x3_codec_destroy(codec);
	fclose(istream);
	fclose(ostream);

	/* report, x3.c:662-693 */
	const size_t *events = st.events;
	const float *sizes = st.sizes;
	size_t dict_hit_count = events[X3_E_CTX0] + events[X3_E_CTX1] + events[X3_E_IDX1];
	size_t stream_size_dict = (size_t)ceil(sizes[X3_E_CTX0] + sizes[X3_E_CTX1] + sizes[X3_E_IDX1]);
	size_t stream_size_all = (size_t)ceil(sizes[X3_E_CTX0] + sizes[X3_E_CTX1] + sizes[X3_E_IDX1] + sizes[X3_E_NEW]);

	fprintf(stderr, "input stream size: %zu\n", size);
	fprintf(stderr, "output stream size: %zu\n", (stream_size_all + 7) / 8);
	fprintf(stderr, "dictionary: hit %zu, miss %zu\n", dict_hit_count, events[X3_E_NEW]);
	fprintf(stderr, "codestream size: dictionary %zu / %f%%, new fragment %zu / %f%%\n",
	        (stream_size_dict + 7) / 8, 100.f * stream_size_dict / stream_size_all,
	        ((size_t)ceil(sizes[X3_E_NEW]) + 7) / 8, 100.f * (size_t)ceil(sizes[X3_E_NEW]) / stream_size_all);
	fprintf(stderr, "\x1b[37;1mest. compression ratio: %f\x1b[0m\n", size / (float)((stream_size_all + 7) / 8));
	fprintf(stderr, "\x1b[37;1mreal compression ratio: %f\x1b[0m\n", size / (float)asize);
	fprintf(stderr, "number of events: ctx0 %zu, ctx1 %zu, miss1 %zu, new %zu\n", events[X3_E_CTX0],
	        events[X3_E_CTX1], events[X3_E_IDX1], events[X3_E_NEW]);
	fprintf(stderr, "event sizes: ctx0 %f%%, ctx1 %f%%, miss1 %f%%, new %f%%\n",
	        100.f * (size_t)ceil(sizes[X3_E_CTX0]) / stream_size_all,
	        100.f * (size_t)ceil(sizes[X3_E_CTX1]) / stream_size_all,
	        100.f * (size_t)ceil(sizes[X3_E_IDX1]) / stream_size_all,
	        100.f * (size_t)ceil(sizes[X3_E_NEW]) / stream_size_all);
	fprintf(stderr, "context entries: ctx0 %zu, ctx1 %zu\n", st.ctx0_entries, st.ctx1_entries);
	return 0;
}
