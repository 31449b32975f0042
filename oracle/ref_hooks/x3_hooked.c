/*
 * x3_hooked.c -- TEST INFRASTRUCTURE.  Compiles the UNMODIFIED reference x3.c
 * (taken from -I<reference dir>, never copied) with the one-line binding a
 * maintainer would add after fload() at reference x3.c:591:
 *
 *     x3_search_prepare(iptr, isize);
 *
 * The binding is injected with a function-like macro so the reference source
 * stays untouched.  file.h is included first so that its prototype of fload()
 * is not rewritten by the macro (its include guard skips the second inclusion).
 *
 * With -DX3_HOOK_CHECK the nine backend.h symbols called by x3.c are redirected
 * to chk_* wrappers (x3_check.c) that run the reference backend and the new
 * backend side by side on every call.
 */
#define _POSIX_C_SOURCE 2 /* as reference x3.c:1, must precede the first libc header */
#include <stddef.h>
#include <stdio.h>
#include "file.h"

void x3_search_prepare(const char *base, size_t isize);

#define fload(ptr, size, stream)                                        \
	do {                                                                \
		(fload)((ptr), (size), (stream));                               \
		if (mode == COMPRESS) {                                         \
			x3_search_prepare((const char *)(ptr), (size));             \
		}                                                               \
	} while (0)

#ifdef X3_HOOK_CHECK
#define find_best_match chk_find_best_match
#define set_forward_window chk_set_forward_window
#define get_forward_window chk_get_forward_window
#define set_max_match_count chk_set_max_match_count
#define get_max_match_count chk_get_max_match_count
#define set_magic_factor1 chk_set_magic_factor1
#define get_magic_factor1 chk_get_magic_factor1
#define set_magic_factor2 chk_set_magic_factor2
#define get_magic_factor2 chk_get_magic_factor2
#endif

#include "x3.c"
