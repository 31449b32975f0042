/*
 * x3_backend.h -- drop-in replacement for the reference's match-search
 * translation unit (reference backend.h:20-31 / backend.c).
 *
 * The reference has no plugin mechanism: backend.c is linked by name
 * (reference Makefile:33), so the swappable unit is "the object that defines
 * these nine symbols".  This header declares exactly those symbols, with the
 * reference's signatures and defaults, plus the three new entry points the GPU
 * search needs.  Host code stays C99.
 *
 * Error convention follows the reference (reference x3.c:552-560,582-588):
 * a message on stderr and abort().  There is no CPU fallback.
 */
#ifndef X3_BACKEND_H
#define X3_BACKEND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference backend.h:7,10 */
#ifndef MATCH_LOGSIZE
#define MATCH_LOGSIZE 5
#endif
#ifndef MAX_MATCH_LEN
#define MAX_MATCH_LEN (1 << MATCH_LOGSIZE)
#endif

/*
 * Replaces reference backend.c:56-100.  p must point into the buffer given to
 * x3_search_prepare().  Returns Lstar[p - base] refined by the dictionary filter
 * of reference backend.c:79-90, evaluated against the live dictionary through
 * dict_find_match()/dict_get_len_by_index() (reference dict.c:105-130,159-162),
 * which the host program provides -- either as ordinary external symbols or via
 * x3_backend_set_dict().
 */
size_t find_best_match(char *p);

/* reference backend.c:8-19; default 8192 */
void set_forward_window(size_t n);
size_t get_forward_window(void);

/* reference backend.c:21-31; default 15 */
void set_max_match_count(int n);
int get_max_match_count(void);

/* reference backend.c:33-54; defaults 4 and 0 */
size_t get_magic_factor1(void);
void set_magic_factor1(size_t factor);
size_t get_magic_factor2(void);
void set_magic_factor2(size_t factor);

/*
 * NEW.  Called once by the host after the input has been loaded (the point right
 * after fload() at reference x3.c:591) and after every setter has run.  base is
 * the padded input buffer: isize data bytes followed by get_forward_window()
 * zero bytes (reference x3.c:579,590).  Uploads it, runs the search on
 * X3_GPUS GPUs (environment variable; default: all visible), and keeps Lstar in
 * host memory until x3_search_release().  Aborts on any CUDA failure.
 */
void x3_search_prepare(const char *base, size_t isize);

/*
 * NEW.  x3_search_prepare() starts the search on a thread of its own and returns; the table lands
 * piece by piece from the left and find_best_match(p) waits until p has landed, so a host that reads
 * positions in increasing order (reference x3.c:379) starts on the first piece.  (X3_PREPARE_SYNC=1 in the
 * environment, or a request for the full table, makes prepare wait for the whole table itself.)
 * x3_search_wait() blocks until the whole table has landed, x3_search_ready() says how many leading
 * positions have, x3_search_landed_ms() how long it took from entering prepare to the last one.
 */
void x3_search_wait(void);
size_t x3_search_ready(void);
double x3_search_landed_ms(void);

/* NEW.  Drops the tables of the last x3_search_prepare() (waits for a search still running). */
void x3_search_release(void);

/*
 * NEW, test-only.  Exposes the tables of the last prepare.  *H is NULL unless
 * the environment variable X3_SEARCH_TABLE=1 asked for the full 32-bin table.
 */
void x3_search_table(const uint8_t **H, const uint8_t **Lstar, size_t *n);

/* NEW.  Milliseconds spent inside the last x3_search_prepare() (wall clock),
 * so a host can add it to the reference's "elapsed time" (x3.c:597-601). */
double x3_search_prepare_ms(void);

/* NEW.  The part of that which was the process's one-off CUDA start-up (driver load and first
 * context; 0 when a previous call had already paid it). */
double x3_search_startup_ms(void);

/* NEW.  The part that page-locked the caller's buffer in place (so that upload, search and copy
 * back can overlap chunk by chunk); the buffer is unlocked again by x3_search_release(). */
double x3_search_register_ms(void);

/*
 * NEW.  Dictionary callbacks for hosts that cannot export dict_find_match /
 * dict_get_len_by_index as dynamic symbols (e.g. FFI hosts).  Passing NULLs
 * restores symbol lookup.
 */
typedef size_t (*x3_dict_find_fn)(const char *p);
typedef size_t (*x3_dict_len_fn)(size_t index);
void x3_backend_set_dict(x3_dict_find_fn find, x3_dict_len_fn len);

#ifdef __cplusplus
}
#endif
#endif /* X3_BACKEND_H */
