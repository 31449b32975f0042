/*
 * x3_search.h -- device-level C ABI of the B200 forward-window match search.
 *
 * This is the FFI surface a host program binds (C, cgo, ctypes ...): plain
 * pointers and sizes, no C++ or torch types.  It computes, for every position p
 * of a zero-padded input buffer, what the histogram loop of the reference
 * find_best_match() computes on its stack (reference backend.c:58-74) and the
 * threshold selection that follows it (reference backend.c:76-78,92,99):
 *
 *   H[p][i]  = min(255, #{ d in [1, W-33] : x[p..p+i] == x[p+d..p+d+i] }), i = 0..31
 *   Lstar[p] = 0 if t <= 0 or H[p][0] < 2, else #{ i : H[p][i] > min(t, H[p][0]-1) }
 *
 * The dictionary-dependent filter (reference backend.c:79-90) is NOT here; it is
 * applied on the host by find_best_match() in x3_backend.h.
 *
 * All functions return X3S_OK (0) or a negative error code; x3s_last_error()
 * describes the last failure of the calling thread.  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with X3S_ERR_CUDA.
 */
#ifndef X3_SEARCH_H
#define X3_SEARCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X3S_OK            0
#define X3S_ERR_CUDA     (-1) /* CUDA runtime / driver failure, or no device */
#define X3S_ERR_ARG      (-2) /* bad argument */
#define X3S_ERR_UNSUPP   (-3) /* parameter outside what the kernels implement */

#define X3S_MAX_MATCH_LEN 32  /* reference backend.h:7-10 */
#define X3S_MAX_T        254  /* 32-bin table H and the brute-force kernels: u8 cells saturate at 255, lossless
                              * while t <= 254.  The Lstar-only rank search takes any t (reference backend.c:21-26) */

/* kernel variants */
#define X3S_KERNEL_DEFAULT   0 /* production choice: for Lstar alone the segment search (W <= 16 KiB, t >= 5) or the
                                * rank search; the stream kernel when the 32-bin table H is requested (or W - 33 > 2^23) */
#define X3S_KERNEL_NAIVE     1 /* one thread per position, byte loop; cross-check only */
#define X3S_KERNEL_BITSLICED 2 /* first bit-sliced version (thread-private u8 histograms); kept for comparison */
#define X3S_KERNEL_STREAM    3 /* brute-force pair-test kernel (bit-sliced, ALU bound); fast path when t <= 15 and no H */
#define X3S_KERNEL_STREAM_FULL 4 /* stream kernel, u8 counters forced (what H != NULL or t > 15 selects) */
#define X3S_KERNEL_RANK      5 /* occurrence-rank search: sorts positions by L-gram level by level; Lstar only
                                * (d_H must be NULL), cost independent of W and t */
#define X3S_KERNEL_SEG       6 /* segment search: the rank formulation with a whole segment (positions + window)
                                * resident in shared memory, one launch; Lstar only, W <= 16 KiB, t >= 5 */

typedef struct x3s_timing {
	double h2d_ms;    /* host -> device copies (max over GPUs) */
	double kernel_ms; /* search kernels (max over GPUs), CUDA events */
	double d2h_ms;    /* device -> host copies of Lstar (and H) */
	                  /* A shard that is pipelined chunk by chunk (see x3s_search_host) overlaps the three;
	                   * the figures then split its critical path: h2d_ms = until the first search can
	                   * start, kernel_ms = from there to the end of the last search, d2h_ms = from there
	                   * to the last byte back. */
	double total_ms;  /* wall time of the whole call */
	int    gpus;      /* GPUs actually used */
	int    launches;  /* kernel launches issued */
} x3s_timing;

/* Number of visible CUDA devices (0 when there is none or no driver). */
int x3s_device_count(void);

const char *x3s_last_error(void);

/* Library / build description, e.g. "x3-b200 search sm_100a". */
const char *x3s_version(void);

/*
 * Bytes of device memory that must be readable behind d_x for a search over
 * n_positions with window W: the data, the W bytes of padding the reference
 * allocates (reference x3.c:579,590) and the tile round-up slack of the kernels.
 * Bytes at index >= n_positions + W are read but never influence a result.
 */
size_t x3s_required_bytes(size_t n_positions, size_t W);

/*
 * Search over a buffer that is already resident on `device`.
 *   d_x      device pointer, 16-byte aligned, x3s_required_bytes() readable
 *   d_lstar  device pointer, n_positions bytes (may not be NULL)
 *   d_H      device pointer, n_positions*32 bytes, or NULL (production mode)
 *   stream   cudaStream_t as void* (NULL = default stream).  The work is queued on it and the
 *            results are ordered behind it; the brute-force kernels return at once, the rank
 *            search returns when its last level has been queued (it reads level sizes back
 *            while queueing, so the call lasts about as long as the search)
 * One search per device at a time: the library keeps its scratch buffers per device
 * (searches issued from different streams are ordered behind each other; do not call this
 * concurrently from several host threads for the same device).  Inside one search the chunks of a
 * large input run on lane streams of the library's own, forked from `stream` and joined back into
 * it before the call returns, so the ordering seen by the caller is that of a single stream.
 */
int x3s_search_device(int device, const void *d_x, size_t n_positions, size_t W, int t,
                      void *d_lstar, void *d_H, void *stream, int variant);

/*
 * Search over a host buffer x[0 .. n+W) (n data bytes then W zero bytes, the
 * layout of the reference's iptr, x3.c:579-591).  Positions are partitioned into
 * `ngpus` contiguous ranges, each shipped with its trailing window halo; results
 * land in lstar[n] (and H[n*32] when H != NULL).  ngpus <= 0 means "all visible".
 * Synchronous.  Device and staging buffers are cached between calls.
 *
 * A shard larger than one chunk of the rank search (2^24 - W positions) is searched as several
 * chunks in flight at once (up to 4 by default, X3_RANK_LANES=1..8 overrides; never changes the
 * result).  When x and lstar are page-locked (x3s_host_alloc, or cudaHostRegister by the caller)
 * the shard is also pipelined: the upload goes chunk by chunk, a chunk is searched as soon as the
 * bytes it reads have arrived, and its Lstar is copied back while later chunks are still searched.
 * Pageable buffers take the plain upload - search - copy back sequence.
 */
int x3s_search_host(const void *x, size_t n, size_t W, int t, int ngpus, int variant,
                    void *lstar, void *H, x3s_timing *timing);

/*
 * ONE input searched by several processes / GPUs (one rank per GPU under torchrun, say): the positions
 * [0, n) are cut into pieces of x3s_part_positions(W) positions (a whole number of the segment search's
 * segments; 0 when the window does not fit the segment search), and part `part` of `parts` takes the pieces
 * part, part + parts, ...: contiguous position ranges, each read with the window behind it, dealt out in
 * turn so that every part sees every region of the input.  x, lstar (host) and d_x, d_lstar (device) are
 * the buffers of the WHOLE input in both calls: n + W readable bytes (x3s_required_bytes(n, W) on the
 * device) and n table bytes, of which a part reads and writes its own pieces only (plus, reading, the W
 * bytes behind each).  The parts may run in different processes over shared host buffers.  Lstar only,
 * segment search only (X3S_ERR_UNSUPP otherwise).  x3s_search_host_part runs on the first device of
 * x3s_set_devices() (default device 0).
 */
size_t x3s_part_positions(size_t W);
int x3s_search_device_part(int device, const void *d_x, size_t n_positions, size_t W, int t, void *d_lstar,
                           void *stream, int part, int parts);
int x3s_search_host_part(const void *x, size_t n, size_t W, int t, void *lstar, x3s_timing *timing, int part, int parts);

/*
 * x3s_search_host() for a consumer that reads the table from left to right while it is still being
 * made (the reference's compress() visits positions in increasing order, x3.c:379): the call itself
 * returns when the whole table has landed, but while it runs -- call it from a thread of its own --
 * *ready is kept at the number of leading positions whose Lstar is final in `lstar` (stored with
 * release order; it only grows, and ends at n).  Lstar only.  With more than one GPU the prefix moves
 * on through the shards in position order.
 */
int x3s_search_host_stream(const void *x, size_t n, size_t W, int t, int ngpus, int variant, void *lstar,
                           x3s_timing *timing, volatile size_t *ready);

/*
 * Restricts x3s_search_host() to the given CUDA device ordinals, in this order
 * (count <= 0 restores "devices 0 .. ngpus-1").  One process per GPU launchers
 * (torchrun) call it with their LOCAL_RANK.
 */
int x3s_set_devices(const int *ids, int count);

/* Pinned host memory helpers (so FFI callers can avoid pageable copies). */
void *x3s_host_alloc(size_t bytes);
void  x3s_host_free(void *p);

/*
 * Page-locks memory the caller owns (malloc, mmap, a shared-memory segment ...) where it lies, so
 * that x3s_search_host() can pipeline a shard between it and the device; the range is rounded out
 * to whole pages.  x3s_host_unregister() takes the same pointer.  X3S_ERR_CUDA when the driver
 * refuses (the buffer then still works, as pageable memory).
 */
int x3s_host_register(void *p, size_t bytes);
int x3s_host_unregister(void *p);

/*
 * Which kernel X3S_KERNEL_DEFAULT runs for window W and max match count t (no device needed):
 * X3S_KERNEL_SEG, X3S_KERNEL_RANK or X3S_KERNEL_STREAM (want_table != 0: the 32-bin table is asked for).
 */
int x3s_default_kernel(size_t W, int t, int want_table);

/*
 * Measurement hook of the rank search.  With the environment variable X3_RANK_PROFILE=1 the
 * library brackets every launch of a rank search with CUDA events (and synchronises at the
 * end of the search); this call returns, for the last such search on `device`, the device
 * time, the number of 8-byte elements the launches were sized for and the launch count of
 * one kernel family: kind 0 = radix passes (x3_rank_radix_kernel), 1 = level kernels
 * (x3_rank_level_kernel), 2 = set-up.  Returns X3S_ERR_ARG for a bad device/kind.
 */
int x3s_rank_profile(int device, int kind, double *ms, double *elements, int *launches);

/*
 * How a rank search over n_positions with window W is cut up (no device needed): the positions are
 * searched in `chunks` chunks of `chunk_positions` positions each (a multiple of 4096; the last one
 * may be shorter; chunk + W - 33 < 2^24, so that ranks and chunk-relative positions fit 24 bits),
 * `lanes_used` of them in flight at once.  lanes = 0 asks for the library's own choice (one lane per
 * chunk, at most 4, X3_RANK_LANES overrides), lanes >= 1 for that many (capped at 8).  Each lane owns
 * 16 B of device scratch per chunk element.  X3S_ERR_UNSUPP when W - 33 > 2^23 (the brute-force
 * kernels take such windows).
 */
int x3s_rank_plan(size_t n_positions, size_t W, int lanes, size_t *chunk_positions, size_t *chunks, int *lanes_used);

/* Frees every cached device/staging buffer. */
void x3s_release(void);

#ifdef __cplusplus
}
#endif
#endif /* X3_SEARCH_H */
