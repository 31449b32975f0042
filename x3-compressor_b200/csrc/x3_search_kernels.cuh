/*
 * x3_search_kernels.cuh -- internal interface between the C-ABI layer
 * (x3_search_api.cu) and the sm_100a kernels (x3_search_kernels.cu).
 */
#ifndef X3_SEARCH_KERNELS_CUH
#define X3_SEARCH_KERNELS_CUH

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct X3SearchParams {
	const uint8_t *x;      /* padded input on the device, 16-byte aligned */
	unsigned long long n;  /* positions to search: [0, n) */
	uint32_t D;            /* number of distances: W > 33 ? W - 33 : 0 (backend.c:66) */
	int t;                 /* g_max_match_count (backend.c:21) */
	uint8_t *lstar;        /* n bytes */
	uint8_t *H;            /* n*32 bytes or NULL */
	/* stream kernel only (owned by the API layer, one set per device) */
	unsigned int *tile_counter; /* dynamic tile scheduler, zeroed before each launch */
	uint8_t *deep;              /* X3K_DEEP_BYTES_PER_CTA bytes per resident CTA */
	unsigned int ntiles;
	int kd;                     /* dense levels: 2, 3, or 0 = choose per tile from a probe */
	/* segment search only: this launch takes pieces part, part + parts, ... of the positions [0, n), a piece
	 * being piece_segments segments (0 = all of [0, n): the plain search) */
	uint32_t part = 0, parts = 1, piece_segments = 0;
	uint32_t piece_first = 0, piece_count = 0; /* of the part's own pieces: the launch takes [first, first + count) (count 0 = all) */
};

/* Scratch the stream kernel needs: the grid it will be launched with and the
 * bytes of `deep` histogram rows behind it (64 B per position of a resident tile;
 * the rows must be ZERO when a launch starts and the kernel hands them back zeroed). */
#define X3K_STREAM_TILE 1984
#define X3K_DEEP_BYTES_PER_CTA ((size_t)X3K_STREAM_TILE * 64)
int x3k_stream_grid(unsigned long long n);
int x3k_stream_max_grid(void); /* valid after x3k_init_device() */

/* Worst-case bytes a kernel reads behind x for n positions and window W. */
size_t x3k_required_bytes(size_t n, size_t W);

/* Launches the chosen variant on `stream`.  Returns the CUDA error of the launch. */
cudaError_t x3k_launch(int variant, const X3SearchParams &prm, cudaStream_t stream, int *launches);

/* Rank search (x3_search_rank.cu): Lstar only, any t <= 254, D <= x3k_rank_max_distances().
 * Issues its launches on `stream` (and on lane streams forked from and joined back into it) but
 * returns only after the last level has reported its size (one small zero-copy report per level).
 * cudaErrorNotSupported when H is requested. */
cudaError_t x3k_launch_rank(const X3SearchParams &prm, cudaStream_t stream, int *launches);

/* The same with the lanes spelled out: positions [0, n) are searched chunk by chunk
 * (x3k_rank_chunking), each chunk on the next free of `lanes` streams of the caller's (no
 * fork/join), one host thread keeping all of them fed.  before_chunk is called before the first
 * launch of a chunk is queued on its lane's stream (the API layer makes the stream wait for the
 * upload of the bytes the chunk reads), after_chunk as soon as its last kernel is queued (the API
 * layer queues the copy back of the chunk's Lstar there, while other chunks are still searched). */
struct X3RankBatch {
	const uint8_t *x;      /* device, 16-byte aligned, x3k_required_bytes(n, W) readable */
	uint8_t *lstar;        /* n bytes */
	unsigned long long n;
	int lanes;             /* 1 .. x3k_rank_max_lanes() */
	cudaStream_t streams[8];
	cudaError_t (*before_chunk)(void *ctx, int lane, unsigned long long a0, unsigned long long len);
	cudaError_t (*after_chunk)(void *ctx, int lane, unsigned long long a0, unsigned long long len);
	void *ctx;
};
cudaError_t x3k_launch_rank_batch(const X3RankBatch &b, uint32_t D, int t, int *launches);
void x3k_rank_chunking(unsigned long long n, uint32_t D, int lanes, unsigned long long *chunk, unsigned long long *count);
int x3k_rank_max_lanes(void);
int x3k_rank_default_lanes(unsigned long long n, uint32_t D); /* X3_RANK_LANES, else one per chunk */
uint32_t x3k_rank_max_distances(void);
void x3k_rank_release(int device);
int x3k_rank_profile(int device, int kind, double *ms, double *elements, int *launches);

/* Segment search (x3_search_seg.cu): Lstar only, windows that fit on chip (1 <= D <=
 * x3k_seg_max_distances()), t >= x3k_seg_min_t(); one launch, returns at once.  Needs
 * prm.tile_counter (4 bytes of device memory, zeroed by the launch itself on `stream`). */
cudaError_t x3k_launch_seg(const X3SearchParams &prm, cudaStream_t stream, int *launches);
uint32_t x3k_seg_max_distances(void);
int x3k_seg_min_t(void);
uint32_t x3k_seg_positions(uint32_t D); /* positions per segment */
/* what X3S_KERNEL_DEFAULT runs for these parameters: 2 = segment search, 1 = rank search,
 * 0 = brute-force stream kernel (X3_NO_SEG=1 in the environment takes the segment search out) */
int x3k_default_kind(uint32_t D, int t, bool want_table);

/* One-time per-device setup (opt-in shared memory size). */
cudaError_t x3k_init_device(void);

#endif
