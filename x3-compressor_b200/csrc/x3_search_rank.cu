/*
 * x3_search_rank.cu -- "rank" search: the forward-window search of reference
 * backend.c:58-78 as an occurrence-rank problem, sm_100a.
 *
 * The brute-force kernels test every (position, distance) pair: N * (W - 33) byte
 * compares.  The selection of backend.c:76-78 never needs the counts themselves, only
 * whether count_L(p) exceeds tc*(p) = min(t, count_1(p) - 1):
 *
 *     count_L(p) > k   <=>   the (k+1)-th next occurrence of the L-gram x[p..p+L)
 *                            starts at most D = W - 33 bytes behind p.
 *
 * So level L keeps an array of positions sorted by (L-gram, position); in that order the
 * test is ONE lookup: the element k+1 places further on has the same L-gram and lies
 * within D.  Lstar(p) is the deepest level at which p passes (counts are monotone in L).
 *
 *   level 2   all M = n + D positions (the D positions behind the searched range are
 *             followers only) sorted from x by (b0 b1, position): x3_rank_bytehist_kernel,
 *             x3_rank_radix_kernel<INIT> (digit b1 = x[p+1], reads x in position order, so the
 *             key carries b2 b3 for free) and one ordinary pass (digit b0).
 *   level 1   no sort of its own: the INIT pass leaves the positions p ordered by
 *             (x[p+1], p), which is the level-1 order of q = p + 1.  x3_rank_first_kernel
 *             tests every position there (Lstar is preset to 1) and settles the positions
 *             whose byte occurs at most t times in their window (tc* < t) by walking their
 *             <= t followers: Lstar = min LCP32 (0 when fewer than 2).
 *   level L   x3_rank_level_kernel, one pass over the sorted array (tiles of 2048):
 *               passed(i) = key[i+t+1] == key[i] and pos[i+t+1] - pos[i] <= D
 *               kept(i)   = i lies within D behind a passed element of its group
 *                           (anything else can never be a follower that matters; any
 *                           superset is exact because every kept element is a real
 *                           position with its real L-gram)
 *               new key   = (group rank << 8) | x[pos + L]; compacted with a single-pass
 *                           chained scan (decoupled look-back over tile aggregates, counts
 *                           published before any slow work)
 *               Lstar[pos] is written once, at the level where pos stops passing.
 *             then a stable LSD radix sort of the survivors by the new key
 *             (x3_rank_radix_kernel: one pass per 8 bits, per-digit decoupled look-back,
 *             warp-level match ranking, tile reordered in shared memory), which restores
 *             (gram, position) order.
 *   tail      x3_rank_tail_kernel: once <= 8192 elements are left, one CTA runs every
 *             remaining level in shared memory in a single launch.
 *
 * All per-level state lives in device memory; the kernels of a chunk are a pure
 * kernel -> kernel chain launched with programmatic dependent launch, and the host only
 * learns level sizes (zero-copy reports) to size grids and to stop queueing.
 *
 * The work is sum_L m_L element visits instead of N * D pair tests: 3.3 N on text at the
 * default window, and it does not grow with the window or with t.  Everything is plain
 * coalesced streaming over 4-byte keys and positions (plus one byte gather per element from
 * level 4 on), i.e. memory-system work, not ALU work.
 *
 * Inputs larger than 2^24 - D positions are searched chunk by chunk (ranks and chunk-relative
 * positions then fit 24 bits, so a key is 32 bits).  Lstar only: the 32-bin table H is the
 * brute-force kernels' job.
 */
#include "x3_search_device.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

#ifndef X3_LV_THREADS
#define X3_LV_THREADS 256
#endif
constexpr int LV_THREADS = X3_LV_THREADS;      /* threads of a level-kernel CTA */
constexpr int F1_THREADS = 256;                 /* level-1 kernel: tile geometry of its own */
constexpr int F1_ITEMS = 8;
constexpr int F1_TILE = F1_THREADS * F1_ITEMS;
#ifndef X3_LV_ITEMS
#define X3_LV_ITEMS 8
#endif
constexpr int LV_ITEMS = X3_LV_ITEMS;           /* elements per thread of a level tile */
constexpr int LV_CTAS = (LV_ITEMS <= 8 ? 6 : 4) * 256 / LV_THREADS;  /* resident CTAs per SM the level kernel is built for */
constexpr int LV_TILE = LV_THREADS * LV_ITEMS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef X3_RS_LOOK
#define X3_RS_LOOK 2
#endif
constexpr int RS_LOOK = X3_RS_LOOK;             /* statuses a look-back step fetches at once */
constexpr int RS_ITEMS = 16;                    /* elements per thread of a radix tile: large arrays */
constexpr int RS_ITEMS_SMALL = 8;               /* ... arrays that would not fill the GPU with large tiles */
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_TILE_SMALL = RS_THREADS * RS_ITEMS_SMALL;
constexpr uint32_t RS_SMALL_BELOW = 3u << 20;   /* elements: below this a pass uses the small tile */
constexpr uint32_t RANK_MAX_M = (1u << 24) - 1u;
constexpr uint32_t PMASK = 0x00ffffffu; /* position bits of a position word */
constexpr uint32_t PFLAG = 0x80000000u; /* "passed the previous level" */
constexpr uint32_t KEY_NONE = 0xffffffffu; /* no real key: ranks stay below 2^24 - 1 */

constexpr unsigned long long ST_AGG = 1ull, ST_INC = 2ull;

struct RankCtrl {
	struct {
		uint32_t m;             /* length of the level-L array (lv[1].m = M) */
		uint32_t groups;        /* rank range of the level-L keys (lv[1].groups = 1) */
		uint32_t src0;          /* buffer the level-(L-1) kernel wrote the unsorted array to */
		uint32_t buf;           /* buffer that holds it sorted: src0 ^ (radix passes & 1) */
	} lv[36];
	uint32_t tickets[256];      /* one tile dispenser per launch */
	uint32_t hist[34][4][256];  /* digit histograms of the level-L keys */
};

struct RankArgs {
	const uint8_t *x;   /* chunk base */
	uint8_t *lstar;     /* chunk base */
	uint32_t n_out;     /* positions searched: [0, n_out) */
	uint32_t M;         /* elements: n_out + D */
	uint32_t D;
	int t;
	RankCtrl *ctrl;
	uint32_t *key0, *key1, *pos0, *pos1; /* the two element buffers the levels and radix passes ping-pong between */
	volatile uint32_t *report;    /* pinned host memory: [4 L] = {size, groups} of level L+1, then `seq` */
	uint32_t seq;                 /* tag of this chunk's reports */
	unsigned long long *st_level; /* [tiles of the level kernel] */
	unsigned long long *st_radix; /* [tiles of the radix kernel][256] */
};

/* ---- block-wide scans over 256 threads ------------------------------------------------ */

__device__ __forceinline__ unsigned long long block_excl_sum(unsigned long long v, unsigned long long *ws,
                                                             unsigned long long *total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long o = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= d) {
			inc += o;
		}
	}
	if (lane == 31) {
		ws[warp] = inc;
	}
	__syncthreads();
	if (warp == 0) {
		const unsigned long long w = lane < LV_THREADS / 32 ? ws[lane] : 0ull;
		unsigned long long wi = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long o = __shfl_up_sync(FULL_MASK, wi, d);
			if (lane >= d) {
				wi += o;
			}
		}
		if (lane < LV_THREADS / 32) {
			ws[lane] = wi - w;
		}
		if (lane == LV_THREADS / 32 - 1) {
			ws[LV_THREADS / 32] = wi;
		}
	}
	__syncthreads();
	const unsigned long long res = ws[warp] + inc - v;
	*total = ws[LV_THREADS / 32];
	__syncthreads();
	return res;
}

/* exclusive running maximum; `ident` is what the first thread sees */
__device__ __forceinline__ int block_excl_max(int v, int ident, int *ws)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int o = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= d) {
			inc = max(inc, o);
		}
	}
	if (lane == 31) {
		ws[warp] = inc;
	}
	__syncthreads();
	int before = ident; /* maximum over the warps in front of mine */
	for (int w = 0; w < warp; ++w) {
		before = max(before, ws[w]);
	}
	int ex = __shfl_up_sync(FULL_MASK, inc, 1);
	if (lane == 0) {
		ex = ident;
	}
	__syncthreads();
	return max(before, ex);
}

/* Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization
 * attribute may start while its predecessor in the stream is still draining; it must not touch
 * anything the predecessor wrote before this returns (predecessor complete, memory flushed). */
__device__ __forceinline__ void pdl_wait()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

/* lets the next kernel of the stream begin launching (it still waits in pdl_wait) */
__device__ __forceinline__ void pdl_launch_dependents()
{
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* 8-bit radix passes that sort keys (rank << 8 | byte) with rank < groups */
__host__ __device__ __forceinline__ int radix_passes(uint32_t groups)
{
	int bits = 8;
	while (groups > 1 && ((groups - 1) >> (bits - 8)) != 0) {
		++bits;
	}
	return (bits + 7) / 8;
}

/* ---- set-up: byte histogram of the M elements (both digits of the level-2 sort) ---------------------------------- */

__global__ void __launch_bounds__(256) x3_rank_bytehist_kernel(RankArgs a)
{
	__shared__ uint32_t h[256];
	pdl_launch_dependents();
	h[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t nvec = a.M / 16;
	const uint4 *xv = reinterpret_cast<const uint4 *>(a.x);
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
		const uint4 q = __ldg(xv + v);
		const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			atomicAdd(&h[w[k] & 255u], 1u);
			atomicAdd(&h[(w[k] >> 8) & 255u], 1u);
			atomicAdd(&h[(w[k] >> 16) & 255u], 1u);
			atomicAdd(&h[w[k] >> 24], 1u);
		}
	}
	if (blockIdx.x == 0) {
		for (uint32_t i = nvec * 16 + threadIdx.x; i < a.M; i += blockDim.x) {
			atomicAdd(&h[a.x[i]], 1u);
		}
		if (threadIdx.x == 0) {
			/* level 2: all M positions, sorted from x by (b0, b1): the first pass (digit b1) reads x and
			 * lands in buffer 1 (level 1 is tested on that array), the second (digit b0) sorts it into
			 * buffer 0 */
			a.ctrl->lv[2].m = a.M;
			a.ctrl->lv[2].groups = 256;
			a.ctrl->lv[2].src0 = 0;
			a.ctrl->lv[2].buf = 0;
			/* the b1 digits are the bytes x[1 .. M]: the same histogram, one byte out, one byte in */
			atomicAdd(&a.ctrl->hist[2][0][a.x[a.M]], 1u);
			atomicSub(&a.ctrl->hist[2][0][a.x[0]], 1u);
		}
	}
	__syncthreads();
	if (h[threadIdx.x] != 0) {
		atomicAdd(&a.ctrl->hist[2][0][threadIdx.x], h[threadIdx.x]);
		atomicAdd(&a.ctrl->hist[2][1][threadIdx.x], h[threadIdx.x]);
	}
}

/* ---- one stable 8-bit radix pass -------------------------------------------------------
 * INIT: the first of the two passes that sort level 2 by (b0, b1): the elements are the positions
 * 0 .. M-1 themselves, the key is made of the 4 bytes x[p..p+3], the digit is b1 = x[p+1].
 * A tile is ranked with warp-level digit matching, put in digit order in shared memory while
 * the chained per-digit prefix of the tiles in front resolves, then copied out in runs. */
template <bool INIT, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS, 5) x3_rank_radix_kernel(RankArgs a, int level, int pass, int ticket,
                                                                       uint32_t epoch)
{
	__shared__ uint32_t gbase[256];
	__shared__ uint32_t wcnt[RS_WARPS][256];
	__shared__ uint32_t lbase[256];
	__shared__ uint32_t tbase[256];
	constexpr int TILE = RS_THREADS * ITEMS;
	__shared__ uint32_t skey[TILE], spos[TILE];
	__shared__ uint32_t s_tile;
	__shared__ uint32_t wsum[RS_WARPS + 1];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	pdl_wait(); /* programmatic dependent launch: everything above overlapped the previous kernel's tail */
	pdl_launch_dependents();
	const uint32_t m = a.ctrl->lv[level].m;
	if (!INIT && (m < (uint32_t)a.t + 2u || pass >= radix_passes(a.ctrl->lv[level].groups))) {
		return; /* nobody can pass this level any more / the keys have no such digit */
	}
	const uint32_t src = INIT ? 0u : a.ctrl->lv[level].src0 ^ (uint32_t)(pass & 1);
	const uint32_t *__restrict__ keyIn = src ? a.key1 : a.key0;
	const uint32_t *__restrict__ posIn = src ? a.pos1 : a.pos0;
	uint32_t *__restrict__ keyOut = src ? a.key0 : a.key1;
	uint32_t *__restrict__ posOut = src ? a.pos0 : a.pos1;
	const uint32_t ntiles = (m + TILE - 1) / TILE;
	const int shift = 8 * pass;
	const uint32_t lt = (1u << lane) - 1u;

	/* exclusive scan of this pass's global digit histogram */
	{
		const uint32_t h = a.ctrl->hist[level][pass][tid];
		uint32_t inc = h;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
			if (lane >= d) {
				inc += o;
			}
		}
		if (lane == 31) {
			wsum[warp] = inc;
		}
		__syncthreads();
		uint32_t before = 0;
		for (int w = 0; w < warp; ++w) {
			before += wsum[w];
		}
		gbase[tid] = before + inc - h;
		__syncthreads();
	}

	for (;;) {
		if (tid == 0) {
			s_tile = atomicAdd(&a.ctrl->tickets[ticket], 1u);
		}
#pragma unroll
		for (int w = 0; w < RS_WARPS; ++w) {
			wcnt[w][tid] = 0;
		}
		__syncthreads();
		const uint32_t tile = s_tile;
		if (tile >= ntiles) {
			break;
		}
		const uint32_t base = tile * TILE + warp * (32 * ITEMS);
		uint32_t key[ITEMS], off[ITEMS];
#pragma unroll
		for (int k = 0; k < ITEMS; ++k) {
			const uint32_t i = base + 32 * k + lane;
			const bool valid = i < m;
			if (INIT) {
				/* level-1 key: the byte (the sort digit) with its three successors on top, so that the
				 * level-1 kernel needs no gathers */
				const uint32_t *xw = reinterpret_cast<const uint32_t *>(a.x) + (i >> 2);
				const uint32_t w4 = valid ? __funnelshift_r(__ldg(xw), __ldg(xw + 1), 8 * (i & 3)) : 0u;
				/* b3 b2 | b0 b1: sorted by b1 here, by b0 in the next pass */
				key[k] = (w4 & 0xffff0000u) | ((w4 & 255u) << 8) | ((w4 >> 8) & 255u);
			} else {
				key[k] = valid ? keyIn[i] : 0u;
			}
		}
#pragma unroll
		for (int k = 0; k < ITEMS; ++k) {
			const bool valid = base + 32 * k + lane < m;
			const uint32_t d = valid ? ((key[k] >> shift) & 255u) : 256u;
			const uint32_t peers = __match_any_sync(FULL_MASK, d);
			const int leader = __ffs(peers) - 1;
			uint32_t old = 0;
			if (lane == leader && valid) {
				old = wcnt[warp][d];
				wcnt[warp][d] = old + __popc(peers);
			}
			old = __shfl_sync(FULL_MASK, old, leader);
			off[k] = valid ? (old + __popc(peers & lt)) | (d << 16) : 0xffffffffu;
			__syncwarp();
		}
		__syncthreads();
		/* digit tid: offsets of the warps inside the tile, the tile total (published at once for the
		 * tiles behind), and the digit's first slot in the tile's sorted order */
		unsigned long long *mine = a.st_radix + (size_t)tile * 256 + tid;
		const unsigned long long ep = (unsigned long long)epoch << 32;
		uint32_t run = 0;
		{
#pragma unroll
			for (int w = 0; w < RS_WARPS; ++w) {
				const uint32_t c = wcnt[w][tid];
				wcnt[w][tid] = run;
				run += c;
			}
			st_status(mine, ep | ((tile == 0 ? ST_INC : ST_AGG) << 30) | run);
			uint32_t inc = run;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
				if (lane >= d) {
					inc += o;
				}
			}
			if (lane == 31) {
				wsum[warp] = inc;
			}
			__syncthreads();
			uint32_t before = 0;
			for (int w = 0; w < warp; ++w) {
				before += wsum[w];
			}
			lbase[tid] = before + inc - run;
		}
		__syncthreads();
		/* the tile in digit order */
#pragma unroll
		for (int k = 0; k < ITEMS; ++k) {
			if (off[k] != 0xffffffffu) {
				const uint32_t d = off[k] >> 16;
				const uint32_t idx = lbase[d] + wcnt[warp][d] + (off[k] & 0xffffu);
				skey[idx] = key[k];
				spos[idx] = INIT ? base + 32 * k + lane : posIn[base + 32 * k + lane];
			}
		}
		/* chained prefix of digit tid over the tiles in front */
		{
			uint32_t excl = 0;
			if (tile != 0) {
				const unsigned long long *look = mine - 256;
				/* RS_LOOK tiles per round trip: the statuses behind the one being waited for are fetched with it
				 * (the walk back to the nearest inclusive prefix is ~14 tiles long on the big arrays, one L2
				 * round trip each when taken one by one; two at a time: C2 0.98 -> 0.93 ms) */
				uint32_t left = tile; /* tiles in front of `look`, itself included */
				for (;;) {
					unsigned long long sv[RS_LOOK];
#pragma unroll
					for (int q = 0; q < RS_LOOK; ++q) {
						sv[q] = (uint32_t)q < left ? ld_status(look - 256 * q) : 0ull;
					}
					uint32_t used = 0;
					bool done = false;
#pragma unroll
					for (int q = 0; q < RS_LOOK; ++q) {
						const unsigned long long sq = sv[q];
						if (!done && used == (uint32_t)q && (sq >> 32) == epoch && ((sq >> 30) & 3ull) != 0ull) {
							excl += (uint32_t)(sq & 0x3fffffffull);
							++used;
							done = ((sq >> 30) & 3ull) == ST_INC;
						}
					}
					if (done) {
						break;
					}
					look -= 256 * used; /* 0: the nearest one is not published yet */
					left -= used;
				}
				st_status(mine, ep | (ST_INC << 30) | (excl + run));
			}
			tbase[tid] = gbase[tid] + excl - lbase[tid];
		}
		__syncthreads();
		const uint32_t cnt = m - tile * TILE < TILE ? m - tile * TILE : TILE;
		for (uint32_t j = tid; j < cnt; j += RS_THREADS) {
			const uint32_t kk = skey[j];
			const uint32_t dst = tbase[(kk >> shift) & 255u] + j;
			keyOut[dst] = kk;
			posOut[dst] = spos[j];
		}
		__syncthreads();
	}
}

/* ---- level 1: the first byte ----------------------------------------------------------------
 * No sort of its own: the first of the two level-2 passes leaves the positions p ordered by
 * (x[p+1], p), which IS the level-1 order of the positions q = p + 1 (key: b3 b2 | b0 b1 with
 * b1 = x[q], b2 b3 = the two bytes behind it).  A position passes level 1 when the (t+1)-th next
 * position with its byte lies within D; Lstar of every searched position was preset to 1, deeper
 * levels raise it.  A position that does NOT pass has c1 <= t followers within D: tc* = c1 - 1,
 * so Lstar = #{L : count_L >= c1} = the smallest LCP32 over those followers, 0 when c1 < 2
 * (backend.c:76-78 collapsed) -- settled here by walking them.  Position 0 has no element (p = -1):
 * its followers are the head of its byte's group, found through the digit histogram. */
__global__ void __launch_bounds__(F1_THREADS, 6) x3_rank_first_kernel(RankArgs a)
{
	__shared__ __align__(16) uint32_t sk[F1_TILE + 256 + 8]; /* + look-ahead (<= 255) + the walk's read-ahead */
	__shared__ __align__(16) uint32_t sp[F1_TILE + 256 + 8];
	const int tid = threadIdx.x;
	pdl_wait();
	pdl_launch_dependents();
	const uint32_t m = a.M;
	const uint32_t ntiles = (m + F1_TILE - 1) / F1_TILE;
	const uint32_t D = a.D, n_out = a.n_out;
	const uint32_t la = (uint32_t)a.t + 1u;
	const uint32_t *__restrict__ keyIn = a.key1; /* the output of x3_rank_radix_kernel<INIT> */
	const uint32_t *__restrict__ posIn = a.pos1;
	const uint8_t *__restrict__ x = a.x;
	if (blockIdx.x == 0 && tid == 0) {
		/* position 0 */
		const uint32_t b = x[0];
		uint32_t gs = 0;
		for (uint32_t v = 0; v < b; ++v) {
			gs += a.ctrl->hist[2][0][v];
		}
		const uint32_t size = a.ctrl->hist[2][0][b];
		if (!(size > (uint32_t)a.t && posIn[gs + (uint32_t)a.t] + 1u <= D)) {
			uint32_t best = 32, c1 = 0;
			for (uint32_t i = 0; i < size && i < (uint32_t)a.t; ++i) {
				const uint32_t q = posIn[gs + i] + 1u;
				if (q > D) {
					break;
				}
				++c1;
				uint32_t l = 1;
				while (l < best && x[l] == x[q + l]) {
					++l;
				}
				best = l;
			}
			a.lstar[0] = (uint8_t)(c1 >= 2 ? best : 0u);
		}
	}
	for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const uint32_t base = tile * F1_TILE;
		const uint32_t i0 = base + tid * F1_ITEMS;
		if (i0 + F1_ITEMS <= m) {
#pragma unroll
			for (int v = 0; v < F1_ITEMS; v += 4) {
				*reinterpret_cast<uint4 *>(sk + tid * F1_ITEMS + v) = *reinterpret_cast<const uint4 *>(keyIn + i0 + v);
				*reinterpret_cast<uint4 *>(sp + tid * F1_ITEMS + v) = *reinterpret_cast<const uint4 *>(posIn + i0 + v);
			}
		} else {
#pragma unroll
			for (int e = 0; e < F1_ITEMS; ++e) {
				const bool v = i0 + e < m;
				sk[tid * F1_ITEMS + e] = v ? keyIn[i0 + e] : 0u;
				sp[tid * F1_ITEMS + e] = v ? posIn[i0 + e] : 0u;
			}
		}
		/* the look-ahead staged behind the tile covers t <= 254; larger t (reference backend.c:21-26
		 * takes any int) reads the elements further on from global memory */
		const bool far = la > 256u;
		if ((uint32_t)tid < (far ? 256u : la)) {
			const uint32_t i = base + F1_TILE + tid;
			sk[F1_TILE + tid] = i < m ? keyIn[i] : 0u;
			sp[F1_TILE + tid] = i < m ? posIn[i] : 0u;
		}
		__syncthreads();
#pragma unroll 1
		for (int e = 0; e < F1_ITEMS; ++e) {
			const uint32_t idx = e * F1_THREADS + tid;
			if (base + idx >= m) {
				break;
			}
			const uint32_t kk = sk[idx], pp = sp[idx];
			const uint32_t qq = pp + 1u; /* the position this element stands for at level 1 */
			if (qq >= n_out) {
				continue; /* the positions behind the searched range are followers only */
			}
			if (base + idx + la < m) {
				const uint32_t kf = far ? keyIn[base + idx + la] : sk[idx + la];
				const uint32_t pf = far ? posIn[base + idx + la] : sp[idx + la];
				if (((kf ^ kk) & 255u) == 0u && pf - pp <= D) {
					continue; /* passes: Lstar >= 1 */
				}
			}
			const uint32_t room = m - 1u - (base + idx);
			const uint32_t lim = room < (uint32_t)a.t ? room : (uint32_t)a.t;
			uint32_t best = 32, c1 = 0;
			bool done = false;
			for (uint32_t j = 1; j <= lim && !done; j += 4) {
				/* four followers per round: their shared-memory loads are independent */
				uint32_t kf[4], q[4];
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const uint32_t o = idx + j + u;
					const bool staged = o < (uint32_t)F1_TILE + 256u;
					kf[u] = staged ? sk[o] : (base + o < m ? keyIn[base + o] : 0u);
					q[u] = staged ? sp[o] : (base + o < m ? posIn[base + o] : 0u);
				}
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					if (done || j + u > lim || ((kf[u] ^ kk) & 255u) != 0u || q[u] - pp > D) {
						done = true;
					} else {
						++c1;
						/* bytes 1 and 2 of both positions sit in the key's upper half */
						const uint32_t df = (kf[u] ^ kk) >> 16;
						uint32_t l = df != 0u ? 1u + ((uint32_t)(__ffs((int)df) - 1) >> 3) : 3u;
						if (l == 3u) {
							while (l < best && x[qq + l] == x[q[u] + 1u + l]) {
								++l;
							}
						}
						best = min(best, l);
					}
				}
			}
			a.lstar[qq] = (uint8_t)(c1 >= 2 ? best : 0u);
		}
		__syncthreads();
	}
}

/* ---- one level: test, prune, re-key, compact ----------------------------------------------
 * Bit 31 of a position word says "this element passed the previous level": Lstar is written
 * once per position, at the level where it stops passing (or by the level-1 rare path, the
 * flush after the last level, or level 32). */
template <int LK>
__global__ void __launch_bounds__(LV_THREADS, LV_CTAS) x3_rank_level_kernel(RankArgs a, int L, int ticket)
{
	/* LK = min(L, 4), L >= 2.  Keys of levels 2 and 3 carry the bytes behind the gram in their upper
	 * bits (level 2: b3 b2 | b0 b1; level 3: b3 | rank16 b2), put there by the position-ordered passes
	 * that read x; the radix passes only sort the low `gram` bits, so the first byte gather happens at
	 * level 4, when the arrays are small. */
	constexpr uint32_t KM = LK == 2 ? 0xffffu : (LK == 3 ? 0xffffffu : 0xffffffffu);
	__shared__ __align__(16) uint32_t sk[LV_TILE + 256 + 8]; /* + look-ahead (<= 255) + the rare walk's read-ahead */
	__shared__ __align__(16) uint32_t sp[LV_TILE + 256 + 8];
	__shared__ uint32_t hist[4][256];
	__shared__ uint32_t actbits[LV_TILE / 32];
	__shared__ unsigned long long ws[LV_THREADS / 32 + 1];
	__shared__ int wsi[LV_THREADS / 32];
	__shared__ uint32_t s_tile;
	__shared__ unsigned long long s_excl;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	pdl_wait();
	pdl_launch_dependents();
	const uint32_t m = a.ctrl->lv[L].m;
	const uint32_t ntiles = (m + LV_TILE - 1) / LV_TILE;
	const uint32_t D = a.D, n_out = a.n_out;
	const uint32_t la = (uint32_t)a.t + 1u; /* look-ahead of the test, <= 255 */
	const bool emit = L < 32;
	if (m < la + 1u) {
		/* nobody can pass this level (the radix passes did not run either): whoever passed the
		 * previous one keeps it, and the search is over */
		if (blockIdx.x == 0) {
			/* (level 2's input carries no flags -- and was never written when its radix passes returned
			 * at once: m < t + 2 -- so there is nothing to flush there) */
			const uint32_t *__restrict__ pu = a.ctrl->lv[L].src0 ? a.pos1 : a.pos0;
			for (uint32_t i = tid; L > 2 && i < m; i += LV_THREADS) {
				const uint32_t pw = pu[i];
				if (pw & PFLAG) {
					a.lstar[pw & PMASK] = (uint8_t)(L - 1);
				}
			}
			if (tid == 0) {
				a.ctrl->lv[L + 1].m = 0;
				a.ctrl->lv[L + 1].groups = 0;
				a.report[4 * L] = 0;
				a.report[4 * L + 1] = 0;
				__threadfence_system();
				a.report[4 * L + 2] = a.seq;
			}
		}
		return;
	}
	const uint32_t inb = a.ctrl->lv[L].buf;
	const uint32_t *__restrict__ keyIn = inb ? a.key1 : a.key0;
	const uint32_t *__restrict__ posIn = inb ? a.pos1 : a.pos0;
	uint32_t *__restrict__ keyOut = inb ? a.key0 : a.key1;
	uint32_t *__restrict__ posOut = inb ? a.pos0 : a.pos1;
	/* digits of the new keys that can be non-zero: ranks are below the number of L-grams */
	const uint32_t rbound = L == 1 ? (m < 256u ? m : 256u) : (L == 2 ? (m < 65536u ? m : 65536u) : m);
	const int ndig = radix_passes(rbound);
	if (tid < 256) {
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			hist[j][tid] = 0;
		}
	}
	__syncthreads();

	for (;;) {
		if (tid == 0) {
			s_tile = atomicAdd(&a.ctrl->tickets[ticket], 1u);
		}
		__syncthreads();
		const uint32_t tile = s_tile;
		if (tile >= ntiles) {
			break;
		}
		const uint32_t base = tile * LV_TILE;
		const uint32_t i0 = base + tid * LV_ITEMS;
		/* my LV_ITEMS consecutive elements stay in registers; the tile and its look-ahead go to shared memory */
		uint32_t k[LV_ITEMS], p[LV_ITEMS];
		if (i0 + LV_ITEMS <= m) {
#pragma unroll
			for (int v = 0; v < LV_ITEMS; v += 4) {
				const uint4 kv = *reinterpret_cast<const uint4 *>(keyIn + i0 + v);
				const uint4 pv = *reinterpret_cast<const uint4 *>(posIn + i0 + v);
				k[v] = kv.x; k[v + 1] = kv.y; k[v + 2] = kv.z; k[v + 3] = kv.w;
				p[v] = pv.x; p[v + 1] = pv.y; p[v + 2] = pv.z; p[v + 3] = pv.w;
			}
		} else {
#pragma unroll
			for (int e = 0; e < LV_ITEMS; ++e) {
				const bool v = i0 + e < m;
				k[e] = v ? keyIn[i0 + e] : KEY_NONE;
				p[e] = v ? posIn[i0 + e] : 0u;
			}
		}
#pragma unroll
		for (int v = 0; v < LV_ITEMS; v += 4) {
			*reinterpret_cast<uint4 *>(sk + tid * LV_ITEMS + v) = make_uint4(k[v], k[v + 1], k[v + 2], k[v + 3]);
			*reinterpret_cast<uint4 *>(sp + tid * LV_ITEMS + v) = make_uint4(p[v], p[v + 1], p[v + 2], p[v + 3]);
		}
		/* the look-ahead staged behind the tile covers t <= 254; larger t reads the element t+1 places
		 * further on from global memory */
		const bool far = la > 256u;
		if (!far && (uint32_t)tid < la) {
			const uint32_t i = base + LV_TILE + tid;
			sk[LV_TILE + tid] = i < m ? keyIn[i] : KEY_NONE;
			sp[LV_TILE + tid] = i < m ? posIn[i] : 0u;
		}
		__syncthreads();
		/* the test, element idx = 256 e + tid (conflict-free): does the (t+1)-th next element of the
		 * array share the gram within D? */
#pragma unroll
		for (int e = 0; e < LV_ITEMS; ++e) {
			const uint32_t idx = e * LV_THREADS + tid;
			const uint32_t kk = sk[idx], pw = sp[idx];
			const uint32_t pp = pw & PMASK;
			const bool valid = base + idx < m;
			const bool out = pp < n_out;
			bool pass;
			/* masked compare: no sentinel value exists, so the array bound is checked */
			pass = valid && out && base + idx + la < m;
			if (pass) {
				const uint32_t kf = far ? keyIn[base + idx + la] : sk[idx + la];
				const uint32_t pf = far ? posIn[base + idx + la] : sp[idx + la];
				pass = ((kf ^ kk) & KM) == 0u && (pf & PMASK) - pp <= D;
			}
			if ((pw & PFLAG) && !pass) {
				a.lstar[pp] = (uint8_t)(L - 1); /* passed level L-1, stops here */
			}
			if (L == 32 && pass) {
				a.lstar[pp] = 32;
			}
			const uint32_t am = __ballot_sync(FULL_MASK, pass);
			if (lane == 0) {
				actbits[e * (LV_THREADS / 32) + warp] = am;
			}
		}
		if (!emit) {
			__syncthreads();
			continue;
		}
		__syncthreads();
		const uint32_t actm = (actbits[(tid * LV_ITEMS) >> 5] >> ((tid * LV_ITEMS) & 31)) & ((1u << LV_ITEMS) - 1u);
		const int mylast = actm != 0 ? (int)i0 + (31 - __clz((int)actm)) : -1;
		/* last passed element in front of mine; in front of the tile: pretend its neighbour passed
		 * (a superset of the exact rule, at most t+1 extra elements per tile) */
		const int prevlast = block_excl_max(mylast, (int)base - 1, wsi);
		bool have = prevlast >= 0;
		uint32_t rk = 0, rp = 0;
		if (have) {
			if (prevlast >= (int)base) {
				rk = sk[prevlast - (int)base];
				rp = sp[prevlast - (int)base] & PMASK;
			} else {
				rk = __ldg(keyIn + prevlast);
				rp = __ldg(posIn + prevlast) & PMASK;
			}
			rk &= KM;
		}
		uint32_t partm = 0, headm = 0;
		uint32_t kprev = tid > 0 ? sk[tid * LV_ITEMS - 1] : (base > 0 ? __ldg(keyIn + base - 1) : ~k[0]);
		kprev &= KM; /* element 0 is a head by its index */
#pragma unroll
		for (int e = 0; e < LV_ITEMS; ++e) {
			if (i0 + e < m) {
				const uint32_t pp = p[e] & PMASK;
				if ((actm >> e) & 1u) {
					have = true;
					rp = pp;
				}
				const uint32_t ke = k[e] & KM;
				if ((actm >> e) & 1u) {
					rk = ke;
				}
				if (have && rk == ke && pp - rp <= D) {
					partm |= 1u << e;
				}
				if (ke != kprev || i0 + e == 0) {
					headm |= 1u << e;
				}
				kprev = ke;
			}
		}
		/* the byte that extends each kept gram (one gather per element; level 1 has it in the key),
		 * fetched before the scans so that its latency overlaps them */
		uint32_t nb[LV_ITEMS / 4] = {0u};
#pragma unroll
		for (int e = 0; e < LV_ITEMS; ++e) {
			if ((partm >> e) & 1u) {
				const uint32_t by = LK == 2 ? (k[e] >> 16) & 255u
				                  : LK == 3 ? k[e] >> 24
				                            : (uint32_t)__ldg(a.x + (p[e] & PMASK) + L);
				nb[e >> 2] |= by << (8 * (e & 3));
			}
		}
		/* chained scan of (kept, heads) over the tiles */
		const unsigned long long mine = (unsigned long long)__popc(partm) | ((unsigned long long)__popc(headm) << 16);
		unsigned long long total;
		const unsigned long long ex = block_excl_sum(mine, ws, &total);
		const unsigned long long pk = (total & 0xffffull) | (((total >> 16) & 0xffffull) << 24);
		const unsigned long long ep = (unsigned long long)L << 50;
		if (tid == 0) {
			/* the tile's own counts are out at once for the tiles behind */
			st_status(a.st_level + tile, ep | ((tile == 0 ? ST_INC : ST_AGG) << 48) | pk);
		}
		/* the kept elements, compacted in tile order into the (now free) staging arrays; keys carry
		 * the tile-local group rank until the prefix over the tiles in front is known */
		{
			uint32_t lidx = (uint32_t)(ex & 0xffffull);
			uint32_t hloc = (uint32_t)((ex >> 16) & 0xffffull);
#pragma unroll
			for (int e = 0; e < LV_ITEMS; ++e) {
				hloc += (headm >> e) & 1u;
				if ((partm >> e) & 1u) {
					/* bytes still ahead stay on top of the new (rank, byte) gram.  Sums, not ORs: the local
					 * rank is -1 for elements of a group that began in an earlier tile, and only becomes a
					 * real rank (mod 2^32) once the tiles in front are added at copy-out */
					const uint32_t upper = LK == 2 ? k[e] & 0xff000000u : 0u;
					sk[lidx] = upper + ((hloc - 1u) << 8) + ((nb[e >> 2] >> (8 * (e & 3))) & 255u);
					sp[lidx] = (p[e] & PMASK) | (((actm >> e) & 1u) ? PFLAG : 0u);
					++lidx;
				}
			}
		}
		if (tid < 32) {
			unsigned long long excl = 0;
			if (tile != 0) {
				int look = (int)tile - 1;
				for (;;) {
					const int idx = look - lane;
					unsigned long long s = 0;
					bool ready;
					do {
						if (idx >= 0) {
							s = ld_status(a.st_level + idx);
							ready = (s >> 50) == (unsigned long long)L && ((s >> 48) & 3ull) != 0ull;
						} else {
							s = ST_INC << 48; /* in front of tile 0: an empty inclusive prefix */
							ready = true;
						}
					} while (__any_sync(FULL_MASK, !ready));
					const uint32_t incm = __ballot_sync(FULL_MASK, ((s >> 48) & 3ull) == ST_INC);
					const int first = incm != 0 ? __ffs(incm) - 1 : 31;
					unsigned long long v = lane <= first ? (s & 0xffffffffffffull) : 0ull;
#pragma unroll
					for (int d = 16; d >= 1; d >>= 1) {
						v += __shfl_xor_sync(FULL_MASK, v, d);
					}
					excl += v; /* both 24-bit fields stay below 2^24: no carry between them */
					if (incm != 0) {
						break;
					}
					look -= 32;
				}
				if (lane == 0) {
					st_status(a.st_level + tile, ep | (ST_INC << 48) | (excl + pk));
				}
			}
			if (lane == 0) {
				s_excl = excl;
				if (tile == ntiles - 1) {
					const unsigned long long tot = excl + pk;
					const uint32_t g = (uint32_t)((tot >> 24) & 0xffffffull);
					a.ctrl->lv[L + 1].m = (uint32_t)(tot & 0xffffffull);
					a.ctrl->lv[L + 1].groups = g;
					a.ctrl->lv[L + 1].src0 = inb ^ 1u;
					a.ctrl->lv[L + 1].buf = inb ^ 1u ^ (uint32_t)(radix_passes(g) & 1);
					/* the host sizes the launches two levels on from this (zero-copy, no stream traffic) */
					a.report[4 * L] = (uint32_t)(tot & 0xffffffull);
					a.report[4 * L + 1] = g;
					__threadfence_system();
					a.report[4 * L + 2] = a.seq;
				}
			}
		}
		__syncthreads();
		{
			const unsigned long long tex = s_excl;
			const uint32_t dst0 = (uint32_t)(tex & 0xffffffull);
			const uint32_t hadd = (uint32_t)((tex >> 24) & 0xffffffull) << 8;
			const uint32_t cnt = (uint32_t)(total & 0xffffull);
			for (uint32_t j0 = 0; j0 < cnt; j0 += LV_THREADS) {
				const uint32_t j = j0 + tid;
				const bool on = j < cnt;
				uint32_t nk = 0;
				if (on) {
					nk = sk[j] + hadd;
					keyOut[dst0 + j] = nk;
					posOut[dst0 + j] = sp[j];
				}
				/* digit histograms of the new keys */
				/* digit 0 is the byte behind the gram: many distinct values per warp, where MATCH.ANY is at its
				 * slowest and plain shared-memory adds collide little; the rank digits are (nearly) uniform
				 * within a warp, where it is the other way round (C2 1.000 -> 0.975 ms) */
				if (on) {
					atomicAdd(&hist[0][nk & 255u], 1u);
				}
				if (ndig > 1) {
					/* one match on the whole rank serves all of its digits */
					const uint32_t r = on ? (nk >> 8) & (0xffffffffu >> (40 - 8 * ndig)) : 0xffffffffu;
					const uint32_t peers = __match_any_sync(FULL_MASK, r);
					if (on && lane == __ffs(peers) - 1) {
						const uint32_t c = (uint32_t)__popc(peers);
						for (int dgt = 1; dgt < ndig; ++dgt) {
							atomicAdd(&hist[dgt][(nk >> (8 * dgt)) & 255u], c);
						}
					}
				}
			}
		}
		__syncthreads();
	}
	if (emit) {
		__syncthreads();
		for (int j = 0; j < ndig; ++j) {
			if (tid < 256 && hist[j][tid] != 0) {
				atomicAdd(&a.ctrl->hist[L + 1][j][tid], hist[j][tid]);
			}
		}
	}
}

/* ---- the tail: every remaining level of a small array in one launch ------------------------
 * Once an array has shrunk to TL_CAP elements the levels cost launches, not bandwidth (a long
 * run of one byte keeps a few thousand elements alive through all 32 levels).  One CTA then
 * takes the sorted level-L0 array into shared memory and runs test / prune / re-key / radix sort
 * for every remaining level there, with the same rules as the level and radix kernels. */
constexpr int TL_THREADS = 1024;
constexpr int TL_WARPS = TL_THREADS / 32;
constexpr int TL_ITEMS = 8;
constexpr int TL_CAP = TL_THREADS * TL_ITEMS;
constexpr size_t TL_SMEM = (size_t)4 * TL_CAP * 4 + (size_t)TL_WARPS * 256 * 4 + 256 * 4 + (TL_CAP / 32) * 4;

__device__ __forceinline__ unsigned long long tl_excl_sum(unsigned long long v, unsigned long long *ws,
                                                          unsigned long long *total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long o = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= d) {
			inc += o;
		}
	}
	if (lane == 31) {
		ws[warp] = inc;
	}
	__syncthreads();
	if (warp == 0) {
		const unsigned long long w = ws[lane];
		unsigned long long wi = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long o = __shfl_up_sync(FULL_MASK, wi, d);
			if (lane >= d) {
				wi += o;
			}
		}
		ws[lane] = wi - w;
		if (lane == 31) {
			ws[32] = wi;
		}
	}
	__syncthreads();
	const unsigned long long res = ws[warp] + inc - v;
	*total = ws[32];
	__syncthreads();
	return res;
}

__device__ __forceinline__ int tl_excl_max(int v, int ident, int *ws)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int o = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= d) {
			inc = max(inc, o);
		}
	}
	if (lane == 31) {
		ws[warp] = inc;
	}
	__syncthreads();
	if (warp == 0) {
		const int w = ws[lane];
		int wi = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int o = __shfl_up_sync(FULL_MASK, wi, d);
			if (lane >= d) {
				wi = max(wi, o);
			}
		}
		/* exclusive over the warps */
		int ex = __shfl_up_sync(FULL_MASK, wi, 1);
		if (lane == 0) {
			ex = ident;
		}
		ws[lane] = max(ex, ident);
	}
	__syncthreads();
	int ex = __shfl_up_sync(FULL_MASK, inc, 1);
	if (lane == 0) {
		ex = ident;
	}
	const int res = max(ws[warp], ex);
	__syncthreads();
	return res;
}

__global__ void __launch_bounds__(TL_THREADS, 1) x3_rank_tail_kernel(RankArgs a, int L0)
{
	extern __shared__ __align__(16) uint8_t tl_smem[];
	/* two (key, position) buffers: buffer c at words [2 c TL_CAP, 2 (c+1) TL_CAP) */
	uint32_t *const sm32 = reinterpret_cast<uint32_t *>(tl_smem);
#define TL_K(c) (sm32 + (size_t)(c) * 2 * TL_CAP)
#define TL_P(c) (sm32 + (size_t)(c) * 2 * TL_CAP + TL_CAP)
	uint32_t(*wcnt)[256] = reinterpret_cast<uint32_t(*)[256]>(sm32 + 4 * TL_CAP);
	uint32_t *dbase = reinterpret_cast<uint32_t *>(wcnt + TL_WARPS);
	uint32_t *actbits = dbase + 256;
	__shared__ unsigned long long ws[33];
	__shared__ int wsi[32];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t D = a.D, n_out = a.n_out;
	const uint32_t la = (uint32_t)a.t + 1u;
	const uint32_t lt = (1u << lane) - 1u;
	pdl_wait();
	uint32_t m = a.ctrl->lv[L0].m;
	if (m == 0 || m > (uint32_t)TL_CAP) {
		return; /* the search ended earlier (m == 0); m > TL_CAP cannot happen (the host queues this
		         * kernel only behind a level whose size was at most TL_CAP) */
	}
	int cur = 0;
	{
		/* fewer than t+2 elements: the radix passes did not run, the array is where the level kernel left it */
		const uint32_t from = m < la + 1u ? a.ctrl->lv[L0].src0 : a.ctrl->lv[L0].buf;
		const uint32_t *__restrict__ gk = from ? a.key1 : a.key0;
		const uint32_t *__restrict__ gp = from ? a.pos1 : a.pos0;
		for (uint32_t i = tid; i < m; i += TL_THREADS) {
			TL_K(0)[i] = gk[i];
			TL_P(0)[i] = gp[i];
		}
	}
	__syncthreads();

	for (int L = L0; L <= 32; ++L) {
		uint32_t *K = TL_K(cur), *P = TL_P(cur), *KO = TL_K(cur ^ 1), *PO = TL_P(cur ^ 1);
		if (m < la + 1u) {
			/* nobody can pass this level: whoever passed the previous one keeps it (level 2's input has no
			 * flags, and is unwritten scratch when its radix passes returned at once) */
			for (uint32_t i = tid; L > 2 && i < m; i += TL_THREADS) {
				const uint32_t pw = P[i];
				if (pw & PFLAG) {
					a.lstar[pw & PMASK] = (uint8_t)(L - 1);
				}
			}
			break;
		}
		const uint32_t KM = L == 1 ? 0xffu : (L == 2 ? 0xffffu : (L == 3 ? 0xffffffu : 0xffffffffu));
		/* the test (striped: conflict-free) */
#pragma unroll
		for (int e = 0; e < TL_ITEMS; ++e) {
			const uint32_t i = e * TL_THREADS + tid;
			bool pass = false;
			if (i < m) {
				const uint32_t kk = K[i], pw = P[i];
				const uint32_t pp = pw & PMASK;
				pass = pp < n_out && i + la < m && ((K[i + la] ^ kk) & KM) == 0u && (P[i + la] & PMASK) - pp <= D;
				if ((pw & PFLAG) && !pass) {
					a.lstar[pp] = (uint8_t)(L - 1);
				}
				if (L == 32 && pass) {
					a.lstar[pp] = 32;
				}
			}
			const uint32_t am = __ballot_sync(FULL_MASK, pass);
			if (lane == 0) {
				actbits[e * TL_WARPS + warp] = am;
			}
		}
		__syncthreads();
		if (L == 32) {
			break;
		}
		/* prune and re-key (blocked: thread owns 8 consecutive elements) */
		const uint32_t i0 = tid * TL_ITEMS;
		const uint32_t actm = (actbits[tid >> 2] >> ((tid & 3) * 8)) & 0xffu;
		const int mylast = actm != 0 ? (int)i0 + (31 - __clz((int)actm)) : -1;
		const int prevlast = tl_excl_max(mylast, -1, wsi);
		bool have = prevlast >= 0;
		uint32_t rk = 0, rp = 0;
		if (have) {
			rk = K[prevlast] & KM;
			rp = P[prevlast] & PMASK;
		}
		uint32_t partm = 0, headm = 0;
		uint32_t kprev = i0 > 0 && i0 <= m ? K[i0 - 1] & KM : 0u;
		uint32_t k[TL_ITEMS], p[TL_ITEMS];
#pragma unroll
		for (int e = 0; e < TL_ITEMS; ++e) {
			k[e] = 0;
			p[e] = 0;
			if (i0 + e < m) {
				k[e] = K[i0 + e];
				p[e] = P[i0 + e];
				const uint32_t pp = p[e] & PMASK;
				const uint32_t ke = k[e] & KM;
				if ((actm >> e) & 1u) {
					have = true;
					rp = pp;
					rk = ke;
				}
				if (have && rk == ke && pp - rp <= D) {
					partm |= 1u << e;
				}
				if (i0 + e == 0 || ke != kprev) {
					headm |= 1u << e;
				}
				kprev = ke;
			}
		}
		const unsigned long long mine = (unsigned long long)__popc(partm) | ((unsigned long long)__popc(headm) << 32);
		unsigned long long total;
		const unsigned long long ex = tl_excl_sum(mine, ws, &total);
		{
			uint32_t dst = (uint32_t)(ex & 0xffffffffull);
			uint32_t hcount = (uint32_t)(ex >> 32);
#pragma unroll
			for (int e = 0; e < TL_ITEMS; ++e) {
				hcount += (headm >> e) & 1u;
				if ((partm >> e) & 1u) {
					const uint32_t pp = p[e] & PMASK;
					const uint32_t by = L == 1 ? (k[e] >> 8) & 255u
					                  : L == 2 ? (k[e] >> 16) & 255u
					                  : L == 3 ? k[e] >> 24
					                           : (uint32_t)__ldg(a.x + pp + L);
					const uint32_t upper = L == 1 ? k[e] & 0xffff0000u : (L == 2 ? k[e] & 0xff000000u : 0u);
					KO[dst] = upper + ((hcount - 1u) << 8) + by;
					PO[dst] = pp | (((actm >> e) & 1u) ? PFLAG : 0u);
					++dst;
				}
			}
		}
		__syncthreads();
		m = (uint32_t)(total & 0xffffffffull);
		const uint32_t groups = (uint32_t)(total >> 32);
		cur ^= 1;
		if (m < la + 1u) {
			continue; /* the next iteration writes the survivors out and stops */
		}
		/* stable LSD radix sort of the survivors by (rank, byte), 8 bits per pass, all in shared memory --
		 * unless they are in that order already (a long run of one byte keeps thousands of elements alive
		 * through all 32 levels, every one of them followed by the same byte: nothing to sort) */
		int np = radix_passes(groups);
		{
			const uint32_t smask = np >= 4 ? 0xffffffffu : (1u << (8 * np)) - 1u;
			const uint32_t *NK = TL_K(cur);
			int unsorted = 0;
			for (uint32_t i = tid; i + 1 < m; i += TL_THREADS) {
				unsorted |= (NK[i] & smask) > (NK[i + 1] & smask);
			}
			if (__syncthreads_or(unsorted) == 0) {
				np = 0;
			}
		}
		for (int pass = 0; pass < np; ++pass) {
			uint32_t *SK = TL_K(cur), *SP = TL_P(cur), *DK = TL_K(cur ^ 1), *DP = TL_P(cur ^ 1);
			const int shift = 8 * pass;
			for (int j = tid; j < TL_WARPS * 256; j += TL_THREADS) {
				(&wcnt[0][0])[j] = 0;
			}
			__syncthreads();
			uint32_t key[TL_ITEMS], pos[TL_ITEMS], off[TL_ITEMS];
#pragma unroll
			for (int q = 0; q < TL_ITEMS; ++q) {
				const uint32_t i = warp * (32 * TL_ITEMS) + 32 * q + lane;
				const bool valid = i < m;
				key[q] = valid ? SK[i] : 0u;
				pos[q] = valid ? SP[i] : 0u;
				const uint32_t d = valid ? ((key[q] >> shift) & 255u) : 256u;
				const uint32_t peers = __match_any_sync(FULL_MASK, d);
				const int leader = __ffs(peers) - 1;
				uint32_t old = 0;
				if (lane == leader && valid) {
					old = wcnt[warp][d];
					wcnt[warp][d] = old + __popc(peers);
				}
				old = __shfl_sync(FULL_MASK, old, leader);
				off[q] = valid ? (old + __popc(peers & lt)) | (d << 16) : 0xffffffffu;
				__syncwarp();
			}
			__syncthreads();
			/* digit tid (first 256 threads): offsets of the warps, then the digit's first slot */
			unsigned long long run = 0;
			if (tid < 256) {
				uint32_t r = 0;
#pragma unroll 8
				for (int w = 0; w < TL_WARPS; ++w) {
					const uint32_t c = wcnt[w][tid];
					wcnt[w][tid] = r;
					r += c;
				}
				run = r;
			}
			unsigned long long tot2;
			const unsigned long long dex = tl_excl_sum(run, ws, &tot2);
			if (tid < 256) {
				dbase[tid] = (uint32_t)dex;
			}
			__syncthreads();
#pragma unroll
			for (int q = 0; q < TL_ITEMS; ++q) {
				if (off[q] != 0xffffffffu) {
					const uint32_t d = off[q] >> 16;
					const uint32_t dst = dbase[d] + wcnt[warp][d] + (off[q] & 0xffffu);
					DK[dst] = key[q];
					DP[dst] = pos[q];
				}
			}
			__syncthreads();
			cur ^= 1;
		}
	}
}

#undef TL_K
#undef TL_P

/* kernel launch behind another kernel of the same stream with programmatic dependent launch */
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, bool pdl,
                       Args... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid);
	cfg.blockDim = dim3((unsigned)block);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

/* ---- per-device scratch ------------------------------------------------------------------ */
constexpr int RANK_MAX_LANES = 8; /* searches (chunks) of one device that may be in flight at once */
constexpr int RANK_DEFAULT_LANES = 4; /* ... unless X3_RANK_LANES says otherwise */

struct RankScratch {
	uint32_t cap = 0; /* elements */
	uint32_t *key[2] = {nullptr, nullptr};
	uint32_t *pos[2] = {nullptr, nullptr};
	RankCtrl *ctrl = nullptr;
	unsigned long long *st_level = nullptr;
	unsigned long long *st_radix = nullptr;
	uint32_t *h_back = nullptr;       /* pinned, device-visible: the level kernels report lv[L + 1] here */
	uint32_t seq = 0;                 /* tag of the current chunk's reports */
	int sms = 0;
	/* X3_RANK_PROFILE: device time per kernel family of the last search */
	cudaEvent_t pev[520];
	int pkind[260];
	int plevel[260], ppass[260];
	int npev = 0;
	bool pev_made = false;
	double prof_ms[4] = {0, 0, 0, 0};     /* [0] radix passes, [1] level kernels, [2] set-up, [3] unused */
	double prof_elems[4] = {0, 0, 0, 0};  /* elements those launches processed */
	int prof_launches[4] = {0, 0, 0, 0};
};
RankScratch g_rank[64][RANK_MAX_LANES];

/* streams the lanes of a device-level search (x3k_launch_rank) run on, forked from and joined
 * back into the caller's stream */
struct RankLanes {
	cudaStream_t st[RANK_MAX_LANES] = {};
	cudaEvent_t fork = nullptr;
	cudaEvent_t join[RANK_MAX_LANES] = {};
	bool made = false;
};
RankLanes g_lanes[64];

cudaError_t rank_ensure(int dev, int lane, uint32_t M)
{
	RankScratch &s = g_rank[dev][lane];
	cudaError_t e;
	if (s.ctrl == nullptr) {
		if ((e = cudaMalloc((void **)&s.ctrl, sizeof(RankCtrl))) != cudaSuccess) return e;
		if ((e = cudaHostAlloc((void **)&s.h_back, 36 * 16, cudaHostAllocMapped)) != cudaSuccess) return e;
		memset(s.h_back, 0, 36 * 16);
		if ((e = cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
		if ((e = cudaFuncSetAttribute(x3_rank_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_SMEM)) != cudaSuccess) return e;
	}
	if (M <= s.cap) {
		return cudaSuccess;
	}
	for (int j = 0; j < 2; ++j) {
		cudaFree(s.key[j]);
		cudaFree(s.pos[j]);
		s.key[j] = s.pos[j] = nullptr;
	}
	cudaFree(s.st_level);
	cudaFree(s.st_radix);
	s.st_level = s.st_radix = nullptr;
	s.cap = 0;
	const size_t cap = ((size_t)M + 4095) & ~(size_t)4095;
	for (int j = 0; j < 2; ++j) {
		if ((e = cudaMalloc((void **)&s.key[j], cap * 4 + 64)) != cudaSuccess) return e;
		if ((e = cudaMalloc((void **)&s.pos[j], cap * 4 + 64)) != cudaSuccess) return e;
	}
	if ((e = cudaMalloc((void **)&s.st_level, (cap / LV_TILE + 1) * 8)) != cudaSuccess) return e;
	if ((e = cudaMalloc((void **)&s.st_radix, (cap / RS_TILE_SMALL + 1) * 256 * 8)) != cudaSuccess) return e;
	s.cap = (uint32_t)cap;
	return cudaSuccess;
}

/* tuning knobs (never change results): CTAs per SM a persistent grid is capped at, for the radix / level-1
 * kernels and for the level kernels; array size below which a radix pass takes the small tile */
struct RankTune {
	uint32_t rs_per_sm = 8, lv_per_sm = 8, small_below = RS_SMALL_BELOW;
	RankTune()
	{
		const char *v;
		if ((v = getenv("X3_RANK_RSGRID")) != nullptr && atoi(v) >= 1) rs_per_sm = (uint32_t)atoi(v);
		if ((v = getenv("X3_RANK_LVGRID")) != nullptr && atoi(v) >= 1) lv_per_sm = (uint32_t)atoi(v);
		if ((v = getenv("X3_RANK_SMALL")) != nullptr && atoi(v) >= 0) small_below = (uint32_t)atoi(v);
	}
};

/* what every chunk of one call shares */
struct RankCfg {
	uint32_t D;
	int t;
	uint32_t lim;     /* fewer elements than this: nobody can pass */
	bool trace, profile, no_tail, pdl;
	int lag;          /* levels of work that stay queued while the host waits for a level size (>= 1) */
	RankTune tune;
};

/* one lane: a stream with scratch of its own that takes the batch's chunks one after the other.
 * Queueing is a resumable state machine, so that one host thread can keep several lanes fed: a
 * lane that waits for a level report hands the turn to the next one. */
struct RankLane {
	const X3RankBatch *b = nullptr;
	cudaStream_t stream = nullptr;
	int index = 0;
	RankScratch *s = nullptr;
	unsigned long long a0 = 0;    /* first position of the chunk being queued */
	bool active = false;          /* a chunk is being queued */
	bool finished = false;
	RankArgs a;
	int L = 0;                    /* level about to be queued */
	int ticket = 0;
	uint32_t known = 0;           /* upper bound of the size of that level */
	unsigned spins = 0;
	int nl = 0;
};

int rank_grid_for(const RankScratch &s, uint32_t tiles, uint32_t per_sm = 8, int threads = 256)
{
	const uint32_t maxgrid = (uint32_t)s.sms * per_sm * 256u / (uint32_t)threads;
	return (int)(tiles < maxgrid ? (tiles > 0 ? tiles : 1) : maxgrid);
}

void rank_mark(const RankCfg &c, RankLane &ln, int kind, int level, int pass)
{
	/* profile mode: an event in front of every launch (and one behind the last) */
	RankScratch &s = *ln.s;
	if (c.profile && s.npev < 259) {
		cudaEventRecord(s.pev[s.npev], ln.stream);
		s.pkind[s.npev] = kind;
		s.plevel[s.npev] = level;
		s.ppass[s.npev] = pass;
		++s.npev;
	}
}

/* the set-up launches of a chunk [a0, a0 + len): byte histogram, the two level-2 passes, level 1 */
cudaError_t rank_chunk_begin(const RankCfg &c, RankLane &ln, unsigned long long a0, unsigned long long len)
{
	cudaError_t e;
	RankScratch &s = *ln.s;
	cudaStream_t stream = ln.stream;
	RankArgs &a = ln.a;
	ln.a0 = a0;
	a.x = ln.b->x + a0;
	a.lstar = ln.b->lstar + a0;
	a.n_out = (uint32_t)len;
	a.M = a.n_out + c.D;
	a.D = c.D;
	a.t = c.t;
	a.ctrl = s.ctrl;
	a.key0 = s.key[0];
	a.key1 = s.key[1];
	a.pos0 = s.pos[0];
	a.pos1 = s.pos[1];
	a.st_level = s.st_level;
	a.st_radix = s.st_radix;
	{
		void *dp = nullptr;
		if ((e = cudaHostGetDevicePointer(&dp, s.h_back, 0)) != cudaSuccess) return e;
		a.report = (volatile uint32_t *)dp;
	}
	s.seq = s.seq + 1u == 0u ? 1u : s.seq + 1u;
	a.seq = s.seq;
	const uint32_t lv_tiles = (a.M + LV_TILE - 1) / LV_TILE, f1_tiles = (a.M + F1_TILE - 1) / F1_TILE, rs_tiles = (a.M + RS_TILE - 1) / RS_TILE, rs_tiles_small = (a.M + RS_TILE_SMALL - 1) / RS_TILE_SMALL;
	if ((e = cudaMemsetAsync(s.ctrl, 0, sizeof(RankCtrl), stream)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(s.st_level, 0, (size_t)lv_tiles * 8, stream)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(s.st_radix, 0, (size_t)rs_tiles_small * 256 * 8, stream)) != cudaSuccess) return e;
	ln.ticket = 0;
	s.npev = 0;
	/* Lstar = 1 wherever level 1 passes and nothing deeper does; the level-1 kernel writes the rest */
	if ((e = cudaMemsetAsync(a.lstar, 1, a.n_out, stream)) != cudaSuccess) return e;
	rank_mark(c, ln, 2, 1, 0);
	x3_rank_bytehist_kernel<<<rank_grid_for(s, (a.M + 65535) / 65536), 256, 0, stream>>>(a);
	/* level 2 is sorted from x by (b0, b1): digit b1 into buffer 1 -- which is also the level-1 order
	 * of the positions p + 1, tested (and, where the byte is rare, settled) by the first kernel --
	 * then digit b0 back into buffer 0 */
	rank_mark(c, ln, 0, 2, 0);
	const bool small_in = a.M < c.tune.small_below;
	if (small_in) {
		e = launch_pdl(x3_rank_radix_kernel<true, RS_ITEMS_SMALL>, rank_grid_for(s, rs_tiles_small, c.tune.rs_per_sm), RS_THREADS, 0, stream, c.pdl, a,
		               2, 0, ln.ticket, (uint32_t)ln.ticket + 1u);
	} else {
		e = launch_pdl(x3_rank_radix_kernel<true, RS_ITEMS>, rank_grid_for(s, rs_tiles, c.tune.rs_per_sm), RS_THREADS, 0, stream, c.pdl, a, 2, 0,
		               ln.ticket, (uint32_t)ln.ticket + 1u);
	}
	if (e != cudaSuccess) return e;
	++ln.ticket;
	rank_mark(c, ln, 1, 2, 0);
	if ((e = launch_pdl(x3_rank_first_kernel, rank_grid_for(s, f1_tiles, c.tune.rs_per_sm), F1_THREADS, 0, stream, c.pdl, a)) != cudaSuccess) return e;
	rank_mark(c, ln, 0, 2, 1);
	if (small_in) {
		e = launch_pdl(x3_rank_radix_kernel<false, RS_ITEMS_SMALL>, rank_grid_for(s, rs_tiles_small, c.tune.rs_per_sm), RS_THREADS, 0, stream, c.pdl, a,
		               2, 1, ln.ticket, (uint32_t)ln.ticket + 1u);
	} else {
		e = launch_pdl(x3_rank_radix_kernel<false, RS_ITEMS>, rank_grid_for(s, rs_tiles, c.tune.rs_per_sm), RS_THREADS, 0, stream, c.pdl, a, 2, 1,
		               ln.ticket, (uint32_t)ln.ticket + 1u);
	}
	if (e != cudaSuccess) return e;
	++ln.ticket;
	ln.nl += 4;
	ln.known = a.M;
	ln.L = 2;
	ln.spins = 0;
	ln.active = true;
	return cudaSuccess;
}

/* the chunk's last launch is queued */
cudaError_t rank_chunk_end(const RankCfg &c, RankLane &ln)
{
	cudaError_t e;
	RankScratch &s = *ln.s;
	ln.active = false;
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	if (c.profile) {
		rank_mark(c, ln, 3, 0, 0);
		if ((e = cudaStreamSynchronize(ln.stream)) != cudaSuccess) return e;
		RankCtrl *hc = (RankCtrl *)malloc(sizeof(RankCtrl));
		if (hc == nullptr) return cudaErrorMemoryAllocation;
		if ((e = cudaMemcpy(hc, s.ctrl, sizeof(RankCtrl), cudaMemcpyDeviceToHost)) != cudaSuccess) {
			free(hc);
			return e;
		}
		for (int i = 0; i + 1 < s.npev; ++i) {
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, s.pev[i], s.pev[i + 1]) != cudaSuccess) {
				continue;
			}
			/* elements the launch really processed, from the level sizes the device recorded */
			const int kind = s.pkind[i], lv = s.plevel[i];
			double el = hc->lv[lv].m;
			if (kind == 0 && lv > 1 && (hc->lv[lv].m < c.lim || s.ppass[i] >= radix_passes(hc->lv[lv].groups))) {
				el = 0; /* a pass that returned at once */
			}
			if (c.trace) {
				fprintf(stderr, "x3k_launch_rank: profile %-6s level %2d pass %d: %8.2f us, %9.0f elements\n",
				        kind == 0 ? "radix" : (kind == 1 ? "level" : "setup"), lv, s.ppass[i], ms * 1e3, el);
			}
			s.prof_ms[kind] += ms;
			s.prof_elems[kind] += el;
			s.prof_launches[kind] += el > 0 ? 1 : 0;
		}
		free(hc);
	}
	return cudaSuccess;
}

/* Queues as much of the lane's current chunk as can be queued without waiting.  *blocked: the
 * lane waits for the size report of a level that is still running. */
cudaError_t rank_chunk_step(const RankCfg &c, RankLane &ln, bool *blocked)
{
	cudaError_t e;
	RankScratch &s = *ln.s;
	RankArgs &a = ln.a;
	cudaStream_t stream = ln.stream;
	*blocked = false;
	for (;;) {
		const int L = ln.L;
		if (L == 2 && ln.known <= (uint32_t)TL_CAP && !c.no_tail) {
			/* a small input: every level from 2 on in one launch */
			rank_mark(c, ln, 1, L, 0);
			if ((e = launch_pdl(x3_rank_tail_kernel, 1, TL_THREADS, TL_SMEM, stream, c.pdl, a, L)) != cudaSuccess) return e;
			++ln.nl;
			return rank_chunk_end(c, ln);
		}
		if (L >= 2 + c.lag) {
			/* lv[L-lag+1] as level L-lag left it: its size bounds level L, and tells whether the levels
			 * queued since were the last ones */
			volatile uint32_t *rep = s.h_back + 4 * (L - c.lag);
			/* acquire: the size words are read only after the tag (the device orders its stores with
			 * __threadfence_system; a weakly ordered host CPU must not read them early) */
			if (__atomic_load_n(&rep[2], __ATOMIC_ACQUIRE) != a.seq) {
				if ((++ln.spins & 0xfffffu) == 0u) {
					/* a kernel that died would never report: do not wait on a failed stream */
					const cudaError_t q = cudaStreamQuery(stream);
					if (q != cudaErrorNotReady && __atomic_load_n(&rep[2], __ATOMIC_ACQUIRE) != a.seq) {
						return q == cudaSuccess ? cudaErrorUnknown : q;
					}
				}
				*blocked = true;
				return cudaSuccess;
			}
			ln.known = rep[0];
			if (c.trace) {
				fprintf(stderr, "x3k_launch_rank: lane %d chunk at %llu level %d: %u elements, %u groups\n", ln.index,
				        ln.a0, L - c.lag + 1, ln.known, s.h_back[4 * (L - c.lag) + 1]);
			}
			if (ln.known < c.lim) {
				return rank_chunk_end(c, ln);
			}
			if (ln.known <= (uint32_t)TL_CAP && !c.no_tail) {
				/* small enough for one CTA: every remaining level in one launch */
				rank_mark(c, ln, 1, L, 0);
				if ((e = launch_pdl(x3_rank_tail_kernel, 1, TL_THREADS, TL_SMEM, stream, c.pdl, a, L)) != cudaSuccess) return e;
				++ln.nl;
				return rank_chunk_end(c, ln);
			}
		}
		const uint32_t known = ln.known;
		rank_mark(c, ln, 1, L, 0);
		const int lgrid = rank_grid_for(s, (known + LV_TILE - 1) / LV_TILE, c.tune.lv_per_sm, LV_THREADS);
		if (L == 2) {
			e = launch_pdl(x3_rank_level_kernel<2>, lgrid, LV_THREADS, 0, stream, c.pdl, a, L, ln.ticket);
		} else if (L == 3) {
			e = launch_pdl(x3_rank_level_kernel<3>, lgrid, LV_THREADS, 0, stream, c.pdl, a, L, ln.ticket);
		} else {
			e = launch_pdl(x3_rank_level_kernel<4>, lgrid, LV_THREADS, 0, stream, c.pdl, a, L, ln.ticket);
		}
		if (e != cudaSuccess) return e;
		++ln.ticket;
		++ln.nl;
		if (L == 32) {
			return rank_chunk_end(c, ln);
		}
		/* the level-(L+1) keys have at most min(256^L, known) ranks: queue that many passes; a pass
		 * the real rank range does not need returns at once */
		const uint32_t rbound = L == 2 ? (known < 65536u ? known : 65536u) : known;
		const int np = radix_passes(rbound);
		for (int pass = 0; pass < np; ++pass) {
			rank_mark(c, ln, 0, L + 1, pass);
			/* the tile size only has to be the same within one pass */
			if (known < c.tune.small_below) {
				e = launch_pdl(x3_rank_radix_kernel<false, RS_ITEMS_SMALL>, rank_grid_for(s, (known + RS_TILE_SMALL - 1) / RS_TILE_SMALL, c.tune.rs_per_sm),
				               RS_THREADS, 0, stream, c.pdl, a, L + 1, pass, ln.ticket, (uint32_t)ln.ticket + 1u);
			} else {
				e = launch_pdl(x3_rank_radix_kernel<false, RS_ITEMS>, rank_grid_for(s, (known + RS_TILE - 1) / RS_TILE, c.tune.rs_per_sm), RS_THREADS, 0,
				               stream, c.pdl, a, L + 1, pass, ln.ticket, (uint32_t)ln.ticket + 1u);
			}
			if (e != cudaSuccess) return e;
			++ln.ticket;
			++ln.nl;
		}
		ln.L = L + 1;
	}
}

} /* namespace */

/* largest number of distances the rank search takes (chunks must keep room for positions) */
uint32_t x3k_rank_max_distances(void)
{
	return 1u << 23;
}

int x3k_rank_max_lanes(void)
{
	return RANK_MAX_LANES;
}

void x3k_rank_release(int dev)
{
	for (int lane = 0; lane < RANK_MAX_LANES; ++lane) {
		RankScratch &s = g_rank[dev][lane];
		for (int j = 0; j < 2; ++j) {
			cudaFree(s.key[j]);
			cudaFree(s.pos[j]);
		}
		cudaFree(s.st_level);
		cudaFree(s.st_radix);
		cudaFree(s.ctrl);
		cudaFreeHost(s.h_back);
		if (s.pev_made) {
			for (int i = 0; i < 520; ++i) {
				cudaEventDestroy(s.pev[i]);
			}
		}
		s = RankScratch();
	}
	RankLanes &ls = g_lanes[dev];
	if (ls.made) {
		for (int j = 0; j < RANK_MAX_LANES; ++j) {
			cudaStreamDestroy(ls.st[j]);
			cudaEventDestroy(ls.join[j]);
		}
		cudaEventDestroy(ls.fork);
	}
	ls = RankLanes();
}

/* Device time of the last profiled search on `dev` (X3_RANK_PROFILE=1): kind 0 = radix passes,
 * 1 = level kernels, 2 = set-up (byte histogram).  `elements` is what the launches were sized
 * for, an upper bound of what they processed. */
int x3k_rank_profile(int dev, int kind, double *ms, double *elements, int *launches)
{
	if (dev < 0 || dev >= 64 || kind < 0 || kind >= 4) {
		return -1;
	}
	const RankScratch &s = g_rank[dev][0];
	*ms = s.prof_ms[kind];
	*elements = s.prof_elems[kind];
	*launches = s.prof_launches[kind];
	return 0;
}

/* chunk length and count of a search over n positions on `lanes` lanes: chunks of equal length
 * (a multiple of 4096), as few as the 24-bit element format allows but at least one per lane */
void x3k_rank_chunking(unsigned long long n, uint32_t D, int lanes, unsigned long long *chunk, unsigned long long *count)
{
	const unsigned long long CHMAX = (unsigned long long)((RANK_MAX_M - (D < (1u << 23) ? D : (1u << 23))) & ~4095u);
	unsigned long long nch = (n + CHMAX - 1) / CHMAX;
	if (nch < (unsigned long long)lanes) {
		nch = (unsigned long long)lanes;
	}
	if (nch < 1) {
		nch = 1;
	}
	unsigned long long CH = (((n + nch - 1) / nch) + 4095ull) & ~4095ull;
	if (CH == 0) {
		CH = 4096;
	}
	*chunk = CH;
	*count = (n + CH - 1) / CH;
}

/*
 * Lstar by the rank method for positions [0, b.n) of the current device, chunk by chunk, on up to
 * RANK_MAX_LANES lanes at once: a lane is a stream of the caller's with scratch of its own, and
 * takes the next chunk whenever its last one is queued.  Everything the levels of a chunk need to
 * know about each other (sizes, buffers, whether the search is over) lives in device memory, so
 * the launches are simply queued; the host only reads the size of level L-lag back before it
 * queues level L (`lag` levels of work stay queued behind that wait), to size the grids and to
 * stop queueing once the search has ended.  While one lane waits for such a report the others are
 * fed: the mid-size levels of a chunk are one wave of tiles each -- latency, not throughput -- and
 * chunks in flight together fill each other's gaps (and one chunk's copies overlap another's kernels).
 */
cudaError_t x3k_launch_rank_batch(const X3RankBatch &b, uint32_t D, int t, int *launches)
{
	cudaError_t e;
	if (b.lanes < 1 || b.lanes > RANK_MAX_LANES || D > x3k_rank_max_distances()) {
		return cudaErrorNotSupported;
	}
	if (b.n == 0) {
		return cudaSuccess;
	}
	int dev = 0;
	if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
	if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
	RankCfg c;
	c.D = D;
	c.t = t;
	c.lim = (uint32_t)t + 2u;
	c.trace = getenv("X3_TRACE") != nullptr;
	/* the profile mode's events sit between the kernels of one stream: one lane only */
	c.profile = getenv("X3_RANK_PROFILE") != nullptr && b.lanes == 1;
	c.no_tail = getenv("X3_RANK_NO_TAIL") != nullptr; /* testing knob: never changes results */
	/* programmatic dependent launch between the kernels of a chunk; the profile mode keeps plain launches */
	c.pdl = getenv("X3_RANK_NO_PDL") == nullptr && !c.profile;
	c.lag = getenv("X3_RANK_LAG") != nullptr && atoi(getenv("X3_RANK_LAG")) >= 1 ? atoi(getenv("X3_RANK_LAG")) : 2;
	const auto wall0 = std::chrono::steady_clock::now();
	unsigned long long CH = 0, nch = 0;
	x3k_rank_chunking(b.n, D, b.lanes, &CH, &nch);
	const bool trivial = t <= 0 || D == 0; /* backend.c:76 never enters the selection / the window holds no
	                                        * distance: return 1 everywhere, a memset per chunk */

	RankLane lanes[RANK_MAX_LANES];
	for (int j = 0; j < b.lanes; ++j) {
		RankLane &ln = lanes[j];
		ln.b = &b;
		ln.stream = b.streams[j];
		ln.index = j;
		ln.s = &g_rank[dev][j];
		if (trivial || (unsigned long long)j >= nch) {
			continue;
		}
		if ((e = rank_ensure(dev, j, (uint32_t)((b.n < CH ? b.n : CH) + D))) != cudaSuccess) return e;
		RankScratch &s = *ln.s;
		if (c.profile && !s.pev_made) {
			for (int i = 0; i < 520; ++i) {
				if ((e = cudaEventCreate(&s.pev[i])) != cudaSuccess) return e;
			}
			s.pev_made = true;
		}
		for (int k = 0; k < 4; ++k) {
			s.prof_ms[k] = s.prof_elems[k] = 0;
			s.prof_launches[k] = 0;
		}
	}
	if (c.trace) {
		fprintf(stderr, "x3k_launch_rank: %llu chunk(s) of %llu positions on %d lane(s), scratch ready after %.3f ms\n", nch, CH,
		        b.lanes, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
	}
	unsigned long long next = 0; /* chunks are handed out in order */
	int open = b.lanes;
	while (open > 0) {
		bool progress = false;
		for (int j = 0; j < b.lanes; ++j) {
			RankLane &ln = lanes[j];
			if (ln.finished) {
				continue;
			}
			if (!ln.active) {
				if (next >= nch) {
					ln.finished = true;
					--open;
					progress = true;
					continue;
				}
				const unsigned long long a0 = next * CH, len = b.n - a0 < CH ? b.n - a0 : CH;
				++next;
				progress = true;
				if (b.before_chunk != nullptr && (e = b.before_chunk(b.ctx, j, a0, len)) != cudaSuccess) return e;
				if (trivial) {
					if ((e = cudaMemsetAsync(b.lstar + a0, 0, len, ln.stream)) != cudaSuccess) return e;
					ln.a0 = a0;
					ln.a.n_out = (uint32_t)len;
				} else if ((e = rank_chunk_begin(c, ln, a0, len)) != cudaSuccess) {
					return e;
				}
			}
			bool blocked = false;
			if (ln.active && (e = rank_chunk_step(c, ln, &blocked)) != cudaSuccess) return e;
			if (!blocked) {
				progress = true;
			}
			if (!ln.active && b.after_chunk != nullptr) {
				/* the chunk's last kernel is queued on the lane's stream */
				if ((e = b.after_chunk(b.ctx, j, ln.a0, ln.a.n_out)) != cudaSuccess) return e;
			}
		}
		if (!progress) {
#if defined(__x86_64__) || defined(__i386__)
			__builtin_ia32_pause(); /* the reports are a few microseconds away: spin politely */
#endif
		}
	}
	int nl = 0;
	for (int j = 0; j < b.lanes; ++j) {
		nl += lanes[j].nl;
	}
	if (launches != nullptr) {
		*launches += nl;
	}
	if (c.trace) {
		fprintf(stderr, "x3k_launch_rank: %d launches queued after %.3f ms\n", nl,
		        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
	}
	return cudaGetLastError();
}

/* Lanes a search over n positions is split into when the caller does not say (X3_RANK_LANES
 * overrides): one per chunk the input needs anyway, at most RANK_MAX_LANES.  Measured on B200
 * (tests/gpu_lanes_ab.py): the chain of ~40 launches of ONE chunk does not get shorter when the chunk
 * is halved (C2: 1.00 -> 0.98 ms with twice the launches), but the chains of different chunks overlap
 * well (60 MB of C5: 12.2 -> 9.0 ms with 4 lanes). */
int x3k_rank_default_lanes(unsigned long long n, uint32_t D)
{
	const char *v = getenv("X3_RANK_LANES");
	unsigned long long lanes;
	if (v != nullptr && atoi(v) >= 1) {
		lanes = (unsigned long long)atoi(v);
	} else {
		const unsigned long long chmax = (unsigned long long)((RANK_MAX_M - (D < (1u << 23) ? D : (1u << 23))) & ~4095u);
		lanes = (n + chmax - 1) / chmax;
		if (lanes > (unsigned long long)RANK_DEFAULT_LANES) {
			lanes = RANK_DEFAULT_LANES;
		}
	}
	if (getenv("X3_RANK_PROFILE") != nullptr) {
		lanes = 1; /* the profile mode's events bracket the launches of one stream */
	}
	if (lanes > (unsigned long long)RANK_MAX_LANES) {
		lanes = RANK_MAX_LANES;
	}
	/* lanes start at multiples of 4096 positions; tiny inputs are not split */
	while (lanes > 1 && n / lanes < 65536ull) {
		--lanes;
	}
	return lanes < 1 ? 1 : (int)lanes;
}

/*
 * Lstar for positions [0, prm.n) of a device-resident buffer, ordered behind `stream` and
 * complete when `stream` reaches the point where this returns: the position range is cut into
 * lanes that run on streams of their own, forked from and joined back into `stream`.
 */
cudaError_t x3k_launch_rank(const X3SearchParams &prm, cudaStream_t stream, int *launches)
{
	cudaError_t e;
	if (prm.H != nullptr || prm.D > x3k_rank_max_distances()) {
		return cudaErrorNotSupported;
	}
	if (prm.n == 0) {
		return cudaSuccess;
	}
	const int lanes = x3k_rank_default_lanes(prm.n, prm.D);
	X3RankBatch b;
	b.x = prm.x;
	b.lstar = prm.lstar;
	b.n = prm.n;
	b.lanes = lanes;
	b.before_chunk = nullptr;
	b.after_chunk = nullptr;
	b.ctx = nullptr;
	if (lanes == 1) {
		b.streams[0] = stream;
		return x3k_launch_rank_batch(b, prm.D, prm.t, launches);
	}
	int dev = 0;
	if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
	if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
	RankLanes &ls = g_lanes[dev];
	if (!ls.made) {
		if ((e = cudaEventCreateWithFlags(&ls.fork, cudaEventDisableTiming)) != cudaSuccess) return e;
		for (int j = 0; j < RANK_MAX_LANES; ++j) {
			if ((e = cudaStreamCreateWithFlags(&ls.st[j], cudaStreamNonBlocking)) != cudaSuccess) return e;
			if ((e = cudaEventCreateWithFlags(&ls.join[j], cudaEventDisableTiming)) != cudaSuccess) return e;
		}
		ls.made = true;
	}
	if ((e = cudaEventRecord(ls.fork, stream)) != cudaSuccess) return e;
	for (int j = 0; j < lanes; ++j) {
		b.streams[j] = ls.st[j];
		if ((e = cudaStreamWaitEvent(ls.st[j], ls.fork, 0)) != cudaSuccess) return e;
	}
	e = x3k_launch_rank_batch(b, prm.D, prm.t, launches);
	/* the caller's stream continues behind every lane, also when queueing failed half way */
	for (int j = 0; j < lanes; ++j) {
		if (cudaEventRecord(ls.join[j], ls.st[j]) == cudaSuccess) {
			cudaStreamWaitEvent(stream, ls.join[j], 0);
		}
	}
	return e;
}
