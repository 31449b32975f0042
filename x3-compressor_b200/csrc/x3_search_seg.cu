/*
 * x3_search_seg.cu -- "segment" search: the forward-window search of reference backend.c:58-78
 * for windows that fit on chip (D = W - 33 <= SG_DMAX), sm_100a.
 *
 * Same occurrence-rank formulation as x3_search_rank.cu -- count_L(p) > t  <=>  the (t+1)-th next
 * occurrence of the L-gram x[p..p+L) starts within D bytes of p; Lstar(p) = the deepest level p
 * passes -- but nothing except the input and the result ever touches HBM.  Lstar of a position
 * depends on x[p .. p+W-2] only, so a CTA takes a SEGMENT of B consecutive positions together with
 * everything they can see, as M = B + D + 3 <= 32768 elements (element e = position a - 3 + e,
 * 15-bit ids), and runs every level in its 227 KB of shared memory:
 *
 *   load      the segment's bytes with one TMA bulk copy (cp.async.bulk + mbarrier, SASS UBLKCP)
 *   levels 1-4  four stable counting sorts of all M elements on the bytes x[p+3], x[p+2], x[p+1],
 *             x[p] (LSD).  After the pass on x[p+j] the order is by (x[p+j..p+3], p): the
 *             level-(4-j) order of the positions q = p + j, so the NEXT pass tests that level
 *             while it reads its input (look-ahead of t+1 by one warp shuffle per value).  Ranking
 *             inside a round: the lanes with lane 0's digit by one vote, the others by shared-memory
 *             atomicOr into a mask row (MATCH.ANY costs 2 cycles per distinct value, SM-wide); two
 *             sub-blocks per warp in flight; offsets by all 1024 threads.  The positions with a rare
 *             first byte (c1 <= t followers: tc* = c1 - 1 < t) are marked in a bitmap and settled
 *             behind the ranking, every thread taking its share, by walking those followers
 *             (Lstar = min LCP32, 0 when c1 < 2: backend.c:76-78).
 *   level >= 4  the array is a set of GROUPS (runs of equal L-grams in position order).  A group
 *             with fewer than t+2 elements can never pass again and is dropped.  A group is
 *             tested, pruned to the elements within D behind a passed one (the only followers
 *             that can matter deeper down; any superset gives the same table), and split by the
 *             next byte x[p+L] with a stable counting sort inside its own index range.  Groups are
 *             independent; by size: up to 256 elements one warp follows the group down to the
 *             level where it ends (CHAINS: registers, a shared-memory queue, no CTA barrier); up
 *             to 7 936 elements WAVES of up to 32 rows of 256 elements, one warp per row, four
 *             CTA barriers per level; above that the whole CTA, one group at a time.  A group
 *             that does not change at a level (all kept, one next byte) JUMPS to the level where
 *             its bytes first differ.
 *   store     Lstar of the B positions, kept in shared memory (the deepest level passed so far).
 *
 * HBM traffic is the algorithmic minimum: every input byte read once (plus the D-byte halo per
 * segment: (B+D)/B = 1.33 at the default window), every Lstar byte written once.  One launch per
 * search, no level reports to the host, no chained scans between CTAs; a persistent grid of one
 * CTA per SM draws segments from a counter (the last, partly filled wave in thirds or quarters of
 * a segment; a launch may also take every parts-th PIECE of the input: x3s_search_device_part).
 * Any t >= x3k_seg_min_t() (queue capacity M / (t+2)); other parameters go to the rank search.
 * Lstar only (the 32-bin table H is the brute-force kernels' job).  tests/seg_model.py states the
 * same rules in numpy; DESIGN.md section 4.1 has the measurements.
 */
#include "x3_search_device.cuh"

#include <cstdio>
#include <cstdlib>

namespace {

constexpr int SG_THREADS = 1024;
constexpr int SG_WARPS = SG_THREADS / 32;
constexpr uint32_t SG_MMAX = 32768;          /* elements of a segment: 15-bit ids */
constexpr uint32_t SG_BMAX = 24592;          /* searched positions of a segment (Lstar staging) */
constexpr uint32_t SG_XS_OFF = 13;           /* xs[k] = xsr[SG_XS_OFF + k]: the TMA source is 16-byte aligned */
constexpr uint32_t SG_XS_BYTES = SG_MMAX + 80;
constexpr uint32_t SG_CHUNK = 256;           /* elements a warp takes per wave: one histogram row */
constexpr uint32_t SG_CHAIN_MAX = SG_CHUNK;  /* groups up to this size are followed down by one warp (8 elements per lane) */
constexpr uint32_t SG_WAVE_MAX = 31 * SG_CHUNK; /* groups up to this size go through the waves (31 rows + the totals row) */
constexpr uint32_t SG_Q = 4864;              /* queue slots of the chains: live groups <= M / (t+2) */
constexpr uint32_t SG_RING = 128;            /* ring slots of the waves: live groups above SG_CHAIN_MAX <= M / 257 */
constexpr uint32_t SG_BIGQ = 16;             /* live groups above SG_WAVE_MAX: <= M / SG_WAVE_MAX */
constexpr uint32_t SG_NONE = 0xffffffffu;
constexpr uint32_t Q_VALID = 0x80000000u;

/* a group of a wave: rows [r0, r0 + nrows) of the histogram, what the ring entry said, and what
 * became of it (state: 0 = its elements are placed, 1 = nothing to place) */
struct SegGroup {
	uint32_t start, len, L, buf, r0, nrows, state, pad;
};

struct SegMisc {
	unsigned long long mbar;
	uint32_t dbase[256];
	uint32_t wsum[36];
	uint32_t lastp[32];
	uint32_t q_head, q_tail, q_done; /* chains: tickets taken, entries pushed, entries finished */
	uint32_t r_head, r_tail;
	uint32_t big_head, big_tail;
	uint32_t seg;
	uint32_t wave_groups, wave_rows;
	uint2 big[SG_BIGQ];
	uint2 ring[SG_RING];
	SegGroup grp[16];            /* (a group of a wave has at least 2 rows) */
};

constexpr size_t SG_OFF_P0 = SG_XS_BYTES;
constexpr size_t SG_OFF_P1 = SG_OFF_P0 + SG_MMAX * 2;
constexpr size_t SG_OFF_L8 = SG_OFF_P1 + SG_MMAX * 2;
constexpr size_t SG_OFF_WH = SG_OFF_L8 + SG_BMAX + 16;
constexpr size_t SG_OFF_QE = SG_OFF_WH + SG_WARPS * 256 * 2;
constexpr size_t SG_OFF_QL = SG_OFF_QE + SG_Q * 4;
constexpr size_t SG_OFF_MISC = (SG_OFF_QL + SG_Q + 15) & ~(size_t)15;
constexpr size_t SG_SMEM = SG_OFF_MISC + sizeof(SegMisc);
static_assert(SG_SMEM + 128 <= 232448, "segment kernel: shared memory over the 227 KB a CTA can have");
static_assert(SG_XS_BYTES % 16 == 0 && SG_OFF_L8 % 16 == 0 && SG_OFF_WH % 16 == 0, "alignment");

struct SegArgs {
	const uint8_t *x;            /* 16-byte aligned, xbytes readable */
	uint8_t *lstar;
	unsigned long long n;        /* positions [0, n) */
	unsigned long long xbytes;   /* multiple of 16 */
	uint32_t D, B;               /* distances; positions per segment (multiple of 16) */
	int t;
	uint32_t nseg;               /* segments of this launch */
	uint32_t part, parts, spp;   /* spp != 0: the launch takes pieces part, part + parts, ... of spp segments each, */
	uint32_t k0;                 /* from the part's k0-th piece on */
	uint32_t split_first, split_k, split_B; /* tickets from split_first on are split_k-ths of a segment (split_B positions) */
	unsigned int *ticket;        /* zeroed before the launch */
	unsigned long long *prof;    /* NULL, or 9 counters per CTA (X3_SEG_PROF=1): cycles of load, pass 0, passes 1-3,
	                              * level-4 groups, big groups, waves, chains, store; segments */
};

/* The CTA's shared memory, addressed through the array itself in every device function, so that the
 * compiler keeps the accesses in the shared window (LDS/STS, 32-bit addresses). */
extern __shared__ __align__(128) uint8_t sg_smem[];
#define SG_XSB (sg_smem)
#define SG_XSW (reinterpret_cast<const uint32_t *>(sg_smem))
#define SG_P(buf) (reinterpret_cast<uint16_t *>(sg_smem + SG_OFF_P0) + (size_t)(buf) * SG_MMAX)
#define SG_L8 (sg_smem + SG_OFF_L8)
#define SG_WH (reinterpret_cast<uint16_t *>(sg_smem + SG_OFF_WH))
#define SG_QENT (reinterpret_cast<volatile uint32_t *>(sg_smem + SG_OFF_QE))
#define SG_QLVL (reinterpret_cast<volatile uint8_t *>(sg_smem + SG_OFF_QL))
#define SG_MI (reinterpret_cast<SegMisc *>(sg_smem + SG_OFF_MISC))

/* what a segment's phases share besides the shared memory: scalars only (registers) */
struct SegCtx {
	uint32_t M, Bs, D, la;
	int t;
};

/* the 4 bytes x[p .. p+3] of element e, byte 0 in the low bits */
__device__ __forceinline__ uint32_t sg_gram(uint32_t e)
{
	const uint32_t off = SG_XS_OFF + e;
	const uint32_t w = off >> 2;
	return __funnelshift_r(SG_XSW[w], SG_XSW[w + 1], (off & 3u) * 8u);
}

/* what pass K of the LSD sort needs of element e: its digit x[p+3-K] in the low byte, the K bytes behind it above
 * (NB = K + 1 bytes from x[p+3-K]); the second word is only fetched by the lanes whose bytes straddle a word */
template <int K>
__device__ __forceinline__ uint32_t sg_gram_from(uint32_t e)
{
	const uint32_t off = SG_XS_OFF + e + 3u - (uint32_t)K;
	const uint32_t w = off >> 2, sh = off & 3u;
	const uint32_t lo = SG_XSW[w];
	const uint32_t hi = sh + (uint32_t)K + 1u > 4u ? SG_XSW[w + 1] : 0u;
	return __funnelshift_r(lo, hi, sh * 8u);
}

__device__ __forceinline__ uint32_t sg_byte(uint32_t e, uint32_t L)
{
	return SG_XSB[SG_XS_OFF + e + L];
}

/* ---- offsets of a CTA-wide counting sort: wh[w][d] = count of digit d in warp w's block  ->
 * wh[w][d] = elements of digit d in the blocks in front of w's, dbase[d] = first slot of digit d.
 * `drop_below`: digits with fewer elements than this get no slots (dbase = SG_NONE); returns the
 * total count of digit tid (threads 0..255) for the caller. */
__device__ __forceinline__ uint32_t sg_offsets(const SegCtx &c, uint32_t drop_below)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t run = 0;
	if (tid < 256) {
#pragma unroll 8
		for (int w = 0; w < SG_WARPS; ++w) {
			const uint32_t v = SG_WH[w * 256 + tid];
			SG_WH[w * 256 + tid] = (uint16_t)run;
			run += v;
		}
	}
	const uint32_t mine = run >= drop_below ? run : 0u;
	uint32_t inc = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= d) {
			inc += o;
		}
	}
	if (lane == 31 && warp < 8) {
		SG_MI->wsum[warp] = inc;
	}
	__syncthreads();
	if (tid < 256) {
		uint32_t before = 0;
		for (int w = 0; w < warp; ++w) {
			before += SG_MI->wsum[w];
		}
		SG_MI->dbase[tid] = run >= drop_below && run > 0 ? before + inc - mine : SG_NONE;
	}
	return run;
}

/* ---- level 1, the rare first bytes: element i of the order by (x[p+3], p) stands for q = p + 3
 * and has fewer than t+1 followers with its byte within D.  Lstar = the smallest LCP32 over
 * them, 0 when there are fewer than 2 (backend.c:76-78 collapsed for tc* = c1 - 1). */
__device__ __noinline__ uint32_t sg_rare(uint32_t inbuf, uint32_t i, uint32_t e, uint32_t M, uint32_t t, uint32_t D)
{
	const uint8_t *xq = SG_XSB + SG_XS_OFF + 3;
	const uint16_t *In = SG_P(inbuf);
	const uint32_t b = xq[e];
	uint32_t c1 = 0, best = 32;
	for (uint32_t j = i + 1; j < M && c1 < t; ++j) {
		const uint32_t ej = In[j];
		if (xq[ej] != b || ej - e > D) {
			break;
		}
		++c1;
		uint32_t l = 1;
		while (l < best && xq[e + l] == xq[ej + l]) {
			++l;
		}
		best = l;
	}
	return c1 >= 2 ? best : 0u;
}

/* ---- one LSD pass over all M elements: stable counting sort on byte 3-K of the gram.
 * K >= 1: the input order is the level-K order of the positions q = p + 4 - K; it is tested on
 * the way (LA32: the look-ahead t+1 <= 32 comes from the neighbouring lanes by shuffle). */
template <int K, bool LA32, int INBUF>
__device__ __forceinline__ void sg_lsd_pass(const SegCtx &c)
{
	const uint16_t *__restrict__ In = SG_P(INBUF);
	uint16_t *__restrict__ Out = SG_P(INBUF ^ 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t M = c.M;
	const uint32_t R = (M + 1023u) >> 10;      /* rounds per warp; a warp's block is 32 R elements */
	const uint32_t blk = (uint32_t)warp * 32u * R;
	uint16_t *myh = SG_WH + warp * 256;
	/* the warp's mask row: 256 words in the output buffer, which nobody writes before the placement */
	uint32_t *mym = reinterpret_cast<uint32_t *>(SG_P(INBUF ^ 1)) + warp * 256;
	const uint32_t lt = (1u << lane) - 1u;
	reinterpret_cast<uint4 *>(myh)[lane] = make_uint4(0u, 0u, 0u, 0u);
	reinterpret_cast<uint4 *>(mym)[lane] = make_uint4(0u, 0u, 0u, 0u);
	reinterpret_cast<uint4 *>(mym)[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
	__syncwarp();
	/* rank of my element of round r among its digit in the warp's block (< 1024): 3 per word.  (Rolled loops
	 * with this array in local memory were measured: 12 % slower -- the unrolled rounds overlap.) */
	uint32_t rk[11];
#pragma unroll
	for (int k = 0; k < 11; ++k) {
		rk[k] = 0;
	}
	uint32_t e_n = 0, g_n = 0;
	if (K >= 1 && blk + lane < M) {
		e_n = In[blk + lane];
		g_n = sg_gram(e_n);
	}
#pragma unroll
	for (int r = 0; r < 32; ++r) {
		if ((uint32_t)r < R) {
			const uint32_t i = blk + 32u * r + lane;
			const bool valid = i < M;
			const uint32_t nvalid = M > blk + 32u * r ? M - (blk + 32u * r) : 0u; /* lanes of the round with an element */
			uint32_t d;
			if (K == 0) {
				d = valid ? SG_XSB[SG_XS_OFF + i + 3] : 0u;
			} else {
				const uint32_t e = e_n, g = g_n;
				if (i + 32u < M) {
					e_n = In[i + 32u];
					g_n = sg_gram(e_n);
				}
				d = (g >> (8 * (3 - K))) & 255u;
				/* level K on the input order: the element t+1 places further on */
				uint32_t ef, gf;
				if (LA32) {
					const int sl = (lane + (int)c.la) & 31;
					const uint32_t e1 = __shfl_sync(FULL_MASK, e, sl), e2 = __shfl_sync(FULL_MASK, e_n, sl);
					const uint32_t g1 = __shfl_sync(FULL_MASK, g, sl), g2 = __shfl_sync(FULL_MASK, g_n, sl);
					const bool here = lane + (int)c.la < 32;
					ef = here ? e1 : e2;
					gf = here ? g1 : g2;
				} else {
					ef = 0;
					gf = 0;
					if (i + c.la < M) {
						ef = In[i + c.la];
						gf = sg_gram(ef);
					}
				}
				const uint32_t q = e + 1u - (uint32_t)K;   /* searched position (relative to the segment) */
				const bool subj = valid && q < c.Bs;
				const bool pass = subj && i + c.la < M && ((g ^ gf) >> (K == 0 ? 0 : 32 - 8 * K)) == 0u && ef - e <= c.D;
				if (pass) {
					SG_L8[q] = (uint8_t)K;
				} else if (K == 1 && subj) {
					SG_L8[q] = (uint8_t)sg_rare(INBUF, i, e, M, (uint32_t)c.t, c.D);
				}
			}
			/* which lanes of the round hold my digit.  MATCH.ANY takes 2 cycles per distinct value of the warp,
			 * SM-wide (profiles/r2_ubench.json: 32 cycles on text, 64 on noise), so it only serves rounds
			 * with few values; the others set their lane bits in the warp's mask row (shared-memory atomics:
			 * the row ends up the same in whatever order they are served) and read the row back. */
			const uint32_t vm = nvalid >= 32u ? FULL_MASK : (1u << nvalid) - 1u;
			const uint32_t d0 = __shfl_sync(FULL_MASK, d, 0); /* (every lane takes part: not inside the && below) */
			const uint32_t m0 = __ballot_sync(FULL_MASK, valid && d == d0);
			uint32_t peers;
			if (m0 == vm) {
				peers = vm;
			} else if (__popc(m0) >= 11) {
				peers = __match_any_sync(FULL_MASK, valid ? d : 256u);
			} else {
				if (valid) {
					atomicOr(&mym[d], 1u << lane);
				}
				__syncwarp();
				peers = valid ? mym[d] : 0u;
				__syncwarp();
				if (valid && (peers & lt) == 0u) {
					mym[d] = 0u;
				}
			}
			const uint32_t old = valid ? (uint32_t)myh[d] : 0u;
			__syncwarp();
			if (valid && (peers & lt) == 0u) {
				myh[d] = (uint16_t)(old + __popc(peers));
			}
			rk[r / 3] |= (old + __popc(peers & lt)) << (10 * (r % 3));
			__syncwarp();
		}
	}
	__syncthreads();
	sg_offsets(c, 0u);
	__syncthreads();
#pragma unroll
	for (int r = 0; r < 32; ++r) {
		if ((uint32_t)r < R) {
			const uint32_t i = blk + 32u * r + lane;
			if (i < M) {
				const uint32_t rank = (rk[r / 3] >> (10 * (r % 3))) & 1023u;
				const uint32_t e = K == 0 ? i : (uint32_t)In[i];
				const uint32_t d = SG_XSB[SG_XS_OFF + e + 3 - K]; /* the digit again: byte 3-K of the gram */
				Out[SG_MI->dbase[d] + myh[d] + rank] = (uint16_t)e;
			}
		}
	}
	__syncthreads();
}

/* ---- the LSD pass, two sub-blocks per warp in flight (the production form; sg_lsd_pass above is the
 * one-block statement of the same pass, kept for -DX3_SEG_LSD1 comparisons).  A round of the pass is one
 * long dependent chain (load, gram, shuffles, atomics, read back, histogram update) and a CTA has only 8
 * warps per scheduler to hide it with, so every warp works on TWO sub-blocks at once -- sub-block s =
 * elements [s * 32 RH, (s + 1) * 32 RH), s = 2 warp and 2 warp + 1, RH = ceil(M / 2048) rounds each -- with
 * a histogram row and a mask row per sub-block: 64 histogram rows (the second 32 lie over the chains'
 * queue, which is empty during the passes and cleared again behind them), 64 mask rows = the whole
 * output buffer.  The offsets come out absolute (first slot of sub-block s's elements of digit d), by all
 * 1024 threads: 16 rows each. */
constexpr int SG_SUB = 64;
/* The level-1 pass only MARKS the positions with a rare first byte -- one bit per element of the level-1
 * order, a round's vote is a word of the map -- and every thread takes its share of them behind the
 * ranking (the k-th marked element goes to thread k mod 1024, found through a prefix sum over the words).
 * Walked where they are met they sit in one or two warps' blocks (the order is by that byte) and hold up the
 * whole CTA; and a call in the unrolled rounds keeps the compiler from overlapping them (the level-1 pass
 * took 98 K cycles of a text segment's 330 K where levels 2 and 3 take 30 K each).  Map and prefix sums lie
 * behind the 64 histogram rows, over the rest of the chains' queue. */
constexpr size_t SG_OFF_RARE = SG_OFF_WH + SG_SUB * 256 * 2;
static_assert(SG_OFF_RARE % 4 == 0 && SG_OFF_RARE + 4096 + 2048 <= SG_OFF_MISC, "rare map");
#define SG_RAREMAP (reinterpret_cast<uint32_t *>(sg_smem + SG_OFF_RARE))
#define SG_RAREPRE (reinterpret_cast<uint16_t *>(sg_smem + SG_OFF_RARE + 4096))

template <int INBUF>
__device__ __forceinline__ void sg_offsets64()
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t d = (uint32_t)tid & 255u, part = (uint32_t)tid >> 8;
	uint32_t *scr = reinterpret_cast<uint32_t *>(SG_P(INBUF ^ 1)); /* the mask rows: all zero again, free until the placement */
	uint16_t *col = SG_WH + (16u * part) * 256u + d;
	uint32_t pre[16];
	uint32_t run = 0;
#pragma unroll
	for (int k = 0; k < 16; ++k) {
		pre[k] = col[k * 256];
	}
#pragma unroll
	for (int k = 0; k < 16; ++k) {
		const uint32_t v = pre[k];
		pre[k] = run;
		run += v;
	}
	scr[part * 256u + d] = run;
	__syncthreads();
	if (tid < 256) {
		const uint32_t p0 = scr[d], p1 = scr[256u + d], p2 = scr[512u + d], p3 = scr[768u + d];
		const uint32_t tot = p0 + p1 + p2 + p3;
		uint32_t inc = tot;
#pragma unroll
		for (int s = 1; s < 32; s <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, inc, s);
			if (lane >= s) {
				inc += o;
			}
		}
		if (lane == 31) {
			SG_MI->wsum[warp] = inc;
		}
		asm volatile("bar.sync 1, 256;" ::: "memory"); /* the 8 warps of the digits only */
		uint32_t before = 0;
		for (int w = 0; w < warp; ++w) {
			before += SG_MI->wsum[w];
		}
		const uint32_t start = before + inc - tot;
		scr[1024u + d] = start;
		scr[1280u + d] = start + p0;
		scr[1536u + d] = start + p0 + p1;
		scr[1792u + d] = start + p0 + p1 + p2;
	}
	__syncthreads();
	const uint32_t base = scr[1024u + part * 256u + d];
#pragma unroll
	for (int k = 0; k < 16; ++k) {
		col[k * 256] = (uint16_t)(base + pre[k]);
	}
	__syncthreads();
}

template <int K, bool LA32, int INBUF, bool PROF>
__device__ __forceinline__ void sg_lsd_pass2(const SegCtx &c, unsigned long long *gp)
{
	/* measurement build: cycles of the pass's parts (ranking, offsets, placement) into the CTA's global counters 16 + 4 K .. */
	unsigned long long pt0 = 0, wstart = 0;
	if (PROF && threadIdx.x == 0) {
		pt0 = clock64();
	}
	if (PROF && K == 1) {
		wstart = clock64();
	}
#define SG_PLAP(k)                                 \
	if (PROF && threadIdx.x == 0) {                \
		const unsigned long long now = clock64();  \
		gp[16 + 4 * K + (k)] += now - pt0;         \
		pt0 = now;                                 \
	}
	const uint16_t *__restrict__ In = SG_P(INBUF);
	uint16_t *__restrict__ Out = SG_P(INBUF ^ 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t M = c.M;
	const uint32_t RH = (M + 2047u) >> 11;     /* rounds per sub-block */
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t bit = 1u << lane;
	uint32_t blk[2];
	uint16_t *myh[2];
	uint32_t *mym[2];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const uint32_t sub = 2u * (uint32_t)warp + h;
		blk[h] = sub * 32u * RH;
		myh[h] = SG_WH + sub * 256u;
		mym[h] = reinterpret_cast<uint32_t *>(SG_P(INBUF ^ 1)) + sub * 256u;
		reinterpret_cast<uint4 *>(myh[h])[lane] = make_uint4(0u, 0u, 0u, 0u);
		reinterpret_cast<uint4 *>(mym[h])[lane] = make_uint4(0u, 0u, 0u, 0u);
		reinterpret_cast<uint4 *>(mym[h])[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
	}
	__syncwarp();
	uint32_t rk[2][6]; /* rank of my element of round r of sub-block h among its digit in the sub-block (< 512): 3 per word */
#pragma unroll
	for (int h = 0; h < 2; ++h) {
#pragma unroll
		for (int k = 0; k < 6; ++k) {
			rk[h][k] = 0;
		}
	}
	uint32_t e_n[2] = {0u, 0u}, g_n[2] = {0u, 0u};
	if (K >= 1) {
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			if (blk[h] + lane < M) {
				e_n[h] = In[blk[h] + lane];
				g_n[h] = sg_gram_from<K>(e_n[h]);
			}
		}
	}
#pragma unroll
	for (int r = 0; r < 16; ++r) {
		if ((uint32_t)r < RH) {
			uint32_t d[2], vm[2], m0[2];
			bool valid[2];
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const uint32_t i0 = blk[h] + 32u * r, i = i0 + lane;
				valid[h] = i < M;
				const uint32_t nvalid = M > i0 ? M - i0 : 0u; /* lanes of the round with an element */
				vm[h] = nvalid >= 32u ? FULL_MASK : (1u << nvalid) - 1u;
				if (K == 0) {
					d[h] = valid[h] ? SG_XSB[SG_XS_OFF + i + 3] : 0u;
				} else {
					const uint32_t e = e_n[h], g = g_n[h];
					if (i + 32u < M) {
						e_n[h] = In[i + 32u];
						g_n[h] = sg_gram_from<K>(e_n[h]);
					}
					d[h] = g & 255u;
					/* level K on the input order: the element t+1 places further on */
					uint32_t ef, gf;
					if (LA32) {
						/* lane j wants the element la places on: lane (j + la) & 31 of this round while j + la < 32,
						 * of the next round otherwise -- so every lane s serves exactly one other, with this round's
						 * element when s >= la and the next round's when s < la: one shuffle per value */
						const int sl = (lane + (int)c.la) & 31;
						const bool mine_now = (uint32_t)lane >= c.la;
						ef = __shfl_sync(FULL_MASK, mine_now ? e : e_n[h], sl);
						gf = __shfl_sync(FULL_MASK, mine_now ? g : g_n[h], sl);
					} else {
						ef = 0;
						gf = 0;
						if (i + c.la < M) {
							ef = In[i + c.la];
							gf = sg_gram_from<K>(ef);
						}
					}
					const uint32_t q = e + 1u - (uint32_t)K;   /* searched position (relative to the segment) */
					const bool subj = valid[h] && q < c.Bs;
					const bool pass = subj && i + c.la < M && (((g ^ gf) & (K >= 3 ? 0xffffffffu : (1u << (8 * (K + 1))) - 1u)) >> 8) == 0u && ef - e <= c.D;
					if (pass) {
						SG_L8[q] = (uint8_t)K;
					}
					if (K == 1) {
						/* a rare first byte: marked in the map (this round's word), walked behind the ranking */
						const uint32_t rm = __ballot_sync(FULL_MASK, subj && !pass);
						if (lane == 0) {
							SG_RAREMAP[i0 >> 5] = rm;
						}
					}
				}
				/* the lanes with lane 0's digit know their peers from one vote */
				const uint32_t d0 = __shfl_sync(FULL_MASK, d[h], 0);
				m0[h] = __ballot_sync(FULL_MASK, valid[h] && d[h] == d0);
			}
			/* the others set their lane bits in the sub-block's mask row (shared-memory atomics: the row ends
			 * up the same in whatever order they are served; MATCH.ANY would take 2 cycles per distinct value,
			 * SM-wide: profiles/r2_ubench.json) and read the row back */
			bool mine0[2];
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				mine0[h] = (m0[h] >> lane) & 1u;
				if (m0[h] != vm[h] && valid[h] && !mine0[h]) {
					atomicOr(&mym[h][d[h]], bit);
				}
			}
			__syncwarp();
			uint32_t peers[2], old[2];
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				peers[h] = mine0[h] ? m0[h] : (valid[h] ? mym[h][d[h]] : 0u);
				old[h] = valid[h] ? (uint32_t)myh[h][d[h]] : 0u;
			}
			__syncwarp();
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				if (valid[h] && (peers[h] & lt) == 0u) {
					if (!mine0[h]) {
						mym[h][d[h]] = 0u;
					}
					myh[h][d[h]] = (uint16_t)(old[h] + __popc(peers[h]));
				}
				rk[h][r / 3] |= (old[h] + __popc(peers[h] & lt)) << (10 * (r % 3));
			}
			__syncwarp();
		}
	}
	if (PROF && K == 1 && (threadIdx.x & 31) == 0) {
		/* measurement build: how long each warp spent in the level-1 ranking: the slowest (24) and the sum (25) */
		const unsigned long long dur = clock64() - wstart;
		atomicMax(&gp[46], dur);
		atomicAdd(&gp[47], dur);
	}
	__syncthreads();
	if (K == 1) {
		if (PROF && threadIdx.x == 0) {
			const unsigned long long now = clock64();
			gp[16 + 3] += now - pt0; /* K = 0's spare slot: the level-1 pass's ranking alone */
			pt0 = now;
		}
		/* words of the map: one per thread (M <= 32768); rounds beyond M never wrote theirs */
		const uint32_t nwords = 64u * RH;
		const uint32_t word = (uint32_t)tid < nwords && 32u * (uint32_t)tid < M ? SG_RAREMAP[tid] : 0u;
		uint32_t inc = __popc(word);
#pragma unroll
		for (int s = 1; s < 32; s <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, inc, s);
			if (lane >= s) {
				inc += o;
			}
		}
		if (lane == 31) {
			SG_MI->wsum[warp] = inc;
		}
		__syncthreads();
		uint32_t before = 0, nr = 0;
		for (int w = 0; w < SG_WARPS; ++w) {
			const uint32_t v = SG_MI->wsum[w];
			before += w < warp ? v : 0u;
			nr += v;
		}
		SG_RAREPRE[tid] = (uint16_t)(before + inc - __popc(word)); /* marked elements in front of my word */
		__syncthreads();
		if (PROF && threadIdx.x == 0) {
			gp[16 + 7] += nr; /* K = 1's spare slot: marked positions */
		}
		for (uint32_t k = tid; k < nr; k += SG_THREADS) {
			/* the k-th marked element: the last word with at most k marked elements in front of it, then the bit */
			uint32_t lo = 0, hi = 1023;
			while (lo < hi) {
				const uint32_t mid = (lo + hi + 1u) >> 1;
				if ((uint32_t)SG_RAREPRE[mid] <= k) {
					lo = mid;
				} else {
					hi = mid - 1u;
				}
			}
			uint32_t wbits = SG_RAREMAP[lo];
			for (uint32_t skip = k - SG_RAREPRE[lo]; skip > 0u; --skip) {
				wbits &= wbits - 1u;
			}
			const uint32_t i = 32u * lo + (uint32_t)(__ffs((int)wbits) - 1);
			const uint32_t e = In[i];
			SG_L8[e] = (uint8_t)sg_rare(INBUF, i, e, M, (uint32_t)c.t, c.D); /* (q = e + 1 - K = e) */
		}
	}
	SG_PLAP(0)
	sg_offsets64<INBUF>();
	SG_PLAP(1)
#pragma unroll
	for (int r = 0; r < 16; ++r) {
		if ((uint32_t)r < RH) {
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const uint32_t i = blk[h] + 32u * r + lane;
				if (i < M) {
					const uint32_t rank = (rk[h][r / 3] >> (10 * (r % 3))) & 1023u;
					const uint32_t e = K == 0 ? i : (uint32_t)In[i];
					const uint32_t dd = SG_XSB[SG_XS_OFF + e + 3 - K]; /* the digit again: byte 3-K of the gram */
					Out[myh[h][dd] + rank] = (uint16_t)e;
				}
			}
		}
	}
	__syncthreads();
	SG_PLAP(2)
#undef SG_PLAP
}

/* ---- pushing a group, by size: the CTA's list (above SG_WAVE_MAX), the ring of the waves (above
 * SG_CHAIN_MAX; read behind CTA barriers: no flags, no fences), or the queue of the chains, whose
 * entries are taken by polling warps: the entry is published behind a fence, with its valid bit. */
__device__ __noinline__ void sg_push(uint32_t start, uint32_t len, uint32_t L, uint32_t buf)
{
	if (len > SG_WAVE_MAX) {
		const uint32_t slot = atomicAdd(&SG_MI->big_tail, 1u) % SG_BIGQ;
		SG_MI->big[slot] = make_uint2(start | (len << 16), L | (buf << 8));
	} else if (len > SG_CHAIN_MAX) {
		const uint32_t slot = atomicAdd(&SG_MI->r_tail, 1u) % SG_RING;
		SG_MI->ring[slot] = make_uint2(start | (len << 16), L | (buf << 8));
	} else {
		const uint32_t slot = atomicAdd(&SG_MI->q_tail, 1u) % SG_Q;
		SG_QLVL[slot] = (uint8_t)L;
		__threadfence_block();
		SG_QENT[slot] = Q_VALID | (buf << 26) | ((len - 1u) << 15) | start;
	}
}

/* how many levels from Lfrom on the n elements In[0 .. n) (all of one group) go on with one and the
 * same byte: the first level at which their bytes differ, or 32 (whole warp; 4 bytes per step) */
__device__ __forceinline__ uint32_t sg_common_levels(const uint16_t *In, uint32_t n, uint32_t Lfrom)
{
	const int lane = threadIdx.x & 31;
	const uint32_t e0 = In[0];
	uint32_t Ln = Lfrom;
	while (Ln < 32u) {
		const uint32_t g0 = sg_gram(e0 + Ln);
		uint32_t m = 4u;
		for (uint32_t i = lane; i < n; i += 32u) {
			const uint32_t x = sg_gram((uint32_t)In[i] + Ln) ^ g0;
			if (x != 0u) {
				m = min(m, (uint32_t)(__ffs((int)x) - 1) >> 3);
			}
		}
		m = __reduce_min_sync(FULL_MASK, m);
		Ln += m;
		if (m < 4u) {
			break;
		}
	}
	return min(Ln, 32u);
}

/* ---- the level-4 groups: heads of the final LSD order, every group of >= t+2 elements pushed */
__device__ __forceinline__ void sg_groups4(const SegCtx &c, uint32_t buf)
{
	const uint16_t *__restrict__ P = SG_P(buf);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t M = c.M;
	const uint32_t R = (M + 1023u) >> 10;
	const uint32_t blk = (uint32_t)warp * 32u * R;
	uint32_t *HB = reinterpret_cast<uint32_t *>(SG_WH);          /* head bits, 32 R words */
	uint32_t carry = blk > 0 && blk - 1u < M ? sg_gram(P[blk - 1u]) : 0u;
	for (uint32_t r = 0; r < R; ++r) {
		const uint32_t i = blk + 32u * r + lane;
		const bool valid = i < M;
		const uint32_t g = valid ? sg_gram(P[i]) : 0u;
		uint32_t gp = __shfl_up_sync(FULL_MASK, g, 1);
		if (lane == 0) {
			gp = carry;
		}
		const bool head = valid && (i == 0 || g != gp);
		const uint32_t hm = __ballot_sync(FULL_MASK, head);
		if (lane == 0) {
			HB[(blk >> 5) + r] = hm;
		}
		carry = __shfl_sync(FULL_MASK, g, 31);
	}
	__syncthreads();
	/* first head at or behind each word, then the first head behind it */
	const uint32_t nwords = 32u * R;
	const uint32_t word = (uint32_t)tid < nwords ? HB[tid] : 0u;
	uint32_t v = word != 0u ? 32u * tid + (uint32_t)(__ffs(word) - 1) : M;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_down_sync(FULL_MASK, v, d);
		if (lane + d < 32) {
			v = min(v, o);
		}
	}
	if (lane == 0) {
		SG_MI->wsum[warp] = v;
	}
	uint32_t nxt = __shfl_down_sync(FULL_MASK, v, 1);
	if (lane == 31) {
		nxt = M;
	}
	__syncthreads();
	/* the first head in the warps behind mine: one load and one warp-wide minimum (a loop over the warps here
	 * cost every thread up to 31 loads: 12 K of the 22 K cycles of this step) */
	nxt = min(nxt, __reduce_min_sync(FULL_MASK, lane > warp ? SG_MI->wsum[lane] : M));
	uint32_t bits = word;
	while (bits != 0u) {
		const uint32_t start = 32u * tid + (uint32_t)(__ffs(bits) - 1);
		bits &= bits - 1u;
		const uint32_t end = bits != 0u ? 32u * tid + (uint32_t)(__ffs(bits) - 1) : nxt;
		if (end - start >= (uint32_t)c.t + 2u) {
			sg_push(start, end - start, 4u, buf);
		}
	}
	__syncthreads();
}

/* ---- one level of one group, whole CTA (groups above SG_WAVE_MAX elements) --------------------
 * Three sweeps over the group's range, each warp on a block of it: (A) the test, (B) kept elements
 * and the histogram of their next bytes, (C) the placement.  The test is cheap and simply
 * re-evaluated; what crosses warps is the last passed position in front of each block. */
__device__ __forceinline__ void sg_big_level(const SegCtx &c, uint32_t s, uint32_t g, uint32_t L, uint32_t buf)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint16_t *__restrict__ In = SG_P(buf) + s;
	uint16_t *__restrict__ Out = SG_P(buf ^ 1u) + s;
	const uint32_t R = (g + 1023u) >> 10;
	const uint32_t blk = (uint32_t)warp * 32u * R;
	uint16_t *myh = SG_WH + warp * 256;
	const uint32_t lt = (1u << lane) - 1u;
	/* the test of element i (relative to s) */
	auto test = [&](uint32_t i, uint32_t &e) -> bool {
		e = i < g ? (uint32_t)In[i] & 0x7fffu : 0u;
		bool pass = false;
		if (i + c.la < g) {
			const uint32_t ef = (uint32_t)In[i + c.la] & 0x7fffu;
			pass = ef - e <= c.D && e - 3u < c.Bs;
		}
		return pass;
	};
	/* (A) */
	uint32_t lastp = SG_NONE;
	for (uint32_t r = 0; r < R; ++r) {
		uint32_t e;
		const bool pass = test(blk + 32u * r + lane, e);
		if (pass) {
			SG_L8[e - 3u] = (uint8_t)L;
		}
		const uint32_t pm = __ballot_sync(FULL_MASK, pass);
		if (pm != 0u) {
			lastp = __shfl_sync(FULL_MASK, e, 31 - __clz((int)pm));
		}
	}
	if (lane == 0) {
		SG_MI->lastp[warp] = lastp;
	}
	reinterpret_cast<uint4 *>(myh)[lane] = make_uint4(0u, 0u, 0u, 0u);
	const int any = __syncthreads_or(lastp != SG_NONE);
	if (!any || L >= 32u) {
		return; /* nobody passed: the group is finished (uniform over the CTA) */
	}
	/* the last passed position in front of my block */
	uint32_t carry0;
	{
		const uint32_t lp = lane < warp ? SG_MI->lastp[lane] : SG_NONE;
		const uint32_t have = __ballot_sync(FULL_MASK, lp != SG_NONE);
		carry0 = SG_NONE;
		if (have != 0u) {
			carry0 = __shfl_sync(FULL_MASK, lp, 31 - __clz((int)have));
		}
	}
	/* kept and next byte of the element of round r: shared by (B) and (C) */
	auto kept_byte = [&](uint32_t r, uint32_t &carry, uint32_t &e, uint32_t &b) -> bool {
		const uint32_t i = blk + 32u * r + lane;
		const bool pass = test(i, e);
		const uint32_t pm = __ballot_sync(FULL_MASK, pass);
		const uint32_t upto = pm & ((2u << lane) - 1u);
		const uint32_t src = upto != 0u ? 31 - __clz((int)upto) : 0;
		const uint32_t lpv = __shfl_sync(FULL_MASK, e, src);
		const uint32_t lp = upto != 0u ? lpv : carry;
		const bool kept = i < g && lp != SG_NONE && e - lp <= c.D;
		if (pm != 0u) {
			carry = __shfl_sync(FULL_MASK, e, 31 - __clz((int)pm));
		}
		b = kept ? sg_byte(e, L) : 256u; /* (one value for all the others: MATCH.ANY takes 2 cycles per distinct value) */
		return kept;
	};
	/* (B) */
	{
		uint32_t carry = carry0;
		for (uint32_t r = 0; r < R; ++r) {
			uint32_t e, b;
			const bool kept = kept_byte(r, carry, e, b);
			const uint32_t peers = __match_any_sync(FULL_MASK, b);
			if (kept && lane == __ffs(peers) - 1) {
				myh[b] = (uint16_t)(myh[b] + __popc(peers));
			}
			__syncwarp();
		}
	}
	__syncthreads();
	const uint32_t total = sg_offsets(c, (uint32_t)c.t + 2u);
	/* all elements kept and followed by one and the same byte: the group stays where it is */
	const int whole = __syncthreads_or(tid < 256 && total == g);
	if (whole) {
		if (tid == 0) {
			sg_push(s, g, L + 1u, buf);
		}
		return;
	}
	/* (C) */
	{
		uint32_t carry = carry0;
		for (uint32_t r = 0; r < R; ++r) {
			uint32_t e, b;
			const bool kept = kept_byte(r, carry, e, b);
			const uint32_t base = kept ? SG_MI->dbase[b] : SG_NONE;
			const bool act = kept && base != SG_NONE;
			const uint32_t peers = __match_any_sync(FULL_MASK, act ? b : 256u);
			if (act) {
				const uint32_t o = myh[b];
				Out[base + o + __popc(peers & lt)] = (uint16_t)e;
			}
			__syncwarp();
			if (act && lane == __ffs(peers) - 1) {
				myh[b] = (uint16_t)(myh[b] + __popc(peers));
			}
			__syncwarp();
		}
	}
	__syncthreads();
	if (tid < 256 && SG_MI->dbase[tid] != SG_NONE) {
		sg_push(s + SG_MI->dbase[tid], total, L + 1u, buf ^ 1u);
	}
}

/* ---- one wave: the next groups of the ring, up to 32 rows, one warp per row.  A group takes one row
 * per SG_CHUNK elements plus one row for its totals.  A group of the ring (any level) is tested,
 * pruned to its kept elements and split by the next byte; its children are pushed by size.
 *   form   warp 0 takes groups while their rows fit
 *   (1)    every row: test, kept elements, next bytes, histogram row (each lane keeps its 8 elements with
 *          byte and rank in registers).  What crosses rows -- the last passed element in front of a row --
 *          is looked for in the 64 elements in front of the row; when they hold none (and are not out of
 *          reach already) the element in front of the row stands in for it: the kept set only grows,
 *          which never changes a result (every kept element is a real occurrence of the group's gram)
 *   (2)    the group's threads, one byte value each: running sums down the group's rows (a row's count ->
 *          the elements of that byte in the rows in front), the totals of the children that live on
 *          (>= t+2 elements) into the group's totals row
 *   (3)    every row: the children's slots from the totals row, its own final offsets, the placement into
 *          the group's own range of the other buffer; the group's first row pushes the children
 * Returns false when the ring is empty. */
constexpr uint32_t SG_WHOLE = 2u;

template <bool PROF>
__device__ __forceinline__ bool sg_wave(const SegCtx &c, unsigned long long *gp)
{
	/* measurement build: cycles of the wave's parts and the wave count, behind the phase counters */
	unsigned long long *wp = reinterpret_cast<unsigned long long *>(sg_smem + SG_SMEM) + 10;
	unsigned long long wt = 0;
#define SG_WLAP(k)                                 \
	if (PROF && threadIdx.x == 0) {                \
		const unsigned long long now = clock64();  \
		wp[k] += now - wt;                         \
		wt = now;                                  \
	}
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t need = (uint32_t)c.t + 2u;
	__syncthreads(); /* the pushes and placements of the wave before */
	const uint32_t head = SG_MI->r_head, tail = SG_MI->r_tail;
	if (head == tail) {
		return false;
	}
	if (PROF && threadIdx.x == 0) {
		wt = clock64();
		wp[5] += 1;
	}
	if (warp == 0) {
		const bool have = (uint32_t)lane < tail - head && lane < 16;
		const uint2 ent = have ? SG_MI->ring[(head + lane) % SG_RING] : make_uint2(0u, 0u);
		const uint32_t len = ent.x >> 16;
		const uint32_t nch = have ? (len + SG_CHUNK - 1u) / SG_CHUNK + 1u : 0u; /* + the totals row */
		uint32_t inc = nch;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
			if (lane >= d) {
				inc += o;
			}
		}
		const bool fits = have && inc <= 32u;
		const uint32_t fm = __ballot_sync(FULL_MASK, fits); /* a prefix of the lanes: the row count only grows */
		if (fits) {
			SegGroup &g = SG_MI->grp[lane];
			g.start = ent.x & 0xffffu;
			g.len = len;
			g.L = ent.y & 255u;
			g.buf = (ent.y >> 8) & 1u;
			g.r0 = inc - nch;
			g.nrows = nch - 1u;
			g.state = 0u;
		}
		const int ng = __popc(fm);
		const uint32_t rows = __shfl_sync(FULL_MASK, inc, ng - 1);
		if (lane == 0) {
			SG_MI->wave_groups = (uint32_t)ng;
			SG_MI->wave_rows = rows;
			SG_MI->r_head = head + (uint32_t)ng;
		}
	}
	__syncthreads();
	SG_WLAP(0)
	const uint32_t ng = SG_MI->wave_groups, nrows = SG_MI->wave_rows;
	/* my row's group */
	uint32_t gi = 0, gstart = 0, glen = 0, L = 0, buf = 0, r0 = 0, gn = 0;
	if ((uint32_t)warp < nrows) {
		const uint32_t gr0 = (uint32_t)lane < ng ? SG_MI->grp[lane].r0 : 0xffffu;
		const uint32_t m = __ballot_sync(FULL_MASK, gr0 <= (uint32_t)warp);
		gi = 31 - __clz((int)m);
		const SegGroup &g = SG_MI->grp[gi];
		gstart = g.start;
		glen = g.len;
		L = g.L;
		buf = g.buf;
		r0 = g.r0;
		gn = g.nrows;
	}
	const uint32_t myrow = (uint32_t)warp - r0;
	const bool rowon = (uint32_t)warp < nrows && myrow < gn; /* (the totals row's warp has no elements) */
	const uint32_t coff = myrow * SG_CHUNK;
	const uint16_t *In = SG_P(buf) + gstart;
	uint16_t *myh = SG_WH + warp * 256;
	uint32_t st[8]; /* my element of round r: id | rank << 15 | byte << 23 | kept << 31 */
	/* (1) */
	unsigned long long st0 = 0;
#define SG_SLAP(k)                                 \
	if (PROF && threadIdx.x == 0) {                \
		const unsigned long long now = clock64();  \
		gp[40 + (k)] += now - st0;                 \
		st0 = now;                                 \
	}
	if (PROF && threadIdx.x == 0) {
		st0 = clock64();
	}
	if (rowon) {
		reinterpret_cast<uint4 *>(myh)[lane] = make_uint4(0u, 0u, 0u, 0u);
		uint32_t ef[8];
#pragma unroll
		for (int r = 0; r < 8; ++r) { /* all the loads first: nothing between them that they could depend on */
			const uint32_t i = coff + 32u * r + lane;
			st[r] = i < glen ? (uint32_t)In[i] : 0u;
			ef[r] = i + c.la < glen ? (uint32_t)In[i + c.la] : 0xffffffu;
		}
		SG_SLAP(0)
		uint32_t pmk[8];
#pragma unroll
		for (int r = 0; r < 8; ++r) {
			const bool pass = ef[r] - st[r] <= c.D && st[r] - 3u < c.Bs; /* (no element t+1 places on: a huge distance) */
			if (pass) {
				SG_L8[st[r] - 3u] = (uint8_t)L;
			}
			pmk[r] = __ballot_sync(FULL_MASK, pass);
		}
		SG_SLAP(1)
		if (L < 32u) {
			/* the last passed element in front of my row */
			uint32_t carry = SG_NONE;
			if (myrow > 0u) {
				const uint32_t first = __shfl_sync(FULL_MASK, st[0], 0);
				bool settled = false;
#pragma unroll
				for (int back = 1; back <= 2; ++back) {
					if (!settled) {
						const uint32_t j = coff - 32u * back + lane;
						const uint32_t ej = In[j];
						const uint32_t efj = j + c.la < glen ? (uint32_t)In[j + c.la] : 0xffffffu;
						const uint32_t pmj = __ballot_sync(FULL_MASK, efj - ej <= c.D && ej - 3u < c.Bs);
						const uint32_t lastj = __shfl_sync(FULL_MASK, ej, pmj != 0u ? 31 - __clz((int)pmj) : 0);
						if (pmj != 0u) {
							carry = lastj;
							settled = true;
						} else if (first - lastj > c.D) { /* (lastj: lane 0's, the smallest position of the 32) */
							settled = true; /* nothing further back can reach into my row */
						}
					}
				}
				if (!settled) {
					carry = In[coff - 1u];
				}
			}
			SG_SLAP(2)
			uint32_t bb[8];
#pragma unroll
			for (int r = 0; r < 8; ++r) {
				const uint32_t i = coff + 32u * r + lane;
				const uint32_t e = st[r];
				const uint32_t pm = pmk[r];
				const uint32_t upto = pm & ((2u << lane) - 1u);
				const uint32_t lpv = __shfl_sync(FULL_MASK, e, upto != 0u ? 31 - __clz((int)upto) : 0);
				const uint32_t lpos = upto != 0u ? lpv : carry;
				const bool kept = i < glen && lpos != SG_NONE && e - lpos <= c.D;
				if (pm != 0u) {
					carry = __shfl_sync(FULL_MASK, e, 31 - __clz((int)pm));
				}
				bb[r] = kept ? sg_byte(e, L) : 256u; /* (one value for all the others: a distinct value costs MATCH.ANY 2 cycles) */
			}
			SG_SLAP(3)
#pragma unroll
			for (int r = 0; r < 8; ++r) {
				/* the leader of a byte value adds its lanes to the row's count with ONE shared-memory atomic and hands
				 * the old count on: nothing to wait for between the rounds (the atomics of a warp on one address
				 * keep their order), so the eight matches, adds and shuffles of a row overlap */
				const uint32_t b = bb[r];
				const bool kept = b < 256u;
				const uint32_t peers = __match_any_sync(FULL_MASK, b);
				uint32_t old = 0;
				if (kept && (peers & lt) == 0u) {
					/* (two 16-bit counts to a word: the add goes to the byte value's half) */
					const uint32_t sh = (b & 1u) * 16u;
					old = (atomicAdd(reinterpret_cast<uint32_t *>(myh) + (b >> 1), (uint32_t)__popc(peers) << sh) >> sh) & 0xffffu;
				}
				old = __shfl_sync(FULL_MASK, old, __ffs((int)peers) - 1);
				st[r] = kept ? st[r] | ((old + __popc(peers & lt)) << 15) | (b << 23) | 0x80000000u : 0u;
			}
			SG_SLAP(4)
		}
	}
	__syncthreads();
	SG_SLAP(5)
#undef SG_SLAP
	SG_WLAP(1)
	/* (2) */
	if (rowon && L < 32u) {
		const uint32_t nthr = gn * 32u;
		uint16_t *col = SG_WH + r0 * 256;
		for (uint32_t dig = myrow * 32u + lane; dig < 256u; dig += nthr) {
			uint32_t run = 0;
#pragma unroll 4
			for (uint32_t rr = 0; rr < gn; ++rr) {
				const uint32_t v = col[rr * 256u + dig];
				col[rr * 256u + dig] = (uint16_t)run;
				run += v;
			}
			col[gn * 256u + dig] = (uint16_t)(run >= need ? run : 0u);
			if (run == glen) {
				SG_MI->grp[gi].state = SG_WHOLE; /* (one byte value at most) */
			}
		}
	}
	__syncthreads();
	SG_WLAP(2)
	/* (3) */
	if (rowon && L < 32u) {
		if (SG_MI->grp[gi].state == SG_WHOLE) {
			/* every element kept and followed by one and the same byte: the group stays where it is, and
			 * nothing changes (same array, same test) down to the level where the bytes differ */
			if (myrow == 0u) {
				const uint32_t Ln = sg_common_levels(In, glen, L + 1u);
				if (lane == 0) {
					sg_push(gstart, glen, Ln, buf);
				}
			}
		} else {
			const uint4 tv = reinterpret_cast<const uint4 *>(SG_WH + (r0 + gn) * 256)[lane];
			const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
			uint32_t tot[8], mine = 0;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				tot[k] = (tw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
				mine += tot[k];
			}
			uint32_t inc = mine;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
				if (lane >= d) {
					inc += o;
				}
			}
			if (__ballot_sync(FULL_MASK, mine != 0u) != 0u) {
				const uint4 hv = reinterpret_cast<const uint4 *>(myh)[lane];
				const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
				uint32_t ow[4] = {0u, 0u, 0u, 0u};
				uint32_t base = inc - mine;
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					const uint32_t pre = (hw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
					ow[k >> 1] |= (tot[k] != 0u ? base + pre : 0xffffu) << (16 * (k & 1));
					if (tot[k] != 0u) {
						if (myrow == 0u) {
							sg_push(gstart + base, tot[k], L + 1u, buf ^ 1u);
						}
						base += tot[k];
					}
				}
				reinterpret_cast<uint4 *>(myh)[lane] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
				__syncwarp();
				uint16_t *Out = SG_P(buf ^ 1u) + gstart;
#pragma unroll
				for (int r = 0; r < 8; ++r) {
					const uint32_t v = st[r];
					if (v & 0x80000000u) {
						const uint32_t off = myh[(v >> 23) & 255u];
						if (off != 0xffffu) {
							Out[off + ((v >> 15) & 255u)] = (uint16_t)(v & 0x7fffu);
						}
					}
				}
			}
		}
	}
	SG_WLAP(3)
#undef SG_WLAP
	return true;
}

/* ---- a group of at most SG_CHAIN_MAX elements: one warp follows it down to the level where it
 * ends, 8 elements per lane in registers, no CTA barrier.  Per level: the test, the kept elements,
 * their next bytes and ranks (histogram row of the warp), the children's slots, the placement in the
 * group's own range of the other buffer.  The warp goes on with one child and queues the others. */
__device__ __forceinline__ void sg_chain(const SegCtx &c, uint32_t start, uint32_t len, uint32_t L, uint32_t buf)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint16_t *myh = SG_WH + warp * 256;
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t need = (uint32_t)c.t + 2u;
	for (;;) {
		const uint16_t *In = SG_P(buf) + start;
		const uint32_t R = (len + 31u) >> 5;
		reinterpret_cast<uint4 *>(myh)[lane] = make_uint4(0u, 0u, 0u, 0u);
		__syncwarp();
		uint32_t st[8]; /* my element of round r: id | rank << 15 | byte << 23 | kept << 31 */
		uint32_t carry = SG_NONE, anypass = 0, nkept = 0;
#pragma unroll
		for (int r = 0; r < 8; ++r) {
			st[r] = 0u;
			if ((uint32_t)r < R) {
				const uint32_t i = 32u * r + lane;
				const uint32_t e = i < len ? (uint32_t)In[i] : 0u;
				bool pass = false;
				if (i + c.la < len) {
					const uint32_t ef = In[i + c.la];
					pass = ef - e <= c.D && e - 3u < c.Bs;
				}
				if (pass) {
					SG_L8[e - 3u] = (uint8_t)L;
				}
				const uint32_t pm = __ballot_sync(FULL_MASK, pass);
				anypass |= pm;
				if (L < 32u) {
					const uint32_t upto = pm & ((2u << lane) - 1u);
					const uint32_t lpv = __shfl_sync(FULL_MASK, e, upto != 0u ? 31 - __clz((int)upto) : 0);
					const uint32_t lpos = upto != 0u ? lpv : carry;
					const bool kept = i < len && lpos != SG_NONE && e - lpos <= c.D;
					if (pm != 0u) {
						carry = __shfl_sync(FULL_MASK, e, 31 - __clz((int)pm));
					}
					const uint32_t b = kept ? sg_byte(e, L) : 256u; /* (one value for all the others: a distinct value costs MATCH.ANY 2 cycles) */
					const uint32_t peers = __match_any_sync(FULL_MASK, b);
					const int leader = __ffs(peers) - 1;
					uint32_t old = 0;
					if (kept && lane == leader) {
						old = myh[b];
						myh[b] = (uint16_t)(old + __popc(peers));
					}
					old = __shfl_sync(FULL_MASK, old, leader);
					st[r] = kept ? e | ((old + __popc(peers & lt)) << 15) | (b << 23) | 0x80000000u : e;
					nkept += __popc(__ballot_sync(FULL_MASK, kept));
					__syncwarp();
				}
			}
		}
		if (anypass == 0u || L >= 32u || nkept < need) {
			return; /* nobody passed / the last level / too few left to pass again */
		}
		uint32_t cnt[8], off[8];
		uint32_t mine = 0, whole = 0;
		{
			const uint4 hv = reinterpret_cast<const uint4 *>(myh)[lane];
			const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				cnt[k] = (hw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
				mine += cnt[k] >= need ? cnt[k] : 0u;
				whole |= cnt[k] == len ? 1u : 0u;
			}
		}
		if (__any_sync(FULL_MASK, whole != 0u)) {
			/* every element kept and followed by one and the same byte: the group stays where it is, and
			 * nothing changes (same array, same test) down to the level where the bytes differ */
			uint32_t Ln = L + 1u;
			const uint32_t e0 = In[0];
			while (Ln < 32u) {
				const uint32_t g0 = sg_gram(e0 + Ln);
				uint32_t m = 4u;
#pragma unroll
				for (int r = 0; r < 8; ++r) {
					if (32u * r + lane < len) {
						const uint32_t x = sg_gram((st[r] & 0x7fffu) + Ln) ^ g0;
						if (x != 0u) {
							m = min(m, (uint32_t)(__ffs((int)x) - 1) >> 3);
						}
					}
				}
				m = __reduce_min_sync(FULL_MASK, m);
				Ln += m;
				if (m < 4u) {
					break;
				}
			}
			L = min(Ln, 32u);
			continue;
		}
		/* slots of the children: bins of >= t+2 elements, in byte order */
		uint32_t inc = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
			if (lane >= d) {
				inc += o;
			}
		}
		uint32_t run = inc - mine, nchild = 0;
		{
			uint32_t hw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const bool al = cnt[k] >= need;
				off[k] = al ? run : 0xffffu;
				hw[k >> 1] |= off[k] << (16 * (k & 1));
				run += al ? cnt[k] : 0u;
				nchild += al ? 1u : 0u;
			}
			reinterpret_cast<uint4 *>(myh)[lane] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
		}
		const uint32_t cm = __ballot_sync(FULL_MASK, nchild != 0u);
		if (cm == 0u) {
			return; /* every child too small to pass again */
		}
		__syncwarp();
		/* placement, from the registers */
		{
			uint16_t *Out = SG_P(buf ^ 1u) + start;
#pragma unroll
			for (int r = 0; r < 8; ++r) {
				const uint32_t v = st[r];
				if ((uint32_t)r < R && (v & 0x80000000u) != 0u) {
					const uint32_t o = myh[(v >> 23) & 255u];
					if (o != 0xffffu) {
						Out[o + ((v >> 15) & 255u)] = (uint16_t)(v & 0x7fffu);
					}
				}
			}
		}
		/* the first child is mine, the others are queued (their elements are written: publish behind a fence) */
		const int fl = __ffs(cm) - 1;
		uint32_t mystart = 0, mylen = 0;
		bool took = false;
		__threadfence_block();
		__syncwarp();
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			if (off[k] != 0xffffu) {
				if (lane == fl && !took) {
					mystart = off[k];
					mylen = cnt[k];
					took = true;
				} else {
					sg_push(start + off[k], cnt[k], L + 1u, buf ^ 1u);
				}
			}
		}
		start += __shfl_sync(FULL_MASK, mystart, fl);
		len = __shfl_sync(FULL_MASK, mylen, fl);
		buf ^= 1u;
		L += 1u;
	}
}

template <bool PROF>
__global__ void __launch_bounds__(SG_THREADS, 1) x3_seg_kernel(SegArgs a)
{
	SegCtx c;
	c.D = a.D;
	c.t = a.t;
	c.la = (uint32_t)a.t + 1u;
	const int tid = threadIdx.x, lane = tid & 31;
	uint64_t *bar = reinterpret_cast<uint64_t *>(&SG_MI->mbar);

	if (tid == 0) {
		mbar_init(bar, 1);
	}
	for (uint32_t i = tid; i < SG_Q; i += SG_THREADS) {
		SG_QENT[i] = 0u; /* (a taken entry is cleared by the warp that took it) */
	}
	__syncthreads();
	uint32_t phase = 0;
	/* (the cycle counters of the measurement build live in shared memory behind the CTA's own data: no registers) */
	unsigned long long *pt = reinterpret_cast<unsigned long long *>(sg_smem + SG_SMEM);
	if (PROF && tid == 0) {
		for (int k = 0; k < 16; ++k) {
			pt[k] = 0;
		}
		pt[9] = clock64();
	}
#define SG_LAP(k)                                  \
	if (PROF && tid == 0) {                        \
		const unsigned long long now = clock64();  \
		pt[k] += now - pt[9];                      \
		pt[9] = now;                               \
	}
	for (;;) {
		if (tid == 0) {
			SG_MI->seg = atomicAdd(a.ticket, 1u);
			SG_MI->q_head = SG_MI->q_tail = 0u;
			SG_MI->q_done = 0u;
			SG_MI->r_head = SG_MI->r_tail = 0u;
			SG_MI->big_head = SG_MI->big_tail = 0u;
		}
		__syncthreads();
		const uint32_t seg = SG_MI->seg;
		if (seg >= a.nseg) {
			break;
		}
		/* the launch's last, partly filled wave of segments (and every segment of a launch that has fewer
		 * segments than SMs) is dealt out in split_k-ths: more SMs finish it in less time */
		uint32_t lseg = seg, sub = 0, nsub = 1;
		if (seg >= a.split_first) {
			lseg = a.split_first + (seg - a.split_first) / a.split_k;
			sub = (seg - a.split_first) % a.split_k;
			nsub = a.split_k;
		}
		unsigned long long gseg = lseg;
		if (a.spp != 0u) {
			gseg = (unsigned long long)(a.part + (a.k0 + lseg / a.spp) * a.parts) * a.spp + lseg % a.spp;
		}
		const unsigned long long a0 = gseg * a.B + (unsigned long long)sub * a.split_B;
		if (a0 >= a.n) {
			__syncthreads(); /* (everybody has read the ticket before thread 0 draws the next one) */
			continue; /* (the input's last piece may be short) */
		}
		{
			unsigned long long a1 = gseg * a.B + (sub + 1u == nsub ? a.B : (unsigned long long)(sub + 1u) * a.split_B);
			if (a1 > a.n) {
				a1 = a.n;
			}
			c.Bs = (uint32_t)(a1 - a0);
		}
		c.M = c.Bs + a.D + 3u;
		/* the segment's bytes: xs[k] = x[a0 - 3 + k], k in [0, M + 48); virtual zeros in front of the input */
		{
			const uint32_t want = ((a0 == 0 ? 0u : 16u) + c.M + 48u + 15u) & ~15u; /* bytes from x[a0 - 16] (or x[0]) */
			const unsigned long long src = a0 == 0 ? 0ull : a0 - 16ull;
			const unsigned long long avail = a.xbytes > src ? a.xbytes - src : 0ull;
			const uint32_t bytes = (uint32_t)(avail < want ? avail : want) & ~15u;
			uint8_t *dst = SG_XSB + (a0 == 0 ? 16 : 0);
			if (tid == 0) {
				/* the previous segment read this memory through the generic proxy */
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				mbar_expect_tx(bar, bytes);
				for (uint32_t o = 0; o < bytes; o += 16384u) {
					tma_load_1d(dst + o, a.x + src + o, bytes - o < 16384u ? bytes - o : 16384u, bar);
				}
			}
			/* what the copy does not cover: the virtual bytes in front, the tail behind the readable range */
			if (a0 == 0 && tid < 16) {
				SG_XSB[tid] = 0;
			}
			for (uint32_t o = (a0 == 0 ? 16u : 0u) + bytes + tid; o < SG_XS_BYTES; o += SG_THREADS) {
				SG_XSB[o] = 0;
			}
			mbar_wait(bar, phase);
			phase ^= 1u;
			__syncthreads();
		}
		SG_LAP(0)
		/* levels 1 .. 4 */
#ifdef X3_SEG_LSD1
		sg_lsd_pass<0, true, 0>(c);  /* (reads the bytes in position order) -> buffer 1 */
		SG_LAP(1)
		if (c.la <= 32u) {
			sg_lsd_pass<1, true, 1>(c);
			sg_lsd_pass<2, true, 0>(c);
			sg_lsd_pass<3, true, 1>(c);
		} else {
			sg_lsd_pass<1, false, 1>(c);
			sg_lsd_pass<2, false, 0>(c);
			sg_lsd_pass<3, false, 1>(c);
		}
		SG_LAP(2)
#else
		sg_lsd_pass2<0, true, 0, PROF>(c, a.prof + blockIdx.x * 64);  /* (reads the bytes in position order) -> buffer 1 */
		SG_LAP(1)
		if (c.la <= 32u) {
			sg_lsd_pass2<1, true, 1, PROF>(c, a.prof + blockIdx.x * 64);
			sg_lsd_pass2<2, true, 0, PROF>(c, a.prof + blockIdx.x * 64);
			sg_lsd_pass2<3, true, 1, PROF>(c, a.prof + blockIdx.x * 64);
		} else {
			sg_lsd_pass2<1, false, 1, PROF>(c, a.prof + blockIdx.x * 64);
			sg_lsd_pass2<2, false, 0, PROF>(c, a.prof + blockIdx.x * 64);
			sg_lsd_pass2<3, false, 1, PROF>(c, a.prof + blockIdx.x * 64);
		}
		/* the histogram rows of the second sub-blocks lay over the chains' queue: empty again */
		for (uint32_t i = tid; i < SG_Q; i += SG_THREADS) {
			SG_QENT[i] = 0u;
		}
		SG_LAP(2)
#endif
		sg_groups4(c, 0u);
		SG_LAP(3)
		/* big groups, whole CTA, one level at a time */
		for (;;) {
			__syncthreads();
			const uint32_t h = SG_MI->big_head, tl = SG_MI->big_tail;
			__syncthreads();
			if (h == tl) {
				break;
			}
			const uint2 be = SG_MI->big[h % SG_BIGQ];
			if (tid == 0) {
				SG_MI->big_head = h + 1u;
			}
			sg_big_level(c, be.x & 0xffffu, be.x >> 16, be.y & 255u, (be.y >> 8) & 1u);
		}
		SG_LAP(4)
		/* groups above SG_CHAIN_MAX elements, wave after wave */
		while (sg_wave<PROF>(c, a.prof + blockIdx.x * 64)) {
		}
		__syncthreads();
		SG_LAP(5)
		/* every other group: the warps drain the queue, each following its group down */
		for (;;) {
			uint32_t ent = 0, L = 0;
			if (lane == 0) {
				const uint32_t slot = atomicAdd(&SG_MI->q_head, 1u) % SG_Q;
				for (;;) {
					ent = SG_QENT[slot];
					if (ent & Q_VALID) {
						__threadfence_block();
						L = SG_QLVL[slot];
						SG_QENT[slot] = 0u;
						break;
					}
					/* every pushed group finished (finished count first: it never overtakes the pushes) */
					const uint32_t done = *(volatile uint32_t *)&SG_MI->q_done;
					if (done == *(volatile uint32_t *)&SG_MI->q_tail) {
						ent = 0;
						break;
					}
					__nanosleep(100); /* nothing queued yet: leave the issue slots to the warps that work */
				}
			}
			ent = __shfl_sync(FULL_MASK, ent, 0);
			L = __shfl_sync(FULL_MASK, L, 0);
			if (ent == 0u) {
				break;
			}
			__threadfence_block();
			sg_chain(c, ent & 0x7fffu, ((ent >> 15) & 0xffu) + 1u, L, (ent >> 26) & 1u);
			__syncwarp();
			if (lane == 0) {
				__threadfence_block();
				atomicAdd(&SG_MI->q_done, 1u);
			}
		}
		__syncthreads();
		SG_LAP(6)
		/* Lstar of the segment */
		{
			uint8_t *dst = a.lstar + a0;
			if (((uintptr_t)dst & 15u) == 0u) {
				const uint32_t nv = c.Bs >> 4;
				for (uint32_t v = tid; v < nv; v += SG_THREADS) {
					reinterpret_cast<uint4 *>(dst)[v] = reinterpret_cast<const uint4 *>(SG_L8)[v];
				}
				for (uint32_t o = (nv << 4) + tid; o < c.Bs; o += SG_THREADS) {
					dst[o] = SG_L8[o];
				}
			} else {
				for (uint32_t o = tid; o < c.Bs; o += SG_THREADS) {
					dst[o] = SG_L8[o];
				}
			}
		}
		__syncthreads();
		SG_LAP(7)
		if (PROF && tid == 0) {
			pt[8] += 1;
		}
	}
	if (PROF && tid == 0) {
		for (int k = 0; k < 16; ++k) {
			a.prof[blockIdx.x * 64 + k] = pt[k];
		}
	}
#undef SG_LAP
}

} /* namespace */

/* largest number of distances and smallest t the segment kernel takes */
uint32_t x3k_seg_max_distances(void)
{
	return 16351u; /* W <= 16 KB: B >= 16 400 positions per segment */
}

int x3k_seg_min_t(void)
{
	/* live groups of >= t+2 elements must fit the queue: (M - 1) / (t + 2) + 1 <= SG_Q */
	int t = 1;
	while ((SG_MMAX - 1u) / (uint32_t)(t + 2) + 1u > SG_Q) {
		++t;
	}
	return t;
}

/* positions per segment for D distances */
uint32_t x3k_seg_positions(uint32_t D)
{
	uint32_t B = (SG_MMAX - D - 3u) & ~15u;
	return B < SG_BMAX ? B : SG_BMAX;
}

cudaError_t x3k_launch_seg(const X3SearchParams &prm, cudaStream_t stream, int *launches)
{
	cudaError_t e;
	if (prm.H != nullptr || prm.D == 0 || prm.D > x3k_seg_max_distances() || prm.t < x3k_seg_min_t() ||
	    prm.tile_counter == nullptr) {
		return cudaErrorNotSupported;
	}
	if (prm.n == 0) {
		return cudaSuccess;
	}
	int dev = 0, sms = 0;
	if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
	if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
	static bool attr_set[64] = {false};
	if (dev >= 0 && dev < 64 && !attr_set[dev]) {
		if ((e = cudaFuncSetAttribute(x3_seg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM)) != cudaSuccess) return e;
		if ((e = cudaFuncSetAttribute(x3_seg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM + 128)) != cudaSuccess) return e;
		attr_set[dev] = true;
	}
	SegArgs a;
	a.x = prm.x;
	a.lstar = prm.lstar;
	a.n = prm.n;
	a.xbytes = (unsigned long long)x3k_required_bytes((size_t)prm.n, (size_t)prm.D + 33u) & ~15ull;
	a.D = prm.D;
	a.B = x3k_seg_positions(prm.D);
	a.t = prm.t;
	unsigned long long nseg = (prm.n + a.B - 1) / a.B;
	a.part = 0;
	a.parts = 1;
	a.spp = 0;
	a.k0 = 0;
	if (prm.piece_segments != 0u) {
		if (prm.parts == 0u || prm.part >= prm.parts) {
			return cudaErrorInvalidValue;
		}
		const unsigned long long npieces = (nseg + prm.piece_segments - 1) / prm.piece_segments;
		unsigned long long mine = npieces > prm.part ? (npieces - prm.part + prm.parts - 1) / prm.parts : 0;
		a.k0 = 0;
		if (prm.piece_count != 0u) {
			a.k0 = prm.piece_first;
			mine = mine > prm.piece_first ? mine - prm.piece_first : 0;
			if (mine > prm.piece_count) {
				mine = prm.piece_count;
			}
		}
		nseg = mine * prm.piece_segments;
		a.part = prm.part;
		a.parts = prm.parts;
		a.spp = prm.piece_segments;
		if (nseg == 0) {
			return cudaSuccess;
		}
	}
	if (nseg > 0xffffffffull) {
		return cudaErrorNotSupported;
	}
	unsigned grid;
	{
		const char *ns = getenv("X3_SEG_NO_SPLIT"); /* tuning/testing knob; never changes results */
		const uint32_t S = (uint32_t)sms, ns32 = (uint32_t)nseg;
		uint32_t k = 1, first = ns32, tickets = ns32;
		if (ns == nullptr && a.B >= 256u) {
			if (ns32 >= S) {
				const uint32_t full = ns32 / S * S, rem = ns32 - full;
				if (rem > 0 && rem * 2u <= S) {
					k = S / rem < 4u ? S / rem : 4u;
					first = full;
					tickets = full + rem * k;
				}
			} else {
				k = S / ns32 < 4u ? S / ns32 : 4u;
				if (k > 1u) {
					first = 0;
					tickets = ns32 * k;
				}
			}
		}
		a.split_first = first;
		a.split_k = k;
		a.split_B = (a.B / k) & ~15u;
		a.nseg = tickets;
		grid = tickets < S ? tickets : S;
	}
	a.ticket = prm.tile_counter;
	a.prof = nullptr;
	if ((e = cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
	const bool prof = getenv("X3_SEG_PROF") != nullptr; /* measurement knob: cycles per phase, printed */
	if (prof) {
		if ((e = cudaMalloc((void **)&a.prof, (size_t)grid * 512)) != cudaSuccess) return e;
		if ((e = cudaMemsetAsync(a.prof, 0, (size_t)grid * 512, stream)) != cudaSuccess) return e;
	}
	if (prof) {
		x3_seg_kernel<true><<<grid, SG_THREADS, SG_SMEM + 128, stream>>>(a);
	} else {
		x3_seg_kernel<false><<<grid, SG_THREADS, SG_SMEM, stream>>>(a);
	}
	if (launches != nullptr) {
		*launches += 1;
	}
	if (prof) {
		unsigned long long *h = (unsigned long long *)malloc((size_t)grid * 512);
		if (h != nullptr && cudaStreamSynchronize(stream) == cudaSuccess &&
		    cudaMemcpy(h, a.prof, (size_t)grid * 512, cudaMemcpyDeviceToHost) == cudaSuccess) {
			static const char *name[8] = {"load", "pass 0", "passes 1-3", "groups4", "big groups", "waves", "chains", "store"};
			double tot[32], mx = 0;
			for (int k = 0; k < 32; ++k) {
				tot[k] = 0;
			}
			for (unsigned b = 0; b < grid; ++b) {
				double cta = 0;
				for (int k = 0; k < 32; ++k) {
					tot[k] += (double)h[b * 64 + k];
					cta += k < 8 ? (double)h[b * 64 + k] : 0;
				}
				if (cta > mx) mx = cta;
			}
			fprintf(stderr, "x3_seg_kernel: %u CTAs, %.0f segments, busiest CTA %.0f cycles; cycles per segment:", grid, tot[8], mx);
			for (int k = 0; k < 8; ++k) {
				fprintf(stderr, "  %s %.0f", name[k], tot[k] / (tot[8] > 0 ? tot[8] : 1));
			}
			fprintf(stderr, ";  %.1f waves per segment, cycles per wave: form %.0f  (1) %.0f  (2) %.0f  (3) %.0f\n",
			        tot[15] / (tot[8] > 0 ? tot[8] : 1), tot[10] / (tot[15] > 0 ? tot[15] : 1), tot[11] / (tot[15] > 0 ? tot[15] : 1),
			        tot[12] / (tot[15] > 0 ? tot[15] : 1), tot[13] / (tot[15] > 0 ? tot[15] : 1));
			fprintf(stderr, "x3_seg_kernel: wave part (1) as warp 0 (row 0) sees it, cycles per wave: loads %.0f  test+votes %.0f  look-back %.0f  kept+bytes %.0f  match+histogram %.0f  barrier %.0f\n",
			        (double)0 + [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 40]; return v;}() / (tot[15] > 0 ? tot[15] : 1),
			        [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 41]; return v;}() / (tot[15] > 0 ? tot[15] : 1),
			        [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 42]; return v;}() / (tot[15] > 0 ? tot[15] : 1),
			        [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 43]; return v;}() / (tot[15] > 0 ? tot[15] : 1),
			        [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 44]; return v;}() / (tot[15] > 0 ? tot[15] : 1),
			        [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 45]; return v;}() / (tot[15] > 0 ? tot[15] : 1));
			fprintf(stderr, "x3_seg_kernel: passes, cycles per segment (ranking / offsets / placement):");
			for (int K = 0; K < 4; ++K) {
				fprintf(stderr, "  K=%d %.0f / %.0f / %.0f", K, tot[16 + 4 * K] / (tot[8] > 0 ? tot[8] : 1),
				        tot[17 + 4 * K] / (tot[8] > 0 ? tot[8] : 1), tot[18 + 4 * K] / (tot[8] > 0 ? tot[8] : 1));
			}
			/* (in this build warp 0 -- the one that keeps the counters -- takes 3x as long over the level-1 ranking as the
			 * other 31, which the production build does not show: its whole segment takes less than this build's passes;
			 * read the per-warp mean for that pass) */
			fprintf(stderr, ";  level-1 ranking per warp: mean %.0f, slowest warp of a CTA %.0f;",
			        [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 47]; return v;}() / 32.0 / (tot[8] > 0 ? tot[8] : 1),
			        [&]{double v=0; for (unsigned b = 0; b < grid; ++b) v += (double)h[b * 64 + 46]; return v;}() / grid);
			fprintf(stderr, "  as thread 0 saw it %.0f, %.0f rare positions marked per segment\n", tot[19] / (tot[8] > 0 ? tot[8] : 1),
			        tot[23] / (tot[8] > 0 ? tot[8] : 1));
		}
		free(h);
		cudaFree(a.prof);
	}
	return cudaGetLastError();
}
