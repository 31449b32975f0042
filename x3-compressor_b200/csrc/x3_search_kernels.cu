/*
 * x3_search_kernels.cu -- forward-window LCP-histogram kernels for sm_100a.
 *
 * What is computed (reference backend.c:58-78, see include/x3_search.h):
 *   count[p][i] = #{ d in [1, D] : LCP32(p, p+d) >= i+1 },  D = W - 33
 *   Lstar[p]    = threshold selection over count[p][*]
 *
 * Kernel "bitsliced" (production).  The work is N*D byte-pair tests; it is
 * bound by integer issue slots, not HBM (2 B/position of traffic).  So the
 * kernel is organised to spend as few issue slots per pair as possible:
 *
 *   - Diagonal formulation.  For a fixed distance d let E_d[p] = (x[p] == x[p+d]).
 *     Then LCP(p, p+d) is the run of ones of E_d starting at p.  One thread owns
 *     32 consecutive positions (one machine word of E_d) and walks d = 1..D.
 *   - Bit-plane input.  The CTA stages its span of the input with one TMA bulk
 *     copy (cp.async.bulk + mbarrier), transposes it once into 8 bit-planes with
 *     warp ballots, and from then on one LOP3 compares 32 byte pairs of one
 *     plane: E = AND_j ~(A_j ^ funnelshift(B_j)).  16 ALU instructions per 32
 *     pairs instead of >= 32.
 *   - Dense levels (LCP >= 1 .. KD) are counted for all 32 positions at once in
 *     bit-sliced counters fed through a carry-save adder tree (2 LOP3 per input).
 *   - Sparse levels (LCP > KD) are rare: the word R_{KD+1} is non-zero for a few
 *     percent of (thread, d) pairs.  Those words are pushed to a per-lane queue in
 *     shared memory (no divergence in the hot loop) and drained in batches into
 *     per-position u8 histograms over the exact LCP value.
 *   - The run of ones may cross into the next word: lane l takes E_d of lane l+1
 *     with one SHFL; lane 31 of each warp is a helper that owns no positions.
 *
 * Kernel "naive" is the obvious one-thread-per-position byte loop, kept as an
 * independent on-device cross-check and as the measured starting point.
 */
#include "x3_search_kernels.cuh"
#include "x3_search_device.cuh"

#include <cstdlib>

/* ------------------------------------------------------------------------- */
/* Naive kernel                                                               */
/* ------------------------------------------------------------------------- */

#define NAIVE_T 128
#define NAIVE_CH 2048

__global__ void __launch_bounds__(NAIVE_T) x3_lcp_naive_kernel(X3SearchParams prm)
{
	__shared__ uint8_t own[NAIVE_T + 32];
	__shared__ uint8_t tile[NAIVE_T + NAIVE_CH + 32];

	const int tid = threadIdx.x;
	const unsigned long long p0 = (unsigned long long)blockIdx.x * NAIVE_T;
	const unsigned long long p = p0 + tid;
	const uint32_t D = prm.D;

	for (int k = tid; k < NAIVE_T + 32; k += NAIVE_T) {
		own[k] = prm.x[p0 + k];
	}

	uint32_t hist[33];
#pragma unroll
	for (int i = 0; i < 33; ++i) {
		hist[i] = 0;
	}

	for (unsigned long long d0 = 0; d0 <= D; d0 += NAIVE_CH) {
		__syncthreads();
		for (int k = tid; k < NAIVE_T + NAIVE_CH + 32; k += NAIVE_T) {
			tile[k] = prm.x[p0 + d0 + k];
		}
		__syncthreads();
		const unsigned long long left = (unsigned long long)D + 1 - d0;
		const int dmax = left < NAIVE_CH ? (int)left : NAIVE_CH;
		for (int dd = (d0 == 0 ? 1 : 0); dd < dmax; ++dd) {
			const int s = tid + dd;
			int l = 0;
			while (l < 32 && own[tid + l] == tile[s + l]) {
				++l;
			}
			hist[l]++;
		}
	}

	if (p < prm.n) {
		uint32_t cnt[32];
		uint32_t acc = 0;
#pragma unroll
		for (int i = 31; i >= 0; --i) {
			acc += hist[i + 1];
			cnt[i] = min(acc, 255u);
		}
		prm.lstar[p] = (uint8_t)lstar_from_counts(cnt, prm.t);
		if (prm.H != nullptr) {
			store_row(prm.H, p, cnt);
		}
	}
}

/* ------------------------------------------------------------------------- */
/* Bit-sliced diagonal kernel                                                 */
/* ------------------------------------------------------------------------- */

template <int NW, int KD, int QCAP>
struct BsCfg {
	static constexpr int T = NW * 32;        /* threads per CTA */
	static constexpr int WORDS = NW * 31;    /* owned 32-position words per CTA */
	static constexpr int P = WORDS * 32;     /* positions per CTA */
	static constexpr int MC = 256;           /* 32-distance blocks per window chunk */
	static constexpr int PWN = WORDS + MC + 1; /* plane words staged per chunk */
	static constexpr int NB = 32 - KD;       /* sparse histogram bins, L = KD+1..32 */

	static constexpr size_t OFF_BYTES = 0;
	static constexpr size_t OFF_PLO = OFF_BYTES + (size_t)32 * PWN;
	static constexpr size_t OFF_PHI = OFF_PLO + (size_t)16 * PWN;
	static constexpr size_t OFF_HIST = OFF_PHI + (size_t)16 * PWN;
	static constexpr size_t OFF_Q = OFF_HIST + (size_t)NB * 8 * T * 4;
	static constexpr size_t OFF_BAR = OFF_Q + (size_t)QCAP * T * 8;
	static constexpr size_t SMEM = OFF_BAR + 16;
};

/* Bit-sliced counter over 32 positions: value planes 1,2,4,...,128 and a sticky
 * overflow plane.  h[] are the pending carries of the carry-save adder tree. */
struct BsLevel {
	uint32_t c[8];
	uint32_t sat;
	uint32_t h[5];
};

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c)
{
	return (a & b) | (a & c) | (b & c); /* one LOP3 */
}

/* Adds the bit-vector v as input number r (0..31) of a 32-input block.  With r
 * a compile-time constant after unrolling this is the Harley-Seal network:
 * 31 full adders (2 LOP3 each) per 32 inputs plus one ripple into planes 32..128. */
__device__ __forceinline__ void bs_add_tree(BsLevel &L, uint32_t v, int r)
{
	if ((r & 1) == 0) {
		L.h[0] = v;
		return;
	}
	uint32_t x = maj3(L.c[0], L.h[0], v);
	L.c[0] ^= L.h[0] ^ v;
	if ((r & 3) == 1) {
		L.h[1] = x;
		return;
	}
	uint32_t y = maj3(L.c[1], L.h[1], x);
	L.c[1] ^= L.h[1] ^ x;
	if ((r & 7) == 3) {
		L.h[2] = y;
		return;
	}
	x = maj3(L.c[2], L.h[2], y);
	L.c[2] ^= L.h[2] ^ y;
	if ((r & 15) == 7) {
		L.h[3] = x;
		return;
	}
	y = maj3(L.c[3], L.h[3], x);
	L.c[3] ^= L.h[3] ^ x;
	if ((r & 31) == 15) {
		L.h[4] = y;
		return;
	}
	x = maj3(L.c[4], L.h[4], y);
	L.c[4] ^= L.h[4] ^ y;
	/* ripple the carry of weight 32 into planes 32, 64, 128 and the sticky bit */
#pragma unroll
	for (int j = 5; j < 8; ++j) {
		const uint32_t tcar = L.c[j] & x;
		L.c[j] ^= x;
		x = tcar;
	}
	L.sat |= x;
}

/* Plain ripple-carry add of one bit-vector (used by the masked edge blocks). */
__device__ __forceinline__ void bs_add_ripple(BsLevel &L, uint32_t v)
{
	uint32_t x = v;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		const uint32_t tcar = L.c[j] & x;
		L.c[j] ^= x;
		x = tcar;
	}
	L.sat |= x;
}

__device__ __forceinline__ uint32_t bs_value(const BsLevel &L, int b)
{
	uint32_t v = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		v |= ((L.c[j] >> b) & 1u) << j;
	}
	return ((L.sat >> b) & 1u) ? 255u : v;
}

/* match bits of 32 byte pairs from 8 bit-planes: one LOP3 per plane */
__device__ __forceinline__ uint32_t eq_planes(const uint32_t (&a)[8], const uint32_t (&b)[8])
{
	uint32_t e = ~(a[0] ^ b[0]);
#pragma unroll
	for (int j = 1; j < 8; ++j) {
		e &= ~(a[j] ^ b[j]);
	}
	return e;
}

/* Drains the per-lane queue of (E, E_next) words into the per-position u8
 * histograms over the exact LCP value (KD+1..32).  One flattened loop: every
 * iteration either fetches the lane's next entry or retires one set bit, so
 * lanes stay busy until their own work runs out. */
template <int T, int KD>
__device__ __noinline__ void bs_drain(const uint2 *q, uint8_t *hist8, int tid, uint32_t qn)
{
	uint32_t s = 0, R = 0, e = 0, eh = 0;
	for (;;) {
		if (R == 0) {
			if (s >= qn) {
				break;
			}
			const uint2 en = q[s * T + tid];
			++s;
			e = en.x;
			eh = en.y;
			R = e;
#pragma unroll
			for (int k = 1; k <= KD; ++k) {
				R &= __funnelshift_r(e, eh, k);
			}
			continue;
		}
		const int b = __ffs(R) - 1;
		R &= R - 1;
		const uint32_t v = __funnelshift_r(e, eh, b);
		const uint32_t run = (v == 0xffffffffu) ? 32u : (uint32_t)(__ffs(~v) - 1);
		const uint32_t idx = (((run - (KD + 1)) * 8 + (b >> 2)) * T + tid) * 4 + (b & 3);
		const uint32_t hv = hist8[idx];
		hist8[idx] = (uint8_t)(hv + (hv != 255u));
	}
}

template <int NW, int KD, int QCAP>
__global__ void __launch_bounds__(NW * 32, 1) x3_lcp_bitsliced_kernel(X3SearchParams prm)
{
	using C = BsCfg<NW, KD, QCAP>;
	constexpr int T = C::T;

	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *sm_bytes = smem + C::OFF_BYTES;
	uint4 *sm_plo = reinterpret_cast<uint4 *>(smem + C::OFF_PLO);
	uint4 *sm_phi = reinterpret_cast<uint4 *>(smem + C::OFF_PHI);
	uint32_t *sm_hist = reinterpret_cast<uint32_t *>(smem + C::OFF_HIST);
	uint2 *sm_q = reinterpret_cast<uint2 *>(smem + C::OFF_Q);
	uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;
	const int wi = warp * 31 + lane; /* this thread's word of positions in the CTA tile */
	const unsigned long long p0 = (unsigned long long)blockIdx.x * C::P;
	const uint32_t D = prm.D;

	for (int i = tid; i < C::NB * 8 * T; i += T) {
		sm_hist[i] = 0;
	}
	if (tid == 0) {
		mbar_init(bar, 1);
	}
	__syncthreads();

	const uint32_t MB = D / 32 + 1; /* 32-distance blocks 0..D/32; d = 32*m + r */
	const uint32_t nchunks = (MB + C::MC - 1) / C::MC;

	uint32_t a[8];
	BsLevel lv[KD];
#pragma unroll
	for (int k = 0; k < KD; ++k) {
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			lv[k].c[j] = 0;
		}
		lv[k].sat = 0;
#pragma unroll
		for (int j = 0; j < 5; ++j) {
			lv[k].h[j] = 0;
		}
	}
	uint32_t qn = 0;
	uint32_t phase = 0;

	if (tid == 0) {
		const uint32_t mcount = min((uint32_t)C::MC, MB);
		const uint32_t bytes = 32u * (C::WORDS + mcount + 1);
		mbar_expect_tx(bar, bytes);
		tma_load_1d(sm_bytes, prm.x + p0, bytes, bar);
	}

	for (uint32_t c = 0; c < nchunks; ++c) {
		const uint32_t mcount = min((uint32_t)C::MC, MB - c * C::MC);
		const uint32_t nstage = C::WORDS + mcount + 1;

		/* stage -> bit-planes.  Bit l of plane word k is bit j of byte 32k+l. */
		mbar_wait(bar, phase);
		phase ^= 1;
		for (uint32_t k = warp; k < nstage; k += NW) {
			const uint32_t byte = sm_bytes[32 * k + lane];
			uint32_t bal[8];
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				bal[j] = __ballot_sync(FULL_MASK, (byte >> j) & 1u);
			}
			if (lane == 0) {
				sm_plo[k] = make_uint4(bal[0], bal[1], bal[2], bal[3]);
			}
			if (lane == 1) {
				sm_phi[k] = make_uint4(bal[4], bal[5], bal[6], bal[7]);
			}
		}
		__syncthreads();

		/* the byte buffer is free again: prefetch the next chunk under the compute */
		if (tid == 0 && c + 1 < nchunks) {
			const uint32_t mnext = min((uint32_t)C::MC, MB - (c + 1) * C::MC);
			const uint32_t bytes = 32u * (C::WORDS + mnext + 1);
			mbar_expect_tx(bar, bytes);
			tma_load_1d(sm_bytes, prm.x + p0 + 32ull * (c + 1) * C::MC, bytes, bar);
		}

		uint32_t cur[8];
		{
			const uint4 lo = sm_plo[wi], hi = sm_phi[wi];
			cur[0] = lo.x; cur[1] = lo.y; cur[2] = lo.z; cur[3] = lo.w;
			cur[4] = hi.x; cur[5] = hi.y; cur[6] = hi.z; cur[7] = hi.w;
		}
		if (c == 0) {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				a[j] = cur[j];
			}
		}

		for (uint32_t m = 0; m < mcount; ++m) {
			uint32_t nxt[8];
			{
				const uint4 lo = sm_plo[wi + m + 1], hi = sm_phi[wi + m + 1];
				nxt[0] = lo.x; nxt[1] = lo.y; nxt[2] = lo.z; nxt[3] = lo.w;
				nxt[4] = hi.x; nxt[5] = hi.y; nxt[6] = hi.z; nxt[7] = hi.w;
			}
			const uint32_t dbase = 32u * (c * C::MC + m);

			if (dbase >= 1 && dbase + 31 <= D) {
				/* interior block: all 32 distances valid, fully unrolled */
#pragma unroll
				for (int r = 0; r < 32; ++r) {
					uint32_t b[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						b[j] = r == 0 ? cur[j] : __funnelshift_r(cur[j], nxt[j], r);
					}
					const uint32_t e = eq_planes(a, b);
					const uint32_t eh = __shfl_down_sync(FULL_MASK, e, 1);
					uint32_t R = e;
					bs_add_tree(lv[0], R, r);
#pragma unroll
					for (int k = 1; k < KD; ++k) {
						R &= __funnelshift_r(e, eh, k);
						bs_add_tree(lv[k], R, r);
					}
					R &= __funnelshift_r(e, eh, KD);
					if (R != 0 && lane != 31) {
						sm_q[qn * T + tid] = make_uint2(e, eh);
						++qn;
					}
				}
			} else {
				/* edge block (d = 0 or d > D inside): masked, rolled */
#pragma unroll 1
				for (int r = 0; r < 32; ++r) {
					const uint32_t d = dbase + r;
					const bool valid = d >= 1 && d <= D;
					uint32_t b[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						b[j] = __funnelshift_r(cur[j], nxt[j], r);
					}
					const uint32_t e = valid ? eq_planes(a, b) : 0u;
					const uint32_t eh = __shfl_down_sync(FULL_MASK, e, 1);
					uint32_t R = e;
					bs_add_ripple(lv[0], R);
#pragma unroll
					for (int k = 1; k < KD; ++k) {
						R &= __funnelshift_r(e, eh, k);
						bs_add_ripple(lv[k], R);
					}
					R &= __funnelshift_r(e, eh, KD);
					if (R != 0 && lane != 31) {
						sm_q[qn * T + tid] = make_uint2(e, eh);
						++qn;
					}
				}
			}

#pragma unroll
			for (int j = 0; j < 8; ++j) {
				cur[j] = nxt[j];
			}
			if (__any_sync(FULL_MASK, qn > QCAP - 32)) {
				bs_drain<T, KD>(sm_q, reinterpret_cast<uint8_t *>(sm_hist), tid, qn);
				qn = 0;
			}
		}
		__syncthreads(); /* planes are rewritten by the next chunk */
	}

	bs_drain<T, KD>(sm_q, reinterpret_cast<uint8_t *>(sm_hist), tid, qn);

	/* epilogue: counts -> Lstar (and the 32-bin row) for this thread's 32 positions */
	const unsigned long long pbase = p0 + 32ull * wi;
	if (lane != 31 && pbase < prm.n) {
		const uint8_t *hist8 = reinterpret_cast<const uint8_t *>(sm_hist);
#pragma unroll 1
		for (int b = 0; b < 32; ++b) {
			const unsigned long long p = pbase + b;
			uint32_t cnt[32];
			uint32_t acc = 0;
#pragma unroll
			for (int L = 32; L > KD; --L) {
				acc += hist8[(((L - KD - 1) * 8 + (b >> 2)) * T + tid) * 4 + (b & 3)];
				acc = min(acc, 255u);
				cnt[L - 1] = acc;
			}
#pragma unroll
			for (int k = 0; k < KD; ++k) {
				cnt[k] = bs_value(lv[k], b);
			}
			if (p < prm.n) {
				prm.lstar[p] = (uint8_t)lstar_from_counts(cnt, prm.t);
				if (prm.H != nullptr) {
					store_row(prm.H, p, cnt);
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------- */
/* Host side                                                                  */
/* ------------------------------------------------------------------------- */

typedef BsCfg<4, 2, 64> BsDefault;

size_t x3k_required_bytes(size_t n, size_t W)
{
	/* worst case over the variants: tile round-up (< 8192) + window + staging slack */
	return n + W + 16384;
}

cudaError_t x3k_stream_init_device(void);
cudaError_t x3k_launch_stream(bool full, X3SearchParams prm, cudaStream_t stream, int *launches);

cudaError_t x3k_init_device(void)
{
	cudaError_t e = cudaFuncSetAttribute(x3_lcp_bitsliced_kernel<4, 2, 64>,
	                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BsDefault::SMEM);
	if (e != cudaSuccess) {
		return e;
	}
	return x3k_stream_init_device();
}

/* variant: 0/3 stream (production), 1 naive, 2 bitsliced (first version), 4 stream with u8 counters, 5 rank */
int x3k_default_kind(uint32_t D, int t, bool want_table)
{
	if (want_table || D > x3k_rank_max_distances()) {
		return 0;
	}
	if (D >= 1 && D <= x3k_seg_max_distances() && t >= x3k_seg_min_t() && getenv("X3_NO_SEG") == nullptr) {
		return 2;
	}
	return 1;
}

cudaError_t x3k_launch(int variant, const X3SearchParams &prm, cudaStream_t stream, int *launches)
{
	if (prm.n == 0) {
		return cudaSuccess;
	}
	if (launches != nullptr && (variant == 1 || variant == 2)) {
		*launches += 1;
	}
	/* production choice when only Lstar is wanted: the segment search while the window fits on chip
	 * (one launch, HBM traffic = input + result), else the rank search (cost independent of the
	 * window and of t); the brute-force stream kernel for the 32-bin table and for windows beyond
	 * the rank search's chunking */
	const int kind = variant == 0 ? x3k_default_kind(prm.D, prm.t, prm.H != nullptr) : -1;
	if (variant == 6 || kind == 2) {
		return x3k_launch_seg(prm, stream, launches);
	}
	if (variant == 5 || kind == 1) {
		return x3k_launch_rank(prm, stream, launches);
	}
	if (variant == 1) {
		const unsigned long long grid = (prm.n + NAIVE_T - 1) / NAIVE_T;
		x3_lcp_naive_kernel<<<(unsigned)grid, NAIVE_T, 0, stream>>>(prm);
	} else if (variant == 2) {
		const unsigned long long grid = (prm.n + BsDefault::P - 1) / BsDefault::P;
		x3_lcp_bitsliced_kernel<4, 2, 64><<<(unsigned)grid, BsDefault::T, BsDefault::SMEM, stream>>>(prm);
	} else {
		/* t <= 0: the selection never runs (backend.c:76) -- the fast path handles it, Lstar = 0 */
		const bool full = prm.H != nullptr || prm.t > 15 || variant == 4;
		return x3k_launch_stream(full, prm, stream, launches);
	}
	return cudaGetLastError();
}
