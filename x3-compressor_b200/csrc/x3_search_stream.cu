/*
 * x3_search_stream.cu -- "stream" kernel: the production forward-window
 * LCP-histogram search for sm_100a.
 *
 * Computes what reference backend.c:58-78 computes (see include/x3_search.h):
 *   count[p][i] = #{ d in [1, D] : LCP32(p, p+d) >= i+1 },  D = W - 33
 *   Lstar[p]    = threshold selection over count[p][*]
 *
 * The work is N*D byte-pair tests and is bound by the integer ALU pipe, so the
 * kernel is organised around ALU instructions per pair, not around bytes:
 *
 *   - Bit-plane input.  A tile of the input is staged with one TMA bulk copy
 *     (cp.async.bulk + mbarrier) and transposed into 8 bit-planes; one LOP3 then
 *     compares 32 byte pairs of one plane, 8 LOP3 give the match word
 *     E_d[w] = (x[p] == x[p+d]) for the 32 positions p of plane word w.
 *   - One warp = one independent worker (a 32-thread CTA with 17 KB of shared
 *     memory, 12 resident per SM): no CTA-wide barrier exists in the search loop,
 *     so a warp that is busy with the rare-event path never stalls its
 *     neighbours, and the 3 warps that share an SM sub-partition fill each
 *     other's latency.
 *   - Distance d = 32 m + r.  For each r the warp keeps the window planes shifted
 *     by r bits in shared memory; going from r to r+1 is one in-place 1-bit funnel
 *     shift of the array (8 SHF per plane word, amortised over all 62 position
 *     words), so the inner loop over m has no alignment shifts at all: 2 LDS.128
 *     bring the next shifted word, which serves both position words of the thread.
 *   - Each lane owns 2 consecutive position words (64 positions): the match word
 *     of word 1 is the run continuation of word 0, the neighbour lane supplies
 *     the continuation of word 1 with one SHFL.  Lane 31 is a helper that owns
 *     no positions.
 *   - Levels LCP >= 1 and >= 2 are counted for all 32 positions at once in
 *     bit-sliced counters fed by a 16-input carry-save adder tree
 *     (~1.9 LOP3 per input).
 *   - Levels LCP >= 3 are rare (0.35 % of pairs on text) but very unevenly spread
 *     over positions (frequent trigrams): such match words go to a per-lane queue
 *     in shared memory, and the warp drains all queues COOPERATIVELY: the entries
 *     are dealt out evenly over the 32 lanes (prefix sum + owner search by SHFL),
 *     and each event updates an exact-LCP histogram with an atomic add: LCP 3..7
 *     (3..4) packed into one shared-memory word per position, deeper ones in a
 *     per-position row of global scratch that stays in L2.  A position whose
 *     LCP-32 bin reaches the cap is masked out of the queue filter, so long
 *     zero/periodic runs stop generating events.
 *   - Shared-memory plane arrays are split by word parity and plane half, so the
 *     LDS.128 of 32 lanes that each own words 2*lane, 2*lane+1 are contiguous
 *     (no bank conflicts).
 *   - Persistent CTAs fetch tiles from an atomic counter (no tail wave).
 *
 * Two instantiations:
 *   <4,6>   FAST: t <= 15, every counter saturates at 16 (lossless for Lstar because
 *           the selection only evaluates count > tc with tc <= t, backend.c:78)
 *   <8,15>  FULL: counters exact up to 255, any t <= 254, exact H rows
 */
#include "x3_search_device.cuh"

#include <cstdio>
#include <cstdlib>

namespace {

template <int CB, int HB, int KD_>
struct SCfg {
	static constexpr int OWN = 62;                 /* position words per warp: 31 lanes x 2 */
	static constexpr int NWORD = 64;               /* + the helper lane's two words */
	static constexpr int P = OWN * 32;             /* positions per tile */
	static constexpr int MCH = 48;                 /* 32-distance blocks per window chunk (3 groups of 16) */
	static constexpr int NPW = NWORD + MCH + 1;    /* plane words staged per chunk */
	static constexpr int SEG0 = (NPW + 1) / 2;     /* uint4 per (even word, half) segment of the plane array */
	static constexpr int SEG1 = NPW / 2;           /* uint4 per (odd word, half) segment */
	static constexpr int KD = KD_;                 /* LCP levels 1..KD are counted densely (bit-sliced): 2 or 3 */
	static constexpr int L0 = KD + 1;              /* first LCP level that goes through the event queue */
	static constexpr int NSH = HB == 6 ? 5 : 2;    /* LCP levels kept in the shared word: L0 .. L0+NSH-1 (bits
	                                                * 0..29; bit 31 flags a touched deep row) */
	static constexpr int NDEEP = 32 - KD - NSH;    /* LCP levels kept in the global row: L0+NSH .. 32 */
	static constexpr int DBITS = 16;               /* bits per deep bin: never overflows while D <= 65535 */
	static constexpr int ROWB = 64;                /* bytes per deep row */
	static constexpr uint32_t CAP = HB == 6 ? 16u : 255u;
	static constexpr uint32_t FMASK = (1u << HB) - 1u;
	static constexpr uint32_t DMASK = (1u << DBITS) - 1u;
	static constexpr int QCAP = 16;                /* queue slots per lane; checked every 4 blocks */

	/* 16 896 bytes: 13 CTAs (warps) per SM (shared memory is granted in 256-byte units on top
	 * of 1 KB per CTA) */
	static constexpr size_t OFF_PW = 0;
	static constexpr size_t OFF_Q = OFF_PW + (size_t)2 * (SEG0 + SEG1) * 16; /* queue; also the byte staging buffer */
	static constexpr size_t OFF_HIST = OFF_Q + (size_t)QCAP * 256;
	static constexpr size_t OFF_TAB = OFF_HIST + (size_t)62 * 32 * 4;  /* drain index table, 31 * QCAP entries */
	static constexpr size_t OFF_DONE = OFF_TAB + (size_t)31 * QCAP * 2;
	static constexpr size_t OFF_BAR = OFF_DONE + (size_t)62 * 4;
	static constexpr size_t SMEM = OFF_BAR + 8;
	static_assert((size_t)NPW * 32 <= (size_t)QCAP * 256, "staging buffer fits the drained queue");
	static_assert(MCH % 16 == 0, "whole groups of 16 blocks");
	static_assert(SMEM <= 16896, "13 CTAs per SM");
	static_assert(OFF_Q % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
	static_assert(NDEEP * DBITS <= ROWB * 8, "deep row holds every deep bin");
};

static_assert(SCfg<4, 6, 2>::P == X3K_STREAM_TILE, "tile size");
static_assert((size_t)SCfg<8, 15, 3>::ROWB * SCfg<8, 15, 3>::P <= X3K_DEEP_BYTES_PER_CTA, "deep scratch per CTA");

/* Bit-sliced counter over 32 positions fed by a 16-input carry-save tree. */
template <int CB>
struct Tree {
	uint32_t c[CB];
	uint32_t sat;
	uint32_t h[4];
};

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c)
{
	return (a & b) | (a & c) | (b & c);
}

template <int CB>
__device__ __forceinline__ void tree_clear(Tree<CB> &T)
{
#pragma unroll
	for (int j = 0; j < CB; ++j) {
		T.c[j] = 0;
	}
	T.sat = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		T.h[j] = 0;
	}
}

/* input number i (0..15, compile-time after unrolling) of a 16-input block */
template <int CB>
__device__ __forceinline__ void tree_add(Tree<CB> &T, uint32_t v, int i)
{
	if ((i & 1) == 0) {
		T.h[0] = v;
		return;
	}
	uint32_t x = maj3(T.c[0], T.h[0], v);
	T.c[0] ^= T.h[0] ^ v;
	if ((i & 3) == 1) {
		T.h[1] = x;
		return;
	}
	uint32_t y = maj3(T.c[1], T.h[1], x);
	T.c[1] ^= T.h[1] ^ x;
	if ((i & 7) == 3) {
		T.h[2] = y;
		return;
	}
	x = maj3(T.c[2], T.h[2], y);
	T.c[2] ^= T.h[2] ^ y;
	if ((i & 15) == 7) {
		T.h[3] = x;
		return;
	}
	y = maj3(T.c[3], T.h[3], x);
	T.c[3] ^= T.h[3] ^ x;
	/* y carries weight 16 */
#pragma unroll
	for (int j = 4; j < CB; ++j) {
		const uint32_t tcar = T.c[j] & y;
		T.c[j] ^= y;
		y = tcar;
	}
	T.sat |= y;
}

template <int CB>
__device__ __forceinline__ uint32_t tree_value(const Tree<CB> &T, int b)
{
	uint32_t v = 0;
#pragma unroll
	for (int j = 0; j < CB; ++j) {
		v |= ((T.c[j] >> b) & 1u) << j;
	}
	return ((T.sat >> b) & 1u) ? (CB == 4 ? 16u : 255u) : v;
}

/* match word of 32 byte pairs; na[] holds the COMPLEMENTED planes of the owned word,
 * so every plane is one LOP3: e &= (na ^ b) */
__device__ __forceinline__ uint32_t eq8(const uint32_t (&na)[8], const uint32_t (&b)[8])
{
	uint32_t e = na[0] ^ b[0];
#pragma unroll
	for (int j = 1; j < 8; ++j) {
		e &= na[j] ^ b[j];
	}
	return e;
}

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b)
{
	asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b));
}

/* 8 bytes (a: bytes 0..3, b: bytes 4..7) -> plo: byte j = bit j of the 8 bytes (byte i -> bit i),
 * j = 0..3, phi: the same for j = 4..7. */
__device__ __forceinline__ void transpose8x8(uint32_t a, uint32_t b, uint32_t &plo, uint32_t &phi)
{
	uint32_t x = b, y = a, t;
	t = (x ^ (x >> 7)) & 0x00AA00AAu;  x = x ^ t ^ (t << 7);
	t = (y ^ (y >> 7)) & 0x00AA00AAu;  y = y ^ t ^ (t << 7);
	t = (x ^ (x >> 14)) & 0x0000CCCCu; x = x ^ t ^ (t << 14);
	t = (y ^ (y >> 14)) & 0x0000CCCCu; y = y ^ t ^ (t << 14);
	t = (x & 0xF0F0F0F0u) | ((y >> 4) & 0x0F0F0F0Fu);
	y = ((x << 4) & 0xF0F0F0F0u) | (y & 0x0F0F0F0Fu);
	plo = y;
	phi = t;
}

/* p[g] byte j -> out_j byte g */
__device__ __forceinline__ void bytes4x4(const uint32_t (&p)[4], uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
	const uint32_t u0 = __byte_perm(p[0], p[1], 0x5140), u1 = __byte_perm(p[2], p[3], 0x5140);
	const uint32_t u2 = __byte_perm(p[0], p[1], 0x7362), u3 = __byte_perm(p[2], p[3], 0x7362);
	o0 = __byte_perm(u0, u1, 0x5410);
	o1 = __byte_perm(u0, u1, 0x7632);
	o2 = __byte_perm(u2, u3, 0x5410);
	o3 = __byte_perm(u2, u3, 0x7632);
}

/* Plane array in shared memory, split by word parity and plane half so that lanes owning
 * words 2*lane, 2*lane+1 read contiguous uint4: word k, half h (planes 0-3 / 4-7) lives at
 * uint4 index segbase(k & 1, h) + (k >> 1). */
template <class C>
__device__ __forceinline__ constexpr int segbase(int par, int h)
{
	return par == 0 ? h * C::SEG0 : 2 * C::SEG0 + h * C::SEG1;
}

template <class C>
__device__ __forceinline__ int pidx(int k, int h)
{
	return segbase<C>(k & 1, h) + (k >> 1);
}

template <class C>
__device__ __forceinline__ void load_word(const uint4 *base, int k, uint32_t (&w)[8])
{
	const uint4 lo = base[pidx<C>(k, 0)], hi = base[pidx<C>(k, 1)];
	w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w;
	w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
}

template <class C>
__device__ __forceinline__ void store_word(uint4 *base, int k, const uint32_t (&w)[8])
{
	base[pidx<C>(k, 0)] = make_uint4(w[0], w[1], w[2], w[3]);
	base[pidx<C>(k, 1)] = make_uint4(w[4], w[5], w[6], w[7]);
}

/*
 * Cooperative drain.  Lane l holds n0 word-0 entries (slots 0 .. n0-1) and n1 word-1
 * entries (slots QCAP-1 .. QCAP-n1) of (E, E_next) match words.  The queues are very
 * uneven (frequent trigrams) and so are the numbers of events per entry, so the work
 * is dealt out dynamically: entries are numbered by an exclusive prefix sum over lanes
 * and listed in a small index table; in every iteration of the loop each idle lane
 * takes the next unclaimed entry (ballot + popc on a warp-uniform cursor) and every
 * lane retires ONE event of the entry it holds.  An event is one set bit of the
 * entry's LCP>=3 word: an atomic add into the exact-LCP bin of that position (several
 * lanes may hit the same position).  Bins may overshoot the cap by at most 31 (one add
 * in flight per lane), which the field widths absorb and the epilogue clamps.
 */
template <int CB, int HB, int KD>
__device__ __noinline__ void st_drain(const uint2 *q, uint32_t n0, uint32_t n1, uint32_t *hist, uint32_t *done_s,
                                      uint16_t *tab, uint8_t *deep_tile, int lane, bool uncond)
{
	using C = SCfg<CB, HB, KD>;
	const uint32_t n = n0 + n1;
	uint32_t inc = n;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= d) {
			inc += t;
		}
	}
	const uint32_t excl = inc - n;
	const uint32_t E = __shfl_sync(FULL_MASK, inc, 31);
	if (E == 0) {
		return;
	}
	/* index table: entry number -> (word j, queue address slot * 32 + owner lane) */
	for (uint32_t s = 0; s < n; ++s) {
		const uint32_t j = s >= n0 ? 1u : 0u;
		const uint32_t slot = j ? (uint32_t)C::QCAP - 1u - (s - n0) : s;
		tab[excl + s] = (uint16_t)((j << 9) | (slot * 32u + (uint32_t)lane));
	}
	__syncwarp();

	const uint32_t lt = (1u << lane) - 1u;
	uint32_t cursor = 0; /* warp-uniform: entries claimed so far */
	uint32_t R = 0, e = 0, eh = 0, w = 0;
	for (;;) {
		const bool need = R == 0;
		const uint32_t nmask = __ballot_sync(FULL_MASK, need);
		if (nmask == FULL_MASK && cursor >= E) {
			break;
		}
		if (need) {
			const uint32_t idx = cursor + __popc(nmask & lt);
			if (idx < E) {
				const uint32_t t = tab[idx];
				const uint2 en = q[t & 511u];
				w = 2u * (t & 31u) + (t >> 9);
				e = en.x;
				eh = en.y;
				R = e & __funnelshift_r(e, eh, 1) & __funnelshift_r(e, eh, 2) & ~done_s[w];
				if (C::KD >= 3) {
					R &= __funnelshift_r(e, eh, 3);
				}
			}
		}
		cursor += __popc(nmask);
		if (R != 0) {
			/* one event: highest set bit b; v = the pair's bits from b upwards (bits 0..KD are set) */
			const int cz = __clz(R);
			const int b = 31 - cz;
			R &= ~(0x80000000u >> cz);
			const uint32_t v = __funnelshift_r(e, eh, b);
			const uint32_t pos = w * 32u + (uint32_t)b;
			const uint32_t word = hist[pos];
			/* zeros among bits L0 .. L0+NSH-1 of v: the lowest one marks the run length */
			const uint32_t y = ~(v >> C::L0) & ((1u << C::NSH) - 1u);
			if (y != 0) {
				uint32_t inc;
				bool full;
				if (HB == 6) {
					/* run - 3 = log2(z) for the one-hot z; the field increment 1 << 6 (run - 3) is z^6,
					 * computed by 3 multiplies on the FMA pipe (FLO/POPC share the slow XU pipe) */
					const uint32_t z = y & (0u - y);
					const uint32_t z3 = z * z * z;
					inc = z3 * z3;
					full = (word & (inc * 48u)) != 0; /* bits 4 and 5 of the field: value >= 16 */
				} else {
					const uint32_t sh = HB * (uint32_t)(__ffs(y) - 1);
					inc = 1u << sh;
					full = ((word >> sh) & C::FMASK) >= C::CAP;
				}
				if (!full) {
					atomicAdd(&hist[pos], inc);
				}
			} else {
				const uint32_t run = (v == 0xffffffffu) ? 32u : (uint32_t)(__ffs(~v) - 1);
				if ((word >> 31) == 0) {
					atomicOr(&hist[pos], 0x80000000u); /* row touched */
				}
				const uint32_t k = run - (uint32_t)(C::L0 + C::NSH);
				constexpr uint32_t PER = 32 / C::DBITS; /* bins per 32-bit word of the row */
				unsigned int *rw = reinterpret_cast<unsigned int *>(deep_tile + (size_t)pos * C::ROWB) + k / PER;
				const uint32_t sh = C::DBITS * (k % PER);
				if (uncond && run != 32u) {
					/* a 16-bit bin cannot overflow while D <= 65535: fire-and-forget reduction, no
					 * L2 round trip on the critical path; the epilogue clamps */
					atomicAdd(rw, 1u << sh);
				} else if (uncond) {
					/* LCP-32 bin: the returned count tells when every level is saturated */
					const uint32_t cv = (atomicAdd(rw, 1u << sh) >> sh) & C::DMASK;
					if (cv + 1u >= C::CAP) {
						atomicOr(&done_s[w], 1u << b);
					}
				} else {
					const uint32_t cv = (__ldcg(rw) >> sh) & C::DMASK;
					if (cv < C::CAP) {
						atomicAdd(rw, 1u << sh);
					}
					if (run == 32u && cv + 1u >= C::CAP) {
						atomicOr(&done_s[w], 1u << b); /* every level of this position is saturated */
					}
				}
			}
		}
	}
	__syncwarp();
}

/* State a lane carries through the search loop. */
template <int CB, int KD>
struct LaneState {
	uint32_t A0[8], A1[8]; /* COMPLEMENTED planes of the two owned position words */
	Tree<CB> T[2 * KD]; /* [word j][level k] at index j * KD + k: counts of LCP >= k + 1 */
	uint2 done;
	bool uncond;           /* D <= 65535: deep bins are added to without a bound check */
	uint32_t q0, q1;       /* shared-space byte address of the next free slot of the word-0 queue (grows
	                        * up) and of the word-1 queue (grows down) */
};

template <int CB, int HB, int KD>
__device__ __forceinline__ void st_flush(LaneState<CB, KD> &st, uint2 *q, uint32_t *hist, uint32_t *done_s,
                                         uint8_t *deep_tile, int lane)
{
	uint16_t *tab = reinterpret_cast<uint16_t *>(hist + 62 * 32); /* OFF_TAB follows the histograms */
	using C = SCfg<CB, HB, KD>;
	const uint32_t lo = smem_u32(q + lane), hi = smem_u32(q + (C::QCAP - 1) * 32 + lane);
	const uint32_t n0 = (st.q0 - lo) / 256u;
	const uint32_t n1 = (hi - st.q1) / 256u;
	/* the helper lane owns no positions: its entries are dropped */
	st_drain<CB, HB, KD>(q, lane == 31 ? 0u : n0, lane == 31 ? 0u : n1, hist, done_s, tab, deep_tile, lane, st.uncond);
	if (lane != 31) {
		st.done = make_uint2(done_s[2 * lane], done_s[2 * lane + 1]);
	}
	st.q0 = lo;
	st.q1 = hi;
}

/*
 * 16 consecutive distance blocks mm0 .. mm0+15 (chunk relative) at a fixed r.
 * MASKED: blocks outside [vlo, vhi] contribute nothing (d = 0 or d > D).
 */
template <int CB, int HB, int KD, bool MASKED>
__device__ __forceinline__ void st_group16(LaneState<CB, KD> &st, const uint4 *sr, uint2 *q, uint32_t *hist,
                                           uint32_t *done_s, uint8_t *deep_tile, int lane, int wA, int mm0, int vlo,
                                           int vhi)
{
	using C = SCfg<CB, HB, KD>;
	uint32_t S0[8], S1[8];
	/* mm0 is a multiple of 16 and wA = 2 * lane: the parity of every plane word index below is a
	 * compile-time constant after unrolling, and the 32 lanes read contiguous uint4 */
	const uint4 *srl = sr + lane + (mm0 >> 1);
	{
		const uint4 lo = srl[segbase<C>(0, 0)], hi = srl[segbase<C>(0, 1)]; /* word wA + mm0: even */
		S0[0] = lo.x; S0[1] = lo.y; S0[2] = lo.z; S0[3] = lo.w;
		S0[4] = hi.x; S0[5] = hi.y; S0[6] = hi.z; S0[7] = hi.w;
	}
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		{
			/* word wA + 1 + mm0 + i */
			const int par = (1 + i) & 1, off = (1 + i) >> 1;
			const uint4 lo = srl[segbase<C>(par, 0) + off], hi = srl[segbase<C>(par, 1) + off];
			S1[0] = lo.x; S1[1] = lo.y; S1[2] = lo.z; S1[3] = lo.w;
			S1[4] = hi.x; S1[5] = hi.y; S1[6] = hi.z; S1[7] = hi.w;
		}
		uint32_t e0 = eq8(st.A0, S0);
		uint32_t e1 = eq8(st.A1, S1);
		if (MASKED) {
			const bool valid = (mm0 + i >= vlo) && (mm0 + i <= vhi);
			e0 = valid ? e0 : 0u;
			e1 = valid ? e1 : 0u;
		}
		const uint32_t en = __shfl_down_sync(FULL_MASK, e0, 1);
		const uint32_t r20 = e0 & __funnelshift_r(e0, e1, 1);
		const uint32_t r21 = e1 & __funnelshift_r(e1, en, 1);
		tree_add(st.T[0 * C::KD + 0], e0, i);
		tree_add(st.T[0 * C::KD + 1], r20, i);
		tree_add(st.T[1 * C::KD + 0], e1, i);
		tree_add(st.T[1 * C::KD + 1], r21, i);
		uint32_t r30, r31; /* the queue filter: LCP >= KD + 1 at a position that is not done */
		if (C::KD == 2) {
			r30 = r20 & __funnelshift_r(e0, e1, 2) & ~st.done.x;
			r31 = r21 & __funnelshift_r(e1, en, 2) & ~st.done.y;
		} else {
			const uint32_t t30 = r20 & __funnelshift_r(e0, e1, 2);
			const uint32_t t31 = r21 & __funnelshift_r(e1, en, 2);
			tree_add(st.T[0 * C::KD + 2], t30, i);
			tree_add(st.T[1 * C::KD + 2], t31, i);
			r30 = t30 & __funnelshift_r(e0, e1, 3) & ~st.done.x;
			r31 = t31 & __funnelshift_r(e1, en, 3) & ~st.done.y;
		}
		if (r30 != 0) {
			sts64(st.q0, e0, e1);
			st.q0 += 256;
		}
		if (r31 != 0) {
			sts64(st.q1, e1, en);
			st.q1 -= 256;
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			S0[j] = S1[j];
		}
		if ((i & 3) == 3) {
			/* 4 more blocks can push 4 entries at each end: flush when fewer than 8 slots are free
			 * (signed compare: a full queue gives -256) */
			if (__any_sync(FULL_MASK, (int)(st.q1 - st.q0) < 256 * 7)) {
				st_flush<CB, HB, KD>(st, q, hist, done_s, deep_tile, lane);
			}
		}
	}
	(void)wA;
}

/* Stages plane words [32 MCH c, +NPW) of the tile: one TMA bulk copy into the (drained)
 * queue memory, then every lane transposes whole 32-byte words into bit-planes: four 8x8
 * bit-matrix transposes (SWAR, Hacker's Delight 7-3) and a 4x4 byte transpose with PRMT. */
template <class C>
__device__ __forceinline__ void st_stage(const X3SearchParams &prm, uint8_t *smem, unsigned long long p0, uint32_t c,
                                         uint32_t &phase, int lane)
{
	uint4 *pw = reinterpret_cast<uint4 *>(smem + C::OFF_PW);
	uint8_t *stage = smem + C::OFF_Q;
	uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
	__syncwarp();
	if (lane == 0) {
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		mbar_expect_tx(bar, C::NPW * 32);
		tma_load_1d(stage, prm.x + p0 + 32ull * C::MCH * c, C::NPW * 32, bar);
	}
	mbar_wait(bar, phase);
	phase ^= 1;
	for (int k = lane; k < C::NPW; k += 32) {
		const uint4 *src = reinterpret_cast<const uint4 *>(stage + 32 * k);
		const uint4 lo = src[0], hi = src[1];
		const uint32_t in[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
		uint32_t pl[4], ph[4];
#pragma unroll
		for (int g = 0; g < 4; ++g) {
			transpose8x8(in[2 * g], in[2 * g + 1], pl[g], ph[g]);
		}
		uint32_t w[8];
		bytes4x4(pl, w[0], w[1], w[2], w[3]);
		bytes4x4(ph, w[4], w[5], w[6], w[7]);
		store_word<C>(pw, k, w);
	}
	__syncwarp();
}

/* Window planes: shifts the whole array by one more bit, in place.  Word k takes its top bit
 * from word k+1; ascending passes read a word before the pass that owns it rewrites it; the last
 * word shifts in zeros. */
template <class C>
__device__ __forceinline__ void st_shift1(uint4 *pw, int lane)
{
	for (int pass = 0; 64 * pass < C::NPW; ++pass) { /* uniform trip count: __syncwarp inside */
		const int k2 = lane + 32 * pass;
		uint32_t a[8], b[8], c2[8], s[8];
		const bool hasa = 2 * k2 < C::NPW, hasb = 2 * k2 + 1 < C::NPW, hasc = 2 * k2 + 2 < C::NPW;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			a[j] = 0;
			b[j] = 0;
			c2[j] = 0;
		}
		if (hasa) {
			load_word<C>(pw, 2 * k2, a);
		}
		if (hasb) {
			load_word<C>(pw, 2 * k2 + 1, b);
		}
		if (hasc) {
			load_word<C>(pw, 2 * k2 + 2, c2);
		}
		__syncwarp();
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			s[j] = __funnelshift_r(a[j], b[j], 1);
		}
		if (hasa) {
			store_word<C>(pw, 2 * k2, s);
		}
		if (hasb) {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				s[j] = __funnelshift_r(b[j], c2[j], 1);
			}
			store_word<C>(pw, 2 * k2 + 1, s);
		}
	}
}

/*
 * Probe: how many (lane, word) units of distance blocks 1..15 at r = 0..7 would push a queue
 * entry with 2 dense levels.  Chunk 0 must be staged; the plane array is left shifted by 7.
 */
#define X3_PROBE_R 8
template <class C>
__device__ __forceinline__ uint32_t st_probe(uint4 *pw, int lane)
{
	uint32_t A0[8], A1[8], S0[8], S1[8];
	load_word<C>(pw, 2 * lane, A0);
	load_word<C>(pw, 2 * lane + 1, A1);
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		A0[j] = ~A0[j];
		A1[j] = ~A1[j];
	}
	uint32_t pushes = 0;
	for (int r = 0; r < X3_PROBE_R; ++r) {
		if (r > 0) {
			__syncwarp();
			st_shift1<C>(pw, lane);
			__syncwarp();
		}
		load_word<C>(pw, 2 * lane + 1, S0); /* block 1 of word 0 */
#pragma unroll 3
		for (int i = 1; i < 16; ++i) {
			load_word<C>(pw, 2 * lane + 1 + i, S1);
			const uint32_t e0 = eq8(A0, S0), e1 = eq8(A1, S1);
			const uint32_t en = __shfl_down_sync(FULL_MASK, e0, 1);
			const uint32_t r30 = e0 & __funnelshift_r(e0, e1, 1) & __funnelshift_r(e0, e1, 2);
			const uint32_t r31 = e1 & __funnelshift_r(e1, en, 1) & __funnelshift_r(e1, en, 2);
			pushes += (r30 != 0) + (r31 != 0);
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				S0[j] = S1[j];
			}
		}
	}
	if (lane == 31) {
		pushes = 0;
	}
	return __reduce_add_sync(FULL_MASK, pushes);
}

/* One tile (1984 positions) with KD dense levels; chunk 0 is already staged. */
template <int CB, int HB, int KD>
__device__ __forceinline__ void st_tile(const X3SearchParams &prm, uint8_t *smem, unsigned long long p0, uint32_t &phase,
                                     uint8_t *deep_tile)
{
	using C = SCfg<CB, HB, KD>;
	uint4 *pw = reinterpret_cast<uint4 *>(smem + C::OFF_PW);
	uint32_t *hist = reinterpret_cast<uint32_t *>(smem + C::OFF_HIST);
	uint32_t *done_s = reinterpret_cast<uint32_t *>(smem + C::OFF_DONE);
	uint2 *q = reinterpret_cast<uint2 *>(smem + C::OFF_Q);
	const int lane = threadIdx.x;
	const int wA = 2 * lane;
	const uint32_t D = prm.D;
	const uint32_t MB = D / 32 + 1; /* distance blocks m = 0 .. D/32 */
	const uint32_t nchunks = (MB + C::MCH - 1) / C::MCH;

	LaneState<CB, KD> st;
#pragma unroll
	for (int k = 0; k < 2 * C::KD; ++k) {
		tree_clear(st.T[k]);
	}
	st.done = make_uint2(0, 0);
	st.uncond = D <= 65535u;
	st.q0 = smem_u32(q + lane);
	st.q1 = smem_u32(q + (C::QCAP - 1) * 32 + lane);
	for (int i = lane; i < 62 * 32; i += 32) {
		hist[i] = 0;
	}
	done_s[lane] = 0;
	if (lane + 32 < 62) {
		done_s[lane + 32] = 0;
	}
	__syncwarp();

	for (uint32_t c = 0; c < nchunks; ++c) {
		if (c > 0) {
			st_flush<CB, HB, KD>(st, q, hist, done_s, deep_tile, lane);
			st_stage<C>(prm, smem, p0, c, phase, lane);
		}
		if (c == 0) {
			load_word<C>(pw, wA, st.A0);
			load_word<C>(pw, wA + 1, st.A1);
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				st.A0[j] = ~st.A0[j];
				st.A1[j] = ~st.A1[j];
			}
		}

		for (int r = 0; r < 32; ++r) {
			/* valid chunk-relative blocks: d = 32 (MCH c + mm) + r in [1, D] */
			if (D < (uint32_t)r) {
				break;
			}
			const int vlo = (c == 0 && r == 0) ? 1 : 0;
			const long long hi = (long long)((D - (uint32_t)r) / 32) - (long long)C::MCH * c;
			const int vhi = hi >= C::MCH ? C::MCH - 1 : (int)hi;
			if (r > 0) {
				st_shift1<C>(pw, lane);
				__syncwarp();
			}
			if (vhi < vlo) {
				continue;
			}

			for (int g = 0; g < C::MCH / 16; ++g) {
				const int mm0 = 16 * g;
				if (mm0 > vhi || mm0 + 15 < vlo) {
					continue;
				}
				if (mm0 >= vlo && mm0 + 15 <= vhi) {
					st_group16<CB, HB, KD, false>(st, pw, q, hist, done_s, deep_tile, lane, wA, mm0, vlo, vhi);
				} else {
					st_group16<CB, HB, KD, true>(st, pw, q, hist, done_s, deep_tile, lane, wA, mm0, vlo, vhi);
				}
			}
		}
	}

	st_flush<CB, HB, KD>(st, q, hist, done_s, deep_tile, lane);
	__syncwarp();

	/* ---- epilogue: counts -> Lstar (and the 32-bin row) ---- */
	if (lane != 31) {
		const unsigned long long pbase = p0 + 64ull * lane;
		const bool want_rows = prm.H != nullptr;
		const bool aligned4 = (reinterpret_cast<uintptr_t>(prm.lstar) & 3u) == 0;
		uint32_t pack = 0;
#pragma unroll 1
		for (int jb = 0; jb < 64; ++jb) {
			const int j = jb >> 5, b = jb & 31;
			const unsigned long long p = pbase + jb;
			const bool live = p < prm.n; /* rows of padding positions are still handed back zeroed */
			const uint32_t pos = (uint32_t)lane * 64u + (uint32_t)jb;
			const uint32_t word = hist[pos];
			uint32_t dense[C::KD];
#pragma unroll
			for (int k = 0; k < C::KD; ++k) {
				dense[k] = j ? tree_value(st.T[C::KD + k], b) : tree_value(st.T[k], b);
			}
			uint32_t rw[C::ROWB / 4];
			const bool deep = (word >> 31) != 0;
			if (deep) {
				/* read the deep row and hand it back zeroed (the scratch invariant) */
				uint4 *r4 = reinterpret_cast<uint4 *>(deep_tile + (size_t)pos * C::ROWB);
#pragma unroll
				for (int v4 = 0; v4 < C::ROWB / 16; ++v4) {
					const uint4 t4 = __ldcg(r4 + v4);
					rw[4 * v4] = t4.x; rw[4 * v4 + 1] = t4.y; rw[4 * v4 + 2] = t4.z; rw[4 * v4 + 3] = t4.w;
					__stcg(r4 + v4, make_uint4(0, 0, 0, 0));
				}
			}
			constexpr int PER = 32 / C::DBITS;
			uint32_t ls;
			if (!want_rows) {
				/* Lstar only: count the levels whose (suffix-summed) count exceeds tc* without
				 * materialising the row (reference backend.c:76-78 collapsed) */
				const uint32_t c0 = dense[0];
				if (prm.t <= 0 || c0 < 2) {
					ls = 0;
				} else {
					const uint32_t tcs = min((uint32_t)prm.t, c0 - 1);
					uint32_t acc = 0;
					ls = 0;
					if (deep) {
#pragma unroll
						for (int L = 32; L >= C::L0 + C::NSH; --L) {
							const int k = L - C::L0 - C::NSH;
							acc += (rw[k / PER] >> (C::DBITS * (k % PER))) & C::DMASK;
							ls += acc > tcs;
						}
					}
#pragma unroll
					for (int L = C::L0 + C::NSH - 1; L >= C::L0; --L) {
						acc += (word >> (HB * (L - C::L0))) & C::FMASK;
						ls += acc > tcs;
					}
#pragma unroll
					for (int k = 0; k < C::KD; ++k) {
						ls += dense[k] > tcs;
					}
				}
			} else {
				uint32_t cnt[32];
				uint32_t acc = 0;
#pragma unroll
				for (int L = 32; L >= C::L0 + C::NSH; --L) {
					const int k = L - C::L0 - C::NSH;
					if (deep) {
						acc = min(acc + ((rw[k / PER] >> (C::DBITS * (k % PER))) & C::DMASK), C::CAP);
					}
					cnt[L - 1] = acc;
				}
#pragma unroll
				for (int L = C::L0 + C::NSH - 1; L >= C::L0; --L) {
					acc = min(acc + ((word >> (HB * (L - C::L0))) & C::FMASK), C::CAP);
					cnt[L - 1] = acc;
				}
#pragma unroll
				for (int k = 0; k < C::KD; ++k) {
					cnt[k] = dense[k];
				}
				ls = lstar_from_counts(cnt, prm.t);
				if (live) {
					store_row(prm.H, p, cnt);
				}
			}
			/* 4 positions per store where the whole group is live and the output is aligned */
			pack |= ls << (8 * (jb & 3));
			if ((jb & 3) == 3) {
				if (aligned4 && live) {
					*reinterpret_cast<uint32_t *>(prm.lstar + (p - 3)) = pack;
				} else {
#pragma unroll
					for (int k = 0; k < 4; ++k) {
						if (p - 3 + k < prm.n) {
							prm.lstar[p - 3 + k] = (uint8_t)(pack >> (8 * k));
						}
					}
				}
				pack = 0;
			}
		}
	}
}

/*
 * Probe kernel: samples the queue push rate of up to gridDim.x tiles spread over the input
 * (first 15 distance blocks at r = 0 each) into prm.tile_counter[1] (pushes) and [2] (units).
 * The two search kernels (KD = 2 and KD = 3) are launched back to back behind it; each reads the
 * sums and only the one that the rate selects does the work -- no host round trip.
 * A third dense level costs 4 ALU instructions per unit and removes the LCP == 3 events (40 % of
 * them on text): worth it when more than ~3 % of the units push an entry (measured break-even: C2-shaped text 3.9 %, C4-shaped binary 2 %).
 */
template <int CB, int HB>
__global__ void __launch_bounds__(32) x3_lcp_probe_kernel(X3SearchParams prm)
{
	using C = SCfg<CB, HB, 2>;
	extern __shared__ __align__(128) uint8_t smem[];
	uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
	const int lane = threadIdx.x;
	if (lane == 0) {
		mbar_init(bar, 1);
	}
	__syncwarp();
	uint32_t phase = 0;
	const unsigned int tile = (unsigned int)(((unsigned long long)blockIdx.x * prm.ntiles) / gridDim.x);
	st_stage<C>(prm, smem, (unsigned long long)tile * C::P, 0, phase, lane);
	const uint32_t pushes = st_probe<C>(reinterpret_cast<uint4 *>(smem + C::OFF_PW), lane);
	if (lane == 0) {
		atomicAdd(prm.tile_counter + 1, pushes);
		atomicAdd(prm.tile_counter + 2, 15u * 62u * X3_PROBE_R);
	}
}

__device__ __forceinline__ int st_choice(const X3SearchParams &prm)
{
	if (prm.kd == 2 || prm.kd == 3) {
		return prm.kd;
	}
	const unsigned int pushes = __ldcg(prm.tile_counter + 1), units = __ldcg(prm.tile_counter + 2);
	return pushes * 32u > units ? 3 : 2; /* more than 1 in 32 sampled units push an entry */
}

template <int CB, int HB, int KD>
__global__ void __launch_bounds__(32, 13) x3_lcp_stream_kernel(X3SearchParams prm)
{
	using C = SCfg<CB, HB, KD>;
	extern __shared__ __align__(128) uint8_t smem[];
	uint64_t *bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
	const int lane = threadIdx.x;
	uint8_t *deep_tile = prm.deep + (size_t)blockIdx.x * X3K_DEEP_BYTES_PER_CTA;

	if (st_choice(prm) != KD) {
		return; /* the other instantiation does this launch */
	}
	if (lane == 0) {
		mbar_init(bar, 1);
	}
	__syncwarp();
	uint32_t phase = 0;

	for (;;) {
		unsigned int tile = 0;
		if (lane == 0) {
			tile = atomicAdd(prm.tile_counter, 1u);
		}
		tile = __shfl_sync(FULL_MASK, tile, 0);
		if (tile >= prm.ntiles) {
			break;
		}
		const unsigned long long p0 = (unsigned long long)tile * C::P;
		st_stage<C>(prm, smem, p0, 0, phase, lane);
		st_tile<CB, HB, KD>(prm, smem, p0, phase, deep_tile);
		__syncwarp();
	}
}

int g_stream_ctas_per_sm = 0;
int g_stream_sms = 0;

} /* namespace */

template <int CB, int HB>
static cudaError_t stream_set_attr(void)
{
	cudaError_t e = cudaFuncSetAttribute(x3_lcp_stream_kernel<CB, HB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                                     (int)SCfg<CB, HB, 2>::SMEM);
	if (e != cudaSuccess) {
		return e;
	}
	e = cudaFuncSetAttribute(x3_lcp_stream_kernel<CB, HB, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                         (int)SCfg<CB, HB, 3>::SMEM);
	if (e != cudaSuccess) {
		return e;
	}
	return cudaFuncSetAttribute(x3_lcp_probe_kernel<CB, HB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                            (int)SCfg<CB, HB, 2>::SMEM);
}

cudaError_t x3k_stream_init_device(void)
{
	cudaError_t e = stream_set_attr<4, 6>();
	if (e != cudaSuccess) {
		return e;
	}
	e = stream_set_attr<8, 15>();
	if (e != cudaSuccess) {
		return e;
	}
	int dev = 0, occ = 0, occ3 = 0;
	e = cudaGetDevice(&dev);
	if (e != cudaSuccess) {
		return e;
	}
	e = cudaDeviceGetAttribute(&g_stream_sms, cudaDevAttrMultiProcessorCount, dev);
	if (e != cudaSuccess) {
		return e;
	}
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, x3_lcp_stream_kernel<4, 6, 2>, 32, SCfg<4, 6, 2>::SMEM);
	if (e != cudaSuccess) {
		return e;
	}
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, x3_lcp_stream_kernel<4, 6, 3>, 32, SCfg<4, 6, 3>::SMEM);
	if (e != cudaSuccess) {
		return e;
	}
	occ = occ > occ3 ? occ : occ3;
	g_stream_ctas_per_sm = occ > 0 ? occ : 1;
	return cudaSuccess;
}

int x3k_stream_max_grid(void)
{
	return g_stream_sms * g_stream_ctas_per_sm;
}

int x3k_stream_grid(unsigned long long n)
{
	const unsigned long long ntiles = (n + X3K_STREAM_TILE - 1) / X3K_STREAM_TILE;
	const unsigned long long cap = (unsigned long long)x3k_stream_max_grid();
	return (int)(ntiles < cap ? ntiles : cap);
}

template <int CB, int HB>
static cudaError_t stream_launch(const X3SearchParams &prm, int grid, cudaStream_t stream, int *launches)
{
	constexpr size_t SM = SCfg<CB, HB, 2>::SMEM;
	int n = 0;
	if (prm.kd != 2 && prm.kd != 3) {
		const int pgrid = prm.ntiles < 128u ? (int)prm.ntiles : 128;
		x3_lcp_probe_kernel<CB, HB><<<pgrid, 32, SM, stream>>>(prm);
		++n;
	}
	if (prm.kd != 3) {
		x3_lcp_stream_kernel<CB, HB, 2><<<grid, 32, SM, stream>>>(prm);
		++n;
	}
	if (prm.kd != 2) {
		x3_lcp_stream_kernel<CB, HB, 3><<<grid, 32, SM, stream>>>(prm);
		++n;
	}
	if (launches != nullptr) {
		*launches += n;
	}
	return cudaGetLastError();
}

/* full: u8-exact counters (any t <= 254, exact H rows); otherwise the t <= 15 fast path */
cudaError_t x3k_launch_stream(bool full, X3SearchParams prm, cudaStream_t stream, int *launches)
{
	prm.ntiles = (unsigned int)((prm.n + X3K_STREAM_TILE - 1) / X3K_STREAM_TILE);
	const int grid = x3k_stream_grid(prm.n);
	if (getenv("X3_TRACE") != nullptr) {
		fprintf(stderr, "x3k_launch_stream: %s path, kd %d, %u tiles, grid %d (%d CTAs/SM x %d SMs), %zu B shared memory per CTA\n",
		        full ? "full" : "fast", prm.kd, prm.ntiles, grid, g_stream_ctas_per_sm, g_stream_sms,
		        SCfg<4, 6, 2>::SMEM);
	}
	/* [0] tile scheduler, [1] probe pushes, [2] probe units */
	cudaError_t e = cudaMemsetAsync(prm.tile_counter, 0, 4 * sizeof(unsigned int), stream);
	if (e != cudaSuccess) {
		return e;
	}
	e = full ? stream_launch<8, 15>(prm, grid, stream, launches) : stream_launch<4, 6>(prm, grid, stream, launches);
	if (e == cudaSuccess && getenv("X3_TRACE") != nullptr) {
		unsigned int c[4] = {0, 0, 0, 0};
		cudaStreamSynchronize(stream);
		cudaMemcpy(c, prm.tile_counter, sizeof(c), cudaMemcpyDeviceToHost);
		fprintf(stderr, "x3k_launch_stream: probe %u pushes / %u units = %.2f %% -> %d dense levels\n", c[1], c[2],
		        c[2] ? 100.0 * c[1] / c[2] : 0.0, prm.kd == 2 || prm.kd == 3 ? prm.kd : (c[1] * 32u > c[2] ? 3 : 2));
	}
	return e;
}
