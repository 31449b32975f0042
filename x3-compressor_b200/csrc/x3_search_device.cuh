/*
 * x3_search_device.cuh -- device-side helpers shared by the search kernels:
 * mbarrier / 1-D TMA bulk copy PTX wrappers, the threshold selection of
 * reference backend.c:76-78,92,99 and the 32-bin row store.
 */
#ifndef X3_SEARCH_DEVICE_CUH
#define X3_SEARCH_DEVICE_CUH

#include "x3_search_kernels.cuh"

#define FULL_MASK 0xffffffffu

/* ------------------------------------------------------------------------- */
/* PTX helpers: mbarrier + 1-D TMA bulk copy                                  */
/* ------------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
	             : "memory");
}

__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done;
	const uint32_t addr = smem_u32(bar);
	do {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done)
		    : "r"(addr), "r"(parity)
		    : "memory");
	} while (!done);
}

/* ------------------------------------------------------------------------- */
/* Shared epilogue: counts (1-based length L -> cnt[L-1]) to Lstar             */
/* ------------------------------------------------------------------------- */

/* reference backend.c:76-78,92,99 collapsed */
__device__ __forceinline__ uint32_t lstar_from_counts(const uint32_t (&cnt)[32], int t)
{
	const uint32_t c0 = cnt[0];
	if (t <= 0 || c0 < 2) {
		return 0;
	}
	const uint32_t tcs = min((uint32_t)t, c0 - 1);
	uint32_t ls = 0;
#pragma unroll
	for (int i = 0; i < 32; ++i) {
		ls += cnt[i] > tcs;
	}
	return ls;
}

__device__ __forceinline__ void store_row(uint8_t *H, unsigned long long p, const uint32_t (&cnt)[32])
{
	uint32_t w[8];
#pragma unroll
	for (int g = 0; g < 8; ++g) {
		w[g] = cnt[4 * g] | (cnt[4 * g + 1] << 8) | (cnt[4 * g + 2] << 16) | (cnt[4 * g + 3] << 24);
	}
	uint4 *row = reinterpret_cast<uint4 *>(H + p * 32);
	row[0] = make_uint4(w[0], w[1], w[2], w[3]);
	row[1] = make_uint4(w[4], w[5], w[6], w[7]);
}


#endif
