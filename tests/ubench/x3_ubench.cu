/*
 * x3_ubench.cu -- micro-benchmarks of the instruction mixes the search kernels are bound by, on the
 * box's own B200 (sm_100a).  TEST/MEASUREMENT INFRASTRUCTURE ("a measured per-GPU
 * peak for the instruction mix used ... record it next to MEASURED_PEAKS.json").  Prints one JSON
 * object; profiles/r2_ubench.json keeps a run.
 *
 *   alu_lop3      LOP3 chains, 32 warps/SM: warp instructions per cycle per SM (INT/ALU pipe peak)
 *   lds128        conflict-free LDS.128, 32 warps/SM: bytes per cycle per SM (shared-memory peak)
 *   pair_mix      the brute-force kernel's inner mix: 2 LDS.128 + 8 LOP3 per 32-bit match word over
 *                 8 bit planes (x3_search_stream.cu), i.e. 32 byte-pair tests per 8 LOP3
 *   match_any_k   __match_any_sync with k distinct values per warp: cycles per instruction per SM
 *   rank_round    one ranking round of the counting sorts (match + leader histogram update + shuffle)
 *   shfl, ballot  warp instructions per cycle per SM
 *   atoms_hist    shared-memory atomicAdd on a warp-private 256-bin histogram, random bytes
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

constexpr int THREADS = 1024;
constexpr int ITERS = 2048;

__global__ void __launch_bounds__(THREADS, 1) k_lop3(uint32_t *out, unsigned long long *cyc)
{
	uint32_t a = threadIdx.x, b = blockIdx.x, c = 0x9e3779b9u, d = 12345u;
	__syncthreads();
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k) { /* 4 independent chains, 32 LOP3 per iteration */
			a = (a & b) ^ c;
			b = (b | c) ^ d;
			c = (c & d) ^ a;
			d = (d | a) ^ b;
		}
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = a ^ b ^ c ^ d;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(THREADS, 1) k_lds128(uint32_t *out, unsigned long long *cyc)
{
	__shared__ uint4 buf[2048];
	for (int i = threadIdx.x; i < 2048; i += THREADS) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
	__syncthreads();
	uint32_t acc = 0;
	int idx = threadIdx.x;
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const uint4 v = buf[(idx + 32 * k) & 2047];
			acc ^= v.x ^ v.y ^ v.z ^ v.w;
		}
		idx += 256;
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

/* 8 bit planes of 32 positions per word: equality of 32 byte pairs = 8 XOR/OR-type LOP3 on two plane sets */
__global__ void __launch_bounds__(THREADS, 1) k_pairmix(uint32_t *out, unsigned long long *cyc)
{
	__shared__ uint4 planes[2048];
	for (int i = threadIdx.x; i < 2048; i += THREADS) planes[i] = make_uint4(i * 2654435761u, i ^ 0x5555u, i * 40503u, ~i);
	__syncthreads();
	uint32_t cnt = 0;
	int idx = threadIdx.x * 2;
	const uint4 p0 = planes[(threadIdx.x * 2) & 2047], p1 = planes[(threadIdx.x * 2 + 1) & 2047];
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS; ++i) {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint4 q0 = planes[(idx + 64 * k) & 2047], q1 = planes[(idx + 64 * k + 1) & 2047];
			uint32_t m = (p0.x ^ q0.x) | (p0.y ^ q0.y);   /* LOP3 x2 (3-input forms) */
			m |= (p0.z ^ q0.z) | (p0.w ^ q0.w);
			m |= (p1.x ^ q1.x) | (p1.y ^ q1.y);
			m |= (p1.z ^ q1.z) | (p1.w ^ q1.w);
			cnt += __popc(~m);
		}
		idx += 2;
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = cnt;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KDIST>
__global__ void __launch_bounds__(THREADS, 1) k_match(uint32_t *out, unsigned long long *cyc)
{
	const int lane = threadIdx.x & 31;
	uint32_t v = (uint32_t)(lane % KDIST) * 7u + 1u, acc = 0;
	__syncthreads();
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS / 4; ++i) {
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const uint32_t m = __match_any_sync(0xffffffffu, v);
			acc += m;
			v = (v & 255u) + (m & 0u); /* dependent chain through the result, value unchanged */
		}
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

/* independent matches (throughput rather than latency): 4 values per thread */
template <int KDIST>
__global__ void __launch_bounds__(THREADS, 1) k_match_tp(uint32_t *out, unsigned long long *cyc)
{
	const int lane = threadIdx.x & 31;
	uint32_t v0 = (uint32_t)(lane % KDIST), v1 = (uint32_t)((lane + 3) % KDIST), v2 = (uint32_t)((lane * 5) % KDIST),
	         v3 = (uint32_t)((lane * 7 + 1) % KDIST), acc = 0;
	__syncthreads();
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS / 4; ++i) {
#pragma unroll
		for (int k = 0; k < 2; ++k) {
			acc += __match_any_sync(0xffffffffu, v0);
			acc += __match_any_sync(0xffffffffu, v1);
			acc += __match_any_sync(0xffffffffu, v2);
			acc += __match_any_sync(0xffffffffu, v3);
		}
		v0 ^= acc & 0u;
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

/* the ranking round of x3_search_seg.cu: digit from shared memory, match, leader updates the warp's
 * histogram row, old count broadcast by shuffle */
__global__ void __launch_bounds__(THREADS, 1) k_rank_round(const uint8_t *bytes, uint32_t *out, unsigned long long *cyc)
{
	__shared__ uint16_t hist[32 * 256];
	__shared__ uint8_t xs[32768];
	for (int i = threadIdx.x; i < 32768; i += THREADS) xs[i] = bytes[i];
	for (int i = threadIdx.x; i < 32 * 256; i += THREADS) hist[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint16_t *myh = hist + warp * 256;
	const uint32_t lt = (1u << lane) - 1u;
	uint32_t acc = 0;
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int rep = 0; rep < ITERS / 32; ++rep) {
#pragma unroll 8
		for (int r = 0; r < 32; ++r) {
			const uint32_t d = xs[warp * 1024 + 32 * r + lane];
			const uint32_t peers = __match_any_sync(0xffffffffu, d);
			const int leader = __ffs(peers) - 1;
			uint32_t old = 0;
			if (lane == leader) {
				old = myh[d];
				myh[d] = (uint16_t)(old + __popc(peers));
			}
			old = __shfl_sync(0xffffffffu, old, leader);
			acc += old + __popc(peers & lt);
			__syncwarp();
		}
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

/* the same round with shared-memory atomics instead of the match (NOT stable inside a round: measured
 * only to know what the match costs) */
__global__ void __launch_bounds__(THREADS, 1) k_atoms_round(const uint8_t *bytes, uint32_t *out, unsigned long long *cyc)
{
	__shared__ uint32_t hist[32 * 256];
	__shared__ uint8_t xs[8192];
	for (int i = threadIdx.x; i < 8192; i += THREADS) xs[i] = bytes[i];
	for (int i = threadIdx.x; i < 32 * 256; i += THREADS) hist[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t *myh = hist + warp * 256;
	uint32_t acc = 0;
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int rep = 0; rep < ITERS / 32; ++rep) {
#pragma unroll 8
		for (int r = 0; r < 32; ++r) {
			const uint32_t d = xs[warp * 256 + ((32 * r + lane) & 255)];
			acc += atomicAdd(&myh[d], 1u);
		}
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(THREADS, 1) k_shfl(uint32_t *out, unsigned long long *cyc)
{
	uint32_t a = threadIdx.x, b = threadIdx.x * 3, c = 7, d = 9;
	const int lane = threadIdx.x & 31;
	__syncthreads();
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS; ++i) {
		a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
		b = __shfl_sync(0xffffffffu, b, (lane + 5) & 31);
		c = __shfl_sync(0xffffffffu, c, (lane + 9) & 31);
		d = __shfl_sync(0xffffffffu, d, (lane + 13) & 31);
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = a ^ b ^ c ^ d;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(THREADS, 1) k_ballot(uint32_t *out, unsigned long long *cyc)
{
	uint32_t a = threadIdx.x, b = threadIdx.x * 3, c = 7, d = 9;
	__syncthreads();
	const unsigned long long t0 = clock64();
#pragma unroll 1
	for (int i = 0; i < ITERS; ++i) {
		a += __ballot_sync(0xffffffffu, (a & 1u) != 0u);
		b += __ballot_sync(0xffffffffu, (b & 2u) != 0u);
		c += __ballot_sync(0xffffffffu, (c & 4u) != 0u);
		d += __ballot_sync(0xffffffffu, (d & 8u) != 0u);
	}
	const unsigned long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * THREADS + threadIdx.x] = a ^ b ^ c ^ d;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double mean_cycles(unsigned long long *d_cyc, int grid)
{
	unsigned long long *h = (unsigned long long *)malloc(grid * 8);
	CK(cudaMemcpy(h, d_cyc, grid * 8, cudaMemcpyDeviceToHost));
	double s = 0;
	for (int i = 0; i < grid; ++i) s += (double)h[i];
	free(h);
	return s / grid;
}

int main(void)
{
	int dev = 0, sms = 0, khz = 0;
	CK(cudaGetDevice(&dev));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
	const int grid = sms;
	uint32_t *d_out;
	unsigned long long *d_cyc;
	uint8_t *d_bytes, *h_bytes = (uint8_t *)malloc(32768);
	CK(cudaMalloc(&d_out, (size_t)grid * THREADS * 4));
	CK(cudaMalloc(&d_cyc, grid * 8));
	CK(cudaMalloc(&d_bytes, 32768));
	const int warps = THREADS / 32;
	printf("{\"device_sms\": %d, \"sm_clock_khz_max\": %d, \"warps_per_sm\": %d", sms, khz, warps);
#define RUN(name, launch, warp_ops_per_thread_loop)                                                  \
	do {                                                                                             \
		for (int w = 0; w < 2; ++w) { launch; }                                                      \
		CK(cudaDeviceSynchronize());                                                                 \
		cudaEvent_t e0, e1;                                                                          \
		CK(cudaEventCreate(&e0));                                                                    \
		CK(cudaEventCreate(&e1));                                                                    \
		CK(cudaEventRecord(e0));                                                                     \
		launch;                                                                                      \
		CK(cudaEventRecord(e1));                                                                     \
		CK(cudaDeviceSynchronize());                                                                 \
		float ms = 0;                                                                                \
		CK(cudaEventElapsedTime(&ms, e0, e1));                                                       \
		const double cyc = mean_cycles(d_cyc, grid);                                                 \
		const double ops = (double)(warp_ops_per_thread_loop) * warps;                               \
		printf(",\n \"%s\": {\"cycles\": %.0f, \"warp_instr_per_sm\": %.0f, \"warp_instr_per_cycle_per_sm\": %.4f, " \
		       "\"cycles_per_warp_instr_per_sm\": %.3f, \"ms\": %.4f}",                              \
		       name, cyc, ops, ops / cyc, cyc / ops, ms);                                            \
	} while (0)

	RUN("alu_lop3", (k_lop3<<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 32);
	RUN("lds128", (k_lds128<<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 8);
	RUN("pair_mix_words", (k_pairmix<<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 4);
	RUN("match_any_dep_k1", (k_match<1><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_dep_k4", (k_match<4><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_dep_k16", (k_match<16><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_dep_k32", (k_match<32><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_k1", (k_match_tp<1><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_k2", (k_match_tp<2><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_k4", (k_match_tp<4><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_k8", (k_match_tp<8><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_k16", (k_match_tp<16><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	RUN("match_any_k32", (k_match_tp<32><<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 2);
	for (int kind = 0; kind < 3; ++kind) {
		/* digits: 0 = text-like skew (16 frequent values), 1 = uniform random bytes, 2 = sorted runs (one value per round) */
		uint32_t s = 12345u;
		for (int i = 0; i < 32768; ++i) {
			s = s * 1664525u + 1013904223u;
			h_bytes[i] = kind == 0 ? (uint8_t)("etaoinshrdlucmfw"[(s >> 24) & 15]) : (kind == 1 ? (uint8_t)(s >> 24) : (uint8_t)(i >> 7));
		}
		CK(cudaMemcpy(d_bytes, h_bytes, 32768, cudaMemcpyHostToDevice));
		const char *nm[3] = {"rank_round_text16", "rank_round_random256", "rank_round_sorted"};
		const char *na[3] = {"atoms_round_text16", "atoms_round_random256", "atoms_round_sorted"};
		RUN(nm[kind], (k_rank_round<<<grid, THREADS>>>(d_bytes, d_out, d_cyc)), (double)ITERS);
		RUN(na[kind], (k_atoms_round<<<grid, THREADS>>>(d_bytes, d_out, d_cyc)), (double)ITERS);
	}
	RUN("shfl", (k_shfl<<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 4);
	RUN("ballot", (k_ballot<<<grid, THREADS>>>(d_out, d_cyc)), (double)ITERS * 4);
	printf("}\n");
	return 0;
}
