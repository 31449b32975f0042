/*
 * x3_search_api.cu -- C ABI of the B200 match search (include/x3_search.h).
 *
 * Host-side plumbing only: device discovery, cached device/staging buffers,
 * halo-sharded multi-GPU dispatch (contiguous position
 * ranges, trailing halo of W-2 bytes, no exchange step) and the copies either
 * side of the kernels in x3_search_kernels.cu.  No CPU implementation of the
 * search exists in this library: without a device every entry point fails.
 */
#include "x3_search.h"
#include "x3_search_kernels.cuh"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int X3S_MAX_PIECES = 8;

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

#define CU_TRY(expr)                                                                              \
	do {                                                                                          \
		cudaError_t e_ = (expr);                                                                  \
		if (e_ != cudaSuccess) {                                                                  \
			return fail(X3S_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),     \
			            __FILE__, __LINE__);                                                      \
		}                                                                                         \
	} while (0)

struct DevState {
	bool inited = false;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	uint8_t *d_x = nullptr;
	size_t cap_x = 0;
	uint8_t *d_l = nullptr;
	size_t cap_l = 0;
	uint8_t *d_h = nullptr;
	size_t cap_h = 0;
	/* pipelined shard (rank search, page-locked host buffers): the upload goes chunk by chunk on
	 * `stream` with an event behind every chunk, the chunks are searched on lane streams, and each
	 * chunk's Lstar goes back on its lane's stream */
	cudaStream_t ps[X3S_MAX_PIECES] = {};
	cudaEvent_t pev[X3S_MAX_PIECES][2] = {}; /* per lane: last search queued so far done, its Lstar back */
	std::vector<cudaEvent_t> upev;           /* per chunk: uploaded */
	bool pinited = false;
	bool follow_pageable = false; /* last search: a host thread brought the pieces home (pageable table) */
	int pieces = 1;  /* lanes of the last search on this device */
	int pfirst = 0;  /* chunk whose upload lets the first search start */
};

/* the host thread that follows a shard's pieces home (segment search, see x3s_search_host) */
struct FollowCtx {
	std::thread thr;
	std::atomic<size_t> queued{0}; /* pieces whose search (and, page-locked table, copy) is queued */
	std::atomic<int> failed{0};
};

/* what the rank search's per-chunk hooks need */
struct PieceCtx {
	DevState *ds;
	uint8_t *h_lstar;        /* host destination of the shard */
	unsigned long long CH;   /* chunk length */
	unsigned long long nch;
	size_t W;
};

/* the lane's stream waits for the upload that brings the last byte the chunk reads */
cudaError_t piece_before(void *vctx, int lane, unsigned long long a0, unsigned long long len)
{
	PieceCtx *c = (PieceCtx *)vctx;
	DevState &ds = *c->ds;
	const unsigned long long last = a0 + x3k_required_bytes((size_t)len, c->W) - 1;
	unsigned long long q = last / c->CH;
	if (q >= c->nch) {
		q = c->nch - 1;
	}
	if (a0 == 0) {
		ds.pfirst = (int)q;
	}
	return cudaStreamWaitEvent(ds.ps[lane], ds.upev[(size_t)q], 0);
}

/* the chunk's last kernel is queued: its Lstar goes back behind it */
cudaError_t piece_after(void *vctx, int lane, unsigned long long a0, unsigned long long len)
{
	PieceCtx *c = (PieceCtx *)vctx;
	DevState &ds = *c->ds;
	cudaError_t e;
	if ((e = cudaEventRecord(ds.pev[lane][0], ds.ps[lane])) != cudaSuccess) return e;
	if ((e = cudaMemcpyAsync(c->h_lstar + a0, ds.d_l + a0, (size_t)len, cudaMemcpyDeviceToHost, ds.ps[lane])) != cudaSuccess) return e;
	return cudaEventRecord(ds.pev[lane][1], ds.ps[lane]);
}

/* Per-device scratch of the stream kernel (tile counter + deep histogram rows).
 * One search is in flight per device at a time: launches on other streams are
 * ordered behind the previous one through `last`. */
struct Scratch {
	unsigned int *counter = nullptr;
	uint8_t *deep = nullptr;
	cudaEvent_t last = nullptr;
};
Scratch g_scratch[64];

std::mutex g_mu;
std::vector<DevState> g_dev;
std::vector<int> g_ids; /* device ordinals for x3s_search_host; empty = 0..n-1 */
bool g_kernel_inited[64] = {false};

int ensure_kernel_init(int device)
{
	if (device < 0 || device >= 64) {
		return fail(X3S_ERR_ARG, "device index %d out of range", device);
	}
	static std::mutex init_mu; /* x3s_search_device / _part take no other lock: two threads' first calls on one device */
	std::lock_guard<std::mutex> lock(init_mu);
	if (!g_kernel_inited[device]) {
		Scratch &sc = g_scratch[device];
		CU_TRY(cudaMalloc((void **)&sc.counter, 256));
		CU_TRY(cudaEventCreateWithFlags(&sc.last, cudaEventDisableTiming));
		g_kernel_inited[device] = true;
	}
	return X3S_OK;
}

/* the brute-force stream kernel's deep histogram rows: only when that kernel is going to run */
int ensure_stream_scratch(int device)
{
	Scratch &sc = g_scratch[device];
	if (sc.deep == nullptr) {
		/* opt-in shared memory sizes of the brute-force kernels (this also loads their modules, which
		 * a search that only runs the rank kernels never needs) */
		CU_TRY(x3k_init_device());
		CU_TRY(cudaMalloc((void **)&sc.deep, (size_t)x3k_stream_max_grid() * X3K_DEEP_BYTES_PER_CTA));
		CU_TRY(cudaMemset(sc.deep, 0, (size_t)x3k_stream_max_grid() * X3K_DEEP_BYTES_PER_CTA));
	}
	return X3S_OK;
}

/* fills the scratch fields and launches behind the device's previous search */
int launch_on(int device, int variant, X3SearchParams &prm, cudaStream_t stream, int *launches)
{
	Scratch &sc = g_scratch[device];
	const int kind = variant == X3S_KERNEL_DEFAULT ? x3k_default_kind(prm.D, prm.t, prm.H != nullptr)
	                 : (variant == X3S_KERNEL_RANK ? 1 : (variant == X3S_KERNEL_SEG ? 2 : 0));
	if (kind == 0 && variant != X3S_KERNEL_NAIVE) {
		const int rc = ensure_stream_scratch(device);
		if (rc != X3S_OK) {
			return rc;
		}
	}
	prm.tile_counter = sc.counter;
	prm.deep = sc.deep;
	prm.ntiles = 0;
	{
		const char *kd = getenv("X3_STREAM_KD"); /* tuning/testing knob; never changes results */
		prm.kd = kd != nullptr ? atoi(kd) : 0;
	}
	CU_TRY(cudaStreamWaitEvent(stream, sc.last, 0));
	CU_TRY(x3k_launch(variant, prm, stream, launches));
	CU_TRY(cudaEventRecord(sc.last, stream));
	return X3S_OK;
}

int grow(uint8_t **ptr, size_t *cap, size_t need)
{
	if (need <= *cap) {
		return X3S_OK;
	}
	if (*ptr != nullptr) {
		CU_TRY(cudaFree(*ptr));
		*ptr = nullptr;
		*cap = 0;
	}
	CU_TRY(cudaMalloc((void **)ptr, need));
	*cap = need;
	return X3S_OK;
}

int check_params(size_t W, int t, const void *H, int variant)
{
	/* the rank search (Lstar only) takes any t, like the reference (backend.c:21-26); the brute-force
	 * kernels count in u8 cells that saturate at 255, which is lossless only while t <= 254 */
	const size_t Dw = W > X3S_MAX_MATCH_LEN + 1 ? W - X3S_MAX_MATCH_LEN - 1 : 0;
	const bool brute = H != nullptr || (variant != X3S_KERNEL_DEFAULT && variant != X3S_KERNEL_RANK && variant != X3S_KERNEL_SEG) ||
	                   Dw > x3k_rank_max_distances();
	if (variant == X3S_KERNEL_SEG && (H != nullptr || Dw < 1 || Dw > x3k_seg_max_distances() || t < x3k_seg_min_t())) {
		return fail(X3S_ERR_UNSUPP, "the segment search takes Lstar only, 34 <= W <= %u and t >= %d", x3k_seg_max_distances() + 33u,
		            x3k_seg_min_t());
	}
	if (t > X3S_MAX_T && brute) {
		return fail(X3S_ERR_UNSUPP, "max match count %d exceeds %d for the 32-bin table / brute-force kernels (u8 cells)", t,
		            X3S_MAX_T);
	}
	if (W > ((size_t)1 << 31)) {
		return fail(X3S_ERR_UNSUPP, "forward window %zu exceeds 2^31", W);
	}
	return X3S_OK;
}

/* page-locked (cudaHostAlloc / cudaHostRegister) host memory?  Async copies of pageable memory
 * hold the calling thread until they are done, which would stall the thread that feeds the lanes. */
bool host_pinned(const void *p)
{
	cudaPointerAttributes at;
	if (p == nullptr || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		(void)cudaGetLastError();
		return false;
	}
	return at.type == cudaMemoryTypeHost;
}

uint32_t distances(size_t W)
{
	/* reference backend.c:66: s in [p+1, p+W-33] */
	return W > X3S_MAX_MATCH_LEN + 1 ? (uint32_t)(W - X3S_MAX_MATCH_LEN - 1) : 0u;
}

} /* namespace */

extern "C" {

int x3s_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		(void)cudaGetLastError();
		return 0;
	}
	return n;
}

const char *x3s_last_error(void)
{
	return g_err;
}

const char *x3s_version(void)
{
	return "x3-b200 search 0.4 (sm_100a; kernels: seg, rank, stream, bitsliced, naive)";
}

int x3s_default_kernel(size_t W, int t, int want_table)
{
	const int kind = x3k_default_kind(distances(W), t, want_table != 0);
	return kind == 2 ? X3S_KERNEL_SEG : (kind == 1 ? X3S_KERNEL_RANK : X3S_KERNEL_STREAM);
}

size_t x3s_required_bytes(size_t n_positions, size_t W)
{
	return x3k_required_bytes(n_positions, W);
}

/* ---- ONE input over several processes / GPUs, piece by piece in turn ------------------------------
 * part p of `parts` takes the pieces p, p + parts, ... of x3s_part_positions(W) positions each: contiguous
 * position ranges, each read together with the window behind it (backend.c:60-74 reads p .. p+W-2), dealt
 * out in turn so that every part gets its share of every region of the input (the cost of a position
 * varies 3x between the members of a mixed corpus: two contiguous halves of C5 take 7.9 and 4.4 ms). */
static size_t part_segments(size_t W)
{
	const uint32_t D = distances(W);
	if (D < 1 || D > x3k_seg_max_distances()) {
		return 0;
	}
	const size_t B = x3k_seg_positions(D);
	const char *pm = getenv("X3_PART_PIECE_KB"); /* tuning/testing knob; never changes results */
	const size_t kb = pm != nullptr && atoi(pm) >= 1 ? (size_t)atoi(pm) : 1024;
	const size_t segs = (kb << 10) / B;
	return segs < 1 ? 1 : segs;
}

size_t x3s_part_positions(size_t W)
{
	const uint32_t D = distances(W);
	return D >= 1 && D <= x3k_seg_max_distances() ? part_segments(W) * (size_t)x3k_seg_positions(D) : 0;
}

int x3s_search_device_part(int device, const void *d_x, size_t n_positions, size_t W, int t, void *d_lstar,
                           void *stream, int part, int parts)
{
	int rc = check_params(W, t, nullptr, X3S_KERNEL_SEG);
	if (rc != X3S_OK) {
		return rc;
	}
	if (d_x == nullptr || d_lstar == nullptr || ((uintptr_t)d_x & 15) != 0) {
		return fail(X3S_ERR_ARG, "null or misaligned device pointer");
	}
	if (parts < 1 || part < 0 || part >= parts) {
		return fail(X3S_ERR_ARG, "part %d of %d", part, parts);
	}
	CU_TRY(cudaSetDevice(device));
	rc = ensure_kernel_init(device);
	if (rc != X3S_OK) {
		return rc;
	}
	Scratch &sc = g_scratch[device];
	X3SearchParams prm;
	prm.x = (const uint8_t *)d_x;
	prm.n = n_positions;
	prm.D = distances(W);
	prm.t = t;
	prm.lstar = (uint8_t *)d_lstar;
	prm.H = nullptr;
	prm.tile_counter = sc.counter;
	prm.deep = nullptr;
	prm.ntiles = 0;
	prm.kd = 0;
	prm.part = (uint32_t)part;
	prm.parts = (uint32_t)parts;
	prm.piece_segments = (uint32_t)part_segments(W);
	cudaStream_t st = (cudaStream_t)stream;
	CU_TRY(cudaStreamWaitEvent(st, sc.last, 0));
	CU_TRY(x3k_launch_seg(prm, st, nullptr));
	CU_TRY(cudaEventRecord(sc.last, st));
	return X3S_OK;
}

int x3s_search_host_part(const void *x, size_t n, size_t W, int t, void *lstar, x3s_timing *timing, int part, int parts)
{
	const auto wall0 = std::chrono::steady_clock::now();
	int rc = check_params(W, t, nullptr, X3S_KERNEL_SEG);
	if (rc != X3S_OK) {
		return rc;
	}
	if (x == nullptr || (lstar == nullptr && n > 0)) {
		return fail(X3S_ERR_ARG, "null host pointer");
	}
	if (parts < 1 || part < 0 || part >= parts) {
		return fail(X3S_ERR_ARG, "part %d of %d", part, parts);
	}
	const int nvis = x3s_device_count();
	if (nvis <= 0) {
		return fail(X3S_ERR_CUDA, "no CUDA device visible (the search has no CPU fallback)");
	}
	std::lock_guard<std::mutex> lock(g_mu);
	const int dev = g_ids.empty() ? 0 : g_ids[0];
	if ((int)g_dev.size() < nvis) {
		g_dev.resize(nvis);
	}
	DevState &ds = g_dev[dev];
	CU_TRY(cudaSetDevice(dev));
	if ((rc = ensure_kernel_init(dev)) != X3S_OK) {
		return rc;
	}
	if (!ds.inited) {
		CU_TRY(cudaStreamCreateWithFlags(&ds.stream, cudaStreamNonBlocking));
		for (int i = 0; i < 4; ++i) {
			CU_TRY(cudaEventCreate(&ds.ev[i]));
		}
		ds.inited = true;
	}
	if (!ds.pinited) {
		for (int p = 0; p < X3S_MAX_PIECES; ++p) {
			CU_TRY(cudaStreamCreateWithFlags(&ds.ps[p], cudaStreamNonBlocking));
			for (int k = 0; k < 2; ++k) {
				CU_TRY(cudaEventCreate(&ds.pev[p][k]));
			}
		}
		ds.pinited = true;
	}
	/* the device holds the input's layout whole (this part's pieces and the windows behind them land at
	 * their own offsets; what other parts search stays untouched), zeroed behind the reference's padding */
	const size_t need = x3k_required_bytes(n, W), total = n + W;
	if (need > ds.cap_x) {
		if ((rc = grow(&ds.d_x, &ds.cap_x, need)) != X3S_OK) {
			return rc;
		}
	}
	if ((rc = grow(&ds.d_l, &ds.cap_l, n > 0 ? n : 1)) != X3S_OK) {
		return rc;
	}
	const size_t PS = x3s_part_positions(W);
	const size_t npieces = PS > 0 ? (n + PS - 1) / PS : 0;
	std::vector<size_t> mine;
	for (size_t q = (size_t)part; q < npieces; q += (size_t)parts) {
		mine.push_back(q);
	}
	x3s_timing tm;
	memset(&tm, 0, sizeof(tm));
	tm.gpus = 1;
	if (!mine.empty()) {
		while (ds.upev.size() < 3 * mine.size()) {
			cudaEvent_t ev = nullptr;
			CU_TRY(cudaEventCreate(&ev));
			ds.upev.push_back(ev);
		}
		Scratch &sc = g_scratch[dev];
		cudaStream_t U = ds.stream, K[2] = {ds.ps[0], ds.ps[1]}, Cc = ds.ps[2];
		const size_t m = mine.size();
		CU_TRY(cudaStreamWaitEvent(U, sc.last, 0));
		CU_TRY(cudaEventRecord(ds.ev[0], U));
		for (int k = 0; k < 2; ++k) {
			CU_TRY(cudaStreamWaitEvent(K[k], ds.ev[0], 0));
		}
		CU_TRY(cudaStreamWaitEvent(Cc, ds.ev[0], 0));
		if (need > total) {
			CU_TRY(cudaMemsetAsync(ds.d_x + total, 0, need - total, U));
		}
		const uint32_t D = distances(W);
		/* batches of pieces: a batch is uploaded piece by piece, searched by ONE launch (the kernel maps the
		 * launch's tickets to the batch's pieces) on one of two streams in turn -- so that the last segments of a
		 * batch share the GPU with the first of the next -- and copied back piece by piece */
		size_t per = 4; /* (measured on C5, 2 and 8 parts: 2 pieces per launch 7.7 / 2.3 ms, 4: 6.3 / 1.8, 8: 6.4 / 2.0, 26: 7.2 / 2.8) */
		const char *pb = getenv("X3_PART_BATCH"); /* tuning/testing knob; never changes results */
		if (pb != nullptr && atoi(pb) >= 1) per = (size_t)atoi(pb);
		size_t nb = 0;
		for (size_t j0 = 0; j0 < m; j0 += per, ++nb) {
			const size_t j1 = j0 + per < m ? j0 + per : m;
			for (size_t j = j0; j < j1; ++j) {
				const size_t p0 = mine[j] * PS, len = n - p0 < PS ? n - p0 : PS;
				size_t b1 = p0 + len + W + 128; /* the piece, the window behind it, the kernel's slack */
				if (b1 > total) b1 = total;
				CU_TRY(cudaMemcpyAsync(ds.d_x + p0, (const uint8_t *)x + p0, b1 - p0, cudaMemcpyHostToDevice, U));
			}
			CU_TRY(cudaEventRecord(ds.upev[nb], U));
			cudaStream_t Ks = K[nb & 1];
			CU_TRY(cudaStreamWaitEvent(Ks, ds.upev[nb], 0));
			X3SearchParams prm;
			prm.x = ds.d_x;
			prm.n = n;
			prm.D = D;
			prm.t = t;
			prm.lstar = ds.d_l;
			prm.H = nullptr;
			prm.tile_counter = sc.counter + 16 * (1 + (nb & 1));
			prm.deep = nullptr;
			prm.ntiles = 0;
			prm.kd = 0;
			prm.part = (uint32_t)part;
			prm.parts = (uint32_t)parts;
			prm.piece_segments = (uint32_t)part_segments(W);
			prm.piece_first = (uint32_t)j0;
			prm.piece_count = (uint32_t)(j1 - j0);
			CU_TRY(x3k_launch_seg(prm, Ks, &tm.launches));
			CU_TRY(cudaEventRecord(ds.upev[m + nb], Ks));
			CU_TRY(cudaStreamWaitEvent(Cc, ds.upev[m + nb], 0));
			for (size_t j = j0; j < j1; ++j) {
				const size_t p0 = mine[j] * PS, len = n - p0 < PS ? n - p0 : PS;
				CU_TRY(cudaMemcpyAsync((uint8_t *)lstar + p0, ds.d_l + p0, len, cudaMemcpyDeviceToHost, Cc));
			}
		}
		CU_TRY(cudaEventRecord(ds.ev[1], U));
		for (int k = 0; k < 2; ++k) {
			CU_TRY(cudaEventRecord(ds.pev[k][0], K[k]));
			CU_TRY(cudaStreamWaitEvent(U, ds.pev[k][0], 0));
		}
		CU_TRY(cudaEventRecord(ds.pev[2][1], Cc));
		CU_TRY(cudaStreamWaitEvent(U, ds.pev[2][1], 0));
		CU_TRY(cudaEventRecord(ds.ev[3], U));
		CU_TRY(cudaEventRecord(sc.last, U));
		CU_TRY(cudaStreamSynchronize(U));
		float up = 0.f, all = 0.f, kend = 0.f, ms = 0.f;
		CU_TRY(cudaEventElapsedTime(&up, ds.ev[0], ds.upev[0])); /* (the first batch uploaded) */
		CU_TRY(cudaEventElapsedTime(&all, ds.ev[0], ds.ev[3]));
		for (int k = 0; k < 2; ++k) {
			CU_TRY(cudaEventElapsedTime(&ms, ds.ev[0], ds.pev[k][0]));
			if (ms > kend) kend = ms;
		}
		tm.h2d_ms = up;
		tm.kernel_ms = kend - up;
		tm.d2h_ms = all - kend;
	}
	tm.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
	if (timing != nullptr) {
		*timing = tm;
	}
	return X3S_OK;
}

int x3s_search_device(int device, const void *d_x, size_t n_positions, size_t W, int t, void *d_lstar,
                      void *d_H, void *stream, int variant)
{
	int rc = check_params(W, t, d_H, variant);
	if (rc != X3S_OK) {
		return rc;
	}
	if (d_x == nullptr || d_lstar == nullptr) {
		return fail(X3S_ERR_ARG, "null device pointer");
	}
	if (((uintptr_t)d_x & 15) != 0) {
		return fail(X3S_ERR_ARG, "d_x must be 16-byte aligned");
	}
	if (d_H != nullptr && ((uintptr_t)d_H & 31) != 0) {
		return fail(X3S_ERR_ARG, "d_H must be 32-byte aligned");
	}
	CU_TRY(cudaSetDevice(device));
	rc = ensure_kernel_init(device);
	if (rc != X3S_OK) {
		return rc;
	}
	X3SearchParams prm;
	prm.x = (const uint8_t *)d_x;
	prm.n = n_positions;
	prm.D = distances(W);
	prm.t = t;
	prm.lstar = (uint8_t *)d_lstar;
	prm.H = (uint8_t *)d_H;
	return launch_on(device, variant, prm, (cudaStream_t)stream, nullptr);
}

} /* extern "C" */

namespace {

int search_host_impl(const void *x, size_t n, size_t W, int t, int ngpus, int variant, void *lstar, void *H,
                     x3s_timing *timing, volatile size_t *ready)
{
	const auto wall0 = std::chrono::steady_clock::now();
	if (ready != nullptr) {
		__atomic_store_n(ready, (size_t)0, __ATOMIC_RELEASE);
	}
	int rc = check_params(W, t, H, variant);
	if (rc != X3S_OK) {
		return rc;
	}
	if (x == nullptr || (lstar == nullptr && n > 0)) {
		return fail(X3S_ERR_ARG, "null host pointer");
	}
	const bool trace = getenv("X3_TRACE") != nullptr;
	auto lap = [&](const char *what) {
		if (trace) {
			fprintf(stderr, "x3s_search_host: %-28s +%.3f ms\n", what,
			        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
		}
	};
	const int nvis = x3s_device_count();
	lap("device count (cuInit)");
	if (nvis <= 0) {
		return fail(X3S_ERR_CUDA, "no CUDA device visible (the search has no CPU fallback)");
	}
	std::lock_guard<std::mutex> lock(g_mu);
	const int ndev = g_ids.empty() ? nvis : (int)g_ids.size();
	int G = ngpus <= 0 ? ndev : (ngpus < ndev ? ngpus : ndev);
	if ((size_t)G > n / 4096 + 1) {
		G = (int)(n / 4096 + 1); /* do not shard tiny inputs */
	}
	if ((int)g_dev.size() < nvis) {
		g_dev.resize(nvis);
	}

	/* contiguous position ranges [a_g, b_g), 16-byte aligned starts */
	std::vector<size_t> a(G + 1);
	for (int g = 0; g <= G; ++g) {
		size_t cut = (size_t)((unsigned __int128)n * g / G);
		cut &= ~(size_t)4095;
		a[g] = g == G ? n : cut;
	}

	const size_t total = n + W; /* bytes the caller guarantees behind x */
	/* a shard is pipelined piece by piece only between page-locked buffers (pageable ones take the plain
	 * upload - search - copy back sequence; the search itself still runs its lanes concurrently) */
	const bool pinned_io = n > 0 && host_pinned(x) && host_pinned(lstar);
	std::vector<int> shard_launches(G, 0);
	std::vector<int> shard_rc(G, X3S_OK);
	std::vector<std::string> shard_err(G);
	/* how far the table has landed: per shard, and as the prefix of the whole input a consumer that reads
	 * left to right (the reference's compress(), x3.c:379) may use */
	std::unique_ptr<FollowCtx[]> follow(new FollowCtx[G]);
	std::unique_ptr<std::atomic<size_t>[]> landed(new std::atomic<size_t>[G]);
	for (int g = 0; g < G; ++g) {
		landed[g].store(0);
	}
	std::mutex pub_mu;
	auto publish = [&](int g, size_t upto) {
		landed[g].store(upto, std::memory_order_release);
		if (ready == nullptr) {
			return;
		}
		std::lock_guard<std::mutex> lk(pub_mu);
		size_t w = 0;
		for (int k = 0; k < G; ++k) {
			const size_t pk = landed[k].load(std::memory_order_acquire);
			w = a[k] + pk;
			if (pk < a[k + 1] - a[k]) {
				break;
			}
		}
		if (w > *ready) {
			__atomic_store_n(ready, w, __ATOMIC_RELEASE);
		}
	};
	auto join_followers = [&]() {
		for (int g = 0; g < G; ++g) {
			if (follow[g].thr.joinable()) {
				follow[g].thr.join();
			}
		}
	};
	/* one shard: upload the slice with its trailing halo, search, bring Lstar back */
	auto shard = [&](int g) -> int {
		const int dev = g_ids.empty() ? g : g_ids[g];
		DevState &ds = g_dev[dev];
		const size_t np = a[g + 1] - a[g];
		if (np == 0) {
			return X3S_OK;
		}
		ds.follow_pageable = false;
		CU_TRY(cudaSetDevice(dev));
		int rc2 = ensure_kernel_init(dev);
		if (rc2 != X3S_OK) {
			return rc2;
		}
		if (!ds.inited) {
			CU_TRY(cudaStreamCreateWithFlags(&ds.stream, cudaStreamNonBlocking));
			for (int i = 0; i < 4; ++i) {
				CU_TRY(cudaEventCreate(&ds.ev[i]));
			}
			ds.inited = true;
		}
		const size_t need = x3k_required_bytes(np, W);
		if ((rc2 = grow(&ds.d_x, &ds.cap_x, need)) != X3S_OK) {
			return rc2;
		}
		if ((rc2 = grow(&ds.d_l, &ds.cap_l, np)) != X3S_OK) {
			return rc2;
		}
		if (H != nullptr && (rc2 = grow(&ds.d_h, &ds.cap_h, np * 32)) != X3S_OK) {
			return rc2;
		}
		/* slice + trailing halo: position p reads x[p .. p+W-2] (backend.c:66-74) */
		size_t have = np + W;
		if (a[g] + have > total) {
			have = total - a[g];
		}
		const uint32_t D = distances(W);
		const int kind = variant == X3S_KERNEL_DEFAULT ? x3k_default_kind(D, t, H != nullptr)
		                 : (variant == X3S_KERNEL_RANK ? 1 : (variant == X3S_KERNEL_SEG ? 2 : 0));
		const bool rank = H == nullptr && kind == 1;
		if (kind == 2 && getenv("X3_HOST_PIECES") == nullptr) {
			/* Segment search, piece by piece: the upload runs ahead on ds.stream, a piece (a whole number of
			 * segments, 16 MB; X3_SEG_PIECE_MB overrides) is
			 * searched on a second stream as soon as the bytes it reads -- its own and the window behind
			 * them -- have arrived, and its Lstar goes back on a third while the next pieces are searched.
			 * Pageable buffers (what the backend.h drop-in gets from the reference's main, x3.c:579) take the
			 * same route: an upload from pageable memory holds this thread only while the driver stages it,
			 * and the copies back to a pageable table -- which hold their caller until they are done -- are
			 * made by a second host thread, which also publishes how far the table has landed (`ready`). */
			const bool x_pinned = host_pinned(x), l_pinned = host_pinned(lstar);
			const size_t B = x3k_seg_positions(D);
			const char *pm = getenv("X3_SEG_PIECE_MB");
			const size_t mb = pm != nullptr && atoi(pm) >= 1 ? (size_t)atoi(pm) : 16; /* (C5, page-locked: 32 MB pieces 13.9 ms host to host, 16 MB 12.7, 8 MB 14.1) */
			size_t PS = ((mb << 20) / B) * B;
			if (PS < B) PS = B;
			const size_t nps = (np + PS - 1) / PS;
			const bool follower = !l_pinned || ready != nullptr; /* a host thread follows the pieces home */
			if (nps >= 2 || (follower && nps >= 1)) {
				if (!ds.pinited) {
					for (int p = 0; p < X3S_MAX_PIECES; ++p) {
						CU_TRY(cudaStreamCreateWithFlags(&ds.ps[p], cudaStreamNonBlocking));
						for (int k = 0; k < 2; ++k) {
							CU_TRY(cudaEventCreate(&ds.pev[p][k]));
						}
					}
					ds.pinited = true;
				}
				while (ds.upev.size() < 3 * nps) {
					cudaEvent_t ev = nullptr;
					CU_TRY(cudaEventCreate(&ev));
					ds.upev.push_back(ev);
				}
				/* searches on two streams in turn (a counter each): the last segments of a piece share the GPU with the
				 * first of the next */
				cudaStream_t U = ds.stream, K2[2] = {ds.ps[0], ds.ps[2]}, Cc = ds.ps[1];
				Scratch &sc = g_scratch[dev];
				CU_TRY(cudaStreamWaitEvent(U, sc.last, 0));
				CU_TRY(cudaEventRecord(ds.ev[0], U));
				CU_TRY(cudaStreamWaitEvent(K2[0], ds.ev[0], 0));
				CU_TRY(cudaStreamWaitEvent(K2[1], ds.ev[0], 0));
				CU_TRY(cudaStreamWaitEvent(Cc, ds.ev[0], 0));
				/* the follower: piece q is taken home (pageable table) or waited for (page-locked table: the copy
				 * is queued by this thread) once this thread has queued its search */
				FollowCtx &fc = follow[g];
				fc.queued.store(0);
				fc.failed.store(0);
				if (follower) {
					uint8_t *dst0 = (uint8_t *)lstar + a[g];
					fc.thr = std::thread([&, dev, nps, PS, np, dst0, l_pinned, Cc, g]() {
						if (cudaSetDevice(dev) != cudaSuccess) {
							fc.failed.store(1);
						}
						for (size_t q = 0; q < nps && fc.failed.load() == 0; ++q) {
							while (fc.queued.load(std::memory_order_acquire) <= q && fc.failed.load() == 0) {
								std::this_thread::yield();
							}
							if (fc.failed.load() != 0) {
								break;
							}
							const size_t p0 = q * PS, len = np - p0 < PS ? np - p0 : PS;
							cudaError_t e;
							if (l_pinned) {
								e = cudaEventSynchronize(ds.upev[2 * nps + q]);
							} else {
								e = cudaEventSynchronize(ds.upev[nps + q]);
								if (e == cudaSuccess) e = cudaMemcpyAsync(dst0 + p0, ds.d_l + p0, len, cudaMemcpyDeviceToHost, Cc);
								if (e == cudaSuccess) e = cudaStreamSynchronize(Cc);
							}
							if (e != cudaSuccess) {
								fc.failed.store(1);
								break;
							}
							publish(g, p0 + len);
						}
						if (!l_pinned && fc.failed.load() == 0) {
							(void)cudaEventRecord(ds.pev[0][1], Cc);
							(void)cudaEventSynchronize(ds.pev[0][1]);
						}
					});
				}
				auto bail = [&](cudaError_t e) -> int { /* never leave the follower spinning */
					fc.failed.store(1);
					return fail(X3S_ERR_CUDA, "segment pieces: %s", cudaGetErrorString(e));
				};
				cudaError_t ce = cudaSuccess;
				size_t up = 0; /* pieces uploaded so far */
				auto upload = [&](size_t q) -> cudaError_t {
					const size_t b0 = q * PS, b1 = q + 1 == nps ? have : (q + 1) * PS;
					cudaError_t e = cudaMemcpyAsync(ds.d_x + b0, (const uint8_t *)x + a[g] + b0, b1 - b0, cudaMemcpyHostToDevice, U);
					if (e == cudaSuccess && q + 1 == nps && need > have) {
						e = cudaMemsetAsync(ds.d_x + have, 0, need - have, U);
					}
					if (e == cudaSuccess) e = cudaEventRecord(ds.upev[q], U);
					return e;
				};
				for (size_t q = 0; q < nps; ++q) {
					const size_t p0 = q * PS, len = np - p0 < PS ? np - p0 : PS;
					size_t last = (p0 + len + W + 64) / PS; /* the piece that brings the last byte this one reads */
					if (last >= nps) last = nps - 1;
					if (q == 0) ds.pfirst = (int)last;
					/* uploads interleaved with the launches: from pageable memory each one holds this thread */
					while (up <= last && ce == cudaSuccess) {
						ce = upload(up++);
					}
					if (ce != cudaSuccess) return bail(ce);
					cudaStream_t K = K2[q & 1];
					if ((ce = cudaStreamWaitEvent(K, ds.upev[last], 0)) != cudaSuccess) return bail(ce);
					X3SearchParams prm;
					prm.x = ds.d_x + p0;
					prm.n = len;
					prm.D = D;
					prm.t = t;
					prm.lstar = ds.d_l + p0;
					prm.H = nullptr;
					prm.tile_counter = sc.counter + 16 * (1 + (q & 1));
					prm.deep = nullptr;
					prm.ntiles = 0;
					prm.kd = 0;
					if ((ce = x3k_launch_seg(prm, K, &shard_launches[g])) != cudaSuccess) return bail(ce);
					if ((ce = cudaEventRecord(ds.upev[nps + q], K)) != cudaSuccess) return bail(ce);
					if (l_pinned) {
						if ((ce = cudaStreamWaitEvent(Cc, ds.upev[nps + q], 0)) != cudaSuccess) return bail(ce);
						if ((ce = cudaMemcpyAsync((uint8_t *)lstar + a[g] + p0, ds.d_l + p0, len, cudaMemcpyDeviceToHost, Cc)) != cudaSuccess) return bail(ce);
						if ((ce = cudaEventRecord(ds.upev[2 * nps + q], Cc)) != cudaSuccess) return bail(ce);
					}
					fc.queued.store(q + 1, std::memory_order_release);
				}
				if ((ce = cudaEventRecord(ds.ev[1], U)) != cudaSuccess) return bail(ce);
				if ((ce = cudaEventRecord(ds.pev[0][0], K2[0])) != cudaSuccess) return bail(ce);
				if ((ce = cudaEventRecord(ds.pev[1][0], K2[1])) != cudaSuccess) return bail(ce);
				if (l_pinned) {
					if ((ce = cudaEventRecord(ds.pev[0][1], Cc)) != cudaSuccess) return bail(ce);
				}
				if ((ce = cudaStreamWaitEvent(U, ds.pev[0][0], 0)) != cudaSuccess) return bail(ce);
				if ((ce = cudaStreamWaitEvent(U, ds.pev[1][0], 0)) != cudaSuccess) return bail(ce);
				if (l_pinned) {
					if ((ce = cudaStreamWaitEvent(U, ds.pev[0][1], 0)) != cudaSuccess) return bail(ce);
				}
				if ((ce = cudaEventRecord(ds.ev[3], U)) != cudaSuccess) return bail(ce);
				if ((ce = cudaEventRecord(sc.last, U)) != cudaSuccess) return bail(ce);
				ds.pieces = 2;
				ds.follow_pageable = !l_pinned;
				return X3S_OK;
			}
		}
		int P = 1;
		if (rank && getenv("X3_RANK_PROFILE") == nullptr && pinned_io) {
			const char *hp = getenv("X3_HOST_PIECES"); /* tuning/testing knob; never changes results */
			P = hp != nullptr && atoi(hp) >= 1 ? atoi(hp) : x3k_rank_default_lanes(np, D);
			if (P > X3S_MAX_PIECES) P = X3S_MAX_PIECES;
			if (P > x3k_rank_max_lanes()) P = x3k_rank_max_lanes();
			while (P > 1 && np / (size_t)P < 65536) --P;
		}
		ds.pieces = P;
		if (P > 1) {
			/* Pipelined shard: the upload goes chunk by chunk on ds.stream, a chunk is searched on the next
			 * free lane as soon as the bytes it reads (its positions and the window behind them) have
			 * arrived, and its Lstar goes back while the chunks behind it are still being searched. */
			if (!ds.pinited) {
				for (int p = 0; p < X3S_MAX_PIECES; ++p) {
					CU_TRY(cudaStreamCreateWithFlags(&ds.ps[p], cudaStreamNonBlocking));
					for (int k = 0; k < 2; ++k) {
						CU_TRY(cudaEventCreate(&ds.pev[p][k]));
					}
				}
				ds.pinited = true;
			}
			PieceCtx ctx;
			ctx.ds = &ds;
			ctx.h_lstar = (uint8_t *)lstar + a[g];
			ctx.W = W;
			x3k_rank_chunking(np, D, P, &ctx.CH, &ctx.nch);
			while (ds.upev.size() < ctx.nch) {
				cudaEvent_t ev = nullptr;
				CU_TRY(cudaEventCreate(&ev));
				ds.upev.push_back(ev);
			}
			Scratch &sc = g_scratch[dev];
			CU_TRY(cudaStreamWaitEvent(ds.stream, sc.last, 0));
			CU_TRY(cudaEventRecord(ds.ev[0], ds.stream));
			for (unsigned long long q = 0; q < ctx.nch; ++q) {
				/* bytes of chunk q; the last upload brings the halo and is followed by the zeroed slack */
				const size_t b0 = (size_t)(q * ctx.CH), b1 = q + 1 == ctx.nch ? have : (size_t)((q + 1) * ctx.CH);
				CU_TRY(cudaMemcpyAsync(ds.d_x + b0, (const uint8_t *)x + a[g] + b0, b1 - b0, cudaMemcpyHostToDevice, ds.stream));
				if (q + 1 == ctx.nch && need > have) {
					CU_TRY(cudaMemsetAsync(ds.d_x + have, 0, need - have, ds.stream));
				}
				CU_TRY(cudaEventRecord(ds.upev[(size_t)q], ds.stream));
			}
			CU_TRY(cudaEventRecord(ds.ev[1], ds.stream));
			X3RankBatch b;
			b.x = ds.d_x;
			b.lstar = ds.d_l;
			b.n = np;
			b.lanes = P;
			for (int p = 0; p < P; ++p) {
				b.streams[p] = ds.ps[p];
				/* lanes that get no chunk (or whose hooks never run) still have recorded events to wait on and time */
				CU_TRY(cudaStreamWaitEvent(ds.ps[p], ds.ev[0], 0));
				CU_TRY(cudaEventRecord(ds.pev[p][0], ds.ps[p]));
				CU_TRY(cudaEventRecord(ds.pev[p][1], ds.ps[p]));
			}
			b.before_chunk = piece_before;
			b.after_chunk = piece_after;
			b.ctx = &ctx;
			CU_TRY(x3k_launch_rank_batch(b, D, t, &shard_launches[g]));
			for (int p = 0; p < P; ++p) {
				/* ds.stream (and with it the device's next search) continues behind every lane */
				CU_TRY(cudaStreamWaitEvent(ds.stream, ds.pev[p][1], 0));
			}
			CU_TRY(cudaEventRecord(ds.ev[3], ds.stream));
			CU_TRY(cudaEventRecord(sc.last, ds.stream));
			return X3S_OK;
		}
		CU_TRY(cudaEventRecord(ds.ev[0], ds.stream));
		CU_TRY(cudaMemcpyAsync(ds.d_x, (const uint8_t *)x + a[g], have, cudaMemcpyHostToDevice, ds.stream));
		if (need > have) {
			CU_TRY(cudaMemsetAsync(ds.d_x + have, 0, need - have, ds.stream));
		}
		CU_TRY(cudaEventRecord(ds.ev[1], ds.stream));
		X3SearchParams prm;
		prm.x = ds.d_x;
		prm.n = np;
		prm.D = D;
		prm.t = t;
		prm.lstar = ds.d_l;
		prm.H = H != nullptr ? ds.d_h : nullptr;
		if ((rc2 = launch_on(dev, variant, prm, ds.stream, &shard_launches[g])) != X3S_OK) {
			return rc2;
		}
		CU_TRY(cudaEventRecord(ds.ev[2], ds.stream));
		CU_TRY(cudaMemcpyAsync((uint8_t *)lstar + a[g], ds.d_l, np, cudaMemcpyDeviceToHost, ds.stream));
		if (H != nullptr) {
			CU_TRY(cudaMemcpyAsync((uint8_t *)H + a[g] * 32, ds.d_h, np * 32, cudaMemcpyDeviceToHost,
			                       ds.stream));
		}
		CU_TRY(cudaEventRecord(ds.ev[3], ds.stream));
		return X3S_OK;
	};
	if (G == 1) {
		rc = shard(0);
		lap("shard queued");
		if (rc != X3S_OK) {
			follow[0].failed.store(1);
			join_followers();
			return rc;
		}
	} else {
		/* one submitting thread per GPU: the rank search waits on its own read-backs while it
		 * queues levels, which must not hold up the other GPUs */
		std::vector<std::thread> thr;
		for (int g = 0; g < G; ++g) {
			thr.emplace_back([&, g]() {
				shard_rc[g] = shard(g);
				if (shard_rc[g] != X3S_OK) {
					shard_err[g] = g_err; /* thread-local message of the worker */
				}
			});
		}
		for (auto &th : thr) {
			th.join();
		}
		lap("shards queued");
		for (int g = 0; g < G; ++g) {
			if (shard_rc[g] != X3S_OK) {
				for (int k = 0; k < G; ++k) {
					follow[k].failed.store(1);
				}
				join_followers();
				return fail(shard_rc[g], "GPU shard %d: %s", g, shard_err[g].c_str());
			}
		}
	}
	join_followers();
	for (int g = 0; g < G; ++g) {
		if (follow[g].failed.load() != 0) {
			return fail(X3S_ERR_CUDA, "GPU shard %d: bringing the table home failed: %s", g, cudaGetErrorString(cudaGetLastError()));
		}
	}
	lap("pieces home");
	int launches = 0;
	for (int g = 0; g < G; ++g) {
		launches += shard_launches[g];
	}

	x3s_timing tm;
	memset(&tm, 0, sizeof(tm));
	tm.gpus = G;
	tm.launches = launches;
	for (int g = 0; g < G; ++g) {
		if (a[g + 1] == a[g]) {
			continue;
		}
		const int dev = g_ids.empty() ? g : g_ids[g];
		DevState &ds = g_dev[dev];
		CU_TRY(cudaSetDevice(dev));
		CU_TRY(cudaStreamSynchronize(ds.stream));
		float ms = 0.f;
		if (ds.pieces > 1) {
			/* pipelined shard: the three figures split the critical path -- upload until the first search
			 * can start, from there to the end of the last search, from there to the last byte back */
			float up = 0.f, all = 0.f, kend = 0.f;
			CU_TRY(cudaEventElapsedTime(&up, ds.ev[0], ds.upev[(size_t)ds.pfirst]));
			CU_TRY(cudaEventElapsedTime(&all, ds.ev[0], ds.follow_pageable ? ds.pev[0][1] : ds.ev[3]));
			for (int p = 0; p < ds.pieces; ++p) {
				CU_TRY(cudaEventElapsedTime(&ms, ds.ev[0], ds.pev[p][0]));
				if (ms > kend) kend = ms;
			}
			if (up > tm.h2d_ms) tm.h2d_ms = up;
			if (kend - up > tm.kernel_ms) tm.kernel_ms = kend - up;
			if (all - kend > tm.d2h_ms) tm.d2h_ms = all - kend;
			continue;
		}
		CU_TRY(cudaEventElapsedTime(&ms, ds.ev[0], ds.ev[1]));
		if (ms > tm.h2d_ms) tm.h2d_ms = ms;
		CU_TRY(cudaEventElapsedTime(&ms, ds.ev[1], ds.ev[2]));
		if (ms > tm.kernel_ms) tm.kernel_ms = ms;
		CU_TRY(cudaEventElapsedTime(&ms, ds.ev[2], ds.ev[3]));
		if (ms > tm.d2h_ms) tm.d2h_ms = ms;
	}
	lap("done");
	if (ready != nullptr) {
		__atomic_store_n(ready, n, __ATOMIC_RELEASE);
	}
	tm.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
	if (timing != nullptr) {
		*timing = tm;
	}
	return X3S_OK;
}

} /* namespace */

extern "C" {

int x3s_search_host(const void *x, size_t n, size_t W, int t, int ngpus, int variant, void *lstar, void *H,
                    x3s_timing *timing)
{
	return search_host_impl(x, n, W, t, ngpus, variant, lstar, H, timing, nullptr);
}

int x3s_search_host_stream(const void *x, size_t n, size_t W, int t, int ngpus, int variant, void *lstar,
                           x3s_timing *timing, volatile size_t *ready)
{
	if (ready == nullptr) {
		return fail(X3S_ERR_ARG, "x3s_search_host_stream: null progress word");
	}
	return search_host_impl(x, n, W, t, ngpus, variant, lstar, nullptr, timing, ready);
}

int x3s_set_devices(const int *ids, int count)
{
	std::lock_guard<std::mutex> lock(g_mu);
	const int nvis = x3s_device_count();
	if (count <= 0 || ids == nullptr) {
		g_ids.clear();
		return X3S_OK;
	}
	for (int i = 0; i < count; ++i) {
		if (ids[i] < 0 || ids[i] >= nvis) {
			return fail(X3S_ERR_ARG, "device ordinal %d not visible (%d devices)", ids[i], nvis);
		}
	}
	g_ids.assign(ids, ids + count);
	return X3S_OK;
}

int x3s_rank_profile(int device, int kind, double *ms, double *elements, int *launches)
{
	if (ms == nullptr || elements == nullptr || launches == nullptr ||
	    x3k_rank_profile(device, kind, ms, elements, launches) != 0) {
		return fail(X3S_ERR_ARG, "x3s_rank_profile: bad device, kind or null pointer");
	}
	return X3S_OK;
}

int x3s_rank_plan(size_t n_positions, size_t W, int lanes, size_t *chunk_positions, size_t *chunks, int *lanes_used)
{
	if (chunk_positions == nullptr || chunks == nullptr || lanes_used == nullptr || lanes < 0) {
		return fail(X3S_ERR_ARG, "x3s_rank_plan: null pointer or negative lane count");
	}
	const uint32_t D = distances(W);
	if (W > ((size_t)1 << 31) || D > x3k_rank_max_distances()) {
		return fail(X3S_ERR_UNSUPP, "forward window %zu is beyond the rank search (W - 33 <= 2^23)", W);
	}
	int L = lanes == 0 ? x3k_rank_default_lanes(n_positions, D) : lanes;
	if (L > x3k_rank_max_lanes()) {
		L = x3k_rank_max_lanes();
	}
	unsigned long long ch = 0, cnt = 0;
	x3k_rank_chunking(n_positions, D, L, &ch, &cnt);
	*chunk_positions = (size_t)ch;
	*chunks = (size_t)cnt;
	*lanes_used = (unsigned long long)L < cnt ? L : (int)(cnt > 0 ? cnt : 1);
	return X3S_OK;
}

void *x3s_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) {
		(void)cudaGetLastError();
		fail(X3S_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes);
		return nullptr;
	}
	return p;
}

void x3s_host_free(void *p)
{
	if (p != nullptr) {
		cudaFreeHost(p);
	}
}

int x3s_host_register(void *p, size_t bytes)
{
	if (p == nullptr || bytes == 0) {
		return fail(X3S_ERR_ARG, "x3s_host_register: null pointer or empty range");
	}
	/* whole pages: the driver locks pages, and callers hand in slices of larger buffers */
	const uintptr_t page = 4096, lo = (uintptr_t)p & ~(page - 1), hi = ((uintptr_t)p + bytes + page - 1) & ~(page - 1);
	const cudaError_t e = cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterPortable);
	if (e != cudaSuccess) {
		(void)cudaGetLastError();
		return fail(X3S_ERR_CUDA, "cudaHostRegister(%zu bytes) failed: %s", (size_t)(hi - lo), cudaGetErrorString(e));
	}
	return X3S_OK;
}

int x3s_host_unregister(void *p)
{
	if (p == nullptr) {
		return fail(X3S_ERR_ARG, "x3s_host_unregister: null pointer");
	}
	const cudaError_t e = cudaHostUnregister((void *)((uintptr_t)p & ~(uintptr_t)4095));
	if (e != cudaSuccess) {
		(void)cudaGetLastError();
		return fail(X3S_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
	}
	return X3S_OK;
}

void x3s_release(void)
{
	std::lock_guard<std::mutex> lock(g_mu);
	for (size_t g = 0; g < g_dev.size(); ++g) {
		DevState &ds = g_dev[g];
		if (!ds.inited && ds.d_x == nullptr) {
			continue;
		}
		if (cudaSetDevice((int)g) != cudaSuccess) {
			continue;
		}
		cudaFree(ds.d_x);
		cudaFree(ds.d_l);
		cudaFree(ds.d_h);
		if (ds.inited) {
			for (int i = 0; i < 4; ++i) {
				cudaEventDestroy(ds.ev[i]);
			}
			cudaStreamDestroy(ds.stream);
		}
		if (ds.pinited) {
			for (int p = 0; p < X3S_MAX_PIECES; ++p) {
				cudaStreamDestroy(ds.ps[p]);
				for (int k = 0; k < 2; ++k) {
					cudaEventDestroy(ds.pev[p][k]);
				}
			}
		}
		for (cudaEvent_t ev : ds.upev) {
			cudaEventDestroy(ev);
		}
		ds = DevState();
	}
	/* scratch of every device a search ran on, also those only x3s_search_device() touched */
	for (int g = 0; g < 64; ++g) {
		if (!g_kernel_inited[g] || cudaSetDevice(g) != cudaSuccess) {
			continue;
		}
		cudaFree(g_scratch[g].counter);
		cudaFree(g_scratch[g].deep);
		cudaEventDestroy(g_scratch[g].last);
		x3k_rank_release(g);
		g_scratch[g] = Scratch();
		g_kernel_inited[g] = false;
	}
}

} /* extern "C" */
