/* pageable host -> device upload rate with 1, 2, 4 host threads (separate streams, 16 MB pieces): does the driver's
 * staging of pageable copies scale with threads?  MEASUREMENT TOOL. */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

int main()
{
	const size_t n = (size_t)212 << 20, piece = (size_t)16 << 20;
	char *h = (char *)malloc(n);
	memset(h, 1, n);
	char *d;
	cudaMalloc(&d, n);
	cudaFree(0);
	for (int T : {1, 2, 4}) {
		for (int rep = 0; rep < 3; ++rep) {
			std::vector<cudaStream_t> st(T);
			for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
			const auto t0 = std::chrono::steady_clock::now();
			std::vector<std::thread> th;
			for (int k = 0; k < T; ++k) {
				th.emplace_back([&, k]() {
					cudaSetDevice(0);
					for (size_t o = (size_t)k * piece; o < n; o += (size_t)T * piece) {
						cudaMemcpyAsync(d + o, h + o, n - o < piece ? n - o : piece, cudaMemcpyHostToDevice, st[k]);
					}
					cudaStreamSynchronize(st[k]);
				});
			}
			for (auto &t : th) t.join();
			const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			if (rep == 2) printf("pageable H2D, %d thread(s): %.1f ms, %.1f GB/s\n", T, ms, n / ms / 1e6);
			for (auto &s : st) cudaStreamDestroy(s);
		}
	}
	/* and back */
	for (int T : {1, 2, 4}) {
		for (int rep = 0; rep < 3; ++rep) {
			std::vector<cudaStream_t> st(T);
			for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
			const auto t0 = std::chrono::steady_clock::now();
			std::vector<std::thread> th;
			for (int k = 0; k < T; ++k) {
				th.emplace_back([&, k]() {
					cudaSetDevice(0);
					for (size_t o = (size_t)k * piece; o < n; o += (size_t)T * piece) {
						cudaMemcpyAsync(h + o, d + o, n - o < piece ? n - o : piece, cudaMemcpyDeviceToHost, st[k]);
					}
					cudaStreamSynchronize(st[k]);
				});
			}
			for (auto &t : th) t.join();
			const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			if (rep == 2) printf("pageable D2H, %d thread(s): %.1f ms, %.1f GB/s\n", T, ms, n / ms / 1e6);
			for (auto &s : st) cudaStreamDestroy(s);
		}
	}
	return 0;
}
