// Micro-benchmark of one radix pass of the packed suffix sorter (suffix_sort.cuh) on random
// words: build with -D switches to compare kernel variants.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -I phylonium_b200/csrc tools/sortbench.cu -o /tmp/sortbench
#include "suffix_sort.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace phy;

int main(int argc, char **argv)
{
	const int64_t m = argc > 1 ? atoll(argv[1]) : 10000001;
	const int reps = argc > 2 ? atoi(argv[2]) : 20;
	cudaStream_t s;
	cudaStreamCreate(&s);
	std::vector<uint64_t> h(m);
	std::mt19937_64 rng(1);
	for (int64_t i = 0; i < m; i++)
		h[i] = (rng() & 0xffffffff00000000ull) | (uint64_t)i;
	DevBuf<uint64_t> a(pk_padded_words(m), s), b(pk_padded_words(m), s);
	cudaMemset(a.get(), 0xff, pk_padded_words(m) * 8);
	DevBuf<uint8_t> flush(512 << 20, s);
	cudaMemcpy(a.get(), h.data(), m * 8, cudaMemcpyHostToDevice);
	const int ntiles = div_up(m, PK_TILE);
	DevBuf<uint32_t> counts((size_t)ntiles * RS_BINS, s), totals(RS_BINS, s);
	const int grid = pk_scatter_grid<false>(ntiles);
	printf("grid %d blocks (%d per SM), %zu B shared\n", grid, grid / 148, sizeof(PkSmem<false>));
	const PkMasks mk = pk_masks(16);
	cudaEvent_t e0, e1, e2, e3;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventCreate(&e2);
	cudaEventCreate(&e3);
	float th = 0, ts = 0, tc = 0;
	for (int r = 0; r < reps + 2; r++) {
		flush.zero();
		const int shift = 32 + 8 * (r % 4);
		cudaEventRecord(e0, s);
		pk_histogram<false><<<ntiles, PK_THREADS, 0, s>>>(a.get(), nullptr, 0, m, shift, mk, counts.get());
		cudaEventRecord(e1, s);
		pk_scan_counts<<<RS_BINS / 8, PK_SCAN_THREADS, 0, s>>>(counts.get(), ntiles, totals.get());
		cudaEventRecord(e2, s);
		pk_scatter<false><<<grid, PKS_THREADS, sizeof(PkSmem<false>), s>>>(a.get(), nullptr, 0, b.get(), m, shift, ntiles, mk,
		                                                                 counts.get(), totals.get(), nullptr, nullptr, 0);
		cudaEventRecord(e3, s);
		cudaStreamSynchronize(s);
		float x, y, z;
		cudaEventElapsedTime(&x, e0, e1);
		cudaEventElapsedTime(&y, e1, e2);
		cudaEventElapsedTime(&z, e2, e3);
		if (r >= 2) {
			th += x;
			ts += y;
			tc += z;
		}
		if (r == 0) { // check: stable sort by that digit
			std::vector<uint64_t> got(m);
			cudaMemcpy(got.data(), b.get(), m * 8, cudaMemcpyDeviceToHost);
			std::vector<uint64_t> want(h);
			std::stable_sort(want.begin(), want.end(),
			                 [shift](uint64_t p, uint64_t q) { return ((p >> shift) & 255) < ((q >> shift) & 255); });
			printf("check %s\n", got == want ? "ok" : "WRONG");
		}
	}
	cudaError_t e = cudaGetLastError();
	printf("m=%lld hist %.1f us scan %.1f us scatter %.1f us (%.0f GB/s)  [%s]\n", (long long)m, 1e3 * th / reps,
	       1e3 * ts / reps, 1e3 * tc / reps, 16.0 * m / (tc / reps * 1e-3) / 1e9, cudaGetErrorString(e));
	return 0;
}
