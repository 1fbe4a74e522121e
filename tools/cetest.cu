// Does a small H2D copy on stream B wait behind bulk H2D copies queued earlier on stream A?
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
__global__ void touch(int *p) { p[0] += 1; }
int main()
{
	const size_t total = 1ull << 30, piece = 16ull << 20;
	char *h, *d;
	int *hs, *ds;
	cudaHostAlloc(&h, total, cudaHostAllocDefault);
	cudaMalloc(&d, total);
	cudaHostAlloc(&hs, 4096, cudaHostAllocDefault);
	cudaMalloc(&ds, 4096);
	cudaStream_t a, b;
	cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
	cudaEvent_t e0, ea, eb;
	cudaEventCreate(&e0);
	cudaEventCreate(&ea);
	cudaEventCreate(&eb);
	for (int rep = 0; rep < 3; rep++) {
		for (int mode = 0; mode < 5; mode++) {
			cudaDeviceSynchronize();
			cudaEventRecord(e0, a);
			for (size_t o = 0; o < total; o += piece)
				cudaMemcpyAsync(d + o, h + o, piece, cudaMemcpyHostToDevice, a);
			cudaEventRecord(ea, a);
			cudaStreamWaitEvent(b, e0, 0);
			if (mode == 0) cudaMemcpyAsync(ds, hs, 4096, cudaMemcpyHostToDevice, b); // small H2D
			if (mode == 1) cudaMemcpyAsync(hs, ds, 4096, cudaMemcpyDeviceToHost, b); // small D2H
			static int pageable[1024];
			if (mode == 3) cudaMemcpyAsync(pageable, ds, 4096, cudaMemcpyDeviceToHost, b); // small D2H, pageable
			if (mode == 4) cudaMemcpyAsync(ds, pageable, 4096, cudaMemcpyHostToDevice, b); // small H2D, pageable
			touch<<<1, 1, 0, b>>>(ds);
			cudaEventRecord(eb, b);
			cudaDeviceSynchronize();
			float ta, tb;
			cudaEventElapsedTime(&ta, e0, ea);
			cudaEventElapsedTime(&tb, e0, eb);
			printf("mode %d (%s): bulk 1 GiB H2D done after %.2f ms (%.1f GB/s), stream B done after %.2f ms\n", mode,
			       mode == 0 ? "small H2D + kernel" : mode == 1 ? "small D2H + kernel" : mode == 2 ? "kernel only" : mode == 3 ? "pageable D2H + kernel" : "pageable H2D + kernel", ta, total / ta / 1e6, tb);
		}
	}
	return 0;
}
