// Does a host-to-device copy in flight slow down a chain of short kernels on another stream?
// (phylo_process builds the index while the query sequences cross PCIe: measured 0.99 ms
// against 0.71 ms for the same 35 launches without the copies.)  Cases: the chain alone, next to
// copies, as a CUDA graph next to copies, and with the launches queued before the copies start.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/launchbench tools/launchbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>

#define CK(x)                                                                                     \
	do {                                                                                          \
		cudaError_t e = (x);                                                                      \
		if (e != cudaSuccess) {                                                                   \
			fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));             \
			exit(1);                                                                              \
		}                                                                                         \
	} while (0)

__global__ void spin(long long ns, unsigned *sink)
{
	unsigned long long t0, t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	do {
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	} while ((long long)(t - t0) < ns);
	if (sink && t == 0) *sink = 1;
}

int main(int argc, char **argv)
{
	const int chain = argc > 1 ? atoi(argv[1]) : 36;
	const long long ns = argc > 2 ? atoll(argv[2]) : 18000;
	const size_t piece = 5000000, pieces = 7;
	cudaStream_t s, c;
	CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
	char *h, *d;
	CK(cudaHostAlloc((void **)&h, piece * pieces, cudaHostAllocDefault));
	CK(cudaMalloc((void **)&d, piece * pieces));
	cudaEvent_t e0, e1, go;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	CK(cudaEventCreateWithFlags(&go, cudaEventDisableTiming));

	cudaGraph_t graph;
	cudaGraphExec_t exec;
	CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
	for (int k = 0; k < chain; k++)
		spin<<<148, 128, 0, s>>>(ns, nullptr);
	CK(cudaStreamEndCapture(s, &graph));
	CK(cudaGraphInstantiate(&exec, graph, 0));
	CK(cudaGraphUpload(exec, s));

	auto copies = [&] {
		for (size_t p = 0; p < pieces; p++)
			CK(cudaMemcpyAsync(d + p * piece, h + p * piece, piece, cudaMemcpyHostToDevice, c));
	};
	auto run = [&](const char *name, int mode) {
		float sum = 0;
		const int reps = 20;
		for (int r = 0; r < reps + 3; r++) {
			CK(cudaDeviceSynchronize());
			CK(cudaEventRecord(go, s));
			CK(cudaStreamWaitEvent(c, go, 0));
			if (mode == 1 || mode == 2) copies();
			CK(cudaEventRecord(e0, s));
			if (mode == 2 || mode == 4) {
				CK(cudaGraphLaunch(exec, s));
			} else {
				for (int k = 0; k < chain; k++)
					spin<<<148, 128, 0, s>>>(ns, nullptr);
			}
			CK(cudaEventRecord(e1, s));
			if (mode == 3) copies(); // the chain is queued first, then the copies
			CK(cudaEventSynchronize(e1));
			float ms;
			CK(cudaEventElapsedTime(&ms, e0, e1));
			if (r >= 3) sum += ms;
		}
		printf("%-44s %.3f ms  (%d kernels of %.0f us: %.3f ms of work)\n", name, sum / reps, chain, ns / 1e3, chain * ns / 1e6);
	};
	{
		// what re-capturing the chain on every call costs the host: capture + exec update + launch
		cudaGraph_t g2;
		double cap = 0, upd = 0, lau = 0;
		const int reps = 50;
		for (int r = 0; r < reps; r++) {
			CK(cudaDeviceSynchronize());
			auto t0 = std::chrono::steady_clock::now();
			CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
			for (int k = 0; k < chain; k++)
				spin<<<148 + (r & 1), 128, 0, s>>>(ns + r, nullptr); // (parameters change from call to call)
			CK(cudaStreamEndCapture(s, &g2));
			auto t1 = std::chrono::steady_clock::now();
			cudaGraphExecUpdateResultInfo info;
			CK(cudaGraphExecUpdate(exec, g2, &info));
			auto t2 = std::chrono::steady_clock::now();
			CK(cudaGraphLaunch(exec, s));
			auto t3 = std::chrono::steady_clock::now();
			CK(cudaGraphDestroy(g2));
			cap += std::chrono::duration<double, std::micro>(t1 - t0).count();
			upd += std::chrono::duration<double, std::micro>(t2 - t1).count();
			lau += std::chrono::duration<double, std::micro>(t3 - t2).count();
		}
		printf("host cost per call, %d nodes: capture %.1f us, exec update %.1f us, launch %.1f us\n", chain, cap / reps, upd / reps, lau / reps);
		CK(cudaDeviceSynchronize());
		auto t0 = std::chrono::steady_clock::now();
		for (int k = 0; k < chain; k++)
			spin<<<148, 128, 0, s>>>(1000, nullptr);
		auto t1 = std::chrono::steady_clock::now();
		printf("host cost of %d direct launches: %.1f us\n", chain, std::chrono::duration<double, std::micro>(t1 - t0).count());
		cudaGraphExec_t e2;
		CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
		for (int k = 0; k < chain; k++)
			spin<<<148, 128, 0, s>>>(ns, nullptr);
		CK(cudaStreamEndCapture(s, &g2));
		t0 = std::chrono::steady_clock::now();
		CK(cudaGraphInstantiate(&e2, g2, 0));
		t1 = std::chrono::steady_clock::now();
		printf("instantiate: %.1f us\n", std::chrono::duration<double, std::micro>(t1 - t0).count());
		CK(cudaDeviceSynchronize());
	}
	run("chain alone", 0);
	run("chain next to H2D copies (35 MB pinned)", 1);
	run("graph next to H2D copies", 2);
	run("chain queued first, then the copies", 3);
	run("graph alone", 4);
	return 0;
}
