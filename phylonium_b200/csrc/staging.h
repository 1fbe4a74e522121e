// Host-to-device upload of the sequences, packed.
//
// phylo_process() receives the sequences where the caller keeps them — in the reference that
// is the std::string inside a `sequence` (src/sequence.h), i.e. ordinary pageable memory — and
// for a thousand genomes the PCIe bus, not the GPU, bounds the call.  So the bytes do not
// cross the bus as they are: a few worker threads pack pieces of the sequences to 2 bits per
// base (host_pack.cpp; the alphabet check happens there too) into small pinned ring buffers,
// copy the packed pieces over — a quarter of the bytes, and from pinned memory whatever the
// caller's buffers are — and a kernel on the worker's stream unpacks each piece to bytes at
// its final place.  The mapping waits, batch by batch, for an event behind the last piece of
// the batch.  Pageable and pinned callers take the same path.
#pragma once
#include "common.cuh"
#include "host_pack.h"

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace phy
{

// packed piece -> bytes; 16 bases (one 32-bit word of codes) per thread
static __global__ void __launch_bounds__(256)
k_unpack_2bit(const uint32_t *__restrict__ packed, uint8_t *__restrict__ dst, uint32_t n)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t i0 = t * 16;
	if (i0 >= n) return;
	const uint32_t w = packed[t];
	uint32_t out[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		uint32_t x = 0;
#pragma unroll
		for (int b = 0; b < 4; b++) {
			const uint32_t code = (w >> (8 * k + 2 * b)) & 3u;
			x |= ((0x47544341u >> (8 * code)) & 0xffu) << (8 * b); // 0 A, 1 C, 2 T, 3 G
		}
		out[k] = x;
	}
	if (i0 + 16 <= n) {
		*reinterpret_cast<uint4 *>(dst + i0) = make_uint4(out[0], out[1], out[2], out[3]); // dst is 16-byte aligned
	} else {
		for (uint32_t i = i0; i < n; i++)
			dst[i] = (uint8_t)(out[(i - i0) >> 2] >> (8 * ((i - i0) & 3)));
	}
}

static __global__ void k_put_bangs(uint8_t *__restrict__ dst, const uint32_t *__restrict__ pos, uint32_t count)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < count) dst[pos[t]] = (uint8_t)'!';
}

// resources of one upload lane: a stream, a pinned ring of packed pieces and its twin on the device
struct UploadLane {
	static constexpr size_t PIECE_BYTES = 2u << 20;         // bases per piece
	static constexpr size_t PACKED_BYTES = PIECE_BYTES / 4; // its packed form
	static constexpr uint32_t BANG_CAP = 4096;              // '!' per piece listed with the packed form
	static constexpr size_t SLOT_BYTES = PACKED_BYTES + BANG_CAP * sizeof(uint32_t);
	static constexpr int SLOTS = 4;
	cudaStream_t stream = nullptr;
	uint8_t *ring = nullptr;  // pinned: SLOTS x (packed piece, '!' list)
	uint8_t *dring = nullptr; // the same on the device
	cudaEvent_t slot_ev[SLOTS] = {};
	size_t used = 0;

	void create()
	{
		CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		CUDA_CHECK(cudaHostAlloc((void **)&ring, SLOT_BYTES * SLOTS, cudaHostAllocDefault));
		CUDA_CHECK(cudaMalloc((void **)&dring, SLOT_BYTES * SLOTS));
		for (auto &e : slot_ev)
			CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	void destroy()
	{
		if (stream) {
			cudaStreamSynchronize(stream);
			cudaStreamDestroy(stream);
		}
		if (ring) cudaFreeHost(ring);
		if (dring) cudaFree(dring);
		for (auto e : slot_ev)
			if (e) cudaEventDestroy(e);
		stream = nullptr;
		ring = dring = nullptr;
		for (auto &e : slot_ev)
			e = nullptr;
	}
	// packs src[0, len), len <= PIECE_BYTES, sends it and unpacks it at dst (16-byte aligned).
	// *bad = 1: a byte outside the alphabet (nothing sent).  The host bytes are not needed any
	// more when this returns.
	cudaError_t upload_piece(uint8_t *dst, const uint8_t *src, uint32_t len, int *bad)
	{
		const int slot = (int)(used++ % SLOTS);
		cudaError_t e = cudaEventSynchronize(slot_ev[slot]); // the copy that last used this slot has left it
		if (e != cudaSuccess) return e;
		uint8_t *hp = ring + (size_t)slot * SLOT_BYTES, *dp = dring + (size_t)slot * SLOT_BYTES;
		uint32_t *hb = reinterpret_cast<uint32_t *>(hp + PACKED_BYTES);
		uint32_t nb = 0;
		*bad = pack_2bit(src, len, hp, hb, BANG_CAP, &nb);
		if (*bad) return cudaSuccess;
		if (nb > BANG_CAP) {
			// a piece that is mostly separators: as it is, through the driver's own staging, and
			// done before we return (the caller may reuse src)
			e = cudaMemcpyAsync(dst, src, len, cudaMemcpyHostToDevice, stream);
			if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
			return e;
		}
		const size_t pbytes = ((size_t)len + 3) / 4;
		e = cudaMemcpyAsync(dp, hp, (pbytes + 3) / 4 * 4, cudaMemcpyHostToDevice, stream);
		if (e != cudaSuccess) return e;
		k_unpack_2bit<<<(unsigned)((len + 16 * 256 - 1) / (16 * 256)), 256, 0, stream>>>(
			reinterpret_cast<const uint32_t *>(dp), dst, len);
		g_kernel_launches++;
		e = cudaGetLastError();
		if (e == cudaSuccess && nb) {
			e = cudaMemcpyAsync(dp + PACKED_BYTES, hb, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
			if (e != cudaSuccess) return e;
			k_put_bangs<<<(nb + 255) / 256, 256, 0, stream>>>(dst, reinterpret_cast<const uint32_t *>(dp + PACKED_BYTES), nb);
			g_kernel_launches++;
			e = cudaGetLastError();
		}
		if (e != cudaSuccess) return e;
		return cudaEventRecord(slot_ev[slot], stream);
	}
};

class HostStager
{
  public:
	struct Piece {
		uint8_t *dst;       // device, 16-byte aligned
		const uint8_t *src; // host
		uint32_t len;       // <= PIECE_BYTES
		int32_t batch;
	};
	static constexpr size_t PIECE_BYTES = UploadLane::PIECE_BYTES;

	// true if a cudaMemcpyAsync from p would be staged by the driver
	static bool is_pageable(const void *p)
	{
		cudaPointerAttributes a;
		if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
			cudaGetLastError();
			return true;
		}
		return a.type == cudaMemoryTypeUnregistered;
	}

	// appends the pieces of one sequence (call in batch order, then start()); dst 16-byte aligned.
	// piece: bases per piece, a multiple of 64, at most PIECE_BYTES (smaller pieces spread one
	// sequence that everything waits for — the reference — over all workers)
	void add(uint8_t *dst, const void *src, uint64_t len, int batch, uint64_t piece = PIECE_BYTES)
	{
		if (piece > PIECE_BYTES || piece < 64 || piece % 64) piece = PIECE_BYTES;
		const uint8_t *p = static_cast<const uint8_t *>(src);
		for (uint64_t o = 0; o < len; o += piece) {
			const uint64_t l = len - o < piece ? len - o : piece;
			pending_.push_back(Piece{dst + o, p + o, (uint32_t)l, batch});
		}
	}
	bool empty() const { return pending_.empty(); }

	// `after`: an event (of another stream) the copies must not overtake, e.g. the clearing of
	// the destination buffer.  The worker threads live as long as the stager (a call of
	// phylo_process on 8 x 5 Mbp takes 2 ms: creating threads per call would show).
	void start(int device, int nbatches, int threads, cudaEvent_t after)
	{
		finish();
		pieces_.swap(pending_);
		pending_.clear();
		if (threads < 1) threads = 1;
		device_ = device;
		nbatches_ = nbatches;
		abort_ = false;
		error_.clear();
		bad_input_ = false;
		if ((int)workers_.size() != threads) {
			release_workers();
			workers_.resize(threads);
			for (auto &w : workers_)
				w.lane.create();
			quit_ = false;
			for (int t = 0; t < threads; t++)
				workers_[t].th = std::thread([this, t, g = generation_] { thread_main(t, g); }); // jobs posted from now on
		}
		for (auto &w : workers_) {
			while ((int)w.batch_ev.size() < nbatches) {
				cudaEvent_t e;
				CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
				w.batch_ev.push_back(e);
			}
			if (after) CUDA_CHECK(cudaStreamWaitEvent(w.lane.stream, after, 0));
		}
		{
			std::lock_guard<std::mutex> lock(mu_);
			for (auto &w : workers_)
				w.issued = 0;
			done_ = 0;
			generation_++;
			running_ = true;
		}
		cv_job_.notify_all();
	}

	// blocks until every worker has issued its copies of batch b, then puts `s` behind them;
	// throws std::invalid_argument if a byte outside the alphabet was met on the way
	void wait_batch(int b, cudaStream_t s)
	{
		if (!running_) return;
		std::unique_lock<std::mutex> lock(mu_);
		for (auto &w : workers_) {
			cv_.wait(lock, [&] { return w.issued > b || !error_.empty() || bad_input_; });
			if (bad_input_) throw std::invalid_argument("a sequence contains bytes outside {A,C,G,T,!}");
			if (!error_.empty()) throw CudaError("staging copy failed: " + error_);
			CUDA_CHECK(cudaStreamWaitEvent(s, w.batch_ev[b], 0));
		}
	}

	// waits until the workers are through with the job; with abort they stop at the next piece.
	// Afterwards nothing reads the caller's buffers any more (the copies out of the pinned
	// rings may still be in flight).
	void finish(bool abort = false)
	{
		if (!running_) return;
		if (abort) abort_ = true;
		std::unique_lock<std::mutex> lock(mu_);
		cv_.wait(lock, [&] { return done_ == (int)workers_.size(); });
		running_ = false;
		pieces_.clear();
	}

	// waits for the copies themselves (error paths, context destruction)
	void drain()
	{
		finish(true);
		pending_.clear();
		for (auto &w : workers_)
			if (w.lane.stream) cudaStreamSynchronize(w.lane.stream);
	}

	void release()
	{
		drain();
		release_workers();
	}
	~HostStager() { release(); }

  private:
	struct Worker {
		std::thread th;
		UploadLane lane;
		std::vector<cudaEvent_t> batch_ev;
		int issued = 0; // batches whose events have been recorded (guarded by mu_)
	};

	void release_workers()
	{
		{
			std::lock_guard<std::mutex> lock(mu_);
			quit_ = true;
		}
		cv_job_.notify_all();
		for (auto &w : workers_)
			if (w.th.joinable()) w.th.join();
		for (auto &w : workers_) {
			w.lane.destroy();
			for (auto e : w.batch_ev)
				cudaEventDestroy(e);
		}
		workers_.clear();
	}

	void thread_main(int t, uint64_t seen)
	{
		for (;;) {
			{
				std::unique_lock<std::mutex> lock(mu_);
				cv_job_.wait(lock, [&] { return quit_ || generation_ != seen; });
				if (quit_) return;
				seen = generation_;
			}
			run(t);
			{
				std::lock_guard<std::mutex> lock(mu_);
				done_++;
			}
			cv_.notify_all();
		}
	}

	void run(int t)
	{
		Worker &w = workers_[t];
		const int T = (int)workers_.size();
		int batch = 0;
		auto publish_up_to = [&](int upto) { // events of batches [batch, upto)
			for (; batch < upto; batch++) {
				const cudaError_t e = cudaEventRecord(w.batch_ev[batch], w.lane.stream);
				std::lock_guard<std::mutex> lock(mu_);
				if (e != cudaSuccess && error_.empty()) error_ = cudaGetErrorString(e);
				w.issued = batch + 1;
				cv_.notify_all();
			}
		};
		cudaError_t e = cudaSetDevice(device_);
		for (size_t i = t; e == cudaSuccess && i < pieces_.size() && !abort_; i += T) {
			const Piece &p = pieces_[i];
			publish_up_to(p.batch);
			int bad = 0;
			e = w.lane.upload_piece(p.dst, p.src, p.len, &bad);
			if (bad) {
				std::lock_guard<std::mutex> lock(mu_);
				bad_input_ = true;
				abort_ = true;
				cv_.notify_all();
				break;
			}
		}
		if (e != cudaSuccess) {
			std::lock_guard<std::mutex> lock(mu_);
			if (error_.empty()) error_ = cudaGetErrorString(e);
			cv_.notify_all();
		}
		publish_up_to(nbatches_);
	}

	std::vector<Piece> pending_; // added, not started yet
	std::vector<Piece> pieces_;  // what the running workers read
	std::vector<Worker> workers_;
	std::mutex mu_;
	std::condition_variable cv_, cv_job_;
	uint64_t generation_ = 0; // guarded by mu_: bumped per job
	int done_ = 0;            // workers through with the current job
	bool quit_ = false;
	std::string error_;
	bool bad_input_ = false; // guarded by mu_
	std::atomic<bool> abort_{false};
	bool running_ = false;
	int device_ = 0, nbatches_ = 0;
};

} // namespace phy
