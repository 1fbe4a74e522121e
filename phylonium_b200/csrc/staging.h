// Host-to-device copies from ordinary (pageable) host memory.
//
// phylo_process() receives the sequences where the caller keeps them — in the reference that
// is the std::string inside a `sequence` (src/sequence.h), i.e. pageable memory.  A
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread, one
// copy after the other, while that thread should be launching the index build and the
// mapping.  Here a few worker threads copy pieces of the sequences into pinned ring buffers
// and issue the asynchronous copies from there, each on a stream of its own; the mapping
// waits, batch by batch, for an event behind the last piece of the batch.
#pragma once
#include "common.cuh"

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace phy
{

class HostStager
{
  public:
	struct Piece {
		uint8_t *dst;       // device
		const uint8_t *src; // host, pageable
		uint32_t len;       // <= SLOT_BYTES
		int32_t batch;
	};
	static constexpr size_t SLOT_BYTES = 4u << 20;
	static constexpr int SLOTS = 4;

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

	// appends the pieces of one sequence (call in batch order, then start())
	void add(uint8_t *dst, const void *src, uint64_t len, int batch)
	{
		const uint8_t *p = static_cast<const uint8_t *>(src);
		for (uint64_t o = 0; o < len; o += SLOT_BYTES) {
			const uint64_t l = len - o < SLOT_BYTES ? len - o : SLOT_BYTES;
			pending_.push_back(Piece{dst + o, p + o, (uint32_t)l, batch});
		}
	}
	bool empty() const { return pending_.empty(); }

	// `after`: an event (of another stream) the copies must not overtake, e.g. the clearing of
	// the destination buffer
	void start(int device, int nbatches, int threads, cudaEvent_t after)
	{
		finish();
		pieces_.swap(pending_);
		pending_.clear();
		if (threads < 1) threads = 1;
		if ((size_t)threads > pieces_.size()) threads = (int)(pieces_.size() ? pieces_.size() : 1);
		device_ = device;
		nbatches_ = nbatches;
		abort_ = false;
		error_.clear();
		if ((int)workers_.size() != threads) {
			release_workers();
			workers_.resize(threads);
			for (auto &w : workers_) {
				CUDA_CHECK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
				CUDA_CHECK(cudaHostAlloc((void **)&w.ring, SLOT_BYTES * SLOTS, cudaHostAllocDefault));
				for (auto &e : w.slot_ev)
					CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			}
		}
		for (auto &w : workers_) {
			while ((int)w.batch_ev.size() < nbatches) {
				cudaEvent_t e;
				CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
				w.batch_ev.push_back(e);
			}
			w.issued = 0;
			if (after) CUDA_CHECK(cudaStreamWaitEvent(w.stream, after, 0));
		}
		running_ = true;
		for (int t = 0; t < threads; t++)
			workers_[t].th = std::thread([this, t] { run(t); });
	}

	// blocks until every worker has issued its copies of batch b, then puts `s` behind them
	void wait_batch(int b, cudaStream_t s)
	{
		if (!running_) return;
		std::unique_lock<std::mutex> lock(mu_);
		for (auto &w : workers_) {
			cv_.wait(lock, [&] { return w.issued > b || !error_.empty(); });
			if (!error_.empty()) throw CudaError("staging copy failed: " + error_);
			CUDA_CHECK(cudaStreamWaitEvent(s, w.batch_ev[b], 0));
		}
	}

	// joins the workers; with abort they stop at the next piece.  Afterwards nothing reads the
	// caller's buffers any more (the copies out of the pinned rings may still be in flight).
	void finish(bool abort = false)
	{
		if (!running_) return;
		if (abort) abort_ = true;
		for (auto &w : workers_)
			if (w.th.joinable()) w.th.join();
		running_ = false;
		pieces_.clear();
	}

	// waits for the copies themselves (error paths, context destruction)
	void drain()
	{
		finish(true);
		pending_.clear();
		for (auto &w : workers_)
			if (w.stream) cudaStreamSynchronize(w.stream);
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
		cudaStream_t stream = nullptr;
		uint8_t *ring = nullptr;
		cudaEvent_t slot_ev[SLOTS] = {};
		std::vector<cudaEvent_t> batch_ev;
		int issued = 0; // batches whose events have been recorded (guarded by mu_)
	};

	void release_workers()
	{
		for (auto &w : workers_) {
			if (w.stream) {
				cudaStreamSynchronize(w.stream);
				cudaStreamDestroy(w.stream);
			}
			if (w.ring) cudaFreeHost(w.ring);
			for (auto e : w.slot_ev)
				if (e) cudaEventDestroy(e);
			for (auto e : w.batch_ev)
				cudaEventDestroy(e);
		}
		workers_.clear();
	}

	void run(int t)
	{
		Worker &w = workers_[t];
		const int T = (int)workers_.size();
		int batch = 0;
		auto publish_up_to = [&](int upto) { // events of batches [batch, upto)
			for (; batch < upto; batch++) {
				const cudaError_t e = cudaEventRecord(w.batch_ev[batch], w.stream);
				std::lock_guard<std::mutex> lock(mu_);
				if (e != cudaSuccess && error_.empty()) error_ = cudaGetErrorString(e);
				w.issued = batch + 1;
				cv_.notify_all();
			}
		};
		cudaError_t e = cudaSetDevice(device_);
		size_t used = 0;
		for (size_t i = t; e == cudaSuccess && i < pieces_.size() && !abort_; i += T) {
			const Piece &p = pieces_[i];
			publish_up_to(p.batch);
			const int slot = (int)(used++ % SLOTS);
			e = cudaEventSynchronize(w.slot_ev[slot]); // the copy that last used this slot has left it
			if (e != cudaSuccess) break;
			uint8_t *stage = w.ring + (size_t)slot * SLOT_BYTES;
			std::memcpy(stage, p.src, p.len);
			e = cudaMemcpyAsync(p.dst, stage, p.len, cudaMemcpyHostToDevice, w.stream);
			if (e != cudaSuccess) break;
			e = cudaEventRecord(w.slot_ev[slot], w.stream);
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
	std::condition_variable cv_;
	std::string error_;
	std::atomic<bool> abort_{false};
	bool running_ = false;
	int device_ = 0, nbatches_ = 0;
};

} // namespace phy
