// Small CUDA plumbing shared by the kernels files: error handling, stream-ordered
// temporary buffers, launch geometry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace phy
{

struct CudaError : std::runtime_error {
	explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line)
{
	if (e != cudaSuccess) {
		char buf[512];
		snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
		throw CudaError(buf);
	}
}

#define CUDA_CHECK(x) ::phy::cuda_check((x), #x, __FILE__, __LINE__)
// every kernel launch of the library goes through KERNEL_CHECK; the count is reported as
// the "launches" statistic
inline std::atomic<uint64_t> g_kernel_launches{0};
#define KERNEL_CHECK()                                                                              \
	do {                                                                                            \
		::phy::g_kernel_launches++;                                                                 \
		::phy::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__);                 \
	} while (0)

constexpr int NUM_SMS_B200 = 148;

// Tuning options (phylo_set_option) belong to a context.  The kernels' host code reads them
// from this thread-local copy, which every C-ABI call refreshes from its context on entry
// (a context is used by one thread at a time): two contexts in one process — one per GPU —
// never see each other's settings.
struct Tuning {
	int scan_single_pass = 1;  // "scan_mode": 1 = one launch with decoupled look-back, 0 = three launches
	int rs_mode = 0;           // "sort_mode": 0 by size, 1 histogram + scan + scatter, 2 single-pass look-back
	int sort_path = 0;         // "sort_path": 0 pick, 1 always the general sorter (3-bit codes, 64-bit keys)
	int table_direct = 0;      // "table_direct": 0 / 1 entry by entry from the root, 2 level by level
	uint64_t map_batch_bytes = 512ull << 20; // "map_batch_bytes"
	int push_kernel = 1;       // "push_kernel": 1 = the last batch of rows goes to the peers by a kernel, 0 = copy engines
	int esa_speculative = 1;   // "esa_speculative": 1 = index build without host round trips (checked at its end)
	int compare_path = 0;      // "compare_path": 0 = TMA + mbarrier pipeline, 1 = cp.async double buffering
	int upload_raw = 0;        // "upload_raw": 1 = sequences cross PCIe as bytes instead of packed to 2 bits
	int esa_graph = 1;         // "esa_graph": 1 = the speculative index build is replayed as a CUDA graph
	int map_graph = 1;         // "map_graph": the mapping of a batch (+ rows, comparison) as one graph: 0 never, 1 the first batch of a call if it has 4 Mbp or more, 2 every batch
};
inline thread_local Tuning g_tuning;

// Per-device "done once" flags (kernel attributes are per device: a process that drives
// several GPUs has to set them on each).
struct PerDeviceOnce {
	std::atomic<bool> done[64] = {};
	// true the first time it is asked on the current device
	bool first()
	{
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
		return !done[dev].exchange(true);
	}
};

// ---- transaction barriers and bulk copies (Blackwell tile movement) -------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "MBAR_WAIT:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	             "@p bra MBAR_DONE;\n"
	             "bra MBAR_WAIT;\n"
	             "MBAR_DONE:\n"
	             "}" ::"r"(smem_u32(bar)),
	             "r"(parity)
	             : "memory");
}
// has the phase with this parity completed?  (does not wait)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
	             "selp.u32 %0, 1, 0, p;\n"
	             "}"
	             : "=r"(ok)
	             : "r"(smem_u32(bar)), "r"(parity)
	             : "memory");
	return ok != 0;
}
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy
// ones (a bulk copy into a buffer the threads have just been reading and writing)
__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// bytes (multiple of 16, 16-byte aligned on both sides) global -> shared, completes on bar
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :
	             : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
#endif

inline int div_up(int64_t a, int64_t b)
{
	return (int)((a + b - 1) / b);
}

// Device memory for the library's buffers.  A step of the pipeline asks for the same few
// dozen buffers in the same order every time, so freed blocks are kept and handed out again:
// after the first step no allocation reaches the driver (cudaMallocAsync's pool re-maps
// memory when the size pattern changes, which showed up as idle gaps of 0.1..40 ms on B200).
// Blocks are cached per (device, stream): a block freed on a stream is only reused on that
// stream, where the earlier work on it is ordered before the later one.
class BlockCache
{
	struct Block {
		size_t size;
		int device;
		cudaStream_t stream;
	};
	std::mutex mu_;
	std::map<std::pair<int, cudaStream_t>, std::multimap<size_t, void *>> free_;
	std::unordered_map<void *, Block> live_;
	size_t cached_bytes_ = 0;

	static size_t round_size(size_t b)
	{
		const size_t q = b <= (1u << 20) ? 512 : (2u << 20);
		return (b + q - 1) / q * q;
	}

  public:
	void *alloc(size_t bytes, cudaStream_t s)
	{
		if (!bytes) return nullptr;
		int dev = 0;
		cuda_check(cudaGetDevice(&dev), "cudaGetDevice", __FILE__, __LINE__);
		const size_t want = round_size(bytes);
		std::lock_guard<std::mutex> lock(mu_);
		auto &pool = free_[{dev, s}];
		auto it = pool.lower_bound(want);
		void *p = nullptr;
		size_t got = want;
		if (it != pool.end() && it->first <= want + want / 2) {
			p = it->second;
			got = it->first;
			pool.erase(it);
			cached_bytes_ -= got;
		} else {
			cudaError_t e = cudaMalloc(&p, want);
			if (e == cudaErrorMemoryAllocation) {
				cudaGetLastError();
				trim_locked(); // give everything cached back to the driver and try once more
				e = cudaMalloc(&p, want);
			}
			cuda_check(e, "cudaMalloc", __FILE__, __LINE__);
		}
		live_[p] = Block{got, dev, s};
		return p;
	}
	void free(void *p)
	{
		if (!p) return;
		std::lock_guard<std::mutex> lock(mu_);
		auto it = live_.find(p);
		if (it == live_.end()) return;
		const Block b = it->second;
		live_.erase(it);
		free_[{b.device, b.stream}].emplace(b.size, p);
		cached_bytes_ += b.size;
	}
	// returns the cached blocks of one stream (or of all, stream == (cudaStream_t)-1) to the driver
	void trim(int device, cudaStream_t stream)
	{
		std::lock_guard<std::mutex> lock(mu_);
		for (auto it = free_.begin(); it != free_.end();) {
			if (it->first.first == device && (stream == (cudaStream_t)-1 || it->first.second == stream)) {
				for (auto &kv : it->second) {
					cudaFree(kv.second);
					cached_bytes_ -= kv.first;
				}
				it = free_.erase(it);
			} else {
				++it;
			}
		}
	}
	size_t cached_bytes() const { return cached_bytes_; }

  private:
	void trim_locked()
	{
		cudaDeviceSynchronize();
		int dev = 0;
		cudaGetDevice(&dev);
		for (auto it = free_.begin(); it != free_.end();) {
			if (it->first.first == dev) {
				for (auto &kv : it->second) {
					cudaFree(kv.second);
					cached_bytes_ -= kv.first;
				}
				it = free_.erase(it);
			} else {
				++it;
			}
		}
	}
};

inline BlockCache g_block_cache;

// Device buffer from the block cache; `stream` is the stream the buffer is used on.
template <typename T> class DevBuf
{
	T *p_ = nullptr;
	size_t n_ = 0;
	cudaStream_t s_ = 0;

  public:
	DevBuf() = default;
	DevBuf(size_t n, cudaStream_t s) { alloc(n, s); }
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	DevBuf(DevBuf &&o) noexcept : p_(o.p_), n_(o.n_), s_(o.s_) { o.p_ = nullptr; o.n_ = 0; }
	DevBuf &operator=(DevBuf &&o) noexcept
	{
		if (this != &o) {
			release();
			p_ = o.p_;
			n_ = o.n_;
			s_ = o.s_;
			o.p_ = nullptr;
			o.n_ = 0;
		}
		return *this;
	}
	~DevBuf() { release(); }
	void alloc(size_t n, cudaStream_t s)
	{
		release();
		s_ = s;
		n_ = n;
		if (n) p_ = static_cast<T *>(g_block_cache.alloc(n * sizeof(T), s));
	}
	void release()
	{
		if (p_) g_block_cache.free(p_);
		p_ = nullptr;
		n_ = 0;
	}
	void zero() { if (n_) CUDA_CHECK(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s_)); }
	T *get() const { return p_; }
	size_t size() const { return n_; }
	size_t bytes() const { return n_ * sizeof(T); }
	void swap(DevBuf &o)
	{
		std::swap(p_, o.p_);
		std::swap(n_, o.n_);
		std::swap(s_, o.s_);
	}
};

// Pinned host memory for the small read-backs of a call (counts, flags, verdicts).  A copy to
// ordinary memory is staged by the driver and costs the calling thread a full round trip to
// the GPU EACH (measured on B200: ~14 us per cudaMemcpyAsync, 13 of them per process() pass);
// copies to pinned memory are queued like kernels and one synchronisation covers them all.
// One arena per host thread, handed out stack-wise: a Scope gives back what was taken inside it.
class PinnedArena
{
	struct Block {
		char *p;
		size_t cap;
	};
	std::vector<Block> blocks_; // the last one is current; earlier (smaller) ones live until the outermost scope ends
	size_t used_ = 0;
	int depth_ = 0;

  public:
	PinnedArena() = default;
	PinnedArena(const PinnedArena &) = delete;
	PinnedArena &operator=(const PinnedArena &) = delete;
	~PinnedArena()
	{
		for (auto &b : blocks_)
			cudaFreeHost(b.p); // (an error at process teardown is of no consequence)
	}
	struct Scope {
		PinnedArena &a;
		size_t mark;
		size_t nblocks;
		explicit Scope(PinnedArena &arena) : a(arena), mark(arena.used_), nblocks(arena.blocks_.size()) { a.depth_++; }
		Scope(const Scope &) = delete;
		Scope &operator=(const Scope &) = delete;
		~Scope()
		{
			if (a.blocks_.size() == nblocks) a.used_ = mark; // same block: pop; a new block started empty
			if (--a.depth_ == 0) {
				while (a.blocks_.size() > 1) {
					cudaFreeHost(a.blocks_.front().p);
					a.blocks_.erase(a.blocks_.begin());
				}
				a.used_ = 0;
			}
		}
	};
	// n objects of T, 64-byte aligned, valid until the enclosing Scope ends; not initialised
	template <typename T> T *take(size_t n)
	{
		const size_t bytes = (n * sizeof(T) + 63) / 64 * 64;
		if (blocks_.empty() || used_ + bytes > blocks_.back().cap) {
			size_t cap = blocks_.empty() ? (size_t)(64 << 10) : 2 * blocks_.back().cap;
			while (cap < bytes)
				cap *= 2;
			void *p = nullptr;
			cuda_check(cudaHostAlloc(&p, cap, cudaHostAllocPortable), "cudaHostAlloc", __FILE__, __LINE__);
			blocks_.push_back(Block{static_cast<char *>(p), cap});
			used_ = 0;
		}
		T *out = reinterpret_cast<T *>(blocks_.back().p + used_);
		used_ += bytes;
		return out;
	}
};

inline thread_local PinnedArena g_pinned;

// A launch sequence submitted as ONE CUDA graph.  Every call captures the sequence anew (the
// host code runs as usual, its launches are recorded instead of submitted), brings the
// instantiated graph of the previous call up to date — same kernels, new parameters — and
// launches that.  Measured on B200 for 30 kernels: capture 10 us + update 7 us + launch 27 us of
// host time against 122 us for 30 launches; on the device 0.7 us from kernel to kernel instead of
// 2.6 us, and no 4..6 us more per kernel while a host-to-device copy is in flight (the GPU fetches
// every launch from host memory over the same bus).  A sequence whose shape changed (other
// kernels, another number of them) is instantiated again (~160 us).
struct GraphSegment {
	cudaGraphExec_t exec = nullptr;
	cudaStream_t capturing = nullptr;
	uint64_t instantiated = 0, updated = 0;
	void begin(cudaStream_t s)
	{
		// relaxed: the block cache may have to call cudaMalloc while the capture is on
		CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
		capturing = s;
	}
	// ends the capture and submits what was captured
	void launch()
	{
		cudaStream_t s = capturing;
		capturing = nullptr;
		cudaGraph_t g = nullptr;
		CUDA_CHECK(cudaStreamEndCapture(s, &g));
		struct Drop {
			cudaGraph_t g;
			~Drop() { cudaGraphDestroy(g); }
		} drop{g};
		if (exec) {
			cudaGraphExecUpdateResultInfo info;
			if (cudaGraphExecUpdate(exec, g, &info) == cudaSuccess) {
				updated++;
			} else {
				cudaGetLastError();
				cudaGraphExecDestroy(exec);
				exec = nullptr;
			}
		}
		if (!exec) {
			CUDA_CHECK(cudaGraphInstantiate(&exec, g, 0));
			instantiated++;
		}
		CUDA_CHECK(cudaGraphLaunch(exec, s));
	}
	// something threw while capturing: leave the stream usable
	void abandon()
	{
		if (!capturing) return;
		cudaGraph_t g = nullptr;
		cudaStreamEndCapture(capturing, &g);
		if (g) cudaGraphDestroy(g);
		cudaGetLastError();
		capturing = nullptr;
	}
	void destroy()
	{
		abandon();
		if (exec) cudaGraphExecDestroy(exec);
		exec = nullptr;
	}
};

template <typename T> inline T d2h_scalar(const T *dptr, cudaStream_t s)
{
	PinnedArena::Scope scope(g_pinned);
	T *v = g_pinned.take<T>(1);
	CUDA_CHECK(cudaMemcpyAsync(v, dptr, sizeof(T), cudaMemcpyDeviceToHost, s));
	CUDA_CHECK(cudaStreamSynchronize(s));
	return *v;
}

} // namespace phy
