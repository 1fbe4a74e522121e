// Device-wide building blocks written for this library: scan, stream compaction and a
// stable LSD radix sort of (64-bit key, 32-bit value) pairs.  No CUB/Thrust on the path.
//
// Radix sort: 8-bit digits.  Each pass is three launches
//   1. rs_histogram   per-tile digit counts            reads  8 B/element
//   2. scan           exclusive sum over (digit, tile)  tiny
//   3. rs_scatter     stable rank inside the tile (warp match_any + per-warp counters),
//                     exchange through shared memory so that every digit's run leaves the
//                     SM as one contiguous, coalesced store   reads 12 B, writes 12 B/element
// HBM-bound: 32 B of traffic per element and pass.
#pragma once
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace phy
{

// ----------------------------------------------------------------------------- scan

// option "scan_mode" (Tuning::scan_single_pass): 1 = one launch per scan (decoupled look-back,
// default), 0 = three

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T, typename Op>
__device__ __forceinline__ T block_scan_exclusive(T v, Op op, T identity, T *total, T *smem /* 32 */)
{
	// exclusive scan of one value per thread over a 256-thread block
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	T inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		T o = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc = op(o, inc);
	}
	if (lane == 31) smem[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		T w = lane < (SCAN_THREADS / 32) ? smem[lane] : identity;
		T winc = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			T o = __shfl_up_sync(0xffffffffu, winc, d);
			if (lane >= d) winc = op(o, winc);
		}
		T wexc = __shfl_up_sync(0xffffffffu, winc, 1);
		if (lane == 0) wexc = identity;
		if (lane < (SCAN_THREADS / 32)) smem[lane] = wexc;
		if (lane == (SCAN_THREADS / 32) - 1) smem[16] = winc; // block total
	}
	__syncthreads();
	T exc = __shfl_up_sync(0xffffffffu, inc, 1);
	if (lane == 0) exc = identity;
	exc = op(smem[warp], exc);
	if (total) *total = smem[16];
	__syncthreads();
	return exc;
}

template <typename T, typename InF, typename Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(int64_t n, InF in, Op op, T identity, T *block_sums)
{
	__shared__ T smem[32];
	const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
	T acc = identity;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		int64_t i = base + k;
		if (i < n) acc = op(acc, in(i));
	}
	T total;
	block_scan_exclusive(acc, op, identity, &total, smem);
	if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <typename T, typename InF, typename OutF, typename Op>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(int64_t n, InF in, OutF out, Op op, T identity, const T *block_prefix, bool inclusive)
{
	__shared__ T smem[32];
	const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
	T v[SCAN_ITEMS];
	T acc = identity;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		int64_t i = base + k;
		v[k] = i < n ? in(i) : identity;
		acc = op(acc, v[k]);
	}
	T exc = block_scan_exclusive(acc, op, identity, (T *)nullptr, smem);
	T run = op(block_prefix ? block_prefix[blockIdx.x] : identity, exc);
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		int64_t i = base + k;
		T incl = op(run, v[k]);
		if (i < n) out(i, inclusive ? incl : run);
		run = incl;
	}
}

template <typename T> struct ScanPtrIn {
	const T *p;
	__device__ __forceinline__ T operator()(int64_t i) const { return p[i]; }
};
template <typename T> struct ScanPtrOut {
	T *p;
	__device__ __forceinline__ void operator()(int64_t i, T v) const { p[i] = v; }
};

// ---- single-pass scan (decoupled look-back) -----------------------------------------------
// One launch instead of reduce + scan + apply: a block takes the next tile from an atomic
// counter, publishes the tile's aggregate, then adds up its predecessors' status words — a
// warp looks at 32 of them at a time — until it meets one that already holds an inclusive
// prefix.  status[t] = epoch << 34 | flag << 32 | value (32-bit T only).  The status array
// belongs to the stream and is never cleared between scans: every launch has its own epoch and
// a word of another epoch reads as "not there yet"; the block that draws the last tile puts
// the tile counter back to zero.  (A memset per scan was a launch of its own: ~4 us each.)
constexpr unsigned long long SP_AGGREGATE = 1ull, SP_INCLUSIVE = 2ull;

template <typename T, typename InF, typename OutF, typename Op>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_single_pass_kernel(int64_t n, InF in, OutF out, Op op, T identity, bool inclusive, unsigned long long *status,
                        uint32_t *tile_counter, uint32_t epoch, const uint32_t *__restrict__ n_dev = nullptr)
{
	static_assert(sizeof(T) == 4, "status words carry 32-bit values");
	__shared__ T smem[32];
	__shared__ uint32_t s_tile;
	__shared__ T s_prefix;
	if (threadIdx.x == 0) {
		s_tile = atomicAdd(tile_counter, 1u);
		if (s_tile == gridDim.x - 1) atomicExch(tile_counter, 0u); // every block has drawn: ready for the next scan
	}
	__syncthreads();
	const uint32_t tile = s_tile;
	if (n_dev) {
		// the length is only known on the device (the grid covers its upper bound n): tiles past
		// the end leave at once — nobody looks back at them
		const int64_t have = (int64_t)*n_dev;
		if (have < n) n = have;
		if ((int64_t)tile * SCAN_TILE >= n) return;
	}
	const unsigned long long tag_agg = ((unsigned long long)epoch << 2) | SP_AGGREGATE;
	const unsigned long long tag_inc = ((unsigned long long)epoch << 2) | SP_INCLUSIVE;
	const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
	T v[SCAN_ITEMS];
	T acc = identity;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		const int64_t i = base + k;
		v[k] = i < n ? in(i) : identity;
		acc = op(acc, v[k]);
	}
	T total;
	const T exc = block_scan_exclusive(acc, op, identity, &total, smem);
	if (threadIdx.x < 32) {
		const int lane = threadIdx.x;
		T prefix = identity;
		if (tile == 0) {
			if (lane == 0) atomicExch(status, (tag_inc << 32) | (unsigned long long)(uint32_t)total);
		} else {
			if (lane == 0) atomicExch(status + tile, (tag_agg << 32) | (unsigned long long)(uint32_t)total);
			// windows of 32 predecessors, nearest first: lane l looks at tile - 1 - l
			int64_t top = (int64_t)tile - 1;
			for (;;) {
				const int64_t t = top - lane;
				unsigned long long tag = tag_inc; // before tile 0: an inclusive prefix of `identity`
				T val = identity;
				if (t >= 0) {
					unsigned long long w;
					do {
						w = *(volatile unsigned long long *)(status + t);
						tag = w >> 32;
					} while (tag != tag_agg && tag != tag_inc);
					val = (T)(uint32_t)w;
				}
				const uint32_t incl = __ballot_sync(0xffffffffu, tag == tag_inc);
				// lanes up to and including the nearest inclusive one contribute
				const int stop = incl ? __ffs(incl) - 1 : 31;
				T part = lane <= stop ? val : identity;
				// (the tree below does not keep the tiles in order: op must commute — sum, max)
#pragma unroll
				for (int d = 16; d > 0; d >>= 1) {
					const T o = __shfl_down_sync(0xffffffffu, part, d);
					part = op(o, part);
				}
				part = __shfl_sync(0xffffffffu, part, 0);
				prefix = op(part, prefix);
				if (incl) break;
				top -= 32;
			}
			if (lane == 0) atomicExch(status + tile, (tag_inc << 32) | (unsigned long long)(uint32_t)op(prefix, total));
		}
		if (lane == 0) s_prefix = prefix;
	}
	__syncthreads();
	T run = op(s_prefix, exc);
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		const int64_t i = base + k;
		const T incl = op(run, v[k]);
		if (i < n) out(i, inclusive ? incl : run);
		run = incl;
	}
}

// status words of the single-pass scans of one stream (scans on a stream run one after the other)
struct ScanState {
	DevBuf<unsigned long long> status; // [tiles] + the tile counter
	uint32_t epoch = 0;
};
class ScanStates
{
	std::mutex mu_;
	std::map<std::pair<int, cudaStream_t>, ScanState> states_;

  public:
	// status array for a scan of nblocks tiles on stream s, and the epoch of that scan
	unsigned long long *get(int nblocks, cudaStream_t s, uint32_t *epoch)
	{
		int dev = 0;
		CUDA_CHECK(cudaGetDevice(&dev));
		std::lock_guard<std::mutex> lock(mu_);
		ScanState &st = states_[{dev, s}];
		if (st.status.size() < (size_t)nblocks + 1 || st.epoch >= (1u << 30) - 1) {
			size_t cap = std::max<size_t>(st.status.size(), 4096);
			while (cap < (size_t)nblocks + 1)
				cap *= 2;
			st.status.alloc(cap, s); // (the old array goes back to this stream's cache: ordered behind its last scan)
			st.status.zero();
			st.epoch = 0;
		}
		*epoch = ++st.epoch;
		return st.status.get();
	}
	void drop(int device, cudaStream_t s)
	{
		std::lock_guard<std::mutex> lock(mu_);
		states_.erase({device, s});
	}
};
inline ScanStates g_scan_states;

// out(i, scan value); exclusive: value before element i, inclusive: including it.
template <typename T, typename InF, typename OutF, typename Op>
void device_scan(int64_t n, InF in, OutF out, Op op, T identity, bool inclusive, cudaStream_t s)
{
	if (n <= 0) return;
	const int nblocks = div_up(n, SCAN_TILE);
	if (nblocks == 1) {
		scan_apply_kernel<T><<<1, SCAN_THREADS, 0, s>>>(n, in, out, op, identity, (const T *)nullptr, inclusive);
		KERNEL_CHECK();
		return;
	}
	if (g_tuning.scan_single_pass) {
		uint32_t epoch = 0;
		unsigned long long *status = g_scan_states.get(nblocks, s, &epoch); // [0]: the tile counter, tiles from [1]
		scan_single_pass_kernel<T><<<nblocks, SCAN_THREADS, 0, s>>>(n, in, out, op, identity, inclusive, status + 1,
		                                                            reinterpret_cast<uint32_t *>(status), epoch);
		KERNEL_CHECK();
		return;
	}
	DevBuf<T> sums(nblocks, s);
	scan_reduce_kernel<T><<<nblocks, SCAN_THREADS, 0, s>>>(n, in, op, identity, sums.get());
	KERNEL_CHECK();
	// plain functors here: a lambda would make every recursion level a new instantiation
	device_scan<T>(nblocks, ScanPtrIn<T>{sums.get()}, ScanPtrOut<T>{sums.get()}, op, identity, false, s);
	scan_apply_kernel<T><<<nblocks, SCAN_THREADS, 0, s>>>(n, in, out, op, identity, sums.get(), inclusive);
	KERNEL_CHECK();
}

// The same with a length that is only known on the device: *n_dev elements, at most n_max (what
// the buffers hold and the grid covers).  Always the single-pass kernel.
template <typename T, typename InF, typename OutF, typename Op>
void device_scan_n(const uint32_t *n_dev, int64_t n_max, InF in, OutF out, Op op, T identity, bool inclusive, cudaStream_t s)
{
	if (n_max <= 0) return;
	const int nblocks = div_up(n_max, SCAN_TILE);
	uint32_t epoch = 0;
	unsigned long long *status = g_scan_states.get(nblocks, s, &epoch);
	scan_single_pass_kernel<T><<<nblocks, SCAN_THREADS, 0, s>>>(n_max, in, out, op, identity, inclusive, status + 1,
	                                                            reinterpret_cast<uint32_t *>(status), epoch, n_dev);
	KERNEL_CHECK();
}

struct OpSum {
	template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMax {
	template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};

inline void exclusive_sum_u32(const uint32_t *in, uint32_t *out, int64_t n, cudaStream_t s)
{
	device_scan<uint32_t>(n, ScanPtrIn<uint32_t>{in}, ScanPtrOut<uint32_t>{out}, OpSum(), 0u, false, s);
}

template <typename F> __global__ void for_kernel(int64_t n, F f)
{
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
		f(i);
}

// f(i) for every i in [0, n)
template <typename F> void device_for(int64_t n, F f, cudaStream_t s)
{
	if (n <= 0) return;
	int64_t blocks = (n + 255) / 256;
	if (blocks > NUM_SMS_B200 * 16) blocks = NUM_SMS_B200 * 16;
	for_kernel<<<(int)blocks, 256, 0, s>>>(n, f);
	KERNEL_CHECK();
}

// Stream compaction: for every i with pred(i), call emit(i, rank) with rank = number of
// selected elements before i.  *d_count receives the total (device pointer).
template <typename PredF, typename EmitF>
void device_select(int64_t n, PredF pred, EmitF emit, uint32_t *d_count, cudaStream_t s)
{
	if (n <= 0) {
		CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(uint32_t), s));
		return;
	}
	device_scan<uint32_t>(
		n, [pred] __device__(int64_t i) { return pred(i) ? 1u : 0u; },
		[pred, emit, d_count, n] __device__(int64_t i, uint32_t before) {
			const bool p = pred(i);
			if (p) emit(i, before);
			if (i == n - 1) *d_count = before + (p ? 1u : 0u);
		},
		OpSum(), 0u, false, s);
}

// device_select over *n_dev (<= n_max) elements.  *d_count must be zero beforehand (it stays
// untouched when there is no element).
template <typename PredF, typename EmitF>
void device_select_n(const uint32_t *n_dev, int64_t n_max, PredF pred, EmitF emit, uint32_t *d_count, cudaStream_t s)
{
	if (n_max <= 0) return;
	device_scan_n<uint32_t>(
		n_dev, n_max, [pred] __device__(int64_t i) { return pred(i) ? 1u : 0u; },
		[pred, emit, d_count, n_dev, n_max] __device__(int64_t i, uint32_t before) {
			const bool p = pred(i);
			if (p) emit(i, before);
			const int64_t n = (int64_t)*n_dev < n_max ? (int64_t)*n_dev : n_max;
			if (i == n - 1) *d_count = before + (p ? 1u : 0u);
		},
		OpSum(), 0u, false, s);
}

// ----------------------------------------------------------------------------- radix sort

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS; // 4096 pairs per tile
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;

__device__ __forceinline__ uint32_t rs_digit(uint64_t key, int shift)
{
	return (uint32_t)(key >> shift) & (RS_BINS - 1);
}

// lanes of the warp that hold the same 8-bit digit: one ballot per bit (cheaper than
// match.any, which the hardware executes as a long micro-coded loop)
__device__ __forceinline__ uint32_t warp_peers8(uint32_t d)
{
	uint32_t peers = 0xffffffffu;
#pragma unroll
	for (int b = 0; b < 8; b++) {
		const bool bit = (d & (1u << b)) != 0; // one LOP3 with predicate output
		const uint32_t bal = __ballot_sync(0xffffffffu, bit);
		peers &= bit ? bal : ~bal;
	}
	return peers;
}

// counts[d * ntiles + tile]
static __global__ void __launch_bounds__(RS_THREADS)
rs_histogram(const uint64_t *__restrict__ keys, int64_t n, int shift, int ntiles, uint32_t *__restrict__ counts)
{
	__shared__ uint32_t h[RS_BINS];
	h[threadIdx.x] = 0;
	__syncthreads();
	const int64_t base = (int64_t)blockIdx.x * RS_TILE;
	uint64_t key[RS_ITEMS];
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++) { // all loads first, then the shared-memory atomics
		const int64_t i = base + r * RS_THREADS + threadIdx.x;
		key[r] = i < n ? keys[i] : 0;
	}
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++) {
		const int64_t i = base + r * RS_THREADS + threadIdx.x;
		if (i < n) atomicAdd(&h[rs_digit(key[r], shift)], 1u);
	}
	__syncthreads();
	counts[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

struct RsSmem {
	uint64_t keys[RS_TILE];
	uint32_t vals[RS_TILE];
	uint32_t warp_cnt[RS_WARPS][RS_BINS]; // per-warp digit counts, then per-warp offsets
	uint32_t tile_start[RS_BINS];         // exclusive scan of the tile's digit counts
	uint32_t gbase[RS_BINS];              // global output offset of (digit, this tile)
	uint32_t scan_tmp[32];
};

// All digit histograms of a sort in one read of the keys: ghist[pass * 256 + digit].
constexpr int RS_MAX_PASSES = 8;

static __global__ void __launch_bounds__(RS_THREADS)
rs_global_hist(const uint64_t *__restrict__ keys, int64_t n, int bit_lo, int passes, uint32_t *__restrict__ ghist)
{
	__shared__ uint32_t h[RS_MAX_PASSES][RS_BINS];
	for (int p = 0; p < passes; p++)
		h[p][threadIdx.x] = 0;
	__syncthreads();
	for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
		const uint64_t key = keys[i];
		for (int p = 0; p < passes; p++)
			atomicAdd(&h[p][rs_digit(key, bit_lo + 8 * p)], 1u);
	}
	__syncthreads();
	for (int p = 0; p < passes; p++) {
		const uint32_t v = h[p][threadIdx.x];
		if (v) atomicAdd(&ghist[p * RS_BINS + threadIdx.x], v);
	}
}

// ghist -> exclusive prefix per pass (one block per pass)
static __global__ void __launch_bounds__(RS_THREADS) rs_digit_bases(uint32_t *__restrict__ ghist)
{
	__shared__ uint32_t tmp[32];
	uint32_t *g = ghist + blockIdx.x * RS_BINS;
	const uint32_t v = g[threadIdx.x];
	const uint32_t e = block_scan_exclusive<uint32_t>(v, OpSum(), 0u, (uint32_t *)nullptr, tmp);
	g[threadIdx.x] = e;
}

// tile status word of the decoupled look-back: flag in the top two bits, count below
constexpr unsigned long long RS_FLAG_LOCAL = 1ull << 62;     // count of this tile only
constexpr unsigned long long RS_FLAG_INCLUSIVE = 2ull << 62; // count of this tile and all before it
constexpr unsigned long long RS_VALUE_MASK = (1ull << 62) - 1;

// One radix pass.  LOOKBACK = false: offsets[d * ntiles + tile] holds the exclusive scan of
// per-tile digit counts made by rs_histogram + scan.  LOOKBACK = true ("onesweep"): offsets
// holds the 256 global digit bases of this pass; tiles are handed out in launch order by an
// atomic counter and every tile gets the number of equal digits in the tiles before it by
// looking back through the status words of its predecessors (which are already running, so
// the wait is bounded).  vals_in == nullptr: value = index.
template <bool LOOKBACK>
static __global__ void __launch_bounds__(RS_THREADS)
rs_scatter(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint64_t *__restrict__ keys_out,
           uint32_t *__restrict__ vals_out, int64_t n, int shift, int ntiles, const uint32_t *__restrict__ offsets,
           unsigned long long *status, uint32_t *tile_counter, int *err)
{
	extern __shared__ __align__(16) unsigned char rs_smem_raw[];
	RsSmem &sm = *reinterpret_cast<RsSmem *>(rs_smem_raw);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int tile = blockIdx.x;
	if (LOOKBACK) {
		if (threadIdx.x == 0) sm.scan_tmp[31] = atomicAdd(tile_counter, 1u);
		__syncthreads();
		tile = (int)sm.scan_tmp[31];
	}
	const int64_t tile_base = (int64_t)tile * RS_TILE;
	const int64_t remaining = n - tile_base;
	const int nvalid = remaining < RS_TILE ? (int)remaining : RS_TILE;

	for (int b = threadIdx.x; b < RS_WARPS * RS_BINS; b += RS_THREADS)
		(&sm.warp_cnt[0][0])[b] = 0;
	if (!LOOKBACK) sm.gbase[threadIdx.x] = offsets[(int64_t)threadIdx.x * ntiles + tile];
	__syncthreads();

	// warp-striped load: item r of lane l sits at warp_base + r*32 + l (tile order)
	uint64_t key[RS_ITEMS];
	uint32_t val[RS_ITEMS];
	uint32_t rank[RS_ITEMS];
	const int warp_base = warp * (32 * RS_ITEMS);
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++) {
		const int t = warp_base + r * 32 + lane;
		const int64_t i = tile_base + t;
		if (t < nvalid) {
			key[r] = keys_in[i];
			val[r] = vals_in ? vals_in[i] : (uint32_t)i;
		} else {
			key[r] = ~0ull; // sorts last inside the tile, never written out
			val[r] = 0;
		}
	}
	// stable rank of every item among the items of its warp with the same digit
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++) {
		const uint32_t d = rs_digit(key[r], shift);
		const uint32_t peers = warp_peers8(d);
		const uint32_t before = sm.warp_cnt[warp][d];
		__syncwarp();
		if (lane == (__ffs(peers) - 1)) sm.warp_cnt[warp][d] = before + __popc(peers);
		__syncwarp();
		rank[r] = before + __popc(peers & ((1u << lane) - 1));
	}
	__syncthreads();
	// thread d: turn per-warp counts of digit d into exclusive offsets, get the tile count
	{
		const int d = threadIdx.x;
		uint32_t run = 0;
#pragma unroll
		for (int w = 0; w < RS_WARPS; w++) {
			uint32_t c = sm.warp_cnt[w][d];
			sm.warp_cnt[w][d] = run;
			run += c;
		}
		if (LOOKBACK) {
			// publish this tile's count of digit d, then add up the predecessors' counts
			unsigned long long *mine = status + (int64_t)tile * RS_BINS + d;
			if (tile == 0) {
				atomicExch(mine, RS_FLAG_INCLUSIVE | run);
				sm.gbase[d] = offsets[d];
			} else {
				atomicExch(mine, RS_FLAG_LOCAL | run);
				unsigned long long before = 0;
				int t = tile - 1;
				long long spins = 0;
				for (;;) {
					const unsigned long long v = *(volatile unsigned long long *)(status + (int64_t)t * RS_BINS + d);
					if ((v >> 62) == 0) {
						if (++spins > (1ll << 28)) { // a bug, not a wait: bail out instead of hanging the GPU
							atomicExch(err, 1);
							break;
						}
						continue;
					}
					before += v & RS_VALUE_MASK;
					if (v & RS_FLAG_INCLUSIVE) break;
					t--;
				}
				atomicExch(mine, RS_FLAG_INCLUSIVE | (before + run));
				sm.gbase[d] = offsets[d] + (uint32_t)before;
			}
		}
		uint32_t start = block_scan_exclusive<uint32_t>(run, OpSum(), 0u, (uint32_t *)nullptr, sm.scan_tmp);
		sm.tile_start[d] = start;
	}
	__syncthreads();
	// exchange: place every pair at its position in the tile's digit-sorted order
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++) {
		const uint32_t d = rs_digit(key[r], shift);
		const uint32_t p = sm.tile_start[d] + sm.warp_cnt[warp][d] + rank[r];
		sm.keys[p] = key[r];
		sm.vals[p] = val[r];
	}
	__syncthreads();
#pragma unroll
	for (int r = 0; r < RS_ITEMS; r++) {
		const int p = r * RS_THREADS + threadIdx.x;
		if (p < nvalid) {
			const uint64_t k = sm.keys[p];
			const uint32_t d = rs_digit(k, shift);
			const int64_t o = (int64_t)sm.gbase[d] + (p - sm.tile_start[d]);
			keys_out[o] = k;
			vals_out[o] = sm.vals[p];
		}
	}
}

// Sorts n pairs by bits [bit_lo, bit_hi) of the key, stable.  Ping-pongs between
// (keys, vals) and (keys_alt, vals_alt); returns true if the result ended up in the
// alternate buffers.  vals may be nullptr on input: values start as 0..n-1 and are written
// to vals_alt/vals from the first pass on (vals must still be a valid buffer then).
// Tuning::rs_mode — 0: pick by size, 1: histogram + scan + scatter per pass, 2: onesweep (option "sort_mode")
constexpr int RS_ONESWEEP_MIN_TILES = 16384; // ~67 M pairs

// optional per-kernel timing of a sort (CUDA events on the launching stream)
struct RsProfile {
	float hist_ms = 0, scan_ms = 0, scatter_ms = 0;
	int passes = 0;
};

inline bool radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_alt, uint32_t *vals_alt, int64_t n,
                             int bit_lo, int bit_hi, bool iota_values, cudaStream_t s, RsProfile *prof = nullptr)
{
	if (n <= 0 || bit_hi <= bit_lo) return false;
	CUDA_CHECK(cudaFuncSetAttribute(rs_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
	CUDA_CHECK(cudaFuncSetAttribute(rs_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
	const int ntiles = div_up(n, RS_TILE);
	const int passes = (bit_hi - bit_lo + 7) / 8;
	if (passes > RS_MAX_PASSES) throw std::invalid_argument("radix sort: more than 8 passes");
	// Two schemes.  Look-back ("onesweep") moves 24 B per element and pass instead of 32, but
	// the first wave of resident tiles pays a serial look-back of ~0.1 ms per pass (measured
	// on B200), so it only wins once there are many waves of tiles.
	const bool onesweep = g_tuning.rs_mode == 2 || (g_tuning.rs_mode == 0 && ntiles >= RS_ONESWEEP_MIN_TILES);
	if (!onesweep) {
		DevBuf<uint32_t> counts((size_t)ntiles * RS_BINS, s);
		std::vector<cudaEvent_t> evs;
		auto mark = [&] {
			if (!prof) return;
			cudaEvent_t e;
			CUDA_CHECK(cudaEventCreate(&e));
			CUDA_CHECK(cudaEventRecord(e, s));
			evs.push_back(e);
		};
		bool flipped = false, first = true;
		for (int p = 0; p < passes; p++) {
			const int shift = bit_lo + 8 * p;
			const uint64_t *kin = flipped ? keys_alt : keys;
			const uint32_t *vin = flipped ? vals_alt : vals;
			uint64_t *kout = flipped ? keys : keys_alt;
			uint32_t *vout = flipped ? vals : vals_alt;
			mark();
			rs_histogram<<<ntiles, RS_THREADS, 0, s>>>(kin, n, shift, ntiles, counts.get());
			KERNEL_CHECK();
			mark();
			exclusive_sum_u32(counts.get(), counts.get(), (int64_t)ntiles * RS_BINS, s);
			mark();
			rs_scatter<false><<<ntiles, RS_THREADS, sizeof(RsSmem), s>>>(kin, (first && iota_values) ? nullptr : vin, kout,
			                                                             vout, n, shift, ntiles, counts.get(), nullptr,
			                                                             nullptr, nullptr);
			KERNEL_CHECK();
			mark();
			flipped = !flipped;
			first = false;
		}
		if (prof) {
			CUDA_CHECK(cudaStreamSynchronize(s));
			for (size_t k = 0; k + 3 < evs.size(); k += 4) {
				float a = 0, b = 0, c = 0;
				CUDA_CHECK(cudaEventElapsedTime(&a, evs[k], evs[k + 1]));
				CUDA_CHECK(cudaEventElapsedTime(&b, evs[k + 1], evs[k + 2]));
				CUDA_CHECK(cudaEventElapsedTime(&c, evs[k + 2], evs[k + 3]));
				prof->hist_ms += a;
				prof->scan_ms += b;
				prof->scatter_ms += c;
				prof->passes++;
			}
			for (auto e : evs)
				cudaEventDestroy(e);
		}
		return flipped;
	}
	// digit bases of every pass from one read of the keys ("onesweep")
	DevBuf<uint32_t> ghist((size_t)passes * RS_BINS, s);
	DevBuf<unsigned long long> status((size_t)ntiles * RS_BINS, s);
	DevBuf<uint32_t> ctl(2, s); // [0] tile counter, [1] error flag
	ghist.zero();
	std::vector<cudaEvent_t> evs;
	auto mark = [&] {
		if (!prof) return;
		cudaEvent_t e;
		CUDA_CHECK(cudaEventCreate(&e));
		CUDA_CHECK(cudaEventRecord(e, s));
		evs.push_back(e);
	};
	mark();
	{
		int blocks = ntiles < NUM_SMS_B200 * 8 ? ntiles : NUM_SMS_B200 * 8;
		rs_global_hist<<<blocks, RS_THREADS, 0, s>>>(keys, n, bit_lo, passes, ghist.get());
		KERNEL_CHECK();
		rs_digit_bases<<<passes, RS_THREADS, 0, s>>>(ghist.get());
		KERNEL_CHECK();
	}
	mark();
	bool flipped = false;
	bool first = true;
	for (int p = 0; p < passes; p++) {
		const int shift = bit_lo + 8 * p;
		const uint64_t *kin = flipped ? keys_alt : keys;
		const uint32_t *vin = flipped ? vals_alt : vals;
		uint64_t *kout = flipped ? keys : keys_alt;
		uint32_t *vout = flipped ? vals : vals_alt;
		status.zero();
		ctl.zero();
		mark();
		rs_scatter<true><<<ntiles, RS_THREADS, sizeof(RsSmem), s>>>(
			kin, (first && iota_values) ? nullptr : vin, kout, vout, n, shift, ntiles, ghist.get() + (size_t)p * RS_BINS,
			status.get(), ctl.get(), (int *)(ctl.get() + 1));
		KERNEL_CHECK();
		mark();
		flipped = !flipped;
		first = false;
	}
	if (prof) {
		CUDA_CHECK(cudaStreamSynchronize(s));
		float a = 0;
		CUDA_CHECK(cudaEventElapsedTime(&a, evs[0], evs[1]));
		prof->hist_ms += a;
		for (size_t k = 2; k + 1 < evs.size(); k += 2) {
			float c = 0;
			CUDA_CHECK(cudaEventElapsedTime(&c, evs[k], evs[k + 1]));
			prof->scatter_ms += c;
			prof->passes++;
		}
		for (auto e : evs)
			cudaEventDestroy(e);
	}
	return flipped;
}

} // namespace phy
