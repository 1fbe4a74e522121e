// Anchoring of all queries against the device ESA.  Replaces hot loop A of
// /root/reference/src/process.cxx:433-458.  See walk.h for why the speculative,
// chunk-parallel walk reproduces the reference's sequential walk exactly.
//
// Launch sequence for one batch of queries (all on one stream):
//   k_walk_chunks     one warp per CH-base chunk: cold walk, events, dead bitmap, exit
//   k_resolve_open    one warp per over-long match: link it to the next open match on the
//                     same diagonal or scan on cooperatively to the mismatch
//   k_open_jump       pointer jumping over those links
//   k_apply_open      final lengths / exit states of open events
//   k_bridge          one warp per chunk: from the exit state to the merge point
//   k_resolve_path    one block per query: pointer doubling from walker 0 marks the true path
//   (k_continue)      only if a give-up sits on a true path: exact serial continuation
//   k_copy_events     true events, compacted per query
//   assemble          right/left classification (max-scan for run heads) -> homologies
//   k_sort_filter     one block per query: bitonic sort by projected start in shared memory,
//                     overlap check, chaining DP, survivors
//   (general path)    a list longer than 2048, or equal starts: global radix sort by (query,
//                     start) + k_filter, or std::sort on the host for the tie case
#include "anchor_device.h"
#include "filter.h"
#include "primitives.cuh"

#include <algorithm>

namespace phy
{

namespace
{

struct Timer {
	cudaEvent_t a, b;
	cudaStream_t s;
	bool on;
	Timer(cudaStream_t st, bool enabled) : s(st), on(enabled)
	{
		if (!on) return;
		CUDA_CHECK(cudaEventCreate(&a));
		CUDA_CHECK(cudaEventCreate(&b));
		CUDA_CHECK(cudaEventRecord(a, s));
	}
	float lap()
	{
		if (!on) return 0.f;
		CUDA_CHECK(cudaEventRecord(b, s));
		CUDA_CHECK(cudaEventSynchronize(b));
		float ms = 0;
		CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
		CUDA_CHECK(cudaEventRecord(a, s));
		return ms;
	}
	~Timer()
	{
		if (!on) return;
		cudaEventDestroy(a);
		cudaEventDestroy(b);
	}
};

// ------------------------------------------------------------------ phase 1

// Walkers are latency-bound pointer chasers that diverge from the first step on.  A thread
// per walker does not work: threads of one warp that sit at different PCs are serialised and
// their memory latencies add up.  A walker therefore gets a group of WALK_LANES lanes
// (esa_search.h, "lane groups"): all of them execute the walker's scalar logic redundantly
// (same addresses: broadcast loads, merged stores) and share the byte comparisons, 32 bases
// per step.  Measured on B200 (8 x 5 Mbp): a whole warp per walker 0.41 ms, 16 lanes 0.50 ms,
// 8 lanes 0.55 ms, one thread 6.7 ms — more walkers in flight per SM do not make up for the
// warp's walkers taking turns, so the default is the whole warp.
#ifndef WALK_LANES
#define WALK_LANES 32
#endif
constexpr int WALK_THREADS = 128;
constexpr int WALKERS_PER_BLOCK = WALK_THREADS / WALK_LANES;
#ifndef WALK_MIN_BLOCKS
#define WALK_MIN_BLOCKS 16
#endif

// Walkers of divergent genomes take several times longer than those of close ones (one takes
// ~190 us at d = 0.05, a few us at d = 0.001).  With one walker per launched group a block keeps
// its slot until its slowest walker is done and the second wave of blocks starts late: the
// kernel took two "slowest walker" times.  So the groups are persistent and pull the next
// chunk from a counter; the launch-order permutation still spreads a genome's chunks (which
// are numbered consecutively) over the queue.
__device__ __forceinline__ int32_t next_walker(int *counter)
{
	int32_t w = 0;
	if (coop_lane<WALK_LANES>() == 0) w = atomicAdd(counter, 1);
	return __shfl_sync(coop_mask<WALK_LANES>(), w, coop_shift<WALK_LANES>());
}

// ctl[0]: any open match, ctl[3]: the work counter (zero at launch)
__global__ void __launch_bounds__(WALK_THREADS, WALK_MIN_BLOCKS) k_walk_chunks(WalkParams P, int *__restrict__ ctl)
{
	if (P.skip && *P.skip) return; // (ctl[0] stays 0: the kernels of phase 2 do nothing either)
	for (;;) {
		const int32_t w = next_walker(ctl + 3);
		if (w >= P.total_chunks) return;
		const int32_t g = (int32_t)(((int64_t)w * P.perm_mul) % P.total_chunks);
		walk_chunk<WALK_LANES>(P, g);
		if (P.rec[g].open) ctl[0] = 1;
	}
}

// ------------------------------------------------------------------ phase 2: open matches

// first mismatch of Q[from, to) against S[from + diag, …), warp-cooperative; returns `to`
// if there is none.  Positions at or beyond m in S count as mismatches.
__device__ int32_t warp_first_mismatch(const uint8_t *__restrict__ q, const uint8_t *__restrict__ S, int64_t diag,
                                       int32_t m, int32_t from, int32_t to)
{
	const int lane = threadIdx.x & 31;
	for (int32_t base = from; base < to; base += 32) {
		const int32_t x = base + lane;
		bool miss = false;
		if (x < to) {
			const int64_t sp = (int64_t)x + diag;
			miss = (sp >= m) || (q[x] != S[sp]);
		}
		const uint32_t b = __ballot_sync(0xffffffffu, miss);
		if (b) return base + (__ffs(b) - 1);
	}
	return to;
}

// lnk[g]: chunk whose open event ends where ours does, or -1; endq[g]: end if known.
// open chunks are also appended to open_list (count in ctl[2]) for the jumping kernel.
__global__ void k_resolve_open(WalkParams P, int *__restrict__ ctl, int32_t *__restrict__ lnk,
                               int32_t *__restrict__ endq, int32_t *__restrict__ open_list)
{
	if (!ctl[0]) return;
	const int32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (g >= P.total_chunks) return;
	const ChunkRec &r = P.rec[g];
	if (!r.open) return;
	int32_t link, end;
	open_resolve_one(P, g, warp_first_mismatch, link, end);
	if (lane == 0) {
		lnk[g] = link;
		endq[g] = end;
		open_list[atomicAdd(&ctl[2], 1)] = g;
	}
}

// resolves chains lnk -> lnk -> … -> end by pointer jumping over the open chunks only; one
// block.  Up to OPEN_SMEM open chunks (the usual case: the reference mapped onto itself is one
// chain of length / CH links) the links live in shared memory under compact indices — a round
// is then two barriers instead of dependent trips to L2; longer lists ping-pong in global
// memory (buffers indexed by chunk, touched only at open chunks).
constexpr int OPEN_SMEM = 8192;

__global__ void __launch_bounds__(1024)
k_open_jump(const int *__restrict__ ctl, const int32_t *__restrict__ open_list, int32_t *lnk_a, int32_t *end_a,
            int32_t *lnk_b, int32_t *end_b, int rounds)
{
	if (!ctl[0]) return;
	const int32_t count = ctl[2];
	if (count <= OPEN_SMEM) {
		extern __shared__ int32_t open_smem[];
		int32_t *sl = open_smem, *se = open_smem + OPEN_SMEM;
		// lnk_b is free here: chunk -> position in open_list
		for (int32_t k = threadIdx.x; k < count; k += blockDim.x)
			lnk_b[open_list[k]] = k;
		__syncthreads();
		for (int32_t k = threadIdx.x; k < count; k += blockDim.x) {
			const int32_t g = open_list[k];
			const int32_t l = lnk_a[g];
			sl[k] = l >= 0 ? lnk_b[l] : -1;
			se[k] = end_a[g];
		}
		__syncthreads();
		constexpr int PER = OPEN_SMEM / 1024;
		for (int r = 0; r < rounds; r++) {
			int32_t nl[PER], ne[PER];
			bool any = false;
#pragma unroll
			for (int u = 0; u < PER; u++) {
				const int32_t k = threadIdx.x + u * 1024;
				if (k >= count) break;
				int32_t l = sl[k], e = se[k];
				if (l >= 0) {
					const int32_t l2 = sl[l];
					if (l2 < 0) {
						e = se[l];
						l = -1;
					} else {
						l = l2;
						any = true;
					}
				}
				nl[u] = l;
				ne[u] = e;
			}
			__syncthreads(); // everybody has read this round's links
#pragma unroll
			for (int u = 0; u < PER; u++) {
				const int32_t k = threadIdx.x + u * 1024;
				if (k >= count) break;
				sl[k] = nl[u];
				se[k] = ne[u];
			}
			if (!__syncthreads_or(any)) break;
		}
		for (int32_t k = threadIdx.x; k < count; k += blockDim.x) {
			const int32_t g = open_list[k];
			const int32_t l = sl[k];
			lnk_a[g] = l >= 0 ? open_list[l] : -1;
			end_a[g] = se[k];
		}
		return;
	}
	int32_t *la = lnk_a, *ea = end_a, *lb = lnk_b, *eb = end_b;
	for (int r = 0; r < rounds; r++) {
		bool any = false;
		for (int32_t k = threadIdx.x; k < count; k += blockDim.x) {
			const int32_t g = open_list[k];
			int32_t l = la[g], e = ea[g];
			if (l >= 0) {
				const int32_t l2 = la[l];
				if (l2 < 0) {
					e = ea[l];
					l = -1;
				} else {
					l = l2;
					any = true;
				}
			}
			lb[g] = l;
			eb[g] = e;
		}
		const int more = __syncthreads_or(any);
		int32_t *t = la;
		la = lb;
		lb = t;
		t = ea;
		ea = eb;
		eb = t;
		if (!more) break;
	}
	// result must end up in the _a buffers
	if (la != lnk_a) {
		for (int32_t k = threadIdx.x; k < count; k += blockDim.x) {
			const int32_t g = open_list[k];
			lnk_a[g] = la[g];
			end_a[g] = ea[g];
		}
	}
}

__global__ void k_apply_open(WalkParams P, const int *__restrict__ any_open, const int32_t *__restrict__ lnk,
                             const int32_t *__restrict__ endq, int *__restrict__ err)
{
	if (!*any_open) return;
	const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= P.total_chunks) return;
	ChunkRec &r = P.rec[g];
	if (!r.open) return;
	if (lnk[g] >= 0 || endq[g] < 0) {
		atomicExch(err, 1);
		return;
	}
	Event &e = P.ev[(int64_t)g * P.cap_ev + r.n_events - 1];
	e.len = endq[g] - e.pos;
	r.exit.lastLen = e.len;
	r.exit.pos = e.pos + e.len + 1;
	r.open = 0;
}

// ------------------------------------------------------------------ phase 3

__global__ void __launch_bounds__(WALK_THREADS) k_bridge(WalkParams P, int *__restrict__ counter)
{
	if (P.skip && *P.skip) return;
	for (;;) {
		const int32_t w = next_walker(counter); // one walker per lane group, pulled from a counter
		if (w >= P.total_chunks) return;
		const int32_t g = (int32_t)(((int64_t)w * P.perm_mul) % P.total_chunks);
		ChunkRec &r = P.rec[g];
		const int32_t link = bridge_walk<WALK_LANES>(P, g, r.exit, 0, r.bridge_ev, P.cap_ev, P.CH, P.CAP);
		__syncwarp(coop_mask<WALK_LANES>());
		r.link = link;
	}
}

// ------------------------------------------------------------------ phase 4

// One block per query.  reach/from are per chunk; status[q] = -1 when the path reaches the
// end of the query, else the (global) chunk whose bridge is unresolved.
__global__ void __launch_bounds__(1024)
k_resolve_path(WalkParams P, int32_t *__restrict__ jump_a, int32_t *__restrict__ jump_b, uint8_t *__restrict__ reach,
               int32_t *__restrict__ from, int32_t *__restrict__ status)
{
	const int32_t qid = blockIdx.x;
	const QueryInfo qi = P.qi[qid];
	const int32_t base = qi.chunk_base, nc = qi.nchunks;
	if (nc == 0 || (P.skip && *P.skip)) {
		if (threadIdx.x == 0) status[qid] = -1;
		return;
	}
	for (int32_t k = threadIdx.x; k < nc; k += blockDim.x) {
		const ChunkRec &r = P.rec[base + k];
		jump_a[base + k] = (r.link == LINK_MERGED) ? r.link_chunk : base + k; // terminal nodes point at themselves
		reach[base + k] = (k == 0);
		from[base + k] = 0;
	}
	__syncthreads();
	int32_t *ja = jump_a, *jb = jump_b;
	for (int32_t span = 1; span < nc; span <<= 1) {
		for (int32_t k = threadIdx.x; k < nc; k += blockDim.x)
			if (reach[base + k]) reach[ja[base + k]] = 1; // benign race: all writers store 1
		__syncthreads();
		for (int32_t k = threadIdx.x; k < nc; k += blockDim.x)
			jb[base + k] = ja[ja[base + k]];
		__syncthreads();
		int32_t *t = ja;
		ja = jb;
		jb = t;
	}
	// one more marking step covers the last doubling
	for (int32_t k = threadIdx.x; k < nc; k += blockDim.x)
		if (reach[base + k]) reach[ja[base + k]] = 1;
	__syncthreads();
	for (int32_t k = threadIdx.x; k < nc; k += blockDim.x) {
		if (!reach[base + k]) continue;
		const ChunkRec &r = P.rec[base + k];
		if (r.link == LINK_MERGED)
			from[r.link_chunk] = r.link_from;
		else
			status[qid] = (r.link == LINK_END) ? -1 : base + k; // exactly one terminal node is reached
	}
}

// exact, serial continuation of an unresolved bridge that lies on a true path
__global__ void k_continue(WalkParams P, const int32_t *__restrict__ stuck, int32_t nstuck, Event *const *__restrict__ ovf,
                           const int64_t *__restrict__ ovf_cap)
{
	const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nstuck) return;
	const int32_t g = stuck[t];
	ChunkRec &r = P.rec[g];
	Event *out = ovf[t];
	for (int32_t k = 0; k < r.n_bridge; k++)
		out[k] = r.bridge_ev[k];
	r.bridge_ev = out;
	r.link = bridge_walk(P, g, r.bstate, r.n_bridge, out, ovf_cap[t], -1, 0x7fffffff);
}

// ------------------------------------------------------------------ events -> homologies

// cap: entries out / out_q hold (the number of true events is only checked against it by the
// host afterwards: nothing may be written past it meanwhile)
__global__ void k_copy_events(WalkParams P, const uint8_t *__restrict__ reach, const int32_t *__restrict__ from,
                              const uint32_t *__restrict__ offs, Event *__restrict__ out, int32_t *__restrict__ out_q,
                              uint32_t cap)
{
	// one warp per chunk, lanes stride over its events
	const int32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (P.skip && *P.skip) return;
	if (g >= P.total_chunks || !reach[g]) return;
	const ChunkRec &r = P.rec[g];
	const int32_t qid = P.chunk_query[g];
	const Event *ev = P.ev + (int64_t)g * P.cap_ev;
	const int32_t f = from[g];
	uint32_t w = offs[2 * (int64_t)g];
	for (int32_t k = f + lane; k < r.n_events && w + (uint32_t)(k - f) < cap; k += 32) {
		out[w + (k - f)] = ev[k];
		out_q[w + (k - f)] = qid;
	}
	w = offs[2 * (int64_t)g + 1];
	for (int32_t k = lane; k < r.n_bridge && w + (uint32_t)k < cap; k += 32) {
		out[w + k] = r.bridge_ev[k];
		out_q[w + k] = qid;
	}
}

__device__ __forceinline__ bool ev_first_of_query(const int32_t *evq, int64_t t)
{
	return t == 0 || evq[t] != evq[t - 1];
}

__device__ __forceinline__ bool ev_is_right(const Event *ev, const int32_t *evq, int64_t t, int32_t border)
{
	const Event prev = ev_first_of_query(evq, t) ? Event{0, 0, 0} : ev[t - 1];
	return event_is_right(prev, ev[t], border);
}

// offsets of each query's list inside an array sorted by query id
__global__ void k_query_offsets(const int32_t *__restrict__ qids, const uint32_t *__restrict__ n_ptr, int32_t nq,
                                int64_t *__restrict__ offs, int32_t *__restrict__ total_out = nullptr)
{
	const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q > nq) return;
	if (q == 0 && total_out) *total_out = (int32_t)*n_ptr; // next to the per-query results the host reads back
	int64_t lo = 0, hi = *n_ptr;
	while (lo < hi) {
		const int64_t mid = (lo + hi) >> 1;
		if (qids[mid] < q)
			lo = mid + 1;
		else
			hi = mid;
	}
	offs[q] = lo;
}

// overlap[q] == 0: no two neighbours of query q's sorted list overlap, hence no two
// elements at all — every homology is kept (keep[] is preset to 1) and the DP is skipped.
__global__ void k_filter(const int64_t *__restrict__ offs, int32_t nq, const int32_t *__restrict__ start,
                         const int32_t *__restrict__ len, int64_t *__restrict__ score, int32_t *__restrict__ pred,
                         uint8_t *__restrict__ keep, int32_t *__restrict__ heap, const int32_t *__restrict__ overlap)
{
	const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= nq || !overlap[q]) return;
	const int64_t o = offs[q];
	const int32_t h = (int32_t)(offs[q + 1] - o);
	filter_overlaps_max(start + o, len + o, h, score + o, pred + o, keep + o, heap + o);
}

// ------------------------------------------------------------------ sort + filter, one block per query
//
// The usual case — a few thousand homologies per query at most — needs no global sort:
// the block sorts (projected start, push index) in shared memory (bitonic), checks for
// equal starts (ties: the reference's unstable std::sort decides, see below) and for
// overlaps, runs the chaining DP only if something overlaps, and writes the survivors.
// Queries with more than FIN_CAP homologies, or ties, are flagged; the caller then takes
// the general path (global radix sort / host std::sort) for the batch.
constexpr int FIN_THREADS = 256;
constexpr int FIN_CAP = 2048;
constexpr int FIN_FLAG_TIES = 1, FIN_FLAG_BIG = 2;

struct FinSmem {
	uint64_t key[FIN_CAP]; // (start << 32) | push index; padding keys are all ones
	int64_t score[FIN_CAP];
	int32_t start[FIN_CAP], len[FIN_CAP], pred[FIN_CAP], heap[FIN_CAP];
	uint8_t keep[FIN_CAP];
	uint32_t scan_tmp[32];
	int flags;
};

__global__ void __launch_bounds__(FIN_THREADS)
k_sort_filter(const int64_t *__restrict__ raw_offs, const Hom *__restrict__ raw, Hom *__restrict__ fin,
              int32_t *__restrict__ fin_count, int32_t *__restrict__ fin_flags, int64_t *__restrict__ d_begin,
              int64_t *__restrict__ d_count)
{
	extern __shared__ __align__(16) unsigned char fin_smem_raw[];
	FinSmem &sm = *reinterpret_cast<FinSmem *>(fin_smem_raw);
	const int32_t q = blockIdx.x;
	const int64_t lo = raw_offs[q];
	const int32_t h = (int32_t)(raw_offs[q + 1] - lo);
	if (h > FIN_CAP) {
		if (threadIdx.x == 0) {
			fin_count[q] = 0;
			fin_flags[q] = FIN_FLAG_BIG;
			d_begin[q] = lo; // an empty list, in case the rows are built before the host has looked
			d_count[q] = 0;
		}
		return;
	}
	if (threadIdx.x == 0) sm.flags = 0;
	// power of two >= h
	int32_t npow = 1;
	while (npow < h)
		npow <<= 1;
	for (int32_t k = threadIdx.x; k < npow; k += FIN_THREADS)
		sm.key[k] = k < h ? (((uint64_t)(uint32_t)raw[lo + k].iproj << 32) | (uint32_t)k) : ~0ull;
	__syncthreads();
	for (int32_t size = 2; size <= npow; size <<= 1) {
		for (int32_t stride = size >> 1; stride > 0; stride >>= 1) {
			for (int32_t t = threadIdx.x; t < (npow >> 1); t += FIN_THREADS) {
				const int32_t i = 2 * t - (t & (stride - 1)); // lower index of the pair
				const int32_t j = i + stride;
				const bool up = (i & size) == 0;
				const uint64_t a = sm.key[i], b = sm.key[j];
				if ((a > b) == up) {
					sm.key[i] = b;
					sm.key[j] = a;
				}
			}
			__syncthreads();
		}
	}
	// sorted order -> start/len; ties and overlaps between neighbours
	int my_flags = 0, overlap = 0;
	for (int32_t k = threadIdx.x; k < h; k += FIN_THREADS) {
		const Hom x = raw[lo + (uint32_t)sm.key[k]];
		sm.start[k] = x.iproj;
		sm.len[k] = x.len;
		sm.keep[k] = 1;
		if (k > 0) {
			const Hom p = raw[lo + (uint32_t)sm.key[k - 1]];
			if (p.iproj == x.iproj) my_flags |= FIN_FLAG_TIES;
			if ((int64_t)p.iproj + p.len > x.iproj) overlap = 1;
		}
	}
	overlap = __syncthreads_or(overlap);
	my_flags = __syncthreads_or(my_flags);
	if (my_flags) {
		if (threadIdx.x == 0) {
			fin_count[q] = 0;
			fin_flags[q] = FIN_FLAG_TIES;
			d_begin[q] = lo;
			d_count[q] = 0;
		}
		return;
	}
	if (overlap) {
		if (threadIdx.x == 0) filter_overlaps_max(sm.start, sm.len, h, sm.score, sm.pred, sm.keep, sm.heap);
		__syncthreads();
	}
	// survivors, in sorted order, to fin[lo ...]
	uint32_t carry = 0;
	for (int32_t base = 0; base < h; base += FIN_THREADS) {
		const int32_t k = base + threadIdx.x;
		const uint32_t kept = (k < h && sm.keep[k]) ? 1u : 0u;
		uint32_t total;
		const uint32_t before = block_scan_exclusive<uint32_t>(kept, OpSum(), 0u, &total, sm.scan_tmp);
		if (kept) fin[lo + carry + before] = raw[lo + (uint32_t)sm.key[k]];
		carry += total;
	}
	if (threadIdx.x == 0) {
		fin_count[q] = (int32_t)carry;
		fin_flags[q] = 0;
		d_begin[q] = lo; // what the row builder reads (no host round trip on this path)
		d_count[q] = carry;
	}
}

int rounds_for(int64_t n)
{
	int r = 1;
	while ((1ll << r) < n)
		r++;
	return r + 1;
}

} // namespace

void host_sort_filter(std::vector<Hom> &list)
{
	// the very call of process.cxx:438-441: unstable, so equal starts keep libstdc++'s order
	std::sort(list.begin(), list.end(), [](const Hom &a, const Hom &b) { return a.iproj < b.iproj; });
	const int32_t h = (int32_t)list.size();
	std::vector<int32_t> start(h), len(h), pred(h), heap(h);
	std::vector<int64_t> score(h);
	std::vector<uint8_t> keep(h);
	for (int32_t k = 0; k < h; k++) {
		start[k] = list[k].iproj;
		len[k] = list[k].len;
	}
	filter_overlaps_max(start.data(), len.data(), h, score.data(), pred.data(), keep.data(), heap.data());
	size_t w = 0;
	for (int32_t k = 0; k < h; k++)
		if (keep[k]) list[w++] = list[k];
	list.resize(w);
}

void anchor_queries_device(const EsaDevice &esa, const uint8_t *d_Q, std::vector<QueryInfo> &qi, int32_t thr,
                           const AnchorOptions &opt, cudaStream_t s, AnchorResult &out, AnchorStats *stats)
{
	AnchorStats local;
	AnchorStats &ST = stats ? *stats : local;
	ST = AnchorStats();
	const int32_t nq = (int32_t)qi.size();
	if (thr < 1) throw std::invalid_argument("anchor threshold must be >= 1");
	int32_t CH = opt.chunk > 0 ? opt.chunk : 4096;
	CH = ((CH + 31) / 32) * 32;
	if (CH <= thr + 1) CH = ((thr + 2 + 31) / 32) * 32;
	int32_t CAP = opt.cap > 0 ? opt.cap : 2 * CH;
	if (CAP < CH) CAP = CH; // an open match must cover the rest of its walker's chunk
	if (CAP < thr + 1) CAP = thr + 1;

	Timer total(s, opt.timings), lap(s, opt.timings);

	// chunk geometry
	int64_t total_chunks64 = 0;
	for (auto &q : qi) {
		q.chunk_base = (int32_t)total_chunks64;
		q.nchunks = (q.qlen + CH - 1) / CH;
		total_chunks64 += q.nchunks;
	}
	if (total_chunks64 > 0x3fffffff) throw std::invalid_argument("too many chunks in one batch");
	const int32_t total_chunks = (int32_t)total_chunks64;
	ST.chunks = total_chunks;
	// (pinned: a copy from ordinary memory would make the host wait for everything queued on the
	// stream before — the index build — and for the copy itself)
	PinnedArena::Scope pinned_scope(g_pinned);
	int32_t *const chunk_query = g_pinned.take<int32_t>((size_t)total_chunks + 1);
	for (int32_t k = 0; k < nq; k++)
		std::fill(chunk_query + qi[k].chunk_base, chunk_query + qi[k].chunk_base + qi[k].nchunks, k);
	QueryInfo *const h_qi = g_pinned.take<QueryInfo>((size_t)nq + 1);
	std::copy(qi.begin(), qi.end(), h_qi);

	out.offs.assign((size_t)nq + 1, 0);
	out.raw_offs.assign((size_t)nq + 1, 0);
	out.homs.release();
	out.raw.release();
	out.d_offs.alloc((size_t)nq + 1, s);
	out.begin.assign((size_t)nq, 0);
	out.count.assign((size_t)nq, 0);
	out.d_begin.alloc((size_t)nq, s);
	out.d_count.alloc((size_t)nq, s);
	if (total_chunks == 0) {
		// (otherwise k_sort_filter or the general path write every entry)
		out.d_offs.zero();
		out.d_begin.zero();
		out.d_count.zero();
		if (opt.input_flags && opt.input_flags_ready) CUDA_CHECK(cudaStreamWaitEvent(s, opt.input_flags_ready, 0));
		if (opt.graph && opt.graph->capturing) opt.graph->launch();
		if (opt.input_flags) ST.input_flags = d2h_scalar(opt.input_flags, s);
		return;
	}

	const int32_t cap_ev = CH / (thr + 1) + 2;
	DevBuf<QueryInfo> d_qi(nq, s);
	DevBuf<int32_t> d_cq(total_chunks, s);
	CUDA_CHECK(cudaMemcpyAsync(d_qi.get(), h_qi, nq * sizeof(QueryInfo), cudaMemcpyHostToDevice, s));
	CUDA_CHECK(cudaMemcpyAsync(d_cq.get(), chunk_query, (size_t)total_chunks * sizeof(int32_t), cudaMemcpyHostToDevice, s));
	DevBuf<Event> ev((size_t)total_chunks * cap_ev, s), bev((size_t)total_chunks * cap_ev, s);
	DevBuf<uint32_t> dead((size_t)total_chunks * (CH / 32), s);
	dead.zero();
	DevBuf<ChunkRec> rec(total_chunks, s);
	// Everything the host reads back at its first stop, in one block (one copy to pinned memory):
	// ctl[0] any_open, [1] error, [2] open chunks, [3] / [4] work counters of walk / bridge,
	// [8] number of true events, [9] copy of *opt.input_flags, [10] number of raw homologies,
	// [11] number of true events the buffers hold (= [8] unless something is wrong);
	// from [CTL_INTS] on the path status of every query.
	constexpr int CTL_INTS = 16, CTL_EVENTS = 8, CTL_INPUT = 9, CTL_RAW = 10, CTL_NEFF = 11;
	DevBuf<int> ctl((size_t)CTL_INTS + nq, s);
	CUDA_CHECK(cudaMemsetAsync(ctl.get(), 0, CTL_INTS * sizeof(int), s));
	int *const flags = ctl.get();
	int32_t *const status = ctl.get() + CTL_INTS;
	int *const h_ctl = g_pinned.take<int>((size_t)CTL_INTS + nq);

	WalkParams P;
	P.esa = esa.view();
	P.Q = d_Q;
	P.qi = d_qi.get();
	P.nq = nq;
	P.thr = thr;
	P.CH = CH;
	P.CAP = CAP;
	P.cap_ev = cap_ev;
	P.total_chunks = total_chunks;
	{
		// multiplier of the launch-order permutation: near the golden ratio, coprime to the count
		auto gcd = [](int64_t a, int64_t b) {
			while (b) {
				const int64_t t = a % b;
				a = b;
				b = t;
			}
			return a;
		};
		int64_t mul = (int64_t)(0.6180339887 * total_chunks) | 1;
		while (mul > 1 && gcd(mul, total_chunks) != 1)
			mul -= 2;
		P.perm_mul = (int32_t)(mul < 1 ? 1 : mul);
	}
	P.ev = ev.get();
	P.bev = bev.get();
	P.dead = dead.get();
	P.rec = rec.get();
	P.chunk_query = d_cq.get();
	P.skip = opt.index_skip;

	// 1. cold walks
	const int walk_blocks = std::min(div_up(total_chunks, WALKERS_PER_BLOCK), NUM_SMS_B200 * WALK_MIN_BLOCKS);
	k_walk_chunks<<<walk_blocks, WALK_THREADS, 0, s>>>(P, flags);
	KERNEL_CHECK();
	ST.walk_ms = lap.lap();

	// 2. open matches
	{
		DevBuf<int32_t> lnk_a(total_chunks, s), end_a(total_chunks, s), lnk_b(total_chunks, s), end_b(total_chunks, s);
		DevBuf<int32_t> open_list(total_chunks, s);
		k_resolve_open<<<div_up((int64_t)total_chunks * 32, 256), 256, 0, s>>>(P, flags, lnk_a.get(), end_a.get(),
		                                                                       open_list.get());
		KERNEL_CHECK();
		CUDA_CHECK(cudaFuncSetAttribute(k_open_jump, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                                2 * OPEN_SMEM * (int)sizeof(int32_t)));
		k_open_jump<<<1, 1024, 2 * OPEN_SMEM * sizeof(int32_t), s>>>(flags, open_list.get(), lnk_a.get(), end_a.get(),
		                                                             lnk_b.get(), end_b.get(), rounds_for(total_chunks));
		KERNEL_CHECK();
		k_apply_open<<<div_up(total_chunks, 256), 256, 0, s>>>(P, flags, lnk_a.get(), end_a.get(), flags + 1);
		KERNEL_CHECK();
	}
	ST.open_ms = lap.lap();

	// 3. bridges
	k_bridge<<<walk_blocks, WALK_THREADS, 0, s>>>(P, flags + 4);
	KERNEL_CHECK();
	ST.bridge_ms = lap.lap();

	// 4. true path; 5. true events, compacted, then homologies; per-query sort + filter.
	// All of it is queued without the host in between: how many true events there are is only
	// known on the device (ctl[CTL_NEFF]), the buffers hold what a true path can carry at most —
	// an accepted anchor covers more than thr bases, a chunk adds its first event and its bridge's
	// last —, the scans read their length on the device.  The host looks once, at the end.  A
	// give-up that sits on a true path (rare) shows there as well: it is continued exactly and
	// serially (k_continue), the path resolved again, and everything behind it run once more.
	DevBuf<int32_t> jump_a(total_chunks, s), jump_b(total_chunks, s), from(total_chunks, s);
	DevBuf<uint8_t> reach(total_chunks, s);
	std::vector<DevBuf<Event>> overflow;
	DevBuf<uint32_t> cnt((size_t)2 * total_chunks + 1, s); // true events per (walker, bridge), scanned
	int64_t bases = 0;
	for (const auto &q : qi)
		bases += q.qlen;
	const int64_t cap_true64 = bases / (thr + 1) + 2 * (int64_t)total_chunks + 64;
	if (cap_true64 > 0x7fffffffll) throw std::invalid_argument("too many anchors possible in one batch: use smaller batches");
	const uint32_t cap_true = (uint32_t)cap_true64;
	DevBuf<Event> tev(cap_true, s);
	DevBuf<int32_t> tevq(cap_true, s);
	DevBuf<uint32_t> headcode(cap_true, s);
	DevBuf<Hom> raw(cap_true, s);
	DevBuf<int32_t> raw_q(cap_true, s);
	DevBuf<Hom> fin(cap_true, s);
	uint32_t *const d_n_eff = reinterpret_cast<uint32_t *>(ctl.get() + CTL_NEFF);
	uint32_t *const d_n_raw = reinterpret_cast<uint32_t *>(ctl.get() + CTL_RAW);
	// What the host wants to see of the filter sits in one block: raw_offs (nq + 1 int64), then
	// as int32 the survivors per query, the per-query flags and the number of raw homologies.
	DevBuf<int64_t> finrep((size_t)nq + 1 + ((size_t)2 * nq + 1 + 1) / 2, s);
	int64_t *const d_raw_offs = finrep.get();
	int32_t *const fin_count = reinterpret_cast<int32_t *>(finrep.get() + nq + 1);
	int32_t *const fin_flags = fin_count + nq;
	int64_t *const h_finrep = g_pinned.take<int64_t>(finrep.size());
	CUDA_CHECK(cudaFuncSetAttribute(k_sort_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FinSmem)));

	auto resolve = [&](bool first) {
		k_resolve_path<<<nq, 1024, 0, s>>>(P, jump_a.get(), jump_b.get(), reach.get(), from.get(), status);
		KERNEL_CHECK();
		if (first && opt.input_flags && opt.input_flags_ready) CUDA_CHECK(cudaStreamWaitEvent(s, opt.input_flags_ready, 0));
		uint32_t *c = cnt.get();
		const ChunkRec *rc = rec.get();
		const uint8_t *rh = reach.get();
		const int32_t *fr = from.get();
		const int64_t n2 = 2 * (int64_t)total_chunks;
		int *report = flags;
		const int *input_flags = opt.input_flags;
		const int *skip = P.skip;
		device_scan<uint32_t>(
			n2 + 1,
			[rc, rh, fr, n2, skip] __device__(int64_t i) -> uint32_t {
				if (i >= n2 || (skip && *skip)) return 0u;
				const int64_t g = i >> 1;
				if (!rh[g]) return 0u;
				return (i & 1) ? (uint32_t)rc[g].n_bridge : (uint32_t)(rc[g].n_events - fr[g]);
			},
			[c, n2, report, input_flags, cap_true] __device__(int64_t i, uint32_t v) {
				c[i] = v;
				if (i == n2) { // the total, and the verdict of the input validation next to it
					report[CTL_EVENTS] = (int)v;
					report[CTL_NEFF] = (int)(v < cap_true ? v : cap_true);
					report[CTL_INPUT] = input_flags ? *input_flags : 0;
					report[CTL_RAW] = 0;
				}
			},
			OpSum(), 0u, false, s);
	};
	auto assemble_and_filter = [&] {
		k_copy_events<<<div_up((int64_t)total_chunks * 32, 256), 256, 0, s>>>(P, reach.get(), from.get(), cnt.get(),
		                                                                      tev.get(), tevq.get(), cap_true);
		KERNEL_CHECK();
		// run heads by inclusive max-scan: 2(t+1)+1 for a left anchor, 2(t+1) for a query's
		// first event that extends the virtual anchor (0,0,0), 0 otherwise
		const Event *E = tev.get();
		const int32_t *EQ = tevq.get();
		uint32_t *HC = headcode.get();
		const int32_t border = esa.n;
		device_scan_n<uint32_t>(
			d_n_eff, cap_true,
			[E, EQ, border] __device__(int64_t t) -> uint32_t {
				const bool right = ev_is_right(E, EQ, t, border);
				if (!right) return 2u * (uint32_t)(t + 1) + 1u;
				return ev_first_of_query(EQ, t) ? 2u * (uint32_t)(t + 1) : 0u;
			},
			[HC] __device__(int64_t t, uint32_t v) { HC[t] = v; }, OpMax(), 0u, true, s);
		// a run ends where the next event is a left anchor or belongs to another query
		const uint32_t *NE = d_n_eff;
		auto run_end_pushed = [E, EQ, HC, border, thr, NE] __device__(int64_t t, Hom *h) -> bool {
			const bool last = (t + 1 == (int64_t)*NE) || EQ[t + 1] != EQ[t] || !ev_is_right(E, EQ, t + 1, border);
			if (!last) return false;
			const uint32_t code = HC[t];
			const int64_t first = (int64_t)(code >> 1) - 1;
			const Event *base = E + first; // run_homology indexes relative to the run's first real event
			const int32_t a = (code & 1) ? 0 : -1;
			const int32_t b = (int32_t)(t - first);
			Hom tmp;
			const bool ok = run_homology(base, a, b, thr, border, tmp);
			if (ok && h) *h = tmp;
			return ok;
		};
		// one pass: there are at most as many homologies as events
		Hom *R = raw.get();
		int32_t *RQ = raw_q.get();
		device_select_n(
			d_n_eff, cap_true, [run_end_pushed] __device__(int64_t t) { return run_end_pushed(t, nullptr); },
			[run_end_pushed, R, RQ, EQ] __device__(int64_t t, uint32_t w) {
				Hom h;
				run_end_pushed(t, &h);
				R[w] = h;
				RQ[w] = EQ[t];
			},
			d_n_raw, s);
		// per-query ranges of the raw lists, then the per-query sort + filter in shared memory
		k_query_offsets<<<div_up(nq + 1, 128), 128, 0, s>>>(raw_q.get(), d_n_raw, nq, d_raw_offs, fin_flags + nq);
		KERNEL_CHECK();
		k_sort_filter<<<nq, FIN_THREADS, sizeof(FinSmem), s>>>(d_raw_offs, raw.get(), fin.get(), fin_count, fin_flags,
		                                                       out.d_begin.get(), out.d_count.get());
		KERNEL_CHECK();
		CUDA_CHECK(cudaMemcpyAsync(h_ctl, ctl.get(), ((size_t)CTL_INTS + nq) * sizeof(int), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaMemcpyAsync(h_finrep, finrep.get(), finrep.bytes(), cudaMemcpyDeviceToHost, s));
	};

	resolve(true);
	ST.path_ms = lap.lap();
	assemble_and_filter();
	// the lists are final unless a query needs the general path or a bridge has to be continued
	// (both rare): let the caller queue its next step behind them while the host waits for the verdict
	if (opt.on_filtered && !opt.keep_raw) opt.on_filtered(fin.get(), out.d_begin.get(), out.d_count.get());
	if (opt.graph && opt.graph->capturing) opt.graph->launch(); // everything queued so far, as one graph
	CUDA_CHECK(cudaStreamSynchronize(s));
	if (opt.index_verdict_host && *opt.index_verdict_host) throw IndexNotBuilt();
	ST.input_flags = h_ctl[CTL_INPUT];
	if (ST.input_flags) return; // the caller reports what is wrong with the input
	if (h_ctl[1]) throw std::runtime_error("internal error: open match left unresolved");
	ST.open_events += h_ctl[0] ? 1 : 0;
	for (int iter = 0;; iter++) {
		std::vector<int32_t> stuck;
		for (int32_t q = 0; q < nq; q++)
			if (h_ctl[CTL_INTS + q] >= 0) stuck.push_back(h_ctl[CTL_INTS + q]);
		if (stuck.empty()) break;
		if (iter > total_chunks + 2) throw std::runtime_error("internal error: path resolution does not terminate");
		ST.unresolved += (int64_t)stuck.size();
		ST.lists_redone = 1; // what was queued behind the lists so far worked on incomplete ones
		std::vector<Event *> h_ptr;
		std::vector<int64_t> h_cap;
		for (int32_t g : stuck) {
			const QueryInfo &q = qi[chunk_query[g]];
			const int64_t capq = q.qlen / (thr + 1) + 2;
			overflow.emplace_back((size_t)capq, s);
			h_ptr.push_back(overflow.back().get());
			h_cap.push_back(capq);
		}
		DevBuf<int32_t> d_stuck(stuck.size(), s);
		DevBuf<Event *> d_ptr(stuck.size(), s);
		DevBuf<int64_t> d_cap(stuck.size(), s);
		CUDA_CHECK(cudaMemcpyAsync(d_stuck.get(), stuck.data(), stuck.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(d_ptr.get(), h_ptr.data(), stuck.size() * sizeof(Event *), cudaMemcpyHostToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(d_cap.get(), h_cap.data(), stuck.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
		k_continue<<<div_up((int64_t)stuck.size(), 32), 32, 0, s>>>(P, d_stuck.get(), (int32_t)stuck.size(), d_ptr.get(),
		                                                            d_cap.get());
		KERNEL_CHECK();
		resolve(false);
		CUDA_CHECK(cudaMemcpyAsync(h_ctl, ctl.get(), ((size_t)CTL_INTS + nq) * sizeof(int), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaStreamSynchronize(s)); // (also: the host vectors above go out of scope)
	}
	if (ST.lists_redone) {
		assemble_and_filter();
		CUDA_CHECK(cudaStreamSynchronize(s));
	}
	ST.events = (uint32_t)h_ctl[CTL_EVENTS];
	if ((uint32_t)h_ctl[CTL_EVENTS] > cap_true) throw std::runtime_error("internal error: more true events than a path can carry");
	ST.assemble_ms = lap.lap();
	uint32_t n_raw = 0;
	const int32_t *const h_fin_count = reinterpret_cast<const int32_t *>(h_finrep + nq + 1);
	const int32_t *const h_fin_flags = h_fin_count + nq;
	std::copy(h_finrep, h_finrep + nq + 1, out.raw_offs.begin());
	n_raw = (uint32_t)h_fin_flags[nq];
	bool general_path = false;
	for (int32_t q = 0; q < nq; q++)
		general_path = general_path || h_fin_flags[q] != 0;
	if (!general_path) {
		// done: the survivors of query q sit at fin[raw_offs[q] ...)
		for (int32_t q = 0; q < nq; q++) {
			out.begin[q] = out.raw_offs[q];
			out.count[q] = h_fin_count[q];
		}
		out.homs = std::move(fin); // d_begin / d_count were written by k_sort_filter
		n_raw = 0;                 // skip the general path below
	}

	// 6. sort by (query, projected start) and keep the heaviest chain per query
	if (n_raw) {
		DevBuf<uint64_t> keys(n_raw, s), keys_alt(n_raw, s);
		DevBuf<uint32_t> idx(n_raw, s), idx_alt(n_raw, s);
		{
			uint64_t *K = keys.get();
			const Hom *R = raw.get();
			const int32_t *RQ = raw_q.get();
			device_for(n_raw, [K, R, RQ] __device__(int64_t i) { K[i] = ((uint64_t)(uint32_t)RQ[i] << 32) | (uint32_t)R[i].iproj; }, s);
		}
		int qbits = 1;
		while ((1ll << qbits) < nq)
			qbits++;
		const bool fl = radix_sort_pairs(keys.get(), idx.get(), keys_alt.get(), idx_alt.get(), n_raw, 0, 32 + qbits, true, s);
		const uint64_t *KS = fl ? keys_alt.get() : keys.get();
		const uint32_t *IS = fl ? idx_alt.get() : idx.get();
		// equal (query, start) keys: the reference's std::sort is unstable there
		DevBuf<uint32_t> d_ties(1, s);
		device_select(
			n_raw, [KS] __device__(int64_t i) { return i > 0 && KS[i] == KS[i - 1]; }, [] __device__(int64_t, uint32_t) {},
			d_ties.get(), s);
		const uint32_t ties = d2h_scalar(d_ties.get(), s);
		if (!ties) {
			DevBuf<int32_t> st(n_raw, s), ln(n_raw, s), pred(n_raw, s), heap(n_raw, s);
			DevBuf<int64_t> score(n_raw, s);
			DevBuf<uint8_t> keep(n_raw, s);
			DevBuf<int32_t> overlap(nq, s);
			overlap.zero();
			{
				int32_t *ST_ = st.get(), *LN = ln.get(), *OV = overlap.get();
				uint8_t *KP_ = keep.get();
				const Hom *R = raw.get();
				device_for(n_raw, [ST_, LN, KP_, OV, R, IS, KS] __device__(int64_t i) {
					const Hom h = R[IS[i]];
					ST_[i] = h.iproj;
					LN[i] = h.len;
					KP_[i] = 1;
					if (i > 0 && (KS[i] >> 32) == (KS[i - 1] >> 32)) {
						const Hom p = R[IS[i - 1]];
						if ((int64_t)p.iproj + p.len > h.iproj) OV[KS[i] >> 32] = 1;
					}
				}, s);
			}
			k_filter<<<div_up(nq, 32), 32, 0, s>>>(d_raw_offs, nq, st.get(), ln.get(), score.get(), pred.get(),
			                                       keep.get(), heap.get(), overlap.get());
			KERNEL_CHECK();
			DevBuf<uint32_t> d_n(1, s);
			const uint8_t *KP = keep.get();
			device_select(
				n_raw, [KP] __device__(int64_t i) { return KP[i] != 0; }, [] __device__(int64_t, uint32_t) {}, d_n.get(), s);
			const uint32_t n_final = d2h_scalar(d_n.get(), s);
			out.homs.alloc(n_final, s);
			DevBuf<int32_t> fq(n_final, s);
			{
				Hom *F = out.homs.get();
				int32_t *FQ = fq.get();
				const Hom *R = raw.get();
				const int32_t *RQ = raw_q.get();
				device_select(
					n_raw, [KP] __device__(int64_t i) { return KP[i] != 0; },
					[F, FQ, R, RQ, IS] __device__(int64_t i, uint32_t w) {
						F[w] = R[IS[i]];
						FQ[w] = RQ[IS[i]];
					},
					d_n.get(), s);
			}
			k_query_offsets<<<div_up(nq + 1, 128), 128, 0, s>>>(fq.get(), d_n.get(), nq, out.d_offs.get());
			KERNEL_CHECK();
			CUDA_CHECK(cudaMemcpyAsync(out.offs.data(), out.d_offs.get(), ((size_t)nq + 1) * sizeof(int64_t),
			                           cudaMemcpyDeviceToHost, s));
			CUDA_CHECK(cudaStreamSynchronize(s));
		} else {
			// Rare: two homologies of one query start at the same reference position. Which one
			// survives in the reference depends on libstdc++'s unstable std::sort, so run that
			// very call on the push-order list (process.cxx:438-443) for the whole batch.
			ST.tie_fallback++;
			std::vector<Hom> h_raw(n_raw);
			CUDA_CHECK(cudaMemcpyAsync(h_raw.data(), raw.get(), n_raw * sizeof(Hom), cudaMemcpyDeviceToHost, s));
			CUDA_CHECK(cudaStreamSynchronize(s));
			std::vector<Hom> fin;
			for (int32_t q = 0; q < nq; q++) {
				std::vector<Hom> list(h_raw.begin() + out.raw_offs[q], h_raw.begin() + out.raw_offs[q + 1]);
				host_sort_filter(list);
				out.offs[q] = (int64_t)fin.size();
				fin.insert(fin.end(), list.begin(), list.end());
			}
			out.offs[nq] = (int64_t)fin.size();
			out.homs.alloc(fin.size(), s);
			if (!fin.empty())
				CUDA_CHECK(cudaMemcpyAsync(out.homs.get(), fin.data(), fin.size() * sizeof(Hom), cudaMemcpyHostToDevice, s));
			CUDA_CHECK(cudaMemcpyAsync(out.d_offs.get(), out.offs.data(), ((size_t)nq + 1) * sizeof(int64_t),
			                           cudaMemcpyHostToDevice, s));
			CUDA_CHECK(cudaStreamSynchronize(s));
		}
	}
	if (general_path) {
		ST.general_path++;
		for (int32_t q = 0; q < nq; q++) {
			out.begin[q] = out.offs[q];
			out.count[q] = out.offs[q + 1] - out.offs[q];
		}
		CUDA_CHECK(cudaMemcpyAsync(out.d_begin.get(), out.begin.data(), (size_t)nq * sizeof(int64_t), cudaMemcpyHostToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(out.d_count.get(), out.count.data(), (size_t)nq * sizeof(int64_t), cudaMemcpyHostToDevice, s));
		CUDA_CHECK(cudaStreamSynchronize(s));
	}
	ST.filter_ms = lap.lap();
	if (opt.keep_raw) {
		out.raw = std::move(raw);
	}
	ST.total_ms = total.lap();
}

} // namespace phy
