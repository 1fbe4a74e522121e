// C ABI of include/phylonium_b200.h: context, argument checking, host<->device copies.
// All numerical work is in esa_build.cu, anchor.cu and compare.cu.
#include "../../include/phylonium_b200.h"

#include "anchor_device.h"
#include "compare_device.h"
#include "esa_device.h"
#include "esa_search.h"
#include "primitives.cuh"
#include "staging.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <functional>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace phy;

struct phylo_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;     // the stream all work is issued on
	cudaStream_t own_stream = nullptr; // created with the context
	cudaStream_t copy_stream = nullptr; // host-to-device copies that overlap the index build
	cudaEvent_t ev_main = nullptr, ev_copy = nullptr;
	cudaEvent_t ev_q_idle = nullptr; // recorded behind the last kernel of a mapping that reads the sequences
	cudaEvent_t ev_mark[4] = {nullptr, nullptr, nullptr, nullptr}; // device timeline of process(): start, index, mapped, done
	std::string err;

	int64_t opt_chunk = 2048, opt_cap = 0, opt_kmer = -1, opt_key_chars = 0, opt_stage_threads = 0;
	bool keep_raw = false, timings = false;
	Tuning tuning; // this context's copy of the tuning options (common.cuh)

	EsaDevice esa;
	bool esa_ready = false;
	EsaTimings esa_t; // of the last build (completed by finish_index when the build was lazy)
	GraphSegment map_graph; // the mapping of a batch up to its host stop, as one graph (do_map)

	DevBuf<uint8_t> q_own;       // queries uploaded by phylo_map_queries
	const uint8_t *dQ = nullptr; // q_own or the caller's device buffer
	std::vector<QueryInfo> qi;
	uint64_t N = 0; // sequences of the last map call
	// the sequences are mapped in batches (bounded scratch memory; in phylo_process a batch is
	// mapped while the next one is still on its way over PCIe)
	struct Batch {
		uint64_t first = 0, count = 0;
		AnchorResult res;
	};
	std::vector<Batch> batches;
	std::vector<cudaEvent_t> batch_events; // phylo_process: "the copies of batch b have landed"
	cudaStream_t check_stream = nullptr;   // input validation next to the walk
	cudaEvent_t ev_check_fork = nullptr, ev_check_done = nullptr;
	bool mapped = false;
	bool last_general_path = false; // the last batch's lists went through the global sort (do_map)

	RowStore rows;
	uint64_t rows_total = 0; // 0: follow N
	uint64_t rows_first = 0;
	// sharded runs: the row stores of the other ranks (peer memory, same layout as ours).  Every
	// batch of rows is copied into all of them as soon as it is built (phylo_rows_ipc_import /
	// phylo_rows_set_peers), on a stream of its own: the exchange runs while the next batch is mapped.
	std::vector<uint32_t *> peer_rows;
	std::vector<void *> peer_ipc; // what cudaIpcOpenMemHandle returned (closed with the context)
	int peer_rank = 0;
	DevBuf<int> db_sent; // device flag of rows_push_kernel
	bool db_sent_host = false;
	bool pushed_async = false; // copy-engine pushes in flight that the mapping stream has not joined yet // ... and what the copy-engine pushes (always all planes) imply
	std::vector<cudaStream_t> push_streams; // one per peer: the copies to different peers run side by side
	std::vector<cudaEvent_t> ev_pushed;
	cudaEvent_t ev_rows = nullptr;

	// sequences of the last phylo_process / phylo_map_queries, still in q_own (phylo_process_again)
	std::vector<uint64_t> q_offs, q_lens;
	bool q_resident = false;
	HostStager stager; // sequences cross PCIe packed to 2 bits per base (staging.h)

	// phylo_ingest_*: sequences handed over one by one, from any thread, as a parser gets
	// them ready; each call packs and uploads its sequence at once
	struct Ingest {
		bool active = false;
		std::vector<uint64_t> caps;
		std::vector<uint8_t> got;
		std::vector<UploadLane> lanes;
		std::vector<uint8_t> lane_busy;
		std::mutex mu;
		std::condition_variable cv;
		std::string err; // first failure of a put (guarded by mu)
	} ingest;

	DevBuf<unsigned long long> d_subst, d_hom;
	uint64_t matN = 0;

	std::map<std::string, double> stats;
};

namespace
{

thread_local std::string g_create_error;

int fail(phylo_ctx *ctx, int code, const std::string &msg)
{
	if (ctx)
		ctx->err = msg;
	else
		g_create_error = msg;
	return code;
}

// The block cache (common.cuh) keeps freed scratch for the next call.  A caller that maps very
// differently sized inputs one after the other would pile up blocks that never fit again: once
// more than half of the device memory sits idle in the cache, give it back.
void trim_scratch_if_large(phylo_ctx *ctx)
{
	static std::atomic<size_t> limit[64] = {};
	const int dev = ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0;
	if (!limit[dev]) {
		size_t free_b = 0, total_b = 0;
		if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return;
		limit[dev] = total_b / 2;
	}
	if (g_block_cache.cached_bytes() > limit[dev]) {
		cudaStreamSynchronize(ctx->stream);
		g_block_cache.trim(ctx->device, ctx->stream);
		if (ctx->stream != ctx->own_stream) g_block_cache.trim(ctx->device, ctx->own_stream);
		cudaStreamSynchronize(ctx->copy_stream);
		g_block_cache.trim(ctx->device, ctx->copy_stream); // (old sequence buffers)
	}
}

void clear_peers(phylo_ctx *c)
{
	for (cudaStream_t st : c->push_streams)
		cudaStreamSynchronize(st);
	for (void *p : c->peer_ipc)
		if (p) cudaIpcCloseMemHandle(p);
	c->peer_ipc.clear();
	c->peer_rows.clear();
	c->db_sent.release(); // new peers (or a new store): nothing has been sent to them yet
	c->db_sent_host = false;
}

template <typename F> int guarded(phylo_ctx *ctx, F &&f)
{
	if (!ctx) return fail(nullptr, PHYLO_ERR_INVALID, "context is NULL");
	try {
		cudaError_t e = cudaSetDevice(ctx->device);
		if (e != cudaSuccess) return fail(ctx, PHYLO_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
		g_tuning = ctx->tuning; // the host code of the kernels reads the calling context's options
		f();
		trim_scratch_if_large(ctx);
		return PHYLO_OK;
	} catch (const std::invalid_argument &e) {
		return fail(ctx, PHYLO_ERR_INVALID, e.what());
	} catch (const CudaError &e) {
		cudaGetLastError();
		return fail(ctx, PHYLO_ERR_CUDA, e.what());
	} catch (const std::exception &e) {
		return fail(ctx, PHYLO_ERR_INTERNAL, e.what());
	}
}

struct WallTimer {
	cudaEvent_t a, b;
	cudaStream_t s;
	bool on; // only with option "timings": stop() synchronises
	WallTimer(cudaStream_t st, bool enabled) : s(st), on(enabled)
	{
		if (!on) return;
		CUDA_CHECK(cudaEventCreate(&a));
		CUDA_CHECK(cudaEventCreate(&b));
		CUDA_CHECK(cudaEventRecord(a, s));
	}
	float stop()
	{
		if (!on) return 0.f;
		CUDA_CHECK(cudaEventRecord(b, s));
		CUDA_CHECK(cudaEventSynchronize(b));
		float ms = 0;
		CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
		return ms;
	}
	~WallTimer()
	{
		if (!on) return;
		cudaEventDestroy(a);
		cudaEventDestroy(b);
	}
};

void record_esa_stats(phylo_ctx *c, const EsaTimings &t)
{
	auto &s = c->stats;
	s["esa.text_ms"] = t.text_ms;
	s["esa.keys_ms"] = t.keys_ms;
	s["esa.sort_ms"] = t.sort_ms;
	s["esa.refine_ms"] = t.refine_ms;
	s["esa.lcp_ms"] = t.lcp_ms;
	s["esa.cld_ms"] = t.cld_ms;
	s["esa.table_ms"] = t.table_ms;
	s["esa.total_ms"] = t.total_ms;
	s["esa.hist_ms_avg"] = t.hist_ms_avg;
	s["esa.scan_ms_avg"] = t.scan_ms_avg;
	s["esa.scatter_ms_avg"] = t.scatter_ms_avg;
	s["esa.scatter_launches"] = t.sort_passes;
	s["esa.key_chars"] = t.key_chars;
	s["esa.packed"] = t.packed ? 1 : 0;
	s["esa.dirty"] = (double)t.dirty;
	s["esa.first_pass_ms"] = t.first_pass_ms;
	s["esa.refine_rounds"] = t.refine_rounds;
	s["esa.tied"] = (double)t.tied;
	s["esa.tie_groups"] = (double)t.tie_groups;
	s["esa.kmer_k"] = c->esa.K;
	s["esa.gc_count"] = (double)c->esa.gc_count;
	s["esa.graph_instantiated"] = (double)c->esa.build_graph.instantiated;
	s["esa.graph_updated"] = (double)c->esa.build_graph.updated;
}

void record_anchor_stats(phylo_ctx *c, const AnchorStats &t)
{
	auto &s = c->stats;
	s["anchor.chunks"] = (double)t.chunks;
	s["anchor.events"] = (double)t.events;
	s["anchor.open"] = (double)t.open_events;
	s["anchor.unresolved"] = (double)t.unresolved;
	s["anchor.tie_fallback"] = (double)t.tie_fallback;
	s["anchor.general_path"] = (double)t.general_path;
	s["anchor.walk_ms"] = t.walk_ms;
	s["anchor.open_ms"] = t.open_ms;
	s["anchor.bridge_ms"] = t.bridge_ms;
	s["anchor.path_ms"] = t.path_ms;
	s["anchor.assemble_ms"] = t.assemble_ms;
	s["anchor.filter_ms"] = t.filter_ms;
	s["anchor.total_ms"] = t.total_ms;
}

// Alphabet check of every sequence plus the zero byte that must follow it (the walk and the
// cooperative comparisons rely on both).  16 bases per thread, one 128-bit load when the
// sequence start is 16-byte aligned (phylo_map_queries lays them out that way).
__global__ void __launch_bounds__(256)
k_validate_queries(const uint8_t *__restrict__ Q, const QueryInfo *__restrict__ qi, int32_t nq, int *__restrict__ bad)
{
	const int32_t k = blockIdx.y;
	if (k >= nq) return;
	const uint8_t *q = Q + qi[k].qoff;
	const int64_t len = qi[k].qlen;
	const bool aligned = (reinterpret_cast<uintptr_t>(q) & 15) == 0;
	for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i <= len; i += (int64_t)gridDim.x * blockDim.x * 16) {
		uint8_t c[16];
		if (aligned && i + 16 <= len + 1) {
			*reinterpret_cast<uint4 *>(c) = *reinterpret_cast<const uint4 *>(q + i);
		} else {
			for (int t = 0; t < 16; t++)
				c[t] = (i + t <= len) ? q[i + t] : (uint8_t)'A';
		}
		int flag = 0;
#pragma unroll
		for (int t = 0; t < 16; t++) {
			const uint8_t x = c[t];
			const bool letter = (x == 'A' || x == 'C' || x == 'G' || x == 'T' || x == '!');
			if (i + t == len)
				flag |= (x == 0) ? 0 : 2;
			else if (i + t < len)
				flag |= letter ? 0 : 1;
		}
		if (flag) atomicOr(bad, flag);
	}
}

// G/C bytes of the reference (first n bytes of S), for ranks that received the index
__global__ void k_count_gc(const uint8_t *__restrict__ S, int32_t n, unsigned long long *__restrict__ out)
{
	unsigned int gc = 0;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
		gc += ((S[i] & 'G' & 'C') == ('G' & 'C'));
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		gc += __shfl_xor_sync(0xffffffffu, gc, d);
	if ((threadIdx.x & 31) == 0 && gc) atomicAdd(out, (unsigned long long)gc);
}

__global__ void k_get_matches(EsaView e, const uint8_t *__restrict__ text, const uint64_t *__restrict__ offs,
                              const uint64_t *__restrict__ lens, uint64_t count, int use_table, int64_t *__restrict__ out)
{
	const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= count) return;
	const uint8_t *q = text + offs[k];
	const int32_t qlen = (int32_t)lens[k];
	Match mt;
	if (qlen <= 0) {
		const Interval r = esa_root(e);
		mt = Match{0, r.i, r.j, 0, -1};
	} else {
		mt = use_table ? esa_match(e, q, qlen, 0x7fffffff) : esa_match_root(e, q, qlen, 0x7fffffff);
	}
	out[3 * k + 0] = mt.l;
	out[3 * k + 1] = mt.i;
	out[3 * k + 2] = mt.j;
}

// One bound for every way an index comes into being (build, build_dev, alloc for import): the
// same as esa_build_device's, so that m = 2n + 1 plus its padding stays inside int32.
void check_ref_length(uint64_t n)
{
	if (n < 1 || 2 * n + 1 > 0x7fffffffull - 128 - 320)
		throw std::invalid_argument("reference length must be in [1, 2^30 - 225)");
}

// query_bases: how much text will be mapped on this index, if the caller knows (0 = unknown)
void do_esa_build(phylo_ctx *c, const uint8_t *d_ref, uint64_t n, uint64_t query_bases = 0)
{
	check_ref_length(n);
	c->esa_ready = false;
	c->mapped = false;
	EsaTimings t;
	t.enabled = c->timings;
	int kmer = (int)c->opt_kmer;
	if (kmer < 0 && query_bases >= 24 * n) {
		// a table one level deeper shortens every descent; its build pays off from a few
		// dozen genomes per index on (measured on B200: 100 x 5 Mbp, walk 4.4 -> 3.5 ms)
		kmer = esa_default_k((int32_t)(2 * n + 1)) + 1;
		if (kmer > 12) kmer = 12;
	}
	// lazy: the call returns with the build queued (and the reference's alphabet checked); whoever
	// uses the index next either queues its work behind the build (do_map) or waits (finish_index)
	esa_build_device(c->esa, d_ref, (int32_t)n, kmer, (int)c->opt_key_chars, c->stream, &t, true);
	c->esa_t = t;
	record_esa_stats(c, t);
	c->esa_ready = true;
}

// Waits for a lazily built index; if what the build took for granted did not hold, the index is
// built again step by step.  Returns true in that case: work queued behind the build with
// EsaDevice::skip() has done nothing and must be queued again.
bool finish_index(phylo_ctx *c)
{
	if (!c->esa.pending) return false;
	const bool rebuilt = esa_finish(c->esa, c->stream, &c->esa_t);
	record_esa_stats(c, c->esa_t);
	return rebuilt;
}

// batch b = sequences [ends[b - 1], ends[b]).  With many short sequences a batch ends on a
// multiple of 16 of them: the all-pairs stage works on tiles of 16 genomes and can then start
// on the tiles of a batch as soon as that batch is mapped.
std::vector<uint64_t> plan_batches(const uint64_t *lens, uint64_t N)
{
	const uint64_t limit = g_tuning.map_batch_bytes;
	uint64_t longest = 0;
	for (uint64_t k = 0; k < N; k++)
		longest = std::max(longest, lens[k]);
	const uint64_t align = (N > 24 && 16 * longest <= limit / 2) ? 16 : 1;
	std::vector<uint64_t> ends;
	uint64_t bytes = 0;
	for (uint64_t k = 0; k < N; k++) {
		bytes += lens[k];
		if ((bytes >= limit && (k + 1) % align == 0) || k + 1 == N) {
			ends.push_back(k + 1);
			bytes = 0;
		}
	}
	return ends;
}

void accumulate(AnchorStats &sum, const AnchorStats &st)
{
	sum.chunks += st.chunks;
	sum.events += st.events;
	sum.open_events += st.open_events;
	sum.unresolved += st.unresolved;
	sum.tie_fallback += st.tie_fallback;
	sum.general_path += st.general_path;
	sum.walk_ms += st.walk_ms;
	sum.open_ms += st.open_ms;
	sum.bridge_ms += st.bridge_ms;
	sum.path_ms += st.path_ms;
	sum.assemble_ms += st.assemble_ms;
	sum.filter_ms += st.filter_ms;
	sum.total_ms += st.total_ms;
}

// What phylo_process hangs into the mapping: its batch plan, a call before batch b is touched
// (queue later copies, make the stream wait for this batch's bytes) and one after the rows of
// batch b have been built (compare what can be compared already).
struct MapHooks {
	const std::vector<uint64_t> *ends = nullptr;
	std::function<void(size_t)> before_batch, after_batch;
	// after_batch(b) may have run on lists that turned out not to be final (do_map queues it
	// before the host has seen the mapping's last read-back): forget what it did, it is called again
	std::function<void()> redo;
	bool validated = false; // the alphabet was checked while the sequences were packed on the host
};

// copies rows [first, first + count) of this context's store into the peers' stores.
// A batch that is followed by another one goes through the copy engines, every peer on a stream
// of its own (NVSwitch gives each pair of GPUs its full bandwidth at the same time), next to the
// mapping of the next batch.  The last batch — the only one nothing can hide — is pushed by a
// kernel on the mapping stream itself: all SMs feed the links, three of five planes where that
// is enough, no extra streams and events to wait for.
void push_rows(phylo_ctx *c, uint64_t first, uint64_t count, bool last)
{
	if (c->peer_rows.empty() || !count) return;
	if (c->peer_rows.size() > 16) throw std::invalid_argument("at most 16 ranks can exchange rows");
	if (!c->db_sent.get()) {
		c->db_sent.alloc(1, c->stream);
		c->db_sent.zero();
	}
	if (last && c->tuning.push_kernel) {
		RowPeers peers;
		peers.n = (int)c->peer_rows.size();
		for (int p = 0; p < 16; p++)
			peers.ptr[p] = (p < peers.n && p != c->peer_rank) ? c->peer_rows[p] : nullptr;
		if (c->db_sent_host) { // a copy-engine push has sent D / B planes before: tell the kernel
			static const int one = 1; // (static: the copy may run later, as a node of a graph)
			CUDA_CHECK(cudaMemcpyAsync(c->db_sent.get(), &one, sizeof one, cudaMemcpyHostToDevice, c->stream));
		}
		rows_push_kernel(c->rows, peers, (int64_t)first, (int32_t)count, c->db_sent.get(), c->stream);
		return;
	}
	if (!c->ev_rows) CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_rows, cudaEventDisableTiming));
	while (c->push_streams.size() < c->peer_rows.size()) {
		cudaStream_t st;
		cudaEvent_t ev;
		CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
		CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		c->push_streams.push_back(st);
		c->ev_pushed.push_back(ev);
	}
	CUDA_CHECK(cudaEventRecord(c->ev_rows, c->stream));
	const size_t off = (size_t)first * (size_t)c->rows.genome_words();
	const size_t bytes = (size_t)count * (size_t)c->rows.genome_words() * sizeof(uint32_t);
	for (size_t p = 0; p < c->peer_rows.size(); p++) {
		if ((int)p == c->peer_rank || !c->peer_rows[p]) continue;
		CUDA_CHECK(cudaStreamWaitEvent(c->push_streams[p], c->ev_rows, 0));
		CUDA_CHECK(cudaMemcpyAsync(c->peer_rows[p] + off, c->rows.data.get() + off, bytes, cudaMemcpyDefault, c->push_streams[p]));
	}
	c->db_sent_host = true; // whole rows went over: D / B planes included
	c->pushed_async = true;
}

// whatever the caller puts on the stream next (its barrier across ranks) is behind our pushes
void join_pushes(phylo_ctx *c)
{
	if (!c->pushed_async) return;
	c->pushed_async = false;
	for (size_t p = 0; p < c->push_streams.size() && p < c->peer_rows.size(); p++) {
		if ((int)p == c->peer_rank || !c->peer_rows[p]) continue;
		CUDA_CHECK(cudaEventRecord(c->ev_pushed[p], c->push_streams[p]));
		CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_pushed[p], 0));
	}
}

void do_map(phylo_ctx *c, const uint8_t *dQ, const uint64_t *offs, const uint64_t *lens, uint64_t N, uint64_t thr,
            const MapHooks &hooks = MapHooks())
{
	if (!c->esa_ready) throw std::invalid_argument("phylo_esa_build has not been called");
	if (thr < 1 || thr > 0x3fffffffull) throw std::invalid_argument("threshold out of range");
	if (N > 0x7fff0000ull) throw std::invalid_argument("too many sequences");
	c->mapped = false;
	c->dQ = dQ;
	c->N = N;
	c->batches.clear();
	c->qi.assign((size_t)N, QueryInfo());
	for (uint64_t k = 0; k < N; k++) {
		if (lens[k] > 0x7fffff00ull) throw std::invalid_argument("sequence too long for 32-bit indices");
		c->qi[k].qoff = (int64_t)offs[k];
		c->qi[k].qlen = (int32_t)lens[k];
	}
	cudaStream_t s = c->stream;
	AnchorOptions opt;
	opt.chunk = (int32_t)c->opt_chunk;
	opt.cap = (int32_t)c->opt_cap;
	opt.keep_raw = c->keep_raw;
	opt.timings = c->timings;

	const uint64_t total = c->rows_total ? c->rows_total : N;
	const uint64_t first_row = c->rows_total ? c->rows_first : 0;
	if (first_row + N > total) throw std::invalid_argument("rows: first_row + N exceeds total_genomes");
	if ((uint64_t)c->rows.genomes != total || c->rows.n != c->esa.n) {
		if (!c->peer_rows.empty())
			throw std::invalid_argument("the row store changed size after the peers exchanged its address: "
			                            "call phylo_rows_configure and exchange the handles again");
		rows_alloc(c->rows, (int64_t)total, c->esa.n, s);
	}

	const std::vector<uint64_t> ends = hooks.ends ? *hooks.ends : plan_batches(lens, N);
	AnchorStats sum;
	float rows_ms = 0;
	c->batches.resize(ends.size());
	uint64_t b0 = 0;
	for (size_t b = 0; b < ends.size(); b++) {
		const uint64_t b1 = ends[b], cnt = b1 - b0;
		phylo_ctx::Batch &B = c->batches[b];
		B.first = b0;
		B.count = cnt;
		if (hooks.before_batch) hooks.before_batch(b);
		std::vector<QueryInfo> qi(c->qi.begin() + (size_t)b0, c->qi.begin() + (size_t)b1);
		// Everything from here to the host's one stop in the mapping — validation, walk, path,
		// lists, rows, exchange, comparison — is recorded and submitted as one graph (GraphSegment,
		// common.cuh): ~25 launches with 0.7 instead of 2.6 us between them.  Only for the first
		// batch: while the host records it the GPU is busy with the index build; before a later
		// batch it would sit idle for the whole recording (geometry, allocations, ~25 launches:
		// measured at 1000 x 5 Mbp in 9 batches, +0.5 ms per batch).  Not when the rows of this
		// batch go to the peers through the copy engines (their streams join the main one only
		// after the last batch), and not for small batches (a new shape costs an instantiation).
		const bool last_batch = b + 1 == ends.size();
		uint64_t batch_bases = 0;
		for (uint64_t k = b0; k < b1; k++)
			batch_bases += lens[k];
		const bool use_graph = (b == 0 || c->tuning.map_graph == 2) && !c->timings && !c->keep_raw && !c->last_general_path &&
		                       (c->tuning.map_graph == 2 || (c->tuning.map_graph == 1 && batch_bases >= (4u << 20))) &&
		                       (c->peer_rows.empty() || (last_batch && c->tuning.push_kernel && !c->pushed_async));
		struct GraphGuard {
			GraphSegment &g;
			~GraphGuard() { g.abandon(); }
		} graph_guard{c->map_graph};
		if (use_graph) c->map_graph.begin(s);
		opt.graph = use_graph ? &c->map_graph : nullptr;
		// the walk relies on the alphabet and on the zero byte behind every sequence; the
		// verdict is read back with the first synchronisation of the mapping
		DevBuf<QueryInfo> d_qi((size_t)cnt, s);
		DevBuf<int> bad(1, s);
		PinnedArena::Scope pinned_scope(g_pinned); // (copies from ordinary memory would wait for the stream)
		QueryInfo *const h_qi = g_pinned.take<QueryInfo>((size_t)cnt + 1);
		std::copy(qi.begin(), qi.end(), h_qi);
		CUDA_CHECK(cudaMemcpyAsync(d_qi.get(), h_qi, cnt * sizeof(QueryInfo), cudaMemcpyHostToDevice, s));
		const bool validate = !hooks.validated; // sequences packed on the host were checked there
		if (validate) {
			bad.zero();
			// ... on a stream of its own, next to the walk
			if (!c->check_stream) {
				CUDA_CHECK(cudaStreamCreateWithFlags(&c->check_stream, cudaStreamNonBlocking));
				CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_check_fork, cudaEventDisableTiming));
				CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_check_done, cudaEventDisableTiming));
			}
			CUDA_CHECK(cudaEventRecord(c->ev_check_fork, s));
			CUDA_CHECK(cudaStreamWaitEvent(c->check_stream, c->ev_check_fork, 0));
			for (uint64_t k0 = 0; k0 < cnt; k0 += 32768) {
				const int32_t part = (int32_t)(cnt - k0 < 32768 ? cnt - k0 : 32768);
				uint64_t longest = 0;
				for (uint64_t k = k0; k < k0 + (uint64_t)part; k++)
					longest = std::max<uint64_t>(longest, lens[b0 + k]);
				dim3 grid((unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (longest / 16 + 256) / 256), 4096), part);
				k_validate_queries<<<grid, 256, 0, c->check_stream>>>(dQ, d_qi.get() + k0, part, bad.get());
				KERNEL_CHECK();
			}
			CUDA_CHECK(cudaEventRecord(c->ev_check_done, c->check_stream));
		}
		opt.input_flags_ready = validate ? c->ev_check_done : nullptr;
		// whatever happens below, the mapping stream is behind the validation before d_qi and
		// bad (declared above, released after this guard) can be handed out again
		struct JoinGuard {
			cudaStream_t s;
			cudaEvent_t e;
			~JoinGuard()
			{
				if (e) cudaStreamWaitEvent(s, e, 0);
			}
		} join_guard{s, validate ? c->ev_check_done : nullptr};
		opt.input_flags = validate ? bad.get() : nullptr;
		// reference-coordinate rows for the all-pairs stage, their way to the peers and whatever
		// the caller does with a finished batch: queued as soon as the filtered lists are, while
		// the host still waits for the mapping's last read-back
		bool batch_done = false;
		auto finish_batch = [&](const Hom *homs, const int64_t *begin, const int64_t *count) {
			WallTimer wt(s, c->timings);
			rows_build(c->rows, (int64_t)(first_row + b0), dQ, d_qi.get(), (int32_t)cnt, homs, begin, count, s);
			rows_ms += wt.stop();
			push_rows(c, first_row + b0, cnt, last_batch);
			if (hooks.after_batch) hooks.after_batch(b);
			batch_done = true;
		};
		// (timed runs keep the phases apart; after a batch that needed the general path the next
		// one probably does too: no point in building its rows twice)
		if (!c->timings && !c->last_general_path)
			opt.on_filtered = finish_batch;
		else
			opt.on_filtered = nullptr;
		AnchorStats st;
		for (;;) {
			// the first batch may find the index still being built: its kernels are queued behind
			// the build and skip their work should the build's assumptions fail
			opt.index_skip = c->esa.skip();
			opt.index_verdict_host = c->esa.host_verdict();
			batch_done = false;
			struct AfterGraph { // the validation was joined inside the graph: its event is not a real one
				AnchorOptions &o;
				cudaEvent_t &guard_event;
				bool on;
				~AfterGraph()
				{
					if (!on) return;
					o.graph = nullptr;
					o.input_flags_ready = nullptr;
					guard_event = nullptr;
				}
			} after_graph{opt, join_guard.e, use_graph};
			try {
				anchor_queries_device(c->esa, dQ, qi, (int32_t)thr, opt, s, B.res, &st);
			} catch (const IndexNotBuilt &) {
				if (!finish_index(c)) throw std::runtime_error("internal error: index verdict inconsistent");
				if (batch_done && hooks.redo) hooks.redo(); // (what was queued behind the lists saw none)
				continue;
			}
			if (finish_index(c)) { // (a batch without a single base does not look at the verdict)
				if (batch_done && hooks.redo) hooks.redo();
				continue;
			}
			break;
		}
		if (st.input_flags & 1) throw std::invalid_argument("a sequence contains bytes outside {A,C,G,T,!}");
		if (st.input_flags & 2) throw std::invalid_argument("a sequence is not followed by a zero byte in the device buffer");
		accumulate(sum, st);
		c->last_general_path = st.general_path != 0;
		if (!batch_done || st.general_path || st.lists_redone) {
			if (batch_done && hooks.redo) hooks.redo();
			finish_batch(B.res.homs.get(), B.res.d_begin.get(), B.res.d_count.get());
		}
		b0 = b1;
	}
	if (!c->peer_rows.empty()) join_pushes(c);
	if (!c->ev_q_idle) CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_q_idle, cudaEventDisableTiming));
	CUDA_CHECK(cudaEventRecord(c->ev_q_idle, s));
	record_anchor_stats(c, sum);
	c->stats["rows.ms"] = rows_ms;
	c->stats["map.batches"] = (double)ends.size();
	c->mapped = true;
}

// batch that holds sequence `index`
const phylo_ctx::Batch &batch_of(const phylo_ctx *c, uint64_t index)
{
	for (const auto &B : c->batches)
		if (index >= B.first && index < B.first + B.count) return B;
	throw std::invalid_argument("sequence index out of range");
}

void do_compare(phylo_ctx *c, int flags, int rank, int world, unsigned long long *d_subst,
                unsigned long long *d_hom)
{
	if (!c->mapped) throw std::invalid_argument("phylo_map_queries has not been called");
	if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad tile rank/world");
	const uint64_t total = c->rows_total ? c->rows_total : c->N;
	WallTimer wt(c->stream, c->timings);
	compare_all_device(c->rows, (int64_t)total, (flags & PHYLO_FLAG_COMPLETE_DELETION) != 0, rank, world, d_subst, d_hom,
	                   c->stream);
	c->stats["compare.ms"] = wt.stop();
	c->matN = total;
}

void ensure_matrix(phylo_ctx *c, uint64_t total)
{
	if (c->d_subst.size() != total * total) {
		c->d_subst.alloc((size_t)(total * total), c->stream);
		c->d_hom.alloc((size_t)(total * total), c->stream);
	}
}

} // namespace

extern "C" {

const char *phylo_version(void)
{
	return "phylonium_b200 0.1 (pipeline of phylonium 1.7)";
}

int phylo_ctx_create(int device, phylo_ctx **out)
{
	if (!out) return fail(nullptr, PHYLO_ERR_INVALID, "out is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		cudaGetLastError();
		return fail(nullptr, PHYLO_ERR_CUDA,
		            std::string("no usable CUDA device (this library has no CPU path): ") +
		                (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
	}
	if (device < 0) {
		e = cudaGetDevice(&device);
		if (e != cudaSuccess) return fail(nullptr, PHYLO_ERR_CUDA, cudaGetErrorString(e));
	}
	if (device >= count) return fail(nullptr, PHYLO_ERR_INVALID, "device ordinal out of range");
	e = cudaSetDevice(device);
	if (e != cudaSuccess) return fail(nullptr, PHYLO_ERR_CUDA, cudaGetErrorString(e));
	auto *c = new phylo_ctx;
	c->device = device;
	e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) {
		delete c;
		return fail(nullptr, PHYLO_ERR_CUDA, cudaGetErrorString(e));
	}
	c->stream = c->own_stream;
	if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess) {
		cudaStreamDestroy(c->own_stream);
		delete c;
		return fail(nullptr, PHYLO_ERR_CUDA, "could not create the copy stream");
	}
	*out = c;
	// PHYLO_B200_OPTIONS="key=value,key=value": options for every context of the process (A/B
	// measurements through programs that do not pass them on); unknown keys are ignored here
	if (const char *env = getenv("PHYLO_B200_OPTIONS")) {
		std::string all(env);
		size_t at = 0;
		while (at < all.size()) {
			size_t end = all.find(',', at);
			if (end == std::string::npos) end = all.size();
			const std::string kv = all.substr(at, end - at);
			const size_t eq = kv.find('=');
			if (eq != std::string::npos) phylo_set_option(c, kv.substr(0, eq).c_str(), atoll(kv.c_str() + eq + 1));
			at = end + 1;
		}
		c->err.clear();
	}
	return PHYLO_OK;
}

void phylo_ctx_destroy(phylo_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	c->esa.release();
	c->esa.destroy_side();
	c->map_graph.destroy();
	if (c->check_stream) {
		cudaStreamSynchronize(c->check_stream);
		cudaStreamDestroy(c->check_stream);
		cudaEventDestroy(c->ev_check_fork);
		cudaEventDestroy(c->ev_check_done);
	}
	c->stager.release();
	for (auto &l : c->ingest.lanes)
		l.destroy();
	c->ingest.lanes.clear();
	clear_peers(c);
	for (cudaStream_t st : c->push_streams) {
		cudaStreamSynchronize(st);
		cudaStreamDestroy(st);
	}
	for (cudaEvent_t ev : c->ev_pushed)
		cudaEventDestroy(ev);
	if (c->ev_rows) cudaEventDestroy(c->ev_rows);
	c->q_own.release();
	c->batches.clear();
	for (cudaEvent_t e : c->batch_events)
		cudaEventDestroy(e);
	c->batch_events.clear();
	c->rows.data.release();
	c->d_subst.release();
	c->d_hom.release();
	cudaStreamSynchronize(c->stream);
	cudaStreamSynchronize(c->copy_stream);
	// blocks cached for this context's streams go back to the driver
	g_scan_states.drop(c->device, c->own_stream);
	g_block_cache.trim(c->device, c->own_stream);
	if (c->stream != c->own_stream) {
		g_scan_states.drop(c->device, c->stream);
		g_block_cache.trim(c->device, c->stream);
	}
	g_block_cache.trim(c->device, c->copy_stream);
	cudaStreamDestroy(c->copy_stream);
	cudaEventDestroy(c->ev_main);
	cudaEventDestroy(c->ev_copy);
	for (cudaEvent_t e : c->ev_mark)
		if (e) cudaEventDestroy(e);
	if (c->ev_q_idle) cudaEventDestroy(c->ev_q_idle);
	cudaStreamDestroy(c->own_stream);
	delete c;
}

int phylo_set_stream(phylo_ctx *c, void *stream)
{
	return guarded(c, [&] {
		CUDA_CHECK(cudaStreamSynchronize(c->stream));
		if (c->stream != c->own_stream) {
			g_scan_states.drop(c->device, c->stream);
			g_block_cache.trim(c->device, c->stream);
		}
		c->stream = stream ? (cudaStream_t)stream : c->own_stream;
	});
}

const char *phylo_last_error(const phylo_ctx *c)
{
	return c ? c->err.c_str() : g_create_error.c_str();
}

int phylo_set_option(phylo_ctx *c, const char *key, int64_t value)
{
	return guarded(c, [&] {
		const std::string k = key ? key : "";
		if (k == "chunk") {
			if (value < 32 || value > (1 << 24)) throw std::invalid_argument("chunk must be in [32, 2^24]");
			c->opt_chunk = value;
		} else if (k == "cap") {
			if (value < 0) throw std::invalid_argument("cap must be >= 0");
			c->opt_cap = value;
		} else if (k == "kmer_k") {
			if (value < -1 || value > 12) throw std::invalid_argument("kmer_k must be in [-1, 12]");
			c->opt_kmer = value;
		} else if (k == "sort_mode") {
			if (value < 0 || value > 2) throw std::invalid_argument("sort_mode must be 0, 1 or 2");
			c->tuning.rs_mode = (int)value;
		} else if (k == "sort_path") {
			if (value < 0 || value > 2) throw std::invalid_argument("sort_path must be 0, 1 or 2");
			c->tuning.sort_path = (int)value;
		} else if (k == "scan_mode") {
			c->tuning.scan_single_pass = value != 0;
		} else if (k == "map_batch_bytes") {
			if (value < 1) throw std::invalid_argument("map_batch_bytes must be >= 1");
			c->tuning.map_batch_bytes = (uint64_t)value;
		} else if (k == "table_direct") {
			if (value < 0 || value > 2) throw std::invalid_argument("table_direct must be 0, 1 or 2");
			c->tuning.table_direct = (int)value;
		} else if (k == "key_chars") {
			if (value < 0 || value > 21) throw std::invalid_argument("key_chars must be in [0, 21]");
			c->opt_key_chars = value;
		} else if (k == "push_kernel") {
			c->tuning.push_kernel = value != 0;
		} else if (k == "esa_speculative") {
			c->tuning.esa_speculative = value != 0;
		} else if (k == "esa_graph") {
			c->tuning.esa_graph = value != 0;
		} else if (k == "map_graph") {
			if (value < 0 || value > 2) throw std::invalid_argument("map_graph must be 0, 1 or 2");
			c->tuning.map_graph = (int)value;
		} else if (k == "compare_path") {
			if (value < 0 || value > 1) throw std::invalid_argument("compare_path must be 0 or 1");
			c->tuning.compare_path = (int)value;
		} else if (k == "upload_raw") {
			if (value < -1 || value > 1) throw std::invalid_argument("upload_raw must be -1, 0 or 1");
			c->tuning.upload_raw = (int)value;
		} else if (k == "stage_threads") {
			if (value < 0 || value > 64) throw std::invalid_argument("stage_threads must be in [0, 64]");
			c->opt_stage_threads = value;
		} else if (k == "keep_raw") {
			c->keep_raw = value != 0;
		} else if (k == "timings") {
			c->timings = value != 0;
		} else {
			throw std::invalid_argument("unknown option: " + k);
		}
	});
}

int phylo_get_stat(const phylo_ctx *c, const char *key, double *out)
{
	if (!c || !key || !out) return PHYLO_ERR_INVALID;
	if (std::string(key) == "launches") {
		*out = (double)g_kernel_launches.load(); // process-wide count of kernel launches so far
		return PHYLO_OK;
	}
	if (c->esa.pending && !strncmp(key, "esa.", 4) && strcmp(key, "esa.gc_count") && strcmp(key, "esa.kmer_k")) {
		// figures the build itself produces: wait for it (gc_count is known as soon as the build is queued)
		auto *m = const_cast<phylo_ctx *>(c);
		const int rc = guarded(m, [&] { finish_index(m); });
		if (rc != PHYLO_OK) return rc;
	}
	auto it = c->stats.find(key);
	*out = it == c->stats.end() ? -1.0 : it->second;
	return PHYLO_OK;
}

/* src/sequence.cxx:152-165: a byte counts as G or C iff it has both bits of 'G' & 'C' */
double phylo_gc_content(const char *seq, uint64_t n)
{
	uint64_t gc = 0;
	for (uint64_t i = 0; i < n; i++)
		gc += ((seq[i] & 'G' & 'C') == ('G' & 'C'));
	return (double)gc / (double)n;
}

namespace
{
/* src/process.cxx:103-125 */
uint64_t binomial(uint64_t n, uint64_t k)
{
	if (n == 0 || k > n) return 0;
	if (k == 0 || k == n) return 1;
	if (k > n - k) k = n - k;
	uint64_t r = 1;
	for (uint64_t i = 1; i <= k; i++) {
		r *= n - k + i;
		r /= i;
	}
	return r;
}

/* src/process.cxx:140-161, same operation order so the doubles agree bit for bit */
double shuprop(uint64_t x, double p, uint64_t l)
{
	const double xx = (double)x, ll = (double)l;
	double s = 0.0;
	for (uint64_t k = 0; k <= x; k++) {
		const double kk = (double)k;
		const double t = pow(p, kk) * pow(0.5 - p, xx - kk);
		s += pow(2, xx) * (t * pow(1 - t, ll)) * (double)binomial(x, k);
		if (s >= 1.0) {
			s = 1.0;
			break;
		}
	}
	return s;
}
} // namespace

int phylo_host_pack_2bit(const char *seq, uint64_t n, uint8_t *packed, uint32_t *bangs, uint32_t cap, uint32_t *nbangs)
{
	uint32_t nb = 0;
	const int bad = pack_2bit(reinterpret_cast<const uint8_t *>(seq), (size_t)n, packed, bangs, cap, &nb);
	if (nbangs) *nbangs = nb;
	return bad;
}

uint64_t phylo_min_anchor_length(double p, double gc, uint64_t l)
{
	uint64_t x = 1;
	while (shuprop(x, gc / 2, l) < 1 - p)
		x++;
	return x;
}

int phylo_esa_build(phylo_ctx *c, const char *ref, uint64_t n)
{
	return guarded(c, [&] {
		if (!ref) throw std::invalid_argument("ref is NULL");
		check_ref_length(n);
		DevBuf<uint8_t> d(n, c->stream);
		CUDA_CHECK(cudaMemcpyAsync(d.get(), ref, n, cudaMemcpyHostToDevice, c->stream));
		do_esa_build(c, d.get(), n);
	});
}

int phylo_esa_build_dev(phylo_ctx *c, const void *d_ref, uint64_t n)
{
	return guarded(c, [&] {
		if (!d_ref) throw std::invalid_argument("d_ref is NULL");
		do_esa_build(c, (const uint8_t *)d_ref, n);
	});
}

int phylo_esa_size(const phylo_ctx *c, uint64_t *m)
{
	if (!c || !m) return PHYLO_ERR_INVALID;
	*m = (uint64_t)c->esa.m;
	return PHYLO_OK;
}

int phylo_esa_get_arrays(const phylo_ctx *cc, int64_t *SA, int64_t *LCP, int64_t *CLD, char *FVC, char *S)
{
	auto *c = const_cast<phylo_ctx *>(cc);
	return guarded(c, [&] {
		if (!c->esa_ready) throw std::invalid_argument("no index");
		finish_index(c);
		const size_t m = (size_t)c->esa.m;
		cudaStream_t s = c->stream;
		std::vector<int32_t> tmp(m + 1);
		auto fetch = [&](const int32_t *d, size_t cnt, int64_t *dst) {
			CUDA_CHECK(cudaMemcpyAsync(tmp.data(), d, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
			CUDA_CHECK(cudaStreamSynchronize(s));
			for (size_t i = 0; i < cnt; i++)
				dst[i] = tmp[i];
		};
		if (SA) fetch(c->esa.SA.get(), m, SA);
		if (LCP) fetch(c->esa.LCP.get(), m + 1, LCP);
		if (CLD) fetch(c->esa.CLD.get(), m + 1, CLD);
		if (FVC) CUDA_CHECK(cudaMemcpyAsync(FVC, c->esa.FVC.get(), m, cudaMemcpyDeviceToHost, s));
		if (S) CUDA_CHECK(cudaMemcpyAsync(S, c->esa.S.get(), m, cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaStreamSynchronize(s));
	});
}

int phylo_esa_get_matches(phylo_ctx *c, const char *text, const uint64_t *offs, const uint64_t *lens, uint64_t count,
                          int use_table, int64_t *out)
{
	return guarded(c, [&] {
		if (!c->esa_ready) throw std::invalid_argument("no index");
		finish_index(c);
		if (!count) return;
		if (!text || !offs || !lens || !out) throw std::invalid_argument("NULL argument");
		uint64_t extent = 0;
		for (uint64_t k = 0; k < count; k++) {
			if (lens[k] > 0x7fffff00ull) throw std::invalid_argument("string too long");
			extent = std::max(extent, offs[k] + lens[k]);
		}
		cudaStream_t s = c->stream;
		DevBuf<uint8_t> d_text(extent + 64, s);
		d_text.zero();
		DevBuf<uint64_t> d_offs(count, s), d_lens(count, s);
		DevBuf<int64_t> d_out(3 * count, s);
		CUDA_CHECK(cudaMemcpyAsync(d_text.get(), text, extent, cudaMemcpyHostToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(d_offs.get(), offs, count * 8, cudaMemcpyHostToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(d_lens.get(), lens, count * 8, cudaMemcpyHostToDevice, s));
		k_get_matches<<<div_up((int64_t)count, 64), 64, 0, s>>>(c->esa.view(), d_text.get(), d_offs.get(), d_lens.get(),
		                                                       count, use_table, d_out.get());
		KERNEL_CHECK();
		CUDA_CHECK(cudaMemcpyAsync(out, d_out.get(), 3 * count * 8, cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaStreamSynchronize(s));
	});
}

int phylo_map_queries_dev(phylo_ctx *c, const void *d_queries, const uint64_t *offs, const uint64_t *lens, uint64_t N,
                          uint64_t threshold)
{
	return guarded(c, [&] {
		if (N && (!d_queries || !offs || !lens)) throw std::invalid_argument("NULL argument");
		c->q_resident = false;
		c->q_own.release();
		do_map(c, (const uint8_t *)d_queries, offs, lens, N, threshold);
	});
}

int phylo_homology_counts(const phylo_ctx *c, uint64_t *counts, int raw)
{
	if (!c || !counts) return PHYLO_ERR_INVALID;
	if (!c->mapped) return PHYLO_ERR_INVALID;
	for (const auto &B : c->batches)
		for (uint64_t k = 0; k < B.count; k++)
			counts[B.first + k] =
				raw ? (uint64_t)(B.res.raw_offs[k + 1] - B.res.raw_offs[k]) : (uint64_t)B.res.count[k];
	return PHYLO_OK;
}

int phylo_get_homologies(const phylo_ctx *cc, uint64_t index, int raw, phylo_homology *out, uint64_t cap,
                         uint64_t *written)
{
	auto *c = const_cast<phylo_ctx *>(cc);
	return guarded(c, [&] {
		if (!c->mapped) throw std::invalid_argument("phylo_map_queries has not been called");
		if (index >= c->N) throw std::invalid_argument("sequence index out of range");
		if (raw && !c->keep_raw) throw std::invalid_argument("raw lists need option keep_raw");
		const phylo_ctx::Batch &B = batch_of(c, index);
		const uint64_t k = index - B.first;
		const Hom *src = raw ? B.res.raw.get() : B.res.homs.get();
		const uint64_t first = (uint64_t)(raw ? B.res.raw_offs[k] : B.res.begin[k]);
		const uint64_t cnt = raw ? (uint64_t)(B.res.raw_offs[k + 1] - B.res.raw_offs[k]) : (uint64_t)B.res.count[k];
		if (written) *written = cnt;
		const uint64_t take = cnt < cap ? cnt : cap;
		if (!take) return;
		if (!out) throw std::invalid_argument("out is NULL");
		std::vector<Hom> h(take);
		CUDA_CHECK(cudaMemcpyAsync(h.data(), src + first, take * sizeof(Hom), cudaMemcpyDeviceToHost, c->stream));
		CUDA_CHECK(cudaStreamSynchronize(c->stream));
		for (uint64_t k = 0; k < take; k++) {
			out[k].direction = h[k].dir;
			out[k].index_reference = h[k].iref;
			out[k].index_reference_projected = h[k].iproj;
			out[k].index_query = h[k].iq;
			out[k].length = h[k].len;
		}
	});
}

int phylo_compare_all(phylo_ctx *c, int flags, uint64_t *subst, uint64_t *homologs)
{
	return guarded(c, [&] {
		if (!subst || !homologs) throw std::invalid_argument("NULL argument");
		if (!c->mapped) throw std::invalid_argument("phylo_map_queries has not been called");
		const uint64_t total = c->rows_total ? c->rows_total : c->N;
		ensure_matrix(c, total);
		do_compare(c, flags, 0, 1, c->d_subst.get(), c->d_hom.get());
		const size_t bytes = (size_t)(total * total) * sizeof(uint64_t);
		CUDA_CHECK(cudaMemcpyAsync(subst, c->d_subst.get(), bytes, cudaMemcpyDeviceToHost, c->stream));
		CUDA_CHECK(cudaMemcpyAsync(homologs, c->d_hom.get(), bytes, cudaMemcpyDeviceToHost, c->stream));
		CUDA_CHECK(cudaStreamSynchronize(c->stream));
	});
}

int phylo_compare_all_dev(phylo_ctx *c, int flags, void *d_subst, void *d_homologs)
{
	return phylo_compare_tiles_dev(c, flags, 0, 1, d_subst, d_homologs);
}

int phylo_compare_tiles_dev(phylo_ctx *c, int flags, int rank, int world, void *d_subst, void *d_homologs)
{
	return guarded(c, [&] {
		if (!d_subst || !d_homologs) throw std::invalid_argument("NULL argument");
		do_compare(c, flags, rank, world, (unsigned long long *)d_subst, (unsigned long long *)d_homologs);
		if (world > 1) {
			// one rank's share of the tiles: not a matrix phylo_estimate could work on
			c->matN = 0;
			return;
		}
		const uint64_t total = c->matN;
		ensure_matrix(c, total);
		const size_t bytes = (size_t)(total * total) * sizeof(uint64_t);
		CUDA_CHECK(cudaMemcpyAsync(c->d_subst.get(), d_subst, bytes, cudaMemcpyDeviceToDevice, c->stream));
		CUDA_CHECK(cudaMemcpyAsync(c->d_hom.get(), d_homologs, bytes, cudaMemcpyDeviceToDevice, c->stream));
	});
}

int phylo_core_sites(phylo_ctx *c, uint32_t *core, uint32_t *border, uint32_t *seg, uint64_t *words)
{
	return guarded(c, [&] {
		if (!words) throw std::invalid_argument("NULL argument");
		if (!c->mapped) throw std::invalid_argument("phylo_map_queries has not been called");
		const uint64_t W = (uint64_t)c->rows.W;
		if (!core && !border && !seg) {
			*words = W;
			return;
		}
		if (!core || !border || !seg) throw std::invalid_argument("NULL argument");
		if (*words < W) throw std::invalid_argument("phylo_core_sites: bitmaps too short");
		if (c->rows_total) throw std::invalid_argument("phylo_core_sites works on an unsharded context");
		*words = W;
		cudaStream_t s = c->stream;
		DevBuf<uint32_t> d_core(W, s), d_border(W, s), d_seg(W, s);
		d_border.zero();
		core_sites_device(c->rows, (int64_t)c->N, d_core.get(), d_seg.get(), s);
		for (const auto &B : c->batches)
			hom_borders_device(B.res.homs.get(), B.res.d_begin.get(), B.res.d_count.get(), (int32_t)B.count, d_border.get(), s);
		CUDA_CHECK(cudaMemcpyAsync(core, d_core.get(), W * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaMemcpyAsync(border, d_border.get(), W * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaMemcpyAsync(seg, d_seg.get(), W * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaStreamSynchronize(s));
	});
}

int phylo_estimate(phylo_ctx *c, int kind, double *dist)
{
	return guarded(c, [&] {
		if (!dist) throw std::invalid_argument("dist is NULL");
		if (!c->matN) throw std::invalid_argument("no full matrix in this context (none computed yet, or only one rank's tiles)");
		if (kind < 0 || kind > 2) throw std::invalid_argument("unknown estimator");
		const uint64_t n2 = c->matN * c->matN;
		DevBuf<double> d(n2, c->stream);
		estimate_device(c->d_subst.get(), c->d_hom.get(), (int64_t)c->matN, kind, d.get(), c->stream);
		CUDA_CHECK(cudaMemcpyAsync(dist, d.get(), n2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_CHECK(cudaStreamSynchronize(c->stream));
	});
}

namespace
{

// index + threshold + mapping + all pairs on sequences that are already in (or on their way
// into) the context's device buffer; hooks.ends / before_batch as for do_map.  The all-pairs
// stage runs tile column by tile column as the batches get mapped, so that what is left to do
// after the last sequence has arrived is the last batch's share only.
void process_resident(phylo_ctx *c, uint64_t N, uint64_t ref_index, int flags, MapHooks hooks, uint64_t *subst,
                      uint64_t *homologs)
{
	const auto t_enter = std::chrono::steady_clock::now();
	auto since = [&](const char *key) { // host clock since the call began, as a statistic
		c->stats[key] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
	};
	const uint8_t *dq = c->q_own.get();
	const uint64_t *offs = c->q_offs.data(), *lens = c->q_lens.data();
	cudaStream_t s = c->stream;
	uint64_t query_bases = 0;
	for (uint64_t k = 0; k < N; k++)
		query_bases += lens[k];
	if (!c->ev_mark[0])
		for (auto &e : c->ev_mark)
			CUDA_CHECK(cudaEventCreate(&e));
	CUDA_CHECK(cudaEventRecord(c->ev_mark[0], s));
	do_esa_build(c, dq + offs[ref_index], lens[ref_index], query_bases);
	CUDA_CHECK(cudaEventRecord(c->ev_mark[1], s));
	// process.cxx:416-417; the G/C count comes out of the text kernel, the division and the
	// threshold search are the reference's double arithmetic on the host
	const double gc = (double)c->esa.gc_count / (double)lens[ref_index];
	const uint64_t thr = phylo_min_anchor_length(0.025, gc, 2 * lens[ref_index] + 1);
	c->stats["threshold"] = (double)thr;
	since("process.host_index_done_ms");

	const uint64_t tot = c->rows_total ? c->rows_total : N;
	ensure_matrix(c, tot);
	const std::vector<uint64_t> ends = hooks.ends ? *hooks.ends : plan_batches(lens, N);
	hooks.ends = &ends;
	const bool complete_deletion = (flags & PHYLO_FLAG_COMPLETE_DELETION) != 0;
	const bool incremental = !complete_deletion && !c->rows_total; // complete deletion needs every row first
	const int64_t CT = compare_tile_side((int64_t)tot);
	const int64_t tiles_side = ((int64_t)tot + CT - 1) / CT;
	int64_t tiles_done = 0;
	float compare_ms = 0;
	bool mirrored = false;
	hooks.after_batch = [&](size_t b) {
		if (!incremental) return;
		const bool final = b + 1 == ends.size(); // the last batch also mirrors the triangle
		const int64_t ready = final ? tiles_side : (int64_t)ends[b] / CT;
		if (ready <= tiles_done && !final) return;
		WallTimer wt(s, c->timings);
		compare_all_device(c->rows, (int64_t)tot, false, 0, 1, c->d_subst.get(), c->d_hom.get(), s, tiles_done, ready,
		                   tiles_done == 0, final);
		compare_ms += wt.stop();
		tiles_done = ready;
		mirrored = final;
	};
	hooks.redo = [&] { // (the next comparison starts over: it clears the matrix first)
		tiles_done = 0;
		mirrored = false;
	};
	do_map(c, dq, offs, lens, N, thr, hooks);
	since("process.host_map_done_ms");
	if (!mirrored) { // (complete deletion, a sharded row store, or no sequence at all)
		WallTimer wt(s, c->timings);
		compare_all_device(c->rows, (int64_t)tot, complete_deletion, 0, 1, c->d_subst.get(), c->d_hom.get(), s, tiles_done,
		                   -1, tiles_done == 0, true);
		compare_ms += wt.stop();
	}
	c->stats["compare.ms"] = compare_ms;
	c->stats["compare.increments"] = (double)ends.size();
	c->matN = tot;
	CUDA_CHECK(cudaEventRecord(c->ev_mark[2], s));
	const size_t bytes = (size_t)(tot * tot) * sizeof(uint64_t);
	if (bytes <= (512u << 10)) {
		// small matrices through pinned memory: two queued copies and one wait instead of two
		// copies that each block the host
		PinnedArena::Scope pinned_scope(g_pinned);
		uint64_t *h = g_pinned.take<uint64_t>((size_t)(2 * tot * tot));
		CUDA_CHECK(cudaMemcpyAsync(h, c->d_subst.get(), bytes, cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaMemcpyAsync(h + tot * tot, c->d_hom.get(), bytes, cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaStreamSynchronize(s));
		memcpy(subst, h, bytes);
		memcpy(homologs, h + tot * tot, bytes);
	} else {
		CUDA_CHECK(cudaMemcpyAsync(subst, c->d_subst.get(), bytes, cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaMemcpyAsync(homologs, c->d_hom.get(), bytes, cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaStreamSynchronize(s));
	}
	since("process.host_done_ms");
	// the same on the device's clock: when the index was built, when everything was compared
	CUDA_CHECK(cudaEventRecord(c->ev_mark[3], s));
	CUDA_CHECK(cudaEventSynchronize(c->ev_mark[3]));
	float ms = 0;
	CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev_mark[0], c->ev_mark[1]));
	c->stats["process.gpu_index_ms"] = ms;
	CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev_mark[0], c->ev_mark[2]));
	c->stats["process.gpu_compared_ms"] = ms;
	CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev_mark[0], c->ev_mark[3]));
	c->stats["process.gpu_done_ms"] = ms;
}

// after a failed call nothing of ours may still be reading the caller's buffers
void quiesce(phylo_ctx *c)
{
	c->stager.drain();
	cudaStreamSynchronize(c->copy_stream);
	cudaStreamSynchronize(c->stream);
	for (cudaStream_t st : c->push_streams)
		cudaStreamSynchronize(st);
	cudaGetLastError();
}

} // namespace

namespace
{

// Lays the sequences out in one device buffer of the context (16-byte aligned starts, a zero
// byte behind every sequence) and gets them on their way: sequence `first` (if < N) on the main
// stream at once — phylo_process needs the reference for the index — the others batch by
// batch next to whatever the main stream does meanwhile.  before_batch(b) puts the main stream
// behind the bytes of batch b.
struct Uploader {
	phylo_ctx *c;
	const char *const *seqs;
	const uint64_t *lens;
	uint64_t N, first;
	std::vector<uint64_t> ends;
	size_t queued = 0;
	bool pageable = false, packed = true, ref_staged = false;
	static constexpr uint64_t COPY_QUEUE = 256;

	Uploader(phylo_ctx *ctx, const char *const *seqs_, const uint64_t *lens_, uint64_t N_, uint64_t first_)
		: c(ctx), seqs(seqs_), lens(lens_), N(N_), first(first_)
	{
		c->q_resident = false;
		c->q_offs.assign((size_t)N, 0);
		c->q_lens.assign(lens, lens + N);
		uint64_t total = 0, bases = 0;
		for (uint64_t k = 0; k < N; k++) {
			if (!seqs[k] && lens[k]) throw std::invalid_argument("NULL sequence");
			c->q_offs[k] = total;
			total = (total + lens[k] + 1 + 15) / 16 * 16; // >= 1 zero byte, 16-byte aligned starts
			bases += lens[k];
		}
		// where do the sequences live?  one probe: callers do not mix pinned and ordinary memory
		for (uint64_t k = 0; k < N; k++)
			if (k != first && lens[k]) {
				pageable = HostStager::is_pageable(seqs[k]);
				break;
			}
		cudaStream_t s = c->stream;
		CUDA_CHECK(cudaStreamSynchronize(c->copy_stream)); // nothing may still write into the old buffer
		c->stager.drain();
		// (a block of the copy stream's own cache: one freed on the main stream may still be written
		// by work queued there — the scratch of a lazily built index)
		c->q_own.alloc(total + 64, c->copy_stream);
		// The buffer is cleared on the copy stream, behind whatever read the old one last (the row
		// builder of the previous mapping) — NOT on the main stream: phylo_map_queries finds an
		// index build queued there, and the sequences would start to cross the bus only after it
		// (measured, 8 x 5 Mbp from host buffers through the stage calls: 2.2 ms, of which 0.7 ms
		// were the uploads waiting behind the index).
		if (c->ev_q_idle) CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_q_idle, 0));
		CUDA_CHECK(cudaMemsetAsync(c->q_own.get(), 0, c->q_own.bytes(), c->copy_stream));
		CUDA_CHECK(cudaEventRecord(c->ev_main, c->copy_stream));
		CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_main, 0)); // (what the main stream does with the buffer comes after)
		uint8_t *dq = c->q_own.get();
		const uint64_t *offs = c->q_offs.data();
		ends = plan_batches(lens, N);
		// Packing pays when the bus is the bottleneck or the memory is pageable (the driver would
		// stage it on this thread), and when there are cores to do it: at ~6 GB/s per core it takes
		// eight of them to outrun the plain copy of pinned memory (55 GB/s).  A small pinned input
		// (measured: 8 x 5 Mbp) is over the bus before the index is built either way, and several
		// ranks that share one host's cores are better off with the copy engines.
		int threads = (int)c->opt_stage_threads;
		if (threads <= 0) {
			const unsigned hw = std::thread::hardware_concurrency();
			threads = hw > 18 ? 16 : hw > 3 ? (int)hw - 2 : 1;
		}
		packed = c->tuning.upload_raw < 0 ||
		         (c->tuning.upload_raw == 0 && (pageable || (bases >= (128ull << 20) && threads >= 8)));
		// The reference is what everything waits for.  From pinned memory it goes over as it is;
		// from ordinary memory the driver would stage it on this thread at a fraction of the bus
		// rate (measured: ~0.5 ms for 5 Mbp), so the workers take it first, in pieces small
		// enough that each of them gets one.
		ref_staged = packed && first < N && lens[first] && HostStager::is_pageable(seqs[first]);
		cudaEvent_t after = c->ev_main; // what the uploads of the other sequences must not overtake
		if (first < N && lens[first] && !ref_staged) {
			// the reference first: everything waits for it
			CUDA_CHECK(cudaMemcpyAsync(dq + offs[first], seqs[first], lens[first], cudaMemcpyHostToDevice, s));
			CUDA_CHECK(cudaEventRecord(c->ev_copy, s));
			CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
			after = c->ev_copy;
		}
		if (packed) {
			// worker threads pack the pieces to 2 bits per base into pinned rings, a kernel unpacks
			// them at their place (staging.h): a quarter of the bytes on the bus, pageable or pinned
			const int shift = ref_staged ? 1 : 0; // the reference's pieces are batch 0 of the stager then
			if (ref_staged) {
				uint64_t piece = (lens[first] + threads - 1) / threads;
				piece = std::max<uint64_t>(64u << 10, (piece + 63) / 64 * 64);
				c->stager.add(dq + offs[first], seqs[first], lens[first], 0, piece);
			}
			uint64_t k = 0;
			for (size_t b = 0; b < ends.size(); b++)
				for (; k < ends[b]; k++)
					if (k != first && lens[k]) c->stager.add(dq + offs[k], seqs[k], lens[k], (int)b + shift);
			if (!c->stager.empty()) c->stager.start(c->device, (int)ends.size() + shift, threads, after);
			if (ref_staged) c->stager.wait_batch(0, s); // the index build is queued behind the reference's pieces
		} else {
			// option "upload_raw": the bytes as they are, plain asynchronous copies on the copy
			// stream (truly asynchronous only from pinned memory) with an event behind every batch.  At most COPY_QUEUE copies are queued ahead of the batch being mapped —
			// a thousand queued copies fill the driver's queue and block the host until they have
			// drained, with the index build not yet launched (measured: 1000 x 3 Mbp) — but never
			// fewer than two batches.
			while (c->batch_events.size() < ends.size()) {
				cudaEvent_t e;
				CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
				c->batch_events.push_back(e);
			}
			queue_copies(0);
		}
		c->stats["process.pageable"] = pageable ? 1 : 0;
		c->stats["process.packed"] = packed ? 1 : 0;
		const uint64_t ref_bytes = first < N ? lens[first] : 0;
		c->stats["process.h2d_bytes"] = packed ? (double)((bases - ref_bytes) / 4 + (ref_staged ? ref_bytes / 4 : ref_bytes)) : (double)bases;
	}

	void queue_copies(size_t current) // `current` = the batch about to be mapped
	{
		uint8_t *dq = c->q_own.get();
		const uint64_t *offs = c->q_offs.data();
		const uint64_t seq0 = current ? ends[current - 1] : 0;
		size_t last = current + 2;
		while (last + 1 < ends.size() && ends[last + 1] - seq0 <= COPY_QUEUE)
			last++;
		for (; queued <= last && queued < ends.size(); queued++) {
			for (uint64_t k = queued ? ends[queued - 1] : 0; k < ends[queued]; k++)
				if (k != first && lens[k])
					CUDA_CHECK(cudaMemcpyAsync(dq + offs[k], seqs[k], lens[k], cudaMemcpyHostToDevice, c->copy_stream));
			CUDA_CHECK(cudaEventRecord(c->batch_events[queued], c->copy_stream));
		}
	}

	void before_batch(size_t b)
	{
		if (packed) {
			c->stager.wait_batch((int)b + (ref_staged ? 1 : 0), c->stream);
		} else {
			queue_copies(b);
			CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->batch_events[b], 0));
		}
	}

	MapHooks hooks()
	{
		MapHooks h;
		h.ends = &ends;
		h.before_batch = [this](size_t b) { before_batch(b); };
		// packed sequences were checked by the packer; the reference of phylo_process (copied as
		// it is) by the text kernel of the index build
		h.validated = packed;
		return h;
	}

	void done()
	{
		c->stager.finish();
		c->q_resident = true;
	}
};

} // namespace

int phylo_process(phylo_ctx *c, const char *const *seqs, const uint64_t *lens, uint64_t N, uint64_t ref_index, int flags,
                  uint64_t *subst, uint64_t *homologs)
{
	if (!c) return fail(nullptr, PHYLO_ERR_INVALID, "context is NULL");
	if (!seqs || !lens || !subst || !homologs || ref_index >= N)
		return fail(c, PHYLO_ERR_INVALID, "bad arguments to phylo_process");
	const int rc = guarded(c, [&] {
		Uploader up(c, seqs, lens, N, ref_index);
		process_resident(c, N, ref_index, flags, up.hooks(), subst, homologs);
		up.done();
	});
	if (rc != PHYLO_OK) quiesce(c);
	return rc;
}

int phylo_ingest_begin(phylo_ctx *c, uint64_t N, const uint64_t *max_lens, int lanes)
{
	const int rc = guarded(c, [&] {
		if (!N || !max_lens) throw std::invalid_argument("NULL argument");
		if (lanes < 1) lanes = 1;
		if (lanes > 64) lanes = 64;
		auto &in = c->ingest;
		cudaStream_t s = c->stream;
		CUDA_CHECK(cudaStreamSynchronize(c->copy_stream));
		c->stager.drain();
		c->q_resident = false;
		c->mapped = false;
		c->q_offs.assign((size_t)N, 0);
		c->q_lens.assign((size_t)N, 0);
		in.caps.assign(max_lens, max_lens + N);
		in.got.assign((size_t)N, 0);
		in.err.clear();
		uint64_t total = 0;
		for (uint64_t k = 0; k < N; k++) {
			if (max_lens[k] > 0x7fffff00ull) throw std::invalid_argument("sequence too long for 32-bit indices");
			c->q_offs[k] = total;
			total = (total + max_lens[k] + 1 + 15) / 16 * 16;
		}
		c->q_own.alloc(total + 64, s);
		c->q_own.zero();
		CUDA_CHECK(cudaEventRecord(c->ev_main, s));
		while ((int)in.lanes.size() < lanes) {
			in.lanes.emplace_back();
			in.lanes.back().create();
		}
		in.lane_busy.assign(in.lanes.size(), 0);
		for (auto &l : in.lanes)
			CUDA_CHECK(cudaStreamWaitEvent(l.stream, c->ev_main, 0)); // not before the buffer is cleared
		in.active = true;
	});
	return rc;
}

int phylo_ingest_put(phylo_ctx *c, uint64_t index, const char *seq, uint64_t len)
{
	if (!c) return PHYLO_ERR_INVALID;
	auto &in = c->ingest;
	auto failed = [&](int code, const std::string &msg) {
		std::lock_guard<std::mutex> lock(in.mu);
		if (in.err.empty()) in.err = msg;
		return code;
	};
	if (!in.active) return failed(PHYLO_ERR_INVALID, "phylo_ingest_begin has not been called");
	if (index >= in.caps.size() || (!seq && len)) return failed(PHYLO_ERR_INVALID, "bad arguments to phylo_ingest_put");
	if (len > in.caps[index]) return failed(PHYLO_ERR_INVALID, "sequence longer than announced to phylo_ingest_begin");
	if (cudaSetDevice(c->device) != cudaSuccess) return failed(PHYLO_ERR_CUDA, "cudaSetDevice failed");
	int lane = -1;
	{
		std::unique_lock<std::mutex> lock(in.mu);
		if (in.got[index]) {
			if (in.err.empty()) in.err = "sequence handed over twice";
			return PHYLO_ERR_INVALID;
		}
		in.got[index] = 1;
		in.cv.wait(lock, [&] {
			for (size_t l = 0; l < in.lane_busy.size(); l++)
				if (!in.lane_busy[l]) {
					lane = (int)l;
					return true;
				}
			return false;
		});
		in.lane_busy[lane] = 1;
	}
	UploadLane &L = in.lanes[lane];
	uint8_t *dst = c->q_own.get() + c->q_offs[index];
	const uint8_t *src = reinterpret_cast<const uint8_t *>(seq);
	int rc = PHYLO_OK;
	for (uint64_t o = 0; o < len && rc == PHYLO_OK; o += UploadLane::PIECE_BYTES) {
		const uint32_t l = (uint32_t)(len - o < UploadLane::PIECE_BYTES ? len - o : UploadLane::PIECE_BYTES);
		int bad = 0;
		const cudaError_t e = L.upload_piece(dst + o, src + o, l, &bad);
		if (bad) rc = failed(PHYLO_ERR_INVALID, "a sequence contains bytes outside {A,C,G,T,!}");
		else if (e != cudaSuccess) rc = failed(PHYLO_ERR_CUDA, std::string("upload failed: ") + cudaGetErrorString(e));
	}
	c->q_lens[index] = len; // one writer per index
	{
		std::lock_guard<std::mutex> lock(in.mu);
		in.lane_busy[lane] = 0;
	}
	in.cv.notify_one();
	return rc;
}

int phylo_ingest_end(phylo_ctx *c)
{
	const int rc = guarded(c, [&] {
		auto &in = c->ingest;
		if (!in.active) throw std::invalid_argument("phylo_ingest_begin has not been called");
		in.active = false;
		for (auto &l : in.lanes)
			CUDA_CHECK(cudaStreamSynchronize(l.stream));
		if (!in.err.empty()) throw std::invalid_argument(in.err);
		for (size_t k = 0; k < in.got.size(); k++)
			if (!in.got[k]) throw std::invalid_argument("phylo_ingest_end: not every sequence was handed over");
		c->N = in.got.size();
		c->q_resident = true;
		uint64_t bases = 0;
		for (uint64_t l : c->q_lens)
			bases += l;
		c->stats["process.h2d_bytes"] = (double)(bases / 4);
		c->stats["process.packed"] = 1;
	});
	if (rc != PHYLO_OK) quiesce(c);
	return rc;
}

int phylo_process_again(phylo_ctx *c, uint64_t ref_index, int flags, uint64_t *subst, uint64_t *homologs)
{
	if (!c) return fail(nullptr, PHYLO_ERR_INVALID, "context is NULL");
	const int rc = guarded(c, [&] {
		if (!subst || !homologs) throw std::invalid_argument("NULL argument");
		if (!c->q_resident) throw std::invalid_argument("phylo_process_again: no sequences resident on the device");
		const uint64_t N = c->q_lens.size();
		if (ref_index >= N) throw std::invalid_argument("reference index out of range");
		if (!c->q_lens[ref_index]) throw std::invalid_argument("reference is empty");
		CUDA_CHECK(cudaStreamSynchronize(c->copy_stream));
		process_resident(c, N, ref_index, flags, MapHooks(), subst, homologs);
		c->stats["process.h2d_bytes"] = 0;
	});
	if (rc != PHYLO_OK) quiesce(c);
	return rc;
}

int phylo_map_queries(phylo_ctx *c, const char *const *queries, const uint64_t *lens, uint64_t N, uint64_t threshold)
{
	if (!c) return fail(nullptr, PHYLO_ERR_INVALID, "context is NULL");
	const int rc = guarded(c, [&] {
		if (N && (!queries || !lens)) throw std::invalid_argument("NULL argument");
		// batch b is mapped while the later ones are still crossing PCIe
		Uploader up(c, queries, lens, N, N);
		do_map(c, c->q_own.get(), c->q_offs.data(), lens, N, threshold, up.hooks());
		up.done();
	});
	if (rc != PHYLO_OK) quiesce(c);
	return rc;
}

int phylo_esa_alloc(phylo_ctx *c, uint64_t n)
{
	return guarded(c, [&] {
		check_ref_length(n);
		cudaStream_t s = c->stream;
		c->esa_ready = false;
		c->mapped = false;
		c->esa.release();
		const int32_t m = (int32_t)(2 * n + 1);
		const int32_t padded = ((m + 64 + 255) / 256) * 256; // check_ref_length leaves room for the padding
		c->esa.n = (int32_t)n;
		c->esa.m = m;
		c->esa.S.alloc(padded, s);
		c->esa.S.zero();
		c->esa.SA.alloc(m, s);
		c->esa.LCP.alloc((size_t)m + 1, s);
		c->esa.CLD.alloc((size_t)m + 1, s);
		c->esa.FVC.alloc(m, s);
		CUDA_CHECK(cudaStreamSynchronize(s));
	});
}

int phylo_esa_device_arrays(const phylo_ctx *c, void **S, uint64_t *S_bytes, void **SA, void **LCP, void **CLD,
                            void **FVC)
{
	if (!c || !c->esa.m) return PHYLO_ERR_INVALID;
	if (c->esa.pending) { // the arrays are about to leave the library: the build must be over (and good)
		auto *m = const_cast<phylo_ctx *>(c);
		const int rc = guarded(m, [&] { finish_index(m); });
		if (rc != PHYLO_OK) return rc;
	}
	if (S) *S = c->esa.S.get();
	if (S_bytes) *S_bytes = c->esa.S.size();
	if (SA) *SA = c->esa.SA.get();
	if (LCP) *LCP = c->esa.LCP.get();
	if (CLD) *CLD = c->esa.CLD.get();
	if (FVC) *FVC = c->esa.FVC.get();
	return PHYLO_OK;
}

int phylo_esa_finish_import(phylo_ctx *c)
{
	return guarded(c, [&] {
		if (!c->esa.m) throw std::invalid_argument("phylo_esa_alloc has not been called");
		esa_build_table(c->esa, (int)c->opt_kmer, c->stream);
		DevBuf<unsigned long long> gc(1, c->stream);
		gc.zero();
		k_count_gc<<<NUM_SMS_B200 * 4, 256, 0, c->stream>>>(c->esa.S.get(), c->esa.n, gc.get());
		KERNEL_CHECK();
		c->esa.gc_count = (int64_t)d2h_scalar(gc.get(), c->stream);
		c->stats["esa.gc_count"] = (double)c->esa.gc_count;
		c->stats["esa.kmer_k"] = c->esa.K;
		c->esa_ready = true;
	});
}

int phylo_rows_configure(phylo_ctx *c, uint64_t total_genomes, uint64_t first_row)
{
	return guarded(c, [&] {
		if (total_genomes && first_row >= total_genomes) throw std::invalid_argument("first_row out of range");
		c->rows_total = total_genomes;
		c->rows_first = first_row;
		if (total_genomes && c->esa.m) {
			if ((uint64_t)c->rows.genomes != total_genomes || c->rows.n != c->esa.n) {
				clear_peers(c); // the store moves: addresses handed to the peers are stale
				rows_alloc(c->rows, (int64_t)total_genomes, c->esa.n, c->stream);
			} else // no slot counts as a written row until a mapping (here or on a peer) fills it
				rows_clear_flags(c->rows, c->stream);
			CUDA_CHECK(cudaStreamSynchronize(c->stream));
		}
	});
}

int phylo_rows_ipc_export(phylo_ctx *c, void *handle)
{
	return guarded(c, [&] {
		static_assert(sizeof(cudaIpcMemHandle_t) == PHYLO_IPC_HANDLE_BYTES, "handle size");
		if (!handle) throw std::invalid_argument("handle is NULL");
		if (!c->rows.data.get()) throw std::invalid_argument("no row store yet: call phylo_rows_configure after the index exists");
		cudaIpcMemHandle_t h;
		CUDA_CHECK(cudaIpcGetMemHandle(&h, c->rows.data.get()));
		memcpy(handle, &h, sizeof h);
	});
}

int phylo_rows_ipc_import(phylo_ctx *c, const void *handles, int world, int rank)
{
	return guarded(c, [&] {
		if (!handles || world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad arguments");
		if (!c->rows.data.get()) throw std::invalid_argument("no row store yet");
		clear_peers(c);
		c->peer_rows.assign((size_t)world, nullptr);
		c->peer_ipc.assign((size_t)world, nullptr);
		c->peer_rank = rank;
		for (int p = 0; p < world; p++) {
			if (p == rank) {
				c->peer_rows[p] = c->rows.data.get();
				continue;
			}
			cudaIpcMemHandle_t h;
			memcpy(&h, (const char *)handles + (size_t)p * sizeof h, sizeof h);
			void *ptr = nullptr;
			const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) {
				cudaGetLastError();
				clear_peers(c);
				throw CudaError(std::string("cudaIpcOpenMemHandle (row store of rank ") + std::to_string(p) + "): " + cudaGetErrorString(e));
			}
			c->peer_ipc[p] = ptr;
			c->peer_rows[p] = (uint32_t *)ptr;
		}
	});
}

int phylo_rows_set_peers(phylo_ctx *c, void *const *peer_rows, int world, int rank)
{
	return guarded(c, [&] {
		clear_peers(c);
		if (!peer_rows || world <= 0) return; // back to a context that keeps its rows to itself
		if (rank < 0 || rank >= world) throw std::invalid_argument("bad rank");
		if (!c->rows.data.get()) throw std::invalid_argument("no row store yet");
		c->peer_rank = rank;
		c->peer_rows.assign((size_t)world, nullptr);
		for (int p = 0; p < world; p++)
			c->peer_rows[p] = p == rank ? c->rows.data.get() : (uint32_t *)peer_rows[p];
	});
}

int phylo_rows_device(const phylo_ctx *c, void **rows, uint64_t *bytes_per_genome, uint64_t *total_genomes)
{
	if (!c || !c->rows.data.get()) return PHYLO_ERR_INVALID;
	if (rows) *rows = c->rows.data.get();
	if (bytes_per_genome) *bytes_per_genome = (uint64_t)c->rows.genome_words() * sizeof(uint32_t);
	if (total_genomes) *total_genomes = (uint64_t)c->rows.genomes;
	return PHYLO_OK;
}

} // extern "C"
