// Device-resident enhanced suffix array and its construction (esa_build.cu).
// Replaces esa::esa of /root/reference/src/esa.cxx:69-81 (divsufsort64 + init_LCP +
// init_CLD + init_FVC + init_cache).
#pragma once
#include "common.cuh"
#include "esa_types.h"

namespace phy
{

struct EsaTimings {
	bool enabled = false; // in: record per-phase device times (adds synchronisation)
	float text_ms = 0, keys_ms = 0, sort_ms = 0, refine_ms = 0, lcp_ms = 0, cld_ms = 0, table_ms = 0, total_ms = 0;
	float hist_ms_avg = 0, scan_ms_avg = 0, scatter_ms_avg = 0; // per radix pass of the main sort
	int sort_passes = 0;
	float first_pass_ms = 0; // packed sorter: the pass that reads the text (histogram + scan + scatter)
	bool packed = false;     // sorted as packed 2-bit words (suffix_sort.cuh) or as 3-bit keys + indices
	int64_t dirty = 0;       // packed sorter: suffixes with a byte below 'A' among their key characters
	int key_chars = 0; // characters per sort key actually used
	int refine_rounds = 0;
	int64_t tie_groups = 0; // groups of suffixes with equal sort keys
	int64_t tied = 0;       // suffixes left to the doubling rounds (groups too large or too similar)
};

struct EsaDevice {
	int32_t n = 0, m = 0, K = 0;
	int64_t gc_count = 0; // G/C bytes in the reference (for the anchor threshold)
	DevBuf<uint8_t> S;   // m + 64 bytes, zero padded
	DevBuf<uint8_t> FVC; // m
	DevBuf<int32_t> SA;  // m
	DevBuf<int32_t> LCP; // m + 1
	DevBuf<int32_t> CLD; // m + 1
	DevBuf<EsaNode> node;   // m + 1: SA/LCP/CLD/FVC interleaved for the descent (esa_search.h)
	DevBuf<TableRec> table; // 4^K
	// side stream of the build (min-pyramid next to the child-table kernel), made on first use
	cudaStream_t side = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	// Verdicts of the speculative build (esa_build.cu).  report (device, 8 ints): [0] bytes outside
	// the alphabet, [1] G/C count, [2] separators — written by the first kernel —, [3] != 0: an
	// assumption of the build failed and the kernels behind the check did nothing, [4..7]
	// {tie groups, hard groups, dirty suffixes, dirty error}.  h_report (pinned, 16 ints): [0..2]
	// arrive early (ev_text: right behind the first kernel, on the side stream), [8..15] = report at
	// the end of the build (ev_done, on the build's stream).  While `pending` the host has not
	// looked at the second half yet: esa_finish() does, and what is queued behind the build in the
	// meantime takes `skip()` along.
	DevBuf<int> report;
	int *h_report = nullptr;
	cudaEvent_t ev_text = nullptr, ev_early = nullptr, ev_done = nullptr;
	bool pending = false;
	GraphSegment build_graph; // the speculative build behind its first radix pass, as one graph
	int pend_kmer_k = -1, pend_key_chars = 0;
	const int *skip() const { return pending ? report.get() + 3 : nullptr; }
	const int *host_verdict() const { return pending ? h_report + 8 + 3 : nullptr; }
	void destroy_side()
	{
		build_graph.destroy();
		if (h_report) {
			cudaFreeHost(h_report);
			h_report = nullptr;
		}
		if (ev_text) {
			cudaEventDestroy(ev_text);
			cudaEventDestroy(ev_early);
			cudaEventDestroy(ev_done);
			ev_text = nullptr;
		}
		if (!side) return;
		cudaStreamSynchronize(side);
		cudaStreamDestroy(side);
		cudaEventDestroy(ev_fork);
		cudaEventDestroy(ev_join);
		side = nullptr;
	}
	EsaView view() const
	{
		EsaView v;
		v.S = S.get();
		v.SA = SA.get();
		v.LCP = LCP.get();
		v.CLD = CLD.get();
		v.FVC = FVC.get();
		v.node = node.get();
		v.table = table.get();
		v.K = K;
		v.m = m;
		v.n = n;
		return v;
	}
	void release()
	{
		S.release();
		FVC.release();
		SA.release();
		LCP.release();
		CLD.release();
		node.release();
		table.release();
		n = m = K = 0;
		pending = false; // (report stays: kernels of an abandoned build may still be reading it)
	}
};

// d_ref: n reference bytes over {A,C,G,T,!} already on the device. Throws CudaError /
// std::invalid_argument. kmer_k < 0 picks K from m.
// key_chars <= 0 picks the number of characters per sort key from m.
// lazy: return as soon as everything is queued and the first kernel's verdict (alphabet, G/C
// count) is in; esa.pending then says that esa_finish() has not been called yet.
void esa_build_device(EsaDevice &esa, const uint8_t *d_ref, int32_t n, int kmer_k, int key_chars, cudaStream_t stream,
                      EsaTimings *timings, bool lazy = false);

// Waits for a lazily built index and looks at its verdict; if an assumption of the speculative
// build did not hold, the index is built again step by step (from the text already on the
// device).  Returns true if that happened: work queued behind the build with esa.skip() did
// nothing and has to be queued again.  No-op (false) unless esa.pending.
bool esa_finish(EsaDevice &esa, cudaStream_t stream, EsaTimings *timings);

// builds only the table (used after importing S/SA/LCP/CLD/FVC from another GPU)
// nodes_ready: esa.node already holds the interleaved records
// skip: device flag; if set when the kernel starts, it does nothing (speculative build)
void esa_build_table(EsaDevice &esa, int kmer_k, cudaStream_t stream, bool nodes_ready = false, const int *skip = nullptr);

int esa_default_k(int32_t m);

// options "sort_path" and "table_direct": see Tuning (common.cuh)
// most dirty suffixes the packed sorter orders by pairwise comparison (about 1000 contigs)
constexpr int64_t PK_DIRTY_CAP = 32768;

} // namespace phy
