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
	void destroy_side()
	{
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
	}
};

// d_ref: n reference bytes over {A,C,G,T,!} already on the device. Throws CudaError /
// std::invalid_argument. kmer_k < 0 picks K from m.
// key_chars <= 0 picks the number of characters per sort key from m.
void esa_build_device(EsaDevice &esa, const uint8_t *d_ref, int32_t n, int kmer_k, int key_chars, cudaStream_t stream,
                      EsaTimings *timings);

// builds only the table (used after importing S/SA/LCP/CLD/FVC from another GPU)
// nodes_ready: esa.node already holds the interleaved records
// skip: device flag; if set when the kernel starts, it does nothing (speculative build)
void esa_build_table(EsaDevice &esa, int kmer_k, cudaStream_t stream, bool nodes_ready = false, const int *skip = nullptr);

int esa_default_k(int32_t m);

// options "sort_path" and "table_direct": see Tuning (common.cuh)
// most dirty suffixes the packed sorter orders by pairwise comparison (about 1000 contigs)
constexpr int64_t PK_DIRTY_CAP = 32768;

} // namespace phy
