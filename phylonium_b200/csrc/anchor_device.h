// Per-query anchoring on the device (anchor.cu): replaces hot loop A of
// /root/reference/src/process.cxx:433-458 — anchor_homologies() + std::sort +
// filter_overlaps_max() for every query.
#pragma once
#include <functional>
#include <stdexcept>
#include <vector>

#include "common.cuh"
#include "esa_device.h"
#include "walk.h"

namespace phy
{

struct AnchorStats {
	int64_t chunks = 0;
	int64_t events = 0;        // accepted anchors on the true paths
	int64_t open_events = 0;   // matches longer than the per-thread cap
	int64_t unresolved = 0;    // bridges that gave up and were continued serially
	int64_t tie_fallback = 0;  // batches whose sort went through std::sort on the host
	int64_t general_path = 0;  // batches sorted/filtered by the global path (a list > 2048, or ties)
	int lists_redone = 0;      // a bridge on a true path had to be continued: on_filtered saw lists that were not final
	int input_flags = 0;       // copy of *AnchorOptions::input_flags; != 0: nothing was mapped
	float walk_ms = 0, open_ms = 0, bridge_ms = 0, path_ms = 0, assemble_ms = 0, filter_ms = 0, total_ms = 0;
};

struct AnchorOptions {
	int32_t chunk = 2048; // CH, multiple of 32
	int32_t cap = 0;      // comparison cap per thread; 0 = 2 * chunk
	bool keep_raw = false; // also keep the unsorted, unfiltered lists (tests)
	bool timings = false;
	// device int set by the caller's input validation kernel (same stream); read back with the
	// first synchronisation of the mapping instead of one of its own
	const int *input_flags = nullptr;
	// if set: recorded (on another stream) behind the validation kernel; the mapping stream
	// waits for it only right before it reads input_flags, so the validation runs next to the walk
	cudaEvent_t input_flags_ready = nullptr;
	// An index whose build has not been looked at yet (EsaDevice::pending): the device flag that
	// makes the kernels do nothing if the build failed, and where the host finds the verdict once
	// the stream has been synchronised (pinned).  The mapping throws IndexNotBuilt then.
	const int *index_skip = nullptr;
	const int *index_verdict_host = nullptr;
	// Called once the filtered lists of the batch are queued (device pointers: lists, per-query
	// begin and count) and before the host waits for them: whatever the caller queues here (the
	// row builder) runs while the host is still looking at the counts.  If the lists turn out to
	// need the general path (AnchorStats::general_path), the final ones replace them and the
	// caller has to redo that work.
	std::function<void(const Hom *, const int64_t *, const int64_t *)> on_filtered;
	// If the caller is capturing the stream into this graph (GraphSegment, common.cuh): the
	// mapping submits it right before its host stop — everything up to there, on_filtered's work
	// included, is then one graph.
	GraphSegment *graph = nullptr;
};

struct IndexNotBuilt : std::runtime_error {
	IndexNotBuilt() : std::runtime_error("the speculative index build failed: build it again step by step") {}
};

struct AnchorResult {
	DevBuf<Hom> homs;                  // filtered lists; query q owns homs[begin[q] .. begin[q] + count[q])
	std::vector<int64_t> begin, count; // (the lists need not be packed back to back)
	DevBuf<int64_t> d_begin, d_count;  // same on the device
	std::vector<int64_t> offs;         // scratch of the general path: nq + 1 packed offsets
	DevBuf<int64_t> d_offs;
	DevBuf<Hom> raw;           // push-order lists before sort/filter (keep_raw)
	std::vector<int64_t> raw_offs;
};

// d_Q: concatenated queries on the device, every query followed by at least one zero byte.
// qi[k].chunk_base / nchunks are filled in here.
void anchor_queries_device(const EsaDevice &esa, const uint8_t *d_Q, std::vector<QueryInfo> &qi, int32_t thr,
                           const AnchorOptions &opt, cudaStream_t s, AnchorResult &out, AnchorStats *stats);

// host std::sort + filter for one list; used when equal starts make the reference's
// unstable sort implementation-defined (see anchor.cu)
void host_sort_filter(std::vector<Hom> &list);

} // namespace phy
