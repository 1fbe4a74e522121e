// filter_overlaps_max of /root/reference/src/process.cxx:354-401 for one sorted list:
// keep the chain of pairwise non-overlapping homologies with the largest total length.
//
// The reference runs an O(h^2) DP: score[i] = len[i] + max{score[k] : k < i, end[k] <=
// start[i]} taking the FIRST k that attains the maximum (strict '>'), a virtual
// predecessor -1 with score 0, then the FIRST global maximum (std::max_element) and a
// backtrack.  We compute exactly the same predecessors but look only at candidates that
// can differ: the list is sorted by start, so once the largest end seen so far is <= the
// current start ("clean cut") every earlier element is compatible with every later one and
// the best of them is a running prefix maximum (first index wins ties, and earlier indices
// win against later ones, as in the reference's left-to-right scan).  Cost: sum of squared
// cluster sizes instead of h^2; identical output.  When the clusters are large that is still
// quadratic; filter_overlaps_heap below is O(h log h) always.
#pragma once
#include "walk.h"

namespace phy
{

// The cluster version described above.  Gives up (returns -1, outputs undefined) once its
// inner loops have done more than `budget` steps: one cluster of thousands of mutually
// overlapping homologies — real, rearranged genomes — is quadratic here.
PHY_HD int32_t filter_overlaps_clusters(const int32_t *start, const int32_t *len, int32_t h, int64_t *score,
                                        int32_t *pred, uint8_t *keep, int64_t budget)
{
	int64_t prefix_best = 0; // best score among elements left of the current cluster
	int32_t prefix_at = -1;
	int32_t c0 = 0;       // first element of the current cluster
	int64_t max_end = 0;  // largest end inside the current cluster
	int64_t top = 0;      // global maximum, slot -1 holds 0 (max_element over score_buffer)
	int32_t top_at = -1;
	int64_t work = 0;
	for (int32_t i = 0; i < h; i++) {
		const int64_t s = start[i];
		if (i > c0 && max_end <= s) {
			// clean cut: fold the finished cluster into the prefix maximum
			for (int32_t k = c0; k < i; k++) {
				if (score[k] > prefix_best) {
					prefix_best = score[k];
					prefix_at = k;
				}
			}
			c0 = i;
			max_end = 0;
		}
		work += i - c0;
		if (work > budget) return -1;
		int64_t best = prefix_best;
		int32_t at = prefix_at;
		for (int32_t k = c0; k < i; k++) {
			if ((int64_t)start[k] + len[k] > s) continue; // !ends_left_of
			if (score[k] > best) {
				best = score[k];
				at = k;
			}
		}
		pred[i] = at;
		score[i] = best + len[i];
		const int64_t e = s + len[i];
		if (e > max_end) max_end = e;
		if (score[i] > top) {
			top = score[i];
			top_at = i;
		}
		keep[i] = 0;
	}
	int32_t kept = 0;
	for (int32_t k = top_at; k >= 0; k = pred[k]) {
		keep[k] = 1;
		kept++;
	}
	return kept;
}

// The same predecessors in O(h log h), whatever the clusters look like.  The list is sorted by
// start, so "k ends left of i" (end[k] <= start[i]), once true, stays true for every later i:
// the elements whose end has not been passed yet wait in a min-heap ordered by end; those the
// current start has passed leave it for good and are folded into ONE running best — largest
// score, smallest index among equals, which is what the reference's left-to-right scan with
// its strict '>' returns over the same set.  heap: scratch of h entries.
PHY_HD int32_t filter_overlaps_heap(const int32_t *start, const int32_t *len, int32_t h, int64_t *score, int32_t *pred,
                                    uint8_t *keep, int32_t *heap)
{
	int32_t hn = 0;
	int64_t best = 0;
	int32_t at = -1;
	int64_t top = 0;
	int32_t top_at = -1;
	for (int32_t i = 0; i < h; i++) {
		const int64_t s = start[i];
		while (hn > 0 && (int64_t)start[heap[0]] + len[heap[0]] <= s) {
			const int32_t k = heap[0];
			if (score[k] > best || (score[k] == best && at >= 0 && k < at)) {
				best = score[k];
				at = k;
			}
			// pop: the last entry sinks from the root
			const int32_t x = heap[--hn];
			const int64_t xe = (int64_t)start[x] + len[x];
			int32_t p = 0;
			for (;;) {
				int32_t c = 2 * p + 1;
				if (c >= hn) break;
				if (c + 1 < hn && (int64_t)start[heap[c + 1]] + len[heap[c + 1]] < (int64_t)start[heap[c]] + len[heap[c]]) c++;
				if ((int64_t)start[heap[c]] + len[heap[c]] >= xe) break;
				heap[p] = heap[c];
				p = c;
			}
			if (hn > 0) heap[p] = x;
		}
		pred[i] = at;
		score[i] = best + len[i];
		if (score[i] > top) {
			top = score[i];
			top_at = i;
		}
		keep[i] = 0;
		// push i
		const int64_t e = s + len[i];
		int32_t p = hn++;
		while (p > 0) {
			const int32_t up = (p - 1) >> 1;
			if ((int64_t)start[heap[up]] + len[heap[up]] <= e) break;
			heap[p] = heap[up];
			p = up;
		}
		heap[p] = i;
	}
	int32_t kept = 0;
	for (int32_t k = top_at; k >= 0; k = pred[k]) {
		keep[k] = 1;
		kept++;
	}
	return kept;
}

// start/len: the list (sorted by start). score (int64), pred (int32), keep (uint8), heap
// (int32): scratch/outputs of h entries. Returns the number of survivors; keep[k] marks them.
// Clusters first (a handful of steps per element on ordinary data), the heap when they
// turn out to be large: O(h log h) in every case.
PHY_HD int32_t filter_overlaps_max(const int32_t *start, const int32_t *len, int32_t h, int64_t *score,
                                   int32_t *pred, uint8_t *keep, int32_t *heap)
{
	if (h < 2) {
		for (int32_t k = 0; k < h; k++)
			keep[k] = 1;
		return h;
	}
	const int32_t r = filter_overlaps_clusters(start, len, h, score, pred, keep, 32 * (int64_t)h + 256);
	if (r >= 0) return r;
	return filter_overlaps_heap(start, len, h, score, pred, keep, heap);
}

} // namespace phy
