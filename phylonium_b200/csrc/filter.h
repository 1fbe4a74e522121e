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
// cluster sizes instead of h^2; identical output.
#pragma once
#include "walk.h"

namespace phy
{

// start/len: the list (sorted by start). score (int64), pred (int32), keep (uint8):
// scratch/outputs of h entries. Returns the number of survivors; keep[k] marks them.
PHY_HD int32_t filter_overlaps_max(const int32_t *start, const int32_t *len, int32_t h, int64_t *score,
                                   int32_t *pred, uint8_t *keep)
{
	if (h < 2) {
		for (int32_t k = 0; k < h; k++)
			keep[k] = 1;
		return h;
	}
	int64_t prefix_best = 0; // best score among elements left of the current cluster
	int32_t prefix_at = -1;
	int32_t c0 = 0;       // first element of the current cluster
	int64_t max_end = 0;  // largest end inside the current cluster
	int64_t top = 0;      // global maximum, slot -1 holds 0 (max_element over score_buffer)
	int32_t top_at = -1;
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

} // namespace phy
