// Child table (CLD) without a stack.
//
// /root/reference/src/esa.cxx:256-298 fills CLD in one sequential pass with a stack.
// Its result has a closed form (re-derived in DESIGN.md):  for 0 <= i < m
//   up   (LCP[i+1] <  LCP[i]):  p = max{q <= i : LCP[q] <= LCP[i+1]},
//                               CLD[i] = leftmost argmin of LCP over (p, i]
//   else                        s = min{q > i : LCP[q] <= LCP[i]};
//        next (LCP[s] == LCP[i]): CLD[i] = s
//        down (LCP[s] <  LCP[i]): CLD[i] = leftmost argmin of LCP over (i, s)
// and CLD[m] = 0.  Every entry is independent, so one thread per entry works, given a way
// to answer "first position to the right / last position to the left with LCP <= v" and
// "range minimum" quickly.  A min-pyramid with fan-out 32 over LCP does that: level 0 is
// LCP itself (m + 1 entries), level k+1 holds the minimum of each group of 32 of level k.
// Nearly all queries finish inside the first group they scan.
#pragma once
#include "esa_types.h"

namespace phy
{

constexpr int PYR_MAX_LEVELS = 8;

struct Pyramid {
	const int32_t *level[PYR_MAX_LEVELS]; // level[0] = LCP
	int32_t size[PYR_MAX_LEVELS];
	int32_t levels;
};

// first q >= from with LCP[q] <= v; the caller guarantees one exists (LCP[m] = -1 <= v)
PHY_HD int32_t pyr_first_le_right(const Pyramid &py, int32_t from, int32_t v)
{
	int32_t lvl = 0, idx = from;
	for (;;) {
		const int32_t *L = py.level[lvl];
		int32_t end = (idx | 31) + 1;
		if (end > py.size[lvl]) end = py.size[lvl];
		bool found = false;
		for (; idx < end; idx++) {
			if (L[idx] <= v) {
				found = true;
				break;
			}
		}
		if (found) break;
		idx >>= 5; // idx is a multiple of 32 here: first group not yet looked at
		lvl++;
	}
	while (lvl > 0) {
		lvl--;
		idx <<= 5;
		const int32_t *L = py.level[lvl];
		while (L[idx] > v)
			idx++;
	}
	return idx;
}

// last q <= from with LCP[q] <= v; the caller guarantees one exists (LCP[0] = -1 <= v)
PHY_HD int32_t pyr_last_le_left(const Pyramid &py, int32_t from, int32_t v)
{
	int32_t lvl = 0, idx = from;
	for (;;) {
		const int32_t *L = py.level[lvl];
		const int32_t begin = idx & ~31;
		bool found = false;
		for (; idx >= begin; idx--) {
			if (L[idx] <= v) {
				found = true;
				break;
			}
		}
		if (found) break;
		idx = (begin >> 5) - 1; // group to the left, one level up
		lvl++;
	}
	while (lvl > 0) {
		lvl--;
		idx = (idx << 5) + 31;
		if (idx >= py.size[lvl]) idx = py.size[lvl] - 1;
		const int32_t *L = py.level[lvl];
		while (L[idx] > v)
			idx--;
	}
	return idx;
}

// minimum of LCP over [a, b], a <= b
PHY_HD int32_t pyr_range_min(const Pyramid &py, int32_t a, int32_t b)
{
	int32_t mn = 0x7fffffff;
	int32_t lo = a, hi = b + 1, lvl = 0;
	while (lo < hi) {
		const int32_t *L = py.level[lvl];
		int32_t lend = (lo + 31) & ~31;
		if (lend > hi) lend = hi;
		for (int32_t t = lo; t < lend; t++)
			mn = L[t] < mn ? L[t] : mn;
		lo = lend;
		int32_t hbeg = hi & ~31;
		if (hbeg < lo) hbeg = lo;
		for (int32_t t = hbeg; t < hi; t++)
			mn = L[t] < mn ? L[t] : mn;
		hi = hbeg;
		lo >>= 5;
		hi >>= 5;
		lvl++;
	}
	return mn;
}

PHY_HD int32_t cld_entry(const Pyramid &py, int32_t i)
{
	const int32_t *LCP = py.level[0];
	const int32_t a = LCP[i], b = LCP[i + 1];
	if (b < a) { // up value of i + 1, stored at i
		const int32_t p = pyr_last_le_left(py, i, b);
		const int32_t mv = pyr_range_min(py, p + 1, i);
		return pyr_first_le_right(py, p + 1, mv);
	}
	const int32_t s = pyr_first_le_right(py, i + 1, a);
	if (LCP[s] == a) return s; // nextlIndex
	const int32_t mv = pyr_range_min(py, i + 1, s - 1);
	return pyr_first_le_right(py, i + 1, mv); // down
}

} // namespace phy
