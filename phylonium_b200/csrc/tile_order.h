// Order in which the all-pairs stage visits the tile pairs of the count matrix (compare.cu,
// mirrored by sharding.py and checked on the host by tests/test_emulation.py).
#pragma once
#include <stdint.h>

#include "esa_types.h"

namespace phy
{

// Order of the tile pairs (ti <= tj, tj in [tile_begin, tile_end)) of one launch: in bands of
// CMP_BAND tile rows, column by column inside a band.  The ~148 blocks that run at the same
// time then work on a 12 x 12 patch of the matrix and walk along their rows in step: 24 tile
// rows stream through L2 for 148 tile pairs, instead of 64 + 2 when a whole column of the matrix
// is in flight (measured at 1000 x 3 Mbp in column order: L2 hit rate 36 %, the copies late).
constexpr int CMP_BAND = 12;

// pairs of band b (rows [b * CMP_BAND, ...)) up to column tj_end (exclusive), columns from tile_begin
PHY_HD int64_t cmp_band_pairs(int32_t b, int32_t tile_begin, int32_t tj_end)
{
	const int32_t ti0 = b * CMP_BAND;
	const int32_t c0 = tile_begin > ti0 ? tile_begin : ti0;
	if (tj_end <= c0) return 0;
	const int32_t tfull = ti0 + CMP_BAND - 1; // first column that crosses the whole band
	int64_t n = 0;
	const int32_t tri_end = tj_end < tfull ? tj_end : tfull;
	if (tri_end > c0) // columns c0 .. tri_end - 1 hold (tj - ti0 + 1) pairs each
		n += (int64_t)(c0 - ti0 + 1 + tri_end - ti0) * (tri_end - c0) / 2;
	const int32_t f0 = c0 > tfull ? c0 : tfull;
	if (tj_end > f0) n += (int64_t)(tj_end - f0) * CMP_BAND;
	return n;
}

PHY_HD void cmp_unrank_pair(int64_t p, int32_t tile_begin, int32_t tile_end, int32_t &ti, int32_t &tj)
{
	int32_t b = 0;
	for (;; b++) {
		const int64_t n = cmp_band_pairs(b, tile_begin, tile_end);
		if (p < n) break;
		p -= n;
	}
	const int32_t ti0 = b * CMP_BAND;
	int32_t c = tile_begin > ti0 ? tile_begin : ti0;
	const int32_t tfull = ti0 + CMP_BAND - 1;
	for (; c < tfull; c++) { // the band's triangular head, column by column
		const int32_t rows = c - ti0 + 1;
		if (p < rows) break;
		p -= rows;
	}
	if (c >= tfull) {
		c += (int32_t)(p / CMP_BAND);
		p %= CMP_BAND;
	}
	tj = c;
	ti = ti0 + (int32_t)p;
}

} // namespace phy
