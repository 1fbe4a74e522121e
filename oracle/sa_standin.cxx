/* TEST INFRASTRUCTURE — CPU oracle only; never linked into the product library.
 *
 * Stand-in for libdivsufsort64's divsufsort64(), which phylonium calls at
 * /root/reference/src/esa.cxx:74 but which is neither vendored in the
 * reference tree nor installed in this image (upstream
 * y-256/libdivsufsort, no version pinned by configure.ac:43-44).
 *
 * Contract restated: SA[0..n) is the permutation of 0..n-1 that orders the
 * suffixes T[i..n) lexicographically by unsigned byte, a proper prefix sorting
 * before the longer string.  All suffixes are distinct, so the answer is unique
 * and any correct sorter yields the same array ("parity pinned by definition").
 *
 * Method: re-code the bytes that occur densely, pack the first K symbols of
 * every suffix into one 63-bit key (K = 21 for the 6-letter ESA text), LSD radix
 * sort (key, index) pairs, then finish every run of equal keys with a
 * comparison sort on the remaining bytes.  Plain, single threaded like the
 * original; only meant to be correct and not embarrassingly slow.
 */
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "divsufsort64.h"

namespace
{

struct item {
	uint64_t key;
	int64_t idx;
};

/* dense re-coding of the bytes that actually occur keeps keys short: with
 * sigma <= 7 symbols (+1 for "past the end") 21 symbols fit into 63 bits. */
struct alphabet {
	int code[256];
	int sigma = 0;
	int bits = 0;
	int per_key = 0;
};

alphabet make_alphabet(const unsigned char *T, int64_t n)
{
	alphabet a;
	bool seen[256] = {false};
	for (int64_t i = 0; i < n; i++)
		seen[T[i]] = true;
	for (int c = 0; c < 256; c++) {
		a.code[c] = 0;
		if (seen[c]) a.code[c] = ++a.sigma; // 0 is reserved for "past the end"
	}
	a.bits = 1;
	while ((1 << a.bits) <= a.sigma)
		a.bits++;
	a.per_key = 63 / a.bits;
	return a;
}

void radix_sort(std::vector<item> &v, int total_bits)
{
	const int DIGIT = 11;
	const size_t BINS = (size_t)1 << DIGIT;
	std::vector<item> tmp(v.size());
	std::vector<size_t> count(BINS);
	for (int shift = 0; shift < total_bits; shift += DIGIT) {
		std::fill(count.begin(), count.end(), 0);
		for (const auto &it : v)
			count[(it.key >> shift) & (BINS - 1)]++;
		size_t sum = 0;
		for (size_t b = 0; b < BINS; b++) {
			size_t c = count[b];
			count[b] = sum;
			sum += c;
		}
		for (const auto &it : v)
			tmp[count[(it.key >> shift) & (BINS - 1)]++] = it;
		v.swap(tmp);
	}
}

} // namespace

extern "C" double po_sa_seconds;

extern "C" int divsufsort64(const unsigned char *T, saidx64_t *SA, saidx64_t n)
{
	if (n < 0 || (n > 0 && (!T || !SA))) return -1;
	if (n == 0) return 0;
	struct stopwatch {
		std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
		~stopwatch()
		{
			po_sa_seconds +=
				std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		}
	} sw;

	const alphabet a = make_alphabet(T, n);
	const int K = a.per_key;
	const uint64_t mask = (a.bits * K == 64) ? ~0ull : ((1ull << (a.bits * K)) - 1);

	std::vector<item> v((size_t)n);
	// rolling key, right to left: key(i) = code(T[i]) . key(i+1) >> bits
	uint64_t key = 0;
	for (int64_t i = n - 1; i >= 0; i--) {
		key = (key >> a.bits) | ((uint64_t)a.code[T[i]] << (a.bits * (K - 1)));
		key &= mask;
		v[(size_t)i] = {key, i};
	}
	radix_sort(v, a.bits * K);

	// finish runs of equal keys by comparing the tails
	auto tail_less = [&](const item &x, const item &y) {
		int64_t p = x.idx + K, q = y.idx + K;
		// equal keys without a 0 code inside means both have >= K bytes
		int64_t lp = n - p, lq = n - q;
		int64_t l = std::min(lp, lq);
		int c = l > 0 ? std::memcmp(T + p, T + q, (size_t)l) : 0;
		if (c != 0) return c < 0;
		return lp < lq;
	};
	size_t i = 0, N = v.size();
	while (i < N) {
		size_t j = i + 1;
		while (j < N && v[j].key == v[i].key)
			j++;
		if (j - i > 1) std::sort(v.begin() + i, v.begin() + j, tail_less);
		i = j;
	}
	for (size_t k = 0; k < N; k++)
		SA[k] = v[k].idx;
	return 0;
}

/* seconds spent inside divsufsort64() since last reset; lets the timed drivers
 * report the stand-in sort separately from init_LCP/CLD/FVC */
extern "C" double po_sa_seconds;
double po_sa_seconds = 0.0;
