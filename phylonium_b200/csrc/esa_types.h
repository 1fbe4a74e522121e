// Shared host/device plain-data types of the device-resident enhanced suffix array.
//
// The layout follows what the reference keeps in class esa
// (/root/reference/src/esa.h:45-64) but with 32-bit indices: the reference's own descent
// truncates to int (src/esa.cxx:374-375,503), so m = 2n+1 < 2^31 is the real limit and
// int32 arrays are exact (SURVEY.md §5 "long-context").
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PHY_HD __host__ __device__ __forceinline__
#else
#define PHY_HD inline
#endif

namespace phy
{

// lcp_interval of src/esa.h:31-40. Also the record type of the k-mer table:
//  - i == j : singleton, l = number of query characters already verified
//  - i <  j : lcp-interval with lcp value l and first l-index m; when it comes out of the
//             k-mer table, min(K, l) characters are verified
struct Interval {
	int32_t l, i, j, m;
};

// everything the descent needs to know about one suffix-array index, in one 16-byte load.
// fvc: low byte FVC[x]; upper 24 bits a hint, LCP[CLD[x]] + 1 (0xffffff: too large, look it
// up) — the lcp value of the child interval CLD[x] points at, which the descent needs right
// after CLD[x] itself; with the hint that is no second dependent load.
struct alignas(16) EsaNode {
	int32_t sa, lcp, cld, fvc;
};
constexpr uint32_t ESA_HINT_NONE = 0xffffffu;

PHY_HD int32_t esa_pack_fvc(uint8_t fvc, int32_t lcp_of_cld)
{
	const uint32_t hint = (uint32_t)(lcp_of_cld + 1) < ESA_HINT_NONE ? (uint32_t)(lcp_of_cld + 1) : ESA_HINT_NONE;
	return (int32_t)((uint32_t)fvc | (hint << 8));
}

// One record of the K-mer table: the descent state for the K-mer (see Interval) and, for a
// proper interval, everything the NEXT step of the descent would have to fetch first — the
// records of index i, of the first l-index m and of m - 1, and the text byte S[SA[i] + l] (kept
// in np.sa, which the descent never reads) — so that a search that starts from the table takes
// its first step down without a single dependent load.  64 bytes, one or two sectors.
struct alignas(64) TableRec {
	Interval ij;
	EsaNode ni, nm, np;
};

struct EsaView {
	const uint8_t *S;   // m bytes, followed by >= 64 zero bytes
	const int32_t *SA;  // m
	const int32_t *LCP; // m + 1, LCP[0] = LCP[m] = -1
	const int32_t *CLD; // m + 1
	const uint8_t *FVC; // m
	const EsaNode *node; // m + 1: the four arrays above interleaved (used by the descent)
	const TableRec *table; // 4^K records, the GPU counterpart of the 6-mer cache (src/esa.cxx:90-228)
	int32_t K;
	int32_t m; // 2n + 1
	int32_t n; // reference length == index of '#'
};

// result of a longest-match search
struct Match {
	int32_t l, i, j;
	int32_t open; // != 0: comparison stopped at the cap while still matching (singletons only)
	int32_t sa;   // SA[i] when i == j (the only case in which the walk uses it), else -1
};

// 3-bit text codes used by the suffix sorter. Order == unsigned byte order of the
// reference's text: end-of-text < '!' < '#' < A < C < G < T  
PHY_HD uint32_t text_code(uint8_t c)
{
	switch (c) {
		case '!': return 1;
		case '#': return 2;
		case 'A': return 3;
		case 'C': return 4;
		case 'G': return 5;
		case 'T': return 6;
	}
	return 0;
}

PHY_HD uint8_t text_char(uint32_t code)
{
	// bytes 0,'!','#','A','C','G','T',0 packed little-end first
	return (uint8_t)(0x0054474341232100ull >> (8 * (code & 7)));
}

// 2-bit code of the k-mer table, same as char2code in src/esa.cxx:48-62; -1 for non-ACGT
PHY_HD int kmer_code(uint8_t c)
{
	switch (c) {
		case 'A': return 0;
		case 'C': return 1;
		case 'G': return 2;
		case 'T': return 3;
	}
	return -1;
}

} // namespace phy
