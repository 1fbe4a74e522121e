// Enhanced suffix array construction on the GPU.
//
// Replaces the single-threaded esa::esa of /root/reference/src/esa.cxx:69-81:
//   S = R '#' revcomp(R)                                   (esa.cxx:72, sequence.cxx:73-103)
//   SA   divsufsort64                                      (esa.cxx:73-75, external library)
//   LCP  PHI/Kasai                                         (esa.cxx:305-347)
//   CLD  stack child table                                 (esa.cxx:256-298)
//   FVC  S[SA[i] + LCP[i]]                                 (esa.cxx:239-250)
//   6-mer interval cache                                   (esa.cxx:90-228)
//
// GPU formulation (all streaming except where noted):
//   1. k_build_text      S and its zero padding, alphabet check, G/C count.
//   2. k_make_keys       key per suffix = its first c characters in 3-bit codes
//                        (end < '!' < '#' < A < C < G < T, i.e. unsigned byte order);
//                        c from the text length (16 for a 5 Mbp reference, at most 21).
//   3. radix sort        (key, index) pairs, ceil(3c / 8) passes of 8 bits (primitives.cuh).
//   4. k_keys_to_lcp     neighbours with different keys give LCP (count of equal leading
//                        codes) and FVC (the next code of the right neighbour) straight
//                        from the sorted keys — no random access.  Equal keys mark a tie;
//                        the first member of every tie group is appended to a list.
//   5a. k_small_groups   tie groups of up to 8 suffixes: sorted by direct comparison, with
//                        their LCP and FVC (all there is on non-repetitive text).
//   5b. refinement       what is left (repeats): prefix doubling on the compacted set —
//                        key = (group rank, rank of suffix + h), h = c, 2c, 4c, … — until
//                        every group is a singleton; ranks live in an ISA array.  Random
//                        access, but over the tied fraction only.  k_tie_lcp: their LCP/FVC.
//   6. k_pyramid_level, k_cld, k_cld_long   child table from its closed form (cld_search.h).
//   7. k_pack_nodes, k_table   interleaved records and K-mer table for the descents
//                        (esa_search.h).
#include "cld_search.h"
#include "esa_device.h"
#include "esa_search.h"
#include "primitives.cuh"
#include "suffix_sort.cuh"

#include <algorithm>
#include <vector>

namespace phy
{

namespace
{

constexpr int KEY_CHARS = 21; // most characters a 63-bit key can hold

// ---------------------------------------------------------------- text

// flags[0]: bad byte seen; flags[1]: number of G/C bytes (gc_content, sequence.cxx:152-165);
// flags[2]: number of '!' (contig separators), which bounds the dirty suffixes of suffix_sort.cuh
__global__ void k_build_text(const uint8_t *__restrict__ ref, int32_t n, uint8_t *__restrict__ S, int32_t padded,
                             int *__restrict__ flags)
{
	const int32_t m = 2 * n + 1;
	int gc = 0, bangs = 0;
	// 16 bytes of S per thread and iteration, one 128-bit store (padded is a multiple of 256)
	for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i0 < padded;
	     i0 += (int64_t)gridDim.x * blockDim.x * 16) {
		uint8_t out[16];
#pragma unroll
		for (int t = 0; t < 16; t++) {
			const int64_t i = i0 + t;
			uint8_t c = 0;
			if (i < n) {
				c = ref[i];
				if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == '!')) atomicExch(flags, 1);
				gc += ((c & 'G' & 'C') == ('G' & 'C'));
				bangs += (c == '!');
			} else if (i == n) {
				c = '#';
			} else if (i < m) {
				c = ref[2 * (int64_t)n - i]; // S[n+1+k] = comp(R[n-1-k])
				if (c >= 'A') c ^= (c & 2) ? 4 : 21;
			}
			out[t] = c;
		}
		*reinterpret_cast<uint4 *>(S + i0) = *reinterpret_cast<const uint4 *>(out);
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		gc += __shfl_xor_sync(0xffffffffu, gc, d);
	if ((threadIdx.x & 31) == 0 && gc) atomicAdd(flags + 1, gc);
	if (__any_sync(0xffffffffu, bangs != 0)) {
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			bangs += __shfl_xor_sync(0xffffffffu, bangs, d);
		if ((threadIdx.x & 31) == 0) atomicAdd(flags + 2, bangs);
	}
}

// ---------------------------------------------------------------- keys

constexpr int KEY_THREADS = 256;
constexpr int KEY_ITEMS = 8;
constexpr int KEY_TILE = KEY_THREADS * KEY_ITEMS;

__global__ void __launch_bounds__(KEY_THREADS)
k_make_keys(const uint8_t *__restrict__ S, int32_t m, int32_t padded, uint64_t *__restrict__ keys, int kc)
{
	const uint64_t key_mask = (1ull << (3 * kc)) - 1; // kc <= 21 characters, 3 bits each
	__shared__ __align__(16) uint8_t sm[KEY_TILE + 32];
	const int64_t base = (int64_t)blockIdx.x * KEY_TILE;
	for (int o = threadIdx.x * 4; o < KEY_TILE + 32; o += KEY_THREADS * 4) {
		uint32_t w = 0;
		if (base + o + 3 < padded) w = *reinterpret_cast<const uint32_t *>(S + base + o);
		*reinterpret_cast<uint32_t *>(sm + o) = w;
	}
	__syncthreads();
	const int p = threadIdx.x * KEY_ITEMS;
	uint64_t key = 0;
#pragma unroll
	for (int t = 0; t < KEY_CHARS; t++)
		if (t < kc) key = (key << 3) | text_code(sm[p + t]);
	uint64_t out[KEY_ITEMS];
	out[0] = key;
#pragma unroll
	for (int k = 1; k < KEY_ITEMS; k++) {
		key = ((key << 3) & key_mask) | text_code(sm[p + k + kc - 1]);
		out[k] = key;
	}
	const int64_t i0 = base + p;
	if (i0 + KEY_ITEMS <= m) {
		ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(keys + i0);
#pragma unroll
		for (int k = 0; k < KEY_ITEMS; k += 2)
			dst[k / 2] = make_ulonglong2(out[k], out[k + 1]);
	} else {
		for (int k = 0; k < KEY_ITEMS; k++)
			if (i0 + k < m) keys[i0 + k] = out[k];
	}
}

// ---------------------------------------------------------------- LCP / FVC from sorted keys

constexpr int32_t LCP_TIE = -2;

// heads[] receives the first SA index of every tie group: the block's heads in one stretch, ONE
// atomic per block.  (One per warp was one atomic on a single address per 32
// suffixes: at m = 5 * 10^8, where a tenth of the suffixes tie on 16 characters, 12 M of them —
// most of the 15 ms this phase took.)  Called by all threads of the block.  Blocks of 256: with
// 1024 the two barriers cost the streaming part ~15 us at m = 10^7 (LCP phase 0.077 -> 0.098 ms).
constexpr int LCP_THREADS = 256;

__device__ __forceinline__ void append_heads(bool head, int64_t j, int32_t *__restrict__ heads,
                                             uint32_t *__restrict__ counters)
{
	__shared__ uint32_t warp_first[LCP_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t bal = __ballot_sync(0xffffffffu, head);
	if (lane == 0) warp_first[warp] = (uint32_t)__popc(bal);
	__syncthreads();
	if (warp == 0) {
		const uint32_t c = lane < (int)(blockDim.x >> 5) ? warp_first[lane] : 0u;
		uint32_t inc = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += o;
		}
		const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
		uint32_t base = 0;
		if (lane == 0 && total) base = atomicAdd(&counters[0], total);
		base = __shfl_sync(0xffffffffu, base, 0);
		warp_first[lane] = base + inc - c;
	}
	__syncthreads();
	if (head) heads[warp_first[warp] + __popc(bal & ((1u << lane) - 1))] = (int32_t)j;
}

__global__ void k_keys_to_lcp(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ sa,
                              const uint8_t *__restrict__ S, int32_t m, int32_t *__restrict__ SA,
                              int32_t *__restrict__ LCP, uint8_t *__restrict__ FVC, int kc,
                              int32_t *__restrict__ heads, uint32_t *__restrict__ counters)
{
	const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool head = false; // first member of a tie group (same key as the next suffix, not as the previous)
	if (j < m) {
		const uint64_t key = keys[j];
		const uint32_t pos = sa[j];
		SA[j] = (int32_t)pos;
		const uint64_t x = j > 0 ? key ^ keys[j - 1] : 1;
		if (j == 0) {
			LCP[0] = -1; // esa.cxx:313-314
			LCP[m] = -1;
			FVC[0] = pos > 0 ? S[pos - 1] : 0; // esa.cxx:247-248 reads S[SA[0] + LCP[0]] with LCP[0] = -1
		} else if (x == 0) {
			LCP[j] = LCP_TIE;
		} else {
			// equal leading 3-bit codes: the key occupies bits [0, 3 kc)
			const int l = (__clzll((long long)x) - (64 - 3 * kc)) / 3;
			LCP[j] = l;
			FVC[j] = text_char((uint32_t)(key >> (3 * (kc - 1 - l))) & 7u);
		}
		head = x != 0 && j + 1 < m && keys[j + 1] == key;
	}
	append_heads(head, j, heads, counters);
}

// The same for the packed words of suffix_sort.cuh (2-bit codes in the top half, dirty flag,
// index).  Next to a dirty suffix LCP and FVC come from the text: its key is not its text.
__global__ void k_words_to_lcp(const uint64_t *__restrict__ words, const uint8_t *__restrict__ S, int32_t m,
                               int32_t *__restrict__ SA, int32_t *__restrict__ LCP, uint8_t *__restrict__ FVC,
                               int32_t *__restrict__ heads, uint32_t *__restrict__ counters)
{
	const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool head = false;
	if (j < m) {
		const uint64_t e = words[j];
		const uint32_t key = (uint32_t)(e >> 32), low = (uint32_t)e, pos = low & PK_INDEX;
		SA[j] = (int32_t)pos;
		bool first_of_key = true;
		if (j == 0) {
			LCP[0] = -1; // esa.cxx:313-314
			LCP[m] = -1;
			FVC[0] = pos > 0 ? S[pos - 1] : 0; // esa.cxx:247-248 reads S[SA[0] + LCP[0]] with LCP[0] = -1
		} else {
			const uint64_t ep = words[j - 1];
			const uint32_t x = key ^ (uint32_t)(ep >> 32);
			if ((low | (uint32_t)ep) & PK_DIRTY) {
				const uint8_t *a = S + ((uint32_t)ep & PK_INDEX), *b = S + pos;
				int32_t l = 0;
				while (a[l] == b[l]) // two different suffixes; S ends in zeros
					l++;
				LCP[j] = l;
				FVC[j] = b[l];
			} else if (x == 0) {
				LCP[j] = LCP_TIE;
				first_of_key = false;
			} else {
				const int l = __clz((int)x) >> 1;
				LCP[j] = l;
				FVC[j] = (uint8_t)(0x54474341u >> (8 * ((key >> (30 - 2 * l)) & 3u)));
			}
		}
		// dirty suffixes lead their key group and are final; a tie group is made of clean ones
		head = first_of_key && !(low & PK_DIRTY) && j + 1 < m && (uint32_t)(words[j + 1] >> 32) == key;
	}
	append_heads(head, j, heads, counters);
}

// ---------------------------------------------------------------- small tie groups
//
// On non-repetitive text the few suffixes that tie on their sort key come in groups of
// two or three whose order is decided a handful of characters later.  Those are finished
// right here — one thread per group, insertion sort by direct comparison, with their LCP
// and FVC — so that the rank array and the doubling rounds are only needed when something
// harder (a real repeat) is left over.

constexpr int SMALL_GROUP = 8;    // largest group handled by direct comparison
// blocks of 128 threads per SM for k_small_groups: as many as its 40 registers per thread allow
// (a thread chases dependent random loads).  At m = 5 * 10^8 — 27 M groups — 12 instead of 4 made
// no measurable difference: the time of that phase went into collecting the heads (append_heads).
constexpr int SMALL_GROUP_BLOCKS = 12;
constexpr int SMALL_COMPARE = 256; // characters beyond the key a comparison may look at

// returns the number of equal characters beyond offset `from`, or -1 when the cap is hit;
// less = suffix a sorts before suffix b (the zeros behind S make the shorter one smaller)
__device__ __forceinline__ int32_t suffix_compare(const uint8_t *__restrict__ S, int32_t a, int32_t b, int32_t from,
                                                  bool &less)
{
	const uint8_t *pa = S + a + from, *pb = S + b + from;
	for (int32_t t = 0; t < SMALL_COMPARE; t++) {
		const uint8_t ca = pa[t], cb = pb[t];
		if (ca != cb) {
			less = ca < cb;
			return t;
		}
	}
	return -1;
}

// counters[0] = number of groups, counters[1] = groups left for the doubling rounds.
// PACKED: keys[] holds the words of suffix_sort.cuh, the sort key is their top half.
template <bool PACKED>
__global__ void k_small_groups(const int32_t *__restrict__ heads, uint32_t *__restrict__ counters,
                               const uint64_t *__restrict__ keys, const uint8_t *__restrict__ S, int32_t m,
                               int32_t *__restrict__ SA, int32_t *__restrict__ LCP, uint8_t *__restrict__ FVC, int kc)
{
	const uint32_t groups = counters[0];
	for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
		const int32_t j0 = heads[g];
		const uint64_t key = PACKED ? keys[j0] >> 32 : keys[j0];
		int32_t size = 1;
		while (size <= SMALL_GROUP && j0 + size < m && (PACKED ? keys[j0 + size] >> 32 : keys[j0 + size]) == key)
			size++;
		bool hard = size > SMALL_GROUP;
		int32_t sa[SMALL_GROUP];
		if (!hard) {
			for (int32_t t = 0; t < size; t++)
				sa[t] = SA[j0 + t];
			for (int32_t t = 1; t < size && !hard; t++) { // insertion sort
				const int32_t x = sa[t];
				int32_t u = t;
				while (u > 0) {
					bool less;
					if (suffix_compare(S, x, sa[u - 1], kc, less) < 0) {
						hard = true;
						break;
					}
					if (!less) break;
					sa[u] = sa[u - 1];
					u--;
				}
				sa[u] = x;
			}
		}
		if (hard) {
			atomicAdd(&counters[1], 1u); // its LCP markers stay: the doubling path picks it up
			continue;
		}
		SA[j0] = sa[0];
		if (j0 == 0) FVC[0] = sa[0] > 0 ? S[sa[0] - 1] : 0; // esa.cxx:247-248 with LCP[0] = -1, for the final SA[0]
		for (int32_t t = 1; t < size; t++) {
			bool less;
			const int32_t l = kc + suffix_compare(S, sa[t - 1], sa[t], kc, less);
			SA[j0 + t] = sa[t];
			LCP[j0 + t] = l;
			FVC[j0 + t] = S[sa[t] + l];
		}
	}
}

// ---------------------------------------------------------------- refinement helpers

__global__ void k_gather_refine_keys(const uint32_t *__restrict__ sa_c, const uint32_t *__restrict__ grank_c,
                                     const int32_t *__restrict__ ISA, int32_t m, int32_t h, int shift, uint32_t count,
                                     uint64_t *__restrict__ keys_c)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= count) return;
	const int64_t nxt = (int64_t)sa_c[r] + h;
	const uint64_t k2 = nxt < m ? (uint64_t)ISA[nxt] + 1 : 0; // a suffix that ends sorts first
	keys_c[r] = ((uint64_t)grank_c[r] << shift) | k2;
}

__global__ void k_tie_lcp(const uint32_t *__restrict__ slots, uint32_t count, const int32_t *__restrict__ SA,
                          const uint8_t *__restrict__ S, int32_t *__restrict__ LCP, uint8_t *__restrict__ FVC, int kc)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= count) return;
	const uint32_t j = slots[r];
	if (j == 0) { // the first suffix was part of a tie group: its FVC quirk value follows the final SA[0]
		FVC[0] = SA[0] > 0 ? S[SA[0] - 1] : 0;
		return;
	}
	if (LCP[j] != LCP_TIE) return; // group head: its LCP came from the keys
	const uint8_t *a = S + SA[j - 1];
	const uint8_t *b = S + SA[j];
	int32_t l = kc;
	// S is followed by zeros; two different suffixes never reach them at the same offset
	while (a[l] == b[l])
		l++;
	LCP[j] = l;
	FVC[j] = b[l];
}

// ---------------------------------------------------------------- CLD

__global__ void k_pyramid_level(const int32_t *__restrict__ in, int32_t n_in, int32_t *__restrict__ out, int32_t n_out)
{
	// one element per thread (coalesced), one output per warp
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int32_t mn = t < n_in ? in[t] : 0x7fffffff;
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
	const int64_t o = t >> 5;
	if ((threadIdx.x & 31) == 0 && o < n_out) out[o] = mn;
}

// Child table, one entry per thread (closed form of cld_search.h).  The block stages its
// stretch of LCP plus a halo in shared memory and builds a small sparse table over it
// (ST[k][t] = min of 2^k entries from t).  "First entry to the right / last entry to the left
// that is <= v" then takes 8 galloping steps, the same for every lane, and a range minimum
// two loads; entries whose answer lies more than 255 positions away (about 1 in 100) are
// queued for k_cld_long.
constexpr int CLD_THREADS = 256;
constexpr int CLD_ITEMS = 4;
constexpr int CLD_TILE = CLD_THREADS * CLD_ITEMS; // entries per block
constexpr int CLD_HALO = 128;
constexpr int CLD_SPAN = CLD_TILE + 2 * CLD_HALO + 1;
constexpr int CLD_NEAR = 4; // neighbours looked at directly before the sparse table is asked

// Range structure over the staged stretch of LCP for the entries whose answer is not next door:
// minima of blocks of CLD_BLK positions and a sparse table over those (windows of 1 .. 32
// blocks).  Searching in steps of blocks and finishing inside one block costs a few probes more
// per query than a sparse table over all positions, but building it costs a quarter — and the
// build, not the queries, was what this kernel spent its time on (profiles/).
constexpr int CLD_BLK = 4;
constexpr int CLD_NB = (CLD_SPAN + CLD_BLK - 1) / CLD_BLK;
constexpr int CLD_BLEVELS = 6;
struct alignas(16) CldTable {
	int32_t v[CLD_NB * CLD_BLK];        // the values; positions outside [0, m] hold INT_MAX
	int32_t st[CLD_BLEVELS][CLD_NB + 1]; // st[k][b] = min of blocks b .. b + 2^k - 1

	// first position >= from (left == false) or last position <= from (left == true) with
	// value <= val, or -1 if it is not within reach
	__device__ __forceinline__ int gallop(int from, int32_t val, bool left) const
	{
		if (from < 0 || from >= CLD_SPAN) return -1;
		int pos = from;
		if (!left) {
			// the rest of from's block, then whole blocks, then inside the block that has it
#pragma unroll
			for (int k = 0; k < CLD_BLK - 1; k++) {
				if ((pos & (CLD_BLK - 1)) == 0) break;
				if (v[pos] <= val) return pos;
				pos++;
			}
			int b = pos / CLD_BLK;
#pragma unroll
			for (int k = CLD_BLEVELS - 1; k >= 0; k--)
				if (b + (1 << k) <= CLD_NB && st[k][b] > val) b += 1 << k;
			if (b >= CLD_NB || st[0][b] > val) return -1;
			pos = b * CLD_BLK;
#pragma unroll
			for (int k = 0; k < CLD_BLK - 1; k++)
				if (v[pos] > val) pos++;
			return pos < CLD_SPAN ? pos : -1;
		}
#pragma unroll
		for (int k = 0; k < CLD_BLK - 1; k++) {
			if ((pos & (CLD_BLK - 1)) == CLD_BLK - 1) break;
			if (v[pos] <= val) return pos;
			if (--pos < 0) return -1;
		}
		int b = pos / CLD_BLK; // blocks b, b - 1, ... lie wholly at or left of `from`
#pragma unroll
		for (int k = CLD_BLEVELS - 1; k >= 0; k--)
			if (b - (1 << k) + 1 >= 0 && st[k][b - (1 << k) + 1] > val) b -= 1 << k;
		if (b < 0 || st[0][b] > val) return -1;
		pos = b * CLD_BLK + CLD_BLK - 1;
#pragma unroll
		for (int k = 0; k < CLD_BLK - 1; k++)
			if (v[pos] > val) pos--;
		return pos;
	}
	// minimum over [a, b], 1 <= b - a + 1 <= 256 + CLD_BLK
	__device__ __forceinline__ int32_t range_min(int a, int b) const
	{
		const int ba = (a + CLD_BLK - 1) / CLD_BLK, bb = (b + 1) / CLD_BLK - 1; // whole blocks inside
		int32_t mn = 0x7fffffff;
		if (ba > bb) {
			for (int q = a; q <= b; q++) // fewer than two blocks' worth
				mn = min(mn, v[q]);
			return mn;
		}
		for (int q = a; q < ba * CLD_BLK; q++)
			mn = min(mn, v[q]);
		for (int q = (bb + 1) * CLD_BLK; q <= b; q++)
			mn = min(mn, v[q]);
		int k = 31 - __clz(bb - ba + 1);
		if (k > CLD_BLEVELS - 1) k = CLD_BLEVELS - 1;
		// two windows of 2^k blocks cover up to 2^(k+1) of them; longer ranges (k capped) take more
		for (int w = ba; w <= bb; w += 1 << k) {
			const int at = w + (1 << k) - 1 <= bb ? w : bb - (1 << k) + 1;
			mn = min(mn, st[k][at]);
		}
		return mn;
	}
};

// Also writes the interleaved records the descent reads (EsaNode: SA, LCP, CLD, FVC), so
// that no separate packing pass re-reads the arrays.
__global__ void __launch_bounds__(CLD_THREADS)
k_cld(Pyramid py, int32_t m, int32_t *__restrict__ CLD, int32_t *__restrict__ long_list, uint32_t *__restrict__ long_count,
      const int32_t *__restrict__ SA, const uint8_t *__restrict__ FVC, EsaNode *__restrict__ node,
      const int *__restrict__ skip)
{
	__shared__ CldTable T;
	if (skip && *skip) return; // the speculative build found out that LCP is not final yet
	const int32_t *__restrict__ LCP = py.level[0];
	const int64_t tile0 = (int64_t)blockIdx.x * CLD_TILE;
	const int64_t lo = tile0 - CLD_HALO; // global index of window position 0
	for (int t = threadIdx.x; t < CLD_NB * CLD_BLK; t += CLD_THREADS) {
		const int64_t g = lo + t;
		T.v[t] = (t < CLD_SPAN && g >= 0 && g <= m) ? LCP[g] : 0x7fffffff;
	}
	__syncthreads();
	for (int b = threadIdx.x; b < CLD_NB; b += CLD_THREADS) {
		const int4 q = *reinterpret_cast<const int4 *>(&T.v[b * CLD_BLK]);
		T.st[0][b] = min(min(q.x, q.y), min(q.z, q.w));
	}
	for (int k = 1; k < CLD_BLEVELS; k++) {
		__syncthreads();
		const int half = 1 << (k - 1);
		for (int b = threadIdx.x; b < CLD_NB; b += CLD_THREADS)
			T.st[k][b] = (b + 2 * half <= CLD_NB) ? min(T.st[k - 1][b], T.st[k - 1][b + half]) : 0x7fffffff;
	}
	__syncthreads();
	// Three out of four answers lie within CLD_NEAR positions (LCP values of neighbouring suffixes
	// are small and close): those are found by looking at the neighbours directly.  The others
	// are queued in shared memory and go through the sparse table afterwards, densely packed —
	// a warp in which one lane needs the 8-level gallops would otherwise make all 32 pay for them.
	__shared__ uint16_t hard[CLD_TILE];
	__shared__ int nhard;
	if (threadIdx.x == 0) nhard = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	auto emit = [&](int64_t i, int32_t a, int res) { // res: window position of the answer, or -1 = far
		const int32_t cld = res >= 0 ? (int32_t)(lo + res) : 0; // far entries: k_cld_long fills it in
		if (res >= 0) CLD[i] = cld;
		const int32_t lcp_of_cld = res >= 0 ? T.v[res] : 0x7ffffff0; // far: no hint yet
		reinterpret_cast<int4 *>(node)[i] = make_int4(SA[i], a, cld, esa_pack_fvc(FVC[i], lcp_of_cld));
	};
	uint32_t my_hard = 0; // bit r: item r of this thread needs the sparse table
#pragma unroll
	for (int r = 0; r < CLD_ITEMS; r++) {
		const int64_t i = tile0 + r * CLD_THREADS + threadIdx.x;
		if (i < m) {
			const int t = (int)(i - lo);
			const int32_t a = T.v[t], b = T.v[t + 1];
			const bool up = b < a;
			// up: last position left of t with LCP <= b; down: first position right of t with LCP <= a
			const int32_t v = up ? b : a;
			const int dir = up ? -1 : 1;
			int x = -1;
#pragma unroll
			for (int k = CLD_NEAR; k >= 1; k--)
				if (T.v[t + dir * k] <= v) x = t + dir * k; // the nearest one wins (k counts down)
			if (x >= 0) {
				int res;
				if (!up && T.v[x] == a) {
					res = x; // the next l-index
				} else {
					// leftmost minimum of (x, t] (up) or (t, x) (down): at most CLD_NEAR entries
					const int from = up ? x + 1 : t + 1, to = up ? t : x - 1;
					res = from;
					int32_t best = T.v[from];
#pragma unroll
					for (int k = 1; k < CLD_NEAR; k++) {
						const int q = from + k;
						if (q <= to) {
							const int32_t val = T.v[q];
							if (val < best) {
								best = val;
								res = q;
							}
						}
					}
				}
				emit(i, a, res);
			} else {
				my_hard |= 1u << r;
			}
		} else if (i == m) {
			CLD[i] = 0;
			reinterpret_cast<int4 *>(node)[i] = make_int4(0, LCP[m], 0, esa_pack_fvc(0, LCP[0]));
		}
	}
	if (my_hard) { // one shared-memory atomic per thread reserves room in the queue
		int at = atomicAdd(&nhard, __popc(my_hard));
#pragma unroll
		for (int r = 0; r < CLD_ITEMS; r++)
			if (my_hard & (1u << r)) hard[at++] = (uint16_t)(r * CLD_THREADS + threadIdx.x);
	}
	__syncthreads();
	const int n_hard = nhard;
	for (int q0 = 0; q0 < n_hard; q0 += CLD_THREADS) {
		const int q = q0 + threadIdx.x;
		bool far = false;
		int64_t i = 0;
		if (q < n_hard) {
			i = tile0 + hard[q];
			const int t = (int)(i - lo);
			const int32_t a = T.v[t], b = T.v[t + 1];
			// One code path for both cases (lanes of a warp are a mix of them, and divergent
			// branches would run one after the other): a gallop to the left for p = last position
			// left of i with LCP <= b ("up", b < a), or to the right for s = first position right
			// of i with LCP <= a; then the leftmost minimum of (p, i] or (i, s) — unless LCP[s] == a,
			// when s itself is the next l-index.
			const bool up = b < a;
			int res = -1;
			const int x = T.gallop(up ? t - 1 : t + 1, up ? b : a, up);
			if (x >= 0) {
				if (!up && T.v[x] == a) {
					res = x;
				} else {
					const int from = up ? x + 1 : t + 1, to = up ? t : x - 1;
					res = T.gallop(from, T.range_min(from, to), false);
				}
			}
			emit(i, a, res);
			far = res < 0;
		}
		// far away: queue the entry for the warp-cooperative kernel (one atomic per warp)
		const uint32_t bal = __ballot_sync(0xffffffffu, far);
		if (bal) {
			uint32_t base = 0;
			if (lane == __ffs(bal) - 1) base = atomicAdd(long_count, (uint32_t)__popc(bal));
			base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
			if (far) long_list[base + __popc(bal & ((1u << lane) - 1))] = (int32_t)i;
		}
	}
}

// ---- warp-cooperative pyramid searches (all 32 lanes call with identical arguments) ----

__device__ __forceinline__ int32_t coop_first_le_right(const Pyramid &py, int32_t from, int32_t v)
{
	const int lane = threadIdx.x & 31;
	int32_t lvl = 0, idx = from;
	for (;;) {
		const int32_t base = idx & ~31;
		const int32_t pos = base + lane;
		const int32_t val = (pos >= idx && pos < py.size[lvl]) ? py.level[lvl][pos] : 0x7fffffff;
		const uint32_t bal = __ballot_sync(0xffffffffu, val <= v);
		if (bal) {
			idx = base + (__ffs(bal) - 1);
			break;
		}
		idx = (base + 32) >> 5;
		lvl++;
	}
	while (lvl > 0) {
		lvl--;
		const int32_t base = idx << 5;
		const int32_t pos = base + lane;
		const int32_t val = pos < py.size[lvl] ? py.level[lvl][pos] : 0x7fffffff;
		const uint32_t bal = __ballot_sync(0xffffffffu, val <= v);
		idx = base + (__ffs(bal) - 1);
	}
	return idx;
}

__device__ __forceinline__ int32_t coop_last_le_left(const Pyramid &py, int32_t from, int32_t v)
{
	const int lane = threadIdx.x & 31;
	int32_t lvl = 0, idx = from;
	for (;;) {
		const int32_t base = idx & ~31;
		const int32_t pos = base + lane;
		const int32_t val = (pos <= idx) ? py.level[lvl][pos] : 0x7fffffff;
		const uint32_t bal = __ballot_sync(0xffffffffu, val <= v);
		if (bal) {
			idx = base + (31 - __clz(bal));
			break;
		}
		idx = (base >> 5) - 1;
		lvl++;
	}
	while (lvl > 0) {
		lvl--;
		const int32_t base = idx << 5;
		const int32_t pos = base + lane;
		const int32_t val = pos < py.size[lvl] ? py.level[lvl][pos] : 0x7fffffff;
		const uint32_t bal = __ballot_sync(0xffffffffu, val <= v);
		idx = base + (31 - __clz(bal));
	}
	return idx;
}

__device__ __forceinline__ int32_t coop_range_min(const Pyramid &py, int32_t a, int32_t b)
{
	const int lane = threadIdx.x & 31;
	int32_t mn = 0x7fffffff;
	int32_t lo = a, hi = b + 1, lvl = 0;
	while (lo < hi) {
		const int32_t *L = py.level[lvl];
		int32_t lend = (lo + 31) & ~31;
		if (lend > hi) lend = hi;
		if (lo + lane < lend) mn = min(mn, L[lo + lane]); // fewer than 32 entries
		lo = lend;
		int32_t hbeg = hi & ~31;
		if (hbeg < lo) hbeg = lo;
		if (hbeg + lane < hi) mn = min(mn, L[hbeg + lane]);
		hi = hbeg;
		lo >>= 5;
		hi >>= 5;
		lvl++;
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
	return mn;
}

// entries whose scans are long: one warp each, 32 positions per step
__global__ void __launch_bounds__(256)
k_cld_long(Pyramid py, const int32_t *__restrict__ long_list, const uint32_t *__restrict__ long_count,
           int32_t *__restrict__ CLD, EsaNode *__restrict__ node, const int *__restrict__ skip)
{
	if (skip && *skip) return;
	const uint32_t count = *long_count;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	const int32_t *__restrict__ LCP = py.level[0];
	for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < count; k += nwarps) {
		const int32_t i = long_list[k];
		const int32_t a = LCP[i], b = LCP[i + 1];
		int32_t res;
		if (b < a) {
			const int32_t p = coop_last_le_left(py, i, b);
			const int32_t mv = coop_range_min(py, p + 1, i);
			res = coop_first_le_right(py, p + 1, mv);
		} else {
			const int32_t s = coop_first_le_right(py, i + 1, a);
			if (LCP[s] == a) {
				res = s;
			} else {
				const int32_t mv = coop_range_min(py, i + 1, s - 1);
				res = coop_first_le_right(py, i + 1, mv);
			}
		}
		if ((threadIdx.x & 31) == 0) {
			CLD[i] = res;
			node[i].cld = res;
			node[i].fvc = esa_pack_fvc((uint8_t)node[i].fvc, LCP[res]);
		}
	}
}

// ---------------------------------------------------------------- table

__device__ __forceinline__ void table_put(TableRec *dst, const TableRec &r)
{
	int4 *d = reinterpret_cast<int4 *>(dst);
	d[0] = make_int4(r.ij.l, r.ij.i, r.ij.j, r.ij.m);
	d[1] = make_int4(r.ni.sa, r.ni.lcp, r.ni.cld, r.ni.fvc);
	d[2] = make_int4(r.nm.sa, r.nm.lcp, r.nm.cld, r.nm.fvc);
	d[3] = make_int4(r.np.sa, r.np.lcp, r.np.cld, r.np.fvc);
}

__global__ void k_table(EsaView e, int32_t K, TableRec *__restrict__ table, const int *__restrict__ skip)
{
	if (skip && *skip) return;
	const uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
	if (code >= (1u << (2 * K))) return;
	table_put(table + code, esa_table_record(e, esa_table_entry(e, code, K)));
}

// The table level by level (esa_search.h, esa_table_extend).  Levels 0 .. k0 are small and
// are made by one block, level L in buf[L & 1]; every further level is one launch with one
// descent step per entry.  final_out: this level is the table itself, only `cur` is kept.
constexpr int TABLE_HEAD_LEVELS = 6;

__device__ __forceinline__ TableBuild table_load(const TableBuild *p)
{
	const int4 a = __ldg(reinterpret_cast<const int4 *>(p)), b = __ldg(reinterpret_cast<const int4 *>(p) + 1);
	TableBuild t;
	t.cur = Interval{a.x, a.y, a.z, a.w};
	t.par = Interval{b.x, b.y, b.z, b.w};
	return t;
}

__device__ __forceinline__ void table_store(TableBuild *p, const TableBuild &t)
{
	reinterpret_cast<int4 *>(p)[0] = make_int4(t.cur.l, t.cur.i, t.cur.j, t.cur.m);
	reinterpret_cast<int4 *>(p)[1] = make_int4(t.par.l, t.par.i, t.par.j, t.par.m);
}

__global__ void __launch_bounds__(1024)
k_table_head(EsaView e, int32_t k0, TableBuild *buf0, TableBuild *buf1, TableRec *__restrict__ final_out)
{
	TableBuild *buf[2] = {buf0, buf1};
	if (threadIdx.x == 0) buf[0][0] = esa_table_root(e);
	__syncthreads();
	for (int32_t k = 0; k < k0; k++) {
		const TableBuild *prev = buf[k & 1];
		TableBuild *next = buf[(k + 1) & 1];
		const uint32_t entries = 1u << (2 * (k + 1));
		for (uint32_t code = threadIdx.x; code < entries; code += blockDim.x) {
			const TableBuild r = esa_table_extend(e, prev[code >> 2], k, (uint8_t)(0x54474341u >> (8 * (code & 3))));
			if (final_out && k + 1 == k0)
				table_put(final_out + code, esa_table_record(e, r.cur));
			else
				next[code] = r;
		}
		__syncthreads(); // one block: the level is visible to all its threads
	}
	if (k0 == 0 && final_out && threadIdx.x == 0) table_put(final_out, esa_table_record(e, buf[0][0].cur));
}

__global__ void __launch_bounds__(256)
k_table_level(EsaView e, int32_t k, const TableBuild *__restrict__ prev, TableBuild *__restrict__ next,
              TableRec *__restrict__ final_out)
{
	const uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
	if (code >= (1u << (2 * (k + 1)))) return;
	const TableBuild r = esa_table_extend(e, table_load(prev + (code >> 2)), k, (uint8_t)(0x54474341u >> (8 * (code & 3))));
	if (final_out)
		table_put(final_out + code, esa_table_record(e, r.cur));
	else
		table_store(next + code, r);
}

int bits_for(uint64_t v)
{
	int b = 0;
	while ((1ull << b) <= v && b < 63)
		b++;
	return b;
}

struct Timer {
	cudaEvent_t a, b;
	cudaStream_t s;
	bool on;
	Timer(cudaStream_t st, bool enabled) : s(st), on(enabled)
	{
		if (!on) return;
		CUDA_CHECK(cudaEventCreate(&a));
		CUDA_CHECK(cudaEventCreate(&b));
		CUDA_CHECK(cudaEventRecord(a, s));
	}
	float lap()
	{
		if (!on) return 0.f;
		CUDA_CHECK(cudaEventRecord(b, s));
		CUDA_CHECK(cudaEventSynchronize(b));
		float ms = 0;
		CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
		CUDA_CHECK(cudaEventRecord(a, s));
		return ms;
	}
	~Timer()
	{
		if (!on) return;
		cudaEventDestroy(a);
		cudaEventDestroy(b);
	}
};

} // namespace

namespace
{
// fail[0] = 1 if what the speculative build took for granted does not hold: valid input, the
// dirty suffixes fit their list and were put in place, no tie group left for the doubling rounds
__global__ void k_spec_check(const int *__restrict__ text_flags, const uint32_t *__restrict__ counters,
                             const uint32_t *__restrict__ dirty_ctl, uint32_t dirty_cap, int *__restrict__ fail,
                             uint32_t *__restrict__ counters_out)
{
	*fail = (text_flags[0] != 0 || counters[1] != 0 || (dirty_ctl && (dirty_ctl[0] > dirty_cap || dirty_ctl[1] != 0))) ? 1 : 0;
	counters_out[0] = counters[0];
	counters_out[1] = counters[1];
	counters_out[2] = dirty_ctl ? dirty_ctl[0] : 0u;
	counters_out[3] = dirty_ctl ? dirty_ctl[1] : 0u;
}
} // namespace

int esa_default_k(int32_t m)
{
	// about one table record per 4..16 suffixes; 4^K records of 16 bytes.  One level more
	// makes the walk 20 % faster on B200 but costs ~5 random sectors per record to build:
	// worth it from a few dozen queries per index on (phylo_process decides, capi.cu)
	int k = 1;
	while (k < 12 && (1ll << (2 * (k + 1))) * 4 <= (int64_t)m)
		k++;
	return k;
}

namespace
{
__global__ void k_pack_nodes(const int32_t *__restrict__ SA, const int32_t *__restrict__ LCP,
                             const int32_t *__restrict__ CLD, const uint8_t *__restrict__ FVC, int32_t m,
                             EsaNode *__restrict__ node)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > m) return;
	int4 v;
	v.x = i < m ? SA[i] : 0;
	v.y = LCP[i];
	v.z = CLD[i];
	v.w = esa_pack_fvc(i < m ? FVC[i] : 0, LCP[CLD[i]]);
	reinterpret_cast<int4 *>(node)[i] = v;
}
} // namespace

void esa_build_table(EsaDevice &esa, int kmer_k, cudaStream_t s, bool nodes_ready, const int *skip)
{
	// the descent reads interleaved records; the build writes them with the child table, an
	// imported index packs them here (13 B read, 16 B written per suffix)
	if (!nodes_ready) {
		esa.node.alloc((size_t)esa.m + 1, s);
		k_pack_nodes<<<div_up((int64_t)esa.m + 1, 256), 256, 0, s>>>(esa.SA.get(), esa.LCP.get(), esa.CLD.get(),
		                                                            esa.FVC.get(), esa.m, esa.node.get());
		KERNEL_CHECK();
	}
	int K = kmer_k < 0 ? esa_default_k(esa.m) : kmer_k;
	if (K > 12) K = 12;
	esa.K = 0;
	esa.table.release();
	if (K <= 0) return;
	const uint32_t entries = 1u << (2 * K);
	esa.table.alloc(entries, s);
	EsaView v = esa.view();
	v.K = 0;
	if (g_tuning.table_direct != 2) {
		// every entry by its own descent from the root: the upper levels of neighbouring
		// entries are the same cache lines, so this is as fast as or faster than the level-wise
		// build with its extra launches (measured on B200: 0.097 vs 0.193 ms at K = 10, 0.262 vs
		// 0.265 ms at K = 11; a head + one-launch tail variant: 0.149 / 0.302); same table
		k_table<<<div_up(entries, 128), 128, 0, s>>>(v, K, esa.table.get(), skip);
		KERNEL_CHECK();
	} else {
		if (skip) throw std::invalid_argument("the level-wise table build does not run speculatively");
		const int32_t k0 = K < TABLE_HEAD_LEVELS ? K : TABLE_HEAD_LEVELS;
		// level L lives in buf[L & 1]; the largest stored level is K - 1
		const size_t big = (size_t)1 << (2 * (K > 0 ? K - 1 : 0)), small = K > 1 ? big / 4 : 1;
		const size_t head = (size_t)1 << (2 * k0);
		DevBuf<TableBuild> buf_a(std::max(((K - 1) & 1) ? small : big, head), s), buf_b(std::max(((K - 1) & 1) ? big : small, head), s);
		TableBuild *buf[2] = {buf_a.get(), buf_b.get()};
		k_table_head<<<1, 1024, 0, s>>>(v, k0, buf[0], buf[1], k0 == K ? esa.table.get() : nullptr);
		KERNEL_CHECK();
		for (int32_t k = k0; k < K; k++) {
			const uint32_t n_next = 1u << (2 * (k + 1));
			k_table_level<<<div_up(n_next, 256), 256, 0, s>>>(v, k, buf[k & 1], buf[(k + 1) & 1],
			                                                  k + 1 == K ? esa.table.get() : nullptr);
			KERNEL_CHECK();
		}
	}
	esa.K = K;
}

int esa_default_key_chars(int32_t m)
{
	// enough characters that ties are rare (4^c >= 16 m), rounded up to whole radix passes
	int need = 2;
	while (need < KEY_CHARS && (1ll << (2 * need)) < (int64_t)m)
		need++;
	need += 2; // with one character less a pass is saved at 5 Mbp, but the ~14 % of suffixes that
	           // then tie cost more in k_small_groups than the pass (measured on B200)
	const int passes = (3 * need + 7) / 8;
	const int c = (8 * passes) / 3;
	return c > KEY_CHARS ? KEY_CHARS : c;
}

int esa_default_packed_chars(int32_t m)
{
	// whole radix passes of four 2-bit characters; 4^c >= 16 m keeps ties rare
	int c = 4;
	while (c < PK_MAX_CHARS && (1ll << (2 * c)) < 16 * (int64_t)m)
		c += 4;
	return c;
}

namespace
{
bool esa_build_impl(EsaDevice &esa, const uint8_t *d_ref, int32_t n, int kmer_k, int key_chars, cudaStream_t s,
                    EsaTimings *tm, bool spec);
}

// The build normally runs without a single host round trip until its end ("speculative"): it
// takes for granted that the input is valid, that the few suffixes next to separators fit
// their list, and that no group of suffixes ties beyond what the direct comparisons settle —
// true for every genome that is not highly repetitive — and the kernels downstream of those
// assumptions skip their work if a device-side check says otherwise.  One read-back at the end
// tells; if an assumption failed the index is built again the careful way, deciding on the host
// after each stage like before.
void esa_build_device(EsaDevice &esa, const uint8_t *d_ref, int32_t n, int kmer_k, int key_chars, cudaStream_t s,
                      EsaTimings *tm, bool lazy)
{
	const bool timed = tm && tm->enabled; // per-phase timers synchronise anyway
	const bool spec = !timed && g_tuning.esa_speculative && g_tuning.table_direct != 2;
	if (spec) {
		esa_build_impl(esa, d_ref, n, kmer_k, key_chars, s, tm, true);
		// the first kernel's verdict has long arrived (the host was busy queueing the rest)
		CUDA_CHECK(cudaEventSynchronize(esa.ev_early));
		if (esa.h_report[0]) {
			esa.pending = false;
			throw std::invalid_argument("reference contains bytes outside {A,C,G,T,!}");
		}
		esa.gc_count = esa.h_report[1];
		if (!lazy) esa_finish(esa, s, tm);
		return;
	}
	esa_build_impl(esa, d_ref, n, kmer_k, key_chars, s, tm, false);
}

bool esa_finish(EsaDevice &esa, cudaStream_t s, EsaTimings *tm)
{
	if (!esa.pending) return false;
	CUDA_CHECK(cudaEventSynchronize(esa.ev_done));
	esa.pending = false;
	const int *h = esa.h_report + 8;
	if (tm) {
		tm->tie_groups = (uint32_t)h[4];
		tm->dirty = (uint32_t)h[6];
	}
	if (!h[3]) return false;
	// separators beyond the list, or repeats: once more, step by step.  The reference is the
	// first n bytes of the text that the failed build has left behind.
	DevBuf<uint8_t> text = std::move(esa.S);
	esa_build_impl(esa, text.get(), esa.n, esa.pend_kmer_k, esa.pend_key_chars, s, tm, false);
	return true;
}

namespace
{
// returns false (speculative mode only) if the index has to be built again without assumptions
bool esa_build_impl(EsaDevice &esa, const uint8_t *d_ref, int32_t n, int kmer_k, int key_chars, cudaStream_t s,
                    EsaTimings *tm, bool spec)
{
	if (n < 1) throw std::invalid_argument("reference is empty");
	if ((int64_t)n * 2 + 1 > 0x7fffffffll - 128) throw std::invalid_argument("reference too long for 32-bit indices");
	const int32_t m = 2 * n + 1;
	const int32_t padded = ((m + 64 + 255) / 256) * 256;
	EsaTimings local;
	EsaTimings &T = tm ? *tm : local;
	const bool timed = tm && tm->enabled; // event timers synchronise; off unless asked for
	T = EsaTimings();
	T.enabled = timed;
	Timer total(s, timed);
	Timer lap(s, timed);

	esa.release();
	esa.n = n;
	esa.m = m;
	esa.S.alloc(padded, s);
	esa.SA.alloc(m, s);
	esa.LCP.alloc((size_t)m + 1, s);
	esa.CLD.alloc((size_t)m + 1, s);
	esa.FVC.alloc(m, s);

	if (!esa.side) {
		CUDA_CHECK(cudaStreamCreateWithFlags(&esa.side, cudaStreamNonBlocking));
		CUDA_CHECK(cudaEventCreateWithFlags(&esa.ev_fork, cudaEventDisableTiming));
		CUDA_CHECK(cudaEventCreateWithFlags(&esa.ev_join, cudaEventDisableTiming));
	}
	if (!esa.h_report) {
		CUDA_CHECK(cudaHostAlloc((void **)&esa.h_report, 16 * sizeof(int), cudaHostAllocPortable));
		CUDA_CHECK(cudaEventCreateWithFlags(&esa.ev_text, cudaEventDisableTiming));
		CUDA_CHECK(cudaEventCreateWithFlags(&esa.ev_early, cudaEventDisableTiming));
		CUDA_CHECK(cudaEventCreateWithFlags(&esa.ev_done, cudaEventDisableTiming));
	}

	// 1. text
	// 8 report words (EsaDevice::report) and, cleared by the same memset, the build's counters:
	// [8,9] dirty_ctl, [10,11] tie-group counters, [12] length of k_cld's list of long entries
	esa.report.alloc(16, s);
	esa.report.zero();
	int *bad = esa.report.get(); // [0..2], see EsaDevice::report
	k_build_text<<<NUM_SMS_B200 * 8, 256, 0, s>>>(d_ref, n, esa.S.get(), padded, bad);
	KERNEL_CHECK();
	int64_t bangs = 0;
	if (!spec) {
		PinnedArena::Scope scope(g_pinned);
		int *h_flags = g_pinned.take<int>(3);
		CUDA_CHECK(cudaMemcpyAsync(h_flags, bad, 3 * sizeof(int), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaStreamSynchronize(s));
		if (h_flags[0]) throw std::invalid_argument("reference contains bytes outside {A,C,G,T,!}");
		esa.gc_count = h_flags[1];
		bangs = h_flags[2];
	} else {
		// the verdict on the input and the G/C count go to the host right away, next to the sort
		CUDA_CHECK(cudaEventRecord(esa.ev_text, s));
		CUDA_CHECK(cudaStreamWaitEvent(esa.side, esa.ev_text, 0));
		CUDA_CHECK(cudaMemcpyAsync(esa.h_report, bad, 3 * sizeof(int), cudaMemcpyDeviceToHost, esa.side));
		CUDA_CHECK(cudaEventRecord(esa.ev_early, esa.side));
	}
	T.text_ms = lap.lap();
	// Speculative builds have no host decision left in them: everything behind the first radix pass
	// is recorded and submitted as one graph (GraphSegment, common.cuh).
	GraphSegment *const graph = (spec && g_tuning.esa_graph && g_tuning.sort_path != 1) ? &esa.build_graph : nullptr;
	struct GraphGuard {
		GraphSegment *g;
		~GraphGuard()
		{
			if (g) g->abandon();
		}
	} graph_guard{graph};
	int *spec_fail = esa.report.get() + 3; // speculative mode: set by k_spec_check, read by the kernels after it
	const int *skip = spec ? spec_fail : nullptr;
	uint32_t *spec_counters = reinterpret_cast<uint32_t *>(esa.report.get() + 4); // copies of {groups, hard groups, dirty suffixes, dirty error}

	// Which sorter: packed words (2-bit codes, <= 16 characters, suffix_sort.cuh) unless the
	// caller asks for longer keys, the text has so many separators that ordering the dirty
	// suffixes pairwise would cost more than it saves, or 16 characters are hopelessly few.
	// 16 suffixes in front of every byte below 'A'; not known yet when speculating: room for ~1000 contigs
	const int64_t dirty_bound = spec ? PK_DIRTY_CAP : 16 * (2 * bangs + 2);
	const bool packed = g_tuning.sort_path != 1 && key_chars <= PK_MAX_CHARS && dirty_bound <= PK_DIRTY_CAP && m <= (1 << 30);
	int kc;
	if (packed) {
		kc = key_chars > 0 ? key_chars : esa_default_packed_chars(m);
	} else {
		kc = key_chars > 0 ? key_chars : esa_default_key_chars(m);
		if (kc > KEY_CHARS) kc = KEY_CHARS;
	}
	T.key_chars = kc;
	T.packed = packed;

	{
		// 2. keys, 3. sort
		const size_t key_words = packed ? pk_padded_words(m) : (size_t)m;
		DevBuf<uint64_t> keys(key_words, s), keys_alt(key_words, s);
		DevBuf<uint32_t> vals, vals_alt, dirty_list, dirty_sorted;
		uint32_t *const dirty_ctl = reinterpret_cast<uint32_t *>(esa.report.get() + 8); // [0] number of dirty suffixes, [1] error flag of k_dirty_fix
		const uint64_t *K1 = nullptr;
		if (packed) {
			const uint32_t cap = (uint32_t)dirty_bound;
			dirty_list.alloc(cap, s);
			dirty_sorted.alloc(cap, s);
			PkProfile prof;
			uint64_t *W = suffix_sort_packed(esa.S.get(), m, padded, kc, keys.get(), keys_alt.get(), dirty_list.get(),
			                                 dirty_ctl, cap, s, timed ? &prof : nullptr, graph);
			if (timed) {
				T.first_pass_ms = prof.first_ms;
				if (prof.passes) {
					T.sort_passes = prof.passes;
					T.hist_ms_avg = prof.hist_ms / prof.passes;
					T.scan_ms_avg = prof.scan_ms / prof.passes;
					T.scatter_ms_avg = prof.scatter_ms / prof.passes;
				}
			}
			const int dirty_blocks = std::min(div_up(cap, 8), 64); // a warp per dirty suffix, grid-stride
			k_dirty_rank<<<dirty_blocks, 256, 0, s>>>(dirty_list.get(), dirty_ctl, cap, esa.S.get(), dirty_sorted.get());
			KERNEL_CHECK();
			k_dirty_fix<<<dirty_blocks, 256, 0, s>>>(dirty_sorted.get(), dirty_ctl, cap, esa.S.get(), m, kc, W,
			                                         (int *)(dirty_ctl + 1));
			KERNEL_CHECK();
			K1 = W;
			T.sort_ms = lap.lap();
		} else {
			vals.alloc(m, s);
			vals_alt.alloc(m, s);
			k_make_keys<<<div_up(m, KEY_TILE), KEY_THREADS, 0, s>>>(esa.S.get(), m, padded, keys.get(), kc);
			KERNEL_CHECK();
			T.keys_ms = lap.lap();
			RsProfile prof;
			const bool flipped = radix_sort_pairs(keys.get(), vals.get(), keys_alt.get(), vals_alt.get(), m, 0, 3 * kc, true, s,
			                                      timed ? &prof : nullptr);
			if (timed && prof.passes) {
				T.sort_passes = prof.passes;
				T.hist_ms_avg = prof.hist_ms / prof.passes;
				T.scan_ms_avg = prof.scan_ms / prof.passes;
				T.scatter_ms_avg = prof.scatter_ms / prof.passes;
			}
			K1 = flipped ? keys_alt.get() : keys.get();
			T.sort_ms = lap.lap();
		}

		// 4. LCP/FVC from neighbouring keys; SA in its final place for all untied suffixes
		DevBuf<int32_t> heads((size_t)m / 2 + 1, s);
		uint32_t *const counters = reinterpret_cast<uint32_t *>(esa.report.get() + 10);
		if (packed) {
			k_words_to_lcp<<<div_up(m, LCP_THREADS), LCP_THREADS, 0, s>>>(K1, esa.S.get(), m, esa.SA.get(), esa.LCP.get(), esa.FVC.get(),
			                                              heads.get(), counters);
		} else {
			const uint32_t *V1 = K1 == keys.get() ? vals.get() : vals_alt.get();
			k_keys_to_lcp<<<div_up(m, LCP_THREADS), LCP_THREADS, 0, s>>>(K1, V1, esa.S.get(), m, esa.SA.get(), esa.LCP.get(),
			                                             esa.FVC.get(), kc, heads.get(), counters);
		}
		KERNEL_CHECK();

		// 5a. small tie groups by direct comparison
		uint32_t h_counters[2] = {0, 0};
		{
			if (packed)
				k_small_groups<true><<<NUM_SMS_B200 * SMALL_GROUP_BLOCKS, 128, 0, s>>>(heads.get(), counters, K1, esa.S.get(), m,
				                                                     esa.SA.get(), esa.LCP.get(), esa.FVC.get(), kc);
			else
				k_small_groups<false><<<NUM_SMS_B200 * SMALL_GROUP_BLOCKS, 128, 0, s>>>(heads.get(), counters, K1, esa.S.get(), m,
				                                                      esa.SA.get(), esa.LCP.get(), esa.FVC.get(), kc);
			KERNEL_CHECK();
			if (spec) {
				// no read-back: a device-side check decides whether the kernels below may run, the
				// host learns about it at the very end (counters travel with the other flags)
				// (it also copies the counters next to the flags for the read-back)
				k_spec_check<<<1, 1, 0, s>>>(bad, counters, packed ? dirty_ctl : nullptr, (uint32_t)dirty_bound,
				                             spec_fail, spec_counters);
				KERNEL_CHECK();
			} else {
				PinnedArena::Scope scope(g_pinned);
				uint32_t *hp = g_pinned.take<uint32_t>(4);
				uint32_t *h_dirty = hp + 2;
				h_dirty[0] = h_dirty[1] = 0;
				if (packed) CUDA_CHECK(cudaMemcpyAsync(h_dirty, dirty_ctl, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
				CUDA_CHECK(cudaMemcpyAsync(hp, counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
				CUDA_CHECK(cudaStreamSynchronize(s));
				h_counters[0] = hp[0];
				h_counters[1] = hp[1];
				T.tie_groups = h_counters[0];
				T.dirty = h_dirty[0];
				if (h_dirty[1] || h_dirty[0] > (uint64_t)dirty_bound)
					throw std::runtime_error("suffix sort: dirty suffixes inconsistent with their key groups");
			}
		}

		// 5b. refinement of the tie groups that are left (repeats): prefix doubling
		const int32_t *LCP = esa.LCP.get();
		const int32_t *SAcur = esa.SA.get();
		DevBuf<uint32_t> d_count(1, s);
		DevBuf<uint32_t> slots0, slots, sa_c, grank_c;
		uint32_t count = 0;
		if (h_counters[1]) {
			// first a count, then buffers of the right size
			auto tied = [LCP, m] __device__(int64_t j) {
				return LCP[j] == LCP_TIE || (j + 1 < m && LCP[j + 1] == LCP_TIE);
			};
			device_select(m, tied, [] __device__(int64_t, uint32_t) {}, d_count.get(), s);
			count = d2h_scalar(d_count.get(), s);
			T.tied = count;
			if (count) {
				slots0.alloc(count, s);
				sa_c.alloc(count, s);
				uint32_t *sl = slots0.get(), *sc = sa_c.get();
				device_select(
					m, tied,
					[sl, sc, SAcur] __device__(int64_t j, uint32_t r) {
						sl[r] = (uint32_t)j;
						sc[r] = (uint32_t)SAcur[j];
					},
					d_count.get(), s);
			}
		}
		T.lcp_ms = lap.lap();
		if (count) {
			// ISA[SA[j]] = index of the first member of j's group (ties share a rank)
			DevBuf<int32_t> ISA(m, s);
			{
				int32_t *isa = ISA.get();
				device_scan<int32_t>(
					m, [LCP] __device__(int64_t j) { return LCP[j] == LCP_TIE ? 0 : (int32_t)j; },
					[isa, SAcur] __device__(int64_t j, int32_t head) { isa[SAcur[j]] = head; }, OpMax(), 0, true, s);
			}
			grank_c.alloc(count, s);
			{
				uint32_t *gr = grank_c.get();
				const uint32_t *sc = sa_c.get();
				const int32_t *isa = ISA.get();
				device_for(count, [gr, sc, isa] __device__(int64_t r) { gr[r] = (uint32_t)isa[sc[r]]; }, s);
			}
			slots.alloc(count, s);
			CUDA_CHECK(cudaMemcpyAsync(slots.get(), slots0.get(), count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
			const uint32_t count0 = count;
			const int shift = bits_for((uint64_t)m + 1);
			const int key_bits = shift + bits_for((uint64_t)m);
			int64_t h = kc;
			while (count) {
				T.refine_rounds++;
				DevBuf<uint64_t> kc(count, s), kc_alt(count, s);
				DevBuf<uint32_t> sa_alt(count, s);
				k_gather_refine_keys<<<div_up(count, 256), 256, 0, s>>>(
					sa_c.get(), grank_c.get(), ISA.get(), m, (int32_t)(h > 0x7fffffff ? 0x7fffffff : h), shift, count,
					kc.get());
				KERNEL_CHECK();
				const bool fl =
					radix_sort_pairs(kc.get(), sa_c.get(), kc_alt.get(), sa_alt.get(), count, 0, key_bits, false, s);
				const uint64_t *ks = fl ? kc_alt.get() : kc.get();
				const uint32_t *ss = fl ? sa_alt.get() : sa_c.get();
				// new ranks: slot of the first member of each (group, key2) run
				int32_t *isa = ISA.get();
				int32_t *SAo = esa.SA.get();
				const uint32_t *sl = slots.get();
				DevBuf<uint32_t> newrank(count, s);
				uint32_t *nr = newrank.get();
				device_scan<uint32_t>(
					count,
					[ks, sl] __device__(int64_t r) { return (r == 0 || ks[r] != ks[r - 1]) ? sl[r] : 0u; },
					[isa, SAo, ss, sl, nr] __device__(int64_t r, uint32_t rank) {
						const uint32_t suf = ss[r];
						SAo[sl[r]] = (int32_t)suf;
						isa[suf] = (int32_t)rank;
						nr[r] = rank;
					},
					OpMax(), 0u, true, s);
				// keep the members of runs that are still longer than one
				DevBuf<uint32_t> slots_n(count, s), sa_n(count, s), gr_n(count, s);
				uint32_t *sln = slots_n.get(), *san = sa_n.get(), *grn = gr_n.get();
				const uint32_t cnt = count;
				device_select(
					count,
					[ks, cnt] __device__(int64_t r) {
						const bool head = (r == 0 || ks[r] != ks[r - 1]);
						const bool next_head = (r + 1 == cnt || ks[r + 1] != ks[r]);
						return !(head && next_head);
					},
					[sln, san, grn, sl, ss, nr] __device__(int64_t r, uint32_t w) {
						sln[w] = sl[r];
						san[w] = ss[r];
						grn[w] = nr[r];
					},
					d_count.get(), s);
				count = d2h_scalar(d_count.get(), s);
				slots.swap(slots_n);
				sa_c.swap(sa_n);
				grank_c.swap(gr_n);
				h *= 2;
			}
			T.refine_ms = lap.lap();
			// 6. LCP/FVC of tied neighbours
			k_tie_lcp<<<div_up(count0, 128), 128, 0, s>>>(slots0.get(), count0, esa.SA.get(), esa.S.get(),
			                                              esa.LCP.get(), esa.FVC.get(), kc);
			KERNEL_CHECK();
			T.lcp_ms += lap.lap();
		}
	}

	// 7. child table.  The min-pyramid is only read by k_cld_long (the entries whose scans are
	// long): it is built on a side stream while k_cld works through the bulk of the entries.
	{
		cudaStream_t side = esa.side;
		cudaEvent_t ev_fork = esa.ev_fork, ev_join = esa.ev_join;
		Pyramid py;
		std::vector<DevBuf<int32_t>> levels;
		py.level[0] = esa.LCP.get();
		py.size[0] = m + 1;
		py.levels = 1;
		CUDA_CHECK(cudaEventRecord(ev_fork, s));
		CUDA_CHECK(cudaStreamWaitEvent(side, ev_fork, 0));
		// Whatever happens below (a throw unwinds `levels` into the main stream's block cache),
		// the main stream is behind the side stream's kernels before any of those buffers can be
		// handed out again.
		struct SideJoin {
			cudaStream_t main, side;
			cudaEvent_t ev;
			~SideJoin()
			{
				if (cudaEventRecord(ev, side) == cudaSuccess) cudaStreamWaitEvent(main, ev, 0);
			}
		} side_join{s, side, ev_join};
		while (py.size[py.levels - 1] > 1) {
			if (py.levels >= PYR_MAX_LEVELS) throw std::runtime_error("pyramid too deep");
			const int32_t n_in = py.size[py.levels - 1];
			const int32_t n_out = (n_in + 31) / 32;
			levels.emplace_back((size_t)n_out, s);
			k_pyramid_level<<<div_up((int64_t)n_out * 32, 256), 256, 0, side>>>(py.level[py.levels - 1], n_in,
			                                                                    levels.back().get(), n_out);
			KERNEL_CHECK();
			py.level[py.levels] = levels.back().get();
			py.size[py.levels] = n_out;
			py.levels++;
		}
		CUDA_CHECK(cudaEventRecord(ev_join, side));
		DevBuf<int32_t> long_list((size_t)m + 1, s);
		uint32_t *const long_count = reinterpret_cast<uint32_t *>(esa.report.get() + 12);
		esa.node.alloc((size_t)m + 1, s);
		k_cld<<<div_up((int64_t)m + 1, CLD_TILE), CLD_THREADS, 0, s>>>(py, m, esa.CLD.get(), long_list.get(),
		                                                               long_count, esa.SA.get(), esa.FVC.get(),
		                                                               esa.node.get(), skip);
		KERNEL_CHECK();
		CUDA_CHECK(cudaStreamWaitEvent(s, ev_join, 0));
		k_cld_long<<<NUM_SMS_B200 * 8, 256, 0, s>>>(py, long_list.get(), long_count, esa.CLD.get(), esa.node.get(), skip);
		KERNEL_CHECK();
		T.cld_ms = lap.lap();
	}

	// 8. K-mer table
	esa_build_table(esa, kmer_k, s, true, skip);
	T.table_ms = lap.lap();
	T.total_ms = total.lap();
	if (spec) {
		// the one read-back of the speculative build: queued here, looked at by esa_finish()
		CUDA_CHECK(cudaMemcpyAsync(esa.h_report + 8, esa.report.get(), 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
		if (graph && graph->capturing) graph->launch();
		CUDA_CHECK(cudaEventRecord(esa.ev_done, s));
		esa.pending = true;
		esa.pend_kmer_k = kmer_k;
		esa.pend_key_chars = key_chars;
	}
	return true;
}
} // namespace

} // namespace phy
