// First-characters suffix sort on packed 64-bit words — the fast path of the index build.
//
// A suffix is represented by ONE 64-bit word:
//     bits 63..32  its first kc <= 16 characters in 2-bit codes (A < C < G < T), first
//                  character in the top bits
//     bit  31      "dirty": one of those characters is not a nucleotide
//     bits 30..0   the suffix index
// so a radix pass moves 8 bytes per suffix in and 8 out (the general path of primitives.cuh
// moves a 64-bit key of 3-bit codes plus a 32-bit index: 12 + 12), and 16 characters need
// four 8-bit passes instead of six.  The first pass computes the words straight from the
// text (1 byte read per suffix), so the keys are never written out unsorted.
//
// Dirty suffixes.  The text S = R '#' revcomp(R) holds a handful of bytes below 'A': the
// '#', the '!' between contigs and the end of the text (S is zero padded).  A suffix with
// such a byte among its first kc characters, at offset dd, gets the key of its dd leading
// nucleotides followed by zeros.  In the true (unsigned byte) order it sorts before every
// clean suffix that shares those dd nucleotides and after every suffix with a smaller
// prefix, i.e. it belongs at the very start of the group of words with its own key, before
// all clean members.  After the radix passes k_dirty_rank orders the dirty suffixes among
// themselves by direct comparison (there are at most 16 per special byte) and k_dirty_fix
// moves them to the front of their key groups.  From then on they are final singletons;
// LCP and FVC next to them are computed from the text instead of from the keys.
//
// Replaces (together with esa_build.cu) divsufsort64 at /root/reference/src/esa.cxx:74.
#pragma once
#include "primitives.cuh"

namespace phy
{

// geometry of the scatter kernel (switchable for tools/sortbench.cu)
#ifndef PKS_THREADS_N
#define PKS_THREADS_N 256
#endif
#ifndef PKS_MIN_BLOCKS
#define PKS_MIN_BLOCKS 3
#endif
// how a lane learns which lanes of its warp hold the same digit: 0 = eight votes (registers
// only), 1 = one OR into a per-warp table of 256 lane masks in shared memory.  The votes are 44 %
// of the kernel's instructions, the table needs five; measured on B200 (tools/sortbench.cu,
// m = 10^7 / 10^8): votes 49.0 / 400 us, table 55.3 / 467 us — the shared-memory pipe, already
// busy with the conflicts of the exchange, is the scarcer resource.
#ifndef PK_PEERS_SMEM
#define PK_PEERS_SMEM 0
#endif
constexpr int PK_THREADS = 256; // histogram and scan kernels
constexpr int PK_ITEMS = 16;
constexpr int PK_TILE = PK_THREADS * PK_ITEMS; // suffixes per tile
constexpr int PK_WARPS = PK_THREADS / 32;
constexpr int PKS_THREADS = PKS_THREADS_N;      // scatter kernel: same tile, more or fewer items per thread
constexpr int PKS_ITEMS = PK_TILE / PKS_THREADS;
constexpr int PKS_WARPS = PKS_THREADS / 32;
static_assert(PKS_THREADS >= RS_BINS && PKS_THREADS % 32 == 0 && PKS_ITEMS * PKS_THREADS == PK_TILE && PKS_ITEMS % 2 == 0,
              "scatter geometry");
constexpr int PK_WORDS = PK_TILE / 16 + 1; // 16-character text words a tile looks at
constexpr int PK_MAX_CHARS = 16;
constexpr uint32_t PK_DIRTY = 0x80000000u;
constexpr uint32_t PK_INDEX = 0x7fffffffu;

// 16 text bytes -> 32 bits of 2-bit codes (first character on top) and, in the top half of
// `spec`, one bit per character that is not a nucleotide (every byte below 'A' has bit 6 clear)
__device__ __forceinline__ void pk_pack16(const uint4 v, uint32_t &codes, uint32_t &spec)
{
	const uint32_t w[4] = {v.x, v.y, v.z, v.w};
	codes = 0;
	spec = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t x = w[k];
		// A 0x41 -> 0, C 0x43 -> 1, G 0x47 -> 2, T 0x54 -> 3, four bytes at a time
		const uint32_t c = ((x >> 1) & 0x03030303u) ^ ((x >> 2) & 0x01010101u);
		const uint32_t c8 = (c * 0x40100401u) >> 24; // gathers the four fields, lowest address first
		const uint32_t s = (~x >> 6) & 0x01010101u;
		const uint32_t s4 = (s * 0x08040201u) >> 24;
		codes |= c8 << (24 - 8 * k);
		spec |= s4 << (28 - 4 * k);
	}
}

struct PkMasks {
	uint32_t key;    // top 2 kc bits
	uint32_t window; // top kc bits
};

inline PkMasks pk_masks(int kc)
{
	PkMasks mk;
	mk.key = kc >= 16 ? 0xffffffffu : ~0u << (32 - 2 * kc);
	mk.window = ~0u << (32 - kc);
	return mk;
}

// the tile's text as packed words in shared memory; spec[w] covers characters 16 w .. 16 w + 31
__device__ __forceinline__ void pk_stage_text(const uint8_t *__restrict__ S, int64_t tile_base, int32_t padded,
                                              uint32_t *codes, uint32_t *spec)
{
	for (int w = threadIdx.x; w <= PK_WORDS; w += PK_THREADS) {
		const int64_t o = tile_base + 16 * (int64_t)w;
		uint4 a = make_uint4(0, 0, 0, 0), b = a;
		if (o + 16 <= padded) a = *reinterpret_cast<const uint4 *>(S + o);
		if (o + 32 <= padded) b = *reinterpret_cast<const uint4 *>(S + o + 16);
		uint32_t ca, sa, cb, sb;
		pk_pack16(a, ca, sa);
		pk_pack16(b, cb, sb);
		codes[w] = ca;
		spec[w] = sa | (sb >> 16);
	}
}

// packed word of the suffix at tile offset t
__device__ __forceinline__ uint64_t pk_element(const uint32_t *codes, const uint32_t *spec, int t, uint32_t index,
                                               const PkMasks mk)
{
	const int w = t >> 4, k = t & 15;
	uint32_t key = __funnelshift_l(codes[w + 1], codes[w], 2 * k);
	const uint32_t win = (spec[w] << k) & mk.window;
	uint32_t flag = 0;
	if (win) {
		const int dd = __clz(win); // offset of the first special byte
		key &= (uint32_t) ~(0xffffffffull >> (2 * dd));
		flag = PK_DIRTY;
	}
	key &= mk.key;
	return ((uint64_t)key << 32) | flag | index;
}

// the same key from the text (for the few dirty suffixes); the loads do not depend on each other
__device__ __forceinline__ uint32_t pk_key_scalar(const uint8_t *__restrict__ S, uint32_t pos, int kc)
{
	uint8_t c[PK_MAX_CHARS];
#pragma unroll
	for (int t = 0; t < PK_MAX_CHARS; t++)
		c[t] = S[pos + t]; // S is followed by at least 64 zero bytes
	uint32_t key = 0;
	bool live = true;
#pragma unroll
	for (int t = 0; t < PK_MAX_CHARS; t++) {
		live = live && t < kc && c[t] >= 'A';
		const uint32_t code = ((c[t] >> 1) & 3u) ^ ((c[t] >> 2) & 1u);
		if (live) key |= code << (30 - 2 * t);
	}
	return key;
}

// radix digit: shifts are whole bytes of the key half, one PRMT
__device__ __forceinline__ uint32_t pk_digit(uint64_t e, int shift)
{
	return __byte_perm((uint32_t)(e >> 32), 0, 0x4440u | (uint32_t)((shift - 32) >> 3));
}

// ---- asynchronous copies into shared memory (LDGSTS) ------------------------------------

__device__ __forceinline__ void pk_cp_async16(void *smem_dst, const void *gsrc, int src_bytes)
{
	// copies src_bytes (0, 8 or 16) and zero-fills the rest of the 16 bytes
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void pk_cp_async_commit()
{
	asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N> __device__ __forceinline__ void pk_cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

constexpr int PK_RAW_CHUNKS = PK_TILE / 16 + 3; // 16-byte pieces of text a tile looks at

// Every kernel below works on whole tiles: the word buffers are padded to a multiple of
// PK_TILE and the pass that reads the text fills the padding with all-ones words, which are
// the largest in every pass and (the sort being stable) stay behind the m real ones.

// counts[tile * 256 + d]
template <bool FROM_TEXT>
static __global__ void __launch_bounds__(PK_THREADS)
pk_histogram(const uint64_t *__restrict__ in, const uint8_t *__restrict__ S, int32_t padded, int64_t n, int shift,
             PkMasks mk, uint32_t *__restrict__ counts)
{
	__shared__ uint32_t h[RS_BINS];
	__shared__ uint32_t codes[FROM_TEXT ? PK_WORDS + 2 : 1], spec[FROM_TEXT ? PK_WORDS + 2 : 1];
	h[threadIdx.x] = 0;
	const int64_t base = (int64_t)blockIdx.x * PK_TILE;
	if (FROM_TEXT) pk_stage_text(S, base, padded, codes, spec);
	__syncthreads();
	if (FROM_TEXT) {
#pragma unroll
		for (int r = 0; r < PK_ITEMS; r++) {
			const int t = r * PK_THREADS + threadIdx.x;
			const uint64_t e = base + t < n ? pk_element(codes, spec, t, 0, mk) : ~0ull;
			atomicAdd(&h[pk_digit(e, shift)], 1u);
		}
	} else {
		// two words per load
		const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in + base);
		ulonglong2 v[PK_ITEMS / 2];
#pragma unroll
		for (int r = 0; r < PK_ITEMS / 2; r++)
			v[r] = src[r * PK_THREADS + threadIdx.x];
#pragma unroll
		for (int r = 0; r < PK_ITEMS / 2; r++) {
			atomicAdd(&h[pk_digit(v[r].x, shift)], 1u);
			atomicAdd(&h[pk_digit(v[r].y, shift)], 1u);
		}
	}
	__syncthreads();
	counts[(int64_t)blockIdx.x * RS_BINS + threadIdx.x] = h[threadIdx.x];
}

// counts[tile * 256 + d] -> number of words with digit d in the tiles before `tile` (in
// place); totals[d] = number of words with digit d.  One block per 8 digits: a thread reads
// the 8 counts of a tile as one 32-byte sector, four tiles per round.
constexpr int PK_SCAN_THREADS = 1024;
constexpr int PK_SCAN_WARPS = PK_SCAN_THREADS / 32;
constexpr int PK_SCAN_TILES = 4; // tiles per thread and round: 4096 tiles (16.7 M suffixes) per round
static __global__ void __launch_bounds__(PK_SCAN_THREADS)
pk_scan_counts(uint32_t *__restrict__ counts, int ntiles, uint32_t *__restrict__ totals)
{
	__shared__ uint32_t wsum[PK_SCAN_WARPS][8], wtot[8];
	static_assert(PK_SCAN_WARPS == 32, "the warp totals are scanned by one warp");
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t carry[8];
#pragma unroll
	for (int k = 0; k < 8; k++)
		carry[k] = 0;
	for (int base = 0; base < ntiles; base += PK_SCAN_THREADS * PK_SCAN_TILES) {
		uint32_t v[PK_SCAN_TILES][8];
		uint32_t sum[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
			sum[k] = 0;
#pragma unroll
		for (int u = 0; u < PK_SCAN_TILES; u++) {
			const int t = base + threadIdx.x * PK_SCAN_TILES + u;
			uint4 a = make_uint4(0, 0, 0, 0), b = a;
			if (t < ntiles) {
				const uint4 *src = reinterpret_cast<const uint4 *>(counts + (int64_t)t * RS_BINS + 8 * blockIdx.x);
				a = src[0];
				b = src[1];
			}
			v[u][0] = a.x, v[u][1] = a.y, v[u][2] = a.z, v[u][3] = a.w;
			v[u][4] = b.x, v[u][5] = b.y, v[u][6] = b.z, v[u][7] = b.w;
#pragma unroll
			for (int k = 0; k < 8; k++)
				sum[k] += v[u][k];
		}
		// exclusive scan of sum[] over the block, 8 sequences at once
		uint32_t inc[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			inc[k] = sum[k];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(0xffffffffu, inc[k], d);
				if (lane >= d) inc[k] += o;
			}
			if (lane == 31) wsum[warp][k] = inc[k];
		}
		__syncthreads();
		if (warp == 0) { // warp totals -> exclusive prefixes, block totals
#pragma unroll
			for (int k = 0; k < 8; k++) {
				const uint32_t x = wsum[lane][k];
				uint32_t y = x;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t o = __shfl_up_sync(0xffffffffu, y, d);
					if (lane >= d) y += o;
				}
				wsum[lane][k] = y - x;
				if (lane == 31) wtot[k] = y;
			}
		}
		__syncthreads();
		uint32_t total[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			total[k] = wtot[k];
			inc[k] += wsum[warp][k] + carry[k] - sum[k]; // now: exclusive prefix of this thread's first tile
		}
#pragma unroll
		for (int u = 0; u < PK_SCAN_TILES; u++) {
			const int t = base + threadIdx.x * PK_SCAN_TILES + u;
			if (t < ntiles) {
				uint4 *dst = reinterpret_cast<uint4 *>(counts + (int64_t)t * RS_BINS + 8 * blockIdx.x);
				dst[0] = make_uint4(inc[0], inc[1], inc[2], inc[3]);
				dst[1] = make_uint4(inc[4], inc[5], inc[6], inc[7]);
			}
#pragma unroll
			for (int k = 0; k < 8; k++)
				inc[k] += v[u][k];
		}
#pragma unroll
		for (int k = 0; k < 8; k++)
			carry[k] += total[k];
		__syncthreads();
	}
	if (threadIdx.x == 0) {
#pragma unroll
		for (int k = 0; k < 8; k++)
			totals[8 * blockIdx.x + k] = carry[k];
	}
}

// lanes of the warp whose 8-bit digit equals this lane's.  Per bit: a vote, a select (all
// ones if this lane's bit is clear) and one three-input logic operation,
// peers &= ballot ^ select; the predicates of all bits come from one R2P.  The kernel is
// bound by the integer/logic pipe, so the instruction count here is what matters.
__device__ __forceinline__ uint32_t pk_warp_peers(uint32_t d)
{
	uint32_t peers;
	asm volatile("{\n\t"
	             ".reg .pred p;\n\t"
	             ".reg .b32 b, t, nm;\n\t"
	             "and.b32 t, %1, 1;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "xor.b32 %0, b, nm;\n\t"
	             "and.b32 t, %1, 2;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "lop3.b32 %0, %0, b, nm, 0x60;\n\t"
	             "and.b32 t, %1, 4;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "lop3.b32 %0, %0, b, nm, 0x60;\n\t"
	             "and.b32 t, %1, 8;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "lop3.b32 %0, %0, b, nm, 0x60;\n\t"
	             "and.b32 t, %1, 16;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "lop3.b32 %0, %0, b, nm, 0x60;\n\t"
	             "and.b32 t, %1, 32;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "lop3.b32 %0, %0, b, nm, 0x60;\n\t"
	             "and.b32 t, %1, 64;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "lop3.b32 %0, %0, b, nm, 0x60;\n\t"
	             "and.b32 t, %1, 128;\n\t"
	             "setp.ne.u32 p, t, 0;\n\t"
	             "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
	             "selp.b32 nm, 0, 0xffffffff, p;\n\t"
	             "lop3.b32 %0, %0, b, nm, 0x60;\n\t"
	             "}"
	             : "=r"(peers)
	             : "r"(d));
	return peers;
}

template <bool FROM_TEXT> struct PkSmem {
	// words: staging of the incoming tile, then its digit-sorted order; double buffered when
	// the input is words (the next tile streams in while this one is ranked)
	uint64_t buf[FROM_TEXT ? 1 : 2][PK_TILE];
	uint16_t warp_cnt[PKS_WARPS][RS_BINS]; // per-warp digit counts, then the warp's first slot in the tile order
	uint32_t gdelta[RS_BINS];              // global position of (digit, tile) minus its slot in the tile order
	uint32_t digit_base[RS_BINS];          // words with a smaller digit
	uint32_t scan_tmp[2][8];               // warp totals of the digit scan, by tile parity
	uint4 raw[FROM_TEXT ? 2 : 1][FROM_TEXT ? PK_RAW_CHUNKS : 1]; // text of this and the next tile
	uint32_t codes[FROM_TEXT ? PK_WORDS + 2 : 1];
	uint32_t spec[FROM_TEXT ? PK_WORDS + 2 : 1];
	uint64_t full[2]; // transaction barriers of the two word buffers (bulk copies)
};

// One stable radix pass over packed words, persistent blocks: block b takes tiles b,
// b + gridDim.x, ...; the input of the next tile (and its 256 offsets) is on its way while
// the current one is ranked, exchanged through shared memory and written out so that every
// digit's run leaves as one contiguous store.  A tile of words (32 KB) comes in as ONE bulk
// copy issued by one thread (cp.async.bulk, completion on a transaction barrier) instead of
// sixteen cp.async per thread; the text of the first pass, with its ragged end, stays with
// cp.async.  offsets[tile * 256 + d] and totals[d] from pk_scan_counts.  FROM_TEXT: the words are made here from the text (all ones from index n
// on), and the index of every dirty suffix is appended to dirty_list.
// Five block barriers per tile; the kernel is bound by the logic pipe and by shared-memory
// latency, not by HBM (profiles/), so the count of both is what was tuned.
template <bool FROM_TEXT>
static __global__ void __launch_bounds__(PKS_THREADS, PKS_MIN_BLOCKS)
pk_scatter(const uint64_t *__restrict__ in, const uint8_t *__restrict__ S, int32_t padded, uint64_t *__restrict__ out,
           int64_t n, int shift, int ntiles, PkMasks mk, const uint32_t *__restrict__ offsets,
           const uint32_t *__restrict__ totals, uint32_t *__restrict__ dirty_list, uint32_t *__restrict__ dirty_count,
           uint32_t dirty_cap)
{
	extern __shared__ __align__(16) unsigned char pk_smem_raw[];
	PkSmem<FROM_TEXT> &sm = *reinterpret_cast<PkSmem<FROM_TEXT> *>(pk_smem_raw);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t lanes_below = (1u << lane) - 1;

	auto prefetch = [&](int tile, int slot) {
		if (FROM_TEXT) {
			const int64_t tile_base = (int64_t)tile * PK_TILE;
			for (int c = threadIdx.x; c < PK_RAW_CHUNKS; c += PKS_THREADS) {
				const int64_t o = tile_base + 16 * (int64_t)c;
				const bool ok = o + 16 <= padded;
				pk_cp_async16(&sm.raw[slot][c], S + (ok ? o : 0), ok ? 16 : 0);
			}
		} else if (threadIdx.x == 0) {
			// the buffer was read and written by everybody a moment ago (before the last barrier)
			fence_proxy_async();
			mbar_expect_tx(&sm.full[slot], PK_TILE * (uint32_t)sizeof(uint64_t));
			bulk_load(sm.buf[slot], in + (size_t)tile * PK_TILE, PK_TILE * (uint32_t)sizeof(uint64_t), &sm.full[slot]);
		}
	};
	auto zero_counters = [&] {
		for (int b = threadIdx.x; b < PKS_WARPS * RS_BINS / 2; b += PKS_THREADS)
			reinterpret_cast<uint32_t *>(&sm.warp_cnt[0][0])[b] = 0;
	};

	int tile = blockIdx.x;
	uint32_t my_offset = 0;
	if (!FROM_TEXT) {
		if (threadIdx.x == 0) {
			mbar_init(&sm.full[0], 1);
			mbar_init(&sm.full[1], 1);
			mbar_fence_init();
		}
		__syncthreads();
	}
	if (tile < ntiles) {
		prefetch(tile, 0);
		if (threadIdx.x < RS_BINS) my_offset = offsets[(size_t)tile * RS_BINS + threadIdx.x];
	}
	pk_cp_async_commit();
	{
		// digit_base = exclusive scan of the digit totals
		uint32_t v = threadIdx.x < RS_BINS ? totals[threadIdx.x] : 0, inc = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += o;
		}
		if (lane == 31 && warp < 8) sm.scan_tmp[0][warp] = inc;
		__syncthreads();
		if (threadIdx.x < RS_BINS) {
			uint32_t before = 0;
#pragma unroll
			for (int w = 0; w < 8; w++)
				if (w < warp) before += sm.scan_tmp[0][w];
			sm.digit_base[threadIdx.x] = before + inc - v;
		}
	}
	zero_counters();
	uint16_t *const my_cnt = sm.warp_cnt[warp];
	const int warp_base = warp * (32 * PKS_ITEMS);

	for (int it = 0; tile < ntiles; tile += gridDim.x, it++) {
		const int slot = FROM_TEXT ? 0 : (it & 1);
		const uint32_t tile_base = (uint32_t)tile * PK_TILE;
		uint64_t *const buf = sm.buf[slot];

		if (FROM_TEXT)
			pk_cp_async_wait<0>(); // this tile's copies have landed
		else
			mbar_wait(&sm.full[slot], (uint32_t)((it >> 1) & 1)); // use it / 2 of this buffer
		__syncthreads(); // (1) ... for everybody; the previous tile has left; the counters are zero
		// the next tile starts to arrive
		const int next = tile + gridDim.x;
		const uint32_t offset_now = my_offset;
		if (next < ntiles) {
			prefetch(next, (it & 1) ^ 1);
			if (threadIdx.x < RS_BINS) my_offset = offsets[(size_t)next * RS_BINS + threadIdx.x];
		}
		pk_cp_async_commit();
		if (FROM_TEXT) {
			const uint4 *raw = sm.raw[it & 1];
			for (int w = threadIdx.x; w <= PK_WORDS; w += PKS_THREADS) {
				uint32_t ca, sa, cb, sb;
				pk_pack16(raw[w], ca, sa);
				pk_pack16(raw[w + 1], cb, sb);
				sm.codes[w] = ca;
				sm.spec[w] = sa | (sb >> 16);
			}
			__syncthreads();
		}

		// warp-striped: item r of lane l sits at warp_base + r * 32 + l (tile order)
		uint64_t e[PKS_ITEMS];
		uint32_t rank[PKS_ITEMS / 2]; // two 16-bit ranks per register
#pragma unroll
		for (int r = 0; r < PKS_ITEMS; r++) {
			const int t = warp_base + r * 32 + lane;
			if (FROM_TEXT) {
				const uint32_t index = tile_base + t;
				if (index < n) {
					e[r] = pk_element(sm.codes, sm.spec, t, index, mk);
					if ((uint32_t)e[r] & PK_DIRTY) {
						const uint32_t at = atomicAdd(dirty_count, 1u);
						if (at < dirty_cap) dirty_list[at] = index;
					}
				} else {
					e[r] = ~0ull; // padding: the largest word in every pass
				}
			} else {
				e[r] = buf[t];
			}
		}
		// stable rank of every item among the items of its warp with the same digit
#if PK_PEERS_SMEM
		// The warp's stretch of the word buffer is free once its 16 words sit in registers: its
		// first KB serves as a table of lane masks, one per digit.  Every lane ORs its bit into
		// the entry of its digit; what it reads back is the set of lanes with that digit.  The
		// lowest of them bumps the digit's counter and clears the entry for the next row.
		uint32_t *const tbl = reinterpret_cast<uint32_t *>(buf + warp_base);
		__syncwarp();
#pragma unroll
		for (int k = 0; k < RS_BINS / 32; k++)
			tbl[k * 32 + lane] = 0;
		__syncwarp();
#endif
#pragma unroll
		for (int r = 0; r < PKS_ITEMS; r++) {
			const uint32_t d = pk_digit(e[r], shift);
#if PK_PEERS_SMEM
			atomicOr(&tbl[d], 1u << lane);
			__syncwarp();
			const uint32_t peers = tbl[d];
#else
			const uint32_t peers = pk_warp_peers(d);
#endif
			const uint32_t before = my_cnt[d];
			__syncwarp();
			if ((peers & lanes_below) == 0) { // lowest peer
				my_cnt[d] = (uint16_t)(before + __popc(peers));
#if PK_PEERS_SMEM
				tbl[d] = 0;
#endif
			}
			__syncwarp();
			const uint32_t rk = before + __popc(peers & lanes_below);
			if (r & 1)
				rank[r / 2] = __byte_perm(rank[r / 2], rk, 0x5410);
			else
				rank[r / 2] = rk;
		}
		__syncthreads(); // (2)
		// thread d: per-warp counts of digit d -> the warps' first slots; digit d's first slot
		uint32_t c[PKS_WARPS], run = 0, inc = 0;
		if (threadIdx.x < RS_BINS) {
#pragma unroll
			for (int w = 0; w < PKS_WARPS; w++) {
				c[w] = sm.warp_cnt[w][threadIdx.x];
				run += c[w];
			}
			inc = run;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
				if (lane >= d) inc += o;
			}
			if (lane == 31) sm.scan_tmp[it & 1][warp] = inc;
		}
		__syncthreads(); // (3)
		if (threadIdx.x < RS_BINS) {
			uint32_t slot0 = inc - run;
#pragma unroll
			for (int w = 0; w < 8; w++)
				if (w < warp) slot0 += sm.scan_tmp[it & 1][w];
			sm.gdelta[threadIdx.x] = sm.digit_base[threadIdx.x] + offset_now - slot0;
#pragma unroll
			for (int w = 0; w < PKS_WARPS; w++) {
				sm.warp_cnt[w][threadIdx.x] = (uint16_t)slot0;
				slot0 += c[w];
			}
		}
		__syncthreads(); // (4)
		// exchange: every word to its place in the tile's digit order
#pragma unroll
		for (int r = 0; r < PKS_ITEMS; r++)
			buf[my_cnt[pk_digit(e[r], shift)] + ((r & 1) ? rank[r / 2] >> 16 : rank[r / 2] & 0xffffu)] = e[r];
		__syncthreads(); // (5)
		zero_counters();
		// every digit's run leaves as one contiguous store
#pragma unroll
		for (int r = 0; r < PKS_ITEMS; r++) {
			const uint32_t p = r * PKS_THREADS + threadIdx.x;
			const uint64_t w = buf[p];
			out[sm.gdelta[pk_digit(w, shift)] + p] = w;
		}
	}
	pk_cp_async_wait<0>();
}

// ---- dirty suffixes ---------------------------------------------------------------------

// a < b as suffixes of S (unsigned bytes; S is followed by zeros, so the shorter one is smaller)
__device__ __forceinline__ bool pk_suffix_less(const uint8_t *__restrict__ S, uint32_t a, uint32_t b)
{
	const uint8_t *pa = S + a, *pb = S + b;
	for (;;) {
		const uint8_t ca = *pa++, cb = *pb++;
		if (ca != cb) return ca < cb;
	}
}

// exact order of the dirty suffixes: one warp per suffix counts the smaller ones
static __global__ void __launch_bounds__(256)
k_dirty_rank(const uint32_t *__restrict__ list, const uint32_t *__restrict__ count, uint32_t cap,
             const uint8_t *__restrict__ S, uint32_t *__restrict__ sorted)
{
	const uint32_t D = *count < cap ? *count : cap; // more than cap: the caller falls back to the general sorter
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	const int lane = threadIdx.x & 31;
	for (uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; x < D; x += nwarps) {
		const uint32_t me = list[x];
		uint32_t smaller = 0;
		for (uint32_t y = lane; y < D; y += 32) {
			const uint32_t other = list[y];
			if (other != me && pk_suffix_less(S, other, me)) smaller++;
		}
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			smaller += __shfl_xor_sync(0xffffffffu, smaller, d);
		if (lane == 0) sorted[smaller] = me;
	}
}

// One warp per run of dirty suffixes with the same key: inside that key's group of the
// sorted words the clean members move up (stable) and the dirty ones take the front, in
// their exact order.  err[0] is set if the group does not hold exactly the run's suffixes.
static __global__ void __launch_bounds__(256)
k_dirty_fix(const uint32_t *__restrict__ sorted, const uint32_t *__restrict__ count, uint32_t cap,
            const uint8_t *__restrict__ S, int32_t m, int kc, uint64_t *__restrict__ words, int *__restrict__ err)
{
	if (*count > cap) return; // the list is incomplete: nothing here can be trusted, the caller starts over
	const uint32_t D = *count;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	const int lane = threadIdx.x & 31;
	for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < D; r += nwarps) {
		const uint32_t key = pk_key_scalar(S, sorted[r], kc);
		if (r > 0 && pk_key_scalar(S, sorted[r - 1], kc) == key) continue; // not the first of its run
		uint32_t cnt = 1;
		while (r + cnt < D && pk_key_scalar(S, sorted[r + cnt], kc) == key)
			cnt++;
		// [a, b): the words with this key.  32 probes per round, one per lane.
		int64_t bound[2];
#pragma unroll
		for (int upper = 0; upper < 2; upper++) {
			int64_t lo = upper ? bound[0] : 0, hi = m; // the answer is in [lo, hi]
			while (lo < hi) {
				const int64_t step = (hi - lo) / 32 + 1;
				const int64_t pos = lo + lane * step;
				bool before = false; // words[pos] sorts before the bound
				if (pos < hi) {
					const uint32_t k = (uint32_t)(words[pos] >> 32);
					before = upper ? k <= key : k < key;
				}
				const int nb = __popc(__ballot_sync(0xffffffffu, before)); // the predicate is monotone
				const int64_t new_lo = nb ? lo + (nb - 1) * step + 1 : lo;
				const int64_t new_hi = nb < 32 && lo + nb * step < hi ? lo + nb * step : hi;
				lo = new_lo;
				hi = new_hi;
			}
			bound[upper] = lo;
		}
		const int64_t a = bound[0], b = bound[1];
		// clean members to the right end, from the right (writes never pass the reads)
		int64_t w = b;
		for (int64_t top = b; top > a; top -= 32) {
			const int64_t j = top - 1 - lane;
			const uint64_t e = j >= a ? words[j] : 0;
			const bool clean = j >= a && !((uint32_t)e & PK_DIRTY);
			const uint32_t bal = __ballot_sync(0xffffffffu, clean);
			__syncwarp();
			if (clean) words[w - 1 - __popc(bal & ((1u << lane) - 1))] = e;
			w -= __popc(bal);
			__syncwarp();
		}
		if (w - a != (int64_t)cnt) {
			if (lane == 0) atomicExch(err, 1);
			continue;
		}
		for (uint32_t t = lane; t < cnt; t += 32)
			words[a + t] = ((uint64_t)key << 32) | PK_DIRTY | sorted[r + t];
	}
}

inline size_t pk_padded_words(int32_t m)
{
	return (size_t)div_up(m, PK_TILE) * PK_TILE;
}

struct PkProfile {
	float hist_ms = 0, scan_ms = 0, scatter_ms = 0; // passes that read packed words
	float first_ms = 0;                               // the pass that reads the text (all three kernels)
	int passes = 0;                                   // passes counted in hist/scan/scatter
};

// Sorts the m suffixes of S by their first kc <= 16 characters.  buf_a and buf_b hold
// pk_padded_words(m) words each.  Returns the buffer (a or b) whose first m words are the
// sorted ones; the dirty suffixes are listed in dirty_list (count on the
// device) and still sit at arbitrary places inside their key groups.
// blocks of pk_scatter<FROM_TEXT> that fit the device at once (persistent grid)
template <bool FROM_TEXT> inline int pk_scatter_grid(int ntiles)
{
	static int resident[64] = {};
	int dev = 0;
	CUDA_CHECK(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64) dev = 0;
	if (!resident[dev]) {
		CUDA_CHECK(cudaFuncSetAttribute(pk_scatter<FROM_TEXT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                                (int)sizeof(PkSmem<FROM_TEXT>)));
		int per_sm = 0, sms = 0;
		CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk_scatter<FROM_TEXT>, PKS_THREADS,
		                                                         sizeof(PkSmem<FROM_TEXT>)));
		CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
		resident[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : NUM_SMS_B200);
	}
	return ntiles < resident[dev] ? ntiles : resident[dev];
}

inline uint64_t *suffix_sort_packed(const uint8_t *S, int32_t m, int32_t padded, int kc, uint64_t *buf_a,
                                    uint64_t *buf_b, uint32_t *dirty_list, uint32_t *dirty_count, uint32_t dirty_cap,
                                    cudaStream_t s, PkProfile *prof, GraphSegment *graph_after_first_pass = nullptr)
{
	const PkMasks mk = pk_masks(kc);
	const int ntiles = div_up(m, PK_TILE);
	const int passes = (2 * kc + 7) / 8;
	const int grid_text = pk_scatter_grid<true>(ntiles), grid_words = pk_scatter_grid<false>(ntiles);
	DevBuf<uint32_t> counts((size_t)ntiles * RS_BINS, s);
	DevBuf<uint32_t> totals((size_t)passes * RS_BINS, s);
	std::vector<cudaEvent_t> evs;
	auto mark = [&] {
		if (!prof) return;
		cudaEvent_t e;
		CUDA_CHECK(cudaEventCreate(&e));
		CUDA_CHECK(cudaEventRecord(e, s));
		evs.push_back(e);
	};
	uint64_t *in = nullptr, *out = buf_a;
	for (int p = 0; p < passes; p++) {
		const int shift = 64 - 8 * (passes - p);
		uint32_t *tot = totals.get() + (size_t)p * RS_BINS;
		mark();
		if (p == 0)
			pk_histogram<true><<<ntiles, PK_THREADS, 0, s>>>(nullptr, S, padded, m, shift, mk, counts.get());
		else
			pk_histogram<false><<<ntiles, PK_THREADS, 0, s>>>(in, nullptr, 0, m, shift, mk, counts.get());
		KERNEL_CHECK();
		mark();
		pk_scan_counts<<<RS_BINS / 8, PK_SCAN_THREADS, 0, s>>>(counts.get(), ntiles, tot);
		KERNEL_CHECK();
		mark();
		if (p == 0)
			pk_scatter<true><<<grid_text, PKS_THREADS, sizeof(PkSmem<true>), s>>>(
				nullptr, S, padded, out, m, shift, ntiles, mk, counts.get(), tot, dirty_list, dirty_count, dirty_cap);
		else
			pk_scatter<false><<<grid_words, PKS_THREADS, sizeof(PkSmem<false>), s>>>(
				in, nullptr, 0, out, m, shift, ntiles, mk, counts.get(), tot, nullptr, nullptr, 0);
		KERNEL_CHECK();
		mark();
		in = out;
		out = (out == buf_a) ? buf_b : buf_a;
		// the first pass keeps the GPU busy while the host records everything behind it
		if (p == 0 && graph_after_first_pass) graph_after_first_pass->begin(s);
	}
	if (prof) {
		CUDA_CHECK(cudaStreamSynchronize(s));
		for (size_t k = 0; k + 3 < evs.size(); k += 4) {
			float a = 0, b = 0, c = 0;
			CUDA_CHECK(cudaEventElapsedTime(&a, evs[k], evs[k + 1]));
			CUDA_CHECK(cudaEventElapsedTime(&b, evs[k + 1], evs[k + 2]));
			CUDA_CHECK(cudaEventElapsedTime(&c, evs[k + 2], evs[k + 3]));
			if (k == 0) {
				prof->first_ms += a + b + c;
			} else {
				prof->hist_ms += a;
				prof->scan_ms += b;
				prof->scatter_ms += c;
				prof->passes++;
			}
		}
		for (auto e : evs)
			cudaEventDestroy(e);
	}
	return in;
}

} // namespace phy
