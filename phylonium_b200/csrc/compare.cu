// All-pairs mismatch / homology counting on reference-coordinate bit-planes, and the
// distance estimators.  Replaces hot loop B of /root/reference/src/process.cxx:524-549.
//
// The reference intersects the two genomes' homology lists on reference coordinates
// (process.cxx:566-611), trims every overlapping pair to the common range (:620-635,
// process.h:119-143) and counts mismatching bytes with seqcmp / revseqcmp.  Because each
// genome's surviving homologies are gap-free diagonals that are disjoint on the
// reference, that is a column-wise comparison of "rows": row_g[p] = the query base
// aligned to reference column p.  k_build_rows materialises the rows as
// bit-planes once per genome (N * n bytes read, 5/8 * N * n written); k_compare_pairs then
// needs only AND/XOR/POPC on 32 columns at a time.
//
// k_compare_tiles: one block of 16 warps per (16 x 16 genome tile, column chunk).  The rows of
// the 32 genomes stream through shared memory in steps of 32 words, double buffered with
// cp.async; every warp owns a 4 x 4 sub-tile whose 16 pair counters live in registers and are
// reduced with shuffles once per chunk.  Tiles none of whose genomes has a reverse-strand
// homology or a '!' (the flag word behind every row) skip the D and B planes: 3 instead of 5
// planes to load and half the logic per pair.  Work units (tile pair, chunk) are dealt
// round-robin to the ranks of a sharded run.
#include "compare_device.h"
#include "tile_order.h"
#include "primitives.cuh"

#include <cuda.h> // CUtensorMap (types only: the encoder is fetched through the runtime)

namespace phy
{

namespace
{

// 32 query bytes (any alignment) -> three 32-bit planes, byte t -> bit t: bit 1 and bit 2 of
// the byte (the 2-bit code (c & 6) >> 1) and "byte is '!'" (the only valid byte with bit 6
// clear).  Reads the aligned words that cover src[0, 32), i.e. up to 3 bytes past src + 32.
__device__ __forceinline__ void planes_of_32_bytes(const uint8_t *src, uint32_t &c0, uint32_t &c1, uint32_t &bang)
{
	const uintptr_t a = reinterpret_cast<uintptr_t>(src);
	const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	uint32_t x[9];
#pragma unroll
	for (int i = 0; i < 9; i++)
		x[i] = __ldg(w + i);
	c0 = c1 = bang = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) {
		const uint32_t y = __funnelshift_r(x[i], x[i + 1], sh); // bytes src[4 i .. 4 i + 3]
		// one flag per byte -> four adjacent bits (multiply gathers them in the top byte)
		const uint32_t f0 = (y >> 1) & 0x01010101u, f1 = (y >> 2) & 0x01010101u, fb = (~y >> 6) & 0x01010101u;
		c0 |= (((f0 * 0x01020408u) >> 24) & 0xfu) << (4 * i);
		c1 |= (((f1 * 0x01020408u) >> 24) & 0xfu) << (4 * i);
		bang |= (((fb * 0x01020408u) >> 24) & 0xfu) << (4 * i);
	}
}

// One thread per 32 reference columns of one genome.  A word that lies inside one homology
// (nearly all of them) is made from 32 contiguous query bytes with word-wide logic; words at
// homology borders go column by column.
__global__ void k_build_rows(uint32_t *__restrict__ rows, int64_t genome_words, int64_t W, int32_t n,
                             int64_t first_row, const uint8_t *__restrict__ Q, const QueryInfo *__restrict__ qi,
                             int32_t count, const Hom *__restrict__ homs, const int64_t *__restrict__ begin,
                             const int64_t *__restrict__ hcount)
{
	const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int32_t k = blockIdx.y;
	if (w >= W || k >= count) return;
	const Hom *H = homs + begin[k];
	const int32_t h = (int32_t)hcount[k];
	const uint8_t *q = Q + qi[k].qoff;
	const int64_t qlen = qi[k].qlen;
	const int64_t c0 = w * 32;
	uint32_t pv = 0, p0 = 0, p1 = 0, pd = 0, pb = 0;
	if (c0 < n && h > 0) {
		// first homology that ends right of c0 (ends increase: the list is sorted and disjoint)
		int32_t lo = 0, hi = h;
		while (lo < hi) {
			const int32_t mid = (lo + hi) >> 1;
			if ((int64_t)H[mid].iproj + H[mid].len <= c0)
				lo = mid + 1;
			else
				hi = mid;
		}
		bool done = false;
		if (lo < h) {
			const Hom hm = H[lo];
			const int64_t start = hm.iproj, end = (int64_t)hm.iproj + hm.len;
			if (start <= c0 && end >= c0 + 32) {
				// query bytes of columns c0 .. c0 + 31, in ascending query order
				const int64_t q0 = hm.dir ? (int64_t)hm.iq + (end - 1 - (c0 + 31)) : (int64_t)hm.iq + (c0 - start);
				if (q0 >= 0 && q0 + 36 <= qlen + 1) { // the aligned loads stay inside the sequence and its terminator
					uint32_t a, b, g;
					planes_of_32_bytes(q + q0, a, b, g);
					if (hm.dir) { // column p holds byte end - 1 - p: reverse, and complement the code
						a = __brev(a);
						b = ~__brev(b);
						g = __brev(g);
						pd = 0xffffffffu;
					}
					pv = 0xffffffffu;
					p0 = a;
					p1 = b;
					pb = g;
					done = true;
				}
			}
		}
		for (int32_t x = lo; !done && x < h && H[x].iproj < c0 + 32; x++) {
			const Hom hm = H[x];
			const int64_t start = hm.iproj, end = (int64_t)hm.iproj + hm.len;
			const int64_t a = start > c0 ? start : c0;
			const int64_t b = end < c0 + 32 ? end : c0 + 32;
			for (int64_t p = a; p < b; p++) {
				const int64_t qpos = hm.dir ? (int64_t)hm.iq + (end - 1 - p) : (int64_t)hm.iq + (p - start);
				const uint8_t c = q[qpos];
				uint32_t code = (c & 6u) >> 1;
				if (hm.dir) code ^= 2u;
				const uint32_t bit = 1u << (uint32_t)(p - c0);
				pv |= bit;
				if (code & 1u) p0 |= bit;
				if (code & 2u) p1 |= bit;
				if (hm.dir) pd |= bit;
				if (c == '!') pb |= bit;
			}
		}
	}
	uint32_t *row = rows + (first_row + k) * genome_words;
	uint32_t *at = row + row_word(0, w); // the five planes of a block are ROW_BLK words apart
	at[PL_V * ROW_BLK] = pv;
	at[PL_C0 * ROW_BLK] = p0;
	at[PL_C1 * ROW_BLK] = p1;
	at[PL_D * ROW_BLK] = pd;
	at[PL_B * ROW_BLK] = pb;
	// flag words behind the planes: does this genome use the D / B planes at all?  The second
	// word marks the row as written (the padding rows of a sharded store never are).
	const uint32_t f = (pd ? ROW_FLAG_D : 0u) | (pb ? ROW_FLAG_B : 0u);
	if (f) {
		uint32_t *flag = row + ROW_PLANES * W;
		if ((*(volatile uint32_t *)flag & f) != f) atomicOr(flag, f);
	}
	if (w == 0) row[ROW_PLANES * W + ROW_WORD_REAL] = 1u;
}

// vall[w] = AND over all genomes of V (complete deletion, process.cxx:725-776 in row form).
// Rows that were never written — the padding slots of a sharded store — are not genomes and
// do not take part (the reference intersects the real sequences only).
__global__ void k_and_valid(const uint32_t *__restrict__ rows, int64_t genome_words, int64_t W, int64_t N,
                            uint32_t *__restrict__ vall)
{
	const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= W) return;
	uint32_t v = 0xffffffffu;
	for (int64_t g = 0; g < N; g++) {
		const uint32_t *row = rows + g * genome_words;
		if (row[ROW_PLANES * W + ROW_WORD_REAL]) v &= row[row_word(PL_V, w)];
	}
	vall[w] = v;
}

constexpr int PT = 4; // genomes per side of a warp's sub-tile
// CT = genomes per side of a block's tile: 16, or 8 when there are so few genomes that 16 x 16
// tiles would be mostly padding.  A tile has (CT / PT)^2 sub-tiles of 4 x 4 pairs, one per warp
// (two with CMP_TILE_WARPS = 8).
// (geometry switchable at compile time for A/B runs: warps per 16 x 16 tile, resident blocks)
// Measured on B200, 1000 x 3 Mbp: 8 warps x 2 blocks per SM x 3 stages 31.4 ms; 16 warps x 1
// block x 6 (or 5) stages 25.7 ms — the same 16 warps per SM, but one deep pipeline instead of
// two shallow ones: the warps that run ahead find their tiles there; 16 warps x 2 blocks (64
// registers) and 8 warps x 3 blocks (80 registers) spill and take 43 / 41 ms.
#ifndef CMP_TILE_WARPS
#define CMP_TILE_WARPS 16
#endif
#ifndef CMP_MIN_BLOCKS
#define CMP_MIN_BLOCKS 1
#endif
#ifndef CMP_STAGES_N
#define CMP_STAGES_N 6
#endif
__host__ __device__ constexpr int cmp_threads(int CT)
{
	return CT == 16 ? 32 * CMP_TILE_WARPS : 128;
}
// words per lane and step: the 3-plane path adds up three words per pair with one carry-save
// step before it counts bits; the 5-plane path (reverse strands, separators) goes word by word
constexpr int CMP_WPL_FAST = 3, CMP_WPL_FULL = 1;
__host__ __device__ constexpr size_t cmp_smem_bytes(int CT, bool tma)
{
	// buffers of 2 CT genomes x planes x (32 * words per lane) words; the larger of the two paths
	const size_t fast = (size_t)(tma ? CMP_STAGES_N : 2) * 2 * CT * 3 * 32 * CMP_WPL_FAST;
	const size_t full = tma ? (size_t)3 * 2 * CT * ROW_PLANES * 32 * CMP_WPL_FAST : (size_t)2 * 2 * CT * ROW_PLANES * 32 * CMP_WPL_FULL;
	return (fast > full ? fast : full) * sizeof(uint32_t);
}

__device__ __forceinline__ void cmp_cp_async16(void *smem_dst, const void *gsrc, int src_bytes)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

// acc += x * K on the multiply-add pipe (the logic pipe is what bounds this kernel)
template <uint32_t K> __device__ __forceinline__ uint32_t cmp_mad(uint32_t x, uint32_t acc)
{
	uint32_t r;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "n"(K), "r"(acc));
	return r;
}

// ---- TMA + mbarrier plumbing (Blackwell tile movement: one thread issues a tensor copy per
// tile side and step, the data lands in shared memory and completes a transaction barrier) ----

// box of the 3-d tensor (word of a block, block, genome) at (0, blk, g) -> shared memory,
// completes on bar
__device__ __forceinline__ void tma_load_rows(void *smem_dst, const CUtensorMap *tm, int32_t blk, int32_t g, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
	             :
	             : "r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(0), "r"(blk), "r"(g)
	             : "memory");
}

constexpr int CMP_STAGES_FAST = CMP_STAGES_N; // shared-memory stages of the TMA pipeline, 3-plane path
constexpr int CMP_STAGES_FULL = 3;            // ... 5-plane path

// One (tile pair, chunk) unit.  P = planes to look at (3: V, C0, C1; 5: all), WPL = words per
// lane and step.  Per pair and word: both = Va & Vb, diff = both & (codes differ [or, on the
// same strand, '!' flags differ]).  The two counts of a pair share one register (homologous
// columns in the low half, substitutions in the high half) that is emptied into the matrix
// before a half can overflow.
//
// Bit counting is the scarce resource: POPC issues at a quarter of the rate of logic
// instructions (measured: the one-word-at-a-time version of this kernel sat on the XU pipe).
// With WPL = 3 the three words of a pair go through one full-adder step first,
//     ones = x0 ^ x1 ^ x2,  twos = maj(x0, x1, x2),  count = popc(ones) + 2 popc(twos),
// two POPC instead of three for two more LOP3: logic and XU pipes come out even.
//
// Tile movement, TMA = true (3-plane path): per step ONE thread issues two tensor copies (the I
// and the J genomes: CT runs of 1152 contiguous bytes each, thanks to the interleaved row
// layout; out-of-range genomes and blocks arrive as zeros) into one of CMP_STAGES stages; the warps wait on the stage's "full"
// transaction barrier, compute, and release the stage through its "empty" barrier — no block-
// wide barrier in the loop, no per-thread address arithmetic.  TMA = false is the Ampere-style
// path (cp.async by all threads, double buffered, one __syncthreads per step), kept for A/B
// measurements (option "compare_path").
template <int P, int CT, int WPL, bool TMA>
__device__ __forceinline__ void compare_tile(uint32_t *stage, const uint32_t *__restrict__ rows, const CUtensorMap *tm,
                                             int64_t genome_words, int64_t W, int64_t N, int64_t gi0, int64_t gj0,
                                             int64_t w_begin, int64_t w_end, const uint32_t *__restrict__ vall,
                                             unsigned long long *__restrict__ subst, unsigned long long *__restrict__ homol)
{
	constexpr int STEP = 32 * WPL;                  // words per genome plane and step
	constexpr int SLOT_WORDS = P * STEP;            // stage[buf][slot][plane][STEP]: slots 0..CT-1 the I genomes, then J
	constexpr int BUF_WORDS = 2 * CT * SLOT_WORDS;
	constexpr int PIECES = 2 * CT * P * (STEP / 4); // 16-byte pieces per step
	constexpr int THREADS = cmp_threads(CT);
	constexpr int SIDE = CT / PT, SUBS = SIDE * SIDE, WARPS = THREADS / 32, SPW = SUBS / WARPS;
	constexpr int FLUSH_STEPS = 65535 / STEP; // a lane adds at most STEP per step to either half
	static_assert(SUBS % WARPS == 0 && SPW >= 1, "sub-tiles per warp");
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	// sub-tiles of padding genomes, or below the diagonal of a diagonal tile, have nothing to count
	bool active[SPW];
	int wi[SPW], wj[SPW];
	uint32_t acc[SPW][PT][PT];
#pragma unroll
	for (int u = 0; u < SPW; u++) {
		const int sub = warp + u * WARPS;
		wi[u] = sub / SIDE;
		wj[u] = sub % SIDE;
		active[u] = gi0 + PT * wi[u] < N && gj0 + PT * wj[u] < N && (gi0 != gj0 || wi[u] <= wj[u]);
#pragma unroll
		for (int a = 0; a < PT; a++)
#pragma unroll
			for (int b = 0; b < PT; b++)
				acc[u][a][b] = 0;
	}
	auto flush = [&] {
#pragma unroll
		for (int u = 0; u < SPW; u++) {
			if (!active[u]) continue;
#pragma unroll
			for (int a = 0; a < PT; a++) {
#pragma unroll
				for (int b = 0; b < PT; b++) {
					uint32_t hv = acc[u][a][b] & 0xffffu, sv = acc[u][a][b] >> 16;
					acc[u][a][b] = 0;
#pragma unroll
					for (int d = 16; d > 0; d >>= 1) {
						sv += __shfl_xor_sync(0xffffffffu, sv, d);
						hv += __shfl_xor_sync(0xffffffffu, hv, d);
					}
					const int64_t i = gi0 + PT * wi[u] + a, j = gj0 + PT * wj[u] + b;
					if (lane == 0 && i < j && j < N && hv) {
						atomicAdd(&subst[i * N + j], (unsigned long long)sv);
						atomicAdd(&homol[i * N + j], (unsigned long long)hv);
					}
				}
			}
		}
	};
	// the 2 CT genomes' words [w0, w0 + STEP) sit in buf: count them into acc
	auto compute = [&](const uint32_t *buf, int64_t w0) {
		const uint32_t *st = buf + lane;
		uint32_t mask[WPL];
#pragma unroll
		for (int k = 0; k < WPL; k++)
			mask[k] = (vall && w0 + 32 * k + lane < w_end) ? vall[w0 + 32 * k + lane] : 0xffffffffu;
#pragma unroll
		for (int u = 0; u < SPW; u++) {
			if (!active[u]) continue;
			// the I side stays in registers, the J side is read genome by genome
			uint32_t av[PT][WPL], a0[PT][WPL], a1[PT][WPL], ad[PT][WPL], ab[PT][WPL];
#pragma unroll
			for (int a = 0; a < PT; a++) {
				const uint32_t *r = st + (PT * wi[u] + a) * SLOT_WORDS;
#pragma unroll
				for (int k = 0; k < WPL; k++) {
					av[a][k] = r[PL_V * STEP + 32 * k] & mask[k];
					a0[a][k] = r[PL_C0 * STEP + 32 * k];
					a1[a][k] = r[PL_C1 * STEP + 32 * k];
					if (P == 5) {
						ad[a][k] = r[PL_D * STEP + 32 * k];
						ab[a][k] = r[PL_B * STEP + 32 * k];
					}
				}
			}
#pragma unroll
			for (int b = 0; b < PT; b++) {
				const uint32_t *r = st + (CT + PT * wj[u] + b) * SLOT_WORDS;
				uint32_t bv[WPL], b0[WPL], b1[WPL], bd[WPL], bb[WPL];
#pragma unroll
				for (int k = 0; k < WPL; k++) {
					bv[k] = r[PL_V * STEP + 32 * k];
					b0[k] = r[PL_C0 * STEP + 32 * k];
					b1[k] = r[PL_C1 * STEP + 32 * k];
					if (P == 5) {
						bd[k] = r[PL_D * STEP + 32 * k];
						bb[k] = r[PL_B * STEP + 32 * k];
					}
				}
#pragma unroll
				for (int a = 0; a < PT; a++) {
					uint32_t both[WPL], diff[WPL];
#pragma unroll
					for (int k = 0; k < WPL; k++) {
						both[k] = av[a][k] & bv[k];
						uint32_t d = (a0[a][k] ^ b0[k]) | (a1[a][k] ^ b1[k]);
						if (P == 5) d |= ~(ad[a][k] ^ bd[k]) & (ab[a][k] ^ bb[k]);
						diff[k] = d & both[k];
					}
					uint32_t c = acc[u][a][b];
					if (WPL == 3) {
						const uint32_t h1 = both[0] ^ both[1] ^ both[2];
						const uint32_t h2 = (both[0] & both[1]) | (both[2] & (both[0] | both[1]));
						const uint32_t s1 = diff[0] ^ diff[1] ^ diff[2];
						const uint32_t s2 = (diff[0] & diff[1]) | (diff[2] & (diff[0] | diff[1]));
						c = cmp_mad<1u>(__popc(h1), c);
						c = cmp_mad<2u>(__popc(h2), c);
						c = cmp_mad<65536u>(__popc(s1), c);
						c = cmp_mad<131072u>(__popc(s2), c);
					} else {
#pragma unroll
						for (int k = 0; k < WPL; k++) {
							c = cmp_mad<1u>(__popc(both[k]), c);
							c = cmp_mad<65536u>(__popc(diff[k]), c);
						}
					}
					acc[u][a][b] = c;
				}
			}
		}
	};

	if constexpr (TMA) {
		// 36 KB per stage with three planes, 60 KB with five: six or three stages fit one SM
		constexpr int CMP_STAGES = P == 3 ? CMP_STAGES_FAST : CMP_STAGES_FULL;
		__shared__ __align__(8) uint64_t full_bar[CMP_STAGES], empty_bar[CMP_STAGES];
		constexpr uint32_t STAGE_BYTES = BUF_WORDS * sizeof(uint32_t);
		const int nsteps = (int)((w_end - w_begin + STEP - 1) / STEP);
		if (threadIdx.x == 0) {
#pragma unroll
			for (int sidx = 0; sidx < CMP_STAGES; sidx++) {
				mbar_init(&full_bar[sidx], 1);
				mbar_init(&empty_bar[sidx], WARPS);
			}
			mbar_fence_init();
		}
		__syncthreads();
		static_assert(STEP == ROW_BLK, "the tensor map's box is one block of the first P planes");
		auto issue = [&](int n) { // thread 0: step n into stage n % CMP_STAGES
			const int sidx = n % CMP_STAGES;
			uint32_t *dst = stage + sidx * BUF_WORDS;
			const int32_t blk = (int32_t)((w_begin + (int64_t)n * STEP) / ROW_BLK); // chunks start on block borders
			mbar_expect_tx(&full_bar[sidx], STAGE_BYTES);
			tma_load_rows(dst, tm, blk, (int32_t)gi0, &full_bar[sidx]);
			tma_load_rows(dst + CT * SLOT_WORDS, tm, blk, (int32_t)gj0, &full_bar[sidx]);
		};
		// Thread 0 refills the stages: at the top of step it, once every warp has let go of the stage
		// step it - 1 used, step it + STAGES - 1 goes there.  (Tried: refilling without waiting —
		// whatever stage mbarrier.test_wait finds free at the top of a step, blocking only for the
		// step the warp itself needs.  Warp 0 then no longer trails the others, but the copies are
		// issued up to a step later and the latency of L2 under this load, 4.4 TB/s, is several
		// steps: 11.8 ms instead of 7.0 ms at 1000 x 0.9 Mbp, waits on the full barrier 37 % of the
		// stall samples.  Depth of the pipeline is what counts here.)
		if (threadIdx.x == 0)
			for (int n = 0; n < CMP_STAGES - 1 && n < nsteps; n++)
				issue(n);
		int steps = 0;
		for (int it = 0; it < nsteps; it++) {
			const int sidx = it % CMP_STAGES;
			if (threadIdx.x == 0) {
				const int nxt = it + CMP_STAGES - 1;
				if (nxt < nsteps) {
					if (it >= 1) mbar_wait(&empty_bar[nxt % CMP_STAGES], (uint32_t)(((it - 1) / CMP_STAGES) & 1));
					issue(nxt);
				}
			}
			__syncwarp();
			mbar_wait(&full_bar[sidx], (uint32_t)((it / CMP_STAGES) & 1));
			compute(stage + sidx * BUF_WORDS, w_begin + (int64_t)it * STEP);
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty_bar[sidx]);
			if (++steps == FLUSH_STEPS) {
				flush();
				steps = 0;
			}
		}
		flush();
	} else {
		auto prefetch = [&](int64_t w0, int buf) {
			for (int piece = threadIdx.x; piece < PIECES; piece += THREADS) {
				const int slot = piece / (P * (STEP / 4)), rem = piece % (P * (STEP / 4)), plane = rem / (STEP / 4),
				          part = rem % (STEP / 4);
				const int64_t g = slot < CT ? gi0 + slot : gj0 + (slot - CT);
				const int64_t w = w0 + 4 * part;
				const bool ok = g < N && w < w_end; // w_end and W are multiples of 4
				const uint32_t *src = rows + (ok ? g * genome_words + row_word(plane, w) : 0); // a step never straddles a block
				cmp_cp_async16(stage + buf * BUF_WORDS + slot * SLOT_WORDS + plane * STEP + 4 * part, src, ok ? 16 : 0);
			}
			asm volatile("cp.async.commit_group;\n" ::: "memory");
		};
		prefetch(w_begin, 0);
		int buf = 0, steps = 0;
		for (int64_t w0 = w_begin; w0 < w_end; w0 += STEP, buf ^= 1) {
			asm volatile("cp.async.wait_group 0;\n" ::: "memory");
			__syncthreads(); // this step's words are in; everybody is done with the other buffer
			if (w0 + STEP < w_end) prefetch(w0 + STEP, buf ^ 1);
			compute(stage + buf * BUF_WORDS, w0);
			if (++steps == FLUSH_STEPS) {
				flush();
				steps = 0;
			}
		}
		flush();
	}
}

// Work units are (tile pair, chunk of columns).  Tile pairs (ti <= tj) are numbered column by
// column, tp = tj (tj + 1) / 2 + ti, so that "all pairs whose later tile is in [tj0, tj1)" —
// what becomes computable when another batch of genomes has been mapped — is one range of
// tp.  Unit u of this launch is tile pair tp_begin + (u * tile_world + tile_rank) / chunks.
// tm_fast / tm_full: tensor maps of the row store whose boxes are one block of the first three
// / of all five planes per genome (TMA = true).  Tiles in which some genome uses the D / B
// planes (reverse strands, separators) take the 5-plane path: same structure, 1920 instead of
// 1152 bytes per genome and step, three stages instead of six.
template <int CT, bool TMA>
__global__ void __launch_bounds__(cmp_threads(CT), CT == 16 ? CMP_MIN_BLOCKS : 4)
k_compare_tiles(const __grid_constant__ CUtensorMap tm_fast, const __grid_constant__ CUtensorMap tm_full,
                const uint32_t *__restrict__ rows, int64_t genome_words, int64_t W, int64_t N, int32_t tile_begin,
                int32_t tile_end, int64_t units, int64_t unit0, int32_t chunks, int64_t chunk_words,
                const uint32_t *__restrict__ vall, unsigned long long *__restrict__ subst,
                unsigned long long *__restrict__ homol)
{
	extern __shared__ __align__(128) uint32_t cmp_stage[];
	__shared__ uint32_t tile_flags;
	const int64_t unit = unit0 + blockIdx.x; // this rank's units are one contiguous range
	if (unit >= units) return;
	const int32_t chunk = (int32_t)(unit % chunks);
	int32_t ti, tj;
	cmp_unrank_pair(unit / chunks, tile_begin, tile_end, ti, tj);
	const int64_t gi0 = (int64_t)ti * CT, gj0 = (int64_t)tj * CT;
	const int64_t w_begin = (int64_t)chunk * chunk_words;
	const int64_t w_end = w_begin + chunk_words < W ? w_begin + chunk_words : W;

	if (threadIdx.x == 0) tile_flags = 0;
	__syncthreads();
	if (threadIdx.x < 2 * CT) {
		const int64_t g = threadIdx.x < CT ? gi0 + threadIdx.x : gj0 + (threadIdx.x - CT);
		if (g < N) {
			const uint32_t f = rows[g * genome_words + ROW_PLANES * W + ROW_WORD_FLAGS];
			if (f) atomicOr(&tile_flags, f);
		}
	}
	__syncthreads();
	if (tile_flags)
		compare_tile<5, CT, TMA ? CMP_WPL_FAST : CMP_WPL_FULL, TMA>(cmp_stage, rows, &tm_full, genome_words, W, N, gi0, gj0,
		                                                            w_begin, w_end, vall, subst, homol);
	else
		compare_tile<3, CT, CMP_WPL_FAST, TMA>(cmp_stage, rows, &tm_fast, genome_words, W, N, gi0, gj0, w_begin, w_end, vall,
		                                       subst, homol);
}

// seg[w] = core columns where some genome differs from genome 0 (process.cxx:484-490:
// get_segsites of queries[0] against every query, OR-ed), core = columns every genome covers
__global__ void k_seg_sites(const uint32_t *__restrict__ rows, int64_t genome_words, int64_t W, int64_t N,
                            const uint32_t *__restrict__ core, uint32_t *__restrict__ seg)
{
	const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= W) return;
	const int64_t at = row_word(0, w);
	const uint32_t *r0 = rows + at;
	const uint32_t a0 = r0[PL_C0 * ROW_BLK], a1 = r0[PL_C1 * ROW_BLK], ad = r0[PL_D * ROW_BLK], ab = r0[PL_B * ROW_BLK];
	uint32_t any = 0;
	for (int64_t g = 1; g < N; g++) {
		const uint32_t *r = rows + g * genome_words + at;
		any |= (a0 ^ r[PL_C0 * ROW_BLK]) | (a1 ^ r[PL_C1 * ROW_BLK]) | (~(ad ^ r[PL_D * ROW_BLK]) & (ab ^ r[PL_B * ROW_BLK]));
	}
	seg[w] = any & core[w];
}

// border bit at the first reference column of every homology of `count` genomes
__global__ void k_hom_borders(const Hom *__restrict__ homs, const int64_t *__restrict__ begin,
                              const int64_t *__restrict__ hcount, int32_t count, uint32_t *__restrict__ border)
{
	const int32_t k = blockIdx.y;
	if (k >= count) return;
	const Hom *H = homs + begin[k];
	const int64_t h = hcount[k];
	for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < h; x += (int64_t)gridDim.x * blockDim.x) {
		const uint32_t c = (uint32_t)H[x].iproj;
		atomicOr(&border[c >> 5], 1u << (c & 31));
	}
}

__global__ void k_symmetrize(unsigned long long *__restrict__ subst, unsigned long long *__restrict__ homol, int64_t N)
{
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= N * N) return;
	const int64_t i = k / N, j = k % N;
	if (i > j) {
		subst[k] = subst[j * N + i];
		homol[k] = homol[j * N + i];
	}
}

__global__ void k_estimate(const unsigned long long *__restrict__ subst, const unsigned long long *__restrict__ homol,
                           int64_t N, int kind, double *__restrict__ out)
{
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= N * N) return;
	const int64_t i = k / N, j = k % N;
	if (i == j) {
		out[k] = 0.0; // io.cxx:157
		return;
	}
	const unsigned long long h = homol[k];
	double d;
	if (h == 0) {
		d = nan("");
	} else {
		const double raw = (double)subst[k] / (double)h; // evo_model.cxx:100-107
		if (kind == 0)
			d = raw;
		else if (kind == 2)
			d = (1.0 - raw) * 100; // :112-119
		else {
			d = -0.75 * log(1.0 - (4.0 / 3.0) * raw); // :124-131
			d = d <= 0.0 ? 0.0 : d;
		}
	}
	out[k] = d;
}

} // namespace

void rows_alloc(RowStore &rs, int64_t genomes, int32_t n, cudaStream_t s)
{
	rs.n = n;
	rs.W = (((int64_t)n + 31) / 32 + ROW_BLK - 1) / ROW_BLK * ROW_BLK;
	rs.genomes = genomes;
	rs.data.alloc((size_t)(genomes * rs.genome_words()), s);
	rs.maps.base = nullptr; // the tensor maps describe the old geometry
	rs.data.zero(); // rows never written (padding genomes of a sharded run) are all-invalid
}

void rows_clear_flags(RowStore &rs, cudaStream_t s)
{
	if (!rs.genomes) return;
	CUDA_CHECK(cudaMemset2DAsync(rs.row(0) + ROW_PLANES * rs.W, (size_t)rs.genome_words() * sizeof(uint32_t), 0,
	                             ROW_FLAG_WORDS * sizeof(uint32_t), (size_t)rs.genomes, s));
}

void rows_build(RowStore &rs, int64_t first_row, const uint8_t *d_Q, const QueryInfo *d_qi, int32_t count,
                const Hom *d_homs, const int64_t *d_begin, const int64_t *d_count, cudaStream_t s)
{
	if (count <= 0) return;
	if (first_row < 0 || first_row + count > rs.genomes) throw std::invalid_argument("row store too small");
	// the flag words behind the rows about to be written
	CUDA_CHECK(cudaMemset2DAsync(rs.row(first_row) + ROW_PLANES * rs.W, (size_t)rs.genome_words() * sizeof(uint32_t), 0,
	                             ROW_FLAG_WORDS * sizeof(uint32_t), (size_t)count, s));
	for (int32_t k0 = 0; k0 < count; k0 += 32768) { // gridDim.y limit
		const int32_t c = count - k0 < 32768 ? count - k0 : 32768;
		dim3 grid(div_up(rs.W, 128), c);
		k_build_rows<<<grid, 128, 0, s>>>(rs.data.get(), rs.genome_words(), rs.W, rs.n, first_row + k0, d_Q, d_qi + k0, c,
		                                  d_homs, d_begin + k0, d_count + k0);
		KERNEL_CHECK();
	}
}

namespace
{
// Rows [first, first + count) of this GPU's store written into the peers' stores with plain
// stores over NVLink (peer memory mapped into this process): the exchange of a sharded run as
// a kernel, for the batch nothing else can hide — 16 bytes per thread and peer, every SM
// feeding the links.  A genome that uses no D / B plane sends only its first three planes
// (1152 of every 1920 bytes), unless D / B planes have ever been sent from this store
// (*db_sent): then the peers may hold stale ones and zeros have to go over as well.
__global__ void __launch_bounds__(256)
k_push_rows(const uint32_t *__restrict__ rows, RowPeers peers, int64_t genome_words, int64_t W, int64_t first,
            int *__restrict__ db_sent)
{
	const int64_t g = first + blockIdx.y;
	const uint32_t *row = rows + g * genome_words;
	const uint32_t flags = row[ROW_PLANES * W + ROW_WORD_FLAGS];
	const bool all_planes = flags != 0 || *db_sent != 0;
	if (flags != 0 && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(db_sent, 1);
	constexpr int UNITS_PER_BLK = ROW_PLANES * ROW_BLK / 4, UNITS_3 = 3 * ROW_BLK / 4; // 16-byte units
	const int64_t units = (W / ROW_BLK) * UNITS_PER_BLK + ROW_FLAG_WORDS / 4;          // + the flag words
	const uint4 *src = reinterpret_cast<const uint4 *>(row);
	for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
		if (!all_planes && u < units - 1 && (u % UNITS_PER_BLK) >= UNITS_3) continue;
		const uint4 v = src[u];
		const int64_t at = g * (genome_words / 4) + u;
#pragma unroll 4
		for (int p = 0; p < peers.n; p++)
			if (peers.ptr[p]) reinterpret_cast<uint4 *>(peers.ptr[p])[at] = v;
	}
}

// Tensor map of the row store: a 3-d tensor of 64-bit elements (element of a block's first
// `planes` planes, block, genome) whose box is one block of one tile side: 144 elements (= 3
// planes x ROW_BLK words, 1152 contiguous bytes; 240 elements for all five planes) x 1 block x
// CT genomes.
// (64-bit elements because a box dimension may not exceed 256 elements.)  Blocks past the row
// and genomes past the store read as zeros.  cuTensorMapEncodeTiled is a driver entry point; it
// is fetched through the runtime so that the library does not link against libcuda.
void rows_tensor_map(const RowStore &rs, int planes, int CT, CUtensorMap *out)
{
	using Encode = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
	                            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
	                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static Encode encode = nullptr;
	if (!encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult q;
		CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
		if (q != cudaDriverEntryPointSuccess || !fn) throw CudaError("cuTensorMapEncodeTiled is not available in this driver");
		encode = (Encode)fn;
	}
	const cuuint32_t RUN = (cuuint32_t)planes * ROW_BLK / 2; // 64-bit elements of the first `planes` planes of a block
	const cuuint64_t dims[3] = {RUN, (cuuint64_t)(rs.W / ROW_BLK), (cuuint64_t)rs.genomes};
	const cuuint64_t strides[2] = {(cuuint64_t)ROW_PLANES * ROW_BLK * sizeof(uint32_t),
	                               (cuuint64_t)rs.genome_words() * sizeof(uint32_t)};
	const cuuint32_t box[3] = {RUN, 1, (cuuint32_t)CT};
	const cuuint32_t elem[3] = {1, 1, 1};
	const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, rs.data.get(), dims, strides, box, elem,
	                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
	                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed for the row store (code " + std::to_string((int)r) + ")");
}
} // namespace

void rows_push_kernel(const RowStore &rs, const RowPeers &peers, int64_t first, int32_t count, int *d_db_sent,
                      cudaStream_t s)
{
	if (count <= 0 || peers.n <= 0) return;
	const int64_t units = (rs.W / ROW_BLK) * (ROW_PLANES * ROW_BLK / 4) + 1;
	int bx = div_up(units, 256 * 4); // four 16-byte units per thread
	if (bx > 2048) bx = 2048;
	for (int32_t k0 = 0; k0 < count; k0 += 32768) {
		const int32_t c = count - k0 < 32768 ? count - k0 : 32768;
		k_push_rows<<<dim3(bx, c), 256, 0, s>>>(rs.data.get(), peers, rs.genome_words(), rs.W, first + k0, d_db_sent);
		KERNEL_CHECK();
	}
}

int compare_tile_side(int64_t N)
{
	return N <= 24 ? 8 : 16; // sharding.py mirrors this choice
}

void compare_all_device(const RowStore &rs, int64_t N, bool complete_deletion, int tile_rank, int tile_world,
                        unsigned long long *d_subst, unsigned long long *d_homologs, cudaStream_t s, int64_t tile_begin,
                        int64_t tile_end, bool first, bool last)
{
	if (N > rs.genomes) throw std::invalid_argument("row store holds fewer genomes than N");
	const int CT = compare_tile_side(N);
	const int64_t tiles_side = (N + CT - 1) / CT;
	if (tile_end < 0 || tile_end > tiles_side) tile_end = tiles_side;
	if (tile_begin < 0 || tile_begin > tile_end) throw std::invalid_argument("bad tile range");
	if (complete_deletion && !(first && last))
		throw std::invalid_argument("complete deletion needs all rows: no incremental comparison");
	if (first) {
		CUDA_CHECK(cudaMemsetAsync(d_subst, 0, (size_t)(N * N) * sizeof(unsigned long long), s));
		CUDA_CHECK(cudaMemsetAsync(d_homologs, 0, (size_t)(N * N) * sizeof(unsigned long long), s));
	}
	if (N < 2) return;
	DevBuf<uint32_t> vall;
	if (complete_deletion) {
		vall.alloc((size_t)rs.W, s);
		k_and_valid<<<div_up(rs.W, 256), 256, 0, s>>>(rs.data.get(), rs.genome_words(), rs.W, N, vall.get());
		KERNEL_CHECK();
	}
	// tile pairs (ti <= tj) with tj in [tile_begin, tile_end), numbered column by column
	const int64_t n_tile_pairs = tile_end * (tile_end + 1) / 2 - tile_begin * (tile_begin + 1) / 2;
	if (n_tile_pairs > 0) {
		// enough blocks to fill the machine a few times over, chunks of at least 256 words
		const int64_t want_blocks = (int64_t)NUM_SMS_B200 * 8 * tile_world;
		int64_t chunks = (want_blocks + n_tile_pairs - 1) / n_tile_pairs;
		const int64_t max_chunks = (rs.W + 255) / 256;
		if (chunks > max_chunks) chunks = max_chunks;
		if (chunks < 1) chunks = 1;
		int64_t chunk_words = (rs.W + chunks - 1) / chunks;
		chunk_words = (chunk_words + 95) / 96 * 96; // whole steps of both paths (96 and 32 words): a tensor copy never straddles a chunk end
		chunks = (rs.W + chunk_words - 1) / chunk_words;
		const int64_t units = n_tile_pairs * chunks;
		// every rank takes one contiguous range of the units (its blocks share rows in L2)
		const int64_t per_rank = (units + tile_world - 1) / tile_world;
		const int64_t unit0 = per_rank * tile_rank;
		const int64_t my_units = unit0 >= units ? 0 : (units - unit0 < per_rank ? units - unit0 : per_rank);
		const bool tma = g_tuning.compare_path == 0;
		static PerDeviceOnce once;
		if (once.first()) {
			CUDA_CHECK(cudaFuncSetAttribute(k_compare_tiles<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			                                (int)cmp_smem_bytes(16, true)));
			CUDA_CHECK(cudaFuncSetAttribute(k_compare_tiles<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			                                (int)cmp_smem_bytes(8, true)));
			CUDA_CHECK(cudaFuncSetAttribute(k_compare_tiles<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			                                (int)cmp_smem_bytes(16, false)));
			CUDA_CHECK(cudaFuncSetAttribute(k_compare_tiles<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			                                (int)cmp_smem_bytes(8, false)));
		}
		if (my_units > 0x7fffffffll) throw std::invalid_argument("too many tile pairs for one launch");
		if (my_units > 0) {
			static_assert(sizeof(CUtensorMap) == 128, "RowStore::MapCache holds two tensor maps");
			CUtensorMap tm_fast, tm_full;
			memset(&tm_fast, 0, sizeof tm_fast);
			memset(&tm_full, 0, sizeof tm_full);
			if (tma) {
				if (rs.maps.base != rs.data.get() || rs.maps.ct != CT) { // (a new allocation or another tile side)
					rows_tensor_map(rs, 3, CT, &tm_fast);
					rows_tensor_map(rs, ROW_PLANES, CT, &tm_full);
					memcpy(rs.maps.bytes[0], &tm_fast, 128);
					memcpy(rs.maps.bytes[1], &tm_full, 128);
					rs.maps.base = rs.data.get();
					rs.maps.ct = CT;
				} else {
					memcpy(&tm_fast, rs.maps.bytes[0], 128);
					memcpy(&tm_full, rs.maps.bytes[1], 128);
				}
			}
			const size_t smem = cmp_smem_bytes(CT, tma);
#define PHY_LAUNCH_COMPARE(CTV, TMAV)                                                                                    \
	k_compare_tiles<CTV, TMAV><<<(unsigned)my_units, cmp_threads(CTV), smem, s>>>(                                       \
		tm_fast, tm_full, rs.data.get(), rs.genome_words(), rs.W, N, (int32_t)tile_begin, (int32_t)tile_end, units,     \
		unit0, (int32_t)chunks, chunk_words, vall.get(), d_subst, d_homologs)
			if (CT == 16 && tma)
				PHY_LAUNCH_COMPARE(16, true);
			else if (CT == 16)
				PHY_LAUNCH_COMPARE(16, false);
			else if (tma)
				PHY_LAUNCH_COMPARE(8, true);
			else
				PHY_LAUNCH_COMPARE(8, false);
#undef PHY_LAUNCH_COMPARE
			KERNEL_CHECK();
		}
	}
	if (last) {
		k_symmetrize<<<div_up(N * N, 256), 256, 0, s>>>(d_subst, d_homologs, N);
		KERNEL_CHECK();
	}
}

void core_sites_device(const RowStore &rs, int64_t N, uint32_t *d_core, uint32_t *d_seg, cudaStream_t s)
{
	if (N > rs.genomes) throw std::invalid_argument("row store holds fewer genomes than N");
	k_and_valid<<<div_up(rs.W, 256), 256, 0, s>>>(rs.data.get(), rs.genome_words(), rs.W, N, d_core);
	KERNEL_CHECK();
	k_seg_sites<<<div_up(rs.W, 256), 256, 0, s>>>(rs.data.get(), rs.genome_words(), rs.W, N, d_core, d_seg);
	KERNEL_CHECK();
}

void hom_borders_device(const Hom *d_homs, const int64_t *d_begin, const int64_t *d_count, int32_t count,
                        uint32_t *d_border, cudaStream_t s)
{
	for (int32_t k0 = 0; k0 < count; k0 += 32768) {
		const int32_t c = count - k0 < 32768 ? count - k0 : 32768;
		k_hom_borders<<<dim3(8, c), 128, 0, s>>>(d_homs, d_begin + k0, d_count + k0, c, d_border);
		KERNEL_CHECK();
	}
}

void estimate_device(const unsigned long long *d_subst, const unsigned long long *d_homologs, int64_t N, int kind,
                     double *d_out, cudaStream_t s)
{
	if (N <= 0) return;
	k_estimate<<<div_up(N * N, 256), 256, 0, s>>>(d_subst, d_homologs, N, kind, d_out);
	KERNEL_CHECK();
}

} // namespace phy
