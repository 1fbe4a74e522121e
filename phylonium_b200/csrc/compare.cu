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
// k_compare_pairs: one warp per (4 x 4 genome tile, column chunk).  The 32 lanes take 32
// consecutive words, so every row-plane load is one coalesced 128-byte request; the 16
// pair counters live in registers and are reduced with shuffles once per chunk.
#include "compare_device.h"
#include "primitives.cuh"

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
	row[PL_V * W + w] = pv;
	row[PL_C0 * W + w] = p0;
	row[PL_C1 * W + w] = p1;
	row[PL_D * W + w] = pd;
	row[PL_B * W + w] = pb;
}

// vall[w] = AND over all genomes of V (complete deletion, process.cxx:725-776 in row form)
__global__ void k_and_valid(const uint32_t *__restrict__ rows, int64_t genome_words, int64_t W, int64_t N,
                            uint32_t *__restrict__ vall)
{
	const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= W) return;
	uint32_t v = 0xffffffffu;
	for (int64_t g = 0; g < N; g++)
		v &= rows[g * genome_words + PL_V * W + w];
	vall[w] = v;
}

constexpr int PT = 4; // genomes per tile side

__global__ void __launch_bounds__(128)
k_compare_pairs(const uint32_t *__restrict__ rows, int64_t genome_words, int64_t W, int64_t N, int32_t tiles_side,
                int64_t n_tile_pairs, int32_t chunks, int64_t chunk_words, int tile_rank, int tile_world,
                const uint32_t *__restrict__ vall, unsigned long long *__restrict__ subst,
                unsigned long long *__restrict__ homol)
{
	const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	const int64_t my_pairs = (n_tile_pairs - tile_rank + tile_world - 1) / tile_world;
	if (warp >= my_pairs * chunks) return;
	const int64_t tp = (warp / chunks) * tile_world + tile_rank; // tile pair index
	const int32_t chunk = (int32_t)(warp % chunks);
	// unrank tp -> (ti <= tj): row ti holds tiles_side - ti pairs
	int32_t ti = 0;
	int64_t rem = tp;
	while (rem >= tiles_side - ti) {
		rem -= tiles_side - ti;
		ti++;
	}
	const int32_t tj = ti + (int32_t)rem;
	const int64_t gi0 = (int64_t)ti * PT, gj0 = (int64_t)tj * PT;

	uint32_t cs[PT][PT], ch[PT][PT];
#pragma unroll
	for (int a = 0; a < PT; a++)
#pragma unroll
		for (int b = 0; b < PT; b++)
			cs[a][b] = ch[a][b] = 0;

	const int64_t w_begin = (int64_t)chunk * chunk_words;
	const int64_t w_end = w_begin + chunk_words < W ? w_begin + chunk_words : W;
	for (int64_t w = w_begin + lane; w < w_end; w += 32) {
		uint32_t av[PT], a0[PT], a1[PT], ad[PT], ab[PT];
		uint32_t bv[PT], b0[PT], b1[PT], bd[PT], bb[PT];
		const uint32_t mask = vall ? vall[w] : 0xffffffffu;
#pragma unroll
		for (int a = 0; a < PT; a++) {
			const bool ok = gi0 + a < N;
			const uint32_t *r = rows + (ok ? gi0 + a : 0) * genome_words + w;
			av[a] = ok ? (r[PL_V * W] & mask) : 0u;
			a0[a] = r[PL_C0 * W];
			a1[a] = r[PL_C1 * W];
			ad[a] = r[PL_D * W];
			ab[a] = r[PL_B * W];
		}
#pragma unroll
		for (int b = 0; b < PT; b++) {
			const bool ok = gj0 + b < N;
			const uint32_t *r = rows + (ok ? gj0 + b : 0) * genome_words + w;
			bv[b] = ok ? r[PL_V * W] : 0u;
			b0[b] = r[PL_C0 * W];
			b1[b] = r[PL_C1 * W];
			bd[b] = r[PL_D * W];
			bb[b] = r[PL_B * W];
		}
#pragma unroll
		for (int a = 0; a < PT; a++) {
#pragma unroll
			for (int b = 0; b < PT; b++) {
				const uint32_t both = av[a] & bv[b];
				const uint32_t diff = (a0[a] ^ b0[b]) | (a1[a] ^ b1[b]) | (~(ad[a] ^ bd[b]) & (ab[a] ^ bb[b]));
				ch[a][b] += __popc(both);
				cs[a][b] += __popc(both & diff);
			}
		}
	}
#pragma unroll
	for (int a = 0; a < PT; a++) {
#pragma unroll
		for (int b = 0; b < PT; b++) {
			uint32_t s = cs[a][b], h = ch[a][b];
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) {
				s += __shfl_xor_sync(0xffffffffu, s, d);
				h += __shfl_xor_sync(0xffffffffu, h, d);
			}
			const int64_t i = gi0 + a, j = gj0 + b;
			if (lane == 0 && i < j && j < N && h) {
				atomicAdd(&subst[i * N + j], (unsigned long long)s);
				atomicAdd(&homol[i * N + j], (unsigned long long)h);
			}
		}
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
	rs.W = (((int64_t)n + 31) / 32 + 3) / 4 * 4;
	rs.genomes = genomes;
	rs.data.alloc((size_t)(genomes * ROW_PLANES * rs.W), s);
	rs.data.zero(); // rows never written (padding genomes of a sharded run) are all-invalid
}

void rows_build(RowStore &rs, int64_t first_row, const uint8_t *d_Q, const QueryInfo *d_qi, int32_t count,
                const Hom *d_homs, const int64_t *d_begin, const int64_t *d_count, cudaStream_t s)
{
	if (count <= 0) return;
	if (first_row < 0 || first_row + count > rs.genomes) throw std::invalid_argument("row store too small");
	for (int32_t k0 = 0; k0 < count; k0 += 32768) { // gridDim.y limit
		const int32_t c = count - k0 < 32768 ? count - k0 : 32768;
		dim3 grid(div_up(rs.W, 128), c);
		k_build_rows<<<grid, 128, 0, s>>>(rs.data.get(), rs.genome_words(), rs.W, rs.n, first_row + k0, d_Q, d_qi + k0, c,
		                                  d_homs, d_begin + k0, d_count + k0);
		KERNEL_CHECK();
	}
}

void compare_all_device(const RowStore &rs, int64_t N, bool complete_deletion, int tile_rank, int tile_world,
                        unsigned long long *d_subst, unsigned long long *d_homologs, cudaStream_t s)
{
	if (N > rs.genomes) throw std::invalid_argument("row store holds fewer genomes than N");
	CUDA_CHECK(cudaMemsetAsync(d_subst, 0, (size_t)(N * N) * sizeof(unsigned long long), s));
	CUDA_CHECK(cudaMemsetAsync(d_homologs, 0, (size_t)(N * N) * sizeof(unsigned long long), s));
	if (N < 2) return;
	DevBuf<uint32_t> vall;
	if (complete_deletion) {
		vall.alloc((size_t)rs.W, s);
		k_and_valid<<<div_up(rs.W, 256), 256, 0, s>>>(rs.data.get(), rs.genome_words(), rs.W, N, vall.get());
		KERNEL_CHECK();
	}
	const int32_t tiles_side = (int32_t)((N + PT - 1) / PT);
	const int64_t n_tile_pairs = (int64_t)tiles_side * (tiles_side + 1) / 2;
	const int64_t my_pairs = (n_tile_pairs - tile_rank + tile_world - 1) / tile_world;
	if (my_pairs > 0) {
		// enough warps to fill the machine a few times over, chunks of at least 256 words
		const int64_t want_warps = (int64_t)NUM_SMS_B200 * 64 * 4;
		int64_t chunks = (want_warps + my_pairs - 1) / my_pairs;
		const int64_t max_chunks = (rs.W + 255) / 256;
		if (chunks > max_chunks) chunks = max_chunks;
		if (chunks < 1) chunks = 1;
		int64_t chunk_words = (rs.W + chunks - 1) / chunks;
		chunk_words = (chunk_words + 31) / 32 * 32;
		chunks = (rs.W + chunk_words - 1) / chunk_words;
		const int64_t warps = my_pairs * chunks;
		k_compare_pairs<<<div_up(warps * 32, 128), 128, 0, s>>>(rs.data.get(), rs.genome_words(), rs.W, N, tiles_side,
		                                                        n_tile_pairs, (int32_t)chunks, chunk_words, tile_rank,
		                                                        tile_world, vall.get(), d_subst, d_homologs);
		KERNEL_CHECK();
	}
	k_symmetrize<<<div_up(N * N, 256), 256, 0, s>>>(d_subst, d_homologs, N);
	KERNEL_CHECK();
}

void estimate_device(const unsigned long long *d_subst, const unsigned long long *d_homologs, int64_t N, int kind,
                     double *d_out, cudaStream_t s)
{
	if (N <= 0) return;
	k_estimate<<<div_up(N * N, 256), 256, 0, s>>>(d_subst, d_homologs, N, kind, d_out);
	KERNEL_CHECK();
}

} // namespace phy
