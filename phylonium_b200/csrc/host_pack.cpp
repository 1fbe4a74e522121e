// Host side of the packed upload: 4 bases per byte.
//
// The sequences reach phylo_process() as byte strings over {A,C,G,T,!}; over PCIe they travel
// as 2-bit codes, a quarter of the bytes (on this pool's boxes the bus, not the GPU, bounds
// process() for 1000 genomes: 3 GB at ~28 GB/s against ~65 ms of kernels).  The code of a base
// is (c >> 1) & 3 — A 0, C 1, T 2, G 3 —, base k of a group of four sits in bits 2k, 2k+1.
// '!' (the contig separator, src/sequence.cxx:171-199) packs as 0 and is listed separately;
// any other byte is an error, so the alphabet check of the queries happens here, for free,
// while the bytes are in registers anyway.
//
// Plain C++ (no CUDA): compiled by g++ with function-level ISA targets and dispatched at run
// time, so the library still loads on a CPU without AVX2/BMI2.
#include "host_pack.h"

#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace phy
{

namespace
{

// 0..3: code; 4: '!'; 255: not in the alphabet
struct CodeTable {
	uint8_t t[256];
	CodeTable()
	{
		memset(t, 255, sizeof t);
		t[(unsigned char)'A'] = 0;
		t[(unsigned char)'C'] = 1;
		t[(unsigned char)'T'] = 2;
		t[(unsigned char)'G'] = 3;
		t[(unsigned char)'!'] = 4;
	}
};
const CodeTable g_codes;

// bases [i0, i1) one by one; i0 is a multiple of 4
inline int pack_scalar(const uint8_t *src, size_t i0, size_t i1, uint8_t *dst, uint32_t *bangs, uint32_t cap,
                       uint32_t &nb)
{
	int bad = 0;
	for (size_t i = i0; i < i1; i += 4) {
		uint32_t byte = 0;
		for (size_t k = 0; k < 4 && i + k < i1; k++) {
			uint8_t c = g_codes.t[src[i + k]];
			if (c == 4) {
				if (nb < cap) bangs[nb] = (uint32_t)(i + k);
				nb++;
				c = 0;
			} else if (c == 255) {
				bad = 1;
				c = 0;
			}
			byte |= (uint32_t)c << (2 * k);
		}
		dst[i >> 2] = (uint8_t)byte;
	}
	return bad;
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"))) int pack_avx2(const uint8_t *src, size_t n, uint8_t *dst, uint32_t *bangs,
                                                     uint32_t cap, uint32_t &nb)
{
	// expected byte for every low nibble: A 0x41 -> 1, C 0x43 -> 3, T 0x54 -> 4, G 0x47 -> 7
	// (a nibble without a letter maps to a byte with ANOTHER low nibble, which no input byte
	// that selected it can equal)
#define PHY_NO(n) (char)(0x80 | (((n) + 1) & 15))
	const __m256i lut = _mm256_setr_epi8(
		PHY_NO(0), 'A', PHY_NO(2), 'C', 'T', PHY_NO(5), PHY_NO(6), 'G', PHY_NO(8), PHY_NO(9), PHY_NO(10), PHY_NO(11),
		PHY_NO(12), PHY_NO(13), PHY_NO(14), PHY_NO(15), PHY_NO(0), 'A', PHY_NO(2), 'C', 'T', PHY_NO(5), PHY_NO(6), 'G',
		PHY_NO(8), PHY_NO(9), PHY_NO(10), PHY_NO(11), PHY_NO(12), PHY_NO(13), PHY_NO(14), PHY_NO(15));
#undef PHY_NO
	const __m256i low = _mm256_set1_epi8(0x0f), bang = _mm256_set1_epi8('!');
	int bad = 0;
	size_t i = 0;
	for (; i + 32 <= n; i += 32) {
		const __m256i v = _mm256_loadu_si256((const __m256i *)(src + i));
		const __m256i expect = _mm256_shuffle_epi8(lut, _mm256_and_si256(v, low));
		const uint32_t ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, expect));
		if (ok != 0xffffffffu) {
			const uint32_t isbang = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, bang));
			if ((ok | isbang) != 0xffffffffu) bad = 1;
			for (uint32_t m = isbang; m; m &= m - 1) {
				if (nb < cap) bangs[nb] = (uint32_t)(i + (size_t)__builtin_ctz(m));
				nb++;
			}
			// '!' is 0x21: bits 1, 2 clear, it packs as 0 like it should; an invalid byte packs as
			// whatever its bits say, the call fails anyway
		}
		uint64_t x[4];
		memcpy(x, src + i, 32);
		uint16_t out[4];
		for (int k = 0; k < 4; k++)
			out[k] = (uint16_t)_pext_u64(x[k], 0x0606060606060606ull);
		memcpy(dst + (i >> 2), out, 8);
	}
	bad |= pack_scalar(src, i, n, dst, bangs, cap, nb);
	return bad;
}
#endif

} // namespace

int pack_2bit(const uint8_t *src, size_t n, uint8_t *dst, uint32_t *bangs, uint32_t cap, uint32_t *nbangs)
{
	uint32_t nb = 0;
	int bad;
#if defined(__x86_64__)
	static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
	if (fast)
		bad = pack_avx2(src, n, dst, bangs, cap, nb);
	else
#endif
		bad = pack_scalar(src, 0, n, dst, bangs, cap, nb);
	*nbangs = nb;
	return bad;
}

} // namespace phy
