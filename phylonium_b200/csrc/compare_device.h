// All-pairs comparison on the device (compare.cu): replaces hot loop B of
// /root/reference/src/process.cxx:524-549 (compare() over all pairs, :566-658, with
// evo_model::account/account_rev -> seqcmp/revseqcmp) and evo_model's estimators.
#pragma once
#include "common.cuh"
#include "walk.h"

namespace phy
{

// Reference-coordinate rows, bit-sliced.  For every genome five planes of W 32-bit words
// (bit b of word w <-> reference column 32 w + b):
//   V   column covered by one of the genome's (filtered, disjoint) homologies
//   C0  low  bit of the nucleotide code  (byte & 6) >> 1: A=0 C=1 T=2 G=3, '!'=0;
//   C1  high bit;  a reverse-strand homology stores code ^ 2, i.e. the complement
//   D   the covering homology is on the reverse strand
//   B   the query byte is the contig separator '!'
// Two genomes differ in a column iff the codes differ, or — same strand only — the '!'
// flags differ (seqcmp compares bytes, revseqcmp only bits 1-2).
// Behind the planes of a genome sit ROW_FLAG_WORDS words; the first says whether any bit of
// the D / B planes is set (genomes without reverse-strand homologies and separators, the
// common case, let the all-pairs kernel skip those planes).  The flags travel with the rows
// in the all-gather of a sharded run.
// In memory the planes of a genome are interleaved in blocks of ROW_BLK words:
//   [block 0: V C0 C1 D B, ROW_BLK words each][block 1: ...] ... [flag words]
// so that what the all-pairs kernel needs of one genome per step — ROW_BLK words of its first
// three planes — is ONE contiguous run of 1152 bytes (a tensor copy then moves 16 such runs per
// tile side; with whole planes back to back it had to fetch 48 runs of 384 bytes and the copy
// engine, not the arithmetic, set the pace).  row_word() gives the place of a word.
constexpr int ROW_PLANES = 5;
constexpr int ROW_BLK = 96;
constexpr int ROW_FLAG_WORDS = 4;
enum : int { PL_V = 0, PL_C0 = 1, PL_C1 = 2, PL_D = 3, PL_B = 4 };
enum : uint32_t { ROW_FLAG_D = 1, ROW_FLAG_B = 2 };
enum : int { ROW_WORD_FLAGS = 0, ROW_WORD_REAL = 1 }; // flag words: D/B use; "this row was written"

// index of word w of plane `plane` inside a genome's row
__host__ __device__ inline int64_t row_word(int plane, int64_t w)
{
	return (w / ROW_BLK) * (ROW_PLANES * ROW_BLK) + plane * ROW_BLK + (w % ROW_BLK);
}

struct RowStore {
	DevBuf<uint32_t> data; // genomes * (ROW_PLANES * W + ROW_FLAG_WORDS)
	int64_t genomes = 0;   // capacity in genomes
	int64_t W = 0;         // words per plane, multiple of ROW_BLK
	int32_t n = 0;         // reference length (columns)
	// tensor maps of the store (compare.cu: rows_tensor_map), encoded once per allocation and tile
	// side: [0] three planes, [1] all five; 128 opaque bytes each (a CUtensorMap)
	struct alignas(64) MapCache {
		unsigned char bytes[2][128];
		const void *base = nullptr;
		int ct = 0;
	};
	mutable MapCache maps;
	uint32_t *row(int64_t g) const { return data.get() + g * genome_words(); }
	int64_t genome_words() const { return ROW_PLANES * W + ROW_FLAG_WORDS; }
};

// the row stores of the other ranks of a sharded run, as mapped into this process (nullptr
// for this rank itself)
struct RowPeers {
	uint32_t *ptr[16];
	int n = 0;
};
// copies rows [first, first + count) into the peers' stores with a kernel (stores over NVLink);
// d_db_sent: device flag "D / B planes have been sent from this store before"
void rows_push_kernel(const RowStore &rs, const RowPeers &peers, int64_t first, int32_t count, int *d_db_sent,
                      cudaStream_t s);

void rows_alloc(RowStore &rs, int64_t genomes, int32_t n, cudaStream_t s);
// marks every slot as "not written" (and without D / B planes)
void rows_clear_flags(RowStore &rs, cudaStream_t s);

// rows of `count` genomes (genome k's sorted, disjoint homologies are
// d_homs[d_begin[k] .. d_begin[k] + d_count[k]), bases in d_Q / d_qi) written to rs.row(first_row + k)
void rows_build(RowStore &rs, int64_t first_row, const uint8_t *d_Q, const QueryInfo *d_qi, int32_t count,
                const Hom *d_homs, const int64_t *d_begin, const int64_t *d_count, cudaStream_t s);

// substitutions / homologs (N*N, row-major, symmetric, zero diagonal).  Work units are (tile
// pair, column chunk); this call computes units tile_rank, tile_rank + tile_world, … and leaves
// the rest of the matrix zero.  Tiles are compare_tile_side(N) genomes wide; only pairs of
// tiles (ti <= tj) with tj in [tile_begin, tile_end) are compared (tile_end < 0: to the last
// one), so a caller can compare as the rows arrive: `first` zeroes the matrix, `last` mirrors
// the upper triangle.  Complete deletion needs all rows at once.
int compare_tile_side(int64_t N);
void compare_all_device(const RowStore &rs, int64_t N, bool complete_deletion, int tile_rank, int tile_world,
                        unsigned long long *d_subst, unsigned long long *d_homologs, cudaStream_t s,
                        int64_t tile_begin = 0, int64_t tile_end = -1, bool first = true, bool last = true);

// Core genome in row form (the reference's -p option, process.cxx:471-513): d_core[w] = columns
// covered by all N genomes, d_seg[w] = core columns where some genome differs from genome 0
// (W words each); hom_borders_device sets the bit of the first column of every homology.
void core_sites_device(const RowStore &rs, int64_t N, uint32_t *d_core, uint32_t *d_seg, cudaStream_t s);
void hom_borders_device(const Hom *d_homs, const int64_t *d_begin, const int64_t *d_count, int32_t count,
                        uint32_t *d_border, cudaStream_t s);

// kind 0 raw, 1 Jukes-Cantor, 2 ANI (evo_model.cxx:100-131); diagonal 0
void estimate_device(const unsigned long long *d_subst, const unsigned long long *d_homologs, int64_t N, int kind,
                     double *d_out, cudaStream_t s);

} // namespace phy
