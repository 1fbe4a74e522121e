/* phylonium_b200 — C ABI of the B200-native distance pipeline.
 *
 * Drop-in boundary.  The reference, EvolBioInf/phylonium v1.7, has no
 * plugin or FFI layer; the seam this library replaces is
 *
 *     std::vector<evo_model> process(const sequence &, const std::vector<sequence> &)
 *                                               /root/reference/src/process.h:12
 *
 * called from main() at src/phylonium.cxx:287,291.  A maintainer keeps FASTA reading,
 * reference choice and printing and calls phylo_process() (or the three stage calls
 * below) from a process() stub; INTEGRATION.md shows that stub.  Every entry point
 * names the reference interface it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the context
 *     owns all device memory;
 *   - sequences are byte strings over {A,C,G,T,!} ('!' joins contigs,
 *     src/sequence.cxx:171-199), not NUL-terminated, lengths given explicitly;
 *   - every call returns PHYLO_OK (0) or a negative status; phylo_last_error() gives the
 *     message.  The reference calls errx() for fatal conditions (src/global.h:29-43);
 *     a host should map a non-zero status to errx(1, "%s", phylo_last_error(ctx));
 *   - a context is bound to one CUDA device and must be used from one thread at a time
 *     (process() is not re-entrant either: it reads globals, src/process.cxx:417-479);
 *   - there is NO CPU fallback: without a usable CUDA device phylo_ctx_create fails.
 */
#ifndef PHYLONIUM_B200_H
#define PHYLONIUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHYLO_OK 0
#define PHYLO_ERR_INVALID (-1) /* bad argument or call order */
#define PHYLO_ERR_CUDA (-2)    /* CUDA runtime failure, including "no device" */
#define PHYLO_ERR_INTERNAL (-3)

/* flags of phylo_compare_all / phylo_process; same bit as flags::complete_deletion,
 * src/global.h:11 */
#define PHYLO_FLAG_COMPLETE_DELETION 4

/* estimator kinds: evo_model::estimate_raw / estimate_JC / estimate_ani,
 * src/evo_model.cxx:100-131 */
#define PHYLO_DIST_RAW 0
#define PHYLO_DIST_JC 1
#define PHYLO_DIST_ANI 2

typedef struct phylo_ctx phylo_ctx;

/* class homology, src/process.h:14-26: direction 0 forward / 1 reverse, start on S,
 * start projected onto the forward strand, start on the query, length. */
typedef struct {
	int64_t direction;
	int64_t index_reference;
	int64_t index_reference_projected;
	int64_t index_query;
	int64_t length;
} phylo_homology;

/* ---- context ------------------------------------------------------------------ */

/* device: CUDA ordinal, or -1 for the current device. */
int phylo_ctx_create(int device, phylo_ctx **out);
void phylo_ctx_destroy(phylo_ctx *ctx);
/* message of the last failed call on ctx (ctx may be NULL for a failed create) */
const char *phylo_last_error(const phylo_ctx *ctx);
const char *phylo_version(void);

/* Issue all further work of ctx on `stream` (a cudaStream_t of ctx's device, e.g. the
 * framework's current stream); NULL restores the context's own stream.  The stream must
 * outlive the buffers allocated while it was set. */
int phylo_set_stream(phylo_ctx *ctx, void *stream);

/* tuning knobs, per context; results never depend on them (tests sweep them):
 *   "chunk"    bases per speculative walker (default 2048)
 *   "cap"      per-thread comparison cap (default 2 * chunk)
 *   "kmer_k"   K of the descent table, 0 = none, -1 = from the text length (default)
 *   "key_chars" characters per suffix-sort key (1..21), 0 = from the text length (default)
 *   "scan_mode" prefix scans: 1 = one launch with decoupled look-back (default), 0 = three
 *               launches (reduce, scan of the tile sums, apply)
 *   "map_batch_bytes" sequences are mapped in batches of about this many bytes (default
 *               512 MiB): bounds the scratch memory
 *   "table_direct" how the descent table is built: 0 / 1 = entry by entry from the root
 *               (default), 2 = level by level (cross-check)
 *   "sort_path" suffix sorter: 0 = packed 2-bit words when the key fits 16 characters and the
 *               reference has at most ~1000 contigs, else the general one (default); 1 = always
 *               the general sorter (3-bit codes, 64-bit keys)
 *   "sort_mode" radix sort scheme: 0 = by size (default), 1 = histogram+scan+scatter per
 *              pass, 2 = single-pass look-back ("onesweep")
 *   "stage_threads" worker threads that pack the sequences of phylo_process / phylo_map_queries
 *               to 2 bits per base for the trip over PCIe; 0 = from the core count (default)
 *   "push_kernel" sharded runs, rows of the last batch of a mapping: 1 (default) = written into
 *               the peers' stores by a kernel (stores over NVLink, three planes where enough),
 *               0 = copy engines like the earlier batches
 *   "esa_speculative" 1 (default) = the index build makes no host round trip until its end and
 *               is redone step by step if what it took for granted (valid input, few separators,
 *               no repeats beyond the direct comparisons) turns out wrong; 0 = always step by step
 *   "esa_graph" 1 (default) = the speculative build is submitted as one CUDA graph behind its
 *               first radix pass (re-captured per call, the instantiated graph updated); 0 =
 *               kernel by kernel
 *   "map_graph" the mapping of a batch up to its one host stop (validation, walk, lists, rows,
 *               comparison) submitted as one CUDA graph: 0 = never, 1 (default) = the first batch of
 *               a call when it holds 4 Mbp or more (the host records it while the index is being
 *               built), 2 = every batch (tests)
 *   "compare_path" how the all-pairs kernel brings the row tiles into shared memory: 0 = tensor
 *               copies (TMA) through a 6-stage transaction-barrier pipeline (default), 1 =
 *               cp.async by all threads, double buffered (kept for comparison)
 *   "upload_raw" 0 (default) = pack on the host unless the input is pinned and under 128 MiB
 *               (then the plain asynchronous copies are hidden behind the index build anyway),
 *               1 = always send the bytes as they are, -1 = always pack
 *   "keep_raw" keep unsorted/unfiltered homology lists for phylo_get_homologies(raw=1) (the rows of
 *               a batch are then built after its lists are final, not before the host has seen them)
 *   "timings"  record per-phase device times (adds synchronisation) */
int phylo_set_option(phylo_ctx *ctx, const char *key, int64_t value);
/* (The environment variable PHYLO_B200_OPTIONS="key=value,key=value" sets options on every
 * context a process creates — for A/B measurements through programs that do not pass them on.) */
/* last recorded value of a named timing/statistic, e.g. "esa.sort_ms", "anchor.walk_ms",
 * "compare.ms"; *out = -1 for unknown names */
int phylo_get_stat(const phylo_ctx *ctx, const char *key, double *out);

/* ---- host-side scalars ----------------------------------------------------------- */

/* gc_content(), src/sequence.cxx:152-165 */
double phylo_gc_content(const char *seq, uint64_t n);
/* min_anchor_length(), src/process.cxx:77-86 with shuprop :140-161; called by process()
 * as min_anchor_length(ANCHOR_P_VALUE = 0.025, gc, 2n + 1) (:416-417) */
uint64_t phylo_min_anchor_length(double p, double gc, uint64_t l);

/* The packing the library applies to sequences before they cross PCIe (host code, no GPU
 * needed; exported for tests): n bytes over {A,C,G,T,!} -> (n + 3) / 4 bytes, base k of a
 * group of four in bits 2k, 2k + 1, code (c >> 1) & 3 (A 0, C 1, T 2, G 3); '!' packs as 0 and
 * its position goes to bangs (at most cap entries are written, *nbangs counts all).  Returns 1
 * if a byte outside the alphabet was met, else 0. */
int phylo_host_pack_2bit(const char *seq, uint64_t n, uint8_t *packed, uint32_t *bangs, uint32_t cap, uint32_t *nbangs);

/* ---- stage 1: index ---------------------------------------------------------------- */

/* esa::esa(const sequence &), src/esa.cxx:69-81: S = ref '#' revcomp(ref), suffix array
 * (replaces divsufsort64, :73-75), LCP (:305-347), CLD (:256-298), FVC (:239-250) and the
 * descent table (replaces init_cache, :90-228).  Result stays on the device.
 * The call returns as soon as the build is queued on the context's stream and the reference's
 * alphabet has been checked (its first kernel's verdict): whatever is called next on the context
 * is ordered behind the build, phylo_map_queries queues its kernels there without waiting.  Should
 * the build's assumptions fail on the device (option "esa_speculative"), the next call that needs
 * the index builds it again step by step first; a caller never sees that. */
int phylo_esa_build(phylo_ctx *ctx, const char *ref, uint64_t n);
/* esa::size(), src/esa.h:78-81: m = 2n + 1 */
int phylo_esa_size(const phylo_ctx *ctx, uint64_t *m);
/* debug copy-out widened to the reference's saidx64_t; any pointer may be NULL.
 * SA: m, LCP: m+1, CLD: m+1 entries; FVC, S: m bytes (src/esa.h:53-66) */
int phylo_esa_get_arrays(const phylo_ctx *ctx, int64_t *SA, int64_t *LCP, int64_t *CLD, char *FVC, char *S);
/* esa::get_match_cached / get_match, src/esa.cxx:525-563, for a batch of strings:
 * string k is text[offs[k] .. offs[k] + lens[k]); out[3k..3k+2] = {l, i, j}.
 * use_table = 0 descends from the root (get_match) */
int phylo_esa_get_matches(phylo_ctx *ctx, const char *text, const uint64_t *offs, const uint64_t *lens,
                          uint64_t count, int use_table, int64_t *out);

/* ---- stage 2: anchoring -------------------------------------------------------------- */

/* hot loop A of process(), src/process.cxx:433-458, for all N sequences:
 * anchor_homologies(ref, threshold, query) (:198-295), std::sort by start (:438-441),
 * filter_overlaps_max (:354-401).  Lists and reference-coordinate rows stay on the
 * device for phylo_compare_all. */
int phylo_map_queries(phylo_ctx *ctx, const char *const *queries, const uint64_t *lens, uint64_t N,
                      uint64_t threshold);
/* number of homologies per sequence after filtering (raw = 0) or before sort/filter
 * (raw = 1, needs option keep_raw) */
int phylo_homology_counts(const phylo_ctx *ctx, uint64_t *counts, int raw);
int phylo_get_homologies(const phylo_ctx *ctx, uint64_t index, int raw, phylo_homology *out, uint64_t cap,
                         uint64_t *written);

/* ---- stage 3: all pairs ---------------------------------------------------------------- */

/* hot loop B of process(), src/process.cxx:524-549: compare() for every pair (:566-658)
 * after the optional complete_delete (:467-469, :725-776).  subst / homologs: N*N,
 * row-major, symmetric, zero diagonal — the two counters of evo_model
 * (src/evo_model.h:17-19). */
int phylo_compare_all(phylo_ctx *ctx, int flags, uint64_t *subst, uint64_t *homologs);
/* The core genome after complete deletion, for the reference's option -p (print reference
 * positions, src/process.cxx:471-513, get_segsites :665-723).  Call after phylo_map_queries
 * or phylo_process.  Three bitmaps of *words 32-bit words each (bit b of word w = reference
 * column 32 w + b; pass words = ceil(n / 32) rounded up to a multiple of 4, or query it with
 * all three pointers NULL):
 *   core    the column is covered by a homology of every sequence
 *   border  some sequence's homology starts at the column: a new "part" starts there
 *   seg     core column where some sequence differs from sequence 0 (bytes on the same
 *           strand, complement rule across strands)
 * A part of the reference's output is a maximal run of core columns without a border inside. */
int phylo_core_sites(phylo_ctx *ctx, uint32_t *core, uint32_t *border, uint32_t *seg, uint64_t *words);
/* evo_model::estimate_raw/JC/ani on the device for the last matrix; dist: N*N doubles,
 * diagonal 0 (src/io.cxx:157).  Hosts that must print bit-identical text should use
 * their own libm on the integer counts instead. */
int phylo_estimate(phylo_ctx *ctx, int kind, double *dist);

/* ---- the process() seam ---------------------------------------------------------------- */

/* process(subject = seqs[ref_index], queries = seqs), src/process.cxx:408-556:
 * phylo_esa_build + threshold + phylo_map_queries + phylo_compare_all. */
int phylo_process(phylo_ctx *ctx, const char *const *seqs, const uint64_t *lens, uint64_t N, uint64_t ref_index,
                  int flags, uint64_t *subst, uint64_t *homologs);

/* The second pass of the reference's --2pass option (src/phylonium.cxx:289-296): process()
 * again with another sequence as the reference.  The sequences of the last phylo_process
 * (or phylo_map_queries) call are still in device memory; nothing crosses the bus again. */
int phylo_process_again(phylo_ctx *ctx, uint64_t ref_index, int flags, uint64_t *subst, uint64_t *homologs);

/* Ingest pipelined with the upload (the reference reads all FASTA files, src/io.cxx:66-104 with
 * libs/pfasta.c, before process() sees a byte): the host announces N sequences with an upper
 * bound of each one's length (the file size will do), then hands every sequence over as soon
 * as its parser is done with it — from any thread, in any order.  phylo_ingest_put packs the
 * sequence to 2 bits per base, checks the alphabet, sends it and returns; the bytes are not
 * referenced afterwards.  `lanes` = how many puts can be in flight at once (one stream and one
 * small pinned ring each).  After phylo_ingest_end the sequences are resident:
 * phylo_process_again(ctx, ref_index, ...) is process() on them. */
int phylo_ingest_begin(phylo_ctx *ctx, uint64_t N, const uint64_t *max_lens, int lanes);
int phylo_ingest_put(phylo_ctx *ctx, uint64_t index, const char *seq, uint64_t len);
int phylo_ingest_end(phylo_ctx *ctx);

/* ---- device-resident variants (benchmarks, multi-GPU plumbing) --------------------------- */

/* Same stages with inputs/outputs already in device memory of ctx's device.
 * d_ref: n bytes.  d_queries: one buffer holding all sequences; sequence k occupies
 * [offs[k], offs[k] + lens[k]) and MUST be followed by at least one zero byte.
 * offs/lens are host arrays.  d_subst / d_homologs: N*N uint64 on the device. */
int phylo_esa_build_dev(phylo_ctx *ctx, const void *d_ref, uint64_t n);
int phylo_map_queries_dev(phylo_ctx *ctx, const void *d_queries, const uint64_t *offs, const uint64_t *lens,
                          uint64_t N, uint64_t threshold);
/* phylo_compare_all_dev / phylo_compare_tiles_dev return as soon as the work is queued on the
 * context's stream (phylo_set_stream): the counts are there in stream order, so a collective
 * or a copy queued on that stream next sees them without the host waiting in between. */
int phylo_compare_all_dev(phylo_ctx *ctx, int flags, void *d_subst, void *d_homologs);

/* Sharding across GPUs (one context per GPU, the exchange itself is the caller's
 * collective, e.g. NCCL through torch.distributed):
 *  - index: rank 0 builds it, the others call phylo_esa_alloc(n); all ranks fetch the
 *    five device arrays with phylo_esa_device_arrays and broadcast them in place; the
 *    receivers then call phylo_esa_finish_import (builds the local descent table);
 *  - rows: phylo_rows_configure(total, first) before phylo_map_queries[_dev] makes the
 *    context keep rows for `total` genomes and write its own at [first, first + N);
 *    phylo_rows_device gives the store for an all-gather (genome-major, contiguous);
 *  - matrix: phylo_compare_tiles_dev computes the tile pairs rank, rank + world, … of
 *    all `total` genomes; summing the partial matrices over ranks gives the full one. */
int phylo_esa_alloc(phylo_ctx *ctx, uint64_t n);
int phylo_esa_device_arrays(const phylo_ctx *ctx, void **S, uint64_t *S_bytes, void **SA, void **LCP, void **CLD,
                            void **FVC);
int phylo_esa_finish_import(phylo_ctx *ctx);
int phylo_rows_configure(phylo_ctx *ctx, uint64_t total_genomes, uint64_t first_row);
int phylo_rows_device(const phylo_ctx *ctx, void **rows, uint64_t *bytes_per_genome, uint64_t *total_genomes);
/* Exchange of the rows without a collective: every rank hands the address of its row store to
 * the others (across processes: a CUDA IPC handle of PHYLO_IPC_HANDLE_BYTES bytes from
 * phylo_rows_ipc_export, all ranks' handles concatenated in rank order into
 * phylo_rows_ipc_import; inside one process: the pointers of phylo_rows_device into
 * phylo_rows_set_peers, NULL to switch it off again).  From then on phylo_map_queries[_dev]
 * copies every batch of rows it builds into all peers' stores over NVLink while it maps the
 * next batch, and leaves the context's stream ordered behind those copies.  The caller needs
 * one barrier across ranks on that stream (any collective) between the mapping and
 * phylo_compare_tiles_dev, and another one before the next mapping overwrites the rows.
 * The handles die with the row store: exchange them again after a phylo_rows_configure that
 * changes the store's size. */
#define PHYLO_IPC_HANDLE_BYTES 64
int phylo_rows_ipc_export(phylo_ctx *ctx, void *handle);
int phylo_rows_ipc_import(phylo_ctx *ctx, const void *handles, int world, int rank);
int phylo_rows_set_peers(phylo_ctx *ctx, void *const *peer_rows, int world, int rank);
int phylo_compare_tiles_dev(phylo_ctx *ctx, int flags, int rank, int world, void *d_subst, void *d_homologs);

#ifdef __cplusplus
}
#endif

#endif /* PHYLONIUM_B200_H */
