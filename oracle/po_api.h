/* TEST INFRASTRUCTURE — C interface shared by the two CPU checkers:
 *   oracle/_build/libphylo_oracle.so  (phylo_oracle.cxx: our CPU restatement, "port")
 *   oracle/_ref/libphylo_ref.so       (ref_shim.cxx + the UNMODIFIED reference sources)
 * Both export exactly these symbols so tests can run either through one ctypes
 * wrapper (tests/oracle_lib.py).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load these libraries.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors class homology, /root/reference/src/process.h:14-26 (40 bytes) */
typedef struct {
	int64_t direction; /* 0 forward, 1 reverse */
	int64_t index_reference;
	int64_t index_reference_projected;
	int64_t index_query;
	int64_t length;
} po_hom;

const char *po_kind(void); /* "port" or "reference" */

/* sequence helpers: src/sequence.cxx:73-103, 109-146, 152-165 */
void po_revcomp(const char *in, int64_t n, char *out);
int64_t po_filter_nucl(const char *in, int64_t n, char *out);
double po_gc_content(const char *seq, int64_t n);

/* leaf comparators: libs/seqcmp.c, libs/revseqcmp.c */
uint64_t po_seqcmp(const char *a, const char *b, uint64_t len);
uint64_t po_revseqcmp(const char *a, const char *b, uint64_t len);

/* threshold maths: src/process.cxx:77-161 */
int64_t po_min_anchor_length(double p, double gc, int64_t l);

/* ESA: src/esa.cxx:69-81 and friends. n = reference length, m = 2n+1 */
void *po_esa_create(const char *ref, int64_t n);
void po_esa_destroy(void *esa);
int64_t po_esa_size(void *esa);
/* any pointer may be NULL. SA:m, LCP:m+1, CLD:m+1 (int64), FVC:m, S:m (bytes) */
void po_esa_arrays(void *esa, int64_t *SA, int64_t *LCP, int64_t *CLD, char *FVC, char *S);
/* out = {l, i, j}; cached != 0 uses get_match_cached (src/esa.cxx:542-563) */
void po_get_match(void *esa, const char *query, int64_t qlen, int cached, int64_t out[3]);

/* anchoring: src/process.cxx:198-295 (raw list, push order) */
int64_t po_anchor_homologies(void *esa, int64_t threshold, const char *query, int64_t qlen,
                             po_hom *out, int64_t cap);
/* std::sort by start + filter_overlaps_max: src/process.cxx:438-443, 354-401. in place */
int64_t po_sort_filter(po_hom *h, int64_t count, int do_sort);
/* list x list comparison: src/process.cxx:566-658. out = {substitutions, homologs} */
void po_compare(const char *qa, const po_hom *ha, int64_t na, const char *qb, const po_hom *hb,
                int64_t nb, uint64_t out[2]);
/* complete deletion: src/process.cxx:725-776. lists concatenated, offs has N+1 entries;
 * out_offs gets N+1 entries; returns total written (<= cap) */
int64_t po_complete_delete(const po_hom *h, const int64_t *offs, int64_t N, po_hom *out,
                           int64_t *out_offs, int64_t cap);

/* whole pipeline: src/process.cxx:408-556. subst/homologs are N*N row-major.
 * flags bit 2 (=4): complete deletion (src/global.h:11).
 * timings (may be NULL) = seconds for {esa build, loop A, loop B, SA sort only}.
 * hom_counts (may be NULL) = number of filtered homologies per query. */
int po_process(const char *const *seqs, const int64_t *lens, int64_t N, int64_t ref_index,
               int flags, int threads, uint64_t *subst, uint64_t *homologs, double *timings,
               int64_t *hom_counts);

/* the same for a sample of the matrix: every sequence is mapped, rows[0..nrows) of the
 * matrix are computed (subst/homologs: nrows*N, row r = sequence rows[r] against all);
 * timings as above */
int po_process_rows(const char *const *seqs, const int64_t *lens, int64_t N, int64_t ref_index,
                    int flags, int threads, const int64_t *rows, int64_t nrows, uint64_t *subst,
                    uint64_t *homologs, double *timings);

/* distances and printing: src/evo_model.cxx:100-131, src/io.cxx:141-163.
 * kind 0 = raw, 1 = JC, 2 = ANI */
double po_estimate(uint64_t subst, uint64_t homologs, int kind);
int64_t po_format_matrix(const char *const *names, const uint64_t *subst,
                         const uint64_t *homologs, int64_t N, int kind, char *out, int64_t cap);

/* test/simf.cxx:93-140: one simulated genome, bases only (no FASTA framing).
 * divergence is the JC distance unless raw != 0 (simf.cxx:62-68). */
void po_simf(uint32_t base_seed, uint32_t mut_seed, int64_t length, double divergence, int raw,
             char *out);

#ifdef __cplusplus
}
#endif
