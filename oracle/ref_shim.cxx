/* TEST INFRASTRUCTURE — builds into oracle/_ref/libphylo_ref.so together with the
 * UNMODIFIED reference sources, compiled where they lie under /root/reference
 * (recipe: oracle/Makefile; nothing from the reference is copied into this
 * repository).  This file only adapts the reference's C++ entry points to the
 * flat C interface of po_api.h so that tests can compare them with the CPU
 * restatement (phylo_oracle.cxx) and with the CUDA path.
 *
 * process.cxx is included textually because anchor_homologies() has a deduced
 * return type and filter_overlaps_max()/compare() have no header
 * (src/process.cxx:198,354,566).  `private`/`protected` are opened for this
 * translation unit only, to read esa::LCP/CLD/FVC (src/esa.h:53-64) and
 * evo_model::substitutions (src/evo_model.h:17-19).
 */
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <err.h>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <sys/types.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public
#define protected public
#define class struct // esa's leading members are private by default (src/esa.h:45-64)
#include "esa.h"     // /root/reference/src/esa.h (no `template <class …>` inside)
#undef class
#include "process.cxx" // /root/reference/src/process.cxx
#undef private
#undef protected

#include "revseqcmp.h"
#include "seqcmp.h"

#include "po_api.h"

/* globals the reference expects from its main(): src/phylonium.cxx:55-64 */
double ANCHOR_P_VALUE = 0.025;
int FLAGS = flags::none;
int THREADS = 1;
long unsigned int BOOTSTRAP = 0;
int RETURN_CODE = EXIT_SUCCESS;
std::string REFPOS_FILE_NAME = "";
std::mt19937 prng;
size_t reference_index = 0;

extern "C" double po_sa_seconds; // accumulated by sa_standin.cxx

namespace
{
double now()
{
	using namespace std::chrono;
	return duration<double>(steady_clock::now().time_since_epoch()).count();
}

homology to_hom(const po_hom &p)
{
	homology h;
	h.direction = p.direction ? homology::dir::reverse : homology::dir::forward;
	h.index_reference = (size_t)p.index_reference;
	h.index_reference_projected = (size_t)p.index_reference_projected;
	h.index_query = (size_t)p.index_query;
	h.length = (size_t)p.length;
	return h;
}

po_hom from_hom(const homology &h)
{
	po_hom p;
	p.direction = h.direction == homology::dir::reverse ? 1 : 0;
	p.index_reference = (int64_t)h.index_reference;
	p.index_reference_projected = (int64_t)h.index_reference_projected;
	p.index_query = (int64_t)h.index_query;
	p.length = (int64_t)h.length;
	return p;
}
} // namespace

extern "C" {

const char *po_kind(void)
{
	return "reference";
}

void po_revcomp(const char *in, int64_t n, char *out)
{
	auto r = reverse(std::string(in, (size_t)n));
	std::memcpy(out, r.data(), r.size());
}

int64_t po_filter_nucl(const char *in, int64_t n, char *out)
{
	auto r = filter_nucl(std::string(in, (size_t)n));
	std::memcpy(out, r.data(), r.size());
	return (int64_t)r.size();
}

double po_gc_content(const char *seq, int64_t n)
{
	return gc_content(std::string(seq, (size_t)n));
}

uint64_t po_seqcmp(const char *a, const char *b, uint64_t len)
{
	return seqcmp(a, b, len);
}

uint64_t po_revseqcmp(const char *a, const char *b, uint64_t len)
{
	return revseqcmp(a, b, len);
}

int64_t po_min_anchor_length(double p, double gc, int64_t l)
{
	return (int64_t)min_anchor_length(p, gc, (size_t)l);
}

void *po_esa_create(const char *ref, int64_t n)
{
	return new esa(sequence("ref", std::string(ref, (size_t)n)));
}

void po_esa_destroy(void *e)
{
	delete static_cast<esa *>(e);
}

int64_t po_esa_size(void *e)
{
	return static_cast<esa *>(e)->size();
}

void po_esa_arrays(void *ev, int64_t *SA, int64_t *LCP, int64_t *CLD, char *FVC, char *S)
{
	auto *e = static_cast<esa *>(ev);
	size_t m = (size_t)e->size();
	if (SA) std::copy(e->SA.get(), e->SA.get() + m, SA);
	if (LCP) std::copy(e->LCP.get(), e->LCP.get() + m + 1, LCP);
	if (CLD) std::copy(e->CLD.get(), e->CLD.get() + m + 1, CLD);
	if (FVC) std::memcpy(FVC, e->FVC.get(), m);
	if (S) std::memcpy(S, e->S.data(), m);
}

void po_get_match(void *ev, const char *query, int64_t qlen, int cached, int64_t out[3])
{
	auto *e = static_cast<esa *>(ev);
	// the reference always hands in NUL-terminated std::string storage
	std::string q(query, (size_t)qlen);
	lcp_interval r = cached ? e->get_match_cached(q.c_str(), (size_t)qlen)
	                        : e->get_match(q.c_str(), (size_t)qlen);
	out[0] = r.l;
	out[1] = r.i;
	out[2] = r.j;
}

int64_t po_anchor_homologies(void *ev, int64_t threshold, const char *query, int64_t qlen,
                             po_hom *out, int64_t cap)
{
	auto *e = static_cast<esa *>(ev);
	auto hv = anchor_homologies(*e, (size_t)threshold, sequence("q", std::string(query, (size_t)qlen)));
	int64_t k = 0;
	for (const auto &h : hv) {
		if (k < cap) out[k] = from_hom(h);
		k++;
	}
	return k;
}

int64_t po_sort_filter(po_hom *h, int64_t count, int do_sort)
{
	std::vector<homology> pile;
	for (int64_t k = 0; k < count; k++)
		pile.push_back(to_hom(h[k]));
	if (do_sort) {
		// same call as src/process.cxx:438-441
		std::sort(begin(pile), end(pile), [](const homology &self, const homology &other) {
			return self.starts_left_of(other);
		});
	}
	filter_overlaps_max(pile);
	for (size_t k = 0; k < pile.size(); k++)
		h[k] = from_hom(pile[k]);
	return (int64_t)pile.size();
}

void po_compare(const char *qa, const po_hom *ha, int64_t na, const char *qb, const po_hom *hb,
                int64_t nb, uint64_t out[2])
{
	// sequences are only dereferenced inside homologous ranges; find the extent
	size_t la = 0, lb = 0;
	std::vector<homology> va, vb;
	for (int64_t k = 0; k < na; k++) {
		va.push_back(to_hom(ha[k]));
		la = std::max(la, va.back().end_query());
	}
	for (int64_t k = 0; k < nb; k++) {
		vb.push_back(to_hom(hb[k]));
		lb = std::max(lb, vb.back().end_query());
	}
	auto sa = sequence("a", std::string(qa, la));
	auto sb = sequence("b", std::string(qb, lb));
	evo_model em = compare(sa, va, sb, vb);
	out[0] = em.substitutions;
	out[1] = em.homologs;
}

int64_t po_complete_delete(const po_hom *h, const int64_t *offs, int64_t N, po_hom *out,
                           int64_t *out_offs, int64_t cap)
{
	std::vector<std::vector<homology>> in((size_t)N);
	for (int64_t g = 0; g < N; g++)
		for (int64_t k = offs[g]; k < offs[g + 1]; k++)
			in[(size_t)g].push_back(to_hom(h[k]));
	auto core = complete_delete(in);
	int64_t w = 0;
	for (int64_t g = 0; g < N; g++) {
		out_offs[g] = w;
		for (const auto &x : core[(size_t)g]) {
			if (w < cap) out[w] = from_hom(x);
			w++;
		}
	}
	out_offs[N] = w;
	return w;
}

int po_process(const char *const *seqs, const int64_t *lens, int64_t N, int64_t ref_index,
               int flags_, int threads, uint64_t *subst, uint64_t *homologs, double *timings,
               int64_t *hom_counts)
{
	std::vector<sequence> queries;
	for (int64_t g = 0; g < N; g++)
		queries.emplace_back("g" + std::to_string(g), std::string(seqs[g], (size_t)lens[g]));

	FLAGS = flags_ & flags::complete_deletion;
	THREADS = threads > 0 ? threads : 1;
	reference_index = (size_t)ref_index;

	std::vector<evo_model> matrix;
	if (!timings && !hom_counts) {
		matrix = process(queries[(size_t)ref_index], queries); // the real thing
	} else {
		/* Same statements as process() (src/process.cxx:408-556) with clocks
		 * between the phases; po_process(…, NULL, NULL) above is the untouched
		 * call and tests check both give identical counts. */
		const sequence &subject = queries[(size_t)ref_index];
		po_sa_seconds = 0;
		double t0 = now();
		auto ref = esa(subject);
		double t1 = now();
		auto gc = gc_content(subject.get_nucl());
		size_t threshold = min_anchor_length(ANCHOR_P_VALUE, gc, ref.size());
		auto homologies = std::vector<std::vector<homology>>((size_t)N);
#pragma omp parallel for num_threads(THREADS)
		for (size_t j = 0; j < (size_t)N; j++) {
			auto query = queries[j];
			auto hvlocal = anchor_homologies(ref, threshold, query);
			std::sort(begin(hvlocal), end(hvlocal), [](const homology &self, const homology &other) {
				return self.starts_left_of(other);
			});
			filter_overlaps_max(hvlocal);
#pragma omp critical
			homologies[j] = std::move(hvlocal);
		}
		double t2 = now();
		if (FLAGS & flags::complete_deletion) homologies = complete_delete(homologies);
		if (hom_counts)
			for (size_t j = 0; j < (size_t)N; j++)
				hom_counts[j] = (int64_t)homologies[j].size();
		matrix = std::vector<evo_model>((size_t)(N * N));
		double t3 = now();
#pragma omp parallel for num_threads(THREADS)
		for (size_t i = 0; i < (size_t)N; i++) {
			for (size_t j = i + 1; j < (size_t)N; j++) {
				matrix[j * N + i] = matrix[i * N + j] =
					compare(queries[i], homologies[i], queries[j], homologies[j]);
			}
		}
		double t4 = now();
		if (timings) {
			timings[0] = t1 - t0;
			timings[1] = t2 - t1;
			timings[2] = t4 - t3;
			timings[3] = po_sa_seconds;
		}
	}
	for (size_t k = 0; k < (size_t)(N * N); k++) {
		subst[k] = matrix[k].substitutions;
		homologs[k] = matrix[k].homologs;
	}
	return 0;
}

/* process() for a sample of the matrix: every sequence is mapped like in process()
 * (src/process.cxx:433-458), then only the listed rows are compared against all
 * sequences (:524-549 for those i) — what a benchmark at 1000 genomes can afford. */
int po_process_rows(const char *const *seqs, const int64_t *lens, int64_t N, int64_t ref_index, int flags_,
                    int threads, const int64_t *rows, int64_t nrows, uint64_t *subst, uint64_t *homologs,
                    double *timings)
{
	std::vector<sequence> queries;
	for (int64_t g = 0; g < N; g++)
		queries.emplace_back("g" + std::to_string(g), std::string(seqs[g], (size_t)lens[g]));
	FLAGS = flags_ & flags::complete_deletion;
	THREADS = threads > 0 ? threads : 1;
	reference_index = (size_t)ref_index;
	const sequence &subject = queries[(size_t)ref_index];
	po_sa_seconds = 0;
	double t0 = now();
	auto ref = esa(subject);
	double t1 = now();
	auto gc = gc_content(subject.get_nucl());
	size_t threshold = min_anchor_length(ANCHOR_P_VALUE, gc, ref.size());
	auto homologies = std::vector<std::vector<homology>>((size_t)N);
#pragma omp parallel for num_threads(THREADS) schedule(dynamic)
	for (size_t j = 0; j < (size_t)N; j++) {
		auto hvlocal = anchor_homologies(ref, threshold, queries[j]);
		std::sort(begin(hvlocal), end(hvlocal), [](const homology &self, const homology &other) {
			return self.starts_left_of(other);
		});
		filter_overlaps_max(hvlocal);
		homologies[j] = std::move(hvlocal);
	}
	double t2 = now();
	if (FLAGS & flags::complete_deletion) homologies = complete_delete(homologies);
	double t3 = now();
#pragma omp parallel for num_threads(THREADS) schedule(dynamic) collapse(2)
	for (int64_t r = 0; r < nrows; r++) {
		for (int64_t j = 0; j < N; j++) {
			const size_t i = (size_t)rows[r];
			evo_model em;
			if ((size_t)j != i) em = compare(queries[i], homologies[i], queries[(size_t)j], homologies[(size_t)j]);
			subst[r * N + j] = em.substitutions;
			homologs[r * N + j] = em.homologs;
		}
	}
	double t4 = now();
	if (timings) {
		timings[0] = t1 - t0;
		timings[1] = t2 - t1;
		timings[2] = t4 - t3;
		timings[3] = po_sa_seconds;
	}
	return 0;
}

double po_estimate(uint64_t subst, uint64_t homologs, int kind)
{
	evo_model em;
	em.substitutions = subst;
	em.homologs = homologs;
	return kind == 0 ? em.estimate_raw() : kind == 2 ? em.estimate_ani() : em.estimate_JC();
}

} // extern "C"
