/* TEST INFRASTRUCTURE — CPU restatement ("port") of phylonium's distance pipeline.
 *
 * This is the oracle the CUDA path is checked against.  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load oracle/_build/libphylo_oracle.so.
 *
 * Every function restates, in its own words, the algorithm of the reference file
 * cited above it (paths relative to /root/reference).  Parity of this restatement
 * is PINNED: tests/test_oracle_vs_reference.py runs it side by side with
 * oracle/_ref/libphylo_ref.so — the unmodified reference sources compiled by
 * oracle/Makefile — on simulated, multi-contig, reverse-strand and repetitive
 * inputs and requires identical SA/LCP/CLD/FVC, matches, homologies, pair counts
 * and PHYLIP text; tests/test_golden.py re-checks committed fixtures generated from
 * the reference, plus the reference's own known-answer tests
 * (test/Tprocess.cxx:19-123, test/Tsequence.cxx:14-42).
 *
 * Third-party piece: divsufsort64 (libdivsufsort, absent from the image) is
 * replaced by oracle/sa_standin.cxx — see the note there.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "divsufsort64.h"
#include "po_api.h"

extern "C" double po_sa_seconds;

namespace
{

using i64 = int64_t;

double now()
{
	using namespace std::chrono;
	return duration<double>(steady_clock::now().time_since_epoch()).count();
}

/* ---------------------------------------------------------------- sequence */

/* src/sequence.cxx:73-103 — reverse complement; bytes below 'A' (the '!' contig
 * separator) are kept as they are; A<->T is a flip of 0x15, C<->G of 0x04. */
std::string revcomp(const char *s, i64 n)
{
	std::string r((size_t)n, '\0');
	for (i64 k = 0; k < n; k++) {
		unsigned char c = (unsigned char)s[n - 1 - k];
		if (c >= 'A') c ^= (c & 2) ? 4 : 21;
		r[(size_t)k] = (char)c;
	}
	return r;
}

/* ---------------------------------------------------------------- ESA */

struct interval {
	i64 l, i, j, m;
};

const int CACHE_K = 6; /* src/esa.cxx:34 */

struct esa_t {
	i64 n = 0, m = 0;
	std::string S; /* R # revcomp(R); std::string keeps S[m] == 0 */
	std::vector<i64> SA, LCP, CLD;
	std::vector<char> FVC;
	std::vector<interval> cache;

	i64 lchild(i64 k) const { return CLD[(size_t)(k - 1)]; } /* src/esa.h:107-110 */
	i64 rchild(i64 k) const { return CLD[(size_t)k]; }       /* src/esa.h:98-101 */
	interval root() const
	{
		i64 mr = lchild(m); /* src/esa.cxx:527-528 */
		return {LCP[(size_t)mr], 0, m - 1, mr};
	}
};

/* src/esa.cxx:305-347 — LCP via the PHI trick: PHI[SA[r]] = SA[r-1], then walk the
 * text left to right carrying l-1 over, finally permute back into SA order. */
void build_lcp(esa_t &e)
{
	const i64 m = e.m;
	e.LCP.assign((size_t)m + 1, 0);
	e.LCP[0] = -1;
	e.LCP[(size_t)m] = -1;
	std::vector<i64> phi((size_t)m);
	phi[(size_t)e.SA[0]] = -1;
	for (i64 r = 1; r < m; r++)
		phi[(size_t)e.SA[(size_t)r]] = e.SA[(size_t)r - 1];
	i64 l = 0;
	const char *S = e.S.c_str();
	for (i64 p = 0; p < m; p++) {
		i64 k = phi[(size_t)p];
		if (k < 0) {
			phi[(size_t)p] = -1;
			continue;
		}
		while (S[k + l] == S[p + l]) /* stops at the NUL after S at the latest */
			l++;
		phi[(size_t)p] = l;
		if (l > 0) l--;
	}
	for (i64 r = 1; r < m; r++)
		e.LCP[(size_t)r] = phi[(size_t)e.SA[(size_t)r]];
}

/* src/esa.cxx:256-298 — child table in one pass with a stack of (index, lcp).
 * up and down values share a slot with the nextlIndex values (Abouelhoda et al.). */
void build_cld(esa_t &e)
{
	const i64 m = e.m;
	e.CLD.assign((size_t)m + 1, 0);
	struct entry {
		i64 idx, lcp;
	};
	std::vector<entry> st;
	st.reserve(64);
	e.CLD[0] = m;
	st.push_back({0, -1});
	for (i64 k = 1; k <= m; k++) {
		const i64 cur = e.LCP[(size_t)k];
		while (cur < st.back().lcp) {
			entry last = st.back();
			st.pop_back();
			while (st.back().lcp == last.lcp) { /* chain equal lcp values */
				e.CLD[(size_t)st.back().idx] = last.idx;
				last = st.back();
				st.pop_back();
			}
			if (cur < st.back().lcp)
				e.CLD[(size_t)st.back().idx] = last.idx; /* down */
			else
				e.CLD[(size_t)k - 1] = last.idx; /* up, stored left of k */
		}
		st.push_back({k, cur});
	}
}

/* src/esa.cxx:239-250 — first variant character. FVC[0] reads S[SA[0]-1]
 * because LCP[0] is -1; SA[0] > 0 always (S[0] is a nucleotide). */
void build_fvc(esa_t &e)
{
	e.FVC.assign((size_t)e.m, 0);
	for (i64 r = 0; r < e.m; r++) {
		i64 at = e.SA[(size_t)r] + e.LCP[(size_t)r];
		e.FVC[(size_t)r] = at >= 0 ? e.S[(size_t)at] : '\0';
	}
}

/* src/esa.cxx:361-427 — child interval of ij whose suffixes continue with `a`. */
interval get_interval(const esa_t &e, interval ij, char a)
{
	const char *S = e.S.c_str();
	i64 i = ij.i, j = ij.j;
	if (i == j) {
		if (S[e.SA[(size_t)i] + ij.l] != a) ij.i = ij.j = -1;
		return ij;
	}
	int m = (int)ij.m; /* int on purpose: src/esa.cxx:374-375 */
	int l = (int)ij.l;
	char c = S[e.SA[(size_t)i] + l];
	for (;;) {
		if (c == a) {
			if (i != m - 1) {
				i64 nm = e.lchild(m);
				return {e.LCP[(size_t)nm], i, (i64)m - 1, nm};
			}
			return {e.LCP[(size_t)i], i, i, -1};
		}
		if (c > a) break;
		i = m;
		if (i == j) break;
		m = (int)e.rchild(m);
		if (e.LCP[(size_t)m] != l) break;
		c = e.FVC[(size_t)i];
	}
	bool hit = (i != ij.i) ? e.FVC[(size_t)i] == a : S[e.SA[(size_t)i] + l] == a;
	if (!hit) {
		ij.i = ij.j = -1;
		return ij;
	}
	ij.i = i;
	ij.j = j;
	ij.l = e.LCP[(size_t)m];
	ij.m = m;
	return ij;
}

/* src/esa.cxx:446-513 — continue a match of query[0..k) that sits in interval ij. */
interval get_match_from(const esa_t &e, const char *query, size_t qlen, i64 k, interval ij)
{
	const char *S = e.S.c_str();
	if (ij.i == -1 && ij.j == -1) return ij;
	if (ij.i == ij.j) {
		i64 p = e.SA[(size_t)ij.i];
		size_t kk = (size_t)ij.l;
		for (; kk < qlen && S[p + (i64)kk]; kk++)
			if (S[p + (i64)kk] != query[kk]) break;
		ij.l = (i64)kk;
		return ij;
	}
	interval res = ij;
	do {
		ij = get_interval(e, ij, query[k]);
		if (ij.i == -1 && ij.j == -1) {
			res.l = k;
			return res;
		}
		res.i = ij.i;
		res.j = ij.j;
		i64 l = (i64)qlen;
		if (ij.i < ij.j && ij.l < l) l = ij.l;
		k++;
		int p = (int)e.SA[(size_t)ij.i]; /* int on purpose: src/esa.cxx:503 */
		for (; k < l; k++) {
			if (S[p + k] != query[k]) {
				res.l = k;
				return res;
			}
		}
	} while (k < (i64)qlen);
	res.l = (i64)qlen;
	return res;
}

interval get_match(const esa_t &e, const char *q, size_t qlen) /* src/esa.cxx:525-531 */
{
	return get_match_from(e, q, qlen, 0, e.root());
}

int code_of(char c) /* src/esa.cxx:48-62 */
{
	switch (c) {
		case 'A': return 0;
		case 'C': return 1;
		case 'G': return 2;
		case 'T': return 3;
	}
	return -1;
}

/* src/esa.cxx:90-228 keeps, per 6-mer, an interval from which get_match_from can be
 * resumed.  SURVEY.md §8a6/A.4: the cache is a pure accelerator — the cached search
 * returns exactly what the uncached one does (tests check this against the reference
 * for every query position).  We therefore restate its CONTRACT rather than its DFS:
 * cache[w] is the deepest interval reached by the plain descent on w that has all of
 * its ij.l <= 6 characters verified, so resuming at k = ij.l repeats no decision. */
void build_cache(esa_t &e)
{
	const size_t entries = (size_t)1 << (2 * CACHE_K);
	e.cache.assign(entries, interval{0, -1, -1, -1});
	const char *S = e.S.c_str();
	char w[CACHE_K + 1] = {0};
	for (size_t code = 0; code < entries; code++) {
		for (int t = 0; t < CACHE_K; t++)
			w[t] = "ACGT"[(code >> (2 * (CACHE_K - 1 - t))) & 3];
		interval ij = e.root();
		while (ij.i != ij.j && ij.l < CACHE_K) {
			interval nx = get_interval(e, ij, w[ij.l]);
			if (nx.i == -1 && nx.j == -1) break;
			if (nx.i == nx.j) {
				/* singleton: verify the rest of w by hand, keep what matched */
				i64 p = e.SA[(size_t)nx.i];
				i64 k = ij.l + 1;
				while (k < CACHE_K && S[p + k] == w[k])
					k++;
				nx.l = k;
				ij = nx;
				break;
			}
			i64 upto = std::min<i64>(nx.l, CACHE_K);
			i64 p = e.SA[(size_t)nx.i];
			i64 k = ij.l + 1;
			while (k < upto && S[p + k] == w[k])
				k++;
			if (k < upto || nx.l > CACHE_K) break; /* mismatch, or too deep to be resumable */
			ij = nx;
		}
		e.cache[code] = ij;
	}
}

interval get_match_cached(const esa_t &e, const char *q, size_t qlen) /* src/esa.cxx:542-563 */
{
	if (qlen <= (size_t)CACHE_K) return get_match(e, q, qlen);
	size_t code = 0;
	for (int t = 0; t < CACHE_K; t++) {
		int c = code_of(q[t]);
		if (c < 0) return get_match(e, q, qlen);
		code = (code << 2) | (size_t)c;
	}
	interval ij = e.cache[code];
	if (ij.i == -1 && ij.j == -1) return get_match(e, q, qlen);
	return get_match_from(e, q, qlen, ij.l, ij);
}

esa_t *build_esa(const char *ref, i64 n) /* src/esa.cxx:69-81 */
{
	auto *e = new esa_t;
	e->n = n;
	e->m = 2 * n + 1;
	e->S.assign(ref, (size_t)n);
	e->S += '#';
	e->S += revcomp(ref, n);
	e->SA.assign((size_t)e->m, 0);
	divsufsort64((const unsigned char *)e->S.c_str(), e->SA.data(), e->m);
	build_lcp(*e);
	build_cld(*e);
	build_fvc(*e);
	build_cache(*e);
	return e;
}

/* ---------------------------------------------------------------- thresholds */

/* src/process.cxx:103-125 */
size_t binom(size_t n, size_t k)
{
	if (n == 0 || k > n) return 0;
	if (k == 0 || k == n) return 1;
	if (k > n - k) k = n - k;
	size_t r = 1;
	for (size_t t = 1; t <= k; t++) {
		r *= n - k + t;
		r /= t;
	}
	return r;
}

/* src/process.cxx:140-161 — P{shustring length <= x}, Haubold et al. 2009 */
double shuprop(size_t x, double p, size_t l)
{
	double xx = (double)x, ll = (double)l, s = 0.0;
	for (size_t k = 0; k <= x; k++) {
		double kk = (double)k;
		double t = pow(p, kk) * pow(0.5 - p, xx - kk);
		s += pow(2, xx) * (t * pow(1 - t, ll)) * (double)binom(x, k);
		if (s >= 1.0) {
			s = 1.0;
			break;
		}
	}
	return s;
}

/* ---------------------------------------------------------------- homologies */

struct hom {
	i64 dir = 0, iref = 0, iproj = 0, iq = 0, len = 0;
	i64 start() const { return iproj; }
	i64 end() const { return iproj + len; }
	/* src/process.h:72-80 */
	void project(i64 n)
	{
		if (iref < n) return;
		iproj = 2 * n + 1 - len - iref;
		dir = 1;
	}
	/* src/process.h:86-117 */
	bool ends_left_of(const hom &o) const { return end() <= o.start(); }
	bool starts_left_of(const hom &o) const { return start() < o.start(); }
	bool overlaps(const hom &o) const
	{
		if (start() == o.start()) return true;
		if (starts_left_of(o)) return !ends_left_of(o);
		return !o.ends_left_of(*this);
	}
	/* src/process.h:119-143 */
	hom trim(i64 s, i64 e) const
	{
		if (e <= s) return *this;
		hom t = *this;
		i64 off = (s > start() && s < end()) ? s - start() : 0;
		i64 drift = (end() > e && e > start()) ? end() - e : 0;
		t.iproj += off;
		if (dir == 0) {
			t.iref += off;
			t.iq += off;
		} else {
			t.iref += drift;
			t.iq += drift;
		}
		t.len = len - off - drift;
		return t;
	}
};

hom make_hom(i64 ir, i64 iq, i64 l)
{
	hom h;
	h.iref = h.iproj = ir;
	h.iq = iq;
	h.len = l;
	return h;
}

/* src/process.cxx:171-184 */
i64 common_prefix(const char *a, const char *b, i64 limit)
{
	i64 k = 0;
	while (k < limit && a[k] == b[k])
		k++;
	return k;
}

/* src/process.cxx:198-295 — the anchor walk. */
std::vector<hom> anchor_homologies(const esa_t &e, i64 thr, const char *Q, i64 qlen)
{
	std::vector<hom> out;
	const i64 border = e.m / 2;
	const char *S = e.S.c_str();
	i64 lastQ = 0, lastS = 0, lastLen = 0;
	bool last_right = false;
	i64 pos = 0;
	hom cur = make_hom(0, 0, 0);

	while (pos < qlen) {
		i64 posS = 0, len = 0;
		bool ok = false;
		/* "lucky": stay on the diagonal of the previous anchor (:227-242) */
		i64 advance = pos - lastQ;
		i64 gap = advance - lastLen;
		i64 tryS = lastS + advance;
		if (tryS < e.m && gap <= thr) {
			posS = tryS;
			len = common_prefix(Q + pos, S + tryS, qlen - pos);
			ok = len >= thr;
		}
		if (!ok) { /* ESA search (:219-225) */
			interval in = get_match_cached(e, Q + pos, (size_t)(qlen - pos));
			len = std::max<i64>(in.l, 0);
			posS = e.SA[(size_t)in.i];
			ok = in.i == in.j && len >= thr;
		}
		if (ok) {
			i64 endS = lastS + lastLen, endQ = lastQ + lastLen;
			if (posS > endS && pos - endQ == posS - endS && (posS < border) == (lastS < border)) {
				cur.len += pos - endQ + len; /* right anchor: extend */
				last_right = true;
			} else {
				if (last_right || lastLen / 2 >= thr) {
					cur.project(border);
					out.push_back(cur);
				}
				cur = make_hom(posS, pos, len);
				last_right = false;
			}
			lastQ = pos;
			lastS = posS;
			lastLen = len;
		}
		pos += len + 1;
	}
	if (lastLen >= qlen) cur = make_hom(lastS, 0, qlen); /* :284-287 */
	if (last_right || lastLen / 2 >= thr) {
		cur.project(border);
		out.push_back(cur);
	}
	return out;
}

/* src/process.cxx:354-401 — keep the heaviest chain of non-overlapping homologies;
 * first-maximum tie rules as in the original (strict '>' over k, max_element). */
void filter_overlaps_max(std::vector<hom> &pile)
{
	const i64 size = (i64)pile.size();
	if (size < 2) return;
	std::vector<i64> pred((size_t)size + 1, -1), score((size_t)size + 1, 0);
	i64 *P = pred.data() + 1, *Sc = score.data() + 1; /* slot -1 holds score 0 */
	P[0] = -1;
	Sc[0] = pile[0].len;
	for (i64 i = 1; i < size; i++) {
		i64 best = 0, at = -1;
		for (i64 k = 0; k < i; k++) {
			if (!pile[(size_t)k].ends_left_of(pile[(size_t)i])) continue;
			if (Sc[k] > best) {
				best = Sc[k];
				at = k;
			}
		}
		P[i] = at;
		Sc[i] = Sc[at] + pile[(size_t)i].len;
	}
	i64 top = (i64)(std::max_element(score.begin(), score.end()) - score.begin()) - 1;
	std::vector<char> keep((size_t)size, 0);
	for (i64 k = top; k >= 0; k = P[k])
		keep[(size_t)k] = 1;
	size_t w = 0;
	for (i64 k = 0; k < size; k++)
		if (keep[(size_t)k]) pile[w++] = pile[(size_t)k];
	pile.resize(w);
}

void sort_by_start(std::vector<hom> &v) /* src/process.cxx:438-441, same std::sort */
{
	std::sort(v.begin(), v.end(), [](const hom &a, const hom &b) { return a.starts_left_of(b); });
}

/* ---------------------------------------------------------------- comparison */

struct counts {
	uint64_t subst = 0, homologs = 0;
};

/* libs/seqcmp.c:13-28 */
uint64_t seqcmp_plain(const char *a, const char *b, uint64_t len)
{
	uint64_t d = 0;
	for (uint64_t k = 0; k < len; k++)
		d += a[k] != b[k];
	return d;
}

/* libs/revseqcmp.c:15-30 with is_complement of libs/revseqcmp.h:19-23 */
uint64_t revseqcmp_plain(const char *a, const char *b, uint64_t len)
{
	uint64_t d = 0;
	for (uint64_t k = 0; k < len; k++)
		d += (((a[k] ^ b[len - 1 - k]) & 6) != 4);
	return d;
}

/* src/process.cxx:620-658 with src/evo_model.cxx:53-75 */
void compare_one(const char *qa, const hom &ha, const char *qb, const hom &hb, counts &c)
{
	if (!ha.overlaps(hb)) return;
	i64 cs = std::max(ha.start(), hb.start());
	i64 ce = std::min(ha.end(), hb.end());
	i64 len = ce - cs;
	hom ta = ha.trim(cs, ce), tb = hb.trim(cs, ce);
	uint64_t mm;
	if (ha.dir == hb.dir)
		mm = seqcmp_plain(qa + ta.iq, qb + tb.iq, (uint64_t)len);
	else if (hb.dir == 1)
		mm = revseqcmp_plain(qa + ta.iq, qb + (tb.iq + tb.len) - len, (uint64_t)len);
	else
		mm = revseqcmp_plain(qb + tb.iq, qa + (ta.iq + ta.len) - len, (uint64_t)len);
	c.homologs += (uint64_t)len;
	c.subst += mm;
}

/* src/process.cxx:566-611 — sweep over two sorted, internally disjoint lists */
counts compare_lists(const char *qa, const std::vector<hom> &ha, const char *qb,
                     const std::vector<hom> &hb)
{
	counts c;
	size_t right = 0;
	std::vector<hom> pile;
	for (const hom &h : ha) {
		pile.erase(std::remove_if(pile.begin(), pile.end(),
		                          [&](const hom &o) { return o.ends_left_of(h); }),
		           pile.end());
		while (right < hb.size() && hb[right].ends_left_of(h))
			right++;
		size_t far = right;
		while (far < hb.size() && hb[far].overlaps(h))
			far++;
		pile.insert(pile.end(), hb.begin() + (long)right, hb.begin() + (long)far);
		right = far;
		for (const hom &o : pile)
			compare_one(qa, h, qb, o, c);
	}
	return c;
}

/* src/process.cxx:725-776 */
std::vector<std::vector<hom>> complete_delete(const std::vector<std::vector<hom>> &in)
{
	const size_t N = in.size();
	std::vector<std::vector<hom>> out(N);
	std::vector<size_t> at(N, 0);
	auto all_left = [&]() {
		for (size_t g = 0; g < N; g++)
			if (at[g] >= in[g].size()) return false;
		return true;
	};
	while (all_left()) {
		i64 cs = in[0][at[0]].start(), ce = in[0][at[0]].end();
		size_t leftmost = 0;
		for (size_t g = 1; g < N; g++) {
			cs = std::max(cs, in[g][at[g]].start());
			if (in[g][at[g]].end() < ce) { /* first minimum, like std::min_element */
				ce = in[g][at[g]].end();
				leftmost = g;
			}
		}
		if (cs < ce)
			for (size_t g = 0; g < N; g++)
				out[g].push_back(in[g][at[g]].trim(cs, ce));
		at[leftmost]++;
	}
	return out;
}

/* src/evo_model.cxx:100-131 */
double estimate(uint64_t subst, uint64_t homologs, int kind)
{
	if (homologs == 0) return NAN; /* NaN survives the JC formula and the `<= 0` fix-up */
	double raw = subst / (double)homologs;
	if (kind == 0) return raw;
	if (kind == 2) return (1.0 - raw) * 100;
	double d = -0.75 * log(1.0 - (4.0 / 3.0) * raw);
	return d <= 0.0 ? 0.0 : d;
}

po_hom to_po(const hom &h)
{
	return po_hom{h.dir, h.iref, h.iproj, h.iq, h.len};
}

hom from_po(const po_hom &p)
{
	hom h;
	h.dir = p.direction;
	h.iref = p.index_reference;
	h.iproj = p.index_reference_projected;
	h.iq = p.index_query;
	h.len = p.length;
	return h;
}

} // namespace

extern "C" {

const char *po_kind(void)
{
	return "port";
}

void po_revcomp(const char *in, int64_t n, char *out)
{
	auto r = revcomp(in, n);
	std::memcpy(out, r.data(), r.size());
}

/* src/sequence.cxx:109-146 — keep ACGT in either case, upper-cased */
int64_t po_filter_nucl(const char *in, int64_t n, char *out)
{
	int64_t w = 0;
	for (int64_t k = 0; k < n; k++) {
		char c = in[k];
		if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
		if (c == 'A' || c == 'C' || c == 'G' || c == 'T') out[w++] = c;
	}
	return w;
}

/* src/sequence.cxx:152-165 — counts bytes with both bits of ('G' & 'C') = 0x43 set */
double po_gc_content(const char *seq, int64_t n)
{
	size_t gc = 0;
	for (int64_t k = 0; k < n; k++)
		gc += ((seq[k] & 'G' & 'C') == ('G' & 'C'));
	return (double)gc / (size_t)n;
}

uint64_t po_seqcmp(const char *a, const char *b, uint64_t len)
{
	return seqcmp_plain(a, b, len);
}

uint64_t po_revseqcmp(const char *a, const char *b, uint64_t len)
{
	return revseqcmp_plain(a, b, len);
}

/* src/process.cxx:77-86 */
int64_t po_min_anchor_length(double p, double gc, int64_t l)
{
	size_t x = 1;
	while (shuprop(x, gc / 2, (size_t)l) < 1 - p)
		x++;
	return (int64_t)x;
}

void *po_esa_create(const char *ref, int64_t n)
{
	return build_esa(ref, n);
}

void po_esa_destroy(void *e)
{
	delete static_cast<esa_t *>(e);
}

int64_t po_esa_size(void *e)
{
	return static_cast<esa_t *>(e)->m;
}

void po_esa_arrays(void *ev, int64_t *SA, int64_t *LCP, int64_t *CLD, char *FVC, char *S)
{
	auto *e = static_cast<esa_t *>(ev);
	size_t m = (size_t)e->m;
	if (SA) std::copy(e->SA.begin(), e->SA.end(), SA);
	if (LCP) std::copy(e->LCP.begin(), e->LCP.end(), LCP);
	if (CLD) std::copy(e->CLD.begin(), e->CLD.end(), CLD);
	if (FVC) std::memcpy(FVC, e->FVC.data(), m);
	if (S) std::memcpy(S, e->S.data(), m);
}

void po_get_match(void *ev, const char *query, int64_t qlen, int cached, int64_t out[3])
{
	auto *e = static_cast<esa_t *>(ev);
	std::string q(query, (size_t)qlen);
	interval r = cached ? get_match_cached(*e, q.c_str(), (size_t)qlen)
	                    : get_match(*e, q.c_str(), (size_t)qlen);
	out[0] = r.l;
	out[1] = r.i;
	out[2] = r.j;
}

int64_t po_anchor_homologies(void *ev, int64_t threshold, const char *query, int64_t qlen,
                             po_hom *out, int64_t cap)
{
	auto *e = static_cast<esa_t *>(ev);
	std::string q(query, (size_t)qlen);
	auto hv = anchor_homologies(*e, threshold, q.c_str(), qlen);
	int64_t k = 0;
	for (const auto &h : hv) {
		if (k < cap) out[k] = to_po(h);
		k++;
	}
	return k;
}

int64_t po_sort_filter(po_hom *h, int64_t count, int do_sort)
{
	std::vector<hom> pile;
	for (int64_t k = 0; k < count; k++)
		pile.push_back(from_po(h[k]));
	if (do_sort) sort_by_start(pile);
	filter_overlaps_max(pile);
	for (size_t k = 0; k < pile.size(); k++)
		h[k] = to_po(pile[k]);
	return (int64_t)pile.size();
}

void po_compare(const char *qa, const po_hom *ha, int64_t na, const char *qb, const po_hom *hb,
                int64_t nb, uint64_t out[2])
{
	std::vector<hom> va, vb;
	for (int64_t k = 0; k < na; k++)
		va.push_back(from_po(ha[k]));
	for (int64_t k = 0; k < nb; k++)
		vb.push_back(from_po(hb[k]));
	counts c = compare_lists(qa, va, qb, vb);
	out[0] = c.subst;
	out[1] = c.homologs;
}

int64_t po_complete_delete(const po_hom *h, const int64_t *offs, int64_t N, po_hom *out,
                           int64_t *out_offs, int64_t cap)
{
	std::vector<std::vector<hom>> in((size_t)N);
	for (int64_t g = 0; g < N; g++)
		for (int64_t k = offs[g]; k < offs[g + 1]; k++)
			in[(size_t)g].push_back(from_po(h[k]));
	auto core = complete_delete(in);
	int64_t w = 0;
	for (int64_t g = 0; g < N; g++) {
		out_offs[g] = w;
		for (const auto &x : core[(size_t)g]) {
			if (w < cap) out[w] = to_po(x);
			w++;
		}
	}
	out_offs[N] = w;
	return w;
}

/* src/process.cxx:408-556 */
int po_process(const char *const *seqs, const int64_t *lens, int64_t N, int64_t ref_index,
               int flags, int threads, uint64_t *subst, uint64_t *homologs, double *timings,
               int64_t *hom_counts)
{
	if (threads < 1) threads = 1;
	/* the reference hands NUL-terminated std::string storage to every routine */
	std::vector<std::string> q((size_t)N);
	for (int64_t g = 0; g < N; g++)
		q[(size_t)g].assign(seqs[g], (size_t)lens[g]);

	po_sa_seconds = 0;
	double t0 = now();
	esa_t *e = build_esa(q[(size_t)ref_index].c_str(), lens[ref_index]);
	double t1 = now();
	double gc = po_gc_content(q[(size_t)ref_index].c_str(), lens[ref_index]);
	int64_t thr = po_min_anchor_length(0.025, gc, e->m); /* ANCHOR_P_VALUE, src/phylonium.cxx:55 */

	std::vector<std::vector<hom>> H((size_t)N);
#pragma omp parallel for num_threads(threads)
	for (int64_t g = 0; g < N; g++) {
		auto hv = anchor_homologies(*e, thr, q[(size_t)g].c_str(), lens[g]);
		sort_by_start(hv);
		filter_overlaps_max(hv);
		H[(size_t)g] = std::move(hv);
	}
	double t2 = now();
	if (flags & 4) H = complete_delete(H);
	if (hom_counts)
		for (int64_t g = 0; g < N; g++)
			hom_counts[g] = (int64_t)H[(size_t)g].size();
	for (int64_t k = 0; k < N * N; k++)
		subst[k] = homologs[k] = 0;
	double t3 = now();
#pragma omp parallel for num_threads(threads)
	for (int64_t i = 0; i < N; i++) {
		for (int64_t j = i + 1; j < N; j++) {
			counts c = compare_lists(q[(size_t)i].c_str(), H[(size_t)i], q[(size_t)j].c_str(),
			                         H[(size_t)j]);
			subst[i * N + j] = subst[j * N + i] = c.subst;
			homologs[i * N + j] = homologs[j * N + i] = c.homologs;
		}
	}
	double t4 = now();
	if (timings) {
		timings[0] = t1 - t0;
		timings[1] = t2 - t1;
		timings[2] = t4 - t3;
		timings[3] = po_sa_seconds;
	}
	delete e;
	return 0;
}

/* process() for a sample of the matrix: all sequences mapped (src/process.cxx:433-458), only
 * the listed rows compared against every sequence (:524-549 for those i) */
int po_process_rows(const char *const *seqs, const int64_t *lens, int64_t N, int64_t ref_index, int flags,
                    int threads, const int64_t *rows, int64_t nrows, uint64_t *subst, uint64_t *homologs,
                    double *timings)
{
	if (threads < 1) threads = 1;
	std::vector<std::string> q((size_t)N);
	for (int64_t g = 0; g < N; g++)
		q[(size_t)g].assign(seqs[g], (size_t)lens[g]);
	po_sa_seconds = 0;
	double t0 = now();
	esa_t *e = build_esa(q[(size_t)ref_index].c_str(), lens[ref_index]);
	double t1 = now();
	double gc = po_gc_content(q[(size_t)ref_index].c_str(), lens[ref_index]);
	int64_t thr = po_min_anchor_length(0.025, gc, e->m);
	std::vector<std::vector<hom>> H((size_t)N);
#pragma omp parallel for num_threads(threads) schedule(dynamic)
	for (int64_t g = 0; g < N; g++) {
		auto hv = anchor_homologies(*e, thr, q[(size_t)g].c_str(), lens[g]);
		sort_by_start(hv);
		filter_overlaps_max(hv);
		H[(size_t)g] = std::move(hv);
	}
	double t2 = now();
	if (flags & 4) H = complete_delete(H);
	double t3 = now();
#pragma omp parallel for num_threads(threads) schedule(dynamic) collapse(2)
	for (int64_t r = 0; r < nrows; r++) {
		for (int64_t j = 0; j < N; j++) {
			const int64_t i = rows[r];
			counts c{0, 0};
			if (j != i) c = compare_lists(q[(size_t)i].c_str(), H[(size_t)i], q[(size_t)j].c_str(), H[(size_t)j]);
			subst[r * N + j] = c.subst;
			homologs[r * N + j] = c.homologs;
		}
	}
	double t4 = now();
	if (timings) {
		timings[0] = t1 - t0;
		timings[1] = t2 - t1;
		timings[2] = t4 - t3;
		timings[3] = po_sa_seconds;
	}
	delete e;
	return 0;
}

double po_estimate(uint64_t subst, uint64_t homologs, int kind)
{
	return estimate(subst, homologs, kind);
}

/* src/io.cxx:141-163 — PHYLIP text: N, then name and "  "-separated cells,
 * scientific with 4 digits (default float format for ANI), diagonal forced to 0 */
int64_t po_format_matrix(const char *const *names, const uint64_t *subst,
                         const uint64_t *homologs, int64_t N, int kind, char *out, int64_t cap)
{
	std::string s = std::to_string(N) + "\n";
	char buf[64];
	for (int64_t i = 0; i < N; i++) {
		s += names[i];
		for (int64_t j = 0; j < N; j++) {
			double d = i == j ? 0.0 : estimate(subst[i * N + j], homologs[i * N + j], kind);
			if (kind == 2)
				snprintf(buf, sizeof buf, "  %.4g", d);
			else
				snprintf(buf, sizeof buf, "  %.4e", d);
			s += buf;
		}
		s += "\n";
	}
	if ((int64_t)s.size() < cap) std::memcpy(out, s.data(), s.size());
	return (int64_t)s.size();
}

/* test/simf.cxx:93-140 — base sequence from default_random_engine{base_seed},
 * substitutions drawn from a second engine seeded mut_seed; the remaining-mutation
 * bookkeeping makes the number of substitutions hit length*p almost exactly. */
void po_simf(uint32_t base_seed, uint32_t mut_seed, int64_t length, double divergence, int raw,
             char *out)
{
	double p = raw ? divergence : 0.75 - 0.75 * exp(-(4.0 / 3.0) * divergence);
	std::default_random_engine base_rand{base_seed};
	std::uniform_int_distribution<int> base_dist{0, 3};
	/* simf.cxx:108 binds the engine BY VALUE: the "mutate here?" draws come from a
	 * copy, the "which base?" draws from the original — two streams, same seed. */
	std::default_random_engine mut_rand_where{mut_seed};
	std::default_random_engine mut_rand_which{mut_seed};
	std::uniform_real_distribution<double> mut_dist{0, 1};
	std::uniform_int_distribution<int> mut_pick{0, 2};
	static const char *others[4] = {"CGT", "AGT", "ACT", "ACG"};
	double nucleotides = (double)length;
	double mutations = nucleotides * p;
	for (int64_t k = 0; k < length; k++) {
		int b = base_dist(base_rand);
		char c = "ACGT"[b];
		if (mut_dist(mut_rand_where) < mutations / nucleotides) {
			c = others[b][mut_pick(mut_rand_which)];
			mutations--;
		}
		out[k] = c;
		nucleotides--;
	}
}

} // extern "C"
