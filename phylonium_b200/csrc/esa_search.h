// Longest-match search on the device-resident ESA: the interval descent of
// /root/reference/src/esa.cxx:361-563 restated for int32 arrays, usable from CUDA
// kernels and (for the CPU emulation tests) from plain C++.
//
// Differences from the reference that do not change results:
//  * the descent reads one 16-byte record per suffix-array index (EsaNode = SA, LCP, CLD,
//    FVC side by side) instead of four arrays: every step of the sibling walk is then ONE
//    dependent load instead of three — the descent is bound by memory latency, not bandwidth;
//  * the 6-mer interval cache (src/esa.cxx:90-228) is replaced by a K-mer table,
//    K chosen from the text length (esa_build.cu), holding for every K-mer the
//    deepest descent state that needs no character beyond the K-mer. SURVEY.md A.4:
//    the cache is a pure accelerator; only (l, i == j, SA[i]) are consumed;
//  * the byte-by-byte extension of a singleton interval can be capped so that very long
//    matches are finished cooperatively (anchor.cu); the Match then has open != 0.
#pragma once
#include "esa_types.h"

namespace phy
{

PHY_HD EsaNode esa_node(const EsaView &e, int32_t idx)
{
#if defined(__CUDA_ARCH__)
	const int4 v = __ldg(reinterpret_cast<const int4 *>(e.node) + idx);
	return EsaNode{v.x, v.y, v.z, v.w};
#else
	return e.node[idx];
#endif
}

// LCP[nx.cld] for the record nx of some index x: from the hint, or by looking it up
PHY_HD int32_t esa_child_lcp(const EsaView &e, const EsaNode &nx)
{
	const uint32_t hint = (uint32_t)nx.fvc >> 8;
	return hint != ESA_HINT_NONE ? (int32_t)hint - 1 : esa_node(e, nx.cld).lcp;
}

PHY_HD Interval esa_root(const EsaView &e)
{
	// src/esa.cxx:527-528: m = left_child(m_size) = CLD[m_size - 1]
	const EsaNode last = esa_node(e, e.m - 1);
	return Interval{esa_child_lcp(e, last), 0, e.m - 1, last.cld};
}

PHY_HD bool interval_empty(const Interval &ij)
{
	return ij.i == -1 && ij.j == -1;
}

// src/esa.cxx:361-427 — child interval of ij (i < j) whose suffixes continue with character a,
// given the records of i (ni), of the first l-index m (nm) and of m - 1 (np) and the text byte
// c0 = S[SA[i] + l].  sa receives SA[result.i] (undefined for an empty result).
PHY_HD Interval esa_get_interval_core(const EsaView &e, Interval ij, uint8_t a, int32_t &sa, EsaNode ni, EsaNode nm,
                                      EsaNode np, uint8_t c0)
{
	int32_t i = ij.i;
	const int32_t j = ij.j;
	sa = ni.sa;
	int32_t m = ij.m;
	const int32_t l = ij.l;
	uint8_t c = c0;
	for (;;) {
		if (c == a) {
			if (i != m - 1) return Interval{esa_child_lcp(e, np), i, m - 1, np.cld}; // left_child(m)
			return Interval{ni.lcp, i, i, -1};
		}
		if (c > a) break;
		i = m;
		ni = nm;
		sa = ni.sa;
		if (i == j) break;
		m = nm.cld; // right_child(m)
		nm = esa_node(e, m);
		np = esa_node(e, m - 1);
		if (nm.lcp != l) break;
		c = (uint8_t)ni.fvc;
	}
	const bool hit = (i != ij.i) ? ((uint8_t)ni.fvc == a) : (c0 == a);
	if (!hit) {
		ij.i = ij.j = -1;
		return ij;
	}
	ij.i = i;
	ij.l = nm.lcp; // LCP[m]; for i == j == m this is LCP[i] like in the reference
	ij.m = m;
	return ij;
}

// src/esa.cxx:361-427 — child interval of ij whose suffixes continue with character a.
// sa receives SA[result.i] (undefined for an empty result).
PHY_HD Interval esa_get_interval(const EsaView &e, Interval ij, uint8_t a, int32_t &sa)
{
	const EsaNode ni = esa_node(e, ij.i);
	sa = ni.sa;
	if (ij.i == ij.j) {
		if (e.S[ni.sa + ij.l] != a) ij.i = ij.j = -1;
		return ij;
	}
	// the record of the next l-index m and, speculatively, of m - 1 (whose CLD and hint give the
	// child [i, m - 1] if that is the one we are after): both are independent of the text load
	// below and of each other, so a step down costs two round trips to memory instead of four
	const EsaNode nm = esa_node(e, ij.m);
	const EsaNode np = esa_node(e, ij.m - 1);
	const uint8_t c0 = e.S[ni.sa + ij.l];
	return esa_get_interval_core(e, ij, a, sa, ni, nm, np, c0);
}

// Lane groups.  A walker of anchor.cu is run by COOP consecutive lanes of a warp (32, 16 or
// 8; 0 = one plain thread, as in the CPU emulation): all of them execute the walker's scalar
// logic redundantly and share the byte comparisons.  With fewer than 32 lanes per walker a
// warp carries several walkers, which may sit at different program counters; the group
// primitives below only ever name the lanes of one group.
#if defined(__CUDACC__)
template <int COOP> __device__ __forceinline__ int coop_lane()
{
	return threadIdx.x & (COOP - 1);
}
template <int COOP> __device__ __forceinline__ int coop_shift()
{
	return (threadIdx.x & 31) & ~(COOP - 1);
}
template <int COOP> __device__ __forceinline__ uint32_t coop_mask()
{
	return COOP == 32 ? 0xffffffffu : ((1u << (COOP & 31)) - 1u) << coop_shift<COOP>();
}
// ballot over the group, bit 0 = the group's first lane
template <int COOP> __device__ __forceinline__ uint32_t coop_ballot(bool p)
{
	return __ballot_sync(coop_mask<COOP>(), p) >> coop_shift<COOP>();
}
#endif

// First k in [from, to) with a[k] != b[k], or `to`.
// COOP > 0: called by all lanes of a group with identical arguments; the lanes then compare
// 32 bytes per step — 32 / COOP each — and agree on the result through a ballot.
// COOP = 0: plain loop.
template <int COOP> PHY_HD int32_t match_run(const uint8_t *a, const uint8_t *b, int32_t from, int32_t to)
{
#if defined(__CUDA_ARCH__)
	if (COOP) {
		constexpr int PER = 32 / (COOP ? COOP : 32); // bytes per lane and step
		const int gl = coop_lane<COOP ? COOP : 32>();
		// position of the first mismatch among the 32 bytes from `base` on, or -1; lanes past
		// `lim` (exclusive) do not count
		auto step = [&](int32_t base, int32_t lim, bool checked) -> int32_t {
			uint32_t miss = 0; // bit t: my byte t differs
#pragma unroll
			for (int t = 0; t < PER; t++) {
				const int32_t x = base + gl * PER + t;
				if (!checked || x < lim) miss |= (uint32_t)(a[x] != b[x]) << t;
			}
			const uint32_t bal = coop_ballot<COOP ? COOP : 32>(miss != 0);
			if (!bal) return -1;
			const int first = __ffs(bal) - 1;
			if (PER == 1) return base + first;
			const uint32_t m = __shfl_sync(coop_mask<COOP ? COOP : 32>(), miss, coop_shift<COOP ? COOP : 32>() + first);
			return base + first * PER + (__ffs(m) - 1);
		};
		// first 32 bytes alone: most comparisons at high divergence end here
		{
			const int32_t r = step(from, to, true);
			if (r >= 0) return r;
		}
		// Whole steps first (no bounds checks: a third of the walk's instructions were spent
		// here), four of them in flight together — the loop is bound by the latency of a round
		// trip to L2/HBM, not by bandwidth — then checked steps for the rest.
		int32_t base = from + 32;
		if (COOP == 32) {
			for (; base + 128 <= to; base += 128) {
				uint8_t va[4], vb[4];
#pragma unroll
				for (int u = 0; u < 4; u++) {
					va[u] = a[base + 32 * u + gl];
					vb[u] = b[base + 32 * u + gl];
				}
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const uint32_t bal = __ballot_sync(0xffffffffu, va[u] != vb[u]);
					if (bal) return base + 32 * u + (__ffs(bal) - 1);
				}
			}
		} else {
			for (; base + 32 <= to; base += 32) {
				const int32_t r = step(base, to, false);
				if (r >= 0) return r;
			}
		}
		for (; base < to; base += 32) {
			const int32_t r = step(base, to, true);
			if (r >= 0) return r;
		}
		return to;
	}
#endif
	int32_t k = from;
	while (k < to && a[k] == b[k])
		k++;
	return k;
}

// Extension of a singleton: characters [k, …) of the query against S[sa + k …).
// Stops at the first mismatch (the NUL after S counts as one), at qlen, or — still
// matching — once `cap` characters are verified (open).
template <int COOP = 0>
PHY_HD Match esa_extend_singleton(const EsaView &e, const uint8_t *q, int32_t qlen, int32_t k, int32_t idx, int32_t sa,
                                  int32_t cap)
{
	const uint8_t *s = e.S + sa;
	const int32_t lim = qlen < cap ? qlen : cap;
	if (k < lim) k = match_run<COOP>(s, q, k, lim);
	Match r;
	r.l = k;
	r.i = r.j = idx;
	r.open = (k >= lim && lim < qlen) ? 1 : 0;
	r.sa = sa;
	return r;
}

// src/esa.cxx:446-513 — continue a match of q[0..k) that sits in the proper interval ij
// (i < j, k == ij.l).
// pre: a table record whose ij is this very interval; its ni / nm / np / text byte serve the
// first step (no loads), later steps fetch their own
template <int COOP = 0>
PHY_HD Match esa_match_from(const EsaView &e, const uint8_t *q, int32_t qlen, int32_t k, Interval ij, int32_t cap,
                            const TableRec *pre = nullptr)
{
	Match res;
	res.i = ij.i;
	res.j = ij.j;
	res.open = 0;
	res.sa = -1;
	do {
		int32_t sa;
		if (pre) {
			ij = esa_get_interval_core(e, ij, q[k], sa, pre->ni, pre->nm, pre->np, (uint8_t)pre->np.sa);
			pre = nullptr;
		} else {
			ij = esa_get_interval(e, ij, q[k], sa);
		}
		if (interval_empty(ij)) {
			res.l = k;
			return res;
		}
		res.i = ij.i;
		res.j = ij.j;
		k++; // by definition the k-th letter matched
		if (ij.i == ij.j) return esa_extend_singleton<COOP>(e, q, qlen, k, ij.i, sa, cap);
		const int32_t l = ij.l < qlen ? ij.l : qlen;
		if (k < l) {
			k = match_run<COOP>(e.S + sa, q, k, l);
			if (k < l) {
				res.l = k;
				return res;
			}
		}
	} while (k < qlen);
	res.l = qlen;
	return res;
}

// src/esa.cxx:525-531
template <int COOP = 0> PHY_HD Match esa_match_root(const EsaView &e, const uint8_t *q, int32_t qlen, int32_t cap)
{
	return esa_match_from<COOP>(e, q, qlen, 0, esa_root(e), cap);
}

// src/esa.cxx:542-563 with the K-mer table in the role of the cache.
// Table records: i == j: singleton with l verified characters and m = SA[i];
//                i <  j: interval with lcp value l (min(K, l) characters verified).
template <int COOP = 0> PHY_HD Match esa_match(const EsaView &e, const uint8_t *q, int32_t qlen, int32_t cap)
{
	const int32_t K = e.K;
	if (K <= 0 || qlen <= K) return esa_match_root<COOP>(e, q, qlen, cap);
	uint32_t code = 0;
#if defined(__CUDA_ARCH__)
	if (COOP) {
		// the group's lanes read the K <= 12 characters (lane t: t, t + COOP, …); the 2-bit
		// codes are OR-ed together across the group
		constexpr int C = COOP ? COOP : 32;
		const int gl = coop_lane<C>();
		uint32_t part = 0;
		bool bad = false;
#pragma unroll
		for (int t0 = 0; t0 < 12; t0 += C) {
			const int t = t0 + gl;
			if (t < K) {
				const int c = kmer_code(q[t]);
				bad = bad || c < 0;
				part |= (uint32_t)(c & 3) << (2 * (K - 1 - t));
			}
		}
		if (__any_sync(coop_mask<C>(), bad)) return esa_match_root<COOP>(e, q, qlen, cap);
		code = __reduce_or_sync(coop_mask<C>(), part);
	} else
#endif
	{
		for (int32_t t = 0; t < K; t++) {
			int c = kmer_code(q[t]);
			if (c < 0) return esa_match_root<COOP>(e, q, qlen, cap);
			code = (code << 2) | (uint32_t)c;
		}
	}
	TableRec rec;
#if defined(__CUDA_ARCH__)
	{
		const int4 *src = reinterpret_cast<const int4 *>(e.table + code);
		const int4 t0 = __ldg(src), t1 = __ldg(src + 1), t2 = __ldg(src + 2), t3 = __ldg(src + 3);
		rec.ij = Interval{t0.x, t0.y, t0.z, t0.w};
		rec.ni = EsaNode{t1.x, t1.y, t1.z, t1.w};
		rec.nm = EsaNode{t2.x, t2.y, t2.z, t2.w};
		rec.np = EsaNode{t3.x, t3.y, t3.z, t3.w};
	}
#else
	rec = e.table[code];
#endif
	const Interval ij = rec.ij;
	if (ij.i == ij.j) return esa_extend_singleton<COOP>(e, q, qlen, ij.l, ij.i, ij.m, cap);
	int32_t k = ij.l;
	if (k > K) {
		// the table verified K characters of this deep interval; finish its label
		const int32_t l = ij.l < qlen ? ij.l : qlen;
		k = match_run<COOP>(e.S + rec.ni.sa, q, K, l);
		if (k < l) return Match{k, ij.i, ij.j, 0, -1};
		if (k >= qlen) return Match{qlen, ij.i, ij.j, 0, -1};
	}
	return esa_match_from<COOP>(e, q, qlen, k, ij, cap, &rec);
}

// The table record of an interval: the interval and what the first step down from it needs.
PHY_HD TableRec esa_table_record(const EsaView &e, const Interval &ij)
{
	TableRec r;
	r.ij = ij;
	r.ni = r.nm = r.np = EsaNode{0, 0, 0, 0};
	if (ij.i != ij.j && ij.i >= 0) {
		r.ni = esa_node(e, ij.i);
		r.nm = esa_node(e, ij.m);
		r.np = esa_node(e, ij.m - 1);
		r.np.sa = (int32_t)e.S[r.ni.sa + ij.l];
	}
	return r;
}

// One record of the K-mer table: descend on the K characters of `code`, stop before a
// character beyond the K-mer would be needed. See the header comment.
PHY_HD Interval esa_table_entry(const EsaView &e, uint32_t code, int32_t K)
{
	uint8_t w[16];
	for (int32_t t = 0; t < K; t++)
		w[t] = (uint8_t)(0x54474341u >> (8 * ((code >> (2 * (K - 1 - t))) & 3))); // "ACGT"
	Interval ij = esa_root(e);
	while (ij.l < K) {
		int32_t sa;
		Interval nx = esa_get_interval(e, ij, w[ij.l], sa);
		if (interval_empty(nx)) break; // resuming repeats the failing step
		const uint8_t *s = e.S + sa;
		int32_t k = ij.l + 1;
		if (nx.i == nx.j) {
			while (k < K && s[k] == w[k])
				k++;
			return Interval{k, nx.i, nx.i, sa};
		}
		const int32_t upto = nx.l < K ? nx.l : K;
		while (k < upto && s[k] == w[k])
			k++;
		if (k < upto) break; // mismatch inside the edge label: resume from the parent
		ij = nx;             // min(K, nx.l) characters verified
	}
	return ij;
}

// ---- the same table built level by level -------------------------------------------------
//
// esa_table_entry walks K characters down from the root for every K-mer.  The entry of a
// (k+1)-mer is a function of the entry of its k-mer prefix and one more character, so the
// table can be grown one level at a time, one descent step per entry instead of k + 1:
// 4/3 * 4^K steps altogether instead of K * 4^K (esa_build.cu).  What the step needs besides
// the prefix's entry is the node that entry was reached from: if the entry sits inside an
// edge label longer than k and the next label character differs from c, the descent falls
// back to that parent (esa_table_entry: "mismatch inside the edge label").
struct TableBuild {
	Interval cur; // esa_table_entry(code, k)
	Interval par; // the explicit node cur was entered from (used while cur.l > k)
};

PHY_HD TableBuild esa_table_root(const EsaView &e)
{
	TableBuild t;
	t.cur = esa_root(e);
	t.par = t.cur;
	return t;
}

// entry of the (k+1)-mer "prefix c" from the entry P of its k-mer prefix
PHY_HD TableBuild esa_table_extend(const EsaView &e, const TableBuild &P, int32_t k, uint8_t c)
{
	TableBuild r = P;
	const Interval cur = P.cur;
	if (cur.i == cur.j) { // singleton {verified, i, i, SA[i]}
		if (cur.l == k && e.S[cur.m + k] == c) r.cur.l = k + 1;
		return r;
	}
	if (cur.l < k) return r; // the descent failed at an earlier character: same state
	if (cur.l > k) {         // inside the label of cur: one more label character
		if (e.S[esa_node(e, cur.i).sa + k] != c) r.cur = P.par;
		return r;
	}
	// cur is an explicit node at depth k: one descent step
	int32_t sa;
	const Interval nx = esa_get_interval(e, cur, c, sa);
	if (interval_empty(nx)) return r; // resuming repeats the failing step
	r.par = cur;
	if (nx.i == nx.j)
		r.cur = Interval{k + 1, nx.i, nx.i, sa};
	else
		r.cur = nx; // nx.l >= k + 1 characters, k + 1 of them verified
	return r;
}

} // namespace phy
