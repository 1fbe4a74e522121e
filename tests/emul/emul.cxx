// TEST INFRASTRUCTURE — CPU emulation of the device-side logic.
//
// The headers under phylonium_b200/csrc that hold the algorithmic core (esa_search.h,
// walk.h, cld_search.h, filter.h) compile for host and device.  This file drives them
// sequentially on the CPU so that the speculative walk, the stack-free child table, the
// K-mer table and the chaining filter can be checked against the oracle WITHOUT a GPU
// (pytest -m "not gpu").  It is never loaded by the product: libphylonium_b200.so has no
// CPU path.  The orchestration here (plain loops instead of kernels, a sequential chain
// walk instead of pointer doubling) mirrors anchor.cu step by step.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../phylonium_b200/csrc/cld_search.h"
#include "../../phylonium_b200/csrc/esa_search.h"
#include "../../phylonium_b200/csrc/filter.h"
#include "../../phylonium_b200/csrc/tile_order.h"
#include "../../phylonium_b200/csrc/walk.h"

using namespace phy;

namespace
{

struct HostEsa {
	std::vector<uint8_t> S, FVC;
	std::vector<int32_t> SA, LCP, CLD;
	std::vector<TableRec> table;
	std::vector<EsaNode> node;
	EsaView view;
};

HostEsa make_esa(const uint8_t *S, const int64_t *SA, const int64_t *LCP, const int64_t *CLD, const uint8_t *FVC, int32_t m,
                 int32_t K)
{
	HostEsa e;
	e.S.assign(S, S + m);
	e.S.resize((size_t)m + 64, 0);
	e.FVC.assign(FVC, FVC + m);
	e.SA.resize(m);
	e.LCP.resize((size_t)m + 1);
	e.CLD.resize((size_t)m + 1);
	for (int32_t i = 0; i < m; i++)
		e.SA[i] = (int32_t)SA[i];
	for (int32_t i = 0; i <= m; i++) {
		e.LCP[i] = (int32_t)LCP[i];
		e.CLD[i] = (int32_t)CLD[i];
	}
	e.node.resize((size_t)m + 1);
	for (int32_t i = 0; i <= m; i++)
		e.node[i] = EsaNode{i < m ? e.SA[i] : 0, e.LCP[i], e.CLD[i],
		                    esa_pack_fvc(i < m ? e.FVC[i] : 0, e.LCP[e.CLD[i]])};
	e.view.node = e.node.data();
	e.view.S = e.S.data();
	e.view.SA = e.SA.data();
	e.view.LCP = e.LCP.data();
	e.view.CLD = e.CLD.data();
	e.view.FVC = e.FVC.data();
	e.view.table = nullptr;
	e.view.K = 0;
	e.view.m = m;
	e.view.n = (m - 1) / 2;
	if (K > 0) {
		e.table.resize((size_t)1 << (2 * K));
		for (uint32_t code = 0; code < e.table.size(); code++)
			e.table[code] = esa_table_record(e.view, esa_table_entry(e.view, code, K));
		// the level-by-level construction of esa_build.cu must give the very same records
		std::vector<TableBuild> level(1, esa_table_root(e.view));
		for (int32_t k = 0; k < K; k++) {
			std::vector<TableBuild> next(level.size() * 4);
			for (size_t code = 0; code < next.size(); code++)
				next[code] = esa_table_extend(e.view, level[code >> 2], k, (uint8_t) "ACGT"[code & 3]);
			level.swap(next);
		}
		for (uint32_t code = 0; code < e.table.size(); code++) {
			const Interval a = e.table[code].ij, b = level[code].cur;
			if (a.l != b.l || a.i != b.i || a.j != b.j || a.m != b.m) {
				fprintf(stderr, "emul: hierarchical table differs at K=%d code=%u: (%d %d %d %d) vs (%d %d %d %d)\n", K, code,
				        a.l, a.i, a.j, a.m, b.l, b.i, b.j, b.m);
				abort();
			}
		}
		e.view.table = e.table.data();
		e.view.K = K;
	}
	return e;
}

int32_t scalar_first_mismatch(const uint8_t *q, const uint8_t *S, int64_t diag, int32_t m, int32_t from, int32_t to)
{
	for (int32_t x = from; x < to; x++) {
		const int64_t sp = (int64_t)x + diag;
		if (sp >= m || q[x] != S[sp]) return x;
	}
	return to;
}

} // namespace

extern "C" {

// order of the tile pairs of the all-pairs stage (tile_order.h): pair p of a launch over the
// tile columns [tile_begin, tile_end)
void emul_unrank_pair(int64_t p, int32_t tile_begin, int32_t tile_end, int32_t *ti, int32_t *tj)
{
	phy::cmp_unrank_pair(p, tile_begin, tile_end, *ti, *tj);
}


// CLD from LCP through the min-pyramid closed form (cld_search.h)
void emul_cld(const int64_t *LCP64, int32_t m, int64_t *CLD_out)
{
	std::vector<std::vector<int32_t>> lv;
	lv.emplace_back((size_t)m + 1);
	for (int32_t i = 0; i <= m; i++)
		lv[0][i] = (int32_t)LCP64[i];
	while (lv.back().size() > 1) {
		const auto &in = lv.back();
		std::vector<int32_t> out((in.size() + 31) / 32);
		for (size_t o = 0; o < out.size(); o++) {
			int32_t mn = 0x7fffffff;
			for (size_t t = o * 32; t < std::min(in.size(), o * 32 + 32); t++)
				mn = std::min(mn, in[t]);
			out[o] = mn;
		}
		lv.push_back(std::move(out));
	}
	Pyramid py;
	py.levels = (int32_t)lv.size();
	for (int k = 0; k < py.levels; k++) {
		py.level[k] = lv[k].data();
		py.size[k] = (int32_t)lv[k].size();
	}
	for (int32_t i = 0; i < m; i++)
		CLD_out[i] = cld_entry(py, i);
	CLD_out[m] = 0;
}

// longest matches through the K-mer table (use_table) or from the root
void emul_matches(const uint8_t *S, const int64_t *SA, const int64_t *LCP, const int64_t *CLD, const uint8_t *FVC,
                  int32_t m, int32_t K, const uint8_t *text, const int64_t *offs, const int64_t *lens, int64_t count,
                  int use_table, int64_t *out)
{
	HostEsa e = make_esa(S, SA, LCP, CLD, FVC, m, K);
	for (int64_t k = 0; k < count; k++) {
		std::vector<uint8_t> q(text + offs[k], text + offs[k] + lens[k]);
		q.push_back(0);
		Match mt = use_table ? esa_match(e.view, q.data(), (int32_t)lens[k], 0x7fffffff)
		                     : esa_match_root(e.view, q.data(), (int32_t)lens[k], 0x7fffffff);
		out[3 * k] = mt.l;
		out[3 * k + 1] = mt.i;
		out[3 * k + 2] = mt.j;
	}
}

// The whole speculative walk for one query, phase by phase as in anchor.cu.
// out: raw homologies in push order as 5 int64 each {dir, iref, iproj, iq, len}.
// stats: {chunks, events, open events, bridges that merged, unresolved continuations}
int64_t emul_anchor(const uint8_t *S, const int64_t *SA, const int64_t *LCP, const int64_t *CLD, const uint8_t *FVC,
                    int32_t m, int32_t K, int32_t thr, int32_t CH, int32_t CAP, const uint8_t *query, int32_t qlen,
                    int64_t *out, int64_t cap, int64_t *stats)
{
	HostEsa e = make_esa(S, SA, LCP, CLD, FVC, m, K);
	CH = ((CH + 31) / 32) * 32;
	if (CH <= thr + 1) CH = ((thr + 2 + 31) / 32) * 32;
	if (CAP <= 0) CAP = 2 * CH;
	if (CAP < CH) CAP = CH;
	if (CAP < thr + 1) CAP = thr + 1;

	std::vector<uint8_t> Q(query, query + qlen);
	Q.resize((size_t)qlen + 64, 0);
	QueryInfo qi;
	qi.qoff = 0;
	qi.qlen = qlen;
	qi.chunk_base = 0;
	qi.nchunks = (qlen + CH - 1) / CH;
	qi.pad = 0;
	const int32_t nc = qi.nchunks;
	const int32_t cap_ev = CH / (thr + 1) + 2;
	std::vector<Event> ev((size_t)nc * cap_ev), bev((size_t)nc * cap_ev);
	std::vector<uint32_t> dead((size_t)nc * (CH / 32), 0);
	std::vector<ChunkRec> rec(nc);
	std::vector<int32_t> cq(nc, 0);

	WalkParams P;
	P.esa = e.view;
	P.Q = Q.data();
	P.qi = &qi;
	P.nq = 1;
	P.thr = thr;
	P.CH = CH;
	P.CAP = CAP;
	P.cap_ev = cap_ev;
	P.total_chunks = nc;
	P.ev = ev.data();
	P.bev = bev.data();
	P.dead = dead.data();
	P.rec = rec.data();
	P.chunk_query = cq.data();

	int64_t n_open = 0, n_merged = 0, n_unres = 0;
	// phase 1
	for (int32_t g = 0; g < nc; g++)
		walk_chunk(P, g);
	// phase 2
	std::vector<int32_t> lnk(nc, -1), endq(nc, -1);
	for (int32_t g = 0; g < nc; g++)
		if (rec[g].open) {
			n_open++;
			open_resolve_one(P, g, scalar_first_mismatch, lnk[g], endq[g]);
		}
	for (int32_t g = nc - 1; g >= 0; g--)
		if (rec[g].open && lnk[g] >= 0) {
			if (endq[lnk[g]] < 0) return -1; // links point forward, so the target is final
			endq[g] = endq[lnk[g]];
		}
	for (int32_t g = 0; g < nc; g++)
		if (rec[g].open) {
			if (endq[g] < 0) return -2;
			Event &x = ev[(size_t)g * cap_ev + rec[g].n_events - 1];
			x.len = endq[g] - x.pos;
			rec[g].exit.lastLen = x.len;
			rec[g].exit.pos = x.pos + x.len + 1;
			rec[g].open = 0;
		}
	// phase 3
	for (int32_t g = 0; g < nc; g++) {
		rec[g].link = bridge_walk(P, g, rec[g].exit, 0, rec[g].bridge_ev, cap_ev, CH, CAP);
		if (rec[g].link == LINK_MERGED) n_merged++;
	}
	// phase 4 + continuation
	std::vector<Event> path;
	std::vector<std::vector<Event>> overflow;
	int32_t cur = 0, from = 0;
	for (int guard = 0; nc > 0 && guard <= nc + 2; guard++) {
		ChunkRec &r = rec[cur];
		for (int32_t k = from; k < r.n_events; k++)
			path.push_back(ev[(size_t)cur * cap_ev + k]);
		if (r.link == LINK_UNRESOLVED) {
			n_unres++;
			overflow.emplace_back((size_t)qlen / (thr + 1) + 2);
			auto &o = overflow.back();
			for (int32_t k = 0; k < r.n_bridge; k++)
				o[k] = r.bridge_ev[k];
			r.bridge_ev = o.data();
			r.link = bridge_walk(P, cur, r.bstate, r.n_bridge, o.data(), (int64_t)o.size(), -1, 0x7fffffff);
			if (r.link == LINK_UNRESOLVED) return -3;
		}
		for (int32_t k = 0; k < r.n_bridge; k++)
			path.push_back(r.bridge_ev[k]);
		if (r.link == LINK_END) break;
		if (r.link != LINK_MERGED || r.link_chunk <= cur) return -4;
		from = r.link_from;
		cur = r.link_chunk;
	}
	// phase 5: run heads and homologies, formulated as in anchor.cu
	const int32_t border = e.view.n;
	int64_t w = 0;
	const int64_t ne = (int64_t)path.size();
	int64_t head = -1; // index of the run's first real event
	bool virt = false;
	for (int64_t t = 0; t < ne; t++) {
		const Event prev = t == 0 ? Event{0, 0, 0} : path[t - 1];
		const bool right = event_is_right(prev, path[t], border);
		if (!right) {
			head = t;
			virt = false;
		} else if (t == 0) {
			head = 0;
			virt = true;
		}
		const bool last = (t + 1 == ne) || !event_is_right(path[t], path[t + 1], border);
		if (!last) continue;
		Hom h;
		if (run_homology(path.data() + head, virt ? -1 : 0, (int32_t)(t - head), thr, border, h)) {
			if (w < cap) {
				out[5 * w + 0] = h.dir;
				out[5 * w + 1] = h.iref;
				out[5 * w + 2] = h.iproj;
				out[5 * w + 3] = h.iq;
				out[5 * w + 4] = h.len;
			}
			w++;
		}
	}
	if (stats) {
		stats[0] = nc;
		stats[1] = ne;
		stats[2] = n_open;
		stats[3] = n_merged;
		stats[4] = n_unres;
	}
	return w;
}

// filter_overlaps_max on a list already sorted by start; returns survivors' indices count
int32_t emul_filter(const int32_t *start, const int32_t *len, int32_t h, uint8_t *keep)
{
	std::vector<int64_t> score((size_t)h + 1);
	std::vector<int32_t> pred((size_t)h + 1), heap((size_t)h + 1);
	return filter_overlaps_max(start, len, h, score.data(), pred.data(), keep, heap.data());
}

// the two implementations behind filter_overlaps_max on their own: which = 0 clusters (with an
// unlimited budget), 1 heap (O(h log h))
int32_t emul_filter_variant(const int32_t *start, const int32_t *len, int32_t h, uint8_t *keep, int32_t which)
{
	std::vector<int64_t> score((size_t)h + 1);
	std::vector<int32_t> pred((size_t)h + 1), heap((size_t)h + 1);
	if (h < 2) {
		for (int32_t k = 0; k < h; k++)
			keep[k] = 1;
		return h;
	}
	if (which == 0) return filter_overlaps_clusters(start, len, h, score.data(), pred.data(), keep, INT64_MAX);
	return filter_overlaps_heap(start, len, h, score.data(), pred.data(), keep, heap.data());
}

} // extern "C"
