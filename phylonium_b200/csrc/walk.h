// The anchor walk of /root/reference/src/process.cxx:198-295, restructured so that one
// query can be walked by many threads at once and still give the reference's result
// bit for bit.  Shared by the CUDA kernels (anchor.cu) and the CPU emulation tests.
//
// Reference semantics.  The walk visits query positions
//     pos -> pos + len(pos) + 1
// where len is the "lucky" match on the diagonal of the last accepted anchor if that
// reaches the threshold, else the ESA longest match.  What happens at pos depends only on
//     (pos, diag, endQ)  with diag = lastS - lastQ, endQ = lastQ + lastLen
// and only while pos - endQ <= thr; beyond that (or before any anchor) the walk is
// "dead": it depends on pos alone.  Two walks in the same effective state have the same
// future.  That is what makes speculation exact:
//
//   1. walk_chunk:   every CH-sized chunk of the query is walked from a cold (dead) state
//                    at its first base; chunk 0 starts in the true initial state.  Each
//                    walker records the anchors it accepts ("events"), a bitmap of the
//                    positions it visited while dead, and its exit state.
//   2. (anchor.cu)   matches longer than CAP are left "open" and finished cooperatively.
//   3. bridge:       from every exit state walk on until the state provably coincides with
//                    a recorded state of the walker that owns the current position
//                    (same dead position, or an accepted anchor on the same diagonal
//                    overlapping a recorded one).  Bounded; may give up (UNRESOLVED).
//   4. (anchor.cu)   follow walker 0 -> bridge -> walker j -> bridge … to mark the true
//                    path; a give-up on the true path is continued exactly (serially).
//   5. assemble:     classify the true events into right/left anchors and emit homologies.
//
// Homology assembly needs only consecutive events: event t is a "right anchor" iff it lies
// on the diagonal of event t-1, to its right, on the same strand (process.cxx:251-253).
#pragma once
#include "esa_search.h"

namespace phy
{

struct Event {
	int32_t pos;  // query position of the anchor
	int32_t posS; // position in S
	int32_t len;  // match length (provisional while the event is open)
};

struct WalkState {
	int32_t pos;
	int32_t lastQ, lastS, lastLen;
	int32_t live; // 0: no anchor accepted yet
};

struct StepOut {
	int32_t accepted, open, posS, len;
};

PHY_HD bool walk_is_dead(const WalkState &st, int32_t thr)
{
	return !st.live || (st.pos - (st.lastQ + st.lastLen) > thr);
}

// One iteration of the while loop at process.cxx:245-282, without the homology
// bookkeeping.  q points at the query's first base.
// COOP: see match_run in esa_search.h.
template <int COOP = 0>
PHY_HD StepOut walk_step(const EsaView &e, const uint8_t *q, int32_t qlen, int32_t thr, int32_t cap,
                         const WalkState &st)
{
	StepOut o;
	o.accepted = 0;
	o.open = 0;
	o.posS = 0;
	o.len = 0;
	const int32_t pos = st.pos;
	if (st.live) { // lucky_anchor, process.cxx:227-242
		const int32_t advance = pos - st.lastQ;
		const int32_t gap = advance - st.lastLen;
		const int64_t tr = (int64_t)st.lastS + advance;
		if (tr < (int64_t)e.m && gap <= thr) {
			const int32_t limit = qlen - pos;
			const int32_t lim = limit < cap ? limit : cap;
			const int32_t k = match_run<COOP>(q + pos, e.S + tr, 0, lim);
			o.posS = (int32_t)tr;
			o.len = k;
			o.accepted = k >= thr;
			o.open = (k >= lim && lim < limit) ? 1 : 0;
		}
	}
	if (!o.accepted) { // anchor, process.cxx:219-225
		Match mt = esa_match<COOP>(e, q + pos, qlen - pos, cap);
		o.len = mt.l > 0 ? mt.l : 0;
		o.posS = mt.sa; // SA[i] of a unique match; unused otherwise
		o.accepted = (mt.i == mt.j && o.len >= thr);
		o.open = o.accepted ? mt.open : 0;
	}
	return o;
}

PHY_HD void walk_accept(WalkState &st, const StepOut &o)
{
	st.lastQ = st.pos;
	st.lastS = o.posS;
	st.lastLen = o.len;
	st.live = 1;
}

// ---------------------------------------------------------------------------------
// Per-query and per-chunk records

struct QueryInfo {
	int64_t qoff;       // offset of the query's first base in the concatenated buffer
	int32_t qlen;
	int32_t chunk_base; // global index of this query's chunk 0
	int32_t nchunks;
	int32_t pad;
};

enum : int32_t {
	LINK_NONE = 0,       // not yet computed
	LINK_MERGED = 1,     // path continues in walker `chunk` at event index `from`
	LINK_END = 2,        // the walk reached the end of the query
	LINK_UNRESOLVED = 3, // gave up; state saved for the exact continuation
};

struct ChunkRec {
	WalkState exit; // state after the walker's last step (pos >= chunk end unless open)
	int32_t n_events;
	int32_t open; // last event hit the cap; exit.pos still at the event
	// bridge result
	int32_t link;       // LINK_*
	int32_t link_chunk; // global chunk index the path continues in
	int32_t link_from;  // first event of that walker that is on the path
	int32_t n_bridge;   // events recorded by the bridge
	WalkState bstate;   // bridge state when it stopped (for the continuation)
	Event *bridge_ev;   // where the bridge events live (pool slot, or an overflow buffer)
};

struct WalkParams {
	EsaView esa;
	const uint8_t *Q;      // all queries, each followed by >= 1 zero byte
	const QueryInfo *qi;
	int32_t nq;
	int32_t thr;
	int32_t CH;      // chunk length, multiple of 32
	int32_t CAP;     // per-thread comparison cap (>= thr + 1)
	int32_t cap_ev;  // event slots per chunk: CH / (thr + 1) + 2
	int32_t total_chunks;
	int32_t perm_mul; // launch-order permutation of the walkers: walker w takes chunk w * perm_mul mod total
	Event *ev;          // total_chunks * cap_ev
	Event *bev;         // total_chunks * cap_ev (bridge events)
	uint32_t *dead;     // total_chunks * CH / 32 bitmap words, zero-initialised
	ChunkRec *rec;      // total_chunks
	int32_t *chunk_query; // total_chunks: owning query
	const int *skip = nullptr; // device flag of a lazily built index (EsaDevice::skip): != 0, the index is not there
};

// ---------------------------------------------------------------------------------
// Phase 1: cold walk of one chunk

template <int COOP = 0> PHY_HD void walk_chunk(const WalkParams &P, int32_t g)
{
	const int32_t qid = P.chunk_query[g];
	const QueryInfo qi = P.qi[qid];
	const int32_t k = g - qi.chunk_base;
	const uint8_t *q = P.Q + qi.qoff;
	const int32_t begin = k * P.CH;
	const int32_t end = (begin + P.CH < qi.qlen) ? begin + P.CH : qi.qlen;
	Event *ev = P.ev + (int64_t)g * P.cap_ev;
	uint32_t *bits = P.dead + (int64_t)g * (P.CH / 32);

	WalkState st;
	st.pos = begin;
	st.lastQ = st.lastS = st.lastLen = 0;
	st.live = (k == 0) ? 1 : 0; // process.cxx:206-213: lucky is live on diagonal 0 from the start

	int32_t nev = 0, open = 0;
	int32_t cur_word = -1;
	uint32_t cur_bits = 0;
	while (st.pos < end) {
		if (walk_is_dead(st, P.thr)) {
			const int32_t rel = st.pos - begin;
			if ((rel >> 5) != cur_word) {
				if (cur_word >= 0) bits[cur_word] = cur_bits;
				cur_word = rel >> 5;
				cur_bits = 0;
			}
			cur_bits |= 1u << (rel & 31);
		}
		StepOut o = walk_step<COOP>(P.esa, q, qi.qlen, P.thr, P.CAP, st);
		if (o.accepted) {
			ev[nev].pos = st.pos;
			ev[nev].posS = o.posS;
			ev[nev].len = o.len;
			nev++;
			walk_accept(st, o);
			if (o.open) {
				open = 1;
				break;
			}
		}
		st.pos += o.len + 1;
	}
	if (cur_word >= 0) bits[cur_word] = cur_bits;

	ChunkRec &r = P.rec[g];
	r.exit = st;
	r.n_events = nev;
	r.open = open;
	r.link = LINK_NONE;
	r.link_chunk = -1;
	r.link_from = 0;
	r.n_bridge = 0;
	r.bridge_ev = P.bev + (int64_t)g * P.cap_ev;
}

// ---------------------------------------------------------------------------------
// Phase 2: an open match (longer than CAP) of walker g.  Either it runs into the open
// match of a later walker on the same diagonal (then both end at the same mismatch:
// link), or we scan on to its end.  first_mismatch(q, S, diag, m, from, to) returns the
// first query position in [from, to) whose base differs from S[pos + diag] (positions
// at or beyond m count as different), or `to`.
template <typename MismatchF>
PHY_HD void open_resolve_one(const WalkParams &P, int32_t g, MismatchF first_mismatch, int32_t &link, int32_t &end)
{
	const ChunkRec &r = P.rec[g];
	const QueryInfo qi = P.qi[P.chunk_query[g]];
	const uint8_t *q = P.Q + qi.qoff;
	const Event e = P.ev[(int64_t)g * P.cap_ev + r.n_events - 1];
	const int64_t diag = (int64_t)e.posS - e.pos;
	int32_t verified_end = e.pos + e.len; // first query position not yet compared
	int32_t c = (g - qi.chunk_base) + 1;  // next walker (local index) to ask
	link = -1;
	end = -1;
	for (;;) {
		while (c < qi.nchunks && (int64_t)c * P.CH <= verified_end) {
			const int32_t g2 = qi.chunk_base + c;
			const ChunkRec &r2 = P.rec[g2];
			if (r2.open) {
				const Event e2 = P.ev[(int64_t)g2 * P.cap_ev + r2.n_events - 1];
				if ((int64_t)e2.posS - e2.pos == diag && e2.pos <= verified_end) {
					link = g2;
					return;
				}
			}
			c++;
		}
		int64_t target = (c < qi.nchunks) ? (int64_t)c * P.CH : qi.qlen;
		if (target > qi.qlen) target = qi.qlen;
		const int32_t z = first_mismatch(q, P.esa.S, diag, P.esa.m, verified_end, (int32_t)target);
		if (z < target || target >= qi.qlen) {
			end = z;
			return;
		}
		verified_end = (int32_t)target;
	}
}

// ---------------------------------------------------------------------------------
// Merge tests used by bridges and continuations

// index of the first event of walker g with pos >= p
PHY_HD int32_t events_lower_bound(const Event *ev, int32_t n, int32_t p)
{
	int32_t lo = 0, hi = n;
	while (lo < hi) {
		int32_t mid = (lo + hi) >> 1;
		if (ev[mid].pos < p)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

// Does the anchor (pos, posS, verified v) coincide with a resolved event of walker g?
// Both are maximal matches on one diagonal; if they overlap they end at the same mismatch.
// Returns the event index or -1.
PHY_HD int32_t match_event(const WalkParams &P, int32_t g, int32_t pos, int32_t posS, int32_t v)
{
	const ChunkRec &r = P.rec[g];
	const Event *ev = P.ev + (int64_t)g * P.cap_ev;
	const int32_t n = r.n_events;
	if (n == 0) return -1;
	const int32_t diag = posS - pos;
	int32_t idx = events_lower_bound(ev, n, pos + 1) - 1; // last event with pos' <= pos
	if (idx >= 0) {
		const Event x = ev[idx];
		if (x.posS - x.pos == diag && pos < x.pos + x.len) return idx;
	}
	idx++;
	if (idx < n) {
		const Event x = ev[idx];
		if (x.posS - x.pos == diag && x.pos <= pos + v) return idx;
	}
	return -1;
}

PHY_HD bool dead_visited(const WalkParams &P, int32_t g, int32_t rel)
{
	return (P.dead[(int64_t)g * (P.CH / 32) + (rel >> 5)] >> (rel & 31)) & 1u;
}

// ---------------------------------------------------------------------------------
// Phase 3: bridge from the exit state of walker g.  `budget` limits how far it may walk
// (in query bases); budget < 0 means until merged or the query ends.  `out` receives
// the accepted anchors (at most out_cap); cap is the comparison cap for this walk.
// Returns LINK_*; fills r.link_chunk / r.link_from / r.n_bridge / r.bstate.
template <int COOP = 0>
PHY_HD int32_t bridge_walk(const WalkParams &P, int32_t g, WalkState st, int32_t nb, Event *out,
                           int64_t out_cap, int32_t budget, int32_t cap)
{
	const int32_t qid = P.chunk_query[g];
	const QueryInfo qi = P.qi[qid];
	const uint8_t *q = P.Q + qi.qoff;
	ChunkRec &r = P.rec[g];
	const int32_t stop = budget < 0 ? qi.qlen : ((int64_t)st.pos + budget < qi.qlen ? st.pos + budget : qi.qlen);
	int32_t result = LINK_UNRESOLVED;

	// state right after an anchor: try to merge on that anchor first (zero steps)
	if (nb == 0 && st.live && st.pos < qi.qlen && st.pos == st.lastQ + st.lastLen + 1) {
		// A recorded anchor on our diagonal that covers our last matched base is maximal
		// like ours, so it ends at the same mismatch; its walker is then in our state.
		const int32_t c = qi.chunk_base + (st.lastQ + st.lastLen - 1) / P.CH;
		if (c > g && st.lastLen > 0) {
			const int32_t idx = match_event(P, c, st.lastQ + st.lastLen - 1, st.lastS + st.lastLen - 1, 0);
			if (idx >= 0) {
				r.link_chunk = c;
				r.link_from = idx + 1;
				r.n_bridge = 0;
				r.bstate = st;
				return LINK_MERGED;
			}
		}
	}

	while (true) {
		if (st.pos >= qi.qlen) {
			result = LINK_END;
			break;
		}
		if (st.pos >= stop) break;
		const int32_t c = qi.chunk_base + st.pos / P.CH;
		const int32_t rel = st.pos - (c - qi.chunk_base) * P.CH;
		if (c > g && walk_is_dead(st, P.thr) && dead_visited(P, c, rel)) {
			r.link_chunk = c;
			r.link_from = events_lower_bound(P.ev + (int64_t)c * P.cap_ev, P.rec[c].n_events, st.pos);
			result = LINK_MERGED;
			break;
		}
		StepOut o = walk_step<COOP>(P.esa, q, qi.qlen, P.thr, cap, st);
		if (o.accepted) {
			int32_t idx = (c > g) ? match_event(P, c, st.pos, o.posS, o.len) : -1;
			if (idx >= 0) {
				const Event x = P.ev[(int64_t)c * P.cap_ev + idx];
				o.len = x.pos + x.len - st.pos; // same end as the recorded anchor
				o.open = 0;
			}
			if (o.open || nb >= out_cap) {
				// cannot finish this anchor here: leave it to the exact continuation
				break;
			}
			out[nb].pos = st.pos;
			out[nb].posS = o.posS;
			out[nb].len = o.len;
			nb++;
			walk_accept(st, o);
			if (idx >= 0) {
				st.pos += o.len + 1;
				r.link_chunk = c;
				r.link_from = idx + 1;
				result = LINK_MERGED;
				break;
			}
		}
		st.pos += o.len + 1;
	}
	r.n_bridge = nb;
	r.bstate = st;
	return result;
}

// ---------------------------------------------------------------------------------
// Phase 5: homology assembly from the true event list of one query.
// process.cxx:246-278 and 284-292, process.h:72-80.

struct Hom {
	int32_t dir;   // 0 forward, 1 reverse
	int32_t iref;  // index_reference (position in S)
	int32_t iproj; // index_reference_projected
	int32_t iq;    // index_query
	int32_t len;
};

// event t (t >= 0) continues the homology of its predecessor; prev == virtual (0,0,0) for t == 0
PHY_HD bool event_is_right(const Event &prev, const Event &cur, int32_t border)
{
	const int32_t endS = prev.posS + prev.len;
	const int32_t endQ = prev.pos + prev.len;
	return cur.posS > endS && (cur.pos - endQ) == (cur.posS - endS) &&
	       ((cur.posS < border) == (prev.posS < border));
}

// Homology of the run [a, b] of events (a == -1 stands for the virtual anchor at (0,0,0)).
// Returns false if the run is not pushed (process.cxx:261,289).
PHY_HD bool run_homology(const Event *ev, int32_t a, int32_t b, int32_t thr, int32_t border, Hom &h)
{
	const Event last = (b >= 0) ? ev[b] : Event{0, 0, 0};
	const bool was_right = b > a;
	if (!was_right && last.len / 2 < thr) return false;
	const Event first = (a >= 0) ? ev[a] : Event{0, 0, 0};
	h.dir = 0;
	h.iref = first.posS;
	h.iproj = first.posS;
	h.iq = first.pos;
	h.len = last.pos + last.len - first.pos;
	if (h.iref >= border) { // reverseEh
		h.iproj = 2 * border + 1 - h.len - h.iref;
		h.dir = 1;
	}
	return true;
}

} // namespace phy
