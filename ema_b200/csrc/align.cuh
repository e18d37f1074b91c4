// Per-read / per-pair control logic of the BWA-MEM bridge, restated for one *warp* per work item:
//   chain2aln           <- mem_chain2aln           (bwa/bwamem.c:658-812)
//   sort_dedup_patch    <- mem_sort_dedup_patch    (bwa/bwamem.c:463-515) + mem_patch_reg (:432-461)
//   matesw              <- mem_matesw              (bwa/bwamem_pair.c:137-206), FR only, fixed insert model
//   mate_sw_pair        <- bwa_mem_mate_sw         (src/bwabridge.c:204-299), the rescue half
//   reg2aln             <- mem_reg2aln             (bwa/bwamem.c:1119-1189) + bwa_gen_cigar2 (bwa/bwa.c:148-234)
//   append_candidates   <- append_alignments       (src/align.c:986-1061) incl. score_alignment (:846-913)
//                          and mem_approx_mapq_se_insist (:959-984)
//
// Everything here is warp-uniform scalar code: all 32 lanes execute it redundantly on identical
// values (global stores are same-address same-value), and the DP steps are delegated to a policy
// object `DP` whose device implementation (WarpDPPolicy, pipeline.cu) runs the warp-cooperative
// kernels of ksw_warp.cuh.  The same templates compile for the host, where tests/hostsim plugs in
// a scalar DP so the control logic can be unit-tested against the reference without a GPU.
#pragma once
#include "chain.cuh"

// Host-precomputed double constants (glibc log/log10, so that no device libm value enters a
// quantity that is later truncated to int or printed).  Filled by emab_ctx / hostsim.
struct ScoreConsts {
	double log_match, log_mismatch, log_indel, log_clip;     // ln(1-eps), ln(eps), ln(1e-4), ln(0.03)  (src/align.c:858-861)
	double log10_mismatch, log10_indel, log10_clip;          // log10 of the same                      (src/align.c:862-864)
	double mapq_len_coef[1024];                               // l < 50 ? 1 : 3 / ln(l)   (src/align.c:971, mapQ_coef_fac is the int 3)
};

EMAB_HD int cal_max_gap(int qlen)
{  // bwa/bwamem.c:647-654
	int l_del = (int)((double)(qlen * opt::a - opt::o_del) / opt::e_del + 1.);
	int l_ins = (int)((double)(qlen * opt::a - opt::o_ins) / opt::e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	l = l > 1 ? l : 1;
	return l < (opt::w << 1) ? l : (opt::w << 1);
}

struct U64Less { EMAB_HD bool operator()(uint64_t a, uint64_t b) const { return a < b; } };

// ---------------------------------------------------------------------------------------------
// mem_chain2aln (bwa/bwamem.c:658-812) in pieces.  The pieces are shared by two drivers: chain2aln below
// (one work item runs the whole function, DP calls delegated to a policy object: the warp-per-read kernels
// and the host build of tests/hostsim) and the thread-per-read state machine of align_lanes.cuh, where
// the DP calls of 32 reads are gathered and run together by the inter-task kernel of ksw_lanes.cuh.
// ---------------------------------------------------------------------------------------------
struct ChainWin { int64_t rmax0, rmax1; };

// the reference window of a chain (bwa/bwamem.c:668-685)
EMAB_HD ChainWin chain_window(const DevIndex &ix, int l_query, const Chain &c, const Seed *seeds)
{
	const int64_t l_pac = ix.l_pac;
	int64_t rmax0 = l_pac << 1, rmax1 = 0;
	for (int i = 0; i < c.n; ++i) {
		const Seed &t = seeds[i];
		int64_t b = t.rbeg - (t.qbeg + cal_max_gap(t.qbeg));
		int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(l_query - t.qbeg - t.len));
		rmax0 = rmax0 < b ? rmax0 : b;
		rmax1 = rmax1 > e ? rmax1 : e;
	}
	rmax0 = rmax0 > 0 ? rmax0 : 0;
	rmax1 = rmax1 < (l_pac << 1) ? rmax1 : (l_pac << 1);
	if (rmax0 < l_pac && l_pac < rmax1) {  // crossing the forward-reverse boundary: keep the seeds' side
		if (seeds[0].rbeg < l_pac) rmax1 = l_pac;
		else rmax0 = l_pac;
	}
	int rid;
	bns_clamp(ix, &rmax0, seeds[0].rbeg, &rmax1, &rid);  // bns_fetch_seq's window (bwa/bwamem.c:685)
	ChainWin w; w.rmax0 = rmax0; w.rmax1 = rmax1;
	return w;
}

// seeds are visited by decreasing score (bwa/bwamem.c:687-690)
EMAB_HD void chain_sort_seeds(const Chain &c, const Seed *seeds, uint64_t *srt)
{
	for (int i = 0; i < c.n; ++i) srt[i] = (uint64_t)seeds[i].score << 32 | (uint64_t)i;
	ks_introsort((size_t)c.n, srt, U64Less());
}

// bwa/bwamem.c:693-729: is seed srt[k] worth extending given the regions found so far?  Clears srt[k] if not.
EMAB_HD bool seed_wants_extension(int l_query, const Chain &c, const Seed *seeds, uint64_t *srt, int k, const Reg *av, int n_av)
{
	const Seed &s = seeds[(uint32_t)srt[k]];
	int i;
	for (i = 0; i < n_av; ++i) {  // has this seed been covered by an earlier extension?
		const Reg &p = av[i];
		if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
		if (s.len - p.seedlen0 > .1 * l_query) continue;
		int qd = s.qbeg - p.qb;
		int64_t rd = s.rbeg - p.rb;
		int max_gap = cal_max_gap(qd < rd ? qd : (int)rd);
		int w = max_gap < p.w ? max_gap : p.w;
		if (qd - rd < w && rd - qd < w) break;
		qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
		max_gap = cal_max_gap(qd < rd ? qd : (int)rd);
		w = max_gap < p.w ? max_gap : p.w;
		if (qd - rd < w && rd - qd < w) break;
	}
	if (i < n_av) {  // (almost) contained: extend only if an overlapping seed suggests a different alignment
		for (i = k + 1; i < c.n; ++i) {
			if (srt[i] == 0) continue;
			const Seed &t = seeds[(uint32_t)srt[i]];
			if (t.len < s.len * .95) continue;
			if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
			if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
		}
		if (i == c.n) { srt[k] = 0; return false; }
	}
	return true;
}

struct SeedExt {  // the region being built from one seed
	Reg a;
	int aw0, aw1, sc0;
};

EMAB_HD void seed_begin(const Chain &c, SeedExt &e)
{
	Reg &a = e.a;
	a.rb = a.re = 0; a.qb = a.qe = 0; a.sub = a.csub = a.sub_n = 0; a.seedcov = 0; a.secondary = 0; a.n_comp = 0;
	e.aw0 = opt::w; e.aw1 = opt::w; e.sc0 = 0;
	a.w = opt::w;
	a.score = a.truesc = -1;
	a.rid = c.rid;
	a.seedlen0 = 0; a.frac_rep = 0;
}

// One ksw_extend2 call of mem_chain2aln, as data: query base j = query[q0 + j*qstep], target base i =
// ref[t0 + i*tstep].
struct ExtTask { int q0, qstep, qlen; int64_t t0; int tstep, tlen, w, end_bonus, h0; };

EMAB_HD ExtTask left_task(const Seed &s, const ChainWin &cw, int t)
{  // reversed query prefix against the reversed reference prefix (bwa/bwamem.c:741-757)
	ExtTask x;
	x.q0 = s.qbeg - 1; x.qstep = -1; x.qlen = s.qbeg;
	x.t0 = s.rbeg - 1; x.tstep = -1; x.tlen = (int)(s.rbeg - cw.rmax0);
	x.w = opt::w << t; x.end_bonus = opt::pen_clip5; x.h0 = s.len * opt::a;
	return x;
}

// consume one left try; returns true if the band must be doubled and the extension redone (MAX_BAND_TRY 2)
EMAB_HD bool left_try_done(SeedExt &e, const ExtResult &r, int t)
{
	const int prev = e.a.score;
	e.aw0 = opt::w << t;
	e.a.score = r.score;
	if (e.a.score == prev || r.max_off < (e.aw0 >> 1) + (e.aw0 >> 2)) return false;
	return t + 1 < 2;
}

EMAB_HD void left_finish(const Seed &s, SeedExt &e, const ExtResult &r)
{
	Reg &a = e.a;
	if (r.gscore <= 0 || r.gscore <= a.score - opt::pen_clip5) {  // local extension
		a.qb = s.qbeg - r.qle; a.rb = s.rbeg - r.tle;
		a.truesc = a.score;
	} else {  // to-end extension
		a.qb = 0; a.rb = s.rbeg - r.gtle;
		a.truesc = r.gscore;
	}
}

EMAB_HD void left_none(const Seed &s, SeedExt &e) { e.a.score = e.a.truesc = s.len * opt::a; e.a.qb = 0; e.a.rb = s.rbeg; }

EMAB_HD ExtTask right_task(int l_query, const Seed &s, const ChainWin &cw, const SeedExt &e, int t)
{  // bwa/bwamem.c:768-785; h0 is the score after the left extension
	const int qe = s.qbeg + s.len;
	const int64_t re = s.rbeg + s.len - cw.rmax0;
	ExtTask x;
	x.q0 = qe; x.qstep = 1; x.qlen = l_query - qe;
	x.t0 = cw.rmax0 + re; x.tstep = 1; x.tlen = (int)(cw.rmax1 - cw.rmax0 - re);
	x.w = opt::w << t; x.end_bonus = opt::pen_clip3; x.h0 = e.sc0;
	return x;
}

EMAB_HD bool right_try_done(SeedExt &e, const ExtResult &r, int t)
{
	const int prev = e.a.score;
	e.aw1 = opt::w << t;
	e.a.score = r.score;
	if (e.a.score == prev || r.max_off < (e.aw1 >> 1) + (e.aw1 >> 2)) return false;
	return t + 1 < 2;
}

EMAB_HD void right_finish(int l_query, const Seed &s, const ChainWin &cw, SeedExt &e, const ExtResult &r)
{
	Reg &a = e.a;
	const int qe = s.qbeg + s.len;
	const int64_t re = s.rbeg + s.len - cw.rmax0;
	if (r.gscore <= 0 || r.gscore <= a.score - opt::pen_clip3) {
		a.qe = qe + r.qle; a.re = cw.rmax0 + re + r.tle;
		a.truesc += a.score - e.sc0;
	} else {
		a.qe = l_query; a.re = cw.rmax0 + re + r.gtle;
		a.truesc += r.gscore - e.sc0;
	}
}

EMAB_HD void right_none(int l_query, const Seed &s, SeedExt &e) { e.a.qe = l_query; e.a.re = s.rbeg + s.len; }

// bwa/bwamem.c:800-810: seed coverage and bookkeeping, then the region joins av[]
EMAB_HD void seed_finish(const Chain &c, const Seed *seeds, const Seed &s, SeedExt &e, Reg *av, int *n_av)
{
	Reg &a = e.a;
	a.seedcov = 0;
	for (int j = 0; j < c.n; ++j) {
		const Seed &t = seeds[j];
		if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
	}
	a.w = e.aw0 > e.aw1 ? e.aw0 : e.aw1;
	a.seedlen0 = s.len;
	a.frac_rep = c.frac_rep;
	av[(*n_av)++] = a;
}

// mem_chain2aln: extend the seeds of one chain into alignment regions appended to av[*n_av..]
template <class DP>
EMAB_HD void chain2aln(const DevIndex &ix, DP &dp, int l_query, const uint8_t *query, const Chain &c, const Seed *seeds,
                       uint64_t *srt, Reg *av, int *n_av)
{
	if (c.n == 0) return;
	const ChainWin cw = chain_window(ix, l_query, c, seeds);
	chain_sort_seeds(c, seeds, srt);
	for (int k = c.n - 1; k >= 0; --k) {
		if (!seed_wants_extension(l_query, c, seeds, srt, k, av, *n_av)) continue;
		const Seed &s = seeds[(uint32_t)srt[k]];
		SeedExt e;
		seed_begin(c, e);
		if (s.qbeg) {  // left extension
			ExtResult r{};
			for (int t = 0;; ++t) {
				const ExtTask x = left_task(s, cw, t);
				r = dp.extend(query, x.q0, x.qstep, x.qlen, x.t0, x.tstep, x.tlen, x.w, x.end_bonus, x.h0);
				if (!left_try_done(e, r, t)) break;
			}
			left_finish(s, e, r);
		} else left_none(s, e);
		if (s.qbeg + s.len != l_query) {  // right extension
			ExtResult r{};
			e.sc0 = e.a.score;
			for (int t = 0;; ++t) {
				const ExtTask x = right_task(l_query, s, cw, e, t);
				r = dp.extend(query, x.q0, x.qstep, x.qlen, x.t0, x.tstep, x.tlen, x.w, x.end_bonus, x.h0);
				if (!right_try_done(e, r, t)) break;
			}
			right_finish(l_query, s, cw, e, r);
		} else right_none(l_query, s, e);
		seed_finish(c, seeds, s, e, av, n_av);
	}
}

// n bases aligned without gaps: substitution score and number of mismatches (the gap-free path of
// bwa_gen_cigar2, bwa/bwa.c:168-176, and its NM count over M runs, :203-210).  The scalar form; a DP policy's
// ungapped() may spread the bases over the lanes of a warp.
EMAB_HD int ungapped_scalar(const DevIndex &ix, const uint8_t *query, int q0, int qstep, int n, int64_t t0, int tstep, int *score)
{
	int sc = 0, mm = 0;
	for (int i = 0; i < n; ++i) {
		const int qc = query[q0 + i * qstep], tc = ref_base(ix, t0 + (int64_t)i * tstep);
		sc += sc_mat(tc, qc);
		mm += qc != tc;
	}
	*score = sc;
	return mm;
}

// ---------------------------------------------------------------------------------------------
// bwa_gen_cigar2's global alignment set-up (bwa/bwa.c:148-234): orientation, band, fast path.
// Returns the score; if cigar != nullptr also the CIGAR (BAM encoding) and NM.
// ---------------------------------------------------------------------------------------------
template <class DP>
EMAB_HD int gen_cigar(const DevIndex &ix, DP &dp, int w_, int l_query, const uint8_t *query, int qb, int64_t rb, int64_t re,
                      uint32_t *cigar, int *n_cigar, int *NM, bool *ok)
{
	const int64_t l_pac = ix.l_pac;
	*ok = false;
	if (n_cigar) *n_cigar = 0;
	if (NM) *NM = -1;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
	// bns_get_seq clamps to [0, 2 l_pac); regions never leave it, so rlen = re - rb
	const int rlen = (int)(re - rb);
	const bool rev = rb >= l_pac;  // reverse both so that indels are left-aligned on the forward strand
	// sequence accessors in DP order
	const int q0 = rev ? qb + l_query - 1 : qb, qstep = rev ? -1 : 1;
	const int64_t t0 = rev ? re - 1 : rb;
	const int tstep = rev ? -1 : 1;
	*ok = true;
	int score;
	if (l_query == rlen && w_ == 0) {  // no gap: no DP
		const int mm = dp.ungapped(query, q0, qstep, l_query, t0, tstep, &score);
		if (cigar) { cigar[0] = (uint32_t)l_query << 4; *n_cigar = 1; if (NM) *NM = mm; }
		return score;
	}
	int max_ins = (int)((double)(((l_query + 1) >> 1) * opt::a - opt::o_ins) / opt::e_ins + 1.);
	int max_del = (int)((double)(((l_query + 1) >> 1) * opt::a - opt::o_del) / opt::e_del + 1.);
	int max_gap = max_ins > max_del ? max_ins : max_del;
	max_gap = max_gap > 1 ? max_gap : 1;
	int dl = rlen - l_query; dl = dl < 0 ? -dl : dl;
	int w = (max_gap + dl + 1) >> 1;
	w = w < w_ ? w : w_;
	int min_w = dl + 3;
	w = w > min_w ? w : min_w;
	if (!cigar) return dp.global(query, q0, qstep, l_query, t0, tstep, rlen, w, nullptr, nullptr);
	score = dp.global(query, q0, qstep, l_query, t0, tstep, rlen, w, cigar, n_cigar);
	if (NM) {  // bwa/bwa.c:196-226 (MD is not needed by EMA)
		int x = 0, y = 0, n_mm = 0, n_gap = 0;
		const int nc = *n_cigar < EMAB_MAX_CIGAR ? *n_cigar : EMAB_MAX_CIGAR;
		for (int k = 0; k < nc; ++k) {
			int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
			if (op == 0) {
				int sc_unused;
				n_mm += dp.ungapped(query, q0 + x * qstep, qstep, len, t0 + (int64_t)y * tstep, tstep, &sc_unused);
				x += len; y += len;
			} else if (op == 2) {
				if (k > 0 && k < *n_cigar - 1) n_gap += len;
				y += len;
			} else if (op == 1) { x += len; n_gap += len; }
		}
		*NM = n_mm + n_gap;
	}
	return score;
}

// mem_patch_reg (bwa/bwamem.c:432-461)
template <class DP>
EMAB_HD int patch_reg(const DevIndex &ix, DP &dp, const uint8_t *query, const Reg &a, const Reg &b, int *_w)
{
	if (query == nullptr) return 0;  // mem_matesw calls the de-duplication with bns = pac = query = 0
	if (a.rb < ix.l_pac && b.rb >= ix.l_pac) return 0;
	if (a.qb >= b.qb || a.qe >= b.qe || a.re >= b.re) return 0;  // not colinear
	int w = (int)((a.re - b.rb) - (a.qe - b.qb));
	w = w > 0 ? w : -w;
	double r = (double)(a.re - b.rb) / (b.re - a.rb) - (double)(a.qe - b.qb) / (b.qe - a.qb);
	r = r > 0. ? r : -r;
	if (a.re < b.rb || a.qe < b.qb) {
		if (w > opt::w << 1 || r >= 0.05f) return 0;   // PATCH_MAX_R_BW
	} else if (w > opt::w << 2 || r >= 0.05f * 2) return 0;
	w += a.w + b.w;
	w = w < opt::w << 2 ? w : opt::w << 2;
	bool ok;
	int score = gen_cigar(ix, dp, w, b.qe - a.qb, query, a.qb, a.rb, b.re, nullptr, nullptr, nullptr, &ok);
	int q_s = (int)((double)(b.qe - a.qb) / ((b.qe - b.qb) + (a.qe - a.qb)) * (b.score + a.score) + .499);
	int r_s = (int)((double)(b.re - a.rb) / ((b.re - b.rb) + (a.re - a.rb)) * (b.score + a.score) + .499);
	if ((double)score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;  // PATCH_MIN_SC_RATIO
	*_w = w;
	return score;
}

struct RegLessRe { EMAB_HD bool operator()(const Reg &a, const Reg &b) const { return a.re < b.re; } };  // alnreg_slt2
struct RegLessScore {  // alnreg_slt (bwa/bwamem.c:420)
	EMAB_HD bool operator()(const Reg &a, const Reg &b) const
	{
		return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb)));
	}
};

// mem_sort_dedup_patch (bwa/bwamem.c:463-515); query == nullptr reproduces the bns=0 call of mem_matesw
template <class DP>
EMAB_HD int sort_dedup_patch(const DevIndex &ix, DP &dp, const uint8_t *query, int n, Reg *a)
{
	if (n <= 1) return n;
	ks_introsort((size_t)n, a, RegLessRe());
	for (int i = 0; i < n; ++i) a[i].n_comp = 1;
	for (int i = 1; i < n; ++i) {
		Reg &p = a[i];
		if (p.rid != a[i - 1].rid || p.rb >= a[i - 1].re + opt::max_chain_gap) continue;
		for (int j = i - 1; j >= 0 && p.rid == a[j].rid && p.rb < a[j].re + opt::max_chain_gap; --j) {
			Reg &q = a[j];
			if (q.qe == q.qb) continue;  // excluded
			int64_t orr = q.re - p.rb;
			int64_t oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
			int64_t mr = q.re - q.rb < p.re - p.rb ? q.re - q.rb : p.re - p.rb;
			int64_t mq = q.qe - q.qb < p.qe - p.qb ? q.qe - q.qb : p.qe - p.qb;
			int score, w;
			if (orr > opt::mask_level_redun * mr && oq > opt::mask_level_redun * mq) {  // one of the hits is redundant
				if (p.score < q.score) { p.qe = p.qb; break; }
				else q.qe = q.qb;
			} else if (q.rb < p.rb && (score = patch_reg(ix, dp, query, q, p, &w)) > 0) {  // merge q into p
				p.n_comp += q.n_comp + 1;
				p.seedcov = p.seedcov > q.seedcov ? p.seedcov : q.seedcov;
				p.sub = p.sub > q.sub ? p.sub : q.sub;
				p.csub = p.csub > q.csub ? p.csub : q.csub;
				p.qb = q.qb; p.rb = q.rb;
				p.truesc = p.score = score;
				p.w = w;
				q.qb = q.qe;
			}
		}
	}
	int m = 0;
	for (int i = 0; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	n = m;
	ks_introsort((size_t)n, a, RegLessScore());
	for (int i = 1; i < n; ++i)
		if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
	m = n > 0 ? 1 : 0;
	for (int i = 1; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	return m;
}

// mem_align1_core (bwa/bwamem.c:1081-1117) after chaining: extend every chain, then de-duplicate.
template <class DP>
EMAB_HD int align1_from_chains(const DevIndex &ix, DP &dp, int l_query, const uint8_t *query, const Chain *chains, int n_chains,
                               const Seed *seeds, uint64_t *srt, Reg *regs)
{
	int n = 0;
	for (int i = 0; i < n_chains; ++i)
		chain2aln(ix, dp, l_query, query, chains[i], seeds + chains[i].seed_beg, srt, regs, &n);
	return sort_dedup_patch(ix, dp, query, n, regs);
}

// mem_infer_dir (bwa/bwamem_pair.c:49-56)
EMAB_HD int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
	int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
	int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

// mem_matesw with EMA's pes[] (src/bwabridge.c:216-229): only orientation 1 (FR) is live, low = -35,
// high = 500.  `a` is a hit of the anchor read, ms the mate's sequence; ma the mate's region list.
// The part of mem_matesw that decides whether a local SW runs and on which window (bwa/bwamem_pair.c:143-172):
// false when a hit of the mate already pairs with the anchor, or the clamped window leaves the anchor's contig
// or is shorter than a seed.  The window depends on the anchor and the mate's length only — not on `ma` — so
// the SW result for an (anchor, mate) is a pure function the pipeline may compute ahead of time.
EMAB_HD bool matesw_window(const DevIndex &ix, const Reg &a, int l_ms, const Reg *ma, int n_ma, int64_t *rb_, int64_t *re_)
{
	const int64_t l_pac = ix.l_pac;
	const int low = -35, high = 500;
	for (int i = 0; i < n_ma; ++i) {
		int64_t dist;
		int r = infer_dir(l_pac, a.rb, ma[i].rb, &dist);
		if (r == 1 && dist >= low && dist <= high) return false;  // a consistent pair exists
	}
	// r = 1: is_rev = 1, is_larger = 1
	int64_t rb = (a.rb + low) - l_ms;
	int64_t re = a.rb + high;
	if (rb < 0) rb = 0;
	if (re > l_pac << 1) re = l_pac << 1;
	int rid = -1;
	if (rb < re) bns_clamp(ix, &rb, (rb + re) >> 1, &re, &rid);
	*rb_ = rb; *re_ = re;
	return a.rid == rid && re - rb >= opt::min_seed_len;
}

template <class DP>
EMAB_HD int matesw(const DevIndex &ix, DP &dp, const Reg &a, int l_ms, const uint8_t *ms, Reg *ma, int *n_ma)
{
	const int64_t l_pac = ix.l_pac;
	int64_t rb, re;
	const bool run = matesw_window(ix, a, l_ms, ma, *n_ma, &rb, &re);
	int n = 0;
	if (run) {
		// query = reverse complement of ms; target = ref[rb, re)
		LocResult aln = dp.local(ms, l_ms, rb, (int)(re - rb));
		if (aln.score >= opt::min_seed_len && aln.qb >= 0) {
			Reg b;
			b.rid = a.rid;
			b.qb = l_ms - (aln.qe + 1);
			b.qe = l_ms - aln.qb;
			b.rb = (l_pac << 1) - (rb + aln.te + 1);
			b.re = (l_pac << 1) - (rb + aln.tb);
			b.score = aln.score;
			b.truesc = 0; b.sub = 0; b.sub_n = 0; b.w = 0; b.seedlen0 = 0; b.n_comp = 0; b.frac_rep = 0;
			b.csub = aln.score2;
			b.secondary = -1;
			b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
			int i;
			for (i = 0; i < *n_ma; ++i)  // insertion point keeps ma sorted by score
				if (ma[i].score < b.score) break;
			for (int j = *n_ma; j > i; --j) ma[j] = ma[j - 1];
			ma[i] = b;
			++*n_ma;
		}
		++n;
	}
	if (n) *n_ma = sort_dedup_patch(ix, dp, (const uint8_t *)nullptr, *n_ma, ma);
	return n;
}

// The rescue half of bwa_mem_mate_sw (src/bwabridge.c:239-283): up to 50 hits of each mate within
// score_delta = 25 of its best are used as anchors to rescue the other mate.
template <class DP>
EMAB_HD void mate_sw_pair(const DevIndex &ix, DP &dp, int l1, const uint8_t *s1, int l2, const uint8_t *s2,
                          Reg *r1, int *n1, Reg *r2, int *n2)
{
	const int score_delta = 25;
	int best1 = 0, best2 = 0;
	for (int i = 0; i < *n1; ++i) if (r1[i].score > best1) best1 = r1[i].score;
	for (int i = 0; i < *n2; ++i) if (r2[i].score > best2) best2 = r2[i].score;
	int num = 0;
	const int n2_0 = *n2;
	for (int i = 0; i < n2_0 && num < opt::max_matesw; ++i)
		if (r2[i].score >= best2 - score_delta) { ++num; matesw(ix, dp, r2[i], l1, s1, r1, n1); }
	num = 0;
	for (int i = 0; i < *n1 && num < opt::max_matesw; ++i) {  // NB results1.n is re-read every iteration in the reference
		if (r1[i].score >= best1 - score_delta) { ++num; Reg anchor = r1[i]; matesw(ix, dp, anchor, l2, s2, r2, n2); }
	}
}

EMAB_HD int infer_bw(int l1, int l2, int score, int a, int q, int r)
{  // bwa/bwamem.c:818-825
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	int w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
	int d = l1 - l2; d = d < 0 ? -d : d;
	if (w < d) w = d;
	return w;
}

// mem_reg2aln (bwa/bwamem.c:1119-1189) without XA/MD.  Fills pos,rid,is_rev,NM,cigar,score.
template <class DP>
EMAB_HD void reg2aln(const DevIndex &ix, DP &dp, int l_query, const uint8_t *query, const Reg &ar, Aln *out)
{
	int qb = ar.qb, qe = ar.qe;
	int64_t rb = ar.rb, re = ar.re;
	int tmp = infer_bw(qe - qb, (int)(re - rb), ar.truesc, opt::a, opt::o_del, opt::e_del);
	int w2 = infer_bw(qe - qb, (int)(re - rb), ar.truesc, opt::a, opt::o_ins, opt::e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt::w) w2 = w2 < ar.w ? w2 : ar.w;
	int i = 0, score = 0, last_sc = -(1 << 30), NM = -1, n_cigar = 0;
	bool ok;
	do {
		w2 = w2 < opt::w << 2 ? w2 : opt::w << 2;
		score = gen_cigar(ix, dp, w2, qe - qb, query, qb, rb, re, out->cigar, &n_cigar, &NM, &ok);
		if (score == last_sc || w2 == opt::w << 2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < ar.truesc - opt::a);
	int is_rev;
	int64_t pos = bns_depos(ix.l_pac, rb < ix.l_pac ? rb : re - 1, &is_rev);
	uint32_t *cg = out->cigar;
	if (n_cigar > EMAB_MAX_CIGAR - 2) n_cigar = EMAB_MAX_CIGAR - 2;  // overflow is flagged by the caller through n_cigar_raw
	if (n_cigar > 0) {  // squeeze out a leading or trailing deletion
		if ((cg[0] & 0xf) == 2) {
			pos += cg[0] >> 4;
			--n_cigar;
			for (int k = 0; k < n_cigar; ++k) cg[k] = cg[k + 1];
		} else if ((cg[n_cigar - 1] & 0xf) == 2) --n_cigar;
	}
	if (qb != 0 || qe != l_query) {  // clipping
		int clip5 = is_rev ? l_query - qe : qb;
		int clip3 = is_rev ? qb : l_query - qe;
		if (clip5) {
			for (int k = n_cigar; k > 0; --k) cg[k] = cg[k - 1];
			cg[0] = (uint32_t)clip5 << 4 | 3;
			++n_cigar;
		}
		if (clip3) cg[n_cigar++] = (uint32_t)clip3 << 4 | 3;
	}
	out->rid = bns_pos2rid(ix, pos);
	out->pos = pos - ix.ann_offset[out->rid];
	out->is_rev = is_rev;
	out->NM = NM;
	out->n_cigar = n_cigar;
	out->score = ar.score;
}

// mem_approx_mapq_se_insist (src/align.c:959-984); sub = sub_n = 0 for every region on this path
EMAB_HD int approx_mapq_insist(const ScoreConsts &sc, const Reg &a)
{
	int sub = a.sub ? a.sub : opt::min_seed_len * opt::a;
	sub = a.csub > sub ? a.csub : sub;
	if (sub >= a.score) return 0;
	int l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
	double identity = 1. - (double)(l * opt::a - a.score) / (opt::a + opt::b) / l;
	int mapq;
	if (a.score == 0) mapq = 0;
	else {
		double tmp = sc.mapq_len_coef[l < 1024 ? l : 1023];
		tmp *= identity * identity;
		mapq = (int)(6.02 * (a.score - sub) / opt::a * tmp * tmp + .499);
	}
	// a.sub_n is 0 on this path (never set: EMA does not call mem_mark_primary_se)
	if (mapq > 254) mapq = 254;
	if (mapq < 0) mapq = 0;
	mapq = (int)(mapq * (1. - a.frac_rep) + .499);
	return mapq;
}

// multiply-add without contraction, to match the reference's x86-64 build (no FMA)
EMAB_HD double emab_dmul(double a, double b)
{
#ifdef __CUDA_ARCH__
	return __dmul_rn(a, b);
#else
	volatile double r = a * b;
	return r;
#endif
}
EMAB_HD double emab_dadd(double a, double b)
{
#ifdef __CUDA_ARCH__
	return __dadd_rn(a, b);
#else
	volatile double r = a + b;
	return r;
#endif
}

// score_alignment (src/align.c:846-913)
EMAB_HD void score_alignment(const ScoreConsts &sc, Aln *s)
{
	int matches = 0, indels = 0, indels_to_count = 0, clipping = 0;
	for (int i = 0; i < s->n_cigar; ++i) {
		uint32_t type = s->cigar[i] & 0xf, n = s->cigar[i] >> 4;
		if (type == 0) matches += n;
		else if (type == 1 || type == 2) { indels += n; ++indels_to_count; }
		else clipping += n;
	}
	int mismatches = s->NM - indels;
	matches -= mismatches;
	s->em_score = emab_dadd(emab_dadd(emab_dadd(emab_dmul(matches, sc.log_match), emab_dmul(mismatches, sc.log_mismatch)), emab_dmul(indels_to_count, sc.log_indel)),
	                   emab_dmul(clipping, sc.log_clip));
	s->score_mapq = (int)emab_dadd(emab_dadd(emab_dadd(60.0, emab_dmul(mismatches, sc.log10_mismatch)), emab_dmul(indels_to_count, sc.log10_indel)),
	                          emab_dmul(clipping, sc.log10_clip));
}

// append_alignments (src/align.c:986-1061) for one mate: CIGAR every region, apply the clip and
// edit-distance filters, score the survivors.  *best_dist carries over from mate 1 to mate 2 as in
// the reference (it is initialised once per pair and only set by hit 0 of each mate).
template <class DP>
EMAB_HD int append_candidates(const DevIndex &ix, DP &dp, const ScoreConsts &sc, int len, const uint8_t *seq, const Reg *regs, int n_regs,
                              Aln *out, int *best_dist)
{
	int added = 0;
	for (int i = 0; i < n_regs; ++i) {
		Aln &r = out[i];
		reg2aln(ix, dp, len, seq, regs[i], &r);
		r.keep = 0;
		r.clip_edit_dist = 0; r.mapq = 0; r.score_mapq = 0; r.em_score = 0;  // defined bytes on the wire for dropped records too (initcheck)
		const int clip = len - (regs[i].qe - regs[i].qb);
		r.clip = clip;
		if (clip >= len / 2) continue;
		const int dist = r.NM + clip;
		r.clip_edit_dist = dist;
		if (i == 0) *best_dist = dist;
		else if (dist - *best_dist > 12) continue;  // EXTRA_SEARCH_DEPTH
		r.mapq = approx_mapq_insist(sc, regs[i]);
		score_alignment(sc, &r);
		r.keep = 1;
		++added;
	}
	return added;
}
