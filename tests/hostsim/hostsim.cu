// tests/hostsim/hostsim.cu — TEST TOOLING ONLY, never loaded by the ema_b200 package.
//
// Compiles the thread-scalar (__host__ __device__) control logic of ema_b200/csrc — seeding,
// chaining, chain filtering, mem_chain2aln's seed-extension scheduling, de-duplication, mate rescue,
// mem_reg2aln and the append_alignments filters — for the HOST, so that it can be unit-tested against
// the reference on machines without a GPU.  The DP steps, which on the device are the
// warp-cooperative kernels of ksw_warp.cuh, are supplied here by the oracle's scalar routines
// (oracle/oracle_ksw.c): this file therefore checks the control flow around the kernels, while
// tests/test_gpu_*.py check the kernels themselves and the assembled pipeline on the B200.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../ema_b200/csrc/seed.cuh"
#include "../../ema_b200/csrc/seed_hot.cuh"
#include "../../ema_b200/csrc/chain.cuh"
#include "../../ema_b200/csrc/align.cuh"
#include "../../ema_b200/csrc/align_lanes.cuh"
#include "../../oracle/oracle.h"
#define RESCUE_ROOM_HS 51

thread_local char emab_errbuf[512] = "";

struct HostIndex {
	orc_index_t *o;
	DevIndex d;
	std::vector<uint4> bwt_aligned;  // uint4 loads need 16-byte alignment; the file image is at +40
	std::vector<uint32_t> sa32;
	std::vector<uint64_t> sa64;
	std::vector<uint4> hot, kmer;   // the derived structures of seed_hot.cuh, built by the same thread-scalar code as on the device
};

static void build_hot_kmer(HostIndex *h, int K)
{
	DevIndex &d = h->d;
	const uint64_t n_hot = (d.seq_len >> 6) + 1;
	h->bwt_aligned.resize(h->bwt_aligned.size() + 8);   // the last one-hot block reads a whole 64-byte block
	d.bwt = h->bwt_aligned.data();
	h->hot.assign(n_hot * 4, make_uint4(0, 0, 0, 0));
	for (uint64_t b = 0; b < n_hot; ++b) hot_build_block(d.bwt, b, h->hot.data() + b * 4);
	d.hot = h->hot.data();
	if (K < 0) K = kmer_default_k(d.seq_len);
	if (K > EMAB_KMER_MAX) K = EMAB_KMER_MAX;
	d.kmer_k = K;
	d.kmer = nullptr;
	if (K <= 0) return;
	h->kmer.assign(kmer_total(K), make_uint4(0, 0, 0, 0));
	for (int b = 0; b < 4; ++b) h->kmer[b] = kmer_level1(d, b);
	Fm fm{d, 0};
	for (int t = 1; t < K; ++t) {
		const uint64_t n = 1ull << (2 * t), po = kmer_level_off(t), co = kmer_level_off(t + 1);
		for (uint64_t i = 0; i < n; ++i) kmer_children(fm, h->kmer[po + i], h->kmer.data() + co + i * 4);
	}
	d.kmer = h->kmer.data();
}

// 1: the data flow of the device pipeline's default forms — intervals from the seed_hot.cuh algorithm (x1 = 0) and the SA
// values of a read's occurrences gathered ahead of chain_read in k_sa_gather's enumeration order (pipeline.cu)
static int g_device_like = 0;

static thread_local long long hs_cnt[6];  // DP calls: extend, global, local; cells: extend, global, local (instrumentation for tools)

struct HostDP {  // scalar stand-in for WarpPolicy (pipeline.cu)
	const DevIndex &ix;
	int8_t mat[25];
	explicit HostDP(const DevIndex &i) : ix(i) { orc_fill_scmat(1, 4, mat); }
	ExtResult extend(const uint8_t *query, int q0, int qstep, int qlen, int64_t t0, int tstep, int tlen, int w, int end_bonus, int h0)
	{
		std::vector<uint8_t> q(qlen + 1), t(tlen + 1);
		for (int j = 0; j < qlen; ++j) q[j] = query[q0 + j * qstep];
		for (int i = 0; i < tlen; ++i) t[i] = (uint8_t)ref_base(ix, t0 + (int64_t)i * tstep);
		ExtResult r;
		int64_t cells = 0;
		r.score = orc_ksw_extend2(qlen, q.data(), tlen, t.data(), 5, mat, 6, 1, 6, 1, w, end_bonus, opt::zdrop, h0, &r.qle, &r.tle, &r.gtle, &r.gscore, &r.max_off, &cells);
		++hs_cnt[0]; hs_cnt[3] += cells;
		return r;
	}
	int global(const uint8_t *query, int q0, int qstep, int qlen, int64_t t0, int tstep, int tlen, int w, uint32_t *cigar, int *n_cigar)
	{
		std::vector<uint8_t> q(qlen + 1), t(tlen + 1);
		for (int j = 0; j < qlen; ++j) q[j] = query[q0 + j * qstep];
		for (int i = 0; i < tlen; ++i) t[i] = (uint8_t)ref_base(ix, t0 + (int64_t)i * tstep);
		int64_t cells = 0;
		int sc = orc_ksw_global2(qlen, q.data(), tlen, t.data(), 5, mat, 6, 1, 6, 1, w, cigar ? n_cigar : 0, cigar, EMAB_MAX_CIGAR, &cells);
		++hs_cnt[1]; hs_cnt[4] += cells;
		return sc;
	}
	int ungapped(const uint8_t *query, int q0, int qstep, int n, int64_t t0, int tstep, int *score)
	{
		return ungapped_scalar(ix, query, q0, qstep, n, t0, tstep, score);
	}
	LocResult local(const uint8_t *ms, int l_ms, int64_t rb, int tlen)
	{
		std::vector<uint8_t> q(l_ms + 1), t(tlen + 1);
		for (int j = 0; j < l_ms; ++j) q[l_ms - 1 - j] = ms[j] < 4 ? 3 - ms[j] : 4;
		for (int i = 0; i < tlen; ++i) t[i] = (uint8_t)ref_base(ix, rb + i);
		int out[7];
		int64_t cells = 0;
		orc_ksw_align2(l_ms, q.data(), tlen, t.data(), 5, mat, 6, 1, 6, 1, 19, l_ms * opt::a < 250, out, &cells);
		++hs_cnt[2]; hs_cnt[5] += cells;
		LocResult r{out[0], out[1], out[2], out[3], out[4], out[5], out[6]};
		return r;
	}
};

static void flatten(const Reg &g, int64_t *o)
{
	union { float f; uint32_t u; } fr; fr.f = g.frac_rep;
	o[0] = g.rb; o[1] = g.re; o[2] = g.qb; o[3] = g.qe; o[4] = g.rid; o[5] = g.score; o[6] = g.truesc; o[7] = g.sub; o[8] = g.csub;
	o[9] = g.sub_n; o[10] = g.w; o[11] = g.seedcov; o[12] = g.secondary; o[13] = g.seedlen0; o[14] = g.n_comp; o[15] = 0; o[16] = fr.u; o[17] = 0;
}

extern "C" {

void *hs_index_load(const char *prefix)
{
	orc_index_t *o = orc_index_load(prefix);
	if (!o) return 0;
	HostIndex *h = new HostIndex();
	h->o = o;
	DevIndex &d = h->d;
	memset(&d, 0, sizeof d);
	h->bwt_aligned.resize(o->bwt_size / 4 + 4);
	memcpy(h->bwt_aligned.data(), o->bwt, o->bwt_size * 4);
	d.bwt = h->bwt_aligned.data(); d.n_blocks = o->bwt_size / 16; d.primary = o->primary; d.seq_len = o->seq_len;
	for (int i = 0; i < 5; ++i) d.L2[i] = o->L2[i];
	d.sa_sampled = o->sa; d.sa_intv = o->sa_intv; d.pac = o->pac; d.l_pac = o->l_pac; d.n_seqs = o->n_seqs;
	d.ann_offset = o->ann_offset; d.ann_len = o->ann_len;
	// dense SA, same walk as k_build_dense_sa (api.cu)
	h->sa32.assign(d.seq_len + 1, 0);
	Fm fm{d, 0};
	uint64_t mask = (uint64_t)d.sa_intv - 1;
	for (uint64_t m = 0; m < o->n_sa; ++m) {
		uint64_t k = m * d.sa_intv, s = m == 0 ? d.seq_len : o->sa[m];
		h->sa32[k] = (uint32_t)s;
		for (;;) {
			k = bwt_invPsi(fm, k);
			if ((k & mask) == 0) break;
			--s;
			h->sa32[k] = (uint32_t)s;
		}
	}
	d.sa32 = h->sa32.data();
	build_hot_kmer(h, -1);
	return h;
}

void hs_set_device_like(int on) { g_device_like = on; }

// rebuild the k-mer table with another K (0 = no table): the tests walk the table logic at several depths
void hs_set_kmer_k(void *h_, int K) { build_hot_kmer((HostIndex *)h_, K); }
int hs_kmer_k(void *h_) { return ((HostIndex *)h_)->d.kmer_k; }

// the default seeding form (seed_hot.cuh): intervals as (x0, x1 = 0, x2, info); *sectors = 32-byte sectors requested
int hs_collect_intv_hot(void *h_, int len, const uint8_t *seq, int64_t *out, int max, int64_t *sectors)
{
	HostIndex *h = (HostIndex *)h_;
	std::vector<Intv> mem(EMAB_MAX_INTV), b0(EMAB_MAX_READ_LEN + 1), b1(EMAB_MAX_READ_LEN + 1);
	int ovf = 0;
	unsigned sec = 0;
	int n = collect_intv_hot(h->d, len, seq, mem.data(), EMAB_MAX_INTV, b0.data(), b1.data(), &ovf, &sec);
	for (int i = 0; i < n && i < max; ++i) { out[i*4] = mem[i].x0; out[i*4+1] = mem[i].x1; out[i*4+2] = mem[i].x2; out[i*4+3] = mem[i].info; }
	if (sectors) *sectors = sec;
	return ovf ? -1 : n;
}

int hs_collect_intv(void *h_, int len, const uint8_t *seq, int64_t *out, int max)
{
	HostIndex *h = (HostIndex *)h_;
	std::vector<Intv> mem(EMAB_MAX_INTV), b0(EMAB_MAX_READ_LEN + 1), b1(EMAB_MAX_READ_LEN + 1);
	Fm fm{h->d, 0};
	int ovf = 0;
	int n = collect_intv(fm, len, seq, mem.data(), EMAB_MAX_INTV, b0.data(), b1.data(), &ovf);
	for (int i = 0; i < n && i < max; ++i) { out[i*4] = mem[i].x0; out[i*4+1] = mem[i].x1; out[i*4+2] = mem[i].x2; out[i*4+3] = mem[i].info; }
	return ovf ? -1 : n;
}

// work profile of the default seeding form for one read: out[10], see HotFm::prof (measurement aid for DESIGN.md)
int hs_hot_profile(void *h_, int len, const uint8_t *seq, int64_t *out)
{
	HostIndex *h = (HostIndex *)h_;
	std::vector<Intv> mem(4096), b0(EMAB_MAX_READ_LEN + 1), b1(EMAB_MAX_READ_LEN + 1);
	int ovf = 0;
	unsigned sec = 0, prof[10];
	int n = collect_intv_hot(h->d, len, seq, mem.data(), 4096, b0.data(), b1.data(), &ovf, &sec, prof);
	for (int k = 0; k < 10; ++k) out[k] = prof[k];
	return n;
}

// the nested (reference-shaped) form, for cross-checking the flattened one; *touches = Occ-block loads
int hs_collect_intv_nested(void *h_, int len, const uint8_t *seq, int64_t *out, int max, int64_t *touches)
{
	HostIndex *h = (HostIndex *)h_;
	std::vector<Intv> mem(EMAB_MAX_INTV), b0(EMAB_MAX_READ_LEN + 1), b1(EMAB_MAX_READ_LEN + 1);
	Fm fm{h->d, 0};
	int ovf = 0;
	int n = collect_intv_nested(fm, len, seq, mem.data(), EMAB_MAX_INTV, b0.data(), b1.data(), &ovf);
	for (int i = 0; i < n && i < max; ++i) { out[i*4] = mem[i].x0; out[i*4+1] = mem[i].x1; out[i*4+2] = mem[i].x2; out[i*4+3] = mem[i].info; }
	if (touches) *touches = fm.touches;
	return ovf ? -1 : n;
}

int64_t hs_collect_intv_touches(void *h_, int len, const uint8_t *seq)
{
	HostIndex *h = (HostIndex *)h_;
	std::vector<Intv> mem(EMAB_MAX_INTV), b0(EMAB_MAX_READ_LEN + 1), b1(EMAB_MAX_READ_LEN + 1);
	Fm fm{h->d, 0};
	int ovf = 0;
	collect_intv(fm, len, seq, mem.data(), EMAB_MAX_INTV, b0.data(), b1.data(), &ovf);
	return fm.touches;
}

// Seeding work profile of one read (measurement aid for DESIGN.md: what bounds k_seed).  out[0] = bwt_extend calls of
// passes 1+2, out[1] = of pass 3, out[2] = backward rounds, out[3] = dependent steps of passes 1+2 if the (independent)
// extensions of one backward round ran four at a time, out[4] = largest prev list.
struct CountingFm {
	Fm &fm;
	int64_t calls = 0;
	Intv extend1(const Intv &ik, int c, int is_back, bool = true) { ++calls; return bwt_extend1(fm, ik, c, is_back); }
	const DevIndex &index() const { return fm.ix; }
};
struct ProfilingLists {
	PtrLists base;
	int last_which = -1, last_idx = -100, run = 0;
	int64_t rounds = 0, par4 = 0, serial = 0, longest = 0;
	void close_run() { if (run) { ++rounds; par4 += (run + 3) / 4; serial += run; if (run > longest) longest = run; } run = 0; }
	void put(int which, int idx, const Intv &v) { base.put(which, idx, v); }
	Intv get(int which, int idx)
	{
		const int d = idx - last_idx;
		if (which != last_which || (d != 1 && d != -1)) close_run();
		++run; last_which = which; last_idx = idx;
		return base.get(which, idx);
	}
	void emit(Intv *out, int idx, const Intv &v) { out[idx] = v; }
	void put_by(int, int which, int idx, const Intv &v) { base.put(which, idx, v); }
	void emit_by(int, Intv *out, int idx, const Intv &v) { out[idx] = v; }
	void sync() const {}
};
int hs_seed_profile(void *h_, int len, const uint8_t *seq, int64_t *out)
{
	HostIndex *h = (HostIndex *)h_;
	std::vector<Intv> mem(4096), p3(EMAB_P3_CAP), b0(EMAB_MAX_READ_LEN + 1), b1(EMAB_MAX_READ_LEN + 1);
	Fm fm{h->d, 0};
	CountingFm c12{fm}, c3{fm};
	ProfilingLists lists{PtrLists{{b0.data(), b1.data()}}};
	OneReadFeeder f12{{seq, len, mem.data(), 4096, 0}, false, 0, 0}, f3{{seq, len, p3.data(), EMAB_P3_CAP, 0}, false, 0, 0};
	seed_p12(c12, f12, lists, SoloCoop());
	lists.close_run();
	seed_p3(c3, f3, lists);
	out[0] = c12.calls; out[1] = c3.calls; out[2] = lists.rounds;
	out[3] = c12.calls - lists.serial + lists.par4;
	out[4] = lists.longest;
	return f12.n;
}

struct ReadWork {
	std::vector<Intv> intv, b0, b1;
	std::vector<Seed> w_seeds, seeds;
	std::vector<Chain> w_chains, chains;
	std::vector<BNode> nodes;
	std::vector<int32_t> ord;
	std::vector<uint64_t> srt;
	std::vector<Reg> regs;
	int n_chains = 0, n_regs = 0, cap = 0;
};

static int do_chain(HostIndex *h, int len, const uint8_t *seq, ReadWork &w)
{
	w.intv.resize(EMAB_MAX_INTV); w.b0.resize(EMAB_MAX_READ_LEN + 1); w.b1.resize(EMAB_MAX_READ_LEN + 1);
	Fm fm{h->d, 0};
	int ovf = 0;
	int n = g_device_like ? collect_intv_hot(h->d, len, seq, w.intv.data(), EMAB_MAX_INTV, w.b0.data(), w.b1.data(), &ovf, nullptr)
	                      : collect_intv(fm, len, seq, w.intv.data(), EMAB_MAX_INTV, w.b0.data(), w.b1.data(), &ovf);
	int cap = 0;
	if (len >= opt::min_seed_len) for (int i = 0; i < n; ++i) cap += intv_occ_count(w.intv[i].x2);
	w.cap = cap;
	w.w_seeds.resize(cap + 1); w.seeds.resize(cap + 1); w.w_chains.resize(cap + 1); w.chains.resize(cap + 1);
	w.nodes.resize(cap / 3 + 3); w.ord.resize(3 * cap + 3); w.srt.resize(cap + 1); w.regs.resize(cap + RESCUE_ROOM_HS + 1);
	ChainWork wk{w.w_seeds.data(), w.w_chains.data(), w.nodes.data(), w.ord.data()};
	std::vector<int64_t> sa_vals;
	if (g_device_like && len >= opt::min_seed_len) {
		for (int i = 0; i < n; ++i) {
			const Intv &p = w.intv[i];
			const int cnt = intv_occ_count(p.x2);
			const int64_t step = p.x2 > (uint64_t)opt::max_occ ? (int64_t)(p.x2 / opt::max_occ) : 1;
			for (int t = 0; t < cnt; ++t) sa_vals.push_back((int64_t)bwt_sa_dense(h->d, p.x0 + (uint64_t)(t * step)));
		}
	}
	w.n_chains = chain_read(h->d, len, w.intv.data(), n, wk, cap, w.chains.data(), w.seeds.data(), g_device_like ? sa_vals.data() : nullptr);
	return w.n_chains;
}

// chains after mem_chain_flt: pos,rid,n,w,kept,first,frac_rep_bits,seed_off ; seeds: rbeg,qbeg,len,score
int hs_chain(void *h_, int len, const uint8_t *seq, int64_t *chains, int maxc, int64_t *seeds, int maxs, int *n_seeds_out)
{
	HostIndex *h = (HostIndex *)h_;
	ReadWork w;
	int n = do_chain(h, len, seq, w), ns = 0;
	for (int i = 0; i < n; ++i) {
		const Chain &c = w.chains[i];
		union { float f; uint32_t u; } fr; fr.f = c.frac_rep;
		if (i < maxc) { chains[i*8] = c.pos; chains[i*8+1] = c.rid; chains[i*8+2] = c.n; chains[i*8+3] = c.w; chains[i*8+4] = c.kept; chains[i*8+5] = c.first; chains[i*8+6] = fr.u; chains[i*8+7] = ns; }
		for (int j = 0; j < c.n; ++j, ++ns)
			if (ns < maxs) { const Seed &s = w.seeds[c.seed_beg + j]; seeds[ns*4] = s.rbeg; seeds[ns*4+1] = s.qbeg; seeds[ns*4+2] = s.len; seeds[ns*4+3] = s.score; }
	}
	*n_seeds_out = ns;
	return n;
}

static int do_align1(HostIndex *h, int len, const uint8_t *seq, ReadWork &w, bool dedup)
{
	do_chain(h, len, seq, w);
	HostDP dp(h->d);
	if (dedup) w.n_regs = align1_from_chains(h->d, dp, len, seq, w.chains.data(), w.n_chains, w.seeds.data(), w.srt.data(), w.regs.data());
	else {
		int n = 0;
		for (int i = 0; i < w.n_chains; ++i) chain2aln(h->d, dp, len, seq, w.chains[i], w.seeds.data() + w.chains[i].seed_beg, w.srt.data(), w.regs.data(), &n);
		w.n_regs = n;
	}
	return w.n_regs;
}

int hs_align1(void *h_, int len, const uint8_t *seq, int64_t *regs, int max, int dedup)
{
	ReadWork w;
	int n = do_align1((HostIndex *)h_, len, seq, w, dedup != 0);
	for (int i = 0; i < n && i < max; ++i) flatten(w.regs[i], regs + (size_t)i * 18);
	return n;
}

int hs_pair(void *h_, int l1, const uint8_t *s1, int l2, const uint8_t *s2, int64_t *regs1, int *n1, int64_t *regs2, int *n2, int max)
{
	HostIndex *h = (HostIndex *)h_;
	ReadWork w1, w2;
	do_align1(h, l1, s1, w1, true);
	do_align1(h, l2, s2, w2, true);
	HostDP dp(h->d);
	mate_sw_pair(h->d, dp, l1, s1, l2, s2, w1.regs.data(), &w1.n_regs, w2.regs.data(), &w2.n_regs);
	*n1 = w1.n_regs; *n2 = w2.n_regs;
	for (int i = 0; i < w1.n_regs && i < max; ++i) flatten(w1.regs[i], regs1 + (size_t)i * 18);
	for (int i = 0; i < w2.n_regs && i < max; ++i) flatten(w2.regs[i], regs2 + (size_t)i * 18);
	return 0;
}

// full per-pair path: candidates as emab_aln_t-compatible records (sizeof(Aln) bytes each), mate 1 then mate 2
int hs_candidates(void *h_, double eps, int l1, const uint8_t *s1, int l2, const uint8_t *s2, void *alns, int *n1, int *n2, int max)
{
	HostIndex *h = (HostIndex *)h_;
	ReadWork w1, w2;
	do_align1(h, l1, s1, w1, true);
	do_align1(h, l2, s2, w2, true);
	HostDP dp(h->d);
	mate_sw_pair(h->d, dp, l1, s1, l2, s2, w1.regs.data(), &w1.n_regs, w2.regs.data(), &w2.n_regs);
	static ScoreConsts sc;
	sc.log_match = log(1 - eps); sc.log_mismatch = log(eps); sc.log_indel = log(0.0001); sc.log_clip = log(0.03);
	sc.log10_mismatch = log10(eps); sc.log10_indel = log10(0.0001); sc.log10_clip = log10(0.03);
	const int fac = (int)log(50.0f);
	for (int l = 0; l < 1024; ++l) sc.mapq_len_coef[l] = l < 50 ? 1. : fac / log((double)l);
	*n1 = w1.n_regs; *n2 = w2.n_regs;
	if (w1.n_regs + w2.n_regs > max) return -1;
	Aln *out = (Aln *)alns;
	memset(out, 0, sizeof(Aln) * (w1.n_regs + w2.n_regs));
	int best_dist = -1;
	append_candidates(h->d, dp, sc, l1, s1, w1.regs.data(), w1.n_regs, out, &best_dist);
	append_candidates(h->d, dp, sc, l2, s2, w2.regs.data(), w2.n_regs, out + w1.n_regs, &best_dist);
	return 0;
}

int hs_sizeof_aln(void) { return (int)sizeof(Aln); }
void hs_dp_counters(long long *out, int reset) { for (int i = 0; i < 6; ++i) { out[i] = hs_cnt[i]; if (reset) hs_cnt[i] = 0; } }

// the thread-scalar score-only global alignment the thread-per-read kernel uses for mem_patch_reg
// (align_lanes.cuh): query against ref[t0, t0 + tlen) with band w; *target receives the bases read
int hs_patch_global(void *h_, int qlen, const uint8_t *q, int64_t t0, int tstep, int tlen, int w, uint8_t *target, int64_t *cells)
{
	HostIndex *h = (HostIndex *)h_;
	int err = 0;
	unsigned long long c = 0;
	ScalarPatchDP dp{h->d, &err, &c};
	for (int i = 0; i < tlen; ++i) target[i] = (uint8_t)ref_base(h->d, t0 + (int64_t)i * tstep);
	int sc = dp.global(q, 0, 1, qlen, t0, tstep, tlen, w, nullptr, nullptr);
	if (cells) *cells = (int64_t)c;
	return err ? -999999 : sc;
}

}  // extern "C"
