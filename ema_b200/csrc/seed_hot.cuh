// SMEM seeding over an index laid out for 180 GB of HBM: the same three passes of mem_collect_intv
// (bwa/bwamem.c:140-188) as seed.cuh, producing the same intervals (x[0], x[2], info — what mem_chain reads,
// bwa/bwamem.c:163-164,294-309), from three derived structures built on the device at index load:
//
//   hot    one-hot Occ blocks: per 64 BWT symbols and per base c one 16-byte entry {u64 cumulative count of c before the
//          block, u64 bit mask "symbol i is c"}.  Occ(c, k) = count + popc(mask & prefix(k)): ONE 16-byte load and one
//          64-bit popcount per (position, base), against a 64-byte block and ~100 instructions of 2-bit plane counting
//          over bwa's layout (bwa/bwt.c:169-220).  Costs seq_len bytes (6.2 GB at 3.1 Gbp) instead of seq_len / 2.
//   kmer   the bidirectional intervals of every k-mer for k = 1..K (K = 14 at 3.1 Gbp: 5.7 GB): the first K steps of a
//          forward sweep of bwt_smem1a (bwa/bwt.c:303-321) or of bwt_seed_strategy1 (bwa/bwt.c:361-378) become
//          INDEPENDENT 16-byte lookups (one per prefix length, because bwt_smem1a records an interval at every change
//          of size) instead of a chain of K dependent FM-index steps.
//   dense SA + pac   once a forward sweep's interval has ONE occurrence, further extension is a string comparison of
//          the read against the reference at that locus: x[0] does not change while a one-row interval extends
//          (bwa/bwt.c:262-275: the other bases' sizes are zero), so the sweep ends where the read first differs from
//          the text — one SA lookup and a few 32-bit compares instead of up to ~130 dependent FM-index steps.
//
// What is NOT carried: x[1] of an emitted interval.  It is bwt_smem1a's working coordinate for FORWARD extension
// only; the backward sweep, pass 2 and mem_chain never read it (bwa/bwt.c:328-345 use x[0]; bwa/bwamem.c:163,294-309
// read x[2], x[0], info).  Backward extension therefore counts one base at two positions instead of four, and emitted
// intervals carry x1 = 0.  The exact forms of seed.cuh (EMAB_SEED_MODE 1-4) still produce x[1] and the reference's
// Occ-block count; tests compare this form with them and with the reference's intervals on (x0, x2, info).
#pragma once
#include "seed_quad.cuh"

// ---------------------------------------------------------------------------------------------
// packed interval (16 bytes): three 39-bit coordinates + a 9-bit query position.  Shared by the interval lists in
// shared memory and the k-mer table (info = 0 there).
// ---------------------------------------------------------------------------------------------
EMAB_HD uint4 intv_pack(const Intv &v)
{
	const uint64_t M39 = (1ull << 39) - 1;
	const uint64_t A = (v.x0 & M39) | (v.x2 & 0x1ffffffull) << 39;
	const uint64_t B = (v.x1 & M39) | ((v.x2 >> 25) & 0x3fffull) << 39 | (v.info & 0x1ffull) << 53;
	return make_uint4((uint32_t)A, (uint32_t)(A >> 32), (uint32_t)B, (uint32_t)(B >> 32));
}
EMAB_HD Intv intv_unpack(const uint4 &u)
{
	const uint64_t M39 = (1ull << 39) - 1;
	const uint64_t A = (uint64_t)u.y << 32 | u.x, B = (uint64_t)u.w << 32 | u.z;
	Intv v;
	v.x0 = A & M39; v.x1 = B & M39;
	v.x2 = (A >> 39) | ((B >> 39) & 0x3fffull) << 25;
	v.info = B >> 53;
	return v;
}
EMAB_HD uint64_t packed_x2(const uint4 &u) { return (uint64_t)(u.y >> 7) | (uint64_t)((u.w >> 7) & 0x3fffu) << 25; }
EMAB_HD uint4 packed_with_info(uint4 u, int info) { u.w = (u.w & 0x001fffffu) | (uint32_t)info << 21; return u; }

#define EMAB_KMER_MAX 14
// first entry of level t (1-based): 4 + 16 + ... + 4^(t-1)
EMAB_HD uint64_t kmer_level_off(int t) { return ((1ull << (2 * t)) - 4) / 3; }
EMAB_HD uint64_t kmer_total(int K) { return kmer_level_off(K + 1); }

// ---------------------------------------------------------------------------------------------
// builders (thread-scalar: the device kernels in api.cu and the host index of tests/hostsim call the same code)
// ---------------------------------------------------------------------------------------------
// one-hot block b (64 symbols) from bwa's 128-symbol block b >> 1
EMAB_HD void hot_build_block(const uint4 *bwt, uint64_t b, uint4 out[4])
{
	const uint4 *p = bwt + ((b >> 1) << 2);
	const uint4 c0 = p[0], c1 = p[1], w0 = p[2], w1 = p[3];
	const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
	uint64_t cnt[4] = {(uint64_t)c0.y << 32 | c0.x, (uint64_t)c0.w << 32 | c0.z, (uint64_t)c1.y << 32 | c1.x, (uint64_t)c1.w << 32 | c1.z};
	const int half = (int)(b & 1);
	if (half)
		for (int i = 0; i < 64; ++i) ++cnt[(w[i >> 4] >> ((~i & 15) << 1)) & 3];
	uint64_t oh[4] = {0, 0, 0, 0};
	for (int i = 0; i < 64; ++i) oh[(w[half * 4 + (i >> 4)] >> ((~i & 15) << 1)) & 3] |= 1ull << i;
	for (int c = 0; c < 4; ++c) out[c] = make_uint4((uint32_t)cnt[c], (uint32_t)(cnt[c] >> 32), (uint32_t)oh[c], (uint32_t)(oh[c] >> 32));
}

// the four children of a k-mer's interval: child b = the k-mer followed by base b (forward extension = bwt_extend on
// the complement strand, bwa/bwt.c:309-310)
EMAB_HD void kmer_children(Fm &fm, const uint4 &parent, uint4 out[4])
{
	Intv ik = intv_unpack(parent);
	ik.info = 0;
	if (ik.x2 == 0) { for (int b = 0; b < 4; ++b) out[b] = make_uint4(0, 0, 0, 0); return; }
	Intv ok[4];
	bwt_extend(fm, ik, ok, 0);
	for (int b = 0; b < 4; ++b) {
		Intv v = ok[3 - b];
		v.info = 0;
		if (v.x2 == 0) v.x0 = v.x1 = 0;
		out[b] = intv_pack(v);
	}
}
EMAB_HD uint4 kmer_level1(const DevIndex &ix, int b)
{
	Intv v;
	bwt_set_intv(ix, b, v);
	return intv_pack(v);
}
// K for an index of seq_len symbols: k-mers of the last level occur ~16-64 times on average
EMAB_HD int kmer_default_k(uint64_t seq_len)
{
	int lg = 0;
	while ((seq_len >> (2 * (lg + 1))) != 0) ++lg;   // floor(log4(seq_len))
	int k = lg - 2;
	return k < 2 ? 2 : (k > EMAB_KMER_MAX ? EMAB_KMER_MAX : k);
}

// ---------------------------------------------------------------------------------------------
// the FM step: Occ(base, .) at the two ends of an interval
// ---------------------------------------------------------------------------------------------
struct HotFm {
	const DevIndex &ix;
	unsigned sectors;   // 32-byte sectors requested (instrumentation: the unit of this form's traffic)
	// host-side work profile (tests/hostsim hs_hot_profile): memory round trips by kind, as the request loop of seed_rq.cuh
	// would take them — 0 table, 1 forward FM steps, 2 SA, 3 text (128 bases each), 4 backward rounds, 5 backward steps of four
	// entries, 6-9 the same kinds in pass 3
	unsigned prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	EMAB_HD void count(int k, unsigned n = 1)
	{
#ifndef __CUDA_ARCH__
		prof[k] += n;
#endif
	}
	EMAB_HD const DevIndex &index() const { return ix; }

	// tk = Occ(base, xa - 2), ns = Occ(base, xa - 2 + xn) - tk with bwt_2occ4's conventions (bwa/bwt.c:189-220: positions
	// are inclusive, shifted past the primary row, and (bwtint_t)-1 means "before the start")
	EMAB_HD void occ_pair(uint64_t xa, uint64_t xn, int base, bool counted, uint64_t &tk, uint64_t &ns)
	{
		const uint64_t NEG1 = ~0ull;
		const uint64_t k = xa - 1, l = xa - 1 + xn;
		const bool kv = k != NEG1, lv = l != NEG1;
		const uint64_t _k = kv ? k - (k >= ix.primary) : 0, _l = lv ? l - (l >= ix.primary) : 0;
		const uint64_t bk = _k >> 6, bl = _l >> 6;
		const uint4 ek = ldg128(ix.hot + (bk << 2) + base);
		uint4 el = ek;
		const bool other = bk != bl;
		ldg128_if(other, ix.hot + (bl << 2) + base, el);
		if (counted) sectors += 1u + (unsigned)other;
		const uint64_t mk = (2ull << (_k & 63)) - 1, ml = (2ull << (_l & 63)) - 1;
		const uint64_t ck = ((uint64_t)ek.y << 32 | ek.x) + (uint64_t)emab_popcll(((uint64_t)ek.w << 32 | ek.z) & mk);
		const uint64_t cl = ((uint64_t)el.y << 32 | el.x) + (uint64_t)emab_popcll(((uint64_t)el.w << 32 | el.z) & ml);
		tk = kv ? ck : 0;
		ns = (lv ? cl : 0) - tk;
	}
};

// ---------------------------------------------------------------------------------------------
// the text at a locus: 16 bases of the forward-reverse reference starting at p, base j at bits 30 - 2j
// ---------------------------------------------------------------------------------------------
EMAB_HD uint32_t text_word16(const DevIndex &ix, int64_t p)
{
	const int64_t L = ix.l_pac;
#ifdef __CUDA_ARCH__
	const uint32_t *pac32 = (const uint32_t *)ix.pac;
	if (p + 16 <= L) {
		const int64_t w = p >> 4;
		const int sh = (int)(p & 15) << 1;
		const uint32_t W0 = __byte_perm(pac32[w], 0, 0x0123), W1 = __byte_perm(pac32[w + 1], 0, 0x0123);
		return __funnelshift_l(W1, W0, sh);
	}
	if (p >= L && 2 * L - p >= 16) {   // reverse half: base j = 3 - F[f0 - j], f0 = 2L - 1 - p
		const int64_t g = 2 * L - 16 - p;  // f0 - 15
		const int64_t w = g >> 4;
		const int sh = (int)(g & 15) << 1;
		const uint32_t W0 = __byte_perm(pac32[w], 0, 0x0123), W1 = __byte_perm(pac32[w + 1], 0, 0x0123);
		const uint32_t f = __funnelshift_l(W1, W0, sh);
		const uint32_t y = __brev(f);
		return ~(((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1));
	}
#endif
	uint32_t r = 0;   // across the strand junction, at the end of the text, and on the host
	for (int j = 0; j < 16; ++j) {
		const int64_t pj = p + j;
		const uint32_t b = pj < 2 * L ? (uint32_t)ref_base(ix, pj) : 0u;
		r |= b << (30 - 2 * j);
	}
	return r;
}

// how many bases of seq[e .. len) equal the text from position p on; stops at an ambiguous read base and at the end of
// the text (nothing extends past it: the next symbol is the sentinel)
EMAB_HD int text_match_run(const DevIndex &ix, int64_t p, const uint8_t *seq, int e, int len, unsigned *sectors)
{
	const int64_t T = (int64_t)ix.seq_len;
	int run = 0;
	while (e < len && p < T) {
		int n = len - e < 16 ? len - e : 16;
		if (T - p < n) n = (int)(T - p);
		uint32_t w = 0;
		int nv = 0;
		for (; nv < n; ++nv) {
			const uint32_t b = seq[e + nv];
			if (b > 3) break;
			w |= b << (30 - 2 * nv);
		}
		const uint32_t x = w ^ text_word16(ix, p);
#ifdef __CUDA_ARCH__
		int mt = __clz(x) >> 1;
#else
		int mt = x ? __builtin_clz(x) >> 1 : 16;
#endif
		if (mt > nv) mt = nv;
		++*sectors;
		run += mt; e += mt; p += mt;
		if (mt < 16) break;
	}
	return run;
}

// ---------------------------------------------------------------------------------------------
// list / cooperation policies of the loops below (host forms; the device forms are in seed_launch.cuh)
// ---------------------------------------------------------------------------------------------
struct HotPtrLists {
	Intv *l[2];
	EMAB_HD void put(int which, int idx, const Intv &v) { l[which][idx] = v; }
	EMAB_HD Intv get(int which, int idx) const { return l[which][idx]; }
	EMAB_HD void emit(Intv *out, int idx, const Intv &v) { out[idx] = v; }
	EMAB_HD void put_by(int, int which, int idx, const Intv &v) { l[which][idx] = v; }
	EMAB_HD void put_packed_by(int, int which, int idx, const uint4 &u) { l[which][idx] = intv_unpack(u); }
	EMAB_HD void emit_by(int, Intv *out, int idx, const Intv &v) { out[idx] = v; }
	EMAB_HD void sync() const {}
};
struct HotSoloCoop {
	static constexpr int WIDTH = 1;
	EMAB_HD int lane() const { return 0; }
	EMAB_HD uint64_t bcast(uint64_t v, int) const { return v; }
	EMAB_HD uint32_t bcast32(uint32_t v, int) const { return v; }
	EMAB_HD uint64_t sum(uint64_t v) const { return v; }
};

// levels 1..m of the k-mer table for the bases seq[sx .. sx + m): slot s of lane q holds level s * WIDTH + q + 1
template <int WIDTH> struct KmerSlots { uint4 e[(EMAB_KMER_MAX + WIDTH - 1) / WIDTH]; };

template <class Coop>
EMAB_HD int kmer_fetch(HotFm &fm, const Coop &coop, const uint8_t *seq, int sx, int len, KmerSlots<Coop::WIDTH> &ks, bool all_levels)
{
	constexpr int W = Coop::WIDTH, S = (EMAB_KMER_MAX + W - 1) / W;
	const DevIndex &ix = fm.ix;
	const int K = ix.kmer_k;
	int m = 1;
	uint32_t code = seq[sx];
	while (m < K && sx + m < len && seq[sx + m] < 4) { code = code << 2 | seq[sx + m]; ++m; }
#pragma unroll
	for (int s = 0; s < S; ++s) {
		const int t = s * W + coop.lane() + 1;
		const bool in = t <= m, want = all_levels ? in : t == m;
		const uint64_t idx = in ? kmer_level_off(t) + (code >> (2 * (m - t))) : 0;
		ks.e[s] = make_uint4(0, 0, 0, 0);
		ldg128_if(want, ix.kmer + idx, ks.e[s]);
	}
	if (coop.lane() == 0) fm.sectors += all_levels ? (unsigned)m : 1u;
	return m;
}
// level t's entry, delivered to every lane of the group (t is uniform across the group)
template <class Coop>
EMAB_HD uint4 kmer_level(const Coop &coop, const KmerSlots<Coop::WIDTH> &ks, int t)
{
	constexpr int W = Coop::WIDTH, S = (EMAB_KMER_MAX + W - 1) / W;
	const int s = (t - 1) / W, owner = (t - 1) % W;
	uint4 e = ks.e[0];
#pragma unroll
	for (int a = 1; a < S; ++a) if (s == a) e = ks.e[a];
	e.x = coop.bcast32(e.x, owner); e.y = coop.bcast32(e.y, owner); e.z = coop.bcast32(e.z, owner); e.w = coop.bcast32(e.w, owner);
	return e;
}

// forward extension of ik by the read base whose complement is c (bwa/bwt.c:262-275 with is_back = 0), the four bases
// of the step spread over the lanes of the group (lane q counts base q) or looped by a single lane.  ALL LANES OF THE
// GROUP CALL.  tk/ns: this lane's occ_pair result for base coop.lane() (WIDTH 4) or unused (WIDTH 1).
template <class Coop>
EMAB_HD Intv hot_forward(HotFm &fm, const Coop &coop, const Intv &ik, int c, uint64_t tk, uint64_t ns)
{
	const DevIndex &ix = fm.ix;
	uint64_t tk_c, ns_c, above;
	if (Coop::WIDTH == 1) {
		tk_c = ns_c = above = 0;
		for (int b = 0; b < 4; ++b) {
			uint64_t t, n;
			fm.occ_pair(ik.x1, ik.x2, b, true, t, n);
			if (b == c) { tk_c = t; ns_c = n; }
			if (b > c) above += n;
		}
	} else {
		tk_c = coop.bcast(tk, c);
		ns_c = coop.bcast(ns, c);
		above = coop.sum(coop.lane() > c ? ns : 0);
	}
	Intv ok;
	ok.x1 = ix.L2[c] + 1 + tk_c;
	ok.x2 = ns_c;
	ok.x0 = ik.x0 + (ik.x1 <= ix.primary && ik.x1 + ik.x2 - 1 >= ix.primary) + above;
	ok.info = 0;
	return ok;
}

// ---------------------------------------------------------------------------------------------
// passes 1 + 2 (the state machine of seed_p12, seed.cuh; states and bookkeeping are the same, the FM step is occ_pair)
// ALL 32 LANES OF A WARP MUST CALL.
// ---------------------------------------------------------------------------------------------
template <class Feeder, class Lists, class Coop>
EMAB_HD void seed_p12_hot(HotFm &fm, Feeder &feed, Lists &lists, const Coop &coop)
{
	constexpr int WIDTH = Coop::WIDTH;
	const DevIndex &ix = fm.ix;
	SeedJob job;
	job.seq = nullptr; job.len = 0; job.out = nullptr; job.cap = 0; job.id = -1;
	int n = 0, ovf = 0;
	int pass = 1, x = 0, old_n = 0, k2 = 0;
	int st = SD_DONE;
	bool have_job = false, drained = false;
	int i = 0, j = 0, n_prev = 0, n_curr = 0, sx = 0, ret = 0, last_start = 0x7fffffff;
	uint64_t min_intv = 1, last_size = 0;
	bool in_p2 = false, rev = false;
	int cur = 1;
	Pass2Queue q2;
	Intv ik;
	ik.x0 = ik.x1 = ik.x2 = ik.info = 0;
	int c = 0;
	bool counted = true;
	for (;;) {
		bool req = false;
		uint64_t xa = 0, xn = 0;
		int base = 0;
		while (!req && !drained) {
			if (st == SD_DONE) {
				if (have_job) { feed.done(job, n, ovf); have_job = false; }
				if (feed.next(job)) { have_job = true; n = 0; ovf = 0; pass = 1; x = 0; st = SD_NEXT; q2.clear(); }
				else drained = true;
			}
			if (st == SD_BWD && j >= n_prev) {  // end of one backward round (bwa/bwt.c:346-348)
				if (n_curr == 0) {
					if (!in_p2) x = ret;
					st = SD_NEXT;
				} else {
					lists.sync();
					cur ^= 1;
					n_prev = n_curr; n_curr = 0;
					--i; rev = false;
					st = SD_BWD0;
				}
			}
			if (st == SD_NEXT) {
				bool start_call = false;
				if (pass == 1) {
					while (x < job.len && job.seq[x] > 3) ++x;
					if (x >= job.len) { pass = 2; old_n = n; k2 = 0; }
					else { sx = x; min_intv = 1; in_p2 = false; start_call = true; }
				} else {  // pass 2: bwa/bwamem.c:157-168
					if (q2.pop(&sx, &min_intv)) { in_p2 = true; start_call = true; }
					else if (q2.spill_from >= 0) {
						if (k2 < q2.spill_from) k2 = q2.spill_from;
						lists.sync();
						while (k2 < old_n) {
							const Intv p = job.out[k2];
							const int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
							if (end - start >= opt::split_len && p.x2 <= (uint64_t)opt::split_width) break;
							++k2;
						}
						if (k2 >= old_n) st = SD_DONE;
						else {
							const Intv p = job.out[k2++];
							sx = ((int)(p.info >> 32) + (int)(uint32_t)p.info) >> 1;
							min_intv = p.x2 + 1; in_p2 = true; start_call = true;
						}
					} else st = SD_DONE;
				}
				if (start_call) {  // bwt_smem1a(sx, min_intv): bwa/bwt.c:289-302
					n_curr = 0; last_start = 0x7fffffff;
					if (ix.kmer_k > 0) {
						// the first m <= K steps of the forward sweep from the table: level t is the interval of seq[sx .. sx + t)
						KmerSlots<WIDTH> ks;
						const int m = kmer_fetch(fm, coop, job.seq, sx, job.len, ks, true);
						fm.count(0);
						uint64_t cur_x2 = ~0ull;
						bool stopped = false;
						int t_last = 1;
#pragma unroll
						for (int t = 2; t <= EMAB_KMER_MAX; ++t) {
							if (t <= m && !stopped) {
								const int s = (t - 1) / WIDTH, owner = (t - 1) % WIDTH, sp = (t - 2) / WIDTH, ownerp = (t - 2) % WIDTH;
								if (t == 2) cur_x2 = coop.bcast(packed_x2(ks.e[0]), 0);
								const uint64_t ok_x2 = coop.bcast(packed_x2(ks.e[s]), owner);
								if (ok_x2 != cur_x2) {  // bwa/bwt.c:311-314: the previous interval is recorded
									lists.put_packed_by(ownerp, cur, n_curr++, packed_with_info(ks.e[sp], sx + t - 1));
									ret = sx + t - 1;
									if (ok_x2 < min_intv) stopped = true;
								}
								if (!stopped) { cur_x2 = ok_x2; t_last = t; }
							}
						}
						if (stopped) {
							lists.sync();
							cur ^= 1;
							n_prev = n_curr; n_curr = 0;
							i = sx - 1; rev = true;
							st = SD_BWD0;
						} else {
							ik = intv_unpack(kmer_level(coop, ks, t_last));
							ik.info = sx + t_last;
							i = sx + t_last;
							st = SD_FWD;
						}
					} else {
						bwt_set_intv(ix, job.seq[sx], ik);
						ik.info = sx + 1;
						i = sx + 1;
						st = SD_FWD;
					}
				}
			}
			if (st == SD_FWD) {
				bool close = true;
				if (i < job.len && job.seq[i] < 4) {
					if (ik.x2 == 1 && min_intv <= 1 && (ix.sa32 || ix.sa64)) {
						// one occurrence: the sweep continues exactly as far as the read equals the text there
						const uint64_t p = bwt_sa_dense(ix, ik.x0);
						unsigned sec = 1;
						const int i_before = i;
						i += text_match_run(ix, (int64_t)p + (i - sx), job.seq, i, job.len, &sec);
						fm.count(2); fm.count(3, 1 + (i - i_before) / 128);
						if (coop.lane() == 0) fm.sectors += sec;
						ik.info = i;
					} else { c = 3 - job.seq[i]; xa = ik.x1; xn = ik.x2; base = WIDTH == 1 ? 0 : coop.lane(); counted = true; req = true; close = false; }
				}
				if (close) {  // ambiguous base, end of the read, or the end of a one-occurrence sweep (bwa/bwt.c:311-321)
					lists.put(cur, n_curr++, ik); ret = (int)ik.info;
					lists.sync();
					cur ^= 1;
					n_prev = n_curr; n_curr = 0;
					i = sx - 1; rev = true;
					st = SD_BWD0;
				}
			}
			if (st == SD_BWD0) {  // a backward round starts at query position i (bwa/bwt.c:324-326)
				const int cc = i < 0 ? -1 : (job.seq[i] < 4 ? job.seq[i] : -1);
				if (cc < 0) {
					Intv p = lists.get(cur ^ 1, rev ? n_prev - 1 : 0);
					if (i + 1 < last_start) {
						p.info = (p.info & 0xffffffffull) | (uint64_t)(i + 1) << 32;
						p.x1 = 0;
						if ((int)(uint32_t)p.info - (i + 1) >= opt::min_seed_len) {
							if (!in_p2) q2.consider(p, i + 1, n);
							if (n < job.cap) lists.emit(job.out, n++, p); else ovf = 1;
						}
					}
					if (!in_p2) x = ret;
					st = SD_NEXT;
				} else { c = cc; j = 0; last_size = 0; st = SD_BWD; fm.count(4); fm.count(5, (n_prev + 3) / 4); }
			}
			if (st == SD_BWD && j < n_prev) {
				int jj = j + coop.lane();
				counted = jj < n_prev;
				jj = counted ? jj : n_prev - 1;
				ik = lists.get(cur ^ 1, rev ? n_prev - 1 - jj : jj);
				xa = ik.x0; xn = ik.x2; base = c;
				req = true;
			}
		}
		if (!EMAB_WARP_ANY(req)) break;
		if (!req) continue;
		uint64_t tk = 0, ns = 0;
		if (WIDTH > 1 || st != SD_FWD) fm.occ_pair(xa, xn, base, counted, tk, ns);   // the one convergent step
		if (st == SD_FWD) {  // bwa/bwt.c:307-315
			fm.count(1);
			const Intv ok = hot_forward(fm, coop, ik, c, tk, ns);
			bool stop = false;
			if (ok.x2 != ik.x2) {
				lists.put(cur, n_curr++, ik); ret = (int)ik.info;
				stop = ok.x2 < min_intv;
			}
			if (stop) {
				lists.sync();
				cur ^= 1;
				n_prev = n_curr; n_curr = 0;
				i = sx - 1; rev = true;
				st = SD_BWD0;
			} else {
				ik = ok;
				ik.info = i + 1;
				++i;
			}
		} else {  // SD_BWD: bwa/bwt.c:328-345 for up to WIDTH entries of the round, in list order
			Intv ok;
			ok.x0 = ix.L2[c] + 1 + tk; ok.x1 = 0; ok.x2 = ns; ok.info = ik.info;
#pragma unroll
			for (int t = 0; t < WIDTH; ++t) {
				if (j + t >= n_prev) break;
				const uint64_t ok_x2 = coop.bcast(ok.x2, t);
				if (ok_x2 < min_intv) {
					if (n_curr == 0 && i + 1 < last_start) {
						const uint64_t p_info = (coop.bcast(ik.info, t) & 0xffffffffull) | (uint64_t)(i + 1) << 32;
						if ((int)(uint32_t)p_info - (i + 1) >= opt::min_seed_len) {
							Intv p = ik;
							p.info = p_info; p.x1 = 0;
							if (!in_p2) { Intv pq; pq.x2 = coop.bcast(ik.x2, t); pq.info = p_info; q2.consider(pq, i + 1, n); }
							if (n < job.cap) lists.emit_by(t, job.out, n++, p); else ovf = 1;
						}
						last_start = i + 1;
					}
				} else if (n_curr == 0 || ok_x2 != last_size) {
					lists.put_by(t, cur, n_curr++, ok);
					last_size = ok_x2;
				}
			}
			j += WIDTH;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// pass 3: bwt_seed_strategy1 from every restart point (bwa/bwt.c:358-379, bwa/bwamem.c:170-185).  A walk starts from
// the k-mer table (no interval can be emitted before 19 bases, and K < 19), steps the FM index while its interval has
// two or more rows, and finishes against the text once it has one; an empty interval stays empty, so such a walk only
// has to find where the reference's loop would leave it.  ALL 32 LANES OF A WARP MUST CALL.
// ---------------------------------------------------------------------------------------------
template <class Feeder, class Lists, class Coop>
EMAB_HD void seed_p3_hot(HotFm &fm, Feeder &feed, Lists &lists, const Coop &coop)
{
	constexpr int WIDTH = Coop::WIDTH;
	const DevIndex &ix = fm.ix;
	SeedJob job;
	job.seq = nullptr; job.len = 0; job.out = nullptr; job.cap = 0; job.id = -1;
	int n = 0, ovf = 0, x = 0, i = 0, c = 0;
	bool have_job = false, drained = false, active = false;
	Intv ik;
	ik.x0 = ik.x1 = ik.x2 = ik.info = 0;
	for (;;) {
		bool req = false;
		while (!req && !drained) {
			if (!active) {
				while (x < job.len && job.seq[x] > 3) ++x;
				if (x >= job.len) {
					if (have_job) { feed.done(job, n, ovf); have_job = false; }
					if (feed.next(job)) { have_job = true; n = 0; ovf = 0; x = 0; }
					else drained = true;
					continue;
				}
				if (ix.kmer_k > 0) {
					KmerSlots<WIDTH> ks;
					const int m = kmer_fetch(fm, coop, job.seq, x, job.len, ks, false);
					fm.count(6);
					ik = intv_unpack(kmer_level(coop, ks, m));
					i = x + m;
				} else {
					bwt_set_intv(ix, job.seq[x], ik);
					i = x + 1;
				}
				active = true;
			}
			// the reference's loop is at index i with interval ik = seq[x .. i)  (bwa/bwt.c:363-378)
			if (i >= job.len) { x = job.len; active = false; continue; }
			if (job.seq[i] > 3) { x = i + 1; active = false; continue; }
			if (ik.x2 == 0) {  // nothing to extend: the loop runs on until an ambiguous base, 19 bases, or the end
				while (i < job.len && job.seq[i] < 4 && i - x < opt::min_seed_len) ++i;
				x = i < job.len ? i + 1 : job.len;
				active = false;
				continue;
			}
			if (ik.x2 == 1 && (ix.sa32 || ix.sa64)) {  // one occurrence: compare with the text; x[0] stays what it is
				const int64_t p = (int64_t)bwt_sa_dense(ix, ik.x0);
				fm.count(8); fm.count(9);
				if (coop.lane() == 0) fm.sectors += 2;
				bool dead = false;
				for (;;) {
					if (i >= job.len) { x = job.len; break; }
					if (job.seq[i] > 3) { x = i + 1; break; }
					const int64_t tp = p + (i - x);
					const bool mt = tp < (int64_t)ix.seq_len && ref_base(ix, tp) == job.seq[i];
					if (i - x >= opt::min_seed_len) {
						if (mt) {
							Intv o = ik;
							o.x1 = 0;
							o.info = (uint64_t)x << 32 | (uint64_t)(i + 1);
							if (n < job.cap) lists.emit(job.out, n++, o); else ovf = 1;
						}
						x = i + 1;
						break;
					}
					if (!mt) { dead = true; break; }
					++i;
				}
				if (dead) { ik.x2 = 0; ++i; continue; }   // the failed step consumed index i
				active = false;
				continue;
			}
			c = 3 - job.seq[i];
			req = true;
		}
		if (!EMAB_WARP_ANY(req)) break;
		if (!req) continue;
		uint64_t tk = 0, ns = 0;
		if (WIDTH > 1) fm.occ_pair(ik.x1, ik.x2, coop.lane(), true, tk, ns);
		fm.count(7);
		Intv ok = hot_forward(fm, coop, ik, c, tk, ns);
		if (ok.x2 < (uint64_t)opt::max_mem_intv && i - x >= opt::min_seed_len) {  // bwa/bwt.c:366-375
			if (ok.x2 > 0) {
				ok.x1 = 0;
				ok.info = (uint64_t)x << 32 | (uint64_t)(i + 1);
				if (n < job.cap) lists.emit(job.out, n++, ok); else ovf = 1;
			}
			x = i + 1; active = false;
		} else { ik = ok; ++i; }
	}
}

// one read, all passes, on the host (tests/hostsim)
EMAB_HD int collect_intv_hot(const DevIndex &ix, int len, const uint8_t *seq, Intv *mem, int mem_cap, Intv *buf0, Intv *buf1, int *overflow, unsigned *sectors,
                             unsigned *prof = nullptr)
{
	Intv p3[EMAB_P3_CAP];
	OneReadFeeder f12{{seq, len, mem, mem_cap, 0}, false, 0, 0}, f3{{seq, len, p3, EMAB_P3_CAP, 0}, false, 0, 0};
	HotFm fm{ix, 0};
	HotPtrLists lists{{buf0, buf1}};
	seed_p12_hot(fm, f12, lists, HotSoloCoop());
	seed_p3_hot(fm, f3, lists, HotSoloCoop());
	if (sectors) *sectors = fm.sectors;
	if (prof) for (int k = 0; k < 10; ++k) prof[k] = fm.prof[k];
	const int n = finish_intv(mem, f12.n, p3, f3.n, mem_cap);
	if (f12.ovf || f3.ovf || n < 0) { *overflow = 1; return f12.n; }
	return n;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device form: FOUR LANES PER READ.  Backward rounds: four list entries per step, one per lane (each lane one occ_pair
// of its own entry for the round's base).  Forward steps: the four bases of the step, one per lane — the same
// instruction sequence, so lanes in either kind of step share the convergent occ_pair.  Lists in shared memory.
// ---------------------------------------------------------------------------------------------
struct HotQuadLists {
	uint4 *sm;           // [2][SEEDQ_CAP] (+1 pad)
	Intv *slab;          // [2][slab_len]: entries past SEEDQ_CAP
	int slab_len, q;
	unsigned qmask;
	__device__ __forceinline__ void put(int which, int idx, const Intv &v)
	{
		if (idx < SEEDQ_CAP) sm[which * SEEDQ_CAP + idx] = intv_pack(v);
		else slab[(size_t)which * slab_len + idx] = v;
	}
	__device__ __forceinline__ Intv get(int which, int idx) const
	{
		if (idx < SEEDQ_CAP) return intv_unpack(sm[which * SEEDQ_CAP + idx]);
		return slab[(size_t)which * slab_len + idx];
	}
	__device__ __forceinline__ void emit(Intv *out, int idx, const Intv &v)
	{
		const uint64_t comp = q == 0 ? v.x0 : (q == 1 ? v.x1 : (q == 2 ? v.x2 : v.info));
		((uint64_t *)(out + idx))[q] = comp;
	}
	__device__ __forceinline__ void put_by(int owner, int which, int idx, const Intv &v) { if (q == owner) put(which, idx, v); }
	__device__ __forceinline__ void put_packed_by(int owner, int which, int idx, const uint4 &u)
	{
		if (q != owner) return;
		if (idx < SEEDQ_CAP) sm[which * SEEDQ_CAP + idx] = u;
		else slab[(size_t)which * slab_len + idx] = intv_unpack(u);
	}
	__device__ __forceinline__ void emit_by(int owner, Intv *out, int idx, const Intv &v) { if (q == owner) out[idx] = v; }
	__device__ __forceinline__ void sync() const { __syncwarp(qmask); }
};

struct HotQuadCoop {
	static constexpr int WIDTH = 4;
	int q;
	unsigned qmask;
	__device__ __forceinline__ int lane() const { return q; }
	__device__ __forceinline__ int src(int t) const { return ((threadIdx.x & 31) & ~3) + t; }
	__device__ __forceinline__ uint64_t bcast(uint64_t v, int t) const { return __shfl_sync(qmask, v, src(t)); }
	__device__ __forceinline__ uint32_t bcast32(uint32_t v, int t) const { return __shfl_sync(qmask, v, src(t)); }
	__device__ __forceinline__ uint64_t sum(uint64_t v) const
	{
		v += __shfl_xor_sync(qmask, v, 1);
		v += __shfl_xor_sync(qmask, v, 2);
		return v;
	}
};

__device__ __forceinline__ void seed_hot_quads(const DevIndex &ix, const SeedBatch &b, uint4 *smem)
{
	const int quad = threadIdx.x >> 2, q = threadIdx.x & 3;
	const unsigned qmask = 0xfu << ((threadIdx.x & 31) & ~3);
	const size_t gquad = (size_t)blockIdx.x * (blockDim.x >> 2) + quad;
	HotFm fm{ix, 0};
	HotQuadLists lists{smem + quad * SEEDQ_STRIDE, b.scratch + gquad * 2 * b.scratch_len, b.scratch_len, q, qmask};
	HotQuadCoop coop{q, qmask};
	QuadFeeder f12{b, 0, q, qmask}, f3{b, 1, q, qmask};
	const unsigned gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	// every fourth warp starts on the pass-3 queue; each warp moves to the other queue when its own is dry
	const bool p3_first = (gwarp & 3) == 3;
#pragma unroll 1
	for (int r = 0; r < 2; ++r) {
		if ((r == 0) == p3_first) seed_p3_hot(fm, f3, lists, coop);
		else seed_p12_hot(fm, f12, lists, coop);
	}
	unsigned sectors = fm.sectors;
	for (int d = 16; d; d >>= 1) sectors += __shfl_xor_sync(0xffffffffu, sectors, d);
	if ((threadIdx.x & 31) == 0 && sectors) atomicAdd(b.touches, (unsigned long long)sectors);
}
#endif
