// SMEM seeding: mem_collect_intv (bwa/bwamem.c:140-188) = three passes of bwt_smem1a
// (bwa/bwt.c:289-351) / bwt_seed_strategy1 (bwa/bwt.c:358-379) over the HBM-resident FM index.
//
// Mapping: one thread per read.  Every step of every pass is one dependent bwt_extend (one or two
// random 64-byte block reads), so throughput comes from having tens of thousands of independent
// reads in flight, not from intra-read parallelism; a warp therefore keeps 32 independent
// pointer chases outstanding.  Per-read interval lists live in a global scratch slab.
#pragma once
#include "fmindex.cuh"

// number of SA occurrences mem_chain enumerates for one interval (bwa/bwamem.c:304-305)
EMAB_HD int intv_occ_count(uint64_t x2)
{
	uint64_t step = x2 > (uint64_t)opt::max_occ ? x2 / opt::max_occ : 1;
	uint64_t cnt = (x2 + step - 1) / step;
	return (int)(cnt < (uint64_t)opt::max_occ ? cnt : (uint64_t)opt::max_occ);
}

struct SeedScratch {  // per read: two interval lists of (len+1) entries for bwt_smem1a's prev/curr
	Intv *a;          // [n_slots][2][EMAB_MAX_READ_LEN + 1]
};

EMAB_HD void reverse_intvs(Intv *p, int n)
{
	for (int j = 0; j < n >> 1; ++j) {
		Intv t = p[n - 1 - j];
		p[n - 1 - j] = p[j];
		p[j] = t;
	}
}

// bwt_smem1a with max_intv = 0 (the only form on the path, bwa/bwt.c:353-356).
// Appends to mem[*n_mem..]; returns the next x.  Entries shorter than min_seed_len are dropped by
// the caller (bwa/bwamem.c:150-155,165-167) — done here so the output stays compact.
EMAB_HD int smem1(Fm &fm, int len, const uint8_t *q, int x, uint64_t min_intv,
                            Intv *mem, int *n_mem, int mem_cap, Intv *buf0, Intv *buf1, int *overflow)
{
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	Intv *prev = buf0, *curr = buf1;
	int n_prev, n_curr = 0;
	Intv ik;
	bwt_set_intv(fm.ix, q[x], ik);
	ik.info = x + 1;
	int i;
	for (i = x + 1; i < len; ++i) {  // forward search
		if (q[i] < 4) {
			int c = 3 - q[i];
			Intv ok = bwt_extend1(fm, ik, c, 0);
			if (ok.x2 != ik.x2) {
				curr[n_curr++] = ik;
				if (ok.x2 < min_intv) break;
			}
			ok.info = i + 1;
			ik = ok;
		} else {
			curr[n_curr++] = ik;
			break;
		}
	}
	if (i == len) curr[n_curr++] = ik;
	reverse_intvs(curr, n_curr);
	int ret = (int)curr[0].info;
	{ Intv *t = curr; curr = prev; prev = t; }
	n_prev = n_curr;

	int m0 = *n_mem;  // this call's segment of mem starts here
	int m = m0;
	uint64_t last_start = 0;
	bool have_last = false;
	for (i = x - 1; i >= -1; --i) {  // backward search for MEMs
		int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		n_curr = 0;
		uint64_t last_size = 0;
		for (int j = 0; j < n_prev; ++j) {
			Intv p = prev[j];
			Intv ok;
			ok.x2 = 0;
			if (c >= 0) ok = bwt_extend1(fm, p, c, 1);
			if (c < 0 || ok.x2 < min_intv) {
				if (n_curr == 0) {
					if (!have_last || (uint64_t)(i + 1) < last_start) {
						p.info |= (uint64_t)(i + 1) << 32;
						if (m < mem_cap) mem[m] = p; else *overflow = 1;
						++m;
						last_start = (uint64_t)(i + 1);
						have_last = true;
					}
				}
			} else if (n_curr == 0 || ok.x2 != last_size) {
				ok.info = p.info;
				curr[n_curr++] = ok;
				last_size = ok.x2;
			}
		}
		if (n_curr == 0) break;
		{ Intv *t = curr; curr = prev; prev = t; }
		n_prev = n_curr;
	}
	if (m > mem_cap) m = mem_cap;
	reverse_intvs(mem + m0, m - m0);
	// keep only seeds >= min_seed_len, preserving order
	int k = m0;
	for (int j = m0; j < m; ++j) {
		Intv p = mem[j];
		int slen = (int)(uint32_t)p.info - (int)(p.info >> 32);
		if (slen >= opt::min_seed_len) mem[k++] = p;
	}
	*n_mem = k;
	return ret;
}

// bwt_seed_strategy1 (bwa/bwt.c:358-379)
EMAB_HD int seed_strategy1(Fm &fm, int len, const uint8_t *q, int x, int min_len, uint64_t max_intv, Intv *out)
{
	out->x0 = out->x1 = out->x2 = out->info = 0;
	if (q[x] > 3) return x + 1;
	Intv ik;
	bwt_set_intv(fm.ix, q[x], ik);
	for (int i = x + 1; i < len; ++i) {
		if (q[i] < 4) {
			int c = 3 - q[i];
			Intv ok = bwt_extend1(fm, ik, c, 0);
			if (ok.x2 < max_intv && i - x >= min_len) {
				*out = ok;
				out->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
				return i + 1;
			}
			ik = ok;
		} else return i + 1;
	}
	return len;
}

// mem_collect_intv for one read, written as the reference writes it (nested loops).  Kept as the
// readable restatement and as the CPU cross-check of collect_intv below (tests/hostsim); the device
// kernel uses the flattened form.
EMAB_HD int collect_intv_nested(Fm &fm, int len, const uint8_t *seq, Intv *mem, int mem_cap,
                                   Intv *buf0, Intv *buf1, int *overflow)
{
	int n = 0, x = 0;
	while (x < len) {  // pass 1: all SMEMs
		if (seq[x] < 4) x = smem1(fm, len, seq, x, 1, mem, &n, mem_cap, buf0, buf1, overflow);
		else ++x;
	}
	int old_n = n;
	for (int k = 0; k < old_n; ++k) {  // pass 2: re-seed inside long, rare SMEMs
		Intv p = mem[k];
		int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
		if (end - start < opt::split_len || p.x2 > (uint64_t)opt::split_width) continue;
		smem1(fm, len, seq, (start + end) >> 1, p.x2 + 1, mem, &n, mem_cap, buf0, buf1, overflow);
	}
	x = 0;
	while (x < len) {  // pass 3: LAST-like
		if (seq[x] < 4) {
			Intv m;
			x = seed_strategy1(fm, len, seq, x, opt::min_seed_len, opt::max_mem_intv, &m);
			if (m.x2 > 0) {
				if (n < mem_cap) mem[n++] = m; else *overflow = 1;
			}
		} else ++x;
	}
	// ks_introsort by info (bwa/bwamem.c:187): equal keys are identical intervals, so any
	// comparison sort gives the reference's byte-identical result.
	for (int i = 1; i < n; ++i) {
		Intv t = mem[i];
		int j = i;
		while (j > 0 && mem[j - 1].info > t.info) { mem[j] = mem[j - 1]; --j; }
		mem[j] = t;
	}
	return n;
}


// ---------------------------------------------------------------------------------------------
// mem_collect_intv for the device: flattened into loops whose body performs exactly ONE bwt_extend.
//
// Why: with one thread per read, the nested form above puts every lane of a warp in a different
// loop nest (forward sweep, backward sweep, pass 3, ...), so the ~200-instruction bwt_extend is
// issued once per distinct program point (ncu: 8.9 of 32 lanes active per instruction).  Here every
// lane reaches the same bwt_extend call site each iteration; what differs between lanes is only the
// cheap bookkeeping before ("which interval/base next") and after ("what to do with the result").
//
// The work of one read is split where the reference's data flow allows it:
//   * passes 1+2 (bwt_smem1a from every x, then re-seeding inside long rare SMEMs) are one state
//     machine (P12): pass 2 consumes pass 1's intervals;
//   * pass 3 (bwt_seed_strategy1, bwa/bwamem.c:170-185) reads nothing the other passes write, so it is
//     its own, perfectly uniform loop (P3) — on the device a different warp, which doubles the number
//     of independent dependent-load chains in flight per read;
//   * the final ks_introsort by info (bwa/bwamem.c:187) runs once both lists exist (finish_intv).
// Because of that final sort — equal keys are identical intervals, so any comparison sort gives the
// reference's byte-identical list — P12 neither reverses each call's intervals (bwa/bwt.c:349) nor
// compacts them afterwards: an interval is tested against min_seed_len (bwa/bwamem.c:150-155,165-167)
// when it is emitted, and the first backward round indexes the forward list back to front instead
// of reversing it (bwa/bwt.c:322).
//
// Lanes are persistent: a lane that finishes a read takes the next one from its Feeder (an atomic
// counter on the device), so a warp's lanes stay busy until the queue is empty instead of idling
// until the warp's slowest read is done.  Semantics are unchanged: tests/hostsim checks the composed
// form against the nested one and against the reference on every read of the fixtures, including the
// count of Occ-block loads.
// ---------------------------------------------------------------------------------------------
enum { SD_NEXT = 0, SD_FWD = 1, SD_BWD0 = 2, SD_BWD = 3, SD_DONE = 4 };

struct SeedJob {  // one read as the seeding loops see it
	const uint8_t *seq;
	int len;
	Intv *out;   // where this pass's intervals go
	int cap;
	int id;      // read index (feeder's business)
};

#ifdef __CUDA_ARCH__
#define EMAB_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define EMAB_WARP_ANY(p) (p)
#endif

// Passes 1 and 2.  feed.next(job) hands out the next read (false = queue empty); feed.done(job, n, ovf)
// receives the number of intervals written to job.out.  buf0/buf1: this lane's prev/curr lists,
// (longest read + 1) entries each.  ALL 32 LANES OF A WARP MUST CALL (the loop votes).
// Policies of the two loops below.
//   FmT   the FM-index step: fm.extend1(ik, c, is_back) = the interval of base c after bwt_extend (bwa/bwt.c:262-275).
//         ThreadFm computes it in one thread (fmindex.cuh); QuadFm (device) spreads one step over four lanes.
//   Lists bwt_smem1a's prev/curr interval lists (bwa/bwt.c:295-297): lists.put(which, idx, iv) / lists.get(which, idx),
//         which in {0, 1}; lists.emit(out, idx, iv) stores an output interval.  PtrLists are two arrays in memory;
//         QuadLists (device) keep them packed in shared memory.
struct ThreadFm {
	Fm &fm;
	EMAB_HD Intv extend1(const Intv &ik, int c, int is_back, bool counted = true)
	{
		const unsigned before = fm.touches;
		const Intv r = bwt_extend1(fm, ik, c, is_back);
		if (!counted) fm.touches = before;
		return r;
	}
	EMAB_HD const DevIndex &index() const { return fm.ix; }
};
struct PtrLists {
	Intv *l[2];
	EMAB_HD void put(int which, int idx, const Intv &v) { l[which][idx] = v; }
	EMAB_HD Intv get(int which, int idx) const { return l[which][idx]; }
	EMAB_HD void emit(Intv *out, int idx, const Intv &v) { out[idx] = v; }
	EMAB_HD void put_by(int, int which, int idx, const Intv &v) { l[which][idx] = v; }
	EMAB_HD void emit_by(int, Intv *out, int idx, const Intv &v) { out[idx] = v; }
	EMAB_HD void sync() const {}
};
//   Coop  how many lanes work on one read.  The entries of one backward round of bwt_smem1a are extended independently of
//         each other (bwa/bwt.c:328-345: only the bookkeeping after each bwt_extend is sequential), so a group of WIDTH
//         lanes takes WIDTH entries per step, one per lane, and shares the few values the bookkeeping needs.  SoloCoop is
//         one lane per read (WIDTH 1); QuadCoop (device) four.
struct SoloCoop {
	static constexpr int WIDTH = 1;
	EMAB_HD int lane() const { return 0; }
	EMAB_HD uint64_t bcast(uint64_t v, int) const { return v; }
};

// pass 2 re-seeds inside long, rare SMEMs of pass 1 (bwa/bwamem.c:157-168).  Whether an interval qualifies is known
// when pass 1 emits it, so the (position, min_intv) pairs are queued in a register then instead of being searched for
// in the output list afterwards (a chain of dependent global loads run by a few lanes while the warp waits).  The
// order of the pass-2 calls does not matter: each is independent and the final sort orders their output.
struct Pass2Queue {
	uint64_t bits = 0;     // up to 5 entries of 12 bits: sx (8) | min_intv (4); the count in the top 4 bits
	int spill_from = -1;   // >= 0: the queue was full; intervals emitted from this output index on are scanned in memory
	EMAB_HD int count() const { return (int)(bits >> 60); }
	EMAB_HD void clear() { bits = 0; spill_from = -1; }
	EMAB_HD void consider(const Intv &p, int start, int out_idx)
	{
		const int end = (int)(uint32_t)p.info;
		if (end - start < opt::split_len || p.x2 > (uint64_t)opt::split_width) return;
		const int n = count();
		if (n == 5 || spill_from >= 0) { if (spill_from < 0) spill_from = out_idx; return; }
		const uint64_t e = (uint64_t)((start + end) >> 1) | (p.x2 + 1) << 8;
		bits = (bits & 0x0fffffffffffffffull) | e << (12 * n) | (uint64_t)(n + 1) << 60;
	}
	EMAB_HD bool pop(int *sx, uint64_t *min_intv)
	{
		const int n = count();
		if (n == 0) return false;
		*sx = (int)(bits & 0xff); *min_intv = (bits >> 8) & 0xf;
		bits = ((bits & 0x0fffffffffffffffull) >> 12) | (uint64_t)(n - 1) << 60;
		return true;
	}
};

template <class Feeder, class FmT, class Lists, class Coop>
EMAB_HD void seed_p12(FmT &fm, Feeder &feed, Lists &lists, const Coop &coop)
{
	constexpr int WIDTH = Coop::WIDTH;
	SeedJob job;
	job.seq = nullptr; job.len = 0; job.out = nullptr; job.cap = 0; job.id = -1;
	int n = 0, ovf = 0;   // intervals kept so far for this read
	int pass = 1, x = 0, old_n = 0, k2 = 0;
	int st = SD_DONE;
	bool have_job = false, drained = false;
	// state of the bwt_smem1a call in progress
	int i = 0, j = 0, n_prev = 0, n_curr = 0, sx = 0, ret = 0, last_start = 0x7fffffff;
	uint64_t min_intv = 1, last_size = 0;
	bool in_p2 = false, rev = false;
	int cur = 1;           // which list is bwt_smem1a's `curr`; the other one is `prev`
	Pass2Queue q2;
	Intv ik;
	ik.x0 = ik.x1 = ik.x2 = ik.info = 0;
	int c = 0, back = 0;
	bool counted = true;   // does this lane's next bwt_extend count as an Occ-block load of the reference's
	for (;;) {
		// ---- bookkeeping until the next bwt_extend is known (or the queue is empty)
		bool req = false;
		while (!req && !drained) {
			if (st == SD_DONE) {  // take the next read
				if (have_job) { feed.done(job, n, ovf); have_job = false; }
				if (feed.next(job)) { have_job = true; n = 0; ovf = 0; pass = 1; x = 0; st = SD_NEXT; q2.clear(); }
				else drained = true;
			}
			if (st == SD_BWD && j >= n_prev) {  // end of one backward round (bwa/bwt.c:346-348)
				if (n_curr == 0) {  // the call is over
					if (!in_p2) x = ret;
					st = SD_NEXT;
				} else {
					lists.sync();   // entries of the next round were stored by their owner lanes
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
					else if (q2.spill_from >= 0) {  // more than five candidates: the rest are looked up in the output list
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
					bwt_set_intv(fm.index(), job.seq[sx], ik);
					ik.info = sx + 1;
					i = sx + 1; n_curr = 0; last_start = 0x7fffffff;
					st = SD_FWD;
				}
			}
			if (st == SD_FWD) {
				if (i < job.len && job.seq[i] < 4) { c = 3 - job.seq[i]; back = 0; req = true; }
				else {  // ambiguous base, or i == len (bwa/bwt.c:316-321)
					lists.put(cur, n_curr++, ik); ret = (int)ik.info;
					cur ^= 1;
					n_prev = n_curr; n_curr = 0;
					i = sx - 1; rev = true;
					st = SD_BWD0;
				}
			}
			if (st == SD_BWD0) {  // a backward round starts at query position i (bwa/bwt.c:324-326)
				const int cc = i < 0 ? -1 : (job.seq[i] < 4 ? job.seq[i] : -1);
				if (cc < 0) {
					// nothing can extend: only the first (longest) interval may be emitted, after which the round
					// leaves curr empty and the call is over (bwa/bwt.c:331-338,346)
					Intv p = lists.get(cur ^ 1, rev ? n_prev - 1 : 0);
					if (i + 1 < last_start) {
						p.info |= (uint64_t)(i + 1) << 32;
						if ((int)(uint32_t)p.info - (i + 1) >= opt::min_seed_len) {
							if (!in_p2) q2.consider(p, i + 1, n);
							if (n < job.cap) lists.emit(job.out, n++, p); else ovf = 1;
						}
					}
					if (!in_p2) x = ret;
					st = SD_NEXT;
				} else { c = cc; back = 1; j = 0; last_size = 0; st = SD_BWD; }
			}
			if (st == SD_BWD && j < n_prev) {  // this lane's entry of the round: j + lane (the last one again past the end)
				int jj = j + coop.lane();
				counted = jj < n_prev;
				jj = counted ? jj : n_prev - 1;
				ik = lists.get(cur ^ 1, rev ? n_prev - 1 - jj : jj);
				req = true;
			}
		}
		// ---- the one convergent step.  On the device the warp votes here every iteration: the vote is
		// the reconvergence point that brings all lanes to the bwt_extend below together.
		if (!EMAB_WARP_ANY(req)) break;
		if (!req) continue;
		if (st == SD_FWD) counted = coop.lane() == 0;   // the lanes of a group walk forward together: one block load to count
		Intv ok = fm.extend1(ik, c, back, counted);
		// ---- consume
		if (st == SD_FWD) {  // bwa/bwt.c:307-315
			bool stop = false;
			if (ok.x2 != ik.x2) {
				lists.put(cur, n_curr++, ik); ret = (int)ik.info;
				stop = ok.x2 < min_intv;
			}
			if (stop) {
				cur ^= 1;
				n_prev = n_curr; n_curr = 0;
				i = sx - 1; rev = true;
				st = SD_BWD0;
			} else {
				ok.info = i + 1;
				ik = ok;
				++i;
			}
		} else {  // SD_BWD: bwa/bwt.c:328-345 for up to WIDTH entries of the round, in list order
#pragma unroll
			for (int t = 0; t < WIDTH; ++t) {
				if (j + t >= n_prev) break;
				const uint64_t ok_x2 = coop.bcast(ok.x2, t);
				if (ok_x2 < min_intv) {
					if (n_curr == 0 && i + 1 < last_start) {
						const uint64_t p_info = coop.bcast(ik.info, t) | (uint64_t)(i + 1) << 32;
						if ((int)(uint32_t)p_info - (i + 1) >= opt::min_seed_len) {
							Intv p = ik;
							p.info = p_info;
							if (!in_p2) { Intv pq; pq.x2 = coop.bcast(ik.x2, t); pq.info = p_info; q2.consider(pq, i + 1, n); }
							if (n < job.cap) lists.emit_by(t, job.out, n++, p); else ovf = 1;
						}
						last_start = i + 1;
					}
				} else if (n_curr == 0 || ok_x2 != last_size) {
					Intv v = ok;
					v.info = ik.info;
					lists.put_by(t, cur, n_curr++, v);
					last_size = ok_x2;
				}
			}
			j += WIDTH;
		}
	}
}

// Pass 3: bwt_seed_strategy1 from every restart point (bwa/bwt.c:358-379, bwa/bwamem.c:170-185).
// At most len / (min_seed_len + 1) + 1 intervals per read.  ALL 32 LANES OF A WARP MUST CALL.
template <class Feeder, class FmT, class Lists>
EMAB_HD void seed_p3(FmT &fm, Feeder &feed, Lists &lists)
{
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
				if (x >= job.len) {  // this read is finished (or there is none yet)
					if (have_job) { feed.done(job, n, ovf); have_job = false; }
					if (feed.next(job)) { have_job = true; n = 0; ovf = 0; x = 0; }
					else drained = true;
					continue;
				}
				bwt_set_intv(fm.index(), job.seq[x], ik);
				i = x + 1; active = true;
			}
			if (i < job.len && job.seq[i] < 4) { c = 3 - job.seq[i]; req = true; }
			else { x = i < job.len ? i + 1 : job.len; active = false; }  // bwa/bwt.c:376-378
		}
		if (!EMAB_WARP_ANY(req)) break;
		if (!req) continue;
		Intv ok = fm.extend1(ik, c, 0);
		if (ok.x2 < (uint64_t)opt::max_mem_intv && i - x >= opt::min_seed_len) {  // bwa/bwt.c:366-375
			if (ok.x2 > 0) {
				ok.info = (uint64_t)x << 32 | (uint64_t)(i + 1);
				if (n < job.cap) lists.emit(job.out, n++, ok); else ovf = 1;
			}
			x = i + 1; active = false;
		} else { ik = ok; ++i; }
	}
}

#define EMAB_P3_CAP 16  // pass 3 emits at most EMAB_MAX_READ_LEN / (min_seed_len + 1) + 1 = 13 intervals

// Appends the pass-3 intervals to the pass-1/2 list and sorts by info (bwa/bwamem.c:187).  Returns the
// total, or -1 if it does not fit `cap`.
EMAB_HD int finish_intv(Intv *mem, int n12, const Intv *p3, int n3, int cap)
{
	if (n12 + n3 > cap) return -1;
	for (int k = 0; k < n3; ++k) mem[n12 + k] = p3[k];
	const int n = n12 + n3;
	for (int a = 1; a < n; ++a) {
		Intv t = mem[a];
		int b = a;
		while (b > 0 && mem[b - 1].info > t.info) { mem[b] = mem[b - 1]; --b; }
		mem[b] = t;
	}
	return n;
}

// one read, all passes: the host-side composition (tests/hostsim) of what the device runs as three roles
struct OneReadFeeder {
	SeedJob j; bool given; int n, ovf;
	EMAB_HD bool next(SeedJob &o) { if (given) return false; given = true; o = j; return true; }
	EMAB_HD void done(const SeedJob &, int n_, int ovf_) { n = n_; ovf = ovf_; }
};

EMAB_HD int collect_intv(Fm &fm, int len, const uint8_t *seq, Intv *mem, int mem_cap,
                         Intv *buf0, Intv *buf1, int *overflow)
{
	Intv p3[EMAB_P3_CAP];
	OneReadFeeder f12{{seq, len, mem, mem_cap, 0}, false, 0, 0}, f3{{seq, len, p3, EMAB_P3_CAP, 0}, false, 0, 0};
	ThreadFm tfm{fm};
	PtrLists lists{{buf0, buf1}};
	seed_p12(tfm, f12, lists, SoloCoop());
	seed_p3(tfm, f3, lists);
	const int n = finish_intv(mem, f12.n, p3, f3.n, mem_cap);
	if (f12.ovf || f3.ovf || n < 0) { *overflow = 1; return f12.n; }
	return n;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device roles.  A launch is a persistent grid of warps; warps with (warp index % 4 == 3) start on the
// pass-3 queue and move to the pass-1/2 queue when it is empty, the others the other way round.
// ---------------------------------------------------------------------------------------------
struct SeedBatch {
	int n_reads;
	const uint8_t *seq;
	const int64_t *off;
	Intv *intv;        // [n_reads][max_intv]: pass 1/2 output, completed in place by seed_finish
	int max_intv;
	Intv *p3;          // [n_reads][EMAB_P3_CAP]
	int32_t *n12, *n3; // [n_reads]
	Intv *scratch;     // [n_lanes][2][scratch_len]: per-LANE prev/curr lists
	int scratch_len;
	int *err;          // set to 3 on overflow
	unsigned long long *queue;  // [0] pass-1/2 read counter, [1] pass-3 read counter
	unsigned long long *touches;
};

struct QueueFeeder {
	const SeedBatch &b;
	int which;  // 0: pass 1/2, 1: pass 3
	__device__ bool next(SeedJob &o)
	{
		const unsigned long long r = atomicAdd(&b.queue[which], 1ull);
		if (r >= (unsigned long long)b.n_reads) return false;
		o.id = (int)r;
		o.seq = b.seq + b.off[r];
		o.len = (int)(b.off[r + 1] - b.off[r]);
		if (which == 0) { o.out = b.intv + (size_t)r * b.max_intv; o.cap = b.max_intv; }
		else { o.out = b.p3 + (size_t)r * EMAB_P3_CAP; o.cap = EMAB_P3_CAP; }
		return true;
	}
	__device__ void done(const SeedJob &j, int n, int ovf)
	{
		(which == 0 ? b.n12 : b.n3)[j.id] = n;
		if (ovf) *b.err = 3;
	}
};

__device__ __forceinline__ void seed_warp(const DevIndex &ix, const SeedBatch &b)
{
	const int gt = blockIdx.x * blockDim.x + threadIdx.x;
	Fm fm{ix, 0};
	Intv *buf0 = b.scratch + (size_t)gt * 2 * b.scratch_len;
	QueueFeeder f12{b, 0}, f3{b, 1};
	ThreadFm tfm{fm};
	PtrLists lists{{buf0, buf0 + b.scratch_len}};
	if (((threadIdx.x >> 5) & 3) == 3) { seed_p3(tfm, f3, lists); seed_p12(tfm, f12, lists, SoloCoop()); }
	else { seed_p12(tfm, f12, lists, SoloCoop()); seed_p3(tfm, f3, lists); }
	unsigned touches = fm.touches;
	for (int d = 16; d; d >>= 1) touches += __shfl_xor_sync(0xffffffffu, touches, d);
	if ((threadIdx.x & 31) == 0 && touches) atomicAdd(b.touches, (unsigned long long)touches);
}
#endif
