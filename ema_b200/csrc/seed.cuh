// SMEM seeding: mem_collect_intv (bwa/bwamem.c:140-188) = three passes of bwt_smem1a
// (bwa/bwt.c:289-351) / bwt_seed_strategy1 (bwa/bwt.c:358-379) over the HBM-resident FM index.
//
// Mapping: one thread per read.  Every step of every pass is one dependent bwt_extend (one or two
// random 64-byte block reads), so throughput comes from having tens of thousands of independent
// reads in flight, not from intra-read parallelism; a warp therefore keeps 32 independent
// pointer chases outstanding.  Per-read interval lists live in a global scratch slab.
#pragma once
#include "fmindex.cuh"

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
// mem_collect_intv flattened into ONE loop whose body performs exactly one bwt_extend.
//
// Why: with one thread per read, the nested form above puts every lane of a warp in a different
// loop nest (forward sweep, backward sweep, pass 3, ...), so the ~250-instruction bwt_extend is
// issued once per distinct program point: ncu showed 8.9 of 32 lanes active per instruction.
// Here every lane reaches the same bwt_extend call site each iteration; what differs between lanes
// is only the cheap bookkeeping before ("which interval/base next") and after ("what to do with
// the result").  Semantics are unchanged: tests/hostsim checks this against the nested form and
// against the reference on every read of the fixtures.
// ---------------------------------------------------------------------------------------------
enum { SD_NEXT = 0, SD_FWD = 1, SD_BWD = 2, SD_P3 = 3, SD_DONE = 4 };

EMAB_HD int collect_intv(Fm &fm, int len, const uint8_t *seq, Intv *mem, int mem_cap,
                         Intv *buf0, Intv *buf1, int *overflow)
{
	int n = 0;            // intervals kept so far (all passes)
	int pass = 1, x = 0;  // pass 1 / 3 cursor
	int old_n = 0, k2 = 0;  // pass 2 cursor over the pass-1 intervals
	int st = SD_NEXT;
	// state of the bwt_smem1a call in progress
	int i = 0, j = 0, n_prev = 0, n_curr = 0, m0 = 0, m = 0, sx = 0, ret = 0;
	uint64_t min_intv = 1, last_start = 0, last_size = 0;
	bool have_last = false, in_p2 = false;
	Intv *prev = buf0, *curr = buf1;
	Intv ik;
	ik.x0 = ik.x1 = ik.x2 = ik.info = 0;
	int c = 0, back = 0;
	for (;;) {
		// ---- bookkeeping until the next bwt_extend is known (or everything is done)
		bool req = false;
		while (!req && st != SD_DONE) {
			if (st == SD_NEXT) {
				bool start_call = false;
				if (pass == 1) {
					while (x < len && seq[x] > 3) ++x;
					if (x >= len) { pass = 2; old_n = n; k2 = 0; }
					else { sx = x; min_intv = 1; in_p2 = false; start_call = true; }
				} else if (pass == 2) {  // bwa/bwamem.c:157-168
					while (k2 < old_n) {
						const Intv p = mem[k2];
						const int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
						if (end - start >= opt::split_len && p.x2 <= (uint64_t)opt::split_width) break;
						++k2;
					}
					if (k2 >= old_n) { pass = 3; x = 0; }
					else {
						const Intv p = mem[k2++];
						sx = ((int)(p.info >> 32) + (int)(uint32_t)p.info) >> 1;
						min_intv = p.x2 + 1; in_p2 = true; start_call = true;
					}
				} else {  // pass 3: bwt_seed_strategy1 from x (bwa/bwamem.c:170-185)
					while (x < len && seq[x] > 3) ++x;
					if (x >= len) st = SD_DONE;
					else { bwt_set_intv(fm.ix, seq[x], ik); i = x + 1; st = SD_P3; }
				}
				if (start_call) {  // bwt_smem1a(sx, min_intv): bwa/bwt.c:289-302
					bwt_set_intv(fm.ix, seq[sx], ik);
					ik.info = sx + 1;
					i = sx + 1; n_curr = 0; m0 = m = n; have_last = false;
					st = SD_FWD;
				}
			} else if (st == SD_FWD) {
				bool fwd_done = false;
				if (i < len && seq[i] < 4) { c = 3 - seq[i]; back = 0; req = true; }
				else { curr[n_curr++] = ik; fwd_done = true; }  // ambiguous base, or i == len (bwa/bwt.c:316-321)
				if (fwd_done) {
					reverse_intvs(curr, n_curr);
					ret = (int)curr[0].info;
					{ Intv *t = curr; curr = prev; prev = t; }
					n_prev = n_curr; n_curr = 0; last_size = 0;
					i = sx - 1; j = 0;
					st = SD_BWD;
				}
			} else if (st == SD_BWD) {
				if (j == n_prev) {  // end of one backward round (bwa/bwt.c:346-348)
					if (n_curr == 0) {  // the call is over: reverse + keep seeds >= min_seed_len
						if (m > mem_cap) m = mem_cap;
						reverse_intvs(mem + m0, m - m0);
						int k = m0;
						for (int jj = m0; jj < m; ++jj) {
							const Intv p = mem[jj];
							const int slen = (int)(uint32_t)p.info - (int)(p.info >> 32);
							if (slen >= opt::min_seed_len) mem[k++] = p;
						}
						n = k;
						if (!in_p2) x = ret;
						st = SD_NEXT;
					} else {
						{ Intv *t = curr; curr = prev; prev = t; }
						n_prev = n_curr; n_curr = 0; last_size = 0;
						--i; j = 0;
					}
				} else {
					const int cc = i < 0 ? -1 : (seq[i] < 4 ? seq[i] : -1);
					if (cc < 0) {  // nothing can extend: the first interval may be emitted (bwa/bwt.c:331-338)
						Intv p = prev[j];
						if (n_curr == 0 && (!have_last || (uint64_t)(i + 1) < last_start)) {
							p.info |= (uint64_t)(i + 1) << 32;
							if (m < mem_cap) mem[m] = p; else *overflow = 1;
							++m;
							last_start = (uint64_t)(i + 1); have_last = true;
						}
						++j;
					} else { ik = prev[j]; c = cc; back = 1; req = true; }
				}
			} else {  // SD_P3
				if (i < len && seq[i] < 4) { c = 3 - seq[i]; back = 0; req = true; }
				else { x = i < len ? i + 1 : len; st = SD_NEXT; }  // bwa/bwt.c:376-378
			}
		}
		// ---- the one convergent step.  On the device the warp votes here every iteration: the vote is
		// the reconvergence point that brings all lanes to the bwt_extend below together (without it the
		// lanes leave the bookkeeping loop one by one and each runs the extend code on its own), and
		// lanes whose read is finished idle until the whole warp is.  ALL 32 LANES MUST CALL collect_intv.
#ifdef __CUDA_ARCH__
		if (!__any_sync(0xffffffffu, st != SD_DONE)) break;
		if (st == SD_DONE) continue;
#else
		if (st == SD_DONE) break;
#endif
		const Intv ok0 = bwt_extend1(fm, ik, c, back);
		Intv ok = ok0;
		// ---- consume
		if (st == SD_FWD) {  // bwa/bwt.c:307-315
			bool stop = false;
			if (ok.x2 != ik.x2) {
				curr[n_curr++] = ik;
				if (ok.x2 < min_intv) stop = true;
			}
			if (stop) {
				reverse_intvs(curr, n_curr);
				ret = (int)curr[0].info;
				{ Intv *t = curr; curr = prev; prev = t; }
				n_prev = n_curr; n_curr = 0; last_size = 0;
				i = sx - 1; j = 0;
				st = SD_BWD;
			} else {
				ok.info = i + 1;
				ik = ok;
				++i;
			}
		} else if (st == SD_BWD) {  // bwa/bwt.c:328-345
			if (ok.x2 < min_intv) {
				if (n_curr == 0 && (!have_last || (uint64_t)(i + 1) < last_start)) {
					Intv p = ik;
					p.info |= (uint64_t)(i + 1) << 32;
					if (m < mem_cap) mem[m] = p; else *overflow = 1;
					++m;
					last_start = (uint64_t)(i + 1); have_last = true;
				}
			} else if (n_curr == 0 || ok.x2 != last_size) {
				ok.info = ik.info;
				curr[n_curr++] = ok;
				last_size = ok.x2;
			}
			++j;
		} else {  // SD_P3 (bwa/bwt.c:366-375)
			if (ok.x2 < (uint64_t)opt::max_mem_intv && i - x >= opt::min_seed_len) {
				if (ok.x2 > 0) {
					ok.info = (uint64_t)x << 32 | (uint64_t)(i + 1);
					if (n < mem_cap) mem[n++] = ok; else *overflow = 1;
				}
				x = i + 1;
				st = SD_NEXT;
			} else { ik = ok; ++i; }
		}
	}
	// ks_introsort by info (bwa/bwamem.c:187): equal keys are identical intervals, so any
	// comparison sort gives the reference's byte-identical result.
	for (int a = 1; a < n; ++a) {
		Intv t = mem[a];
		int b = a;
		while (b > 0 && mem[b - 1].info > t.info) { mem[b] = mem[b - 1]; --b; }
		mem[b] = t;
	}
	return n;
}
