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

// mem_collect_intv for one read.  Returns the number of intervals (sorted by info).
EMAB_HD int collect_intv(Fm &fm, int len, const uint8_t *seq, Intv *mem, int mem_cap,
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
