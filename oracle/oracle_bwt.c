/*
 * oracle/oracle_bwt.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * Plain-C restatement of the FM-index routines on the `ema align` path, reading the files written
 * by `bwa index` directly:
 *   orc_index_load   <- bwt_restore_bwt / bwt_restore_sa / bns_restore / .pac  (bwa/bwt.c:421-462, bwa/bntseq.c:97-170, bwa/bwa.c:271-316)
 *   orc_occ4         <- bwt_occ4            (bwa/bwt.c:169-186)
 *   orc_extend       <- bwt_extend          (bwa/bwt.c:262-275)
 *   orc_sa           <- bwt_sa / bwt_invPsi (bwa/bwt.c:53-59,86-96)
 *   orc_smem1        <- bwt_smem1a          (bwa/bwt.c:289-351), max_intv = 0
 *   orc_collect_intv <- mem_collect_intv    (bwa/bwamem.c:140-188) incl. bwt_seed_strategy1 (bwa/bwt.c:358-379)
 * Pinned against the compiled reference by tests/test_oracle_vs_ref.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static void *slurp(const char *path, size_t *size)
{
	FILE *f = fopen(path, "rb");
	void *buf;
	long n;
	if (!f) return 0;
	fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET);
	buf = malloc(n + 1);
	if (fread(buf, 1, n, f) != (size_t)n) { free(buf); fclose(f); return 0; }
	fclose(f);
	*size = n;
	return buf;
}

orc_index_t *orc_index_load(const char *prefix)
{
	char path[4096];
	size_t sz;
	orc_index_t *ix = (orc_index_t *)calloc(1, sizeof(*ix));
	uint64_t *raw;
	/* .bwt: primary, L2[1..4], then the interleaved Occ/BWT body (bwa/bwt.c:443-462) */
	snprintf(path, sizeof path, "%s.bwt", prefix);
	raw = (uint64_t *)slurp(path, &sz);
	if (!raw) { free(ix); return 0; }
	ix->primary = raw[0];
	ix->L2[0] = 0; memcpy(&ix->L2[1], raw + 1, 32);
	ix->seq_len = ix->L2[4];
	ix->bwt = (const uint32_t *)(raw + 5);
	ix->bwt_size = (sz - 40) >> 2;
	/* .sa: primary, 4 skipped words, sa_intv, seq_len, then samples 1..n_sa-1 (bwa/bwt.c:421-441) */
	snprintf(path, sizeof path, "%s.sa", prefix);
	raw = (uint64_t *)slurp(path, &sz);
	if (!raw) { free(ix); return 0; }
	ix->sa_intv = (int)raw[5];
	ix->n_sa = (ix->seq_len + ix->sa_intv) / ix->sa_intv;
	{
		uint64_t *sa = (uint64_t *)malloc(ix->n_sa * 8);
		sa[0] = (uint64_t)-1;
		memcpy(sa + 1, raw + 7, (ix->n_sa - 1) * 8);
		ix->sa = sa;
		free(raw);
	}
	/* .ann (bwa/bntseq.c:109-137) */
	{
		FILE *f;
		long long xx; int n_seqs, i; unsigned seed;
		int64_t *off; int32_t *len;
		char name[8192];
		snprintf(path, sizeof path, "%s.ann", prefix);
		if (!(f = fopen(path, "r"))) { free(ix); return 0; }
		if (fscanf(f, "%lld%d%u", &xx, &n_seqs, &seed) != 3) { fclose(f); free(ix); return 0; }
		ix->l_pac = xx; ix->n_seqs = n_seqs;
		off = (int64_t *)malloc(n_seqs * 8); len = (int32_t *)malloc(n_seqs * 4);
		for (i = 0; i < n_seqs; ++i) {
			unsigned gi; int c, l, na;
			if (fscanf(f, "%u%s", &gi, name) != 2) break;
			while ((c = fgetc(f)) != '\n' && c != EOF) ;
			if (fscanf(f, "%lld%d%d", &xx, &l, &na) != 3) break;
			off[i] = xx; len[i] = l;
		}
		fclose(f);
		ix->ann_offset = off; ix->ann_len = len;
	}
	snprintf(path, sizeof path, "%s.pac", prefix);
	ix->pac = (const uint8_t *)slurp(path, &sz);
	return ix;
}

void orc_index_free(orc_index_t *ix)
{
	if (!ix) return;
	free((void *)((const uint64_t *)ix->bwt - 5));
	free((void *)ix->sa); free((void *)ix->pac); free((void *)ix->ann_offset); free((void *)ix->ann_len);
	free(ix);
}

/* number of symbols equal to c among the first n (0..16) symbols of a 32-bit BWT word (MSB first) */
static int word_count(uint32_t w, int n, int c)
{
	int i, k = 0;
	for (i = 0; i < n; ++i) k += (int)((w >> (30 - 2 * i)) & 3) == c;
	return k;
}

/* bwa/bwt.c:169-186 */
void orc_occ4(orc_index_t *ix, uint64_t k, uint64_t cnt[4])
{
	const uint32_t *blk;
	int c, j, n;
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= ix->primary);                       /* '$' is not stored */
	blk = ix->bwt + ((k >> 7) << 4);               /* bwa/bwt.h:75 */
	ix->occ_touches++;
	n = (int)(k & 127) + 1;                        /* symbols of this block inside B[0..k] */
	for (c = 0; c < 4; ++c) {
		uint64_t v = (uint64_t)blk[2 * c] | (uint64_t)blk[2 * c + 1] << 32;
		for (j = 0; j < 8 && j * 16 < n; ++j)
			v += word_count(blk[8 + j], n - j * 16 < 16 ? n - j * 16 : 16, c);
		cnt[c] = v;
	}
}

/* bwa/bwt.c:262-275 */
void orc_extend(orc_index_t *ix, const orc_intv_t *ik, orc_intv_t ok[4], int is_back)
{
	uint64_t tk[4], tl[4];
	int a = !is_back, b = is_back, c;
	orc_occ4(ix, ik->x[a] - 1, tk);
	orc_occ4(ix, ik->x[a] - 1 + ik->x[2], tl);
	{   /* bwt_2occ4 touches one block when both positions share it (bwa/bwt.c:194-219) */
		uint64_t k = ik->x[a] - 1, l = k + ik->x[2];
		if (k != (uint64_t)-1 && l != (uint64_t)-1) {
			uint64_t k2 = k - (k >= ix->primary), l2 = l - (l >= ix->primary);
			if ((k2 >> 7) == (l2 >> 7)) ix->occ_touches--;
		}
	}
	for (c = 0; c < 4; ++c) {
		ok[c].x[a] = ix->L2[c] + 1 + tk[c];
		ok[c].x[2] = tl[c] - tk[c];
	}
	ok[3].x[b] = ik->x[b] + (ik->x[a] <= ix->primary && ik->x[a] + ik->x[2] - 1 >= ix->primary);
	ok[2].x[b] = ok[3].x[b] + ok[3].x[2];
	ok[1].x[b] = ok[2].x[b] + ok[2].x[2];
	ok[0].x[b] = ok[1].x[b] + ok[1].x[2];
}

/* bwa/bwt.c:53-59 */
static uint64_t inv_psi(orc_index_t *ix, uint64_t k)
{
	uint64_t x, cnt[4];
	int c;
	if (k == ix->primary) return 0;
	x = k - (k > ix->primary);
	c = (ix->bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)] >> ((~x & 15) << 1)) & 3;   /* bwt_B0, bwa/bwt.h:80 */
	if (k == ix->seq_len) return ix->L2[c] + (ix->L2[c + 1] - ix->L2[c]);              /* bwt_occ, bwa/bwt.c:112 */
	orc_occ4(ix, k, cnt);
	return ix->L2[c] + cnt[c];
}

/* bwa/bwt.c:86-96 */
uint64_t orc_sa(orc_index_t *ix, uint64_t k)
{
	uint64_t steps = 0, mask = (uint64_t)ix->sa_intv - 1;
	while (k & mask) { ++steps; k = inv_psi(ix, k); }
	return steps + ix->sa[k / ix->sa_intv];
}

static void set_intv(orc_index_t *ix, int c, orc_intv_t *ik)
{   /* bwa/bwt.h:82 */
	ik->x[0] = ix->L2[c] + 1;
	ik->x[2] = ix->L2[c + 1] - ix->L2[c];
	ik->x[1] = ix->L2[3 - c] + 1;
	ik->info = 0;
}

static void reverse(orc_intv_t *a, int n)
{
	int j;
	for (j = 0; j < n / 2; ++j) { orc_intv_t t = a[j]; a[j] = a[n - 1 - j]; a[n - 1 - j] = t; }
}

/* bwa/bwt.c:289-351 with max_intv = 0.  mem must hold len+1 entries. */
int orc_smem1(orc_index_t *ix, int len, const uint8_t *q, int x, int min_intv, orc_intv_t *mem, int *n_mem)
{
	orc_intv_t *prev, *curr, *swap, ik, ok[4];
	int i, j, n_prev, n_curr = 0, n_out = 0, ret;
	*n_mem = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	prev = (orc_intv_t *)malloc((len + 1) * sizeof(*prev));
	curr = (orc_intv_t *)malloc((len + 1) * sizeof(*curr));
	set_intv(ix, q[x], &ik);
	ik.info = x + 1;
	for (i = x + 1; i < len; ++i) {                /* forward: record the interval each time its size changes */
		if (q[i] > 3) { curr[n_curr++] = ik; break; }
		orc_extend(ix, &ik, ok, 0);
		{
			orc_intv_t nx = ok[3 - q[i]];
			if (nx.x[2] != ik.x[2]) {
				curr[n_curr++] = ik;
				if (nx.x[2] < (uint64_t)min_intv) break;
			}
			nx.info = i + 1;
			ik = nx;
		}
	}
	if (i == len) curr[n_curr++] = ik;
	reverse(curr, n_curr);                         /* longest match first */
	ret = (int)curr[0].info;
	swap = curr; curr = prev; prev = swap; n_prev = n_curr;
	for (i = x - 1; i >= -1; --i) {                /* backward: prune */
		int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		n_curr = 0;
		for (j = 0; j < n_prev; ++j) {
			orc_intv_t *p = &prev[j];
			int dead = 1;
			if (c >= 0) { orc_extend(ix, p, ok, 1); dead = ok[c].x[2] < (uint64_t)min_intv; }
			if (dead) {
				if (n_curr == 0 && (n_out == 0 || (uint64_t)(i + 1) < (mem[n_out - 1].info >> 32))) {
					mem[n_out] = *p;
					mem[n_out].info |= (uint64_t)(i + 1) << 32;
					++n_out;
				}
			} else if (n_curr == 0 || ok[c].x[2] != curr[n_curr - 1].x[2]) {
				ok[c].info = p->info;
				curr[n_curr++] = ok[c];
			}
		}
		if (n_curr == 0) break;
		swap = curr; curr = prev; prev = swap; n_prev = n_curr;
	}
	reverse(mem, n_out);
	*n_mem = n_out;
	free(prev); free(curr);
	return ret;
}

/* bwa/bwt.c:358-379 */
static int seed_strategy1(orc_index_t *ix, int len, const uint8_t *q, int x, int min_len, int max_intv, orc_intv_t *mem)
{
	orc_intv_t ik, ok[4];
	int i;
	memset(mem, 0, sizeof(*mem));
	if (q[x] > 3) return x + 1;
	set_intv(ix, q[x], &ik);
	for (i = x + 1; i < len; ++i) {
		if (q[i] > 3) return i + 1;
		orc_extend(ix, &ik, ok, 0);
		ik = ok[3 - q[i]];
		if (ik.x[2] < (uint64_t)max_intv && i - x >= min_len) {
			*mem = ik;
			mem->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
			return i + 1;
		}
	}
	return len;
}

static int cmp_info(const void *a, const void *b)
{
	uint64_t x = ((const orc_intv_t *)a)->info, y = ((const orc_intv_t *)b)->info;
	return (x > y) - (x < y);
}

/* bwa/bwamem.c:140-188 with min_seed_len 19, split_factor 1.5, split_width 10, max_mem_intv 20.
 * Equal `info` keys denote identical intervals, so qsort reproduces ks_introsort's output. */
int orc_collect_intv(orc_index_t *ix, int len, const uint8_t *seq, orc_intv_t *out, int max)
{
	orc_intv_t *tmp = (orc_intv_t *)malloc((len + 1) * sizeof(*tmp));
	int n = 0, x = 0, k, i, nt, old_n;
	const int min_seed_len = 19, split_len = (int)(19 * 1.5 + .499), split_width = 10, max_mem_intv = 20;
	while (x < len) {
		if (seq[x] > 3) { ++x; continue; }
		x = orc_smem1(ix, len, seq, x, 1, tmp, &nt);
		for (i = 0; i < nt; ++i)
			if ((int)(uint32_t)tmp[i].info - (int)(tmp[i].info >> 32) >= min_seed_len && n < max) out[n++] = tmp[i];
	}
	old_n = n;
	for (k = 0; k < old_n; ++k) {
		int start = (int)(out[k].info >> 32), end = (int)(uint32_t)out[k].info;
		if (end - start < split_len || out[k].x[2] > (uint64_t)split_width) continue;
		orc_smem1(ix, len, seq, (start + end) >> 1, (int)out[k].x[2] + 1, tmp, &nt);
		for (i = 0; i < nt; ++i)
			if ((int)(uint32_t)tmp[i].info - (int)(tmp[i].info >> 32) >= min_seed_len && n < max) out[n++] = tmp[i];
	}
	x = 0;
	while (x < len) {
		orc_intv_t m;
		if (seq[x] > 3) { ++x; continue; }
		x = seed_strategy1(ix, len, seq, x, min_seed_len, max_mem_intv, &m);
		if (m.x[2] > 0 && n < max) out[n++] = m;
	}
	qsort(out, n, sizeof(*out), cmp_info);
	free(tmp);
	return n;
}

/* flat ctypes helpers */
int orc_collect_intv_flat(orc_index_t *ix, int len, const uint8_t *seq, int64_t *out, int max)
{
	orc_intv_t *buf = (orc_intv_t *)malloc((size_t)max * sizeof(*buf));
	int i, n = orc_collect_intv(ix, len, seq, buf, max);
	for (i = 0; i < n; ++i) { out[i*4] = buf[i].x[0]; out[i*4+1] = buf[i].x[1]; out[i*4+2] = buf[i].x[2]; out[i*4+3] = buf[i].info; }
	free(buf);
	return n;
}

void orc_sa_batch(orc_index_t *ix, int n, const int64_t *k, int64_t *out)
{
	int i;
	for (i = 0; i < n; ++i) out[i] = (int64_t)orc_sa(ix, (uint64_t)k[i]);
}

int64_t orc_touches(orc_index_t *ix, int reset) { int64_t t = ix->occ_touches; if (reset) ix->occ_touches = 0; return t; }
void orc_index_info(orc_index_t *ix, int64_t *out)
{
	int i;
	out[0] = ix->l_pac; out[1] = ix->n_seqs; out[2] = ix->primary; out[3] = ix->seq_len;
	for (i = 0; i < 5; ++i) out[4 + i] = ix->L2[i];
	out[9] = ix->sa_intv; out[10] = ix->n_sa; out[11] = ix->bwt_size;
}
