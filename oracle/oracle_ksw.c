/*
 * oracle/oracle_ksw.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * Plain-C restatement of the three Smith-Waterman routines on the `ema align` path:
 *   orc_ksw_extend2  <- ksw_extend2  (bwa/ksw.c:416-515)
 *   orc_ksw_global2  <- ksw_global2  (bwa/ksw.c:540-642)
 *   orc_ksw_align2   <- ksw_align2 -> ksw_u8 / ksw_i16 (bwa/ksw.c:122-253,255-370,379-401)
 * Pinned against the compiled reference by tests/test_oracle_vs_ref.py.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define NEG_INF (-0x40000000) /* MINUS_INF, bwa/ksw.c:526 */

static int imax2(int a, int b) { return a > b ? a : b; }
static int imin2(int a, int b) { return a < b ? a : b; }

/* bwa/bwa.c:136-146 */
void orc_fill_scmat(int a, int b, int8_t mat[25])
{
	int i, j, k = 0;
	for (i = 0; i < 4; ++i) {
		for (j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b);
		mat[k++] = -1;
	}
	for (j = 0; j < 5; ++j) mat[k++] = -1;
}

/* Band limit shared by extension: bwa/ksw.c:435-443 */
static int clamp_band(int w, int qlen, int m, const int8_t *mat, int end_bonus, int o_ins, int e_ins, int o_del, int e_del)
{
	int i, best = 0, lim;
	for (i = 0; i < m * m; ++i) best = imax2(best, mat[i]);
	lim = (int)((double)(qlen * best + end_bonus - o_ins) / e_ins + 1.);
	w = imin2(w, imax2(lim, 1));
	lim = (int)((double)(qlen * best + end_bonus - o_del) / e_del + 1.);
	w = imin2(w, imax2(lim, 1));
	return w;
}

int orc_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                    int *qle, int *tle, int *gtle, int *gscore_, int *max_off_, int64_t *cells)
{
	/* Hd[j] holds H(i-1, j-1) and Ev[j] holds E(i, j) on entry to row i (the eh_t pair of bwa/ksw.c:412-414) */
	int32_t *Hd = (int32_t *)calloc(qlen + 1, sizeof(int32_t));
	int32_t *Ev = (int32_t *)calloc(qlen + 1, sizeof(int32_t));
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int i, j, lo = 0, hi = qlen;
	int best = h0, best_i = -1, best_j = -1, g_i = -1, g = -1, off = 0;
	int64_t visited = 0;

	/* first row: bwa/ksw.c:431-433 */
	Hd[0] = h0;
	if (qlen >= 1) Hd[1] = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && Hd[j - 1] > e_ins; ++j) Hd[j] = Hd[j - 1] - e_ins;
	w = clamp_band(w, qlen, m, mat, end_bonus, o_ins, e_ins, o_del, e_del);

	for (i = 0; i < tlen; ++i) { /* bwa/ksw.c:448-507 */
		const int8_t *srow = mat + target[i] * m;
		int f = 0, left, rowmax = 0, rowarg = -1;
		if (lo < i - w) lo = i - w;
		if (hi > i + w + 1) hi = i + w + 1;
		if (hi > qlen) hi = qlen;
		if (lo == 0) { left = h0 - (o_del + e_del * (i + 1)); if (left < 0) left = 0; }
		else left = 0;
		visited += hi > lo ? hi - lo : 0;
		for (j = lo; j < hi; ++j) {
			int diag = Hd[j], e = Ev[j], M, h, t;
			Hd[j] = left;                                   /* H(i, j-1) becomes the next row's diagonal */
			M = diag ? diag + srow[query[j]] : 0;           /* :469 — no restart from a dead cell */
			h = imax2(imax2(M, e), f);
			left = h;
			if (!(rowmax > h)) rowarg = j;                  /* :473 — the last j wins ties */
			rowmax = imax2(rowmax, h);
			t = imax2(M - oe_del, 0);
			Ev[j] = imax2(e - e_del, t);                    /* E(i+1, j), opened from M only */
			t = imax2(M - oe_ins, 0);
			f = imax2(f - e_ins, t);                        /* F(i, j+1), opened from M only */
		}
		Hd[hi] = left; Ev[hi] = 0;
		if (j == qlen) {                                    /* :486-489 — later rows win ties */
			if (!(g > left)) g_i = i;
			g = imax2(g, left);
		}
		if (rowmax == 0) break;
		if (rowmax > best) {
			int d = rowarg - i; if (d < 0) d = -d;
			best = rowmax; best_i = i; best_j = rowarg;
			off = imax2(off, d);
		} else if (zdrop > 0) {                             /* :494-500 */
			int di = i - best_i, dj = rowarg - best_j;
			if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
			else { if (best - rowmax - (dj - di) * e_ins > zdrop) break; }
		}
		for (j = lo; j < hi && Hd[j] == 0 && Ev[j] == 0; ++j) ;   /* :502-505 */
		lo = j;
		for (j = hi; j >= lo && Hd[j] == 0 && Ev[j] == 0; --j) ;
		hi = j + 2 < qlen ? j + 2 : qlen;
	}
	free(Hd); free(Ev);
	if (qle) *qle = best_j + 1;
	if (tle) *tle = best_i + 1;
	if (gtle) *gtle = g_i + 1;
	if (gscore_) *gscore_ = g;
	if (max_off_) *max_off_ = off;
	if (cells) *cells += visited;
	return best;
}

int orc_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar_, uint32_t *cigar_out, int max_cigar,
                    int64_t *cells)
{
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	const int want_bt = n_cigar_ && cigar_out;
	int ncol = qlen < 2 * w + 1 ? qlen : 2 * w + 1;           /* bwa/ksw.c:548 */
	int32_t *Hd = (int32_t *)malloc((qlen + 1) * sizeof(int32_t));
	int32_t *Ev = (int32_t *)malloc((qlen + 1) * sizeof(int32_t));
	uint8_t *dir = want_bt ? (uint8_t *)malloc((size_t)ncol * tlen + 1) : 0;
	int i, j, score;
	int64_t visited = 0;
	if (n_cigar_) *n_cigar_ = 0;

	Hd[0] = 0; Ev[0] = NEG_INF;                               /* :558-561 */
	for (j = 1; j <= qlen && j <= w; ++j) { Hd[j] = -(o_ins + e_ins * j); Ev[j] = NEG_INF; }
	for (; j <= qlen; ++j) Hd[j] = Ev[j] = NEG_INF;

	for (i = 0; i < tlen; ++i) {                              /* :563-622 */
		const int8_t *srow = mat + target[i] * m;
		int lo = i > w ? i - w : 0;
		int hi = i + w + 1 < qlen ? i + w + 1 : qlen;
		int32_t f = NEG_INF, left = lo == 0 ? -(o_del + e_del * (i + 1)) : NEG_INF;
		uint8_t *drow = want_bt ? dir + (size_t)i * ncol : 0;
		visited += hi > lo ? hi - lo : 0;
		for (j = lo; j < hi; ++j) {
			int32_t M = Hd[j] + srow[query[j]], e = Ev[j], h, t;
			uint8_t d;
			Hd[j] = left;
			d = M >= e ? 0 : 1;            /* M preferred over E, E over F (>=) */
			h = M >= e ? M : e;
			d = h >= f ? d : 2;
			h = h >= f ? h : f;
			left = h;
			t = M - oe_del; e -= e_del;
			if (e > t) d |= 1 << 2; else e = t;   /* continuing a gap needs strict > */
			Ev[j] = e;
			t = M - oe_ins; f -= e_ins;
			if (f > t) d |= 2 << 4; else f = t;
			if (drow) drow[j - lo] = d;
		}
		Hd[hi] = left; Ev[hi] = NEG_INF;
	}
	score = Hd[qlen];
	if (want_bt) {                                            /* :624-638 */
		uint32_t *rev = (uint32_t *)malloc((size_t)(qlen + tlen + 2) * sizeof(uint32_t));
		int n = 0, state = 0, k;
		i = tlen - 1;
		k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
#define PUSH(op, len) do { if (n && (rev[n-1] & 0xf) == (uint32_t)(op)) rev[n-1] += (uint32_t)(len) << 4; else rev[n++] = (uint32_t)(len) << 4 | (op); } while (0)
		while (i >= 0 && k >= 0) {
			int lo = i > w ? i - w : 0;
			state = dir[(size_t)i * ncol + (k - lo)] >> (state << 1) & 3;
			if (state == 0) { PUSH(0, 1); --i; --k; }
			else if (state == 1) { PUSH(2, 1); --i; }
			else { PUSH(1, 1); --k; }
		}
		if (i >= 0) PUSH(2, i + 1);
		if (k >= 0) PUSH(1, k + 1);
#undef PUSH
		*n_cigar_ = n;
		for (j = 0; j < n && j < max_cigar; ++j) cigar_out[j] = rev[n - 1 - j];
		free(rev);
	}
	free(Hd); free(Ev); free(dir);
	if (cells) *cells += visited;
	return score;
}

/*
 * One pass of the striped local SW (ksw_u8 / ksw_i16), restated as a plain row-by-row DP.
 * The striped kernels pad the query to a multiple of 16 (u8) or 8 (i16) lanes with symbols that
 * score 0 against everything (bwa/ksw.c:95-113); the padding columns take part in the row maxima
 * that feed the second-best bookkeeping, so they are modelled here too.  H is exact because the
 * lazy-F loop (bwa/ksw.c:201-211) converges to the true F and an insertion directly followed by a
 * deletion is never on an optimal path with these penalties.
 */
typedef struct { int score, te, qe, score2, te2; } pass_t;

static pass_t local_pass(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                         int oe_del, int e_del, int oe_ins, int e_ins, int lanes, int is_u8, int minsc, int endsc,
                         int want_sub, int64_t *cells)
{
	int slen = (qlen + lanes - 1) / lanes, qpad = slen * lanes;
	int *H = (int *)calloc(qpad + 1, sizeof(int)), *Hn = (int *)calloc(qpad + 1, sizeof(int));
	int *E = (int *)calloc(qpad + 1, sizeof(int)), *Hbest = (int *)calloc(qpad + 1, sizeof(int));
	int *brow = 0, *bval = 0, nb = 0;
	int i, j, gmax = 0, te = -1, shift = 0, mx = 0;
	pass_t r;
	if (is_u8) { /* shift = -(min score), bwa/ksw.c:85-91 */
		int a, mn = 127;
		for (a = 0; a < m * m; ++a) { if (mat[a] < mn) mn = mat[a]; }
		shift = -mn;
	}
	{ int a; for (a = 0; a < m * m; ++a) if (mat[a] > mx) mx = mat[a]; }
	brow = (int *)malloc((tlen + 1) * sizeof(int)); bval = (int *)malloc((tlen + 1) * sizeof(int));
	for (i = 0; i < tlen; ++i) {
		const int8_t *srow = mat + target[i] * m;
		int f = 0, imax = 0;
		/* H[] is indexed by query position + 1; H[0] = 0 is the column before the query */
		for (j = 0; j < qpad; ++j) {
			int s = j < qlen ? srow[query[j]] : 0;
			int h = H[j] + s, e = E[j], t;
			if (h < 0) h = 0;
			if (h < e) h = e;
			if (h < f) h = f;
			Hn[j + 1] = h;
			if (h > imax) imax = h;
			t = h - oe_del; if (t < 0) t = 0;
			e -= e_del; if (e < t) e = t;
			E[j] = e;
			t = h - oe_ins; if (t < 0) t = 0;
			f -= e_ins; if (f < t) f = t;
		}
		if (cells) *cells += qlen;
		if (imax >= minsc) {                                   /* bwa/ksw.c:215-223 */
			if (nb == 0 || brow[nb - 1] + 1 != i) { brow[nb] = i; bval[nb] = imax; ++nb; }
			else if (bval[nb - 1] < imax) { brow[nb - 1] = i; bval[nb - 1] = imax; }
		}
		if (imax > gmax) {                                     /* :224-229 */
			gmax = imax; te = i;
			memcpy(Hbest, Hn, (qpad + 1) * sizeof(int));
			if ((is_u8 && gmax + shift >= 255) || gmax >= endsc) break;
		}
		{ int *t2 = H; H = Hn; Hn = t2; }
	}
	r.score = (is_u8 && gmax + shift >= 255) ? 255 : gmax;
	r.te = te; r.qe = -1; r.score2 = -1; r.te2 = -1;
	if (!(is_u8 && r.score == 255)) {                          /* :234-250 */
		int best = -1;
		for (j = 0; j < qpad; ++j)
			if (Hbest[j + 1] > best) { best = Hbest[j + 1]; r.qe = j; }   /* ascending j: smallest index at the max */
		if (want_sub && nb) {
			int k = (r.score + mx - 1) / mx, lo = te - k, hi = te + k;
			for (i = 0; i < nb; ++i)
				if ((brow[i] < lo || brow[i] > hi) && bval[i] > r.score2) { r.score2 = bval[i]; r.te2 = brow[i]; }
		}
	}
	free(H); free(Hn); free(E); free(Hbest); free(brow); free(bval);
	return r;
}

void orc_ksw_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int minsc, int use_u8, int out[7], int64_t *cells)
{
	int lanes = use_u8 ? 16 : 8, i;
	pass_t r = local_pass(qlen, query, tlen, target, m, mat, o_del + e_del, e_del, o_ins + e_ins, e_ins,
	                      lanes, use_u8, minsc, 0x10000, 1, cells);
	out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2; out[5] = -1; out[6] = -1;
	if (r.score < minsc) return;                               /* bwa/ksw.c:392 */
	{
		int ql = r.qe + 1, tl = r.te + 1;
		uint8_t *rq = (uint8_t *)malloc(ql), *rt = (uint8_t *)malloc(tl);
		pass_t rr;
		for (i = 0; i < ql; ++i) rq[i] = query[ql - 1 - i];
		for (i = 0; i < tl; ++i) rt[i] = target[tl - 1 - i];
		/* NB the reference passes the full tlen to the second pass (bwa/ksw.c:395): the reversed prefix
		 * is followed by the untouched rest of the target. */
		{
			uint8_t *tt = (uint8_t *)malloc(tlen);
			memcpy(tt, rt, tl);
			memcpy(tt + tl, target + tl, tlen - tl);
			rr = local_pass(ql, rq, tlen, tt, m, mat, o_del + e_del, e_del, o_ins + e_ins, e_ins,
			                lanes, use_u8, 0x10000, r.score, 0, cells);
			free(tt);
		}
		if (r.score == rr.score) { out[5] = r.te - rr.te; out[6] = r.qe - rr.qe; }
		free(rq); free(rt);
	}
}

/* ---- batched forms ------------------------------------------------------------------------ */
void orc_extend_batch(int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                      const int32_t *h0, int w, int end_bonus, int zdrop, int32_t *out, int64_t *cells, int n_threads)
{
	int8_t mat[25];
	int64_t total = 0;
	int i;
	orc_fill_scmat(1, 4, mat);
	#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64) reduction(+:total)
	for (i = 0; i < n; ++i) {
		int qle, tle, gtle, gscore, max_off;
		int64_t c = 0;
		int sc = orc_ksw_extend2((int)(qoff[i + 1] - qoff[i]), q + qoff[i], (int)(toff[i + 1] - toff[i]), t + toff[i], 5, mat,
		                         6, 1, 6, 1, w, end_bonus, zdrop, h0[i], &qle, &tle, &gtle, &gscore, &max_off, &c);
		out[i * 6] = sc; out[i * 6 + 1] = qle; out[i * 6 + 2] = tle; out[i * 6 + 3] = gtle; out[i * 6 + 4] = gscore; out[i * 6 + 5] = max_off;
		total += c;
	}
	if (cells) *cells = total;
}

void orc_global_batch(int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                      const int32_t *w, int32_t *out, uint32_t *cigar_out, int max_cigar, int64_t *cells, int n_threads)
{
	int8_t mat[25];
	int64_t total = 0;
	int i;
	orc_fill_scmat(1, 4, mat);
	#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64) reduction(+:total)
	for (i = 0; i < n; ++i) {
		int nc = 0;
		int64_t c = 0;
		int sc = orc_ksw_global2((int)(qoff[i + 1] - qoff[i]), q + qoff[i], (int)(toff[i + 1] - toff[i]), t + toff[i], 5, mat,
		                         6, 1, 6, 1, w[i], &nc, cigar_out + (size_t)i * max_cigar, max_cigar, &c);
		out[i * 2] = sc; out[i * 2 + 1] = nc;
		total += c;
	}
	if (cells) *cells = total;
}

void orc_local_batch(int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                     int32_t *out, int64_t *cells, int n_threads)
{
	int8_t mat[25];
	int64_t total = 0;
	int i;
	orc_fill_scmat(1, 4, mat);
	#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64) reduction(+:total)
	for (i = 0; i < n; ++i) {
		int ql = (int)(qoff[i + 1] - qoff[i]);
		int64_t c = 0;
		orc_ksw_align2(ql, q + qoff[i], (int)(toff[i + 1] - toff[i]), t + toff[i], 5, mat, 6, 1, 6, 1, 19, ql * 1 < 250,
		               out + (size_t)i * 7, &c);
		total += c;
	}
	if (cells) *cells = total;
}
