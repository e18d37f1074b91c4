// Warp-cooperative banded Smith-Waterman: one warp per alignment task, rows processed 32 cells at a
// time.  Bit-exact restatements of
//   ksw_extend2 (bwa/ksw.c:416-515)   -> warp_extend
//   ksw_global2 (bwa/ksw.c:540-642)   -> warp_global
//   ksw_align2 / ksw_u8 / ksw_i16 (bwa/ksw.c:122-253,255-370,379-401) -> warp_local
//
// Why a row is data-parallel here: in ksw_extend2 and ksw_global2 gaps open from M (the diagonal
// term), not from H, so within a row E depends only on the previous row and F is a max-plus prefix
// scan of M:  F(j) = max_{k<j} (M(k) - oe_ins - (j-1-k) * e_ins).  The scan runs in 5 shuffle steps
// on (value + k*e_ins).  ksw_align2 opens gaps from H, but F opened from an F-derived H never wins
// (oe > e), so the same scan over H' = max(0, M, E) is exact there too.  The reference's adaptive
// band [beg,end), its tie-breaking (last j / later i win) and its z-drop are row-level scalars
// computed uniformly by the warp, so the cells visited are exactly the reference's cells.
//
// DP rows live in per-warp shared memory (int32 h/e pairs, the eh_t of bwa/ksw.c:412-414); target
// bases are prefetched 32 rows at a time and passed lane-to-lane with shuffles.
#pragma once
#include "common.cuh"

#define FULL_MASK 0xffffffffu
#define KSW_NEG_INF (-0x40000000)  // MINUS_INF, bwa/ksw.c:526
#define KSW_QPAD (EMAB_MAX_READ_LEN + 16)

#define KSW_MAX_TLEN 1024           // local-SW windows are <= 535 + read length (bwa/bwamem_pair.c:156-172)

struct WarpDP {  // per-warp shared-memory working set
	int32_t H[KSW_QPAD + 2];
	int32_t E[KSW_QPAD + 2];
	uint16_t rowmax[KSW_MAX_TLEN];  // per-row maxima of the local pass (second-best bookkeeping)
	uint8_t q[KSW_QPAD + 2];
};


// target fetchers: operator()(i) returns base i of the target in DP order
struct SeqFetch {  // explicit byte string (batch API / tests)
	const uint8_t *p; int step;  // base i = p[i*step]
	__device__ __forceinline__ int operator()(int i) const { return p[(int64_t)i * step]; }
};
struct RefFetch {  // straight from the packed reference: base i = ref[t0 + i*step]
	const DevIndex *ix; int64_t t0; int step;
	__device__ __forceinline__ int operator()(int i) const { return ref_base(*ix, t0 + (int64_t)i * step); }
};

__device__ __forceinline__ int warp_max(int v)
{
	return __reduce_max_sync(FULL_MASK, v);  // one REDUX instead of a 5-step shuffle butterfly
}

// inclusive prefix max across lanes
__device__ __forceinline__ int warp_scan_max(int v, int lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		int n = __shfl_up_sync(FULL_MASK, v, d);
		if (lane >= d) v = max(v, n);
	}
	return v;
}

// per-row substitution bytes of bwa_fill_scmat(1,4) (bwa/bwa.c:136-146) for a byte permute: byte q of the pair
// (lo, hi) is the score of target base t against query code q = A,C,G,T, N (-1) and 5 = the zero-scoring
// padding of the striped query profile (bwa/ksw.c:95-113)
__device__ __forceinline__ uint32_t local_row_scores(int t)
{
	const uint32_t mis = (uint32_t)(uint8_t)(-opt::b) * 0x01010101u;
	const uint32_t flip = (uint32_t)(uint8_t)(-opt::b) ^ (uint32_t)(uint8_t)opt::a;
	return t < 4 ? mis ^ (flip << (t << 3)) : 0xffffffffu;
}
__device__ __forceinline__ int local_score(uint32_t lo, int q)
{
	uint32_t b, r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(b) : "r"(lo), "r"(0x000000ffu), "r"((uint32_t)q));   // byte 4 = -1 (N), byte 5 = 0 (padding)
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(0u), "r"(0x8880u));                 // sign-extend byte 0
	return (int)r;
}

// ---------------------------------------------------------------------------------------------
// ksw_extend2.  sm.q[0..qlen) must hold the query in extension order (and be visible: callers
// __syncwarp() after filling it).  All arguments and the result are warp-uniform.
// ---------------------------------------------------------------------------------------------
template <class TF>
__device__ ExtResult warp_extend(WarpDP &sm, int qlen, const TF &tf, int tlen, int w, int end_bonus, int zdrop, int h0,
                                 unsigned long long *cells)
{
	const int lane = threadIdx.x & 31;
	const int o_del = opt::o_del, e_del = opt::e_del, e_ins = opt::e_ins, oe_del = opt::oe_del, oe_ins = opt::oe_ins;
	// first row (bwa/ksw.c:431-433): H(-1,j) = max(h0 - oe_ins - (j-1)*e_ins, 0), E = 0
	for (int j = lane; j <= qlen; j += 32) {
		int v = j == 0 ? h0 : h0 - oe_ins - (j - 1) * e_ins;
		sm.H[j] = v > 0 ? v : 0;
		sm.E[j] = 0;
	}
	{  // band clamp (bwa/ksw.c:435-443); max(mat) = a
		int max_ins = (int)((double)(qlen * opt::a + end_bonus - opt::o_ins) / e_ins + 1.);
		int max_del = (int)((double)(qlen * opt::a + end_bonus - o_del) / e_del + 1.);
		max_ins = max_ins > 1 ? max_ins : 1;
		max_del = max_del > 1 ? max_del : 1;
		w = w < max_ins ? w : max_ins;
		w = w < max_del ? w : max_del;
	}
	__syncwarp();
	int best = h0, best_i = -1, best_j = -1, g_i = -1, g = -1, max_off = 0;
	int beg = 0, end = qlen;
	unsigned long long visited = 0;
	int tchunk = 0;  // target bases of rows [i & ~31, +32), one per lane
	for (int i = 0; i < tlen; ++i) {
		if ((i & 31) == 0) tchunk = (i + lane < tlen) ? tf(i + lane) : 4;
		const uint32_t rlo = local_row_scores(__shfl_sync(FULL_MASK, tchunk, i & 31));   // byte q = score(target base, q)
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		int carry_h = 0;  // H(i, beg-1)
		if (beg == 0) { carry_h = h0 - (o_del + e_del * (i + 1)); if (carry_h < 0) carry_h = 0; }
		int carry_f = 0;  // F(i, j0)
		int key = -1;                 // lane-local row maximum and its column as (h << 12 | j): a max over keys is "last j wins ties"
		int nz_first = 0x7fffffff, nz_last = -1;  // first / last j in [beg,end] whose new (h,e) is non-zero (warp-uniform)
		const int h_first = carry_h;
		if (end > beg) visited += end - beg;
		for (int j0 = beg; j0 < end; j0 += 32) {
			const int j = j0 + lane;
			const bool act = j < end;
			int diag = 0, e = 0, M = 0;
			if (act) {
				diag = sm.H[j]; e = sm.E[j];
				M = diag ? diag + local_score(rlo, sm.q[j]) : 0;   // bwa/ksw.c:469
			}
			int tI = M - oe_ins; tI = tI > 0 ? tI : 0;       // opens F(i, j+1)
			// F(i,j) = max(carry_f - (j-j0)*e_ins, max_{j0<=k<j} (tI_k - (j-1-k)*e_ins))
			int v = tI + j * e_ins;
			int p = warp_scan_max(v, lane);
			int px = __shfl_up_sync(FULL_MASK, p, 1);
			int f = carry_f - lane * e_ins;
			if (lane) f = max(f, px - (j - 1) * e_ins);
			int h = max(max(M, e), f);
			int hl = __shfl_up_sync(FULL_MASK, h, 1);        // H(i, j-1)
			if (lane == 0) hl = carry_h;
			int nzf = 0;
			if (act) {
				int t = M - oe_del; t = t > 0 ? t : 0;
				int en = max(e - e_del, t);                   // E(i+1, j)
				sm.H[j] = hl;
				sm.E[j] = en;
				key = max(key, (h << 12) | j);
				nzf = hl | en;
			}
			const unsigned nzb = __ballot_sync(FULL_MASK, nzf != 0);
			if (nzb) { nz_first = min(nz_first, j0 + __ffs(nzb) - 1); nz_last = j0 + 31 - __clz(nzb); }
			const int last = min(31, end - 1 - j0);
			carry_h = __shfl_sync(FULL_MASK, h, last);
			carry_f = __shfl_sync(FULL_MASK, max(f - e_ins, tI), 31);
		}
		// eh[end] = {h1, 0}  (bwa/ksw.c:485)
		if (lane == 0) { sm.H[end] = carry_h; sm.E[end] = 0; }
		const int h1 = (end > beg) ? carry_h : h_first;
		if (end <= beg && lane == 0) sm.H[end] = h_first;
		// row max with "last j wins ties"
		const int rkey = warp_max(key);
		const int rm = rkey < 0 ? 0 : rkey >> 12, rj = rkey < 0 ? -1 : (rkey & 4095);
		const int jfin = end > beg ? end : beg;
		if (jfin == qlen) {  // bwa/ksw.c:486-489: later rows win ties
			if (!(g > h1)) g_i = i;
			g = g > h1 ? g : h1;
		}
		if (rm == 0) break;
		if (rm > best) {
			best = rm; best_i = i; best_j = rj;
			int d = rj - i; d = d < 0 ? -d : d;
			max_off = max_off > d ? max_off : d;
		} else if (zdrop > 0) {  // bwa/ksw.c:494-500
			int di = i - best_i, dj = rj - best_j;
			if (di > dj) { if (best - rm - (di - dj) * e_del > zdrop) break; }
			else { if (best - rm - (dj - di) * e_ins > zdrop) break; }
		}
		// next row's [beg,end): first / last non-zero cell of eh[beg..end]  (bwa/ksw.c:502-505)
		int nf = nz_first, nl = nz_last;
		if (h1 != 0) { nl = max(nl, end); nf = min(nf, end); }   // eh[end].h = h1
		int nbeg = nf < end ? nf : end;             // loop stops at j == end
		int jl = nl >= nbeg ? nl : nbeg - 1;        // downward scan stops below beg
		beg = nbeg;
		end = jl + 2 < qlen ? jl + 2 : qlen;
		__syncwarp();
	}
	if (cells && lane == 0) atomicAdd(cells, visited);
	ExtResult r;
	r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = g_i + 1; r.gscore = g; r.max_off = max_off;
	return r;
}

// ---------------------------------------------------------------------------------------------
// ksw_global2.  sm.q[0..qlen) holds the query.  If z != nullptr the direction bytes are written to
// z[i * ncol + (j - beg)] (global scratch) for the backtrack; returns the score.
// ---------------------------------------------------------------------------------------------
template <class TF>
__device__ int warp_global(WarpDP &sm, int qlen, const TF &tf, int tlen, int w, uint8_t *z, unsigned long long *cells)
{
	const int lane = threadIdx.x & 31;
	const int e_del = opt::e_del, e_ins = opt::e_ins, oe_del = opt::oe_del, oe_ins = opt::oe_ins;
	const int ncol = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	for (int j = lane; j <= qlen; j += 32) {  // bwa/ksw.c:558-561
		int hv = KSW_NEG_INF;
		if (j == 0) hv = 0;
		else if (j <= w) hv = -(opt::o_ins + e_ins * j);
		sm.H[j] = hv;
		sm.E[j] = KSW_NEG_INF;
	}
	__syncwarp();
	unsigned long long visited = 0;
	int tchunk = 0;
	for (int i = 0; i < tlen; ++i) {
		if ((i & 31) == 0) tchunk = (i + lane < tlen) ? tf(i + lane) : 4;
		const uint32_t rlo = local_row_scores(__shfl_sync(FULL_MASK, tchunk, i & 31));   // byte q = score(target base, q)
		const int beg = i > w ? i - w : 0;
		const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
		int carry_h = beg == 0 ? -(opt::o_del + e_del * (i + 1)) : KSW_NEG_INF;
		int carry_f = KSW_NEG_INF;
		if (end > beg) visited += end - beg;
		for (int j0 = beg; j0 < end; j0 += 32) {
			const int j = j0 + lane;
			const bool act = j < end;
			int M = KSW_NEG_INF, e = KSW_NEG_INF;
			if (act) { M = sm.H[j] + local_score(rlo, sm.q[j]); e = sm.E[j]; }
			const int tI = M - oe_ins;
			int v = tI + j * e_ins;
			int p = warp_scan_max(v, lane);
			int px = __shfl_up_sync(FULL_MASK, p, 1);
			int f = carry_f - lane * e_ins;
			if (lane) f = max(f, px - (j - 1) * e_ins);
			// bwa/ksw.c:587-600: M preferred over E over F; continuing a gap needs strict >
			int d = M >= e ? 0 : 1;
			int h = M >= e ? M : e;
			d = h >= f ? d : 2;
			h = h >= f ? h : f;
			int hl = __shfl_up_sync(FULL_MASK, h, 1);
			if (lane == 0) hl = carry_h;
			int t = M - oe_del;
			int en = e - e_del;
			d |= en > t ? 1 << 2 : 0;
			en = en > t ? en : t;
			int fn = f - e_ins;
			d |= fn > tI ? 2 << 4 : 0;
			fn = fn > tI ? fn : tI;
			if (act) {
				sm.H[j] = hl;
				sm.E[j] = en;
				if (z) z[(size_t)i * ncol + (j - beg)] = (uint8_t)d;
			}
			const int last = min(31, end - 1 - j0);
			carry_h = __shfl_sync(FULL_MASK, h, last);
			carry_f = __shfl_sync(FULL_MASK, fn, 31);
		}
		if (lane == 0) { sm.H[end] = carry_h; sm.E[end] = KSW_NEG_INF; }
		__syncwarp();
	}
	if (cells && lane == 0) atomicAdd(cells, visited);
	return sm.H[qlen];
}

// Backtrack of ksw_global2 (bwa/ksw.c:624-638), run by every lane redundantly on warp-uniform
// state (z was written by this warp; callers __syncwarp()/fence before).  cigar receives the ops
// in forward order; returns n_cigar (ops beyond max_cigar are counted but not stored).
__device__ inline int global_backtrack(const uint8_t *z, int qlen, int tlen, int w, uint32_t *cigar, int max_cigar, uint32_t *tmp)
{
	const int ncol = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	int n = 0, state = 0;
	int i = tlen - 1;
	int k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
	uint32_t cur = 0;  // op being accumulated (len<<4|op), 0 = none
	auto push = [&](int op, int len) {
		if (cur && (int)(cur & 0xf) == op) cur += (uint32_t)len << 4;
		else {
			if (cur) { if (n < max_cigar) tmp[n] = cur; ++n; }
			cur = (uint32_t)len << 4 | (uint32_t)op;
		}
	};
	while (i >= 0 && k >= 0) {
		int lo = i > w ? i - w : 0;
		state = z[(size_t)i * ncol + (k - lo)] >> (state << 1) & 3;
		if (state == 0) { push(0, 1); --i; --k; }
		else if (state == 1) { push(2, 1); --i; }
		else { push(1, 1); --k; }
	}
	if (i >= 0) push(2, i + 1);
	if (k >= 0) push(1, k + 1);
	if (cur) { if (n < max_cigar) tmp[n] = cur; ++n; }
	int ns = n < max_cigar ? n : max_cigar;
	for (int a = 0; a < ns; ++a) cigar[a] = tmp[ns - 1 - a];
	return n;
}

// ---------------------------------------------------------------------------------------------
// One pass of the striped local SW (ksw_u8 / ksw_i16) as a plain row-parallel DP.  The query is
// padded to a multiple of `lanes` (16 for u8, 8 for i16) with symbols scoring 0 (bwa/ksw.c:95-113):
// sm.q[j] for qlen <= j < qpad must be 5 ("pad").  Row maxima include the padding, as in the
// reference.  Each lane owns the columns j = lane (mod 32), so H/E need no cross-lane ordering.
// ---------------------------------------------------------------------------------------------
struct PassResult { int score, te, qe, score2, te2; };

// One ksw_u8 / ksw_i16 pass (bwa/ksw.c:122-253,255-370) with the query laid out in STRIPS: lane l owns the S
// consecutive columns [l*S, l*S+S) and keeps their H(i-1,.) and E(i,.) in registers.  A row is then
//   1 shuffle   H(i-1, l*S-1) from the lane below,
//   S cells     H' = max(0, diag+s, E) and the gap-open term tI = max(H'-oe, 0)            (registers only)
//   1 scan      exclusive max-plus prefix of the lanes' max(tI_j + j*e): F entering each strip (5 shuffles)
//   S cells     F(j) = max_{k<j}(tI_k + k*e) - (j-1)*e, H = max(H', F), E for the next row   (registers only)
//   1 REDUX     the row maximum,
// i.e. 7 shuffles per row whatever the query length, where 32-column chunks need 8 per chunk and a shared-memory
// round trip per cell.  Same values as the chunked form: F opened from an F-derived H never wins (oe > e), see
// the header.  Columns >= qpad are computed but never stored nor counted.
template <int S, class TF>
__device__ PassResult warp_local_pass(WarpDP &sm, int qlen, int qpad, const TF &tf, int tlen, bool is_u8, int minsc, int endsc,
                                      bool want_sub, unsigned long long *cells)
{
	const int lane = threadIdx.x & 31;
	constexpr int e_del = opt::e_del, e_ins = opt::e_ins, oe_del = opt::oe_del, oe_ins = opt::oe_ins;
	const int shift = opt::b;  // -(min score) = 4 with bwa_fill_scmat(1,4)  (bwa/ksw.c:85-91)
	const int j0 = lane * S;
	int H[S], E[S], Q[S];      // H(i-1, j0+k), E(i, j0+k), query code
#pragma unroll
	for (int k = 0; k < S; ++k) { H[k] = 0; E[k] = 0; Q[k] = j0 + k < qpad ? sm.q[j0 + k] : 5; }
	int gmax = 0, te = -1, qe = 0;
	int tchunk = 0;
	int i;
	for (i = 0; i < tlen; ++i) {
		if ((i & 31) == 0) tchunk = (i + lane < tlen) ? tf(i + lane) : 4;
		const uint32_t rlo = local_row_scores(__shfl_sync(FULL_MASK, tchunk, i & 31));
		int dl = __shfl_up_sync(FULL_MASK, H[S - 1], 1);   // H(i-1, j0-1)
		if (lane == 0) dl = 0;
		int hq[S], tI[S];
		int top = KSW_NEG_INF;
#pragma unroll
		for (int k = 0; k < S; ++k) {
			const int diag = k ? H[k - 1] : dl;
			int v = diag + local_score(rlo, Q[k]);
			v = v > 0 ? v : 0;
			v = v > E[k] ? v : E[k];                       // H' = max(0, diag+s, E)
			hq[k] = v;
			int t = v - oe_ins; t = t > 0 ? t : 0;
			tI[k] = t;
			top = max(top, t + (j0 + k) * e_ins);
		}
		const int incl = warp_scan_max(top, lane);
		int run = __shfl_up_sync(FULL_MASK, incl, 1);      // max_{j < j0} (tI_j + j*e)
		if (lane == 0) run = KSW_NEG_INF;
		int m = 0;
#pragma unroll
		for (int k = 0; k < S; ++k) {
			const int j = j0 + k;
			int f = run - (j - 1) * e_ins; f = f > 0 ? f : 0;
			const int h = hq[k] > f ? hq[k] : f;
			run = max(run, tI[k] + j * e_ins);
			if (j < qpad) {
				int t = h - oe_del; t = t > 0 ? t : 0;
				int en = E[k] - e_del; en = en > t ? en : t;   // gaps open from H here (bwa/ksw.c:189-196)
				E[k] = en;
				H[k] = h;
				m = max(m, h);
			}
		}
		const int imax = warp_max(m);
		if (want_sub && lane == 0 && i < KSW_MAX_TLEN) sm.rowmax[i] = (uint16_t)imax;
		if (imax > gmax) {  // bwa/ksw.c:224-229
			gmax = imax; te = i;
			int c = 0x7fffffff;   // qe = smallest query index holding the row maximum (bwa/ksw.c:235-239)
#pragma unroll
			for (int k = S - 1; k >= 0; --k) if (j0 + k < qpad && H[k] == imax) c = j0 + k;
			qe = -warp_max(-c);
			if ((is_u8 && gmax + shift >= 255) || gmax >= endsc) { ++i; break; }
		}
	}
	if (cells && lane == 0) atomicAdd(cells, (unsigned long long)qlen * (unsigned long long)i);
	PassResult r;
	const bool sat = is_u8 && gmax + shift >= 255;
	r.score = sat ? 255 : gmax;
	r.te = te; r.qe = sat ? -1 : qe; r.score2 = -1; r.te2 = -1;
	if (want_sub && !sat) {
		// bwa/ksw.c:215-223: consecutive rows with imax >= minsc form a run represented by its first
		// maximum; :241-249: best run (first on ties) whose row lies outside te +- ceil(score/a).
		__syncwarp();
		const int k = (r.score + opt::a - 1) / opt::a, lo = te - k, hi = te + k;
		const int nrow = i < KSW_MAX_TLEN ? i : KSW_MAX_TLEN;
		// NB "consecutive" is judged against the row stored in the last entry, which is the row of that
		// entry's maximum, not the previous row (b[n_b-1] + 1 != i, bwa/ksw.c:216).
		int run_row = -2, run_val = -1;
		for (int a = 0; a < nrow; ++a) {
			const int v = sm.rowmax[a];
			if (v < minsc) continue;
			if (run_val < 0 || run_row + 1 != a) {
				if (run_val >= 0 && (run_row < lo || run_row > hi) && run_val > r.score2) { r.score2 = run_val; r.te2 = run_row; }
				run_val = v; run_row = a;
			} else if (run_val < v) { run_val = v; run_row = a; }
		}
		if (run_val >= 0 && (run_row < lo || run_row > hi) && run_val > r.score2) { r.score2 = run_val; r.te2 = run_row; }
	}
	return r;
}

// strip width by query length: the smallest instantiated S with 32*S >= qpad (qpad <= EMAB_MAX_READ_LEN + 16)
template <class TF>
__device__ __forceinline__ PassResult warp_local_pass(WarpDP &sm, int qlen, int qpad, const TF &tf, int tlen, bool is_u8, int minsc, int endsc,
                                                      bool want_sub, unsigned long long *cells)
{
	static_assert(EMAB_MAX_READ_LEN + 16 <= 32 * 9, "widest strip");
	if (qpad <= 32 * 4) return warp_local_pass<4>(sm, qlen, qpad, tf, tlen, is_u8, minsc, endsc, want_sub, cells);
	if (qpad <= 32 * 5) return warp_local_pass<5>(sm, qlen, qpad, tf, tlen, is_u8, minsc, endsc, want_sub, cells);
	if (qpad <= 32 * 7) return warp_local_pass<7>(sm, qlen, qpad, tf, tlen, is_u8, minsc, endsc, want_sub, cells);
	return warp_local_pass<9>(sm, qlen, qpad, tf, tlen, is_u8, minsc, endsc, want_sub, cells);
}

template <class TF>
struct RevPrefixFetch {  // target of the second ksw_align2 pass: first te+1 bases reversed, rest untouched (bwa/ksw.c:393-395)
	TF tf; int te;
	__device__ __forceinline__ int operator()(int i) const { return tf(i <= te ? te - i : i); }
};

// ksw_align2 with KSW_XSUBO|KSW_XSTART (| KSW_XBYTE when is_u8), as mem_matesw calls it
// (bwa/bwamem_pair.c:176-177).  sm.q[0..qlen) holds the query; it is clobbered.
template <class TF>
__device__ LocResult warp_local(WarpDP &sm, int qlen, const TF &tf, int tlen, int minsc, bool is_u8, unsigned long long *cells)
{
	const int lane = threadIdx.x & 31;
	const int lanes = is_u8 ? 16 : 8;
	int qpad = (qlen + lanes - 1) / lanes * lanes;
	for (int j = qlen + lane; j < qpad; j += 32) sm.q[j] = 5;
	__syncwarp();
	PassResult r = warp_local_pass(sm, qlen, qpad, tf, tlen, is_u8, minsc, 0x10000, true, cells);
	LocResult o;
	o.score = r.score; o.te = r.te; o.qe = r.qe; o.score2 = r.score2; o.te2 = r.te2; o.tb = -1; o.qb = -1;
	if (r.score < minsc || r.qe < 0) return o;  // bwa/ksw.c:392
	// second pass over the reversed prefixes, stopping once the score is reached (bwa/ksw.c:393-399)
	const int ql2 = r.qe + 1;
	__syncwarp();
	for (int j = lane; j < ql2 / 2; j += 32) { uint8_t t = sm.q[j]; sm.q[j] = sm.q[ql2 - 1 - j]; sm.q[ql2 - 1 - j] = t; }
	qpad = (ql2 + lanes - 1) / lanes * lanes;
	__syncwarp();
	for (int j = ql2 + lane; j < qpad; j += 32) sm.q[j] = 5;
	__syncwarp();
	RevPrefixFetch<TF> rf{tf, r.te};
	PassResult rr = warp_local_pass(sm, ql2, qpad, rf, tlen, is_u8, 0x10000, r.score, false, cells);
	if (r.score == rr.score) { o.tb = r.te - rr.te; o.qb = r.qe - rr.qe; }
	return o;
}
