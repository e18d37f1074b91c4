// Inter-read wave scheduling of mem_reg2aln's ksw_global2 calls (bwa/bwamem.c:1119-1150, bwa/bwa.c:148-234,
// bwa/ksw.c:540-642): plan, batch, replay — the same shape as ext_wave.cuh and the mate-rescue plan.
//
// The first bwa_gen_cigar2 call of a region is a pure function of (read, region): which bases are aligned, in which
// orientation, with which band.  So before the per-pair kernel walks append_alignments:
//
//   k_glob_plan  thread / read   for every region of the read: the first call's arguments -> one GlobTask, or none when
//                                the call takes the gap-free path (no DP)
//   (sort, scan)                 tasks ordered by (band width, rows); backtrack-matrix offsets per warp of 32 tasks
//   k_glob_wave  thread / task   ksw_global2 + backtrack, 32 similar-sized tasks per warp, DP row {H, E} as two int16 in
//                                one shared-memory word per column, one direction byte per cell (bwa/ksw.c:587-600)
//   k_finalize   warp / pair     as before; a ksw_global2 call whose arguments match a task copies its score and CIGAR,
//                                anything else (a band retry, bwa/bwamem.c:1136-1142) is computed inline
//
// int16 is exact here: a DP value is either reachable — then it lies in [-(o + e*rows) - b*cols, a*cols], a few thousand
// at most — or it is MINUS_INF plus the same small offsets the reference accumulates on top of its -2^30; with
// MINUS_INF16 = -16384 and offsets below 8 * (rows + cols) < 14000 the two ranges never meet and never wrap, so every
// comparison, hence every direction byte and the score, is the reference's.
#pragma once
#include "align_lanes.cuh"

struct GlobTask {
	const uint8_t *query;          // the read (null: this region needs no DP or is not planned)
	int64_t t0;
	int32_t q0, qlen, tlen, w;
	int8_t qstep, tstep;
	int16_t n_cigar;               // result: CIGAR operations (may exceed EMAB_MAX_CIGAR; only the last EMAB_MAX_CIGAR are kept)
	int32_t score;                 // result
	uint32_t cells;
	uint32_t pad;
};
#define GLOB_NEG16 (-16384)
#define GLOB_RING_COLS 32         // bands of up to 31 columns are planned; the DP row of a lane is a ring of this many columns
#define GLOB_WIDE_COLS 28         // bands of this many columns or more: one warp per task (k_glob_wide)
#define GLOB_MAX_DIM 1024          // rows + columns the int16 argument above covers; larger calls stay inline

#ifdef __CUDACC__

// the arguments of reg2aln's first gen_cigar call (align.cuh: reg2aln, gen_cigar); false when it does not reach the DP
EMAB_HD bool glob_first_call(const DevIndex &ix, int l_query_read, const Reg &ar, int *q0, int *qstep, int *qlen, int64_t *t0, int *tstep, int *tlen, int *w_out)
{
	(void)l_query_read;
	const int qb = ar.qb, qe = ar.qe;
	const int64_t rb = ar.rb, re = ar.re;
	int tmp = infer_bw(qe - qb, (int)(re - rb), ar.truesc, opt::a, opt::o_del, opt::e_del);
	int w2 = infer_bw(qe - qb, (int)(re - rb), ar.truesc, opt::a, opt::o_ins, opt::e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt::w) w2 = w2 < ar.w ? w2 : ar.w;
	w2 = w2 < opt::w << 2 ? w2 : opt::w << 2;
	const int l_query = qe - qb;
	if (l_query <= 0 || rb >= re || (rb < ix.l_pac && re > ix.l_pac)) return false;
	const int rlen = (int)(re - rb);
	if (l_query == rlen && w2 == 0) return false;   // gap-free: no DP
	const bool rev = rb >= ix.l_pac;
	*q0 = rev ? qb + l_query - 1 : qb; *qstep = rev ? -1 : 1; *qlen = l_query;
	*t0 = rev ? re - 1 : rb; *tstep = rev ? -1 : 1; *tlen = rlen;
	int max_ins = (int)((double)(((l_query + 1) >> 1) * opt::a - opt::o_ins) / opt::e_ins + 1.);
	int max_del = (int)((double)(((l_query + 1) >> 1) * opt::a - opt::o_del) / opt::e_del + 1.);
	int max_gap = max_ins > max_del ? max_ins : max_del;
	max_gap = max_gap > 1 ? max_gap : 1;
	int dl = rlen - l_query; dl = dl < 0 ? -dl : dl;
	int w = (max_gap + dl + 1) >> 1;
	w = w < w2 ? w : w2;
	const int min_w = dl + 3;
	w = w > min_w ? w : min_w;
	*w_out = w;
	return true;
}

template <class Pools_>
__global__ void __launch_bounds__(128)
k_glob_plan(DevIndex ix, int n_reads, const uint8_t *seq, const int64_t *off, const int32_t *occ_off, Pools_ p, int rescue_room, const int32_t *aln_off,
            GlobTask *tasks, uint16_t *keys)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const Reg *regs = p.regs + (occ_off[r] + (size_t)rescue_room * r);
	const int n = p.n_regs[r], l_query = (int)(off[r + 1] - off[r]);
	for (int i = 0; i < n; ++i) {
		const int slot = aln_off[r] + i;
		GlobTask &t = tasks[slot];
		int q0, qstep, qlen, tstep, tlen, w;
		int64_t t0;
		t.query = nullptr; t.n_cigar = 0; t.score = 0; t.cells = 0; t.pad = 0;
		keys[slot] = 0;
		if (!glob_first_call(ix, l_query, regs[i], &q0, &qstep, &qlen, &t0, &tstep, &tlen, &w)) continue;
		if (qlen + tlen > GLOB_MAX_DIM || qlen > EMAB_MAX_READ_LEN) continue;
		t.query = seq + off[r]; t.t0 = t0; t.q0 = q0; t.qlen = qlen; t.tlen = tlen; t.w = w; t.qstep = (int8_t)qstep; t.tstep = (int8_t)tstep;
		// sort key: band width first, rows second — a warp's row costs its widest band, its length its longest task
		const int ncol = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
		keys[slot] = (uint16_t)((ncol > 255 ? 255 : ncol) << 8 | (((tlen + 3) >> 2) > 255 ? 255 : ((tlen + 3) >> 2)));
	}
}

// Upper bound of a warp's backtrack matrix: its 32 tasks (in sorted order) advance row by row together and a row is
// as wide as its widest lane, so 32 * max(ncol) * max(tlen) bytes always suffice.
__global__ void k_glob_zsize(const GlobTask *tasks, const int32_t *order, const uint16_t *keys_sorted, int n_tasks, unsigned long long *zsize_warp, int n_warps, int wide_cols)
{
	const int wi = blockIdx.x * blockDim.x + threadIdx.x;
	if (wi >= n_warps) return;
	int ncol = 0, tlen = 0;
	for (int l = 0; l < 32; ++l) {
		const int ti = wi * 32 + l;
		if (ti >= n_tasks || keys_sorted[ti] == 0) break;   // sorted descending: nothing valid follows
		if ((keys_sorted[ti] >> 8) >= wide_cols) continue;   // wide bands: k_glob_wide, with its own scratch
		const GlobTask &t = tasks[order[ti]];
		const int nc = t.qlen < 2 * t.w + 1 ? t.qlen : 2 * t.w + 1;
		ncol = nc > ncol ? nc : ncol;
		tlen = t.tlen > tlen ? t.tlen : tlen;
	}
	zsize_warp[wi] = 32ull * ncol * tlen;
}

// ksw_global2 (bwa/ksw.c:540-622) + backtrack (:624-638), one task per thread, the 32 tasks of a warp row by row
// together.  Everything a cell touches is laid out so that a warp instruction is ONE memory transaction:
//   * DP row and query in shared memory, lane-interleaved (word j of lane l at [j * 32 + l]): conflict-free;
//   * direction bytes at zw[(row_base + jj) * 32 + lane], jj = column offset inside the row's band and row_base the
//     running sum of the rows' widest bands: the 32 lanes of a store instruction write 32 consecutive bytes (one sector).
//     With one private matrix per lane every store would be a transaction of its own — 32 L1 wavefronts and 32
//     partial-sector writes in L2 per instruction, which is what bounded the first version of this kernel.
__global__ void __launch_bounds__(32)
k_glob_wave(DevIndex ix, GlobTask *tasks, const int32_t *order, const uint16_t *keys_sorted, int n_tasks, const unsigned long long *zoff_warp, uint8_t *zpool,
            uint32_t *cigars, unsigned long long *planned_cells, int qcap, int ring_cols, int ncol_lo, int ncol_hi)
{
	// One launch serves the tasks whose band (columns per row, the high byte of the sort key) lies in (ncol_lo, ncol_hi].
	// ring_cols != 0: the DP row is kept as a RING of that many columns (a power of two > the band): a row only ever
	// touches columns [beg, end], at most band + 1 of them, so a narrow band — most tasks — needs 4 KB of shared memory
	// per warp instead of 19 KB and twice as many warps fit an SM.  ring_cols == 0: one word per query column.
	extern __shared__ uint32_t glob_smem[];
	const int lane = threadIdx.x, ti = blockIdx.x * 32 + lane;
	const int kcol = ti < n_tasks ? keys_sorted[ti] >> 8 : 0;
	const bool valid = ti < n_tasks && keys_sorted[ti] != 0 && kcol > ncol_lo && kcol <= ncol_hi;
	if (!__any_sync(FULL_MASK, valid)) return;
	const int he_cols = ring_cols ? ring_cols : qcap + 1;
	const unsigned rmask = ring_cols ? (unsigned)ring_cols - 1 : ~0u;
	uint32_t *he = glob_smem + lane;                          // column j at he[(j & rmask) * 32]: {H(i-1, j-1) : lo16, E(i, j) : hi16}
	uint32_t *qw = glob_smem + he_cols * 32 + lane;           // query bases j..j+3 in word (j >> 2)
	uint32_t *rowbase = glob_smem + he_cols * 32 + ((qcap + 4) >> 2) * 32;   // [GLOB_MAX_DIM] warp-uniform
	uint8_t *zw = zpool + zoff_warp[blockIdx.x] + lane;
	unsigned long long cells = 0;
	int slot = 0, qlen = 0, tlen = 0, w = 0, tstep = 1;
	int64_t t0 = 0;
	constexpr int e_del = opt::e_del, e_ins = opt::e_ins, oe_del = opt::oe_del, oe_ins = opt::oe_ins;
	if (valid) {
		slot = order[ti];
		const GlobTask &t = tasks[slot];
		qlen = t.qlen; tlen = t.tlen; w = t.w; tstep = t.tstep; t0 = t.t0;
		const uint8_t *query = t.query;
		const int q0 = t.q0, qstep = t.qstep;
		// bwa/ksw.c:558-561.  Column j > w + 1 is written (end of row j - w - 1) before row j - w first reads it
		const int jinit = qlen < w + 1 ? qlen : w + 1;
		for (int j = 0; j <= jinit; ++j) {
			const int h = j == 0 ? 0 : (j <= w ? -(opt::o_ins + e_ins * j) : GLOB_NEG16);
			he[(j & rmask) * 32] = ((uint32_t)(uint16_t)GLOB_NEG16 << 16) | ((uint32_t)h & 0xffffu);
		}
		for (int j = 0; j < qlen; j += 4) {
			uint32_t v = 0;
			for (int k = 0; k < 4 && j + k < qlen; ++k) v |= (uint32_t)query[q0 + (j + k) * qstep] << (8 * k);
			qw[(j >> 2) * 32] = v;
		}
	}
	const int tl_max = (int)__reduce_max_sync(FULL_MASK, (unsigned)tlen);
	unsigned row_base = 0;
	for (int i = 0; i < tl_max; ++i) {
		const bool act = i < tlen;
		int beg = 0, end = 0, tb = 4;
		if (act) {
			beg = i > w ? i - w : 0;
			end = i + w + 1 < qlen ? i + w + 1 : qlen;
			tb = ref_base(ix, t0 + (int64_t)i * tstep);
		}
		const int width = end > beg ? end - beg : 0;
		const int mw = (int)__reduce_max_sync(FULL_MASK, (unsigned)width);
		if (lane == 0) rowbase[i] = row_base;
		int h1 = beg == 0 ? -(opt::o_del + e_del * (i + 1)) : GLOB_NEG16, f = GLOB_NEG16;
		cells += width;
		uint8_t *zr = zw + (size_t)row_base * 32;
		for (int jj = 0; jj < mw; ++jj) {
			if (jj < width) {
				const int j = beg + jj;
				uint32_t *hp = he + ((unsigned)j & rmask) * 32;
				const uint32_t wv = *hp;
				const int qb = (int)(qw[(j >> 2) * 32] >> ((j & 3) << 3)) & 0xff;
				const int M = (int)(short)(wv & 0xffffu) + sc_mat(tb, qb);
				int e = (int)wv >> 16;
				int d = M >= e ? 0 : 1;
				int h = M >= e ? M : e;
				d = h >= f ? d : 2;
				h = h >= f ? h : f;
				int tt = M - oe_del;
				e -= e_del;
				d |= e > tt ? 1 << 2 : 0;
				e = e > tt ? e : tt;
				*hp = ((uint32_t)e << 16) | ((uint32_t)h1 & 0xffffu);
				h1 = h;
				tt = M - oe_ins;
				f -= e_ins;
				d |= f > tt ? 2 << 4 : 0;
				f = f > tt ? f : tt;
				zr[(size_t)jj * 32] = (uint8_t)d;
			}
		}
		if (act) he[((unsigned)end & rmask) * 32] = ((uint32_t)(uint16_t)GLOB_NEG16 << 16) | ((uint32_t)h1 & 0xffffu);
		row_base += (unsigned)mw;
	}
	__syncwarp();
	if (valid) {
		GlobTask &t = tasks[slot];
		t.score = (int)(short)(he[((unsigned)qlen & rmask) * 32] & 0xffffu);
		t.cells = (uint32_t)cells;
		// backtrack: operations are met last to first and written back to front, so the kept ones are in forward order at
		// the end of the task's EMAB_MAX_CIGAR slots
		uint32_t *cg = cigars + (size_t)slot * EMAB_MAX_CIGAR;
		int n = 0, state = 0, i = tlen - 1;
		int k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		uint32_t cur = 0;
		auto push = [&](int op, int len) {
			if (cur && (int)(cur & 0xf) == op) cur += (uint32_t)len << 4;
			else {
				if (cur) { if (n < EMAB_MAX_CIGAR) cg[EMAB_MAX_CIGAR - 1 - n] = cur; ++n; }
				cur = (uint32_t)len << 4 | (uint32_t)op;
			}
		};
		while (i >= 0 && k >= 0) {
			const int lo = i > w ? i - w : 0;
			state = zw[((size_t)rowbase[i] + (k - lo)) * 32] >> (state << 1) & 3;
			if (state == 0) { push(0, 1); --i; --k; }
			else if (state == 1) { push(2, 1); --i; }
			else { push(1, 1); --k; }
		}
		if (i >= 0) push(2, i + 1);
		if (k >= 0) push(1, k + 1);
		if (cur) { if (n < EMAB_MAX_CIGAR) cg[EMAB_MAX_CIGAR - 1 - n] = cur; ++n; }
		t.n_cigar = (int16_t)(n > 32767 ? 32767 : n);
	}
	for (int d = 16; d; d >>= 1) cells += __shfl_xor_sync(FULL_MASK, cells, d);
	if (lane == 0 && cells) atomicAdd(planned_cells, cells);
}

// The wide bands (GLOB_RING_COLS columns or more: a tenth of the tasks, half of the cells), ONE WARP PER TASK: a wave lasts as
// long as its longest lane, and one lane walking a 150 x 250 matrix alone is ~1 ms — what the bucket used to wait for.
// The 32 lanes take a row together (warp_global, ksw_warp.cuh), tens of microseconds per task; tasks are the head of the
// sorted order and are handed out through a counter.  Runs beside k_glob_wave on the ctx's second stream.  Results in
// the same places and format as k_glob_wave's.
__global__ void __launch_bounds__(256)
k_glob_wide(DevIndex ix, GlobTask *tasks, const int32_t *order, const uint16_t *keys_sorted, int n_tasks, uint32_t *cigars, uint8_t *zbuf, size_t z_cap,
            uint32_t *tmpbuf, unsigned long long *planned_cells, unsigned long long *queue, int wide_cols)
{
	__shared__ WarpDP sm_all[8];
	const int lane = threadIdx.x & 31, gw = blockIdx.x * 8 + (threadIdx.x >> 5);
	WarpDP &sm = sm_all[threadIdx.x >> 5];
	uint8_t *z = zbuf + (size_t)gw * z_cap;
	uint32_t *tmp = tmpbuf + (size_t)gw * EMAB_MAX_CIGAR;
	unsigned long long total = 0;
	for (;;) {
		unsigned long long pos = 0;
		if (lane == 0) pos = atomicAdd(queue, 1ull);
		pos = __shfl_sync(FULL_MASK, pos, 0);
		if (pos >= (unsigned long long)n_tasks || (keys_sorted[pos] >> 8) < wide_cols) break;   // sorted descending
		const int slot = order[pos];
		GlobTask &t = tasks[slot];
		const int qlen = t.qlen, tlen = t.tlen, w = t.w;
		const int ncol = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
		if ((size_t)ncol * tlen > z_cap) { __syncwarp(); if (lane == 0) t.query = nullptr; continue; }   // not planned after all: k_finalize's business
		const uint8_t *query = t.query;
		const int q0 = t.q0, qstep = t.qstep;
		__syncwarp();
		for (int j = lane; j < qlen; j += 32) sm.q[j] = query[q0 + j * qstep];
		__syncwarp();
		RefFetch tf{&ix, t.t0, t.tstep};
		const int score = warp_global(sm, qlen, tf, tlen, w, z, nullptr);
		__syncwarp();
		uint32_t fwd[EMAB_MAX_CIGAR];
		const int n = global_backtrack(z, qlen, tlen, w, fwd, EMAB_MAX_CIGAR, tmp);
		const int ns = n < EMAB_MAX_CIGAR ? n : EMAB_MAX_CIGAR;
		uint32_t *cg = cigars + (size_t)slot * EMAB_MAX_CIGAR + (EMAB_MAX_CIGAR - ns);
		for (int a = lane; a < ns; a += 32) cg[a] = fwd[a];
		unsigned cells = 0;   // the cells ksw_global2 visits: the band of every row (bwa/ksw.c:571-573)
		for (int i = lane; i < tlen; i += 32) {
			const int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
			cells += end > beg ? (unsigned)(end - beg) : 0u;
		}
		cells = __reduce_add_sync(FULL_MASK, cells);
		if (lane == 0) { t.score = score; t.n_cigar = (int16_t)(n > 32767 ? 32767 : n); t.cells = cells; }
		total += cells;
		__syncwarp();
	}
	if (lane == 0 && total) atomicAdd(planned_cells, total);
}

// shared memory of one k_glob_wave warp for queries up to qcap bases
__host__ __device__ inline size_t glob_smem_bytes(int qcap, int ring_cols = 0)
{
	return ((size_t)(ring_cols ? ring_cols : qcap + 1) * 32 + (size_t)((qcap + 4) >> 2) * 32 + GLOB_MAX_DIM) * 4;
}

// a ksw_global2 call of the replay against the read's tasks; on a hit copies the CIGAR (if wanted) and returns true
__device__ __forceinline__ bool glob_plan_lookup(const GlobTask *tasks, const uint32_t *cigars, int n_tasks, const uint8_t *query, int q0, int qstep, int qlen,
                                                 int64_t t0, int tlen, int w, uint32_t *cigar, int *n_cigar, int *score, uint32_t *cells)
{
	for (int k = 0; k < n_tasks; ++k) {
		const GlobTask &t = tasks[k];
		if (t.query != query || t.q0 != q0 || t.qstep != qstep || t.qlen != qlen || t.t0 != t0 || t.tlen != tlen || t.w != w) continue;
		*score = t.score; *cells = t.cells;
		if (cigar) {
			const int n = t.n_cigar, ns = n < EMAB_MAX_CIGAR ? n : EMAB_MAX_CIGAR;
			const uint32_t *cg = cigars + (size_t)k * EMAB_MAX_CIGAR + (EMAB_MAX_CIGAR - ns);
			for (int a = threadIdx.x & 31; a < ns; a += 32) cigar[a] = cg[a];
			__syncwarp();
			*n_cigar = n;
		}
		return true;
	}
	return false;
}

#endif
