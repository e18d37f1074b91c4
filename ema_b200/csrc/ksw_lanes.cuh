// Inter-task banded Smith-Waterman: ONE THREAD PER ALIGNMENT TASK, 32 tasks per warp advancing row by
// row together.  Bit-exact restatement of ksw_extend2 (bwa/ksw.c:416-515) — each lane literally runs
// the reference's scalar loop over its own [beg,end) — laid out for the B200 integer pipes.
//
// Measured on B200 (bench_sw.py / emab_int_peak): LOP3, SHF, PRMT, VIMNMX, VIMNMX3 run on the ALU pipe at
// 16 lanes/clk/SMSP, IMAD (and IADD as IMAD.IADD) on the FMA pipe at 16 lanes/clk/SMSP, the DPX
// add-then-min/max (VIADDMNMX) occupies both; only add-class work reaches 32 lanes/clk/SMSP.  A cell is
// max-dominated, so the kernel is bound by the ALU pipe and the cell is written to minimise ALU-pipe
// instructions:
//
//   * the per-column DP state eh_t{h,e} (bwa/ksw.c:412-414) AND the query base of that column are ONE
//     32-bit shared-memory word per column, stored lane-interleaved: word (j, lane) at [j*32 + lane] —
//     lane L only ever touches bank L whatever column each lane is at: conflict-free without any
//     alignment of bands, and a cell is one LDS + one STS.
//   * all DP values are kept multiplied by 16 in registers.  Word layout: bits 0-3 query base, bits 4-15
//     h*16 (so `word & 0xfff0` IS h*16), bits 16-31 e*16 (so `word >> 16` IS e*16, done by IMAD.HI on
//     the FMA pipe).  Scores are < 4096 (checked by the caller).
//   * the substitution score is a byte permute: the low nibble of the word is the selector of a PRMT over
//     a per-row register holding 16*score(t, q) for q = A,C,G,T,N; a second PRMT sign-extends.
//   * the dead-cell rule `M = M ? M + s : 0` (bwa/ksw.c:469) is min(M + s, M * 1024) (a non-positive M
//     behaves as 0 everywhere downstream); E and F share the gap-open term (o_del+e_del == o_ins+e_ins in
//     BWA-MEM's defaults) and use the three-input RELU forms max(x - e, M - oe, 0);
//     the row maximum with "last j wins ties" (bwa/ksw.c:473) is one max over keys h*2^16 + j.
//   * rows are synchronous across the warp: cells below the narrowest live band of the 32 lanes run
//     unpredicated, 8 per iteration with the loads of the next group in flight (lanes whose task is over
//     run along as zombies on their own dead columns); the few cells between the narrowest and the widest
//     band run four at a time with their effects predicated per lane; the per-row scalar work (band shrink, z-drop, end-of-query score) is
//     the reference's code per lane.
//
// Used by emab_extend_batch (config 5 of BASELINE.json) and by the pipeline's extension stage.
#pragma once
#include "common.cuh"

#ifndef FULL_MASK
#define FULL_MASK 0xffffffffu
#endif

namespace lanes {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}

constexpr int SCALE = 16;        // DP values live in registers multiplied by 16 (see the word layout above)
static_assert(opt::b * SCALE <= 127, "scaled scores must fit the int8 lanes of the per-row register");

// x * 1024 on the FMA pipe (opaque to the compiler, which would otherwise re-derive it from the packed word
// with one more ALU-pipe instruction)
__device__ __forceinline__ int mul1024(int x)
{
	int r;
	asm("mul.lo.s32 %0, %1, 1024;" : "=r"(r) : "r"(x));
	return r;
}

// The key of the row maximum, "last j wins ties" (bwa/ksw.c:473): (h*16) * KEY_MUL + j as ONE IMAD on the FMA
// pipe.  KEY_MUL is deliberately not a power of two: ptxas turns a shift-and-add into ALU-pipe instructions.
constexpr int KEY_MUL = 4097;    // > any column index; 4000*16*4097 < 2^31
__device__ __forceinline__ int row_key(int h, int j)
{
	int r;
	asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(h), "n"(KEY_MUL), "r"(j));
	return r;
}

// per-row substitution register: byte q (0..3) = 16*score(t, q) as int8; a query N selects byte 4 of the
// pair (-16); a target N scores -1 against everything (bwa/bwa.c:136-146, bwa_fill_scmat(1,4))
__device__ __forceinline__ uint32_t row_scores(int tb)
{
	const uint32_t mis = (uint32_t)(uint8_t)(-opt::b * SCALE) * 0x01010101u;
	const uint32_t flip = (uint32_t)(uint8_t)(-opt::b * SCALE) ^ (uint32_t)(uint8_t)(opt::a * SCALE);
	return tb < 4 ? mis ^ (flip << (tb << 3)) : (uint32_t)(uint8_t)(-SCALE) * 0x01010101u;
}

// bytes of shared memory one warp needs for queries up to qcap bases
__host__ __device__ inline size_t smem_per_warp(int qcap) { return (size_t)(qcap + 1) * 128; }
constexpr int MAX_SCORE = 4000;  // h0 + qlen*a must stay below this (12-bit h field)

struct State {  // per-lane scalars of one ksw_extend2 call
	int best, best_i, best_j, g, g_i, max_off, beg, end;
};

// One ksw_extend2 per lane.  eh already points at this lane's column 0 (stride 32 words).
// qf(j) / tf(i) return base j of the query / base i of the target in extension order.  `valid` lanes
// run a task; the others only take part in the warp-wide votes.  All lanes of the warp must call.
template <class QF, class TF>
__device__ __forceinline__ ExtResult extend(uint32_t *eh, bool valid, int qlen, int tlen, int h0, int w, int end_bonus,
                                            int zdrop, const QF &qf, const TF &tf, unsigned long long &visited)
{
	constexpr int o_del = opt::o_del, e_del = opt::e_del, e_ins = opt::e_ins, oe_del = opt::oe_del, oe_ins = opt::oe_ins;
	static_assert(oe_del == oe_ins && e_del == e_ins, "the cell below shares the gap-open term between E and F (BWA-MEM defaults)");
	State s;
	s.best = h0; s.best_i = -1; s.best_j = -1; s.g = -1; s.g_i = -1; s.max_off = 0; s.beg = 0; s.end = qlen;
	if (valid) {
		// first row (bwa/ksw.c:431-433); e = 0 everywhere; bits 0-3 carry the query base of the column
		int v = h0;
		for (int j = 0; j <= qlen; ++j) {
			const uint32_t qb = j < qlen ? (uint32_t)qf(j) & 7u : 0u;
			eh[j * 32] = (uint32_t)(v * SCALE) | qb;
			v = j == 0 ? (h0 > oe_ins ? h0 - oe_ins : 0) : (v > e_ins ? v - e_ins : 0);
		}
		// band clamp (bwa/ksw.c:435-443); max(mat) = a
		int max_ins = (int)((double)(qlen * opt::a + end_bonus - opt::o_ins) / e_ins + 1.);
		int max_del = (int)((double)(qlen * opt::a + end_bonus - o_del) / e_del + 1.);
		max_ins = max_ins > 1 ? max_ins : 1;
		max_del = max_del > 1 ? max_del : 1;
		w = w < max_ins ? w : max_ins;
		w = w < max_del ? w : max_del;
	}
	bool alive = valid;
	unsigned long long cells = 0;
	// 1, but not to the compiler: `x * one - c` stays an IMAD (FMA pipe).  A *_sync vote over the full mask (all lanes
	// must call anyway) — __activemask() is NOT the same thing: after the divergent set-up above the warp need not
	// have reconverged, and a partial mask here silently corrupts every gap-open term.
	const int one = __popc(__ballot_sync(FULL_MASK, true)) - 31;
	int tb_next = valid && tlen > 0 ? tf(0) : 0;   // the target base is fetched one row ahead of its use
	for (int i = 0;; ++i) {
		alive = alive && i < tlen;
		if (!__any_sync(FULL_MASK, alive)) break;
		int width = 0, h1 = 0, f = 0, mkey = -1, j = 0;   // h1, f are SCALEd
		uint32_t rlo = 0;
		uint32_t *p = eh;
		const int tb = tb_next;
		if (alive && i + 1 < tlen) tb_next = tf(i + 1);
		if (alive) {
			if (s.beg < i - w) s.beg = i - w;
			if (s.end > i + w + 1) s.end = i + w + 1;
			if (s.end > qlen) s.end = qlen;
			if (s.beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); h1 = h1 > 0 ? h1 * SCALE : 0; }
			width = s.end > s.beg ? s.end - s.beg : 0;
			rlo = row_scores(tb);
			p = eh + s.beg * 32;
			j = s.beg;
			cells += width;
		}
		// the first group's words are requested before the warp-wide reductions below (nfast >= 4 implies width >= 4
		// on every live lane; the dead lanes compute on zeros)
		uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
		if (width >= 4) { a0 = p[0]; a1 = p[32]; a2 = p[64]; a3 = p[96]; }
		// Lanes whose task is over (or that never had one) run along as zombies on their own, now dead,
		// columns, so that the cells up to the narrowest LIVE band need no predicate at all.
		const int maxw = (int)__reduce_max_sync(FULL_MASK, (unsigned)width);
		const int minw = (int)__reduce_min_sync(FULL_MASK, alive ? (unsigned)width : 0x7fffffffu);
		const int nfast = minw < maxw ? minw : maxw;
		const uint32_t rhi = (uint32_t)(uint8_t)(-SCALE);
#define EMAB_LANES_CELL(WV, JJ, KN)                                                                                \
		{                                                                                                        \
			const uint32_t wv = (WV);                                                                            \
			const int diag = (int)(wv & 0xfff0u), e = (int)__umulhi(wv, 65536u);       /* h*16, e*16 */          \
			const uint32_t sb = prmt(rlo, rhi, wv);          /* byte 0 = 16*score(t, q); the rest is garbage */  \
			const int sc = (int)prmt(sb, 0u, 0x8880u);       /* sign-extend byte 0 */                            \
			const int M = __viaddmin_s32(diag, sc, mul1024(diag));  /* diag ? diag + sc : <= 0  (ksw.c:469) */   \
			const int h = __vimax3_s32(M, e, f);                                                                 \
			key##KN = row_key(h, KN);    /* position inside the group; last j wins ties (ksw.c:473) */             \
			const int mo = M * one - oe_del * SCALE;              /* IMAD: keeps the subtraction off the ALU pipe */ \
			const int en = __vimax_s32_relu(e - e_del * SCALE, mo);   /* E(i+1, j), opened from M only */        \
			f = __vimax_s32_relu(f - e_ins * SCALE, mo);              /* F(i, j+1) */                            \
			p[(JJ) * 32] = (uint32_t)(en * 65536 + (int)((wv & 0xfu) | (uint32_t)h1));  /* {base, H(i,j-1), E} */ \
			h1 = h;                                                                                              \
		}
		int jj = 0;
		{   // cells below the narrowest live band: no predicates; loads run one 4-cell group ahead of the math
			int key0, key1, key2, key3, key4, key5, key6, key7;
			uint32_t b0, b1, b2, b3;
#pragma unroll 1
			for (; jj + 8 <= nfast; jj += 8) {
				b0 = p[(jj + 4) * 32]; b1 = p[(jj + 5) * 32]; b2 = p[(jj + 6) * 32]; b3 = p[(jj + 7) * 32];
				EMAB_LANES_CELL(a0, jj, 0) EMAB_LANES_CELL(a1, jj + 1, 1) EMAB_LANES_CELL(a2, jj + 2, 2) EMAB_LANES_CELL(a3, jj + 3, 3)
				if (jj + 12 <= nfast) { a0 = p[(jj + 8) * 32]; a1 = p[(jj + 9) * 32]; a2 = p[(jj + 10) * 32]; a3 = p[(jj + 11) * 32]; }
				EMAB_LANES_CELL(b0, jj + 4, 4) EMAB_LANES_CELL(b1, jj + 5, 5) EMAB_LANES_CELL(b2, jj + 6, 6) EMAB_LANES_CELL(b3, jj + 7, 7)
				/* row maximum: the group's best key (one ALU-pipe max per two cells), then its offset in the row */
				mkey = max(mkey, __vimax3_s32(__vimax3_s32(key0, key1, key2), __vimax3_s32(key3, key4, key5), max(key6, key7)) + (j + jj));
			}
			if (jj + 4 <= nfast) {
				EMAB_LANES_CELL(a0, jj, 0) EMAB_LANES_CELL(a1, jj + 1, 1) EMAB_LANES_CELL(a2, jj + 2, 2) EMAB_LANES_CELL(a3, jj + 3, 3)
				mkey = max(mkey, __vimax3_s32(__vimax3_s32(key0, key1, key2), key3, -1) + (j + jj));
				jj += 4;
			}
		}
#undef EMAB_LANES_CELL
		// cells between the narrowest and the widest band of the warp: four at a time, branch-free.  A lane whose
		// band is over computes on a zero word and its effects (store, row maximum, H carried to the next column)
		// are predicated off, so the four cells of a group still overlap in the pipes; the band is contiguous,
		// so a garbage f past a lane's last cell is never consumed.
#define EMAB_LANES_TAIL(PK, WV, JJ)                                                                              \
		{                                                                                                        \
			const uint32_t wv = (WV);                                                                            \
			const int diag = (int)(wv & 0xfff0u), e = (int)__umulhi(wv, 65536u);                                 \
			const uint32_t sb = prmt(rlo, rhi, wv);                                                              \
			const int sc = (int)prmt(sb, 0u, 0x8880u);                                                           \
			const int M = __viaddmin_s32(diag, sc, mul1024(diag));                                               \
			const int h = __vimax3_s32(M, e, f);                                                                 \
			const int key = row_key(h, j + (JJ));                                                                \
			mkey = max(mkey, (PK) ? key : -1);                                                                   \
			const int mo = M * one - oe_del * SCALE;              /* IMAD: keeps the subtraction off the ALU pipe */ \
			const int en = __vimax_s32_relu(e - e_del * SCALE, mo);                                              \
			f = __vimax_s32_relu(f - e_ins * SCALE, mo);                                                         \
			if (PK) p[(JJ) * 32] = (uint32_t)(en * 65536 + (int)((wv & 0xfu) | (uint32_t)h1));                   \
			h1 = (PK) ? h : h1;                                                                                  \
		}
#pragma unroll 1
		for (; jj < maxw; jj += 4) {
			const bool p0 = jj < width, p1 = jj + 1 < width, p2 = jj + 2 < width, p3 = jj + 3 < width;
			uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
			if (p0) c0 = p[jj * 32];
			if (p1) c1 = p[(jj + 1) * 32];
			if (p2) c2 = p[(jj + 2) * 32];
			if (p3) c3 = p[(jj + 3) * 32];
			EMAB_LANES_TAIL(p0, c0, jj) EMAB_LANES_TAIL(p1, c1, jj + 1) EMAB_LANES_TAIL(p2, c2, jj + 2) EMAB_LANES_TAIL(p3, c3, jj + 3)
		}
#undef EMAB_LANES_TAIL
		if (alive) {
			// eh[end] = {h1, 0} (bwa/ksw.c:485); the base of that column stays.  The first word of the band is
			// fetched right behind the store (beg may equal end) so that it is there when the band is trimmed.
			uint32_t *pe = eh + s.end * 32;
			*pe = (*pe & 0xfu) | (uint32_t)h1;
			const uint32_t wb = eh[s.beg * 32];
			const int h1u = h1 / SCALE;
			const int jfin = width > 0 ? s.end : s.beg;
			if (jfin == qlen) {                                            // later rows win ties (bwa/ksw.c:486-489)
				if (!(s.g > h1u)) s.g_i = i;
				s.g = s.g > h1u ? s.g : h1u;
			}
			const int m16 = mkey < 0 ? 0 : mkey / KEY_MUL, m = m16 / SCALE, mj = mkey < 0 ? -1 : mkey - m16 * KEY_MUL;
			// best cell / z-drop (bwa/ksw.c:490-500) without branches; e_del == e_ins (asserted above)
			const bool better = m > s.best;
			int gap = (i - s.best_i) - (mj - s.best_j); gap = gap < 0 ? -gap : gap;
			const bool drop = !better && zdrop > 0 && s.best - m - gap * e_del > zdrop;
			int d = mj - i; d = d < 0 ? -d : d;
			s.max_off = better && d > s.max_off ? d : s.max_off;
			s.best_i = better ? i : s.best_i;
			s.best_j = better ? mj : s.best_j;
			s.best = better ? m : s.best;
			alive = m != 0 && !drop;
			if (alive) {                                                   // bwa/ksw.c:502-505
				int a = s.beg;
				if (a < s.end && (wb & 0xfffffff0u) == 0) {
					++a;
					while (a < s.end && (eh[a * 32] & 0xfffffff0u) == 0) ++a;
				}
				s.beg = a;
				a = s.end;
				if (h1 == 0) {   // eh[end] was just written: {h1, 0}
					--a;
					while (a >= s.beg && (eh[a * 32] & 0xfffffff0u) == 0) --a;
				}
				s.end = a + 2 < qlen ? a + 2 : qlen;
			}
		}
	}
	visited += cells;
	ExtResult r;
	r.score = s.best; r.qle = s.best_j + 1; r.tle = s.best_i + 1; r.gtle = s.g_i + 1; r.gscore = s.g; r.max_off = s.max_off;
	return r;
}

struct BytesFetch {  // explicit byte string: base i = p[i * step]
	const uint8_t *p; int step;
	__device__ __forceinline__ int operator()(int i) const { return p[(int64_t)i * step]; }
};

}  // namespace lanes
