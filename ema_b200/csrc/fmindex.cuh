// FM-index walks over the BWT/Occ blocks resident in HBM.
// Semantics: bwa/bwt.c:169-220 (bwt_occ4 / bwt_2occ4), :262-275 (bwt_extend), :53-59,86-96 (bwt_sa).
#pragma once
#include "common.cuh"

// One 64-byte Occ block = 4 x 128-bit loads.  ld.global.nc keeps the (read-only) index on the
// non-coherent path; blocks are 64-byte aligned so each load is one fully used 32-byte sector pair.
EMAB_HD uint4 ldg128(const uint4 *p)
{
#ifdef __CUDA_ARCH__
	uint4 r;
	asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
#else
	return *p;
#endif
}

struct Fm {  // per-thread view of the index + instrumentation (64-byte block loads)
	const DevIndex &ix;
	unsigned touches;
};

struct OccBlock {
	uint4 c0, c1;  // cumulative counts of A,C / G,T before the block (4 x u64)
	uint4 b0, b1;  // 8 x u32, 16 symbols each, symbol i at bits ((~i)&15)<<1
};

EMAB_HD OccBlock load_block(Fm &fm, uint64_t blk)
{
	const uint4 *p = fm.ix.bwt + (blk << 2);
	++fm.touches;
	OccBlock o;
	o.c0 = ldg128(p); o.c1 = ldg128(p + 1); o.b0 = ldg128(p + 2); o.b1 = ldg128(p + 3);
	return o;
}

// Mask keeping the first n symbols (MSB first) of the 16-symbol word that starts at symbol `base`
// of the block: n2 = 2 * (number of symbols to count in the whole block).
EMAB_HD uint32_t prefix_mask(int n2, int base2)
{
	int s = n2 - base2;            // 2 * symbols of this word inside the prefix; may be < 0 or > 32
	s = s < 0 ? 0 : s;
#ifdef __CUDA_ARCH__
	uint32_t r;
	asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(0xffffffffu), "r"(s));   // PTX shr clamps shifts >= 32 to "all out"
	return ~r;
#else
	return s >= 32 ? 0xffffffffu : ~(0xffffffffu >> s);
#endif
}

// Symbols are 2 bits, so the "high bit" / "low bit" planes of a word occupy the even bit positions
// only: the planes of TWO words are interleaved into one register (word A on even bits, word B on
// odd bits) and counted with one popc each — 3 popc per 32 symbols instead of 6.
EMAB_HD void pair_counts(uint32_t wa, uint32_t wb, int n2, int base2, uint32_t &nhi, uint32_t &nlo, uint32_t &nt)
{
	wa &= prefix_mask(n2, base2);
	wb &= prefix_mask(n2, base2 + 32);
	const uint32_t hi = ((wa >> 1) & 0x55555555u) | (wb & 0xaaaaaaaau);
	const uint32_t lo = (wa & 0x55555555u) | ((wb << 1) & 0xaaaaaaaau);
	nhi += emab_popc(hi);
	nlo += emab_popc(lo);
	nt += emab_popc(hi & lo);
}

// Occ of all four bases in B[0..k] inclusive, k already adjusted for the primary and != -1.
// `idx` = k & 127 within block `o`.
EMAB_HD void block_occ4(const OccBlock &o, int idx, uint64_t cnt[4])
{
	uint32_t nhi = 0, nlo = 0, nt = 0;
	const int n = idx + 1, n2 = n << 1;  // symbols to count
	pair_counts(o.b0.x, o.b0.y, n2, 0, nhi, nlo, nt);
	pair_counts(o.b0.z, o.b0.w, n2, 64, nhi, nlo, nt);
	pair_counts(o.b1.x, o.b1.y, n2, 128, nhi, nlo, nt);
	pair_counts(o.b1.z, o.b1.w, n2, 192, nhi, nlo, nt);
	const uint32_t ng = nhi - nt, nc = nlo - nt;   // 2 = G (hi only), 1 = C (lo only), 3 = T (both)
	const uint32_t na = (uint32_t)n - nc - ng - nt;
	cnt[0] = ((uint64_t)o.c0.y << 32 | o.c0.x) + na;
	cnt[1] = ((uint64_t)o.c0.w << 32 | o.c0.z) + nc;
	cnt[2] = ((uint64_t)o.c1.y << 32 | o.c1.x) + ng;
	cnt[3] = ((uint64_t)o.c1.w << 32 | o.c1.z) + nt;
}

// 128-bit load issued only when `pred` holds; otherwise r keeps the value it came in with.  A predicated-off lane
// issues no memory request at all: no sector, no L1 wavefront.
EMAB_HD void ldg128_if(bool pred, const uint4 *p, uint4 &r)
{
#ifdef __CUDA_ARCH__
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
	             : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p), "r"((unsigned)pred));
#else
	if (pred) r = *p;
#endif
}

// bwt_2occ4(k, l): two positions (bwa/bwt.c:189-220).  Branch-free: with one thread per read the lanes of a
// warp disagree on "k and l share a block", and a branch here made the warp run both arms of the ~200-instruction
// counting code every step (ncu, profiles/r1b_ncu_summary_c2.md).  Both positions are always counted, but the
// second block is REQUESTED only by the lanes whose l lies in another block than k (predicated loads): once an
// interval is narrow — most steps of a read — k and l share a block, and with one read per lane every request is a
// wavefront of its own in the L1 pipe, the unit this kernel runs out of first (profiles/r2*_ncu_seed*.md).
// fm.touches counts 64-byte block loads as the reference would issue them (the roofline unit of
// SURVEY.md §8d): one when k and l share a block, none for a position equal to (bwtint_t)-1.
EMAB_HD void bwt_2occ4(Fm &fm, uint64_t k, uint64_t l, uint64_t ck[4], uint64_t cl[4])
{
	const uint64_t NEG1 = ~0ull;
	const bool kv = k != NEG1, lv = l != NEG1;
	const uint64_t _k = kv ? k - (k >= fm.ix.primary) : 0, _l = lv ? l - (l >= fm.ix.primary) : 0;
	const uint64_t bk = _k >> 7, bl = _l >> 7;
	const uint4 *pk = fm.ix.bwt + (bk << 2), *pl = fm.ix.bwt + (bl << 2);
	OccBlock a, b;
	a.c0 = ldg128(pk); a.c1 = ldg128(pk + 1); a.b0 = ldg128(pk + 2); a.b1 = ldg128(pk + 3);
	const bool other = bk != bl || fm.ix.seed_load_both;
	b.c0 = b.c1 = b.b0 = b.b1 = make_uint4(0, 0, 0, 0);
	ldg128_if(other, pl, b.c0); ldg128_if(other, pl + 1, b.c1); ldg128_if(other, pl + 2, b.b0); ldg128_if(other, pl + 3, b.b1);
	fm.touches += (unsigned)kv + (unsigned)(lv && !(kv && bk == bl));
	block_occ4(a, (int)(_k & 127), ck);
	if (!other) b = a;
	block_occ4(b, (int)(_l & 127), cl);
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		ck[i] = kv ? ck[i] : 0;
		cl[i] = lv ? cl[i] : 0;
	}
}

EMAB_HD void bwt_set_intv(const DevIndex &ix, int c, Intv &ik)
{  // bwa/bwt.h:82
	ik.x0 = ix.L2[c] + 1;
	ik.x2 = ix.L2[c + 1] - ix.L2[c];
	ik.x1 = ix.L2[3 - c] + 1;
	ik.info = 0;
}

// bwt_extend (bwa/bwt.c:262-275).  x[] is indexed as {x0,x1,x2}; is_back selects which coordinate
// is the one being walked.  ok[c].info is left untouched (callers set it).
EMAB_HD void bwt_extend(Fm &fm, const Intv &ik, Intv ok[4], int is_back)
{
	uint64_t tk[4], tl[4];
	uint64_t xa = is_back ? ik.x0 : ik.x1;  // x[!is_back]
	uint64_t xb = is_back ? ik.x1 : ik.x0;  // x[is_back]
	bwt_2occ4(fm, xa - 1, xa - 1 + ik.x2, tk, tl);
	uint64_t na[4], nb[4], ns[4];
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		na[i] = fm.ix.L2[i] + 1 + tk[i];
		ns[i] = tl[i] - tk[i];
	}
	nb[3] = xb + (xa <= fm.ix.primary && xa + ik.x2 - 1 >= fm.ix.primary);
	nb[2] = nb[3] + ns[3];
	nb[1] = nb[2] + ns[2];
	nb[0] = nb[1] + ns[1];
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		if (is_back) { ok[i].x0 = na[i]; ok[i].x1 = nb[i]; }
		else { ok[i].x1 = na[i]; ok[i].x0 = nb[i]; }
		ok[i].x2 = ns[i];
	}
}

// Only the interval for base c (what every caller on the path consumes).
EMAB_HD Intv bwt_extend1(Fm &fm, const Intv &ik, int c, int is_back)
{
	Intv ok[4];
	bwt_extend(fm, ik, ok, is_back);
	Intv r = ok[0];
	if (c == 1) r = ok[1];
	if (c == 2) r = ok[2];
	if (c == 3) r = ok[3];
	return r;
}

// bwt_sa through the dense suffix array: identical values to bwa/bwt.c:86-96 with no LF walk
// (bwt_sa is a pure function of the index; SURVEY.md §7 hard part 6).
EMAB_HD uint64_t bwt_sa_dense(const DevIndex &ix, uint64_t k)
{
	if (k == 0) return ~0ull;  // SA[0] = (u64)-1 by construction (bwa/bwt.c:83,437)
	return ix.sa32 ? (uint64_t)ix.sa32[k] : ix.sa64[k];
}

// bwt_invPsi (bwa/bwt.c:53-59): LF-mapping step, used by the dense-SA builder and the walking bwt_sa.
EMAB_HD uint64_t bwt_invPsi(Fm &fm, uint64_t k)
{
	if (k == fm.ix.primary) return 0;
	uint64_t x = k - (k > fm.ix.primary);
	OccBlock o = load_block(fm, x >> 7);
	int idx = (int)(x & 127);
	uint32_t w;
	{
		int wi = idx >> 4;
		uint32_t ws[8] = {o.b0.x, o.b0.y, o.b0.z, o.b0.w, o.b1.x, o.b1.y, o.b1.z, o.b1.w};
		w = ws[0];
#pragma unroll
		for (int j = 1; j < 8; ++j) if (wi == j) w = ws[j];
	}
	int c = (w >> ((~idx & 15) << 1)) & 3;
	// occ(k, c): bwt_occ adjusts k by (k >= primary); k != primary here so x is that adjusted value
	uint64_t cnt[4];
	block_occ4(o, idx, cnt);
	uint64_t occ = cnt[0];
	if (c == 1) occ = cnt[1];
	if (c == 2) occ = cnt[2];
	if (c == 3) occ = cnt[3];
	return fm.ix.L2[c] + occ;
}

// bwt_sa by walking to a sampled slot (bwa/bwt.c:86-96); kept for parity tests and as the
// builder's cross-check.
EMAB_HD uint64_t bwt_sa_walk(Fm &fm, uint64_t k)
{
	uint64_t sa = 0, mask = (uint64_t)fm.ix.sa_intv - 1;
	while (k & mask) { ++sa; k = bwt_invPsi(fm, k); }
	return sa + fm.ix.sa_sampled[k / fm.ix.sa_intv];
}
