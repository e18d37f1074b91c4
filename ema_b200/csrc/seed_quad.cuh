// SMEM seeding with FOUR LANES PER READ (device only): the same loops as seed.cuh (seed_p12 / seed_p3), with
//   * the FM-index step spread over a quad: lane q loads the q-th 16 bytes of the 64-byte Occ block, so a read's block
//     is ONE coalesced 64-byte request (two sectors) instead of four 16-byte requests of one lane, each a wavefront of
//     its own in the L1 pipe; lanes 0/1 contribute the cumulative counts they loaded, lanes 2/3 the popcounts of their
//     64 symbols, and three 64-bit quad sums give what bwt_extend (bwa/bwt.c:262-275) needs for base c.  The second
//     block is requested only when l lies in another block than k;
//   * bwt_smem1a's prev/curr lists packed to 16 bytes per interval in SHARED memory (the thread-per-read form kept
//     them in a global slab: a dependent L2/DRAM round trip in front of every backward step and 0.75 GB of DRAM
//     writes per bucket at 3.1 Gbp, profiles/r2c_ncu_summary_c3.md);
//   * a quarter of the registers per lane for block data, hence more resident reads per register.
// The loops run redundantly on the four lanes of a quad (identical state); memory side effects are issued once.
#pragma once
#include "seed.cuh"

#ifdef __CUDACC__
#define SEEDQ_CAP 20                        // intervals per list kept in shared memory; longer lists continue in a global slab
#define SEEDQ_STRIDE (2 * SEEDQ_CAP + 1)    // uint4 per quad; odd, so neighbouring quads start in different banks

struct QuadFm {
	const DevIndex &ix;
	int q;               // lane within the quad
	unsigned qmask;      // the quad's four lanes
	unsigned touches;

	__device__ const DevIndex &index() const { return ix; }

	// partial counts of one position from this lane's 16 bytes of its block: out[i] = this lane's share of Occ(i, pos)
	__device__ __forceinline__ void contrib(const uint4 &d, int idx, bool valid, uint64_t out[4]) const
	{
		const int n = idx + 1, n2 = n << 1;
		const int base2 = (q - 2) * 128;               // bits before this lane's first symbol (lanes 2, 3)
		uint32_t nhi = 0, nlo = 0, nt = 0;
		pair_counts(d.x, d.y, n2, base2, nhi, nlo, nt);
		pair_counts(d.z, d.w, n2, base2 + 64, nhi, nlo, nt);
		int mine = n - (q - 2) * 64;                     // symbols of the prefix that fall into this lane's 64
		mine = mine < 0 ? 0 : (mine > 64 ? 64 : mine);
		const uint32_t ng = nhi - nt, nc = nlo - nt, na = (uint32_t)mine - nc - ng - nt;
		const uint64_t lo = (uint64_t)d.y << 32 | d.x, hi = (uint64_t)d.w << 32 | d.z;   // the two u64 counts of lanes 0, 1
		const bool words = q >= 2;
		out[0] = words ? na : (q == 0 ? lo : 0);
		out[1] = words ? nc : (q == 0 ? hi : 0);
		out[2] = words ? ng : (q == 1 ? lo : 0);
		out[3] = words ? nt : (q == 1 ? hi : 0);
#pragma unroll
		for (int i = 0; i < 4; ++i) out[i] = valid ? out[i] : 0;
	}

	__device__ __forceinline__ uint64_t quad_sum(uint64_t v) const
	{
		v += __shfl_xor_sync(qmask, v, 1);
		v += __shfl_xor_sync(qmask, v, 2);
		return v;
	}

	// the interval of base c after bwt_extend(ik) (bwa/bwt.c:262-275 + bwt_2occ4 :189-220); info is left 0
	__device__ Intv extend1(const Intv &ik, int c, int is_back, bool = true)
	{
		const uint64_t NEG1 = ~0ull;
		const uint64_t xa = is_back ? ik.x0 : ik.x1, xb = is_back ? ik.x1 : ik.x0;
		const uint64_t k = xa - 1, l = xa - 1 + ik.x2;
		const bool kv = k != NEG1, lv = l != NEG1;
		const uint64_t _k = kv ? k - (k >= ix.primary) : 0, _l = lv ? l - (l >= ix.primary) : 0;
		const uint64_t bk = _k >> 7, bl = _l >> 7;
		const uint4 a = ldg128(ix.bwt + (bk << 2) + q);
		const bool other = bk != bl || ix.seed_load_both;
		uint4 b = a;
		ldg128_if(other, ix.bwt + (bl << 2) + q, b);
		if (q == 0) touches += (unsigned)kv + (unsigned)(lv && !(kv && bk == bl));
		uint64_t ck[4], cl[4];
		contrib(a, (int)(_k & 127), kv, ck);
		contrib(b, (int)(_l & 127), lv, cl);
		// what base c needs: Occ(c, k), the size of its interval, and the sizes of the intervals of the larger bases
		uint64_t tk = ck[0], ns = cl[0] - ck[0], above = 0;
#pragma unroll
		for (int i = 1; i < 4; ++i) {
			const uint64_t d = cl[i] - ck[i];
			if (c == i) { tk = ck[i]; ns = d; }
			if (c < i) above += d;
		}
		tk = quad_sum(tk); ns = quad_sum(ns); above = quad_sum(above);
		const uint64_t na = ix.L2[c] + 1 + tk;
		const uint64_t nb = xb + (xa <= ix.primary && xa + ik.x2 - 1 >= ix.primary) + above;
		Intv ok;
		ok.x0 = is_back ? na : nb;
		ok.x1 = is_back ? nb : na;
		ok.x2 = ns;
		ok.info = 0;
		return ok;
	}
};

// prev/curr of bwt_smem1a: 16-byte packed intervals (three 39-bit coordinates + the 9-bit end position) in the quad's
// shared-memory slice; entries past SEEDQ_CAP go to the quad's global slab as plain Intv.
struct QuadLists {
	uint4 *sm;           // [2][SEEDQ_CAP] (+1 pad)
	Intv *slab;          // [2][slab_len]
	int slab_len, q;
	unsigned qmask;

	static __device__ __forceinline__ uint4 pack(const Intv &v)
	{
		const uint64_t M39 = (1ull << 39) - 1;
		const uint64_t A = (v.x0 & M39) | (v.x2 & 0x1ffffffull) << 39;
		const uint64_t B = (v.x1 & M39) | ((v.x2 >> 25) & 0x3fffull) << 39 | (v.info & 0x1ffull) << 53;
		return make_uint4((uint32_t)A, (uint32_t)(A >> 32), (uint32_t)B, (uint32_t)(B >> 32));
	}
	static __device__ __forceinline__ Intv unpack(const uint4 &u)
	{
		const uint64_t M39 = (1ull << 39) - 1;
		const uint64_t A = (uint64_t)u.y << 32 | u.x, B = (uint64_t)u.w << 32 | u.z;
		Intv v;
		v.x0 = A & M39; v.x1 = B & M39;
		v.x2 = (A >> 39) | ((B >> 39) & 0x3fffull) << 25;
		v.info = B >> 53;
		return v;
	}
	// all four lanes store the same value to the same address: a lane always reads back at least its own write
	__device__ __forceinline__ void put(int which, int idx, const Intv &v)
	{
		if (idx < SEEDQ_CAP) sm[which * SEEDQ_CAP + idx] = pack(v);
		else slab[(size_t)which * slab_len + idx] = v;
	}
	__device__ __forceinline__ Intv get(int which, int idx) const
	{
		if (idx < SEEDQ_CAP) return unpack(sm[which * SEEDQ_CAP + idx]);
		return slab[(size_t)which * slab_len + idx];
	}
	// an output interval: lane q stores component q, the quad one 32-byte sector
	__device__ __forceinline__ void emit(Intv *out, int idx, const Intv &v)
	{
		const uint64_t comp = q == 0 ? v.x0 : (q == 1 ? v.x1 : (q == 2 ? v.x2 : v.info));
		((uint64_t *)(out + idx))[q] = comp;
	}
	// one lane of the quad stores its own value (the others hold different ones); readers sync() first
	__device__ __forceinline__ void put_by(int owner, int which, int idx, const Intv &v) { if (q == owner) put(which, idx, v); }
	__device__ __forceinline__ void emit_by(int owner, Intv *out, int idx, const Intv &v) { if (q == owner) out[idx] = v; }
	__device__ __forceinline__ void sync() const { __syncwarp(qmask); }   // before reading what another lane stored
};

struct QuadCoop {
	static constexpr int WIDTH = 4;
	int q;
	unsigned qmask;
	__device__ __forceinline__ int lane() const { return q; }
	__device__ __forceinline__ uint64_t bcast(uint64_t v, int t) const { return __shfl_sync(qmask, v, ((threadIdx.x & 31) & ~3) + t); }
};

struct QuadFeeder {
	const SeedBatch &b;
	int which;  // 0: pass 1/2, 1: pass 3
	int q;
	unsigned qmask;
	__device__ bool next(SeedJob &o)
	{
		unsigned long long r = 0;
		if (q == 0) r = atomicAdd(&b.queue[which], 1ull);
		r = __shfl_sync(qmask, r, (threadIdx.x & 31) & ~3);
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
		if (q == 0) {
			(which == 0 ? b.n12 : b.n3)[j.id] = n;
			if (ovf) *b.err = 3;
		}
	}
};

// ---------------------------------------------------------------------------------------------
// FOUR LANES PER READ, ONE LIST ENTRY PER LANE (the form the pipeline runs).  What bounds seeding is not bandwidth but
// the longest reads' chains of dependent bwt_extend calls (a kernel of 3 blocks per SM is as fast as one of 6,
// profiles/r2f_*), and most links of those chains are backward rounds of bwt_smem1a over a dozen intervals whose
// extensions do not depend on each other.  A quad takes four entries of a round per step — each lane a whole
// bwt_extend of its own entry, registers and loads as in the one-lane form — and exchanges only the interval sizes
// the bookkeeping needs (QuadCoop::bcast).  On BASELINE configs[1] reads that shortens the passes-1/2 chain from 457
// to 257 steps on average and from 1369 to 570 at the maximum (tests/hostsim hs_seed_profile).  Forward sweeps have one
// extension per step: the four lanes compute it together (same addresses, one request).  Pass 3 has no such
// parallelism and keeps one lane per read on its own warps.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void seed_quads_wide(const DevIndex &ix, const SeedBatch &b, uint4 *smem)
{
	const int quad = threadIdx.x >> 2, q = threadIdx.x & 3;
	const unsigned qmask = 0xfu << ((threadIdx.x & 31) & ~3);
	const size_t gquad = (size_t)blockIdx.x * (blockDim.x >> 2) + quad;
	Fm fm{ix, 0};
	ThreadFm tfm{fm};
	QuadLists lists{smem + quad * SEEDQ_STRIDE, b.scratch + gquad * 2 * b.scratch_len, b.scratch_len, q, qmask};
	QuadCoop coop{q, qmask};
	QuadFeeder f12{b, 0, q, qmask};
	QueueFeeder f3{b, 1};
	PtrLists none{{nullptr, nullptr}};
	const unsigned gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // roles by warp whatever the block size: every fourth warp starts on pass 3
	if ((gwarp & 3) == 3) { seed_p3(tfm, f3, none); seed_p12(tfm, f12, lists, coop); }
	else { seed_p12(tfm, f12, lists, coop); seed_p3(tfm, f3, none); }
	unsigned touches = fm.touches;
	for (int d = 16; d; d >>= 1) touches += __shfl_xor_sync(0xffffffffu, touches, d);
	if ((threadIdx.x & 31) == 0 && touches) atomicAdd(b.touches, (unsigned long long)touches);
}

// ---------------------------------------------------------------------------------------------
// One lane per read, Occ blocks STAGED IN SHARED MEMORY by cp.async: the eight 16-byte pieces of a step's two blocks
// travel global -> shared without occupying destination registers while they are in flight (the register form holds
// 32 registers for them, which is what pins the kernel at 80 registers / 24 warps per SM), and are then consumed
// piece by piece.  Slot layout: piece p of thread t at [(p * blockDim + t)] as uint4 — every warp access is 512
// consecutive bytes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint4 *dst_smem, const uint4 *src, bool pred)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q cp.async.ca.shared.global [%0], [%1], 16;\n\t}" ::"r"(d), "l"(src), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
	asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct StagedFm {
	Fm &fm;
	uint4 *slot;         // this thread's piece 0; piece p at slot[p * stride]
	int stride;          // blockDim.x
	__device__ const DevIndex &index() const { return fm.ix; }

	__device__ __forceinline__ void occ4(int base_piece, int idx, bool valid, uint64_t cnt[4]) const
	{
		uint32_t nhi = 0, nlo = 0, nt = 0;
		const int n = idx + 1, n2 = n << 1;
		{
			const uint4 w = slot[(base_piece + 2) * stride];
			pair_counts(w.x, w.y, n2, 0, nhi, nlo, nt);
			pair_counts(w.z, w.w, n2, 64, nhi, nlo, nt);
		}
		{
			const uint4 w = slot[(base_piece + 3) * stride];
			pair_counts(w.x, w.y, n2, 128, nhi, nlo, nt);
			pair_counts(w.z, w.w, n2, 192, nhi, nlo, nt);
		}
		const uint32_t ng = nhi - nt, nc = nlo - nt, na = (uint32_t)n - nc - ng - nt;
		const uint4 c0 = slot[base_piece * stride], c1 = slot[(base_piece + 1) * stride];
		cnt[0] = valid ? ((uint64_t)c0.y << 32 | c0.x) + na : 0;
		cnt[1] = valid ? ((uint64_t)c0.w << 32 | c0.z) + nc : 0;
		cnt[2] = valid ? ((uint64_t)c1.y << 32 | c1.x) + ng : 0;
		cnt[3] = valid ? ((uint64_t)c1.w << 32 | c1.z) + nt : 0;
	}

	__device__ Intv extend1(const Intv &ik, int c, int is_back, bool counted = true)
	{
		const DevIndex &ix = fm.ix;
		const uint64_t NEG1 = ~0ull;
		const uint64_t xa = is_back ? ik.x0 : ik.x1, xb = is_back ? ik.x1 : ik.x0;
		const uint64_t k = xa - 1, l = xa - 1 + ik.x2;
		const bool kv = k != NEG1, lv = l != NEG1;
		const uint64_t _k = kv ? k - (k >= ix.primary) : 0, _l = lv ? l - (l >= ix.primary) : 0;
		const uint64_t bk = _k >> 7, bl = _l >> 7;
		const uint4 *pk = ix.bwt + (bk << 2), *pl = ix.bwt + (bl << 2);
		const bool other = bk != bl || ix.seed_load_both;
#pragma unroll
		for (int p = 0; p < 4; ++p) cp_async16(slot + p * stride, pk + p, true);
#pragma unroll
		for (int p = 0; p < 4; ++p) cp_async16(slot + (4 + p) * stride, pl + p, other);
		if (counted) fm.touches += (unsigned)kv + (unsigned)(lv && !(kv && bk == bl));
		cp_async_wait_all();
		uint64_t tk[4], tl[4];
		occ4(0, (int)(_k & 127), kv, tk);
		occ4(other ? 4 : 0, (int)(_l & 127), lv, tl);
		uint64_t na = tk[0], ns = tl[0] - tk[0], above = 0;
#pragma unroll
		for (int i = 1; i < 4; ++i) {
			const uint64_t d = tl[i] - tk[i];
			if (c == i) { na = tk[i]; ns = d; }
			if (c < i) above += d;
		}
		na += ix.L2[c] + 1;
		const uint64_t nb = xb + (xa <= ix.primary && xa + ik.x2 - 1 >= ix.primary) + above;
		Intv ok;
		ok.x0 = is_back ? na : nb;
		ok.x1 = is_back ? nb : na;
		ok.x2 = ns;
		ok.info = 0;
		return ok;
	}
};
#define SEEDS_SMEM(threads) ((threads) * 8 * 16)

__device__ __forceinline__ void seed_staged(const DevIndex &ix, const SeedBatch &b, uint4 *smem)
{
	const int gt = blockIdx.x * blockDim.x + threadIdx.x;
	Fm fm{ix, 0};
	StagedFm sfm{fm, smem + threadIdx.x, (int)blockDim.x};
	Intv *buf0 = b.scratch + (size_t)gt * 2 * b.scratch_len;
	PtrLists lists{{buf0, buf0 + b.scratch_len}};
	QueueFeeder f12{b, 0}, f3{b, 1};
	if (((threadIdx.x >> 5) & 3) == 3) { seed_p3(sfm, f3, lists); seed_p12(sfm, f12, lists, SoloCoop()); }
	else { seed_p12(sfm, f12, lists, SoloCoop()); seed_p3(sfm, f3, lists); }
	unsigned touches = fm.touches;
	for (int d = 16; d; d >>= 1) touches += __shfl_xor_sync(0xffffffffu, touches, d);
	if ((threadIdx.x & 31) == 0 && touches) atomicAdd(b.touches, (unsigned long long)touches);
}

// shared memory of one 128-thread block of k_seed_quad
#define SEEDQ_SMEM (32 * SEEDQ_STRIDE * 16)

__device__ __forceinline__ void seed_quads(const DevIndex &ix, const SeedBatch &b, uint4 *smem)
{
	const int quad = threadIdx.x >> 2, q = threadIdx.x & 3;
	const unsigned qmask = 0xfu << ((threadIdx.x & 31) & ~3);
	const size_t gquad = (size_t)blockIdx.x * (blockDim.x >> 2) + quad;
	QuadFm fm{ix, q, qmask, 0};
	QuadLists lists{smem + quad * SEEDQ_STRIDE, b.scratch + gquad * 2 * b.scratch_len, b.scratch_len, q, qmask};
	QuadFeeder f12{b, 0, q, qmask}, f3{b, 1, q, qmask};
	if (((threadIdx.x >> 5) & 3) == 3) { seed_p3(fm, f3, lists); seed_p12(fm, f12, lists, SoloCoop()); }
	else { seed_p12(fm, f12, lists, SoloCoop()); seed_p3(fm, f3, lists); }
	unsigned touches = fm.touches;
	for (int d = 16; d; d >>= 1) touches += __shfl_xor_sync(0xffffffffu, touches, d);
	if ((threadIdx.x & 31) == 0 && touches) atomicAdd(b.touches, (unsigned long long)touches);
}
#endif
