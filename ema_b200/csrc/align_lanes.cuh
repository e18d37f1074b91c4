// mem_align1_core after chaining (mem_chain2aln over every chain + mem_sort_dedup_patch,
// bwa/bwamem.c:1081-1117) with ONE THREAD PER READ and the ksw_extend2 calls of the 32 reads of a warp run
// together by the inter-task kernel of ksw_lanes.cuh.
//
// mem_chain2aln is sequential per read (every seed is tested against the regions earlier seeds produced,
// the right extension starts from the left one's score, a band retry depends on the first try), so the
// parallelism is across reads.  Each lane walks the reference's control flow as a small state machine built
// from the same pieces chain2aln uses (align.cuh) and stops whenever it needs a ksw_extend2 result; the
// warp then runs all pending extensions at once, one per lane, and every lane consumes its result.  A lane
// that finishes its read takes the next one from a global counter, so lanes always arrive with a task until
// the bucket runs dry.  Control code therefore costs one instruction slot per 32 reads instead of one per
// read (the warp-per-read kernel runs it redundantly on all lanes), and the DP costs ~18 instructions per
// cell instead of ~50.
#pragma once
#include "align.cuh"
#include "ksw_lanes.cuh"
#include "ksw_warp.cuh"

// Thread-scalar ksw_global2 score (bwa/ksw.c:540-622 without the backtrack), used for mem_patch_reg's
// score-only bwa_gen_cigar2 call (bwa/bwamem.c:448): rare (two colinear regions of one read), so it runs
// inline on the lane that needs it.  h/e live in local memory.
struct ScalarPatchDP {
	const DevIndex &ix;
	int *err;
	unsigned long long *cells;
	EMAB_HD ExtResult extend(const uint8_t *, int, int, int, int64_t, int, int, int, int, int)
	{
		*err = 4;  // never reached: mem_sort_dedup_patch only aligns globally
		return ExtResult{};
	}
	EMAB_HD LocResult local(const uint8_t *, int, int64_t, int)
	{
		*err = 4;
		return LocResult{};
	}
	EMAB_HD int ungapped(const uint8_t *query, int q0, int qstep, int n, int64_t t0, int tstep, int *score)
	{
		return ungapped_scalar(ix, query, q0, qstep, n, t0, tstep, score);
	}
	EMAB_HD int global(const uint8_t *query, int q0, int qstep, int qlen, int64_t t0, int tstep, int tlen, int w, uint32_t *cigar, int *n_cigar)
	{
		const int NEG = -0x40000000;
		const int e_del = opt::e_del, e_ins = opt::e_ins, oe_del = opt::oe_del, oe_ins = opt::oe_ins;
		if (cigar != nullptr || qlen > EMAB_MAX_READ_LEN) { *err = 4; if (n_cigar) *n_cigar = 0; return 0; }
		int H[EMAB_MAX_READ_LEN + 1], E[EMAB_MAX_READ_LEN + 1];
		H[0] = 0; E[0] = NEG;
		for (int j = 1; j <= qlen; ++j) { H[j] = j <= w ? -(opt::o_ins + e_ins * j) : NEG; E[j] = NEG; }
		unsigned long long visited = 0;
		for (int i = 0; i < tlen; ++i) {
			const int tb = ref_base(ix, t0 + (int64_t)i * tstep);
			const int beg = i > w ? i - w : 0;
			const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
			int h1 = beg == 0 ? -(opt::o_del + e_del * (i + 1)) : NEG, f = NEG;
			if (end > beg) visited += end - beg;
			for (int j = beg; j < end; ++j) {
				const int M = H[j] + sc_mat(tb, query[q0 + j * qstep]);
				int e = E[j];
				H[j] = h1;
				int h = M >= e ? M : e;
				h = h >= f ? h : f;
				h1 = h;
				int t = M - oe_del;
				e -= e_del;
				e = e > t ? e : t;
				E[j] = e;
				t = M - oe_ins;
				f -= e_ins;
				f = f > t ? f : t;
			}
			H[end] = h1; E[end] = NEG;
		}
#ifdef __CUDA_ARCH__
		if (cells && visited) atomicAdd(cells, visited);
#else
		if (cells) *cells += visited;
#endif
		return H[qlen];
	}
};

#ifdef __CUDACC__
namespace lanes {

struct QueryFetch {  // query base j of an ExtTask
	const uint8_t *q; int q0, qstep;
	__device__ __forceinline__ int operator()(int j) const { return q[q0 + j * qstep]; }
};
struct RefLaneFetch {  // target base i of an ExtTask, straight from the packed reference
	const DevIndex *ix; int64_t t0; int tstep;
	__device__ __forceinline__ int operator()(int i) const { return ref_base(*ix, t0 + (int64_t)i * tstep); }
};

enum { A_FETCH = 0, A_CHAIN, A_SEED, A_LEFT, A_RIGHT_BEGIN, A_RIGHT, A_IDLE };

// What a planned extension looks like to the thread-per-read walk: answered without touching the DP (ext_wave.cuh
// provides the implementation; a null cache answers nothing).
struct NoExtCache {
	static constexpr bool COOP_MISSES = false;   // every extension is for the warp to run: 32 tasks at a time, one per lane
	__device__ __forceinline__ void set_read(int, int) {}
	__device__ __forceinline__ bool find(const ExtTask &, ExtResult *, unsigned long long *) const { return false; }
};

// One warp: 32 reads in flight.  `next` is the global read counter, eh this lane's DP column 0.  `cache` answers the
// ksw_extend2 calls that were computed ahead (the bucket-wide waves); the others are gathered and run by the warp.
template <class Pools_, class Cache>
__device__ void align1_warp(const DevIndex &ix, int n_reads, const uint8_t *seq, const int64_t *off, const int32_t *occ_off, const Pools_ &p,
                            int rescue_room, uint32_t *eh, unsigned long long *next, int *err, unsigned long long *cells_ext,
                            unsigned long long *cells_glo, Cache cache, unsigned long long *n_inline)
{
	int st = A_FETCH;
	// read
	int r = 0, l_query = 0, n_chains = 0, ci = 0, n_av = 0;
	const uint8_t *query = nullptr;
	const Chain *chains = nullptr;
	const Seed *seeds_all = nullptr;
	uint64_t *srt = nullptr;
	Reg *regs = nullptr;
	// chain
	ChainWin cw; cw.rmax0 = cw.rmax1 = 0;
	const Seed *seeds = nullptr;
	int k = -1, cn = 0;
	// seed
	SeedExt e;
	seed_begin(Chain{}, e);
	Seed s{};
	int t_try = 0;
	ExtTask x{};
	ExtResult res{};
	unsigned long long visited = 0, cached_cells = 0, inline_calls = 0;
	// what happens once the result of the pending extension (state A_LEFT / A_RIGHT, task x) is known
	auto consume = [&](const ExtResult &rr) {
		if (st == A_LEFT) {
			if (left_try_done(e, rr, t_try)) { ++t_try; x = left_task(s, cw, t_try); }   // again, with the band doubled
			else { left_finish(s, e, rr); st = A_RIGHT_BEGIN; }
		} else {  // A_RIGHT
			if (right_try_done(e, rr, t_try)) { ++t_try; x = right_task(l_query, s, cw, e, t_try); }
			else {
				right_finish(l_query, s, cw, e, rr);
				seed_finish(chains[ci], seeds, s, e, regs, &n_av);
				--k; st = A_SEED;
			}
		}
	};
	for (;;) {
		bool req = false;
		while (!req && st != A_IDLE) {
			if (st == A_LEFT || st == A_RIGHT) {  // an extension is pending: planned ahead, or for the warp to run
				unsigned long long cc = 0;
				if (cache.find(x, &res, &cc)) { cached_cells += cc; consume(res); continue; }
				req = true;
				break;
			}
			if (st == A_FETCH) {
				const unsigned long long rr = atomicAdd(next, 1ull);
				if (rr >= (unsigned long long)n_reads) { st = A_IDLE; break; }
				r = (int)rr;
				const int o = occ_off[r];
				l_query = (int)(off[r + 1] - off[r]);
				query = seq + off[r];
				chains = p.chains + o; seeds_all = p.seeds + o; srt = p.srt + o;
				regs = p.regs + (o + (size_t)rescue_room * r);
				n_chains = p.n_chains[r];
				cache.set_read(r, n_chains);
				ci = 0; n_av = 0;
				st = A_CHAIN;
			} else if (st == A_CHAIN) {
				if (ci >= n_chains) {  // all chains extended: mem_sort_dedup_patch, then the next read
					ScalarPatchDP sdp{ix, err, cells_glo};
					p.n_regs[r] = sort_dedup_patch(ix, sdp, query, n_av, regs);
					st = A_FETCH;
				} else {
					const Chain &c = chains[ci];
					cn = c.n;
					if (cn == 0) { ++ci; continue; }
					seeds = seeds_all + c.seed_beg;
					cw = chain_window(ix, l_query, c, seeds);
					chain_sort_seeds(c, seeds, srt);
					k = cn - 1;
					st = A_SEED;
				}
			} else if (st == A_SEED) {
				if (k < 0) { ++ci; st = A_CHAIN; continue; }
				const Chain &c = chains[ci];
				if (!seed_wants_extension(l_query, c, seeds, srt, k, regs, n_av)) { --k; continue; }
				s = seeds[(uint32_t)srt[k]];
				seed_begin(c, e);
				if (s.qbeg) { t_try = 0; x = left_task(s, cw, 0); st = A_LEFT; }
				else { left_none(s, e); st = A_RIGHT_BEGIN; }
			} else if (st == A_RIGHT_BEGIN) {
				if (s.qbeg + s.len != l_query) {
					e.sc0 = e.a.score;
					t_try = 0; x = right_task(l_query, s, cw, e, 0); st = A_RIGHT;
				} else {
					right_none(l_query, s, e);
					seed_finish(chains[ci], seeds, s, e, regs, &n_av);
					--k; st = A_SEED;
				}
			}
		}
		unsigned pend = __ballot_sync(FULL_MASK, req);
		if (!pend) break;   // every lane idle: the bucket is done
		if (Cache::COOP_MISSES) {
			// A replay asks for an extension the waves did not compute a few times per bucket.  One lane running it alone
			// (lanes::extend) is ~0.3 ms of a whole warp waiting; the 32 lanes running it TOGETHER (warp_extend: a row at a
			// time) are done in a few tens of microseconds.  The warp's shared memory is free between extensions.
			WarpDP &wsm = *reinterpret_cast<WarpDP *>(eh - (threadIdx.x & 31));
			const int lane = threadIdx.x & 31;
			while (pend) {
				const int src = __ffs(pend) - 1;
				pend &= pend - 1;
				const unsigned long long qp = __shfl_sync(FULL_MASK, (unsigned long long)(uintptr_t)query, src);
				const int q0 = __shfl_sync(FULL_MASK, x.q0, src), qstep = __shfl_sync(FULL_MASK, x.qstep, src), qlen = __shfl_sync(FULL_MASK, x.qlen, src);
				const long long t0 = __shfl_sync(FULL_MASK, (long long)x.t0, src);
				const int tstep = __shfl_sync(FULL_MASK, x.tstep, src), tlen = __shfl_sync(FULL_MASK, x.tlen, src);
				const int h0 = __shfl_sync(FULL_MASK, x.h0, src), w = __shfl_sync(FULL_MASK, x.w, src), eb = __shfl_sync(FULL_MASK, x.end_bonus, src);
				const uint8_t *qsrc = reinterpret_cast<const uint8_t *>((uintptr_t)qp);
				__syncwarp();
				for (int jq = lane; jq < qlen; jq += 32) wsm.q[jq] = qsrc[q0 + jq * qstep];
				__syncwarp();
				RefFetch rf{&ix, (int64_t)t0, tstep};
				const ExtResult rr = warp_extend(wsm, qlen, rf, tlen, w, eb, opt::zdrop, h0, cells_ext);
				if (lane == src) res = rr;
				__syncwarp();
			}
		} else {
			QueryFetch qf{query, x.q0, x.qstep};
			RefLaneFetch tf{&ix, x.t0, x.tstep};
			res = extend(eh, req, x.qlen, x.tlen, x.h0, x.w, x.end_bonus, opt::zdrop, qf, tf, visited);
		}
		if (req) { ++inline_calls; consume(res); }
	}
	visited += cached_cells;
	for (int d = 16; d; d >>= 1) { visited += __shfl_xor_sync(FULL_MASK, visited, d); inline_calls += __shfl_xor_sync(FULL_MASK, inline_calls, d); }
	if ((threadIdx.x & 31) == 0 && visited) atomicAdd(cells_ext, visited);
	if ((threadIdx.x & 31) == 0 && inline_calls && n_inline) atomicAdd(n_inline, inline_calls);
}

}  // namespace lanes
#endif
