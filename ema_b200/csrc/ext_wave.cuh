// Inter-read wave scheduling of mem_chain2aln's ksw_extend2 calls (bwa/bwamem.c:693-810, bwa/ksw.c:416-515).
//
// mem_chain2aln is sequential per read: whether a seed is extended depends on the regions earlier seeds produced.
// But WHAT an extension computes does not: the left extension of a seed is a pure function of (read, seed, chain
// window), the right extension of those plus the left score.  And the seed a chain visits first — the one with the
// highest score — is extended unless an earlier chain's region already contains it.  So, as for mate rescue (plan,
// batch, replay):
//
//   k_ext_plan   thread / read   for every chain: its reference window and its top seed -> one ExtPlan record
//   (sort)                       plans ordered by the query length of their left (then right) extension
//   k_ext_wave   thread / task   ALL left extensions of the bucket, 32 similar-sized tasks per warp advancing row by row
//                                (ksw_lanes.cuh: 21 instructions per DP cell instead of ~190 for a warp per task);
//                                a band retry (bwa/bwamem.c:748-757) runs in the same lane right away
//   k_ext_wave   thread / task   all right extensions, h0 = the score the left one reached
//   k_align1     warp / read     the reference's control flow as before; its ksw_extend2 calls are answered from the
//                                read's ExtPlan records when the arguments match exactly, and computed inline
//                                (warp_extend) otherwise — second seeds of a chain, rare
//
// The result is the reference's whatever the plan guessed: a planned extension the replay never asks for is wasted
// work and nothing else (test_ext_plan_is_transparent: identical regions and visited cells with the plan on and off).
#pragma once
#include "align_lanes.cuh"

struct ExtPlan {   // one chain's top seed and its two extensions
	int64_t rbeg, rmax0, rmax1;
	int32_t read, qbeg, len, l_query;
	int32_t sc0;                   // score after the left extension = h0 of the right one
	int8_t n_left, n_right;        // tries computed (0: not planned / not needed, 1, 2 with the band doubled)
	int16_t pad;
	ExtResult left[2], right[2];
	uint32_t cells_left[2], cells_right[2];
};

#ifdef __CUDACC__

// k_ext_plan: Pools_ is pipeline.cu's Pools.  chain_off[r] = first plan slot of read r (exclusive sum of n_chains).
// lkey/rkey = query lengths of the two extensions (0: none), the sort keys of the two waves.
template <class Pools_>
__global__ void __launch_bounds__(128)
k_ext_plan(DevIndex ix, int n_reads, const int64_t *off, const int32_t *occ_off, Pools_ p, const int32_t *chain_off, ExtPlan *plans,
           uint8_t *lkey, uint8_t *rkey)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int o = occ_off[r], l_query = (int)(off[r + 1] - off[r]);
	const Chain *chains = p.chains + o;
	const Seed *seeds_all = p.seeds + o;
	const int nc = p.n_chains[r];
	for (int ci = 0; ci < nc; ++ci) {
		const int slot = chain_off[r] + ci;
		ExtPlan &pl = plans[slot];
		const Chain &c = chains[ci];
		pl.read = r; pl.l_query = l_query; pl.n_left = pl.n_right = 0; pl.pad = 0;
		if (c.n == 0) { pl.qbeg = -1; pl.len = 0; lkey[slot] = rkey[slot] = 0; continue; }
		const Seed *seeds = seeds_all + c.seed_beg;
		const ChainWin cw = chain_window(ix, l_query, c, seeds);
		// the seed mem_chain2aln visits first: largest (score << 32 | index) (bwa/bwamem.c:687-693)
		int top = 0;
		for (int i = 1; i < c.n; ++i) if (seeds[i].score >= seeds[top].score) top = i;
		const Seed &s = seeds[top];
		pl.rbeg = s.rbeg; pl.rmax0 = cw.rmax0; pl.rmax1 = cw.rmax1; pl.qbeg = s.qbeg; pl.len = s.len;
		pl.sc0 = s.len * opt::a;   // left_none
		lkey[slot] = (uint8_t)s.qbeg;
		rkey[slot] = (uint8_t)(l_query - (s.qbeg + s.len));
	}
}

// One wave: task t of the warp = plans[order[32 * block + lane]]; lanes past n_tasks or without this side idle.
template <bool RIGHT>
__global__ void __launch_bounds__(32)
k_ext_wave(DevIndex ix, const uint8_t *seq, const int64_t *off, ExtPlan *plans, const int32_t *order, const uint8_t *keys_sorted, int n_tasks,
           unsigned long long *planned_cells, int qlo, int qhi)
{
	// One launch serves the tasks whose query length lies in (qlo, qhi]: shared memory is sized for qhi, so the short
	// extensions — most of them — run at several times the occupancy the longest read would allow.  The list is sorted
	// by length, so all but a boundary block of a launch either take all their lanes or leave at once.
	extern __shared__ uint32_t wave_smem[];
	const int lane = threadIdx.x, t = blockIdx.x * 32 + lane;
	const int key = t < n_tasks ? keys_sorted[t] : 0;
	bool valid = key > qlo && key <= qhi;
	if (!__any_sync(FULL_MASK, valid)) return;
	ExtPlan *pl = valid ? plans + order[t] : nullptr;
	Seed s{};
	ChainWin cw{};
	SeedExt e;
	e.a.score = -1; e.aw0 = e.aw1 = opt::w; e.sc0 = 0;
	const uint8_t *query = nullptr;
	int l_query = 0;
	if (valid) {
		s.rbeg = pl->rbeg; s.qbeg = pl->qbeg; s.len = pl->len;
		cw.rmax0 = pl->rmax0; cw.rmax1 = pl->rmax1;
		l_query = pl->l_query;
		query = seq + off[pl->read];
		if (RIGHT) { e.sc0 = pl->sc0; e.a.score = pl->sc0; }
	}
	unsigned long long visited = 0, total = 0;
	int tr = 0;
	bool req = valid;
	for (;;) {
		if (!__any_sync(FULL_MASK, req)) break;
		ExtTask x{};
		if (req) x = RIGHT ? right_task(l_query, s, cw, e, tr) : left_task(s, cw, tr);
		lanes::QueryFetch qf{query, x.q0, x.qstep};
		lanes::RefLaneFetch tf{&ix, x.t0, x.tstep};
		visited = 0;
		const ExtResult res = lanes::extend(wave_smem + lane, req, x.qlen, x.tlen, x.h0, x.w, x.end_bonus, opt::zdrop, qf, tf, visited);
		if (req) {
			total += visited;
			if (RIGHT) { pl->right[tr] = res; pl->cells_right[tr] = (uint32_t)visited; pl->n_right = (int8_t)(tr + 1); }
			else { pl->left[tr] = res; pl->cells_left[tr] = (uint32_t)visited; pl->n_left = (int8_t)(tr + 1); }
			const bool again = RIGHT ? right_try_done(e, res, tr) : left_try_done(e, res, tr);
			if (again) ++tr;
			else {
				if (!RIGHT) pl->sc0 = e.a.score;
				req = false;
			}
		}
	}
	for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(FULL_MASK, total, d);
	if (lane == 0 && total) atomicAdd(planned_cells, total);
}

// Looks an extension call up in a read's plans.  Arguments as WarpPolicy::extend receives them (ExtTask of align.cuh).
__device__ __forceinline__ bool ext_plan_lookup(const ExtPlan *plans, int n_plans, int q0, int qstep, int qlen, int64_t t0, int tlen, int w, int h0,
                                                ExtResult *res, uint32_t *cells)
{
	for (int k = 0; k < n_plans; ++k) {
		const ExtPlan &pl = plans[k];
		if (pl.qbeg < 0) continue;
		if (qstep < 0) {  // left_task: q0 = qbeg - 1, t0 = rbeg - 1, tlen = rbeg - rmax0, h0 = len * a
			if (q0 + 1 != pl.qbeg || qlen != pl.qbeg || t0 + 1 != pl.rbeg || (int64_t)tlen != pl.rbeg - pl.rmax0 || h0 != pl.len * opt::a) continue;
			const int tr = w == opt::w ? 0 : (w == opt::w << 1 ? 1 : 2);
			if (tr >= pl.n_left) continue;
			*res = pl.left[tr]; *cells = pl.cells_left[tr];
			return true;
		} else {          // right_task: q0 = qbeg + len, t0 = rbeg + len, tlen = rmax1 - t0, h0 = sc0
			if (q0 != pl.qbeg + pl.len || qlen != pl.l_query - q0 || t0 != pl.rbeg + pl.len || (int64_t)tlen != pl.rmax1 - t0 || h0 != pl.sc0) continue;
			const int tr = w == opt::w ? 0 : (w == opt::w << 1 ? 1 : 2);
			if (tr >= pl.n_right) continue;
			*res = pl.right[tr]; *cells = pl.cells_right[tr];
			return true;
		}
	}
	return false;
}

// the same lookup as a cache object for the thread-per-read walk (lanes::align1_warp)
struct PlanCache {
	static constexpr bool COOP_MISSES = true;    // misses are rare: the warp runs each one together (align_lanes.cuh)
	const ExtPlan *plans;
	const int32_t *chain_off;
	const ExtPlan *mine = nullptr;
	int n_mine = 0;
	__device__ __forceinline__ void set_read(int r, int n_chains) { mine = plans + chain_off[r]; n_mine = n_chains; }
	__device__ __forceinline__ bool find(const ExtTask &x, ExtResult *res, unsigned long long *cells) const
	{
		uint32_t c = 0;
		if (!ext_plan_lookup(mine, n_mine, x.q0, x.qstep, x.qlen, x.t0, x.tlen, x.w, x.h0, res, &c)) return false;
		*cells = c;
		return true;
	}
};

#endif
