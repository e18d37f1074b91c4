// The seeding kernel the pipeline runs: the algorithm of seed_hot.cuh (one-hot Occ blocks, k-mer start table, text
// comparison at a unique locus; the thread-scalar statement there is what tests/hostsim checks against the reference)
// arranged as a REQUEST LOOP for a warp of eight reads, four lanes per read:
//
//     for (;;) {  prepare: every quad walks its state machine — latency-free code only — until it needs memory,
//                          and posts up to four 16-byte loads per lane;
//                 issue:   ONE convergent point where all 32 lanes issue their loads;
//                 consume: per-state code on the loaded values, again latency-free.  }
//
// Why this shape: the first device form of seed_hot.cuh kept the k-mer table lookups, the SA lookup and the text
// compare inside divergent per-state code, so one quad's DRAM round trip stalled the seven other reads of its warp, and
// the per-state code paths were long enough to thrash the instruction cache (profiles/r2q_*: 1.6 G warp instructions at
// 12 live lanes, "no instruction" the second stall reason).  Here every DRAM access of every state goes through the
// one issue point — a warp has eight reads' loads in flight at once — and the per-state code is short.
//
// States of one read: READ (its 2-bit packed bases into shared memory) -> passes 1+2: TAB (k-mer table levels of a
// bwt_smem1a call), FWD (FM-index forward step, the four bases spread over the quad), SA + TEXT (one-occurrence
// sweep: locus, then 128 bases of text per request), BWD (one step of a backward round: four list entries, one per
// lane) -> pass 3: P3_TAB, P3_FWD, P3_SA, P3_TEXT -> next read.  Interval lists live in shared memory as
// {x0 | end << 48, x2}.  Results: (x0, 0, x2, info) exactly as seed_hot.cuh (bwa/bwt.c:289-379, bwa/bwamem.c:140-188).
#pragma once
#include "seed_hot.cuh"

#ifdef __CUDACC__
#define RQ_CAP 24                           // list entries kept in shared memory; longer lists continue in the quad's global slab
#define RQ_READ_WORDS 20                    // 16 words of bases (256) + a zero pad; 17..19 unused
#define RQ_QUAD_U32 (2 * RQ_CAP * 4 + RQ_READ_WORDS + 8)   // lists + bases + N mask
#define RQ_SMEM_BYTES(quads) ((quads) * RQ_QUAD_U32 * 4)
#define RQ_PACKED_BYTES 96                  // per read in the packed-read buffer: 64 bytes of bases, 32 of N mask

// reads as 2-bit words (base b at word b >> 4, bits 30 - 2 (b & 15); N stored as 0) + an N bit mask (base b at word
// b >> 5, bit 31 - (b & 31)): thread per (read, word)
static __global__ void k_pack_reads(const uint8_t *seq, const int64_t *off, int n_reads, uint32_t *packed)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int r = t / 24, w = t % 24;
	if (r >= n_reads) return;
	const uint8_t *s = seq + off[r];
	const int len = (int)(off[r + 1] - off[r]);
	uint32_t v = 0;
	if (w < 16) {
		for (int k = 0; k < 16; ++k) {
			const int b = w * 16 + k;
			const uint32_t x = b < len ? s[b] : 0u;
			v |= (x < 4 ? x : 0u) << (30 - 2 * k);
		}
	} else {
		for (int k = 0; k < 32; ++k) {
			const int b = (w - 16) * 32 + k;
			if (b < len && s[b] > 3) v |= 1u << (31 - k);
		}
	}
	packed[(size_t)r * 24 + w] = v;
}

struct RqBatch {
	SeedBatch b;
	const uint32_t *packed;
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint32_t rev2_32(uint32_t x)   // reverses the order of the sixteen 2-bit groups
{
	const uint32_t y = __brev(x);
	return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}

struct RqQuad {   // everything one lane keeps about its quad's read; fields are identical on the four lanes unless noted
	// environment
	const DevIndex &ix;
	const RqBatch &rb;
	uint4 *lst;          // [2][RQ_CAP] list entries {x0 | end << 48, x2}
	uint32_t *rw;        // [RQ_READ_WORDS] packed bases, then [8] N mask
	uint4 *slab;         // [2][slab_len] entries past RQ_CAP
	int slab_len, q;
	unsigned qmask, qshift;
	unsigned sectors;
	// job
	int st, rid, job_p3, len, has_n, n12, n3, ovf, cap12;
	Intv *out12, *out3;
	// passes
	int pass, x, old_n, k2;
	Pass2Queue q2;
	// the bwt_smem1a call in progress
	int sx, i, ret, last_start, n_prev, n_curr, j, cur, rev, c, in_p2, m;
	uint64_t min_intv, last_size;
	uint64_t f_x0, f_x1, f_x2;   // the forward sweep's interval [sx, i) (passes 1+2) or [x, i) (pass 3)
	int64_t text_p;              // locus of the sweep's first base once the interval has one row
	// BWD: this lane's entry (differs per lane)
	uint64_t e_x0, e_x2;
	int e_end, e_valid;

	// states that end in a memory request (the three Occ kinds first), then the latency-free ones
	enum { FWD = 0, P3_FWD, BWD, READ, TAB, P3_TAB, SA, P3_SA, TEXT, P3_TEXT,
	       JOB, NEXT, P3_NEXT, AFTER_FWD, CLOSE, START_BWD, ROUND, P3_ADV, DRAINED };

	// ---- the read in shared memory
	__device__ __forceinline__ int base(int k) const { return (rw[k >> 4] >> (30 - 2 * (k & 15))) & 3; }
	__device__ __forceinline__ bool is_n(int k) const { return has_n && ((rw[RQ_READ_WORDS + (k >> 5)] >> (31 - (k & 31))) & 1); }
	__device__ __forceinline__ uint32_t word16(int k) const { return __funnelshift_l(rw[(k >> 4) + 1], rw[k >> 4], 2 * (k & 15)); }
	__device__ __forceinline__ int valid_end(int k) const   // first index >= k that is N or len
	{
		if (!has_n) return len;
		while (k < len && !is_n(k)) ++k;
		return k;
	}
	// ---- lists
	__device__ __forceinline__ void lst_put(int which, int idx, uint64_t x0, uint64_t x2, int end)
	{
		const uint64_t A = x0 | (uint64_t)end << 48;
		const uint4 u = make_uint4((uint32_t)A, (uint32_t)(A >> 32), (uint32_t)x2, (uint32_t)(x2 >> 32));
		if (idx < RQ_CAP) lst[which * RQ_CAP + idx] = u; else slab[(size_t)which * slab_len + idx] = u;
	}
	__device__ __forceinline__ void lst_get(int which, int idx, uint64_t &x0, uint64_t &x2, int &end) const
	{
		const uint4 u = idx < RQ_CAP ? lst[which * RQ_CAP + idx] : slab[(size_t)which * slab_len + idx];
		x0 = ((uint64_t)(u.y & 0xffffu) << 32) | u.x;
		end = (int)(u.y >> 16);
		x2 = (uint64_t)u.w << 32 | u.z;
	}
	__device__ __forceinline__ unsigned qballot(bool p) const { return (__ballot_sync(qmask, p) >> qshift) & 0xfu; }
	__device__ __forceinline__ uint64_t qbcast(uint64_t v, int t) const { return __shfl_sync(qmask, v, (int)qshift + t); }
	__device__ __forceinline__ uint64_t qsum(uint64_t v) const
	{
		v += __shfl_xor_sync(qmask, v, 1);
		v += __shfl_xor_sync(qmask, v, 2);
		return v;
	}
	// ---- output
	__device__ __forceinline__ void emit12(uint64_t x0, uint64_t x2, int start, int end)
	{
		if (end - start < opt::min_seed_len) return;
		if (!in_p2) { Intv pq; pq.x2 = x2; pq.info = (uint64_t)(uint32_t)end; q2.consider(pq, start, n12); }
		if (n12 < cap12) {
			const uint64_t comp = q == 0 ? x0 : (q == 1 ? 0ull : (q == 2 ? x2 : ((uint64_t)start << 32 | (uint64_t)end)));
			((uint64_t *)(out12 + n12))[q] = comp;
			++n12;
		} else ovf = 1;
	}
	__device__ __forceinline__ void emit3(uint64_t x0, uint64_t x2, int start, int end)
	{
		if (n3 < EMAB_P3_CAP) {
			const uint64_t comp = q == 0 ? x0 : (q == 1 ? 0ull : (q == 2 ? x2 : ((uint64_t)start << 32 | (uint64_t)end)));
			((uint64_t *)(out3 + n3))[q] = comp;
			++n3;
		} else ovf = 1;
	}
};

// Occ(base, .) at the two ends of [xa, xa + xn): addresses (prepare) and counts (consume); HotFm::occ_pair split in two
struct OccReq { uint64_t _k, _l; bool kv, lv, other; };
__device__ __forceinline__ OccReq occ_addr(const DevIndex &ix, uint64_t xa, uint64_t xn, int base, const uint4 *&a0, const uint4 *&a1)
{
	const uint64_t NEG1 = ~0ull;
	const uint64_t k = xa - 1, l = xa - 1 + xn;
	OccReq r;
	r.kv = k != NEG1; r.lv = l != NEG1;
	r._k = r.kv ? k - (k >= ix.primary) : 0; r._l = r.lv ? l - (l >= ix.primary) : 0;
	const uint64_t bk = r._k >> 6, bl = r._l >> 6;
	r.other = bk != bl;
	a0 = ix.hot + (bk << 2) + base;
	a1 = ix.hot + (bl << 2) + base;
	return r;
}
__device__ __forceinline__ void occ_count(const OccReq &r, const uint4 &ek, const uint4 &el_, uint64_t &tk, uint64_t &ns)
{
	const uint4 el = r.other ? el_ : ek;
	const uint64_t mk = (2ull << (r._k & 63)) - 1, ml = (2ull << (r._l & 63)) - 1;
	const uint64_t ck = ((uint64_t)ek.y << 32 | ek.x) + (uint64_t)__popcll(((uint64_t)ek.w << 32 | ek.z) & mk);
	const uint64_t cl = ((uint64_t)el.y << 32 | el.x) + (uint64_t)__popcll(((uint64_t)el.w << 32 | el.z) & ml);
	tk = r.kv ? ck : 0;
	ns = (r.lv ? cl : 0) - tk;
}

// 32 bases of the forward-reverse text starting at tp from the two 16-byte pac chunks loaded for them (fast == true), as
// a 64-bit word, base j at bits 62 - 2j.  Chunk addresses: text_chunks().
__device__ __forceinline__ bool text_chunks(const DevIndex &ix, int64_t tp, const uint4 *&a0, const uint4 *&a1, int64_t &g)
{
	const int64_t L = ix.l_pac;
	if (tp >= 0 && tp + 32 <= L) g = tp;
	else if (tp >= L && tp + 32 <= 2 * L) g = 2 * L - 32 - tp;   // forward-strand start of the 32 complementary bases
	else return false;
	a0 = (const uint4 *)ix.pac + (g >> 6);
	a1 = a0 + 1;
	return true;
}
__device__ __forceinline__ uint64_t text_word32(const DevIndex &ix, int64_t tp, int64_t g, const uint4 &c0, const uint4 &c1)
{
	const int wi = (int)(g & 63) >> 4, sh = 2 * (int)(g & 15);
	uint32_t w0 = c0.x, w1 = c0.y, w2 = c0.z, w3 = c0.w, w4 = c1.x, w5 = c1.y, w6 = c1.z;
	uint32_t a = w0, b = w1, d = w2;
	if (wi == 1) { a = w1; b = w2; d = w3; }
	if (wi == 2) { a = w2; b = w3; d = w4; }
	if (wi == 3) { a = w3; b = w4; d = w5; }
	(void)w6;
	a = bswap32(a); b = bswap32(b); d = bswap32(d);
	uint32_t hi = __funnelshift_l(b, a, sh), lo = __funnelshift_l(d, b, sh);
	if (tp >= ix.l_pac) {   // reverse half: reverse the 32 groups and complement
		const uint32_t h2 = ~rev2_32(lo), l2 = ~rev2_32(hi);
		hi = h2; lo = l2;
	}
	return (uint64_t)hi << 32 | lo;
}
// the same without the loaded chunks: across the strand junction and at the ends of the text (rare)
static __device__ __noinline__ uint64_t text_word32_slow(const DevIndex &ix, int64_t tp)
{
	uint64_t r = 0;
	for (int k = 0; k < 32; ++k) {
		const int64_t p = tp + k;
		const uint64_t b = p >= 0 && p < 2 * ix.l_pac ? (uint64_t)ref_base(ix, p) : 0ull;
		r |= b << (62 - 2 * k);
	}
	return r;
}

// PROF: per-state iteration counts (quads x iterations) into rb.b.touches[1 + state] — a measurement build of the same loop
template <bool PROF>
__device__ __forceinline__ void seed_rq_warp(const DevIndex &ix, const RqBatch &rb, uint32_t *smem)
{
	const int lane = threadIdx.x & 31;
	const int quad = threadIdx.x >> 2, qi = threadIdx.x & 3;
	const size_t gquad = (size_t)blockIdx.x * (blockDim.x >> 2) + quad;
	const SeedBatch &b = rb.b;
	uint32_t *qs = smem + quad * RQ_QUAD_U32;
	RqQuad s{ix, rb, (uint4 *)qs, qs + 2 * RQ_CAP * 4, (uint4 *)(b.scratch + gquad * 2 * b.scratch_len), 2 * b.scratch_len, qi,
	         0xfu << (lane & ~3), (unsigned)(lane & ~3), 0};
	s.st = RqQuad::JOB; s.rid = -1; s.job_p3 = 0; s.len = 0; s.has_n = 0; s.n12 = s.n3 = s.ovf = 0; s.cap12 = 0; s.out12 = s.out3 = nullptr;
	s.pass = 1; s.x = 0; s.old_n = 0; s.k2 = 0; s.q2.clear();
	s.sx = s.i = s.ret = 0; s.last_start = 0x7fffffff; s.n_prev = s.n_curr = s.j = 0; s.cur = 1; s.rev = 0; s.c = 0; s.in_p2 = 0; s.m = 0;
	s.min_intv = 1; s.last_size = 0; s.f_x0 = s.f_x1 = s.f_x2 = 0; s.text_p = 0; s.e_x0 = s.e_x2 = 0; s.e_end = 0; s.e_valid = 0;
	const int K = ix.kmer_k;
	for (;;) {
		// what the quad's lanes stored in the last iteration's consume (list entries, read words) is ordered before what any
		// of them reads from here on; all 32 lanes pass here every iteration
		__syncwarp();
		// ------------------------------------------------------------------ latency-free transitions, each written once
		while (s.st >= RqQuad::JOB && s.st != RqQuad::DRAINED) {
			switch (s.st) {
			case RqQuad::JOB: {
				// A read is TWO jobs: passes 1+2 (~165 iterations) and pass 3 (~40).  The queue hands out all the long jobs
				// first, then the short ones: the kernel's tail — quads running their last job while the others are dry —
				// is as long as a short job instead of a whole read (~0.9 ms).
				if (s.rid >= 0 && qi == 0) { if (s.job_p3) b.n3[s.rid] = s.n3; else b.n12[s.rid] = s.n12; if (s.ovf) *b.err = 3; }
				unsigned long long r = 0;
				if (qi == 0) r = atomicAdd(&b.queue[0], 1ull);
				r = __shfl_sync(s.qmask, r, (int)s.qshift);
				if (r >= 2ull * (unsigned long long)b.n_reads) { s.rid = -1; s.st = RqQuad::DRAINED; break; }
				s.job_p3 = r >= (unsigned long long)b.n_reads;
				if (s.job_p3) r -= (unsigned long long)b.n_reads;
				s.rid = (int)r;
				s.len = (int)(b.off[r + 1] - b.off[r]);
				s.out12 = b.intv + (size_t)r * b.max_intv; s.cap12 = b.max_intv;
				s.out3 = b.p3 + (size_t)r * EMAB_P3_CAP;
				s.n12 = s.n3 = s.ovf = 0; s.pass = 1; s.x = 0; s.q2.clear();
				s.st = RqQuad::READ;
				break;
			}
			case RqQuad::NEXT: {
				bool start_call = false;
				if (s.pass == 1) {
					while (s.x < s.len && s.is_n(s.x)) ++s.x;
					if (s.x >= s.len) { s.pass = 2; s.old_n = s.n12; s.k2 = 0; }
					else { s.sx = s.x; s.min_intv = 1; s.in_p2 = 0; start_call = true; }
				} else {   // pass 2: bwa/bwamem.c:157-168
					if (s.q2.pop(&s.sx, &s.min_intv)) { s.in_p2 = 1; start_call = true; }
					else if (s.q2.spill_from >= 0) {   // more than five candidates: the rest are looked up in the output list
						if (s.k2 < s.q2.spill_from) s.k2 = s.q2.spill_from;
						__syncwarp(s.qmask);
						while (s.k2 < s.old_n) {
							const Intv p = s.out12[s.k2];
							const int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
							if (end - start >= opt::split_len && p.x2 <= (uint64_t)opt::split_width) break;
							++s.k2;
						}
						if (s.k2 >= s.old_n) s.st = RqQuad::JOB;
						else {
							const Intv p = s.out12[s.k2++];
							s.sx = ((int)(p.info >> 32) + (int)(uint32_t)p.info) >> 1;
							s.min_intv = p.x2 + 1; s.in_p2 = 1; start_call = true;
						}
					} else s.st = RqQuad::JOB;
				}
				if (start_call) {   // bwt_smem1a(sx, min_intv): bwa/bwt.c:289-302
					s.n_curr = 0; s.last_start = 0x7fffffff;
					if (K > 0) s.st = RqQuad::TAB;
					else {
						Intv ik;
						bwt_set_intv(ix, s.base(s.sx), ik);
						s.f_x0 = ik.x0; s.f_x1 = ik.x1; s.f_x2 = ik.x2;
						s.i = s.sx + 1;
						s.st = RqQuad::AFTER_FWD;
					}
				}
				break;
			}
			case RqQuad::P3_NEXT:
				while (s.x < s.len && s.is_n(s.x)) ++s.x;
				if (s.x >= s.len) s.st = RqQuad::JOB;
				else if (K > 0) s.st = RqQuad::P3_TAB;
				else {
					Intv ik;
					bwt_set_intv(ix, s.base(s.x), ik);
					s.f_x0 = ik.x0; s.f_x1 = ik.x1; s.f_x2 = ik.x2;
					s.i = s.x + 1;
					s.st = RqQuad::P3_ADV;
				}
				break;
			case RqQuad::AFTER_FWD:   // the forward sweep's interval now covers [sx, i)
				if (s.i >= s.len || s.is_n(s.i)) s.st = RqQuad::CLOSE;
				else s.st = s.f_x2 == 1 && s.min_intv <= 1 ? RqQuad::SA : RqQuad::FWD;
				break;
			case RqQuad::CLOSE:   // end of the forward sweep: the last interval is recorded (bwa/bwt.c:316-321)
				s.lst_put(s.cur, s.n_curr, s.f_x0, s.f_x2, s.i);   // all four lanes, the same value
				++s.n_curr;
				s.ret = s.i;
				s.st = RqQuad::START_BWD;
				break;
			case RqQuad::START_BWD:   // (the list's entries were stored by all four lanes, or before the loop's last __syncwarp)
				s.cur ^= 1;
				s.n_prev = s.n_curr; s.n_curr = 0;
				s.i = s.sx - 1; s.rev = 1;
				s.st = RqQuad::ROUND;
				break;
			case RqQuad::ROUND: {   // a backward round at query position i over n_prev entries (bwa/bwt.c:324-326)
				const int cc = s.i < 0 || s.is_n(s.i) ? -1 : s.base(s.i);
				if (cc < 0) {   // nothing extends: only the longest interval may be emitted, and the call is over (bwa/bwt.c:331-338,346)
					if (s.i + 1 < s.last_start) {
						uint64_t x0, x2; int end;
						s.lst_get(s.cur ^ 1, s.rev ? s.n_prev - 1 : 0, x0, x2, end);
						s.emit12(x0, x2, s.i + 1, end);
					}
					if (!s.in_p2) s.x = s.ret;
					s.st = RqQuad::NEXT;
				} else { s.c = cc; s.j = 0; s.last_size = 0; s.st = RqQuad::BWD; }
				break;
			}
			case RqQuad::P3_ADV:   // the reference's loop is at index i with interval [x, i)  (bwa/bwt.c:363-378)
				if (s.i >= s.len) { s.x = s.len; s.st = RqQuad::P3_NEXT; }
				else if (s.is_n(s.i)) { s.x = s.i + 1; s.st = RqQuad::P3_NEXT; }
				else if (s.f_x2 == 0) {   // an empty interval stays empty: the loop runs on until an ambiguous base, 19 bases, or the end
					while (s.i < s.len && !s.is_n(s.i) && s.i - s.x < opt::min_seed_len) ++s.i;
					s.x = s.i < s.len ? s.i + 1 : s.len;
					s.st = RqQuad::P3_NEXT;
				} else s.st = s.f_x2 == 1 ? RqQuad::P3_SA : RqQuad::P3_FWD;
				break;
			default: break;
			}
		}
		// ------------------------------------------------------------------ the request of this iteration
		const bool req = s.st != RqQuad::DRAINED;
		const bool is_occ = s.st <= RqQuad::BWD;
		if (PROF && req && qi == 0) atomicAdd(rb.b.touches + 1 + s.st, 1ull);
		const uint4 *a0 = nullptr, *a1 = nullptr, *a2 = nullptr, *a3 = nullptr;
		bool p0 = false, p1 = false, p2 = false, p3 = false, uni = false;   // uni: the quad's lanes ask for the same bytes
		OccReq oq; oq._k = oq._l = 0; oq.kv = oq.lv = oq.other = false;
		int64_t tg = 0; bool tfast = false;
		if (is_occ) {   // FWD, P3_FWD: the four bases of one forward step over the quad; BWD: four list entries, one per lane
			uint64_t xa = s.f_x1, xn = s.f_x2;
			int ob = qi;
			bool valid = true;
			if (s.st == RqQuad::BWD) {
				int jj = s.j + qi;
				valid = jj < s.n_prev;
				jj = valid ? jj : s.n_prev - 1;
				s.lst_get(s.cur ^ 1, s.rev ? s.n_prev - 1 - jj : jj, s.e_x0, s.e_x2, s.e_end);
				s.e_valid = valid;
				xa = s.e_x0; xn = s.e_x2; ob = s.c;
			} else s.c = 3 - s.base(s.i);
			oq = occ_addr(ix, xa, xn, ob, a0, a1);
			p0 = valid; p1 = valid && oq.other;
		} else if (req) {
			switch (s.st) {
			case RqQuad::READ:
				a0 = (const uint4 *)(rb.packed + (size_t)s.rid * 24) + qi; p0 = true;
				a1 = (const uint4 *)(rb.packed + (size_t)s.rid * 24 + 16) + (qi & 1); p1 = qi < 2;
				break;
			case RqQuad::TAB: case RqQuad::P3_TAB: {
				const int from = s.st == RqQuad::TAB ? s.sx : s.x;
				int m = s.valid_end(from) - from;
				m = m < K ? m : K;
				s.m = m;
				const uint32_t code = s.word16(from) >> (32 - 2 * m);
				if (s.st == RqQuad::TAB) {   // lane q: levels q + 1, q + 5, q + 9, q + 13
					const int t0 = qi + 1, t1 = qi + 5, t2 = qi + 9, t3 = qi + 13;
					p0 = t0 <= m; p1 = t1 <= m; p2 = t2 <= m; p3 = t3 <= m;
					a0 = ix.kmer + (p0 ? kmer_level_off(t0) + (code >> (2 * (m - t0))) : 0);
					a1 = ix.kmer + (p1 ? kmer_level_off(t1) + (code >> (2 * (m - t1))) : 0);
					a2 = ix.kmer + (p2 ? kmer_level_off(t2) + (code >> (2 * (m - t2))) : 0);
					a3 = ix.kmer + (p3 ? kmer_level_off(t3) + (code >> (2 * (m - t3))) : 0);
				} else { a0 = ix.kmer + kmer_level_off(m) + code; p0 = true; uni = true; }
				break;
			}
			case RqQuad::SA: case RqQuad::P3_SA:
				if (ix.sa32) a0 = (const uint4 *)(ix.sa32 + (s.f_x0 & ~3ull)); else a0 = (const uint4 *)(ix.sa64 + (s.f_x0 & ~1ull));
				p0 = true; uni = true;
				break;
			default: {   // TEXT, P3_TEXT
				const int from = s.st == RqQuad::TEXT ? s.sx : s.x;
				const int64_t tp = s.text_p + (s.i - from) + 32 * qi;
				tfast = text_chunks(ix, tp, a0, a1, tg);
				p0 = p1 = tfast && s.i + 32 * qi < s.len && (s.st == RqQuad::TEXT || qi == 0);
				break;
			}
			}
		}
		if (!__any_sync(0xffffffffu, req)) break;
		__syncwarp();   // ... and what the transitions above stored (all four lanes the same value) before the consume below reads it
		// ------------------------------------------------------------------ issue: the one place where DRAM is read
		uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
		ldg128_if(p0, a0, r0); ldg128_if(p1, a1, r1); ldg128_if(p2, a2, r2); ldg128_if(p3, a3, r3);
		if (!uni || qi == 0) s.sectors += (unsigned)p0 + (unsigned)p1 + (unsigned)p2 + (unsigned)p3;
		// ------------------------------------------------------------------ consume
		// The Occ states' cross-lane traffic, for ALL 32 lanes at this convergent point and with the full mask: a shuffle or
		// vote over a quad's own mask inside per-state code costs a MATCH/WARPSYNC sequence each (a tenth of the kernel's
		// instructions and 15 % of its stall samples before this).  Lanes in other states take part with values nobody reads.
		uint64_t tk, ns;
		occ_count(oq, r0, r1, tk, ns);
		const bool bwd = is_occ && s.st == RqQuad::BWD;
		const bool alive = bwd && s.e_valid && !(ns < s.min_intv);
		const unsigned aliveb = (__ballot_sync(0xffffffffu, alive) >> s.qshift) & 0xfu;
		uint64_t prev_x2 = __shfl_up_sync(0xffffffffu, ns, 1, 4);
		bool prev_alive = qi ? (aliveb >> (qi - 1)) & 1 : s.n_curr > 0;
		if (qi == 0) prev_x2 = s.last_size;
		// survivors of a backward round: recorded when the first of the round or different in size from the survivor before
		// (sizes never decrease along the list: every entry's string is a prefix of the one before it)
		const bool push = alive && (!prev_alive || ns != prev_x2);
		const unsigned pushb = (__ballot_sync(0xffffffffu, push) >> s.qshift) & 0xfu;
		const uint64_t last_x2 = __shfl_sync(0xffffffffu, ns, (int)s.qshift + 3);
		const int cq = (int)s.qshift + (s.c & 3);
		const uint64_t tk_c = __shfl_sync(0xffffffffu, tk, cq), ns_c = __shfl_sync(0xffffffffu, ns, cq);
		uint64_t above = qi > s.c ? ns : 0;
		above += __shfl_xor_sync(0xffffffffu, above, 1);
		above += __shfl_xor_sync(0xffffffffu, above, 2);
		if (!req) continue;
		if (is_occ) {
			if (bwd) {   // bwa/bwt.c:328-345 for four entries of the round, in list order
				const uint64_t ok_x0 = ix.L2[s.c] + 1 + tk, ok_x2 = ns;
				if (s.j == 0 && !(aliveb & 1)) {
					// the round's first entry did not survive; only it can be emitted: a later one finds either a survivor
					// before it or last_start already set (bwa/bwt.c:331-338)
					if (s.i + 1 < s.last_start) {
						uint64_t x0, x2; int end;
						s.lst_get(s.cur ^ 1, s.rev ? s.n_prev - 1 : 0, x0, x2, end);
						s.emit12(x0, x2, s.i + 1, end);
						s.last_start = s.i + 1;
					}
				}
				if (push) s.lst_put(s.cur, s.n_curr + __popc(pushb & ((1u << qi) - 1)), ok_x0, ok_x2, s.e_end);
				s.n_curr += __popc(pushb);
				s.j += 4;
				if (s.j >= s.n_prev) {   // end of the round (bwa/bwt.c:346-348)
					if (s.n_curr == 0) { if (!s.in_p2) s.x = s.ret; s.st = RqQuad::NEXT; }
					else {
						s.cur ^= 1;
						s.n_prev = s.n_curr; s.n_curr = 0;
						--s.i; s.rev = 0;
						s.st = RqQuad::ROUND;
					}
				} else s.last_size = last_x2;
			} else {   // bwt_extend forward (bwa/bwt.c:262-275), lane q counted base q
				const uint64_t ok_x1 = ix.L2[s.c] + 1 + tk_c, ok_x2 = ns_c;
				const uint64_t ok_x0 = s.f_x0 + (s.f_x1 <= ix.primary && s.f_x1 + s.f_x2 - 1 >= ix.primary) + above;
				if (s.st == RqQuad::FWD) {   // bwa/bwt.c:307-315
					bool stop = false;
					if (ok_x2 != s.f_x2) {
						s.lst_put(s.cur, s.n_curr, s.f_x0, s.f_x2, s.i);   // all four lanes, the same value: each reads back at least its own
						++s.n_curr; s.ret = s.i;
						stop = ok_x2 < s.min_intv;
					}
					if (stop) s.st = RqQuad::START_BWD;
					else { s.f_x0 = ok_x0; s.f_x1 = ok_x1; s.f_x2 = ok_x2; ++s.i; s.st = RqQuad::AFTER_FWD; }
				} else {   // bwa/bwt.c:366-375
					if (ok_x2 < (uint64_t)opt::max_mem_intv && s.i - s.x >= opt::min_seed_len) {
						if (ok_x2 > 0) s.emit3(ok_x0, ok_x2, s.x, s.i + 1);
						s.x = s.i + 1; s.st = RqQuad::P3_NEXT;
					} else { s.f_x0 = ok_x0; s.f_x1 = ok_x1; s.f_x2 = ok_x2; ++s.i; s.st = RqQuad::P3_ADV; }
				}
			}
			continue;
		}
		switch (s.st) {
		case RqQuad::READ: {
			s.rw[4 * qi] = r0.x; s.rw[4 * qi + 1] = r0.y; s.rw[4 * qi + 2] = r0.z; s.rw[4 * qi + 3] = r0.w;
			if (qi == 0) { s.rw[16] = 0; s.rw[17] = 0; }
			if (qi < 2) { uint32_t *nm = s.rw + RQ_READ_WORDS + 4 * qi; nm[0] = r1.x; nm[1] = r1.y; nm[2] = r1.z; nm[3] = r1.w; }
			s.has_n = s.qballot(qi < 2 && (r1.x | r1.y | r1.z | r1.w)) != 0;
			s.st = s.job_p3 ? RqQuad::P3_NEXT : RqQuad::NEXT;   // the words are read from the next iteration on
			break;
		}
		case RqQuad::TAB: {
			// levels 1..m of the forward sweep (bwa/bwt.c:303-315): every level goes to curr[t - 1]; the common case — all
			// sizes differ — leaves the list as it must be, otherwise lane 0 compacts it
			const int m = s.m;
			const uint4 rr[4] = {r0, r1, r2, r3};
			uint64_t my_x1 = 0;
#pragma unroll
			for (int sl = 0; sl < 4; ++sl) {
				const int t = 4 * sl + qi + 1;
				if (t <= m) {
					const Intv v = intv_unpack(rr[sl]);
					s.lst_put(s.cur, t - 1, v.x0, v.x2, s.sx + t);
					if (t == m) my_x1 = v.x1;
				}
			}
			__syncwarp(s.qmask);
			// below: bit t - 1 set when level t (t >= 2) differs from level t - 1 and is smaller than min_intv; same: level t equals level t + 1 in size
			unsigned below = 0, same = 0;
#pragma unroll
			for (int sl = 0; sl < 4; ++sl) {
				const int t = 4 * sl + qi + 1;
				bool bl = false, sm = false;
				if (t <= m) {
					uint64_t x0, x2, y0, y2; int e;
					s.lst_get(s.cur, t - 1, x0, x2, e);
					if (t >= 2) { s.lst_get(s.cur, t - 2, y0, y2, e); bl = x2 < s.min_intv && x2 != y2; }   // the sweep stops only at a change of size
					if (t < m) { s.lst_get(s.cur, t, y0, y2, e); sm = x2 == y2; }
				}
				below |= s.qballot(bl) << (4 * sl);
				same |= s.qballot(sm) << (4 * sl);
			}
			const bool stopped = below != 0;
			const int m_eff = stopped ? __ffs(below) - 1 : m;           // levels consumed by the sweep
			const int n_push = stopped ? m_eff : m_eff - 1;             // candidates for the list: levels 1 .. n_push
			const unsigned cand = (1u << n_push) - 1;
			if (same & cand & (stopped ? ~(1u << (m_eff - 1)) : ~0u)) {     // some level repeats the next one's size: it is not recorded
				if (qi == 0) {
					int k = 0;
					for (int t = 1; t <= n_push; ++t) {
						const bool rec = (stopped && t == m_eff) || !((same >> (t - 1)) & 1);
						if (rec) { if (k != t - 1) s.lst[s.cur * RQ_CAP + k] = s.lst[s.cur * RQ_CAP + t - 1]; ++k; }
					}
					s.n_curr = k;
				}
				s.n_curr = __shfl_sync(s.qmask, s.n_curr, (int)s.qshift);
				__syncwarp(s.qmask);
			} else s.n_curr = n_push;
			if (stopped) {
				s.ret = s.sx + m_eff;
				s.st = RqQuad::START_BWD;
			} else {
				uint64_t x0, x2; int e;
				s.lst_get(s.cur, m - 1, x0, x2, e);
				s.f_x0 = x0; s.f_x2 = x2;
				s.f_x1 = s.qbcast(my_x1, (m - 1) & 3);
				s.i = s.sx + m;
				s.st = RqQuad::AFTER_FWD;
			}
			break;
		}
		case RqQuad::P3_TAB: {
			const Intv v = intv_unpack(r0);
			s.f_x0 = v.x0; s.f_x1 = v.x1; s.f_x2 = v.x2;
			s.i = s.x + s.m;
			s.st = RqQuad::P3_ADV;
			break;
		}
		case RqQuad::SA: case RqQuad::P3_SA: {
			uint64_t p;
			if (ix.sa32) { const unsigned k = (unsigned)(s.f_x0 & 3); p = k == 0 ? r0.x : (k == 1 ? r0.y : (k == 2 ? r0.z : r0.w)); }
			else p = (s.f_x0 & 1) ? ((uint64_t)r0.w << 32 | r0.z) : ((uint64_t)r0.y << 32 | r0.x);
			s.text_p = (int64_t)p;
			s.st = s.st == RqQuad::SA ? RqQuad::TEXT : RqQuad::P3_TEXT;
			break;
		}
		default: {   // TEXT, P3_TEXT
			// lane q compares read bases [i + 32q, i + 32q + 32) with the text at the locus; the sweep goes on as far as they agree
			const bool p12 = s.st == RqQuad::TEXT;
			const int from = p12 ? s.sx : s.x;
			int lim = s.valid_end(s.i);                                   // first index the sweep cannot consume: N or len
			if (!p12 && lim > s.x + opt::min_seed_len + 1) lim = s.x + opt::min_seed_len + 1;   // pass 3 decides at index x + 19
			const int k0 = s.i + 32 * qi;
			const int64_t tp = s.text_p + (k0 - from);
			int n = lim - k0;
			n = n < 0 ? 0 : (n > 32 ? 32 : n);
			const int64_t room = (int64_t)ix.seq_len - tp;             // nothing extends past the end of the text
			if (room < n) n = room < 0 ? 0 : (int)room;
			int mt = 0;
			if (n > 0 && (p12 || qi == 0)) {
				const uint64_t tw = tfast ? text_word32(ix, tp, tg, r0, r1) : text_word32_slow(ix, tp);
				const uint64_t rwd = (uint64_t)s.word16(k0) << 32 | s.word16(k0 + 16);
				const uint64_t xr = tw ^ rwd;
				mt = xr ? __clzll((long long)xr) >> 1 : 32;
				if (mt > n) mt = n;
			}
			const unsigned full = s.qballot(mt == 32);
			const int first = __ffs(~full & 0xf) - 1;   // first lane that did not match all 32; -1: all four did
			const int run = first < 0 ? 128 : 32 * first + (int)s.qbcast((uint64_t)mt, first);
			s.i += run;
			if (p12) {
				if (run == 128 && s.i < lim) break;   // more of the read to compare: stay in TEXT
				s.st = RqQuad::CLOSE;
			} else {
				if (s.i >= s.x + opt::min_seed_len + 1) {   // matched through index x + 19: the interval is emitted there (bwa/bwt.c:366-375)
					s.emit3(s.f_x0, 1, s.x, s.x + opt::min_seed_len + 1);
					s.x = s.x + opt::min_seed_len + 1; s.st = RqQuad::P3_NEXT;
				} else if (s.i < lim) {   // a mismatch at index i
					if (s.i - s.x >= opt::min_seed_len) { s.x = s.i + 1; s.st = RqQuad::P3_NEXT; }
					else { s.f_x2 = 0; ++s.i; s.st = RqQuad::P3_ADV; }
				} else s.st = RqQuad::P3_ADV;   // the end of the read or an ambiguous base
			}
			break;
		}
		}
	}
	unsigned sectors = s.sectors;
	for (int d = 16; d; d >>= 1) sectors += __shfl_xor_sync(0xffffffffu, sectors, d);
	if (lane == 0 && sectors) atomicAdd(b.touches, (unsigned long long)sectors);
}
#endif
