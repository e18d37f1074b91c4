// emab_index_build: the FM index `ema align` loads (<ref>.bwt/.sa/.pac/.ann/.amb), built on the GPU.
//
// Replaces `bwa index` (bwa/bwtindex.c:255-323 bwa_idx_build): bns_fasta2bntseq (host, fasta2pac.cpp), then
// is_bwt / bwt_bwtgen2 (the BWT of forward + reverse complement, bwa/is.c:205-222, bwa/bwtindex.c:63-117),
// bwt_bwtupdate_core (Occ interleaving, :150-172) and bwt_cal_sa (:62-84 of bwa/bwt.c), with files written
// exactly as bwt_dump_bwt / bwt_dump_sa write them (bwa/bwt.c:385-407).  The files are byte-identical to the
// reference's (tests/test_gpu_index.py compares them with `bwa index` output), so either program loads either index.
//
// The reference builds the BWT with induced sorting (SA-IS) below 50 Mbp and incrementally (bwtsw) above — both
// sequential, about one CPU-hour for an hg38-sized genome.  Here the suffix array of the 2 x l_pac text is sorted
// directly, which is what the GPU is good at:
//
//   text      the 2-bit text T (forward, then reverse complement) as big-endian 64-bit words, so that the 29 bases at
//             any position are two loads, a funnel shift and a mask
//   key       key(p) = 29 bases at p (58 bits) << 6 | number of real bases (< 29 only at the end of T): comparing keys
//             orders suffixes by their first 29 bases with the end-of-text sentinel below every base — the order
//             is_sa uses (bwa/is.c:192-200: SA[0] = n, then the suffixes of T)
//   chunks    the key space is cut on its top bits into chunks of at most ~400 M suffixes (a histogram pass sizes them),
//             so one chunk's (key, position) pairs fit a double-buffered cub::DeviceRadixSort however long T is, and
//             chunk c's suffixes are final before chunk c+1 is touched
//   ties      suffixes that share their key (repeats longer than 29 bp) go to a worklist as (first rank of the tie
//             group, position); each round re-keys them 29 bases further on, sorts by (group, key), writes the
//             now-unique ones to their rank = group + offset inside the group, and keeps the rest with their sub-group's
//             first rank.  A round costs time proportional to the suffixes still tied.
//   emit      BWT symbol T[SA-1] per rank, every 32nd SA value, then the 2-bit packing around the primary row, the
//             per-128-symbol cumulative counts (one scan per symbol) and the interleaved layout of bwt_bwtupdate_core.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>
#include <cub/cub.cuh>
#include "../../include/ema_b200.h"
#include "runtime.cuh"
#include "host/indexbuild.hpp"

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

namespace {

constexpr int KEY_BASES = 29;

struct Text {
	const uint64_t *w;   // bases 32i .. 32i+31 of T in word i, base 32i in the top two bits; zero beyond n (two spare words)
	uint64_t n;          // 2 * l_pac
};

__device__ __forceinline__ uint64_t key_at(const Text &t, uint64_t p)
{
	if (p >= t.n) return 0;
	const uint64_t wi = p >> 5;
	const int sh = (int)(p & 31) << 1;
	const uint64_t hi = t.w[wi], lo = t.w[wi + 1];
	const uint64_t win = sh ? (hi << sh) | (lo >> (64 - sh)) : hi;
	const uint64_t rem = t.n - p;
	return (win & ~0x3full) | (rem < KEY_BASES ? rem : KEY_BASES);
}

__device__ __forceinline__ int text_base(const Text &t, uint64_t p) { return (int)(t.w[p >> 5] >> ((~p & 31) << 1)) & 3; }

__device__ __forceinline__ int pac_base(const uint8_t *pac, uint64_t l) { return (pac[l >> 2] >> ((~l & 3) << 1)) & 3; }

// T = forward strand followed by its reverse complement (bns_fasta2bntseq with for_only = 0, bwa/bntseq.c:306-312)
__global__ void k_pack_text(const uint8_t *pac, uint64_t l_pac, uint64_t *w, uint64_t n_words)
{
	const uint64_t n = l_pac << 1;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t v = 0;
		const uint64_t b0 = i << 5;
		if (b0 + 32 <= l_pac) {  // eight whole bytes of the forward strand
			for (int k = 0; k < 8; ++k) v = v << 8 | pac[(b0 >> 2) + k];
		} else {
			for (int k = 0; k < 32; ++k) {
				const uint64_t p = b0 + k;
				int c = 0;
				if (p < l_pac) c = pac_base(pac, p);
				else if (p < n) c = 3 - pac_base(pac, n - 1 - p);
				v = v << 2 | (uint64_t)c;
			}
		}
		w[i] = v;
	}
}

__global__ void k_hist(Text t, int bits, unsigned long long *hist)
{
	extern __shared__ unsigned int sh[];
	const int nb = 1 << bits;
	for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < t.n; p += (uint64_t)gridDim.x * blockDim.x)
		atomicAdd(&sh[key_at(t, p) >> (64 - bits)], 1u);
	__syncthreads();
	for (int i = threadIdx.x; i < nb; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// append with one atomic per warp; the order inside the output does not matter (it is sorted next)
__device__ __forceinline__ unsigned long long warp_slot(bool take, unsigned long long *counter)
{
	const unsigned mask = __ballot_sync(0xffffffffu, take);
	if (!mask) return 0;
	const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
	unsigned long long base = 0;
	if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + __popc(mask & ((1u << lane) - 1));
}

template <class PosT>
__global__ void k_select(Text t, int bits, unsigned chunk, uint64_t *keys, PosT *vals, unsigned long long *counter)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t n_round = (t.n + 31) & ~31ull;
	for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n_round; p += stride) {
		uint64_t key = 0;
		bool take = false;
		if (p < t.n) {
			key = key_at(t, p);
			take = bits == 0 || (unsigned)(key >> (64 - bits)) == chunk;
		}
		const unsigned long long slot = warp_slot(take, counter);
		if (take) { keys[slot] = key; vals[slot] = (PosT)p; }
	}
}

struct HeadOfKeys {  // k if keys[k] starts a run of equal keys, else 0: the input of the max-scan that names tie groups
	const uint64_t *keys;
	__host__ __device__ unsigned operator()(unsigned k) const { return (k == 0 || keys[k] != keys[k - 1]) ? k : 0u; }
};

// after the first sort of a chunk: unique keys are final, the others join the worklist with their group's first rank
template <class PosT>
__global__ void k_split_first(const uint64_t *keys, const PosT *vals, const unsigned *first, unsigned m, PosT *sa,
                              unsigned *u_grp, PosT *u_pos, unsigned long long *counter)
{
	const unsigned m_round = (m + 31) & ~31u;
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m_round; k += gridDim.x * blockDim.x) {
		bool tied = false;
		if (k < m) {
			const bool head = k == 0 || keys[k] != keys[k - 1];
			const bool next_head = k + 1 == m || keys[k + 1] != keys[k];
			tied = !(head && next_head);
			if (!tied) sa[k] = vals[k];
		}
		const unsigned long long slot = warp_slot(tied, counter);
		if (tied) { u_grp[slot] = first[k]; u_pos[slot] = vals[k]; }
	}
}

template <class PosT>
__global__ void k_rekey(Text t, const PosT *pos, unsigned m, uint64_t depth, uint64_t *keys, unsigned *perm)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
		keys[k] = key_at(t, (uint64_t)pos[k] + depth);
		perm[k] = k;
	}
}

template <class A, class B>
__global__ void k_gather2(const unsigned *perm, unsigned m, const A *a_in, A *a_out, const B *b_in, B *b_out)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
		const unsigned s = perm[k];
		a_out[k] = a_in[s]; b_out[k] = b_in[s];
	}
}

__global__ void k_iota(unsigned *p, unsigned m)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) p[k] = k;
}

struct HeadOfGroups {
	const unsigned *grp;
	__host__ __device__ unsigned operator()(unsigned k) const { return (k == 0 || grp[k] != grp[k - 1]) ? k : 0u; }
};
struct HeadOfSubgroups {
	const unsigned *grp; const uint64_t *keys;
	__host__ __device__ unsigned operator()(unsigned k) const { return (k == 0 || grp[k] != grp[k - 1] || keys[k] != keys[k - 1]) ? k : 0u; }
};

// one refinement round after the (group, key) sort: g_first / s_first = worklist index where the element's group /
// sub-group starts.  A sub-group of one is final at rank group + (s_first - g_first); larger ones stay, renamed.
template <class PosT>
__global__ void k_split_round(const unsigned *grp, const uint64_t *keys, const PosT *pos, const unsigned *g_first, const unsigned *s_first,
                              unsigned m, PosT *sa, unsigned *o_grp, PosT *o_pos, unsigned long long *counter)
{
	const unsigned m_round = (m + 31) & ~31u;
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m_round; k += gridDim.x * blockDim.x) {
		bool tied = false;
		unsigned rank = 0;
		if (k < m) {
			const bool head = k == 0 || grp[k] != grp[k - 1] || keys[k] != keys[k - 1];
			const bool next_head = k + 1 == m || grp[k + 1] != grp[k] || keys[k + 1] != keys[k];
			rank = grp[k] + (s_first[k] - g_first[k]);
			tied = !(head && next_head);
			if (!tied) sa[rank] = pos[k];
		}
		const unsigned long long slot = warp_slot(tied, counter);
		if (tied) { o_grp[slot] = rank; o_pos[slot] = pos[k]; }
	}
}

// rank r of the chunk is row base + r + 1 of the suffix array with the sentinel row 0 (bwa/is.c:210-220)
template <class PosT>
__global__ void k_emit(Text t, const PosT *sa, unsigned m, uint64_t base, uint8_t *bwt_sym, uint64_t *samples, unsigned long long *primary)
{
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
		const uint64_t row = base + k + 1, p = (uint64_t)sa[k];
		if (p == 0) { *primary = row; bwt_sym[row] = 0; }
		else bwt_sym[row] = (uint8_t)text_base(t, p - 1);
		if ((row & 31) == 0) samples[row >> 5] = p;   // bwt_cal_sa with intv = 32 (bwa/bwt.c:76-82)
	}
}

__global__ void k_first_row(Text t, uint8_t *bwt_sym, uint64_t *samples)
{
	bwt_sym[0] = (uint8_t)text_base(t, t.n - 1);
	samples[0] = ~0ull;   // bwt->sa[0] = (bwtint_t)-1 (bwa/bwt.c:83)
}

// the BWT without its primary row, 16 symbols per word, first symbol in the top bits (bwa/is.c:219-220, bwtindex.c:113-114)
__global__ void k_pack_bwt(const uint8_t *sym, uint64_t n, uint64_t primary, uint32_t *words, uint64_t n_words)
{
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t v = 0;
		for (int k = 0; k < 16; ++k) {
			const uint64_t j = (w << 4) + k;
			uint32_t c = 0;
			if (j < n) c = sym[j < primary ? j : j + 1];
			v = v << 2 | c;
		}
		words[w] = v;
	}
}

__global__ void k_block_counts(const uint32_t *words, uint64_t n, uint64_t n_blocks, unsigned long long *c0, unsigned long long *c1,
                               unsigned long long *c2, unsigned long long *c3)
{
	for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_blocks; b += (uint64_t)gridDim.x * blockDim.x) {
		unsigned c[4] = {0, 0, 0, 0};
		const uint64_t beg = b << 7, end = beg + 128 < n ? beg + 128 : n;
		for (uint64_t j = beg; j < end; ++j) ++c[(words[j >> 4] >> ((~j & 15) << 1)) & 3];
		c0[b] = c[0]; c1[b] = c[1]; c2[b] = c[2]; c3[b] = c[3];
	}
}

// bwt_bwtupdate_core's layout (bwa/bwtindex.c:159-168): before every 128 symbols the four cumulative counts as
// u64, then the symbols' words; after the last symbol the totals
__global__ void k_interleave(const uint32_t *words, uint64_t n, uint64_t n_blocks, const unsigned long long *c0, const unsigned long long *c1,
                             const unsigned long long *c2, const unsigned long long *c3, const unsigned long long *totals, uint32_t *out)
{
	const uint64_t n_words = (n + 15) >> 4;
	for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= n_blocks; b += (uint64_t)gridDim.x * blockDim.x) {
		if (b == n_blocks) {  // "the last element"; n_words may be odd, so the u64 totals are stored as u32 halves
			uint32_t *o = out + n_words + 8 * n_blocks;
			for (int k = 0; k < 4; ++k) { o[2 * k] = (uint32_t)totals[k]; o[2 * k + 1] = (uint32_t)(totals[k] >> 32); }
			continue;
		}
		uint32_t *o = out + 16 * b;
		unsigned long long *oc = (unsigned long long *)o;
		oc[0] = c0[b]; oc[1] = c1[b]; oc[2] = c2[b]; oc[3] = c3[b];
		for (int k = 0; k < 8; ++k) {
			const uint64_t w = 8 * b + k;
			if (w < n_words) o[8 + k] = words[w];
		}
	}
}

struct Dev {  // owns a device allocation for the length of the build
	void *p = nullptr;
	~Dev() { if (p) cudaFree(p); }
	int alloc(size_t bytes)
	{
		cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
		if (e != cudaSuccess) { snprintf(emab_errbuf, sizeof emab_errbuf, "index build: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); p = nullptr; return EMAB_ERR_NOMEM; }
		return EMAB_OK;
	}
	template <class T> T *as() const { return (T *)p; }
};

double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int grid_for(uint64_t n, int block = 256)
{
	uint64_t g = (n + block - 1) / block;
	const uint64_t cap = 148 * 16;   // a few waves of grid-stride blocks
	return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <class PosT>
int build_sa_and_emit(const Text &text, int chunk_bits, uint8_t *d_sym, uint64_t *d_samples, unsigned long long *d_primary,
                      emab_index_build_stats_t *st)
{
	const int n_chunks = 1 << chunk_bits;
	std::vector<unsigned long long> hist(n_chunks, text.n);
	Dev d_hist, d_counter;
	TRY(d_counter.alloc(8));
	if (chunk_bits) {
		TRY(d_hist.alloc(n_chunks * 8));
		CUDA_TRY(cudaMemset(d_hist.p, 0, n_chunks * 8));
		k_hist<<<grid_for(text.n), 256, n_chunks * 4>>>(text, chunk_bits, d_hist.as<unsigned long long>());
		CUDA_TRY(cudaMemcpy(hist.data(), d_hist.p, n_chunks * 8, cudaMemcpyDeviceToHost));
	}
	unsigned long long cap = 0;
	for (auto h : hist) cap = h > cap ? h : cap;
	if (cap >= 0x7fffffffull) { snprintf(emab_errbuf, sizeof emab_errbuf, "index build: a chunk of %llu suffixes is too large; raise EMAB_INDEX_CHUNK_BITS", cap); return EMAB_ERR_OVERFLOW; }
	st->max_chunk = (int64_t)cap;
	// chunk buffers: double-buffered (key, position) pairs, the chunk's suffix array, tie-group names, the worklist
	Dev k0, k1, v0, v1, sa, first, scan2, u_grp0, u_grp1, u_pos0, u_pos1, perm0, perm1, tmp;
	TRY(k0.alloc(cap * 8)); TRY(k1.alloc(cap * 8));
	TRY(v0.alloc(cap * sizeof(PosT))); TRY(v1.alloc(cap * sizeof(PosT)));
	TRY(sa.alloc(cap * sizeof(PosT)));
	TRY(first.alloc(cap * 4));
	size_t tmp_bytes = 0, need = 0;
	{
		cub::DoubleBuffer<uint64_t> kb(k0.as<uint64_t>(), k1.as<uint64_t>());
		cub::DoubleBuffer<PosT> vb(v0.as<PosT>(), v1.as<PosT>());
		cub::DeviceRadixSort::SortPairs(nullptr, need, kb, vb, (int)cap, 0, 64);
		tmp_bytes = need;
		cub::DoubleBuffer<unsigned> pb(nullptr, nullptr);
		cub::DeviceRadixSort::SortPairs(nullptr, need, kb, pb, (int)cap, 0, 64);
		tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
		cub::DeviceScan::InclusiveScan(nullptr, need, (unsigned *)nullptr, (unsigned *)nullptr, cub::Max(), (int)cap);
		tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
	}
	TRY(tmp.alloc(tmp_bytes + 256));
	uint64_t base = 0;
	size_t u_cap = 0;
	for (int c = 0; c < n_chunks; ++c) {
		CUDA_TRY(cudaMemset(d_counter.p, 0, 8));
		k_select<PosT><<<grid_for(text.n), 256>>>(text, chunk_bits, (unsigned)c, k0.as<uint64_t>(), v0.as<PosT>(), d_counter.as<unsigned long long>());
		unsigned long long m64 = 0;
		CUDA_TRY(cudaMemcpy(&m64, d_counter.p, 8, cudaMemcpyDeviceToHost));
		if (m64 != hist[c] && chunk_bits) { snprintf(emab_errbuf, sizeof emab_errbuf, "index build: chunk %d selected %llu of %llu suffixes", c, m64, hist[c]); return EMAB_ERR_CUDA; }
		const unsigned m = (unsigned)m64;
		if (m == 0) continue;
		cub::DoubleBuffer<uint64_t> kb(k0.as<uint64_t>(), k1.as<uint64_t>());
		cub::DoubleBuffer<PosT> vb(v0.as<PosT>(), v1.as<PosT>());
		size_t tb = tmp_bytes;
		cub::DeviceRadixSort::SortPairs(tmp.p, tb, kb, vb, (int)m, 0, 64 - chunk_bits);
		const uint64_t *keys = kb.Current();
		const PosT *vals = vb.Current();
		{
			cub::CountingInputIterator<unsigned> cnt(0);
			cub::TransformInputIterator<unsigned, HeadOfKeys, cub::CountingInputIterator<unsigned>> in(cnt, HeadOfKeys{keys});
			tb = tmp_bytes;
			cub::DeviceScan::InclusiveScan(tmp.p, tb, in, first.as<unsigned>(), cub::Max(), (int)m);
		}
		CUDA_TRY(cudaMemset(d_counter.p, 0, 8));
		if (u_cap == 0) {  // the worklist can hold a whole chunk (a text that is one long repeat)
			u_cap = cap;
			TRY(u_grp0.alloc(u_cap * 4)); TRY(u_grp1.alloc(u_cap * 4));
			TRY(u_pos0.alloc(u_cap * sizeof(PosT))); TRY(u_pos1.alloc(u_cap * sizeof(PosT)));
			TRY(perm0.alloc(u_cap * 4)); TRY(perm1.alloc(u_cap * 4));
			TRY(scan2.alloc(u_cap * 4));
		}
		k_split_first<PosT><<<grid_for(m), 256>>>(keys, vals, first.as<unsigned>(), m, sa.as<PosT>(), u_grp0.as<unsigned>(), u_pos0.as<PosT>(),
		                                         d_counter.as<unsigned long long>());
		unsigned long long mu64 = 0;
		CUDA_TRY(cudaMemcpy(&mu64, d_counter.p, 8, cudaMemcpyDeviceToHost));
		st->n_tied += (int64_t)mu64;
		// refinement rounds; from here k0/k1 (free again) hold the round's keys, `first` the group starts
		unsigned mu = (unsigned)mu64;
		unsigned *g_cur = u_grp0.as<unsigned>(), *g_alt = u_grp1.as<unsigned>();
		PosT *p_cur = u_pos0.as<PosT>(), *p_alt = u_pos1.as<PosT>();
		int round = 0;
		while (mu) {
			++round;
			if ((uint64_t)round * KEY_BASES > text.n + KEY_BASES) { snprintf(emab_errbuf, sizeof emab_errbuf, "index build: tie refinement did not terminate"); return EMAB_ERR_CUDA; }
			const int g = grid_for(mu);
			// 1. sort by the next 29 bases, carrying (group, position) through a permutation
			k_rekey<PosT><<<g, 256>>>(text, p_cur, mu, (uint64_t)round * KEY_BASES, k0.as<uint64_t>(), perm0.as<unsigned>());
			cub::DoubleBuffer<uint64_t> rk(k0.as<uint64_t>(), k1.as<uint64_t>());
			cub::DoubleBuffer<unsigned> rp(perm0.as<unsigned>(), perm1.as<unsigned>());
			tb = tmp_bytes;
			cub::DeviceRadixSort::SortPairs(tmp.p, tb, rk, rp, (int)mu, 0, 64);
			k_gather2<unsigned, PosT><<<g, 256>>>(rp.Current(), mu, g_cur, g_alt, p_cur, p_alt);
			// 2. stable sort by group: (group, key) order
			unsigned *perm_a = rp.Alternate(), *perm_b = rp.Current();
			k_iota<<<g, 256>>>(perm_a, mu);
			cub::DoubleBuffer<unsigned> gk(g_alt, g_cur);
			cub::DoubleBuffer<unsigned> gp(perm_a, perm_b);
			tb = tmp_bytes;
			cub::DeviceRadixSort::SortPairs(tmp.p, tb, gk, gp, (int)mu, 0, 32);
			unsigned *grp_sorted = gk.Current(), *grp_free = gk.Alternate();   // the old worklist's group array is free again
			uint64_t *key_sorted = rk.Alternate();
			PosT *pos_sorted = p_cur;                                         // likewise its positions
			k_gather2<uint64_t, PosT><<<g, 256>>>(gp.Current(), mu, rk.Current(), key_sorted, p_alt, pos_sorted);
			// 3. group / sub-group starts, then split
			{
				cub::CountingInputIterator<unsigned> cnt(0);
				cub::TransformInputIterator<unsigned, HeadOfGroups, cub::CountingInputIterator<unsigned>> in_g(cnt, HeadOfGroups{grp_sorted});
				tb = tmp_bytes;
				cub::DeviceScan::InclusiveScan(tmp.p, tb, in_g, first.as<unsigned>(), cub::Max(), (int)mu);
				cub::TransformInputIterator<unsigned, HeadOfSubgroups, cub::CountingInputIterator<unsigned>> in_s(cnt, HeadOfSubgroups{grp_sorted, key_sorted});
				tb = tmp_bytes;
				cub::DeviceScan::InclusiveScan(tmp.p, tb, in_s, scan2.as<unsigned>(), cub::Max(), (int)mu);
			}
			CUDA_TRY(cudaMemset(d_counter.p, 0, 8));
			k_split_round<PosT><<<g, 256>>>(grp_sorted, key_sorted, pos_sorted, first.as<unsigned>(), scan2.as<unsigned>(), mu, sa.as<PosT>(),
			                               grp_free, p_alt, d_counter.as<unsigned long long>());
			CUDA_TRY(cudaMemcpy(&mu64, d_counter.p, 8, cudaMemcpyDeviceToHost));
			mu = (unsigned)mu64;
			// next round reads (grp_free, p_alt)
			g_cur = grp_free; g_alt = grp_sorted;
			PosT *t2 = p_cur; p_cur = p_alt; p_alt = t2;
			// k0/k1: the key double buffer is reused as is next round (rekey writes k0)
		}
		if (round > st->max_rounds) st->max_rounds = round;
		k_emit<PosT><<<grid_for(m), 256>>>(text, sa.as<PosT>(), m, base, d_sym, d_samples, d_primary);
		CUDA_TRY(cudaGetLastError());
		base += m;
		++st->n_chunks;
	}
	CUDA_TRY(cudaDeviceSynchronize());
	if (base != text.n) { snprintf(emab_errbuf, sizeof emab_errbuf, "index build: %llu of %llu suffixes placed", (unsigned long long)base, (unsigned long long)text.n); return EMAB_ERR_CUDA; }
	return EMAB_OK;
}

int write_all(FILE *f, const void *p, size_t bytes) { return fwrite(p, 1, bytes, f) == bytes ? 0 : -1; }

// device -> file through a pinned staging buffer
int dump_device(FILE *f, const void *d, size_t bytes)
{
	const size_t piece = 64u << 20;
	void *h = nullptr;
	if (cudaMallocHost(&h, piece) != cudaSuccess) return -1;
	int rc = 0;
	for (size_t o = 0; o < bytes && !rc; o += piece) {
		const size_t n = bytes - o < piece ? bytes - o : piece;
		if (cudaMemcpy(h, (const char *)d + o, n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
		else rc = write_all(f, h, n);
	}
	cudaFreeHost(h);
	return rc;
}

}  // namespace

extern "C" int emab_index_pack_fasta(const char *fasta_path, const char *prefix)
{
	if (!fasta_path) return EMAB_ERR_ARG;
	if (!prefix) prefix = fasta_path;
	emab::PackedRef ref;
	std::string err;
	int rc = emab::pack_fasta(fasta_path, &ref, &err);
	if (!rc) rc = emab::write_pac_ann_amb(ref, prefix, &err);
	if (rc) snprintf(emab_errbuf, sizeof emab_errbuf, "%s", err.c_str());
	return rc;
}

extern "C" int emab_index_build(const char *fasta_path, const char *prefix, int device, emab_index_build_stats_t *stats)
{
	if (!fasta_path) return EMAB_ERR_ARG;
	if (!prefix) prefix = fasta_path;
	emab_index_build_stats_t local;
	emab_index_build_stats_t *st = stats ? stats : &local;
	memset(st, 0, sizeof *st);
	CUDA_TRY(cudaSetDevice(device));
	const double t0 = now_ms();
	emab::PackedRef ref;
	std::string err;
	int rc = emab::pack_fasta(fasta_path, &ref, &err);
	if (!rc) rc = emab::write_pac_ann_amb(ref, prefix, &err);
	if (rc) { snprintf(emab_errbuf, sizeof emab_errbuf, "%s", err.c_str()); return rc; }
	const double t1 = now_ms();
	st->l_pac = ref.l_pac; st->n_seqs = (int32_t)ref.contigs.size(); st->n_holes = (int32_t)ref.holes.size();
	st->ms_pack = t1 - t0;
	const uint64_t l_pac = (uint64_t)ref.l_pac, n = l_pac << 1;
	const uint64_t n_text_words = (n >> 5) + 3;
	Dev d_pac, d_text, d_sym, d_samples, d_primary;
	TRY(d_pac.alloc(ref.pac.size() + 8));
	CUDA_TRY(cudaMemset(d_pac.p, 0, ref.pac.size() + 8));
	CUDA_TRY(cudaMemcpy(d_pac.p, ref.pac.data(), ref.pac.size(), cudaMemcpyHostToDevice));
	TRY(d_text.alloc(n_text_words * 8));
	k_pack_text<<<grid_for(n_text_words), 256>>>(d_pac.as<uint8_t>(), l_pac, d_text.as<uint64_t>(), n_text_words);
	const uint64_t n_sa = (n + 32) / 32;   // bwt_cal_sa (bwa/bwt.c:73)
	TRY(d_sym.alloc(n + 2));
	TRY(d_samples.alloc(n_sa * 8));
	TRY(d_primary.alloc(8));
	CUDA_TRY(cudaMemset(d_primary.p, 0xff, 8));
	Text text{d_text.as<uint64_t>(), n};
	k_first_row<<<1, 1>>>(text, d_sym.as<uint8_t>(), d_samples.as<uint64_t>());
	int chunk_bits = 0;
	while ((n >> chunk_bits) > 400000000ull && chunk_bits < 12) ++chunk_bits;
	if (const char *e = getenv("EMAB_INDEX_CHUNK_BITS")) { chunk_bits = atoi(e); chunk_bits = chunk_bits < 0 ? 0 : (chunk_bits > 12 ? 12 : chunk_bits); }
	if (n < (1ull << 32)) rc = build_sa_and_emit<uint32_t>(text, chunk_bits, d_sym.as<uint8_t>(), d_samples.as<uint64_t>(), d_primary.as<unsigned long long>(), st);
	else rc = build_sa_and_emit<uint64_t>(text, chunk_bits, d_sym.as<uint8_t>(), d_samples.as<uint64_t>(), d_primary.as<unsigned long long>(), st);
	if (rc) return rc;
	unsigned long long primary = 0;
	CUDA_TRY(cudaMemcpy(&primary, d_primary.p, 8, cudaMemcpyDeviceToHost));
	if (primary == ~0ull) { snprintf(emab_errbuf, sizeof emab_errbuf, "index build: primary row not found"); return EMAB_ERR_CUDA; }
	const double t2 = now_ms();
	st->ms_sort = t2 - t1;
	// pack, count, interleave
	const uint64_t n_words = (n + 15) >> 4, n_blocks = (n + 127) >> 7, n_occ = n_blocks + 1;
	const uint64_t bwt_size = n_words + n_occ * 8;   // in u32 (bwa/bwtindex.c:155-156)
	Dev d_words, d_c[4], d_tot, d_out, d_tmp;
	TRY(d_words.alloc(n_words * 4 + 64));
	k_pack_bwt<<<grid_for(n_words), 256>>>(d_sym.as<uint8_t>(), n, primary, d_words.as<uint32_t>(), n_words);
	CUDA_TRY(cudaDeviceSynchronize());
	cudaFree(d_sym.p); d_sym.p = nullptr;
	for (int k = 0; k < 4; ++k) TRY(d_c[k].alloc((n_blocks + 1) * 8));
	k_block_counts<<<grid_for(n_blocks), 256>>>(d_words.as<uint32_t>(), n, n_blocks, d_c[0].as<unsigned long long>(), d_c[1].as<unsigned long long>(),
	                                           d_c[2].as<unsigned long long>(), d_c[3].as<unsigned long long>());
	size_t tb = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, tb, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)(n_blocks + 1));
	TRY(d_tmp.alloc(tb + 256));
	for (int k = 0; k < 4; ++k) {  // element n_blocks of the exclusive sum is the total
		CUDA_TRY(cudaMemset((char *)d_c[k].p + n_blocks * 8, 0, 8));
		size_t t = tb;
		cub::DeviceScan::ExclusiveSum(d_tmp.p, t, d_c[k].as<unsigned long long>(), d_c[k].as<unsigned long long>(), (int)(n_blocks + 1));
	}
	unsigned long long tot[4];
	for (int k = 0; k < 4; ++k) CUDA_TRY(cudaMemcpy(&tot[k], (char *)d_c[k].p + n_blocks * 8, 8, cudaMemcpyDeviceToHost));
	TRY(d_tot.alloc(32));
	CUDA_TRY(cudaMemcpy(d_tot.p, tot, 32, cudaMemcpyHostToDevice));
	TRY(d_out.alloc(bwt_size * 4));
	k_interleave<<<grid_for(n_blocks + 1), 256>>>(d_words.as<uint32_t>(), n, n_blocks, d_c[0].as<unsigned long long>(), d_c[1].as<unsigned long long>(),
	                                             d_c[2].as<unsigned long long>(), d_c[3].as<unsigned long long>(), d_tot.as<unsigned long long>(), d_out.as<uint32_t>());
	CUDA_TRY(cudaDeviceSynchronize());
	CUDA_TRY(cudaGetLastError());
	const double t3 = now_ms();
	st->ms_occ = t3 - t2;
	uint64_t head[5] = {primary, tot[0], tot[0] + tot[1], tot[0] + tot[1] + tot[2], tot[0] + tot[1] + tot[2] + tot[3]};
	if (head[4] != n) { snprintf(emab_errbuf, sizeof emab_errbuf, "index build: symbol counts do not add up"); return EMAB_ERR_CUDA; }
	const std::string pre(prefix);
	FILE *f = fopen((pre + ".bwt").c_str(), "wb");
	if (!f) { snprintf(emab_errbuf, sizeof emab_errbuf, "cannot write %s.bwt", prefix); return EMAB_ERR_IO; }
	int wrc = write_all(f, head, 40);
	if (!wrc) wrc = dump_device(f, d_out.p, bwt_size * 4);
	if (fclose(f) || wrc) { snprintf(emab_errbuf, sizeof emab_errbuf, "write error on %s.bwt", prefix); return EMAB_ERR_IO; }
	f = fopen((pre + ".sa").c_str(), "wb");
	if (!f) { snprintf(emab_errbuf, sizeof emab_errbuf, "cannot write %s.sa", prefix); return EMAB_ERR_IO; }
	const uint64_t tail[2] = {32, n};   // sa_intv, seq_len (bwa/bwt.c:402-403)
	wrc = write_all(f, head, 40);
	if (!wrc) wrc = write_all(f, tail, 16);
	if (!wrc) wrc = dump_device(f, d_samples.as<uint64_t>() + 1, (n_sa - 1) * 8);
	if (fclose(f) || wrc) { snprintf(emab_errbuf, sizeof emab_errbuf, "write error on %s.sa", prefix); return EMAB_ERR_IO; }
	st->ms_write = now_ms() - t3;
	st->ms_total = now_ms() - t0;
	st->primary = (int64_t)primary;
	st->chunk_bits = chunk_bits;
	return EMAB_OK;
}
