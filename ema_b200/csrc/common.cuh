// Shared types and helpers for the ema_b200 device code (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define EMAB_OK 0
#define EMAB_ERR_IO (-1)
#define EMAB_ERR_CUDA (-2)
#define EMAB_ERR_ARG (-3)
#define EMAB_ERR_OVERFLOW (-4)
#define EMAB_ERR_NOMEM (-5)

#define EMAB_MAX_READ_LEN 256   // reads longer than this are rejected (the reference caps at 200, include/align.h:61)
#define EMAB_MAX_INTV 128       // SA intervals kept per read after mem_collect_intv (overflow is reported, never truncated silently)
#define EMAB_MAX_CIGAR 64       // same bound the reference asserts for XA (include/align.h:41, src/samdict.c:199)

extern thread_local char emab_errbuf[512];

#define CUDA_TRY(expr)                                                                      \
	do {                                                                                    \
		cudaError_t e__ = (expr);                                                           \
		if (e__ != cudaSuccess) {                                                           \
			snprintf(emab_errbuf, sizeof emab_errbuf, "%s:%d: %s -> %s", __FILE__, __LINE__, \
			         #expr, cudaGetErrorString(e__));                                       \
			return EMAB_ERR_CUDA;                                                           \
		}                                                                                   \
	} while (0)

// BWA-MEM options exactly as mem_opt_init() leaves them (bwa/bwamem.c:74-110) with EMA's single
// override max_occ = 3000 (src/align.c:185).  Kept as compile-time constants: the reference never
// changes them, and constants let ptxas fold the band/penalty arithmetic.
namespace opt {
constexpr int a = 1, b = 4;
constexpr int o_del = 6, e_del = 1, o_ins = 6, e_ins = 1;
constexpr int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
constexpr int w = 100;
constexpr int zdrop = 100;
constexpr int pen_clip5 = 5, pen_clip3 = 5;
constexpr int min_seed_len = 19;
constexpr int split_width = 10;
constexpr int split_len = 28;        // (int)(19 * 1.5 + .499)
constexpr int max_mem_intv = 20;
constexpr int max_occ = 3000;
constexpr int max_chain_gap = 10000;
constexpr float mask_level = 0.50f;
constexpr float drop_ratio = 0.50f;
constexpr float mask_level_redun = 0.95f;
constexpr int min_chain_weight = 0;
constexpr int max_matesw = 50;
constexpr int mapQ_coef_len = 50;    // float 50 in the reference
constexpr int mapQ_coef_fac = 3;     // (int)log(50)  (bwa/bwamem.c:107, bwa/bwamem.h:79)
}  // namespace opt

// ---------------------------------------------------------------------------------------------
// Index resident in HBM.  Layouts are the on-disk layouts of `bwa index` so the
// files are uploaded verbatim; the only derived structure is the dense suffix array.
// ---------------------------------------------------------------------------------------------
struct DevIndex {
	const uint4 *bwt;        // 64-byte Occ blocks: 4 x u64 counts + 8 x u32 of 16 symbols  (bwa/bwt.h:74-75)
	uint64_t n_blocks;
	uint64_t primary;
	uint64_t seq_len;        // 2 * l_pac
	uint64_t L2[5];
	const uint32_t *sa32;    // dense SA (interval 1) when seq_len < 2^32, else null
	const uint64_t *sa64;    // dense SA otherwise
	const uint64_t *sa_sampled;  // the .sa file's samples (interval sa_intv), kept for bwt_sa parity tests
	int sa_intv;
	const uint8_t *pac;      // forward strand, 2 bits/base, MSB first  (bwa/bntseq.c:229-230)
	int64_t l_pac;
	int n_seqs;
	const int64_t *ann_offset;   // n_seqs
	const int32_t *ann_len;      // n_seqs
	int seed_load_both;          // tuning/measurement knob (EMAB_SEED_LOAD_BOTH=1): bwt_2occ4 requests its second block even when it is the first
	// derived at load for the default seeding form (seed_hot.cuh); null / 0 when not built
	const uint4 *hot;            // one-hot Occ blocks: per 64 symbols, per base {u64 count before the block, u64 "symbol i is this base"}
	const uint4 *kmer;           // packed intervals of every k-mer, levels 1..kmer_k back to back (kmer_level_off)
	int kmer_k;
};

struct Intv {  // bwtintv_t (bwa/bwt.h:62-64)
	uint64_t x0, x1, x2, info;
};

struct Seed {  // mem_seed_t (bwa/bwamem.c:194-198)
	int64_t rbeg;
	int32_t qbeg, len;
	int32_t score;
	int32_t next;  // chain-local linked list while chaining
};

struct Chain {  // mem_chain_t (bwa/bwamem.c:200-206) in SoA-friendly form
	int64_t pos;
	int32_t n, rid;
	int32_t w, kept, first;
	int32_t seed_beg;   // first seed (index into the read's seed segment) ; list head while chaining
	int32_t seed_last;  // list tail while chaining
	int32_t qbeg0, qend_last;  // chn_beg / chn_end caches
	float frac_rep;
};

struct Reg {  // mem_alnreg_t (bwa/bwamem.h:92-110)
	int64_t rb, re;
	int32_t qb, qe;
	int32_t rid;
	int32_t score, truesc;
	int32_t sub, csub, sub_n;
	int32_t w, seedcov;
	int32_t secondary, seedlen0;
	int32_t n_comp;
	float frac_rep;
};

struct ExtResult { int score, qle, tle, gtle, gscore, max_off; };  // ksw_extend2 outputs
struct LocResult { int score, te, qe, score2, te2, tb, qb; };     // kswr_t (bwa/ksw.h:14-19)

struct Aln {  // what append_alignments keeps of mem_aln_t / SingleReadAlignment / SAMRecord (src/align.c:915-956)
	int64_t pos;       // 0-based leftmost position on the contig
	int32_t rid;
	int32_t is_rev;
	int32_t NM;
	int32_t n_cigar;
	int32_t score;     // SW score of the region
	int32_t mapq;      // mem_approx_mapq_se_insist
	int32_t score_mapq;
	int32_t clip, clip_edit_dist;
	int32_t keep;      // 1 if it survives the append_alignments filters
	double em_score;
	uint32_t cigar[EMAB_MAX_CIGAR];
};

// ---------------------------------------------------------------------------------------------
// small helpers.  EMAB_HD functions are thread-scalar and also compile for the host, where
// tests/hostsim drives them for GPU-less unit tests of the control logic (never a product path).
// ---------------------------------------------------------------------------------------------
#define EMAB_HD __host__ __device__ inline
#ifdef __CUDA_ARCH__
#define emab_popc(x) __popc(x)
#define emab_popcll(x) __popcll(x)
#else
#define emab_popc(x) __builtin_popcount(x)
#define emab_popcll(x) __builtin_popcountll(x)
#endif
EMAB_HD int imax(int a, int b) { return a > b ? a : b; }
EMAB_HD int imin(int a, int b) { return a < b ? a : b; }
EMAB_HD int64_t lmax(int64_t a, int64_t b) { return a > b ? a : b; }
EMAB_HD int64_t lmin(int64_t a, int64_t b) { return a < b ? a : b; }

EMAB_HD int64_t bns_depos(int64_t l_pac, int64_t pos, int *is_rev)
{  // bwa/bntseq.h:87-90
	*is_rev = pos >= l_pac;
	return *is_rev ? (l_pac << 1) - 1 - pos : pos;
}

// base of the forward-reverse reference at coordinate p in [0, 2*l_pac)  (bwa/bntseq.c:403-424)
EMAB_HD int ref_base(const DevIndex &ix, int64_t p)
{
	if (p >= ix.l_pac) {
		int64_t f = (ix.l_pac << 1) - 1 - p;
		return 3 - ((ix.pac[f >> 2] >> ((~f & 3) << 1)) & 3);
	}
	return (ix.pac[p >> 2] >> ((~p & 3) << 1)) & 3;
}

EMAB_HD int bns_pos2rid(const DevIndex &ix, int64_t pos_f)
{  // bwa/bntseq.c:354-368 (same bisection so that out-of-range behaviour is identical)
	if (pos_f >= ix.l_pac) return -1;
	int left = 0, mid = 0, right = ix.n_seqs;
	while (left < right) {
		mid = (left + right) >> 1;
		if (pos_f >= ix.ann_offset[mid]) {
			if (mid == ix.n_seqs - 1) break;
			if (pos_f < ix.ann_offset[mid + 1]) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}

EMAB_HD int bns_intv2rid(const DevIndex &ix, int64_t rb, int64_t re)
{  // bwa/bntseq.c:370-378
	if (rb < ix.l_pac && re > ix.l_pac) return -2;
	int is_rev;
	int rid_b = bns_pos2rid(ix, bns_depos(ix.l_pac, rb, &is_rev));
	int rid_e = rb < re ? bns_pos2rid(ix, bns_depos(ix.l_pac, re - 1, &is_rev)) : rid_b;
	return rid_b == rid_e ? rid_b : -1;
}

// bns_fetch_seq's clamping (bwa/bntseq.c:426-451) without materialising the sequence: the DP
// kernels read bases straight from the packed reference.
EMAB_HD void bns_clamp(const DevIndex &ix, int64_t *beg, int64_t mid, int64_t *end, int *rid)
{
	int is_rev;
	*rid = bns_pos2rid(ix, bns_depos(ix.l_pac, mid, &is_rev));
	int64_t far_beg = ix.ann_offset[*rid];
	int64_t far_end = far_beg + ix.ann_len[*rid];
	if (is_rev) {
		int64_t tmp = far_beg;
		far_beg = (ix.l_pac << 1) - far_end;
		far_end = (ix.l_pac << 1) - tmp;
	}
	*beg = lmax(*beg, far_beg);
	*end = lmin(*end, far_end);
}

// substitution score of bwa_fill_scmat(1, 4) (bwa/bwa.c:136-146): 1 / -4, and -1 against N
EMAB_HD int sc_mat(int t, int q)
{
	return (t > 3 || q > 3) ? -1 : (t == q ? opt::a : -opt::b);
}
