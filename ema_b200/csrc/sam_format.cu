// SAM text on the device: print_sam_record (src/samrecord.c:104-284 as called from src/align.c:585-611) for a whole
// batch.  The host decides WHAT is printed (which candidate of each read, its mate, the XA alternative, flag, MAPQ,
// cloud id, the %.5g text of the posterior) and sends one 48-byte descriptor per record; everything of length — read
// names, bases (reverse-complemented for reverse-strand hits), qualities, CIGAR strings — is assembled here from the
// batch's text and candidate records, which are already in HBM.  Formatting and concatenating the text was the largest
// consumer of host CPU on the path (38 % of ~100 thread-ms per 40 000-pair bucket, profiles/r1n_host_profile.log) and is
// what kept several GPUs from scaling on one node.
//
//   k_sam_len    thread / record   exact byte length of the record (the emitter below with a counting writer)
//   (scan)                         record offsets, total
//   k_sam_write  warp / record     the same emitter with a writing writer: short fields by lane 0, strings lane-strided
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cub/device/device_scan.cuh>
#include "../../include/ema_b200.h"
#include "runtime.cuh"

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)
#ifndef FULL_MASK
#define FULL_MASK 0xffffffffu
#endif

namespace {

struct SamTables {   // device pointers
	const char *text;                      // the batch's input text (ids, reads, qualities)
	const emab_pair_text_t *pairs;
	const emab_cand_t *cands;
	const uint32_t *cigars;
	const char *chrom_names;               // .fai names, concatenated
	const int32_t *chrom_off;              // [n_chrom + 1]
	const int32_t *rid2chrom;              // contig of the index -> .fai name (src/main.c:41-55)
	const char *bc_text;                   // BX strings of the batch's barcodes, concatenated
	const int32_t *bc_off;                 // [n_bc + 1]
	char bx_index[16]; int bx_index_len;   // the -i suffix
	char rg_id[64]; int rg_len;            // RG:Z value; rg_len < 0: no read group
	int is_haplotag;
};

struct CountWriter {
	unsigned long long n = 0;
	__device__ void ch(char) { ++n; }
	__device__ void bytes(const char *, int len) { n += len; }
	__device__ void rev(const char *, int len) { n += len; }
	__device__ void revcomp(const char *, int len) { n += len; }
	__device__ void lit(const char *, int len) { n += len; }
};

struct WarpWriter {   // all lanes hold the same cursor; lane 0 writes single characters, strings are spread over the lanes
	char *p;
	int lane;
	__device__ void ch(char c) { if (lane == 0) *p = c; ++p; }
	__device__ void bytes(const char *s, int len) { for (int i = lane; i < len; i += 32) p[i] = s[i]; p += len; }
	__device__ void rev(const char *s, int len) { for (int i = lane; i < len; i += 32) p[i] = s[len - 1 - i]; p += len; }
	__device__ void revcomp(const char *s, int len)
	{
		for (int i = lane; i < len; i += 32) {
			const char c = s[len - 1 - i];
			p[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';   // anything else prints as N
		}
		p += len;
	}
	__device__ void lit(const char *s, int len) { if (lane == 0) for (int i = 0; i < len; ++i) p[i] = s[i]; p += len; }
};

template <class W>
__device__ void put_int(W &w, long long v)
{
	char buf[24];
	int n = 0;
	const bool neg = v < 0;
	unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
	do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	if (neg) w.ch('-');
	while (n) w.ch(buf[--n]);
}

template <class W>
__device__ void put_cigar(W &w, const emab_cand_t &a, const uint32_t *cigars)
{
	const uint32_t *cg = cigars + a.cigar_off;
	for (int i = 0; i < a.n_cigar; ++i) {
		put_int(w, cg[i] >> 4);
		const int op = cg[i] & 0xf;
		w.ch(op == 0 ? 'M' : op == 1 ? 'I' : op == 2 ? 'D' : 'S');
	}
}

__device__ int ref_len(const emab_cand_t &a, const uint32_t *cigars)
{
	const uint32_t *cg = cigars + a.cigar_off;
	int l = 0;
	for (int k = 0; k < a.n_cigar; ++k) { const int op = cg[k] & 0xf; if (op == 0 || op == 2) l += cg[k] >> 4; }
	return l;
}

template <class W>
__device__ void put_chrom(W &w, const SamTables &t, int rid)
{
	const int c = t.rid2chrom[rid];
	w.bytes(t.chrom_names + t.chrom_off[c], t.chrom_off[c + 1] - t.chrom_off[c]);
}

#define LIT(wr, s) (wr).lit(s, (int)sizeof(s) - 1)

// one record, field by field as print_sam_record prints it
template <class W>
__device__ void emit_record(W &w, const SamTables &t, const emab_sam_rec_t &d)
{
	const emab_pair_text_t &pt = t.pairs[d.pair];
	const emab_cand_t *rec = d.rec_cand >= 0 ? &t.cands[d.rec_cand] : nullptr;
	const emab_cand_t *mate = d.mate_cand >= 0 ? &t.cands[d.mate_cand] : nullptr;
	const emab_cand_t *alt = d.alt_cand >= 0 ? &t.cands[d.alt_cand] : nullptr;
	const int m = d.which;
	w.bytes(t.text + pt.id_off[m], (int)pt.id_len[m]); w.ch('\t');
	put_int(w, d.flag); w.ch('\t');
	if (rec) put_chrom(w, t, rec->rid); else w.ch('*');
	w.ch('\t'); put_int(w, rec ? rec->pos + 1 : 0); w.ch('\t'); put_int(w, d.mapq); w.ch('\t');
	if (rec) put_cigar(w, *rec, t.cigars); else w.ch('*');
	if (mate) {
		const bool same = rec && t.rid2chrom[mate->rid] == t.rid2chrom[rec->rid];
		w.ch('\t');
		if (same) w.ch('='); else put_chrom(w, t, mate->rid);
		w.ch('\t'); put_int(w, (int)(mate->pos + 1));
		if (same) {
			const long long p0 = (rec->pos + 1) + (rec->is_rev ? ref_len(*rec, t.cigars) - 1 : 0);
			const long long p1 = (mate->pos + 1) + (mate->is_rev ? ref_len(*mate, t.cigars) - 1 : 0);
			w.ch('\t');
			if (mate->n_cigar == 0 || rec->n_cigar == 0) w.ch('0');
			else put_int(w, -(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
		} else LIT(w, "\t0");
	} else LIT(w, "\t*\t0\t0");
	w.ch('\t');
	const char *read = t.text + pt.read_off[m], *qual = t.text + pt.qual_off[m];
	if (rec && rec->is_rev) { w.revcomp(read, (int)pt.read_len[m]); w.ch('\t'); w.rev(qual, (int)pt.qual_len[m]); }
	else { w.bytes(read, (int)pt.read_len[m]); w.ch('\t'); w.bytes(qual, (int)pt.qual_len[m]); }
	const char *bc = t.bc_text + t.bc_off[d.bc];
	const int bc_len = t.bc_off[d.bc + 1] - t.bc_off[d.bc];
	if (rec) {
		LIT(w, "\tNM:i:"); put_int(w, rec->NM);
		LIT(w, "\tBX:Z:"); w.bytes(bc, bc_len);
		if (!t.is_haplotag) { w.ch('-'); w.lit(t.bx_index, t.bx_index_len); }
		LIT(w, "\tXG:f:"); w.lit(d.gamma, d.gamma_len);
		LIT(w, "\tMI:i:"); put_int(w, d.mi);
		LIT(w, "\tXF:i:"); w.ch(d.xf ? '1' : '0');
	} else {
		LIT(w, "\tBX:Z:"); w.bytes(bc, bc_len);
		if (!t.is_haplotag) LIT(w, "-1");
	}
	if (t.rg_len >= 0) { LIT(w, "\tRG:Z:"); w.lit(t.rg_id, t.rg_len); }
	if (alt) {
		LIT(w, "\tXA:Z:"); put_chrom(w, t, alt->rid); w.ch(','); w.ch(alt->is_rev ? '-' : '+'); put_int(w, (int)(alt->pos + 1)); w.ch(',');
		put_cigar(w, *alt, t.cigars);
		w.ch(','); put_int(w, alt->NM); w.ch(';');
	}
	w.ch('\n');
}

__global__ void k_sam_len(SamTables t, const emab_sam_rec_t *recs, int n, unsigned long long *len)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	if (i == n) { len[i] = 0; return; }
	CountWriter w;
	emit_record(w, t, recs[i]);
	len[i] = w.n;
}

__global__ void __launch_bounds__(256)
k_sam_write(SamTables t, const emab_sam_rec_t *recs, int n, const unsigned long long *off, char *out)
{
	const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (i >= n) return;
	WarpWriter w{out + off[i], (int)(threadIdx.x & 31)};
	emit_record(w, t, recs[i]);
}

}  // namespace

// slots: 45 chrom tables (per ctx, once), 46 record descriptors + barcode strings, 47 lengths/offsets, 33.. free after the
// pipeline (the ExtPlan slot is reused for the output text: it is the largest buffer that is dead by now)
extern "C" int emab_sam_tables(emab_ctx_t *c, int n_chrom, const char *const *names, int n_rid, const int32_t *rid2chrom)
{
	CTX_ENTER(c);
	if (!c || n_chrom < 0 || n_rid < 0) return EMAB_ERR_ARG;
	std::string cat;
	std::vector<int32_t> off(n_chrom + 1, 0);
	for (int i = 0; i < n_chrom; ++i) { cat += names[i]; off[i + 1] = (int32_t)cat.size(); }
	const size_t o_off = (cat.size() + 15) & ~(size_t)15, o_rid = o_off + (size_t)(n_chrom + 1) * 4, total = o_rid + (size_t)n_rid * 4;
	std::vector<char> blob(total + 16, 0);
	memcpy(blob.data(), cat.data(), cat.size());
	memcpy(blob.data() + o_off, off.data(), off.size() * 4);
	if (n_rid) memcpy(blob.data() + o_rid, rid2chrom, (size_t)n_rid * 4);
	TRY(c->b[45].ensure(total + 16));
	CUDA_TRY(cudaMemcpy(c->b[45].p, blob.data(), total, cudaMemcpyHostToDevice));
	c->sam_chrom_off = o_off; c->sam_rid_off = o_rid; c->sam_tables_ready = true;
	return EMAB_OK;
}

extern "C" int emab_sam_format(emab_ctx_t *c, const emab_sam_job_t *job, char *out, uint64_t out_cap, uint64_t *out_len)
{
	CTX_ENTER(c);
	if (!c || !job || !out_len) return EMAB_ERR_ARG;
	*out_len = 0;
	if (!c->sam_tables_ready || !c->text_ready) { snprintf(emab_errbuf, sizeof emab_errbuf, "emab_sam_format: tables or batch text not on the device"); return EMAB_ERR_ARG; }
	const int n = job->n_recs;
	if (n == 0) return EMAB_OK;
	cudaStream_t st = c->stream;
	// one staging block: descriptors | barcode offsets | barcode text
	const size_t b_recs = (size_t)n * sizeof(emab_sam_rec_t), o_bcoff = (b_recs + 15) & ~(size_t)15, b_bcoff = (size_t)(job->n_bc + 1) * 4,
	             o_bctext = o_bcoff + ((b_bcoff + 15) & ~(size_t)15), b_bctext = (size_t)job->bc_off[job->n_bc], total = o_bctext + b_bctext;
	TRY(c->h[6].ensure(total + 16));
	memcpy(c->h[6].p, job->recs, b_recs);
	memcpy((char *)c->h[6].p + o_bcoff, job->bc_off, b_bcoff);
	memcpy((char *)c->h[6].p + o_bctext, job->bc_text, b_bctext);
	TRY(c->b[46].ensure(total + 16));
	CUDA_TRY(cudaMemcpyAsync(c->b[46].p, c->h[6].p, total, cudaMemcpyHostToDevice, st));
	SamTables t;
	memset(&t, 0, sizeof t);
	t.text = c->b[31].as<char>();
	t.pairs = c->b[27].as<emab_pair_text_t>();
	t.cands = c->b[21].as<emab_cand_t>();
	t.cigars = c->b[19].as<uint32_t>();
	t.chrom_names = c->b[45].as<char>();
	t.chrom_off = (const int32_t *)(c->b[45].as<char>() + c->sam_chrom_off);
	t.rid2chrom = (const int32_t *)(c->b[45].as<char>() + c->sam_rid_off);
	t.bc_off = (const int32_t *)(c->b[46].as<char>() + o_bcoff);
	t.bc_text = c->b[46].as<char>() + o_bctext;
	t.bx_index_len = (int)strnlen(job->bx_index ? job->bx_index : "", sizeof t.bx_index - 1);
	memcpy(t.bx_index, job->bx_index ? job->bx_index : "", (size_t)t.bx_index_len);
	t.rg_len = job->rg_id ? (int)strnlen(job->rg_id, sizeof t.rg_id - 1) : -1;
	if (t.rg_len > 0) memcpy(t.rg_id, job->rg_id, (size_t)(t.rg_len & 63));
	t.is_haplotag = job->is_haplotag;
	const emab_sam_rec_t *d_recs = c->b[46].as<emab_sam_rec_t>();
	TRY(c->b[47].ensure((size_t)(n + 1) * 8 * 2));
	unsigned long long *d_len = c->b[47].as<unsigned long long>(), *d_off = d_len + (n + 1);
	CUDA_TRY(cudaEventRecord(c->ev0, st));
	k_sam_len<<<(n + 1 + 127) / 128, 128, 0, st>>>(t, d_recs, n, d_len);
	size_t tb = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, tb, d_len, d_off, n + 1, st);
	TRY(c->b[6].ensure(tb + 16));
	cub::DeviceScan::ExclusiveSum(c->b[6].p, tb, d_len, d_off, n + 1, st);
	unsigned long long total_len = 0;
	CUDA_TRY(cudaMemcpyAsync(&total_len, d_off + n, 8, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx_wait(c));
	if (total_len > out_cap) { *out_len = total_len; snprintf(emab_errbuf, sizeof emab_errbuf, "emab_sam_format: %llu bytes do not fit the output buffer", total_len); return EMAB_ERR_OVERFLOW; }
	TRY(c->b[33].ensure((size_t)total_len + 16));
	k_sam_write<<<(n * 32 + 255) / 256, 256, 0, st>>>(t, d_recs, n, d_off, c->b[33].as<char>());
	CUDA_TRY(cudaEventRecord(c->ev1, st));
	CUDA_TRY(cudaMemcpyAsync(out, c->b[33].p, (size_t)total_len, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx_wait(c));
	CUDA_TRY(cudaGetLastError());
	float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
	c->last_ms = ms; c->last_launches = 4;
	*out_len = total_len;
	return EMAB_OK;
}
