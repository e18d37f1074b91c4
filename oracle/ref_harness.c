/*
 * oracle/ref_harness.c — TEST INFRASTRUCTURE ONLY.
 *
 * A ctypes-friendly shim around the UNMODIFIED reference BWA-MEM code.  It textually includes
 * /root/reference/bwa/bwamem.c (never copied into this repo) so that the static functions on the
 * hot path (mem_collect_intv, bwa/bwamem.c:140) can be called stage by stage, and flattens the
 * reference's structs into int64 arrays.  Built by oracle/Makefile into oracle/_ref/libemaref.so.
 * Nothing in the product path links or loads this file.
 */
#include "bwamem.c" /* the reference's own source, via -I$(REF)/bwa */

#define HREG_N 18

extern int mem_matesw(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], const mem_alnreg_t *a, int l_ms, const uint8_t *ms, mem_alnreg_v *ma);

static mem_opt_t *h_opts(void)
{
	static mem_opt_t *o = 0;
	if (!o) { o = mem_opt_init(); o->max_occ = 3000; /* src/align.c:184-185 */ }
	return o;
}

void *ref_idx_load(const char *prefix) { return bwa_idx_load(prefix, BWA_IDX_ALL); }
void ref_idx_destroy(void *idx) { bwa_idx_destroy((bwaidx_t*)idx); }

/* out[0]=l_pac out[1]=n_seqs out[2]=primary out[3]=seq_len out[4..8]=L2[0..4] out[9]=sa_intv out[10]=n_sa out[11]=bwt_size */
void ref_idx_info(void *idx_, int64_t *out)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	int i;
	out[0] = idx->bns->l_pac; out[1] = idx->bns->n_seqs; out[2] = idx->bwt->primary; out[3] = idx->bwt->seq_len;
	for (i = 0; i < 5; ++i) out[4+i] = idx->bwt->L2[i];
	out[9] = idx->bwt->sa_intv; out[10] = idx->bwt->n_sa; out[11] = idx->bwt->bwt_size;
}

void ref_contig(void *idx_, int i, int64_t *offset, int32_t *len, char *name, int name_cap)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	*offset = idx->bns->anns[i].offset; *len = idx->bns->anns[i].len;
	strncpy(name, idx->bns->anns[i].name, name_cap - 1); name[name_cap-1] = 0;
}

/* raw FM-index primitives (bwa/bwt.c) */
void ref_occ4(void *idx_, int64_t k, int64_t cnt[4]) { bwt_occ4(((bwaidx_t*)idx_)->bwt, (bwtint_t)k, (bwtint_t*)cnt); }
int64_t ref_sa(void *idx_, int64_t k) { return (int64_t)bwt_sa(((bwaidx_t*)idx_)->bwt, (bwtint_t)k); }
void ref_sa_batch(void *idx_, int n, const int64_t *k, int64_t *out)
{
	int i;
	for (i = 0; i < n; ++i) out[i] = (int64_t)bwt_sa(((bwaidx_t*)idx_)->bwt, (bwtint_t)k[i]);
}
/* in/out: x0,x1,x2 ; out 4 intervals x 3 */
void ref_extend(void *idx_, const int64_t ik[3], int is_back, int64_t ok[12])
{
	bwtintv_t a, o[4]; int c;
	a.x[0] = ik[0]; a.x[1] = ik[1]; a.x[2] = ik[2]; a.info = 0;
	bwt_extend(((bwaidx_t*)idx_)->bwt, &a, o, is_back);
	for (c = 0; c < 4; ++c) { ok[c*3] = o[c].x[0]; ok[c*3+1] = o[c].x[1]; ok[c*3+2] = o[c].x[2]; }
}

/* bwt_smem1 (bwa/bwt.c:353): returns next x; intervals as x0,x1,x2,info */
int ref_smem1(void *idx_, int len, const uint8_t *q, int x, int min_intv, int64_t *out, int max, int *n_out)
{
	bwtintv_v mem = {0,0,0};
	int i, ret = bwt_smem1(((bwaidx_t*)idx_)->bwt, len, q, x, min_intv, &mem, 0);
	*n_out = (int)mem.n;
	for (i = 0; i < (int)mem.n && i < max; ++i) {
		out[i*4] = mem.a[i].x[0]; out[i*4+1] = mem.a[i].x[1]; out[i*4+2] = mem.a[i].x[2]; out[i*4+3] = mem.a[i].info;
	}
	free(mem.a);
	return ret;
}

/* mem_collect_intv (bwa/bwamem.c:140): sorted SA intervals of one read (nt4 input) */
int ref_collect_intv(void *idx_, int len, const uint8_t *seq, int64_t *out, int max)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	smem_aux_t *a = smem_aux_init();
	int i, n;
	mem_collect_intv(h_opts(), idx->bwt, len, seq, a);
	n = (int)a->mem.n;
	for (i = 0; i < n && i < max; ++i) {
		out[i*4] = a->mem.a[i].x[0]; out[i*4+1] = a->mem.a[i].x[1]; out[i*4+2] = a->mem.a[i].x[2]; out[i*4+3] = a->mem.a[i].info;
	}
	smem_aux_destroy(a);
	return n;
}

/* mem_chain (+ optionally mem_chain_flt).  chains: pos,rid,n,w,kept,first,frac_rep_bits,seed_off ; seeds: rbeg,qbeg,len,score */
int ref_chain(void *idx_, int len, const uint8_t *seq, int do_flt, int64_t *chains, int maxc, int64_t *seeds, int maxs, int *n_seeds_out)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	mem_chain_v chn = mem_chain(h_opts(), idx->bwt, idx->bns, len, seq, 0);
	int i, j, ns = 0;
	if (do_flt) chn.n = mem_chain_flt(h_opts(), chn.n, chn.a);
	for (i = 0; i < (int)chn.n; ++i) {
		mem_chain_t *c = &chn.a[i];
		union { float f; uint32_t u; } fr; fr.f = c->frac_rep;
		if (i < maxc) {
			chains[i*8] = c->pos; chains[i*8+1] = c->rid; chains[i*8+2] = c->n; chains[i*8+3] = do_flt? c->w : (int64_t)mem_chain_weight(c);
			chains[i*8+4] = do_flt? c->kept : 0; chains[i*8+5] = do_flt? c->first : -1; chains[i*8+6] = fr.u; chains[i*8+7] = ns;
		}
		for (j = 0; j < c->n; ++j, ++ns)
			if (ns < maxs) { seeds[ns*4] = c->seeds[j].rbeg; seeds[ns*4+1] = c->seeds[j].qbeg; seeds[ns*4+2] = c->seeds[j].len; seeds[ns*4+3] = c->seeds[j].score; }
		free(c->seeds);
	}
	free(chn.a);
	*n_seeds_out = ns;
	return (int)chn.n;
}

static void h_flatten_reg(const mem_alnreg_t *p, int64_t *o)
{
	union { float f; uint32_t u; } fr; fr.f = p->frac_rep;
	o[0] = p->rb; o[1] = p->re; o[2] = p->qb; o[3] = p->qe; o[4] = p->rid; o[5] = p->score; o[6] = p->truesc;
	o[7] = p->sub; o[8] = p->csub; o[9] = p->sub_n; o[10] = p->w; o[11] = p->seedcov; o[12] = p->secondary;
	o[13] = p->seedlen0; o[14] = p->n_comp; o[15] = p->is_alt; o[16] = fr.u; o[17] = p->secondary_all;
}

static void h_unflatten_reg(const int64_t *o, mem_alnreg_t *p)
{
	union { float f; uint32_t u; } fr;
	memset(p, 0, sizeof(*p));
	p->rb = o[0]; p->re = o[1]; p->qb = o[2]; p->qe = o[3]; p->rid = o[4]; p->score = o[5]; p->truesc = o[6];
	p->sub = o[7]; p->csub = o[8]; p->sub_n = o[9]; p->w = o[10]; p->seedcov = o[11]; p->secondary = o[12];
	p->seedlen0 = o[13]; p->n_comp = o[14]; p->is_alt = o[15]; fr.u = (uint32_t)o[16]; p->frac_rep = fr.f; p->secondary_all = o[17];
}

/* mem_align1_core (bwa/bwamem.c:1081) on an nt4 read; regs flattened HREG_N int64 each */
int ref_align1(void *idx_, int len, const uint8_t *seq, int64_t *regs, int max)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	char *s = (char*)malloc(len);
	mem_alnreg_v r;
	int i;
	memcpy(s, seq, len);
	r = mem_align1_core(h_opts(), idx->bwt, idx->bns, idx->pac, len, s, 0);
	for (i = 0; i < (int)r.n && i < max; ++i) h_flatten_reg(&r.a[i], regs + (size_t)i * HREG_N);
	free(r.a); free(s);
	return (int)r.n;
}

/* mem_chain2aln only (no dedup): regs in creation order */
int ref_chain2aln(void *idx_, int len, const uint8_t *seq, int64_t *regs, int max)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	mem_chain_v chn = mem_chain(h_opts(), idx->bwt, idx->bns, len, seq, 0);
	mem_alnreg_v av;
	int i;
	chn.n = mem_chain_flt(h_opts(), chn.n, chn.a);
	mem_flt_chained_seeds(h_opts(), idx->bns, idx->pac, len, seq, chn.n, chn.a);
	kv_init(av);
	for (i = 0; i < (int)chn.n; ++i) {
		mem_chain2aln(h_opts(), idx->bns, idx->pac, len, seq, &chn.a[i], &av);
		free(chn.a[i].seeds);
	}
	free(chn.a);
	for (i = 0; i < (int)av.n && i < max; ++i) h_flatten_reg(&av.a[i], regs + (size_t)i * HREG_N);
	free(av.a);
	return (int)av.n;
}

/* mem_matesw (bwa/bwamem_pair.c:137) with EMA's fixed insert model (src/bwabridge.c:216-229) */
int ref_matesw(void *idx_, const int64_t *anchor, int l_ms, const uint8_t *ms, int64_t *regs, int n_regs, int max, int *n_out)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	mem_pestat_t pes[4];
	mem_alnreg_t a;
	mem_alnreg_v ma;
	int i, n;
	for (i = 0; i < 4; ++i) { pes[i].failed = (i != 1); pes[i].low = -35; pes[i].high = 500; pes[i].avg = 200.0; pes[i].std = 100.0; }
	h_unflatten_reg(anchor, &a);
	kv_init(ma);
	for (i = 0; i < n_regs; ++i) { mem_alnreg_t b; h_unflatten_reg(regs + (size_t)i * HREG_N, &b); kv_push(mem_alnreg_t, ma, b); }
	n = mem_matesw(h_opts(), idx->bns, idx->pac, pes, &a, l_ms, ms, &ma);
	for (i = 0; i < (int)ma.n && i < max; ++i) h_flatten_reg(&ma.a[i], regs + (size_t)i * HREG_N);
	*n_out = (int)ma.n;
	free(ma.a);
	return n;
}

/* mem_reg2aln (bwa/bwamem.c:1119). seq is nt4. out: pos,rid,is_rev,NM,score,sub,n_cigar,mapq,flag */
int ref_reg2aln(void *idx_, int len, const uint8_t *seq, const int64_t *reg, int64_t *out, uint32_t *cigar, int max_cigar)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	mem_alnreg_t r;
	mem_aln_t a;
	int i;
	h_unflatten_reg(reg, &r);
	a = mem_reg2aln(h_opts(), idx->bns, idx->pac, len, (const char*)seq, &r);
	out[0] = a.pos; out[1] = a.rid; out[2] = a.is_rev; out[3] = a.NM; out[4] = a.score; out[5] = a.sub; out[6] = a.n_cigar; out[7] = a.mapq; out[8] = a.flag;
	for (i = 0; i < a.n_cigar && i < max_cigar; ++i) cigar[i] = a.cigar[i];
	free(a.cigar);
	return a.n_cigar;
}

/* bwa_gen_cigar2 score-only path used by mem_patch_reg (bwa/bwamem.c:454) */
int ref_gen_cigar_score(void *idx_, int w, int l_query, const uint8_t *query, int64_t rb, int64_t re)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	mem_opt_t *o = h_opts();
	int score = -(1<<30);
	uint8_t *q = (uint8_t*)malloc(l_query);
	memcpy(q, query, l_query);
	bwa_gen_cigar2(o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, w, idx->bns->l_pac, idx->pac, l_query, q, rb, re, &score, 0, 0);
	free(q);
	return score;
}

/* reference fetch (bwa/bntseq.c:426): returns length, clamps beg/end in place */
int ref_fetch_seq(void *idx_, int64_t *beg, int64_t mid, int64_t *end, int *rid, uint8_t *out, int max)
{
	bwaidx_t *idx = (bwaidx_t*)idx_;
	uint8_t *s = bns_fetch_seq(idx->bns, idx->pac, beg, mid, end, rid);
	int n = (int)(*end - *beg);
	memcpy(out, s, n < max? n : max);
	free(s);
	return n;
}

/* batched ksw_extend2 with the call-site parameters of bwa/bwamem.c:757,785.  Sequences are
 * concatenated; offsets are int64.  out: score,qle,tle,gtle,gscore,max_off per task.  Returns
 * the number of DP cells the reference loop visits (sum of end-beg at bwa/ksw.c:460), counted by
 * re-deriving beg/end is not possible from outside, so callers that need cell counts use the
 * port (oracle_ksw.c), which is verified equal to this function. */
void ref_extend_batch(int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                      const int32_t *h0, int w, int end_bonus, int zdrop, int32_t *out, int n_threads)
{
	mem_opt_t *o = h_opts();
	int i;
	#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64)
	for (i = 0; i < n; ++i) {
		int qle, tle, gtle, gscore, max_off, sc;
		sc = ksw_extend2((int)(qoff[i+1] - qoff[i]), q + qoff[i], (int)(toff[i+1] - toff[i]), t + toff[i], 5, o->mat,
		                 o->o_del, o->e_del, o->o_ins, o->e_ins, w, end_bonus, zdrop, h0[i], &qle, &tle, &gtle, &gscore, &max_off);
		out[i*6] = sc; out[i*6+1] = qle; out[i*6+2] = tle; out[i*6+3] = gtle; out[i*6+4] = gscore; out[i*6+5] = max_off;
	}
}

/* batched ksw_global2; cigar_out has max_cigar u32 per task; out: score,n_cigar */
void ref_global_batch(int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                      const int32_t *w, int32_t *out, uint32_t *cigar_out, int max_cigar, int n_threads)
{
	mem_opt_t *o = h_opts();
	int i;
	#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64)
	for (i = 0; i < n; ++i) {
		int n_cigar = 0, k; uint32_t *cigar = 0;
		int sc = ksw_global2((int)(qoff[i+1] - qoff[i]), q + qoff[i], (int)(toff[i+1] - toff[i]), t + toff[i], 5, o->mat,
		                     o->o_del, o->e_del, o->o_ins, o->e_ins, w[i], &n_cigar, &cigar);
		out[i*2] = sc; out[i*2+1] = n_cigar;
		for (k = 0; k < n_cigar && k < max_cigar; ++k) cigar_out[(size_t)i * max_cigar + k] = cigar[k];
		free(cigar);
	}
}

/* batched ksw_align2 as mem_matesw calls it (bwa/bwamem_pair.c:176-177). out: score,te,qe,score2,te2,tb,qb */
void ref_local_batch(int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                     int32_t *out, int n_threads)
{
	mem_opt_t *o = h_opts();
	int i;
	#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64)
	for (i = 0; i < n; ++i) {
		int ql = (int)(qoff[i+1] - qoff[i]), tl = (int)(toff[i+1] - toff[i]);
		int xtra = KSW_XSUBO | KSW_XSTART | (ql * o->a < 250? KSW_XBYTE : 0) | (o->min_seed_len * o->a);
		uint8_t *qq = (uint8_t*)malloc(ql), *tt = (uint8_t*)malloc(tl);
		kswr_t r;
		memcpy(qq, q + qoff[i], ql); memcpy(tt, t + toff[i], tl);
		r = ksw_align2(ql, qq, tl, tt, 5, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, xtra, 0);
		out[i*7] = r.score; out[i*7+1] = r.te; out[i*7+2] = r.qe; out[i*7+3] = r.score2; out[i*7+4] = r.te2; out[i*7+5] = r.tb; out[i*7+6] = r.qb;
		free(qq); free(tt);
	}
}
