/*
 * oracle/ref_harness_ema.c — TEST INFRASTRUCTURE ONLY.
 *
 * Second half of the reference harness: textually includes the reference's src/align.c (never
 * copied into this repo) to reach its static functions (append_alignments, src/align.c:986) and
 * interposes find_best_record (src/samdict.c:166) so that the per-read EM posteriors can be
 * dumped at full double precision (the SAM only carries %.5g).  Built into oracle/_ref/libemaref.so.
 */
#include <stdio.h>
struct sam_dict_ent;
static FILE *h_dump = 0;

#define find_best_record(e) h_find_best_record(e)
#include "src/align.c" /* the reference's own source, via -I$(REF) */
#undef find_best_record

extern SAMRecord *find_best_record(SAMDictEnt *e);

/* called in place of find_best_record by find_clouds_and_align (src/align.c:551-552) */
SAMRecord *h_find_best_record(SAMDictEnt *e)
{
	SAMRecord *best = find_best_record(e);
	if (h_dump) {
		size_t i;
		fprintf(h_dump, "%s\t%d\t%zu", e->key->ident, (int)e->key->mate, e->num_cands);
		for (i = 0; i < e->num_cands; ++i)
			fprintf(h_dump, "\t%u:%u:%d:%d:%.17g", (unsigned)e->cand_records[i]->chrom, (unsigned)e->cand_records[i]->pos,
			        (int)e->cand_records[i]->rev, (int)e->cand_records[i]->active, e->gammas[i]);
		fprintf(h_dump, "\t|\t%u:%u:%.17g:%d:%d\n", (unsigned)best->chrom, (unsigned)best->pos, best->gamma, best->cloud->id, (int)best->cloud->bad);
	}
	return best;
}

static int h_inited = 0;

/* mirrors the set-up half of main() for `align` (src/main.c:316-369) */
int ref_ema_init(const char *ref_path, const char *platform)
{
	char fai[4096];
	FILE *f;
	if (h_inited) return 0;
	if ((tech = get_platform_profile_by_name(platform)) == NULL) return -1;
	BC_LEN = tech->bc_len;
	snprintf(fai, sizeof fai, "%s.fai", ref_path);
	if (!(f = fopen(fai, "r"))) return -2;
	read_fai(f);
	fclose(f);
	bwa_init(ref_path);
	arena_init();
	h_inited = 1;
	return 0;
}

/* find_clouds_and_align over one special-FASTQ bucket, with the gamma dump switched on */
int ref_ema_run_bucket(const char *bucket_path, const char *sam_path, const char *dump_path, int apply_opt, int n_threads)
{
	FILE *in = fopen(bucket_path, "r"), *out = fopen(sam_path, "w");
	if (!in || !out) return -1;
	h_dump = dump_path ? fopen(dump_path, "w") : 0;
	num_threads_per_file = n_threads;
	find_clouds_and_align(NULL, NULL, in, out, apply_opt, NULL, NULL);
	fclose(in); fclose(out);
	if (h_dump) { fclose(h_dump); h_dump = 0; }
	arena_init(); /* find_clouds_and_align destroys the arena on exit */
	return 0;
}

/* bwa_mem_mate_sw (src/bwabridge.c:204) for one pair of ASCII reads; regs flattened like ref_harness.c */
#define HREG_N 18
static void h_flat(const mem_alnreg_t *p, int64_t *o)
{
	union { float f; uint32_t u; } fr; fr.f = p->frac_rep;
	o[0] = p->rb; o[1] = p->re; o[2] = p->qb; o[3] = p->qe; o[4] = p->rid; o[5] = p->score; o[6] = p->truesc;
	o[7] = p->sub; o[8] = p->csub; o[9] = p->sub_n; o[10] = p->w; o[11] = p->seedcov; o[12] = p->secondary;
	o[13] = p->seedlen0; o[14] = p->n_comp; o[15] = p->is_alt; o[16] = fr.u; o[17] = p->secondary_all;
}

int ref_ema_pair(char *r1, int l1, char *r2, int l2, int64_t *regs1, int *n1, int64_t *regs2, int *n2, int max)
{
	EasyAlignmentPairs p = bwa_mem_mate_sw(ref, opts, r1, l1, r2, l2, 25);
	size_t i;
	*n1 = (int)p.len1; *n2 = (int)p.len2;
	for (i = 0; i < p.len1 && (int)i < max; ++i) h_flat(p.a1[i].chained_hit, regs1 + i * HREG_N);
	for (i = 0; i < p.len2 && (int)i < max; ++i) h_flat(p.a2[i].chained_hit, regs2 + i * HREG_N);
	arena_clear();
	return 0;
}

/* append_alignments (src/align.c:986) for one pair: the candidate SAMRecords in reference order.
 * ints per record: chrom,pos,rev,mate,mapq,score_mapq,clip,clip_edit_dist,NM,n_cigar,unique ; score separately */
#define HCAND_N 11
int ref_ema_candidates(const char *id, const char *read1, const char *qual1, const char *read2, const char *qual2,
                       int64_t *ints, double *scores, uint32_t *cigars, int max_cigar, int max)
{
	static FASTQRecord m1, m2;
	static SAMRecord *recs = 0;
	static size_t cap = 0;
	size_t n = 0, i;
	int k;
	if (!recs) { cap = 4096; recs = safe_malloc(cap * sizeof(*recs)); }
	memset(&m1, 0, sizeof m1); memset(&m2, 0, sizeof m2);
	m1.bc = m2.bc = 1;
	snprintf(m1.id, sizeof m1.id, "@%s", id); strcpy(m2.id, m1.id);
	strcpy(m1.read, read1); strcpy(m1.qual, qual1); m1.rlen = strlen(read1);
	strcpy(m2.read, read2); strcpy(m2.qual, qual2); m2.rlen = strlen(read2);
	append_alignments(ref, opts, &m1, &m2, &recs, &n, &cap);
	for (i = 0; i < n && (int)i < max; ++i) {
		SAMRecord *s = &recs[i];
		int64_t *o = ints + i * HCAND_N;
		o[0] = s->chrom; o[1] = s->pos; o[2] = s->rev; o[3] = s->mate; o[4] = s->mapq; o[5] = s->score_mapq;
		o[6] = s->clip; o[7] = s->clip_edit_dist; o[8] = s->aln.edit_dist; o[9] = s->aln.n_cigar; o[10] = s->unique;
		scores[i] = s->score;
		for (k = 0; k < s->aln.n_cigar && k < max_cigar; ++k) cigars[i * max_cigar + k] = s->aln.cigar[k];
	}
	arena_clear();
	return (int)n;
}
