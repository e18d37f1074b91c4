// FASTA -> .pac/.ann/.amb, the host half of index construction (see indexbuild.hpp).
// Restates bns_fasta2bntseq / add1 / bns_dump (bwa/bntseq.c:64-96,213-330) and the part of kseq_read a FASTA
// reaches (bwa/kseq.h:176-204) over a memory image of the file instead of a 16 KB stream buffer.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "indexbuild.hpp"
#include "../../../include/ema_b200.h"

namespace emab {

static const uint8_t *nt4_of()
{  // nst_nt4_table (bwa/bntseq.c:45-62): ACGT in either case -> 0..3, everything else >= 4
	static uint8_t t[256];
	static bool ready = false;
	if (!ready) {
		memset(t, 4, sizeof t);
		t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; t['-'] = 5;
		ready = true;
	}
	return t;
}

int pack_fasta(const char *path, PackedRef *out, std::string *err)
{
	FILE *f = fopen(path, "rb");
	if (!f) { *err = std::string("cannot open ") + path; return EMAB_ERR_IO; }
	fseek(f, 0, SEEK_END);
	const long long size = ftell(f);
	fseek(f, 0, SEEK_SET);
	std::vector<char> text((size_t)size + 1);
	if (size && fread(text.data(), 1, (size_t)size, f) != (size_t)size) { fclose(f); *err = std::string("short read on ") + path; return EMAB_ERR_IO; }
	fclose(f);
	text[(size_t)size] = 0;
	const char *p = text.data(), *end = p + size;
	const uint8_t *nt4 = nt4_of();
	struct drand48_data rng;
	srand48_r(11, &rng);   // bns->seed = 11 (bwa/bntseq.c:262-263)
	out->pac.assign((size_t)(size / 4 + 2), 0);
	out->l_pac = 0;
	out->contigs.clear(); out->holes.clear();
	int64_t l = 0;
	// the first header may start anywhere (kseq skips to the first '>' or '@'); later ones only at a line start
	while (p < end && *p != '>' && *p != '@') ++p;
	while (p < end) {
		++p;  // the header character
		RefContig c;
		const char *q = p;
		while (q < end && !isspace((unsigned char)*q)) ++q;
		c.name.assign(p, q);
		c.anno = "(null)";
		if (q < end && *q != '\n') {  // the rest of the line is the comment
			const char *r = q + 1;
			const char *nl = (const char *)memchr(r, '\n', (size_t)(end - r));
			const char *e = nl ? nl : end;
			size_t n = (size_t)(e - r);
			if (n > 1 && r[n - 1] == '\r') --n;
			if (n) c.anno.assign(r, n);
			q = e;
		}
		p = q < end ? q + 1 : end;
		c.offset = l; c.n_ambs = 0;
		const int64_t l0 = l;
		int lasts = 0;
		while (p < end) {  // sequence lines up to the next header
			const char first = *p;
			if (first == '>' || first == '@') break;
			if (first == '+') { *err = "FASTQ input is not supported by the index builder"; return EMAB_ERR_ARG; }
			if (first == '\n') { ++p; continue; }
			const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
			const char *e = nl ? nl : end;
			size_t n = (size_t)(e - p);
			if ((l - l0) + (int64_t)n > 1 && p[n - 1] == '\r') --n;   // ks_getuntil2's CR rule
			if ((size_t)((l + (int64_t)n) / 4 + 2) > out->pac.size()) out->pac.resize((size_t)((l + (int64_t)n) / 4 + 2) * 2, 0);
			uint8_t *pac = out->pac.data();
			for (size_t i = 0; i < n; ++i) {
				const int ch = (unsigned char)p[i];
				int code = nt4[ch];
				if (code >= 4) {  // an ambiguity code: a new hole unless it continues a run of the same character
					if (lasts == ch) ++out->holes.back().len;
					else { out->holes.push_back(RefHole{l, 1, (char)ch}); ++c.n_ambs; }
					long r;
					lrand48_r(&rng, &r);
					code = (int)(r & 3);
				}
				lasts = ch;
				pac[l >> 2] |= (uint8_t)(code << ((~l & 3) << 1));
				++l;
			}
			p = nl ? nl + 1 : end;
		}
		if (l - l0 > 0x7fffffff) { *err = "contig longer than 2^31-1"; return EMAB_ERR_ARG; }
		c.len = (int32_t)(l - l0);
		out->contigs.push_back(std::move(c));
	}
	out->l_pac = l;
	out->pac.resize((size_t)(l / 4 + 2));
	if (out->contigs.empty() || l == 0) { *err = std::string("no sequence in ") + path; return EMAB_ERR_ARG; }
	return EMAB_OK;
}

int write_pac_ann_amb(const PackedRef &ref, const char *prefix, std::string *err)
{
	const std::string pre(prefix);
	FILE *f = fopen((pre + ".pac").c_str(), "wb");
	if (!f) { *err = "cannot write " + pre + ".pac"; return EMAB_ERR_IO; }
	const int64_t l = ref.l_pac;
	fwrite(ref.pac.data(), 1, (size_t)((l >> 2) + ((l & 3) == 0 ? 0 : 1)), f);
	uint8_t ct = 0;
	if (l % 4 == 0) fwrite(&ct, 1, 1, f);   // the file is always l_pac/4 + 2 bytes (bwa/bntseq.c:318-324)
	ct = (uint8_t)(l % 4);
	fwrite(&ct, 1, 1, f);
	if (fclose(f)) { *err = "write error on " + pre + ".pac"; return EMAB_ERR_IO; }
	f = fopen((pre + ".ann").c_str(), "w");
	if (!f) { *err = "cannot write " + pre + ".ann"; return EMAB_ERR_IO; }
	fprintf(f, "%lld %d %u\n", (long long)l, (int)ref.contigs.size(), 11u);
	for (const RefContig &c : ref.contigs) {
		fprintf(f, "%d %s", 0, c.name.c_str());
		if (!c.anno.empty()) fprintf(f, " %s\n", c.anno.c_str());
		else fprintf(f, "\n");
		fprintf(f, "%lld %d %d\n", (long long)c.offset, c.len, c.n_ambs);
	}
	if (fclose(f)) { *err = "write error on " + pre + ".ann"; return EMAB_ERR_IO; }
	f = fopen((pre + ".amb").c_str(), "w");
	if (!f) { *err = "cannot write " + pre + ".amb"; return EMAB_ERR_IO; }
	fprintf(f, "%lld %d %u\n", (long long)l, (int)ref.contigs.size(), (unsigned)ref.holes.size());
	for (const RefHole &h : ref.holes) fprintf(f, "%lld %d %c\n", (long long)h.offset, h.len, h.amb);
	if (fclose(f)) { *err = "write error on " + pre + ".amb"; return EMAB_ERR_IO; }
	return EMAB_OK;
}

}  // namespace emab
