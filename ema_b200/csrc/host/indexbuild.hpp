// Index construction (the `bwa index` step `ema align` depends on), host half: FASTA -> packed forward strand +
// the .ann/.amb/.pac files exactly as bns_fasta2bntseq writes them (bwa/bntseq.c:248-330, bns_dump :64-96).
// The device half (indexbuild.cu) sorts the suffixes of forward+reverse-complement on the GPU and writes .bwt/.sa.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace emab {

struct RefContig { std::string name, anno; int64_t offset; int32_t len, n_ambs; };   // bntann1_t
struct RefHole { int64_t offset; int32_t len; char amb; };                             // bntamb1_t

struct PackedRef {
	std::vector<uint8_t> pac;        // forward strand, 2 bits/base, base l in byte l>>2 at bits ((~l&3)<<1)
	int64_t l_pac = 0;
	std::vector<RefContig> contigs;
	std::vector<RefHole> holes;
};

// Reads a FASTA file the way kseq_read + add1 do (names to the first space, the rest of the header line as the
// comment, non-ACGT bases replaced by lrand48()&3 of the generator seeded with 11, runs of one ambiguity character
// recorded as holes).  Returns 0 or EMAB_ERR_*; *err explains.
int pack_fasta(const char *path, PackedRef *out, std::string *err);
// <prefix>.pac/.ann/.amb, byte for byte what `bwa index` leaves (the forward-only second pass, bwa/bwtindex.c:303-310)
int write_pac_ann_amb(const PackedRef &ref, const char *prefix, std::string *err);

}  // namespace emab
