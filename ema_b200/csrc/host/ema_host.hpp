// Host side of `ema align` above the device pipeline: the session (index + options), the barcode
// batcher, cloud construction, best-alignment choice, duplicate marking and SAM text.
// Mirrors include/align.h, include/techs.h, include/samrecord.h of the reference; see ema_host.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../../include/ema_b200.h"

namespace emab {

enum BcKind { BC_HAPLOTAG, BC_10X, BC_TELLSEQ, BC_TRUSEQ, BC_CPTSEQ };

struct Platform {  // PlatformProfile (include/techs.h:10-22, src/techs.c:71-127)
	const char *name;
	BcKind kind;
	int many_clouds;
	uint32_t bc_len;
	uint32_t dist_thresh;
	double error_rate;
	size_t n_density_probs;
	double density_probs[16];
};

const Platform *platform_by_name(const char *name);

struct Session {  // the reference's process globals (src/main.c:23-34, src/align.c:177-178) as one object
	emab_index_t *ix = nullptr;
	emab_ctx_t *ctx = nullptr;
	const Platform *tech = nullptr;
	int bc_len = 16;
	bool is_haplotag = false;
	std::string rg = "@RG\tID:rg1\tSM:sample1";  // src/main.c:25
	bool has_rg = true;
	std::string bx_index = "1";
	int apply_opt = 0;
	int n_threads = 1;
	std::vector<std::string> fai_names;   // read_fai (src/main.c:57-71)
	std::vector<int> rid2chrom;           // chrom_index(contig name) (src/main.c:41-55)
	std::vector<std::string> sq_names;    // @SQ from the .ann
	std::vector<int32_t> sq_len;
	int cloud_id = 0;                     // init_cloud's static counter (src/align.c:19-23)
	std::string err;
	std::string gamma_dump;               // test hook: path for full-precision posteriors of the chosen alignments
	emab_run_stats_t last{};
};

int session_open(const char *ref_path, const char *platform, int device, Session **out, std::string *err);
void session_close(Session *s);
void sam_header(const Session *s, int argc, const char *const *argv, std::string *out);
// find_clouds_and_align (src/align.c:214) over the *contents* of the input file(s); appends SAM text to out
int align_special_fastq(Session *s, const char *data, size_t len, std::string *out);
int align_fastq(Session *s, const char *d1, size_t l1, const char *d2, size_t l2, std::string *out);  // d2 == nullptr: interleaved

}  // namespace emab
