// ema-b200: the `ema align` command line (src/main.c:243-425) over libema_b200.so.
// Same option grammar (getopt "r:1:2:s:xo:R:dp:i:t:"), same messages, same SAM bytes; `count` and
// `preproc` are out of scope and are reported as such.
#include <getopt.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>
#include "../../../include/ema_b200.h"

// -R takes the read group with its tabs written as the two characters `\t` (src/util.c:22-39 expands \t \n \r \\ and
// drops the backslash of any other pair).  A lone trailing backslash is dropped too.
static std::string unescape(const char *in)
{
	std::string o;
	for (const char *p = in; *p; ++p) {
		if (*p != '\\') { o.push_back(*p); continue; }
		const char c = p[1];
		if (c == 0) break;
		++p;
		switch (c) {
		case 't': o.push_back('\t'); break;
		case 'n': o.push_back('\n'); break;
		case 'r': o.push_back('\r'); break;
		case '\\': o.push_back('\\'); break;
		default: break;
		}
	}
	return o;
}

static bool slurp(const char *path, std::string *out)
{
	FILE *f = fopen(path, "r");
	if (!f) return false;
	char buf[1 << 16];
	size_t n;
	while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->append(buf, n);
	fclose(f);
	return true;
}

static void usage(const char *argv0, int error)
{
	FILE *out = error ? stderr : stdout;
	fprintf(out, "usage: %s <align|index|help> [options]\n\n", argv0);
	fprintf(out, "align: choose best alignments based on barcodes\n");
	fprintf(out, "  -1 <FASTQ1 path>: first (preprocessed and sorted) FASTQ file [none]\n");
	fprintf(out, "  -2 <FASTQ2 path>: second (preprocessed and sorted) FASTQ file [none]\n");
	fprintf(out, "  -s <EMA-FASTQ path>: specify special FASTQ path [none]\n");
	fprintf(out, "  -x: multi-input mode; takes input files after flags [off]\n");
	fprintf(out, "  -r <FASTA path>: indexed reference [required]\n");
	fprintf(out, "  -o <SAM file>: output SAM file [stdout]\n");
	fprintf(out, "  -R <RG string>: full read group string (e.g. '@RG\\tID:foo\\tSM:bar') [none]\n");
	fprintf(out, "  -d: apply fragment read density optimization [off]\n");
	fprintf(out, "  -p <platform>: sequencing platform (one of '10x', 'tru', 'haplotag', 'dbs', 'cpt', 'tellseq') [10x]\n");
	fprintf(out, "  -i <index>: index to follow 'BX' tag in SAM output [1]\n");
	fprintf(out, "  -t <threads>: set number of host threads [1]\n");
	fprintf(out, "  all other arguments (only for -x): list of all preprocessed inputs\n\n");
	fprintf(out, "index [-p <prefix>] <FASTA path>: build the BWA index `align -r` loads (= `bwa index`), on the GPU\n\n");
	fprintf(out, "count / preproc: not part of this build (use the reference's `ema preproc` to make buckets)\n");
	exit(error ? EXIT_FAILURE : EXIT_SUCCESS);
}

#define IOERROR(fn) do { fprintf(stderr, "error: file %s could not be opened\n", (fn)); exit(EXIT_FAILURE); } while (0)

int main(int argc, char *argv[])
{
	const char *argv0 = argv[0];
	if (argc < 2) {
		fprintf(stderr, "EMA version 0.6.2 (ema-b200)\nnote: use '%s help' for usage information.\n", argv0);
		return EXIT_SUCCESS;
	}
	const char *mode = argv[1];
	if (strcmp(mode, "help") == 0) usage(argv0, 0);
	if (strcmp(mode, "index") == 0) {
		// `ema-b200 index [-p prefix] <ref.fa>`: what `bwa index` does for `ema align -r`, on the GPU (same five files, same bytes)
		const char *prefix = NULL;
		int device = 0, c;
		if (const char *d = getenv("EMAB_DEVICE")) device = atoi(d);
		while ((c = getopt(argc - 1, &argv[1], "p:")) != -1) {
			if (c == 'p') prefix = optarg;
			else usage(argv0, 1);
		}
		if (optind + 1 >= argc) { fprintf(stderr, "error: specify the reference FASTA\n"); exit(EXIT_FAILURE); }
		emab_index_build_stats_t st;
		if (emab_index_build(argv[optind + 1], prefix, device, &st)) { fprintf(stderr, "%s\n", emab_last_error()); exit(EXIT_FAILURE); }
		fprintf(stderr, "[index] %lld bp in %d sequences: pack %.1f s, suffix sort %.1f s (%d chunks, %lld tied suffixes, %d rounds), occ %.1f s, write %.1f s\n",
		        (long long)st.l_pac, st.n_seqs, st.ms_pack / 1e3, st.ms_sort / 1e3, st.n_chunks, (long long)st.n_tied, st.max_rounds, st.ms_occ / 1e3, st.ms_write / 1e3);
		return EXIT_SUCCESS;
	}
	if (strcmp(mode, "align") != 0) {
		fprintf(stderr, "error: unrecognized mode\n");
		usage(argv0, 1);
	}
	char *ref = NULL, *fq1 = NULL, *fq2 = NULL, *fqx = NULL, *out = NULL, *rg = NULL, *bx = NULL;
	const char *platform = "10x";
	int apply_opt = 0, multi = 0, t = 1, device = 0;
	int c;
	if (const char *d = getenv("EMAB_DEVICE")) device = atoi(d);
	// EMAB_DEVICES=0,1,...: the GPUs -x mode spreads its buckets over (work stealing); the first one serves -s / -1
	std::vector<int> devices;
	if (const char *d = getenv("EMAB_DEVICES")) {
		for (const char *p = d; *p;) { devices.push_back(atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
		if (!devices.empty()) device = devices[0];
	}
	while ((c = getopt(argc - 1, &argv[1], "r:1:2:s:xo:R:dp:i:t:")) != -1) {
		switch (c) {
		case 'r': ref = strdup(optarg); break;
		case '1': fq1 = strdup(optarg); break;
		case '2': fq2 = strdup(optarg); break;
		case 's': fqx = strdup(optarg); break;
		case 'x': multi = 1; break;
		case 'o': out = strdup(optarg); break;
		case 'R': rg = strdup(unescape(optarg).c_str()); break;
		case 'd': apply_opt = 1; break;
		case 'p': platform = strdup(optarg); break;
		case 'i': bx = strdup(optarg); break;
		case 't': t = atoi(optarg); break;
		default: usage(argv0, 1);
		}
	}
	if (multi + (fqx != NULL) + (fq1 != NULL || fq2 != NULL) != 1) {
		fprintf(stderr, "error: must specify *exactly one* of -1/-2, -s or -x\n");
		exit(EXIT_FAILURE);
	}
	if (fq1 == NULL && fq2 != NULL) { fprintf(stderr, "error: cannot specify -2 without -1\n"); exit(EXIT_FAILURE); }
	if (ref == NULL) { fprintf(stderr, "error: specify reference FASTA with -r\n"); exit(EXIT_FAILURE); }
	FILE *out_file = out == NULL ? stdout : fopen(out, "w");
	if (!out_file) IOERROR(out);
	setenv("EMAB_MALLOC_TUNE", "1", 0);   // keep the per-bucket work buffers in the heap arenas (see tune_malloc in ema_host.cpp)
	fprintf(stderr, "BWA initialization...\n");
	emab_session_t *s = NULL;
	if (emab_session_open(ref, platform, device, &s)) { fprintf(stderr, "%s\n", emab_last_error()); exit(EXIT_FAILURE); }
	if (emab_session_config(s, rg, bx, apply_opt, t)) { fprintf(stderr, "%s\n", emab_last_error()); exit(EXIT_FAILURE); }
	char *text = NULL;
	uint64_t len = 0;
	emab_sam_header(s, argc, argv, &text, &len);
	fwrite(text, 1, len, out_file);
	emab_free(text);
	auto emit = [&](int rc) {
		if (rc) { fprintf(stderr, "%s\n", emab_last_error()); exit(EXIT_FAILURE); }
		fwrite(text, 1, len, out_file);
		emab_free(text);
	};
	if (multi) {
		const int n_inputs = argc - optind - 1;
		if (n_inputs == 0) { fprintf(stderr, "warning: no input files specified; nothing to do\n"); exit(EXIT_SUCCESS); }
		// buckets are emitted in argument order; EMAB_WORKERS of them (default 3) are in flight on the GPU
		for (size_t d = 1; d < devices.size(); ++d)
			if (emab_session_add_device(s, devices[d])) { fprintf(stderr, "%s\n", emab_last_error()); exit(EXIT_FAILURE); }
		int workers = 3 * (int)std::max<size_t>(1, devices.size());
		if (const char *w = getenv("EMAB_WORKERS")) workers = atoi(w);
		emab_session_workers(s, workers);
		// a window of 2 x workers buckets is resident at a time: read, align, write and release before the next
		// window is read, so host memory is bounded by the window whatever the number of inputs
		const int window = std::max(2, 2 * workers);
		for (int base = 0; base < n_inputs; base += window) {
			const int n = std::min(window, n_inputs - base);
			std::vector<std::string> datas(n);
			std::vector<const char *> ptrs(n);
			std::vector<uint64_t> lens(n), olens(n);
			std::vector<char *> outs(n);
			for (int i = 0; i < n; ++i) {
				if (!slurp(argv[optind + 1 + base + i], &datas[i])) IOERROR(argv[optind + 1 + base + i]);
				ptrs[i] = datas[i].data(); lens[i] = datas[i].size();
				fprintf(stderr, "Processing reads...\n");
			}
			if (emab_align_buckets(s, n, ptrs.data(), lens.data(), outs.data(), olens.data())) { fprintf(stderr, "%s\n", emab_last_error()); exit(EXIT_FAILURE); }
			for (int i = 0; i < n; ++i) { fwrite(outs[i], 1, olens[i], out_file); emab_free(outs[i]); }
		}
	} else if (fqx) {
		std::string data;
		if (!slurp(fqx, &data)) IOERROR(fqx);
		fprintf(stderr, "Processing reads...\n");
		emit(emab_align_bucket(s, data.data(), data.size(), &text, &len));
	} else {
		// -1 / -2: streamed — barcode groups are cut into device batches, memory stays bounded by the batches in flight
		FILE *f1 = fopen(fq1, "r");
		if (!f1) IOERROR(fq1);
		FILE *f2 = fq2 ? fopen(fq2, "r") : NULL;
		if (fq2 && !f2) IOERROR(fq2);
		fprintf(stderr, "Processing reads...\n");
		int workers = 3;
		if (const char *w = getenv("EMAB_WORKERS")) workers = atoi(w);
		emab_session_workers(s, workers);
		int batch = 0;
		if (const char *b = getenv("EMAB_FASTQ_BATCH")) batch = atoi(b);
		auto rd = [](void *u, char *buf, int64_t cap) -> int64_t { FILE *f = (FILE *)u; size_t n = fread(buf, 1, (size_t)cap, f); return ferror(f) ? -1 : (int64_t)n; };
		auto wr = [](void *u, const char *t, uint64_t n) -> int { return fwrite(t, 1, (size_t)n, (FILE *)u) == n ? 0 : -1; };
		if (emab_align_fastq_stream(s, rd, f1, f2 ? (emab_read_cb)rd : NULL, f2, wr, out_file, batch)) { fprintf(stderr, "%s\n", emab_last_error()); exit(EXIT_FAILURE); }
		fclose(f1);
		if (f2) fclose(f2);
	}
	if (out_file != stdout) fclose(out_file);
	emab_session_close(s);
	return EXIT_SUCCESS;
}
