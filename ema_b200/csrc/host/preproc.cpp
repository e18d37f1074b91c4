// `ema count` and `ema preproc`: the producers of the bucket files `align` reads.
//   emab_count    = count()   (cpp/count.cc:38-182, cpp/count.h):   barcode census of an interleaved FASTQ ->
//                   <prefix>.ema-ncnt (whitelisted barcodes seen, with counts) and <prefix>.ema-fcnt (every observed
//                   barcode + quality string, with counts)
//   emab_preproc  = correct() (cpp/correct.cc:66-633, cpp/correct.h): whitelist priors from the census, Hamming-1 (-h:
//                   Hamming-2) correction of every observed barcode string by posterior, barcodes dealt to the emptiest of
//                   n bucket files, then the FASTQ rewritten line by line into ema-bin-NNN / ema-nobc
// Host code: the work is a hash probe per read and file writes, and the OUTPUT ORDER is part of the format — the census
// is written, and the barcodes are dealt to buckets, in the iteration order of the reference's
// std::unordered_map<uint32_t, .> (cpp/count.cc:151-165, cpp/correct.cc:406-411).  The same container filled by the same
// sequence of insertions reproduces it; everything order-independent here is laid out differently from the reference
// (one pass over a block-buffered input, packed 16-byte observation keys, a flat sorted census instead of a node map).
// 10x barcodes (16 bases + 7 trimmed bases in front of read 1) and, with is_haplotag, haplotag barcodes (a BX:Z:AxxCxxBxxDxx
// tag on the name line, every one of the 96^4 combinations whitelisted, no correction, reads untrimmed).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <queue>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>
#include <sys/stat.h>
#include "../../../include/ema_b200.h"

extern thread_local char emab_errbuf[512];

namespace {

constexpr int BCL = 16;          // cpp/common.h:60-67
constexpr int TRIM = 7;
constexpr int QOFF = 33;
constexpr int QBASE = QOFF + 1;
constexpr size_t MIN_READ = 32;

inline int base_code(unsigned char c)   // hash_dna: A/C/G/T in either case -> 0..3, anything else 0 (cpp/common.h:83-97)
{
	switch (c | 0x20) { case 'c': return 1; case 'g': return 2; case 't': return 3; default: return 0; }
}
inline int base_code_n(unsigned char c)  // hash_dna_n: N/n -> 4 (cpp/common.h:99-113)
{
	return (c | 0x20) == 'n' ? 4 : base_code(c);
}

using ObsKey = std::array<unsigned char, BCL>;   // per position: base (0..4) * 34 + capped quality

int fail(int rc, const std::string &msg) { snprintf(emab_errbuf, sizeof emab_errbuf, "%s", msg.c_str()); return rc; }

// getline over a FILE, block-buffered; a final line without '\n' counts (std::getline's behaviour)
struct Lines {
	FILE *f;
	std::vector<char> buf;
	size_t pos = 0, end = 0;
	bool eof = false;
	explicit Lines(FILE *f_) : f(f_), buf(1 << 22) {}
	bool next(std::string &out)
	{
		out.clear();
		bool any = false;
		for (;;) {
			if (pos == end) {
				if (eof) return any;
				end = fread(buf.data(), 1, buf.size(), f);
				pos = 0;
				if (end == 0) { eof = true; return any; }
			}
			const char *p = buf.data() + pos;
			const char *nl = (const char *)memchr(p, '\n', end - pos);
			if (nl) { out.append(p, nl - p); pos = (size_t)(nl - buf.data()) + 1; return true; }
			out.append(p, end - pos);
			any = true;
			pos = end;
		}
	}
};

// the 16 barcode bases of read 1 with their qualities -> observation key + 2-bit code; false = the read is ignored
// (a quality character below '!': cpp/count.cc:114-128, cpp/correct.cc:452-466)
bool observe(const std::string &r, std::string &q, ObsKey &key, uint32_t &code, bool &has_n)
{
	code = 0; has_n = false;
	for (int k = 0; k < BCL; ++k) {
		if (q[k] < QOFF) {
			fprintf(stderr, "Ignoring long read--- quality score %s less than %d\n", q.c_str(), QOFF);
			return false;
		}
		if (q[k] - QOFF >= QBASE) q[k] = (char)(QOFF + QBASE - 1);
		key[k] = (unsigned char)(base_code_n((unsigned char)r[k]) * QBASE + std::min(QBASE - 1, q[k] - QOFF));
		code = code << 2 | (uint32_t)base_code((unsigned char)r[k]);
		has_n |= r[k] == 'N';
	}
	return true;
}

// haplotag: "BX:Z:AxxCxxBxxDxx" after the first blank of the name line -> (A << 24 | C << 16 | B << 8 | D), two decimal
// digits each taken as they are (cpp/common.h:69-75); `bound` is the length the reference compares the tag's end with
// (cpp/count.cc:93, cpp/correct.cc:441: the name line in count, a stale line in correct)
bool haplotag_code(const std::string &name, size_t bound, uint32_t &code, std::string &tag)
{
	size_t at = name.find_first_of(" \t");
	if (at == std::string::npos) return false;
	at = name.find("BX:Z:", at);
	if (at == std::string::npos || !(at + 16 < bound)) return false;
	tag = name.substr(at + 5, 12);
	auto two = [&](size_t k) { return 10 * ((k < tag.size() ? tag[k] : 0) - '0') + ((k + 1 < tag.size() ? tag[k + 1] : 0) - '0'); };
	code = (uint32_t)two(1) << 24 | (uint32_t)two(4) << 16 | (uint32_t)two(7) << 8 | (uint32_t)two(10);
	return true;
}
// all 96^4 haplotag barcodes in the order the reference inserts them (cpp/common.h:76-77: a, b, c, d nested, key a|c|b|d)
template <class Map, class Init>
void all_haplotags(Map &m, Init init)
{
	for (uint32_t a = 1; a <= 96; ++a) for (uint32_t b = 1; b <= 96; ++b) for (uint32_t c = 1; c <= 96; ++c) for (uint32_t d = 1; d <= 96; ++d)
		init(m[a << 24 | c << 16 | b << 8 | d]);
}

int load_whitelist(const char *path, std::vector<uint32_t> &codes)
{
	FILE *f = fopen(path, "r");
	if (!f) return fail(EMAB_ERR_IO, std::string("Cannot open file ") + path);
	Lines in(f);
	std::string s;
	while (in.next(s)) {
		uint32_t code = 0;
		for (int k = 0; k < BCL; ++k) code = code << 2 | (uint32_t)base_code(k < (int)s.size() ? (unsigned char)s[k] : 0);
		if (code == 0) { fclose(f); return fail(EMAB_ERR_ARG, "Invalid barcode AAA...AA whitelisted"); }
		codes.push_back(code);
	}
	fclose(f);
	return EMAB_OK;
}

bool put(FILE *f, const void *p, size_t n) { return fwrite(p, 1, n, f) == n; }

// one block of the full census: count, then (key, count) in key order (cpp/count.cc:17-34)
bool write_block(FILE *f, std::vector<std::pair<ObsKey, int64_t>> &block)
{
	std::sort(block.begin(), block.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
	const int64_t n = (int64_t)block.size();
	if (!put(f, &n, 8)) return false;
	for (const auto &e : block) if (!put(f, e.first.data(), BCL) || !put(f, &e.second, 8)) return false;
	fflush(f);
	block.clear();
	return true;
}

struct KeyHash {
	size_t operator()(const ObsKey &k) const { uint64_t a, b; memcpy(&a, k.data(), 8); memcpy(&b, k.data() + 8, 8); return (size_t)((a * 0x9e3779b97f4a7c15ull) ^ (b + (a >> 29))); }
};

}  // namespace

extern "C" int emab_count(const char *whitelist_path, const char *output_prefix, uint64_t max_map_bytes, int is_haplotag, FILE *in_stream)
{
	if ((!whitelist_path && !is_haplotag) || !output_prefix) return fail(EMAB_ERR_ARG, "count: whitelist and output prefix are required");
	// the reference's container, filled in the reference's order: the census file is written in its iteration order
	std::unordered_map<uint32_t, int64_t> seen;
	if (is_haplotag) all_haplotags(seen, [](int64_t &v) { v = 0; });
	else {
		std::vector<uint32_t> wl;
		if (int rc = load_whitelist(whitelist_path, wl)) return rc;
		for (uint32_t c : wl) seen[c] = 0;
	}
	const std::string pre(output_prefix);
	FILE *ff = is_haplotag ? nullptr : fopen((pre + ".ema-fcnt").c_str(), "wb");   // haplotag: no full census (nothing is corrected)
	if (!ff && !is_haplotag) return fail(EMAB_ERR_IO, "Cannot open file " + pre + ".ema-fcnt");
	FILE *fn = fopen((pre + ".ema-ncnt").c_str(), "wb");
	if (!fn) { if (ff) fclose(ff); return fail(EMAB_ERR_IO, "Cannot open file " + pre + ".ema-ncnt"); }
	// the full census of the current block: the reference dumps its std::map when (sizeof(key) + sizeof(count) + 32) * size
	// reaches max_map_size (cpp/common.h:120-125, cpp/count.cc:140-143), i.e. at a fixed number of distinct keys
	const size_t dump_at = (size_t)((max_map_bytes + 71) / 72);
	std::unordered_map<ObsKey, int64_t, KeyHash> census;
	std::vector<std::pair<ObsKey, int64_t>> block;
	auto dump = [&]() {
		block.assign(census.begin(), census.end());
		census.clear();
		return write_block(ff, block);
	};
	Lines in(in_stream ? in_stream : stdin);
	std::string name, r, q, skip, tag;
	ObsKey key;
	int64_t total = 0, nice = 0, ignored = 0;
	bool ok = true;
	while (ok && in.next(name)) {
		in.next(r); in.next(q); in.next(q);
		uint32_t code = 0;
		bool has_n = false;
		bool process = (!is_haplotag || haplotag_code(name, name.size(), code, tag)) && r.size() >= MIN_READ;
		if (process && !is_haplotag) process = observe(r, q, key, code, has_n);
		if (process) {
			if (!has_n) {
				auto it = seen.find(code);
				if (it != seen.end()) { ++it->second; ++nice; }
			}
			if (!is_haplotag) {
				const bool fresh = ++census[key] == 1;
				if (fresh && census.size() >= dump_at) ok = dump();
			}
			++total;
		} else ++ignored;
		for (int k = 0; k < 4; ++k) in.next(skip);
	}
	int64_t n_seen = 0;
	for (const auto &e : seen) n_seen += e.second != 0;
	ok = ok && put(fn, &n_seen, 8);
	for (const auto &e : seen)
		if (e.second) ok = ok && put(fn, &e.first, 4) && put(fn, &e.second, 8);
	fclose(fn);
	if (ff) { ok = ok && dump(); fclose(ff); }
	fprintf(stderr, ":: Reads with OK barcode: %lld out of %lld\n:: Ignored %lld reads\n", (long long)nice, (long long)total, (long long)ignored);
	return ok ? EMAB_OK : fail(EMAB_ERR_IO, "fwrite failed");
}

namespace {

struct Known { int64_t n_reads = 0; double prior = 0; int bucket = 0; };
using KnownMap = std::unordered_map<uint32_t, Known>;

struct Observed { ObsKey key; uint32_t fixed; int64_t n; };

constexpr double CONF = 0.975;   // BC_CONF_THRESH (cpp/correct.cc:24)

// The whitelisted barcode an observation most likely came from (cpp/correct.cc:66-167): exact hit, else the Hamming-1
// neighbours (or the four completions of a single N), each weighted prior x P(error at that base quality); with -h and
// an exact hit also the Hamming-2 neighbours.  0 = no confident assignment.  kind: 0 unchanged, 1 H1, 2 H2, 3 none.
uint32_t assign(const ObsKey &q, const KnownMap &known, const double *perr, bool do_h2, int *kind)
{
	uint32_t code = 0;
	int ns = 0;
	for (int k = 0; k < BCL; ++k) {
		const int b = q[k] / QBASE;
		code = code << 2 | (uint32_t)(b == 4 ? 0 : b);
		ns += b == 4;
	}
	*kind = 3;
	if (ns > 1) return 0;
	auto exact = ns == 0 ? known.find(code) : known.end();
	uint32_t best = 0;
	double best_p = -1, total = 0;
	auto shift_of = [](int k) { return (BCL - k - 1) * 2; };
	if (exact != known.end()) {
		best_p = exact->second.prior; best = code; total += best_p; *kind = 0;
		if (do_h2) {
			for (int i1 = 0; i1 < BCL; ++i1) for (int j1 = 0; j1 < 4; ++j1) {
				if (j1 == q[i1] / QBASE) continue;
				for (int i2 = i1 + 1; i2 < BCL; ++i2) for (int j2 = 0; j2 < 4; ++j2) {
					if (j2 == q[i2] / QBASE) continue;
					const uint32_t alt = (code & ~(3u << shift_of(i1)) & ~(3u << shift_of(i2))) | (uint32_t)j1 << shift_of(i1) | (uint32_t)j2 << shift_of(i2);
					auto it = known.find(alt);
					if (it == known.end()) continue;
					const double p1 = perr[(int)std::max(3.0, q[i1] % QBASE - 1.0)], p2 = perr[(int)std::max(3.0, q[i2] % QBASE - 1.0)];
					const double p = it->second.prior * (p1 * p2);
					total += p;
					if (p > best_p) { best_p = p; best = alt; *kind = 2; }
				}
			}
		}
	} else {
		for (int i = 0; i < BCL; ++i) {
			if (ns && q[i] / QBASE != 4) continue;
			for (int j = 0; j < 4; ++j) {
				if (ns == 0 && j == q[i] / QBASE) continue;
				const uint32_t alt = (code & ~(3u << shift_of(i))) | (uint32_t)j << shift_of(i);
				auto it = known.find(alt);
				if (it == known.end()) continue;
				const double p = it->second.prior * perr[q[i] % QBASE];
				total += p;
				if (p > best_p) { best_p = p; best = alt; *kind = 1; }
			}
		}
	}
	if (best_p / total > CONF) return best;
	*kind = 3;
	return 0;
}

}  // namespace

extern "C" int emab_preproc(const char *whitelist_path, const char *const *count_files, int n_count_files, const char *output_dir,
                            int do_h2, uint64_t buffer_size, int do_bx_format, int n_threads, int n_buckets, int is_haplotag, FILE *in_stream)
{
	if ((!whitelist_path && !is_haplotag) || !output_dir || n_buckets < 1 || n_count_files < 0) return fail(EMAB_ERR_ARG, "preproc: bad arguments");
	if (n_threads < 1) n_threads = 1;
	double perr[128];
	for (int i = 0; i < 128; ++i) perr[i] = pow(10.0, -std::min(QBASE - 1, i) / 10.0);
	// ---- whitelist and priors (cpp/correct.cc:291-336); the same container and insertion order as the reference: the
	// barcodes are dealt to the bucket files in its iteration order
	KnownMap known;
	if (is_haplotag) all_haplotags(known, [](Known &k) { k.prior = 0; });
	else {
		std::vector<uint32_t> wl;
		if (int rc = load_whitelist(whitelist_path, wl)) return rc;
		for (uint32_t c : wl) known[c].prior = 0;
	}
	std::vector<std::string> full_paths;
	for (int i = 0; i < n_count_files; ++i) {
		std::string s(count_files[i]);
		struct stat sb;
		if (stat(s.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) return fail(EMAB_ERR_IO, s + " is not a file");
		if (s.size() < 9 || s.compare(s.size() - 9, 9, ".ema-ncnt") != 0) return fail(EMAB_ERR_ARG, s + " is not an ema-ncnt file");
		if (is_haplotag) continue;   // no full census to correct from
		std::string f = s;
		f[f.size() - 4] = 'f';
		if (stat(f.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) return fail(EMAB_ERR_IO, f + " is not a file");
		full_paths.push_back(f);
	}
	for (int i = 0; i < n_count_files; ++i) {
		FILE *f = fopen(count_files[i], "rb");
		if (!f) return fail(EMAB_ERR_IO, std::string("Cannot open file ") + count_files[i]);
		int64_t n = 0;
		bool ok = fread(&n, 8, 1, f) == 1;
		while (ok && n-- > 0) {
			uint32_t code; int64_t cnt;
			ok = fread(&code, 4, 1, f) == 1 && fread(&cnt, 8, 1, f) == 1;
			if (ok) { if (is_haplotag) known[code].n_reads += cnt; else known[code].prior += (double)cnt; }
		}
		fclose(f);
		if (!ok) return fail(EMAB_ERR_IO, "fread failed (corrupted input?)");
	}
	if (!is_haplotag) {
		double sum = 0;
		for (const auto &e : known) sum += e.second.prior + 1;
		for (auto &e : known) e.second.prior = (e.second.prior + 1) / sum;
	}
	// ---- every observed barcode string gets its whitelisted barcode (or none); reads per whitelisted barcode
	std::unordered_map<ObsKey, uint32_t, KeyHash> fixed;   // observations that change: key -> corrected code
	int64_t stats[4] = {0, 0, 0, 0};
	for (const std::string &path : full_paths) {
		FILE *f = fopen(path.c_str(), "rb");
		if (!f) return fail(EMAB_ERR_IO, "Cannot open file " + path);
		int64_t n = 0;
		while (fread(&n, 8, 1, f) == 1) {
			std::vector<Observed> obs((size_t)n);
			for (auto &o : obs) {
				if (fread(o.key.data(), 1, BCL, f) != (size_t)BCL || fread(&o.n, 8, 1, f) != 1) { fclose(f); return fail(EMAB_ERR_IO, "fread failed (corrupted input?)"); }
				o.fixed = 0;
			}
			std::vector<std::array<int64_t, 4>> part((size_t)n_threads, std::array<int64_t, 4>{0, 0, 0, 0});
			std::vector<std::thread> team;
			const size_t per = (obs.size() + (size_t)n_threads - 1) / (size_t)n_threads;
			std::vector<int> kinds(obs.size());
			for (int t = 0; t < n_threads; ++t)
				team.emplace_back([&, t]() {
					const size_t a = std::min(obs.size(), (size_t)t * per), b = std::min(obs.size(), a + per);
					for (size_t i = a; i < b; ++i) {
						obs[i].fixed = assign(obs[i].key, known, perr, do_h2 != 0, &kinds[i]);
						part[(size_t)t][kinds[i]] += obs[i].n;
					}
				});
			for (auto &th : team) th.join();
			for (const auto &p : part) for (int k = 0; k < 4; ++k) stats[k] += p[k];
			for (size_t i = 0; i < obs.size(); ++i) {
				if (!obs[i].fixed) continue;
				if (kinds[i] == 1 || kinds[i] == 2) fixed[obs[i].key] = obs[i].fixed;
				known[obs[i].fixed].n_reads += obs[i].n;
			}
		}
		fclose(f);
	}
	if (!is_haplotag) fprintf(stderr, ":: Stats: no change: %lld \n         no barcode: %lld \n       H1-corrected: %lld \n       H2-corrected: %lld \n",
	        (long long)stats[0], (long long)stats[3], (long long)stats[1], (long long)stats[2]);
	// ---- output files; barcodes to the emptiest bucket so far, ties to the lower index (cpp/correct.cc:372-411)
	{
		struct stat sb;
		const int have = stat(output_dir, &sb);
		if (have == 0 && S_ISREG(sb.st_mode)) return fail(EMAB_ERR_IO, std::string(output_dir) + " exists but is not a directory");
		if (have != 0 && mkdir(output_dir, S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH) == -1) return fail(EMAB_ERR_IO, std::string("Cannot create directory ") + output_dir);
	}
	struct Out { FILE *f = nullptr; int64_t load = 0; std::string buf; };
	std::vector<Out> outs((size_t)n_buckets + 1);
	auto close_all = [&]() { for (auto &o : outs) if (o.f) { fclose(o.f); o.f = nullptr; } };
	for (int i = 0; i <= n_buckets; ++i) {
		char nm[64];
		if (i == 0) snprintf(nm, sizeof nm, "ema-nobc"); else snprintf(nm, sizeof nm, "ema-bin-%03d", i - 1);
		outs[(size_t)i].f = fopen((std::string(output_dir) + "/" + nm).c_str(), "wb");
		if (!outs[(size_t)i].f) { close_all(); return fail(EMAB_ERR_IO, std::string("Cannot open file ") + output_dir + "/" + nm); }
		outs[(size_t)i].buf.reserve((size_t)buffer_size + 10 * 1024);
	}
	{
		using Slot = std::pair<int64_t, int>;   // (reads so far, file index): the smallest first
		std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> emptiest;
		for (int i = 1; i <= n_buckets; ++i) emptiest.push({0, i});
		for (auto &e : known) {
			Slot s = emptiest.top(); emptiest.pop();
			s.first += e.second.n_reads;
			e.second.bucket = s.second;
			emptiest.push(s);
		}
	}
	// ---- the FASTQ, pair by pair (cpp/correct.cc:413-620)
	Lines in(in_stream ? in_stream : stdin);
	std::string name, r, q, l, tag;
	ObsKey key;
	char bc_text[BCL + 1];
	bc_text[BCL] = 0;
	bool ok = true;
	size_t stale_len = 0;   // haplotag: the reference bounds the tag by the length of the LAST line it read into another
	                        // variable (the previous pair's last line; empty before the first pair): cpp/correct.cc:441
	auto word = [](const std::string &s) { size_t k = 0; while (k < s.size() && !isspace((unsigned char)s[k])) ++k; return std::string_view(s.data(), k); };
	while (ok && in.next(name)) {
		in.next(r); in.next(q); in.next(q);
		bool process = r.size() >= MIN_READ;
		uint32_t code = 0;
		bool has_n = false;
		if (is_haplotag) process = haplotag_code(name, stale_len, code, tag) && process;
		else if (process) process = observe(r, q, key, code, has_n);
		if (!process) { for (int k = 0; k < 4; ++k) in.next(l); stale_len = l.size(); continue; }
		if (!is_haplotag) {
			auto fx = fixed.find(key);
			if (fx != fixed.end()) { code = fx->second; has_n = false; }
		}
		int fidx = 0;
		auto kn = has_n ? known.end() : known.find(code);
		if (kn != known.end()) fidx = kn->second.bucket; else code = 0;
		std::string &o = outs[(size_t)fidx].buf;
		auto put_bc = [&]() {
			if (!code) return;
			if (is_haplotag) { o.append(tag, 0, 12); o.resize(o.size() + (12 - std::min<size_t>(12, tag.size())), '\0'); return; }
			uint32_t c = code;
			for (int k = 0; k < BCL; ++k) { bc_text[BCL - k - 1] = "ACGT"[c & 3]; c >>= 2; }
			o.append(bc_text, BCL);
		};
		const bool flat = fidx && !do_bx_format;   // one line per pair: "BARCODE name read qual mate-read mate-qual"
		if (flat) { put_bc(); o.push_back(' '); }
		o.append(word(name));
		if (fidx) {
			o.push_back(' ');
			if (do_bx_format) { o.append("BX:Z:"); put_bc(); o.append(is_haplotag ? "\n" : "-1\n"); }
		} else o.push_back('\n');
		const size_t cut = is_haplotag ? 0 : BCL + TRIM, keep = r.size() - cut;   // haplotag reads are not trimmed
		o.append(r, cut, keep);
		if (flat) o.push_back(' '); else o.append("\n+\n");
		{   // the quality is copied for its own length but the cursor moves by the READ's (cpp/correct.cc:556-557)
			const size_t at = o.size();
			o.append(q, std::min(cut, q.size()), std::string::npos);
			o.resize(at + keep, '\0');
		}
		o.push_back(flat ? ' ' : '\n');
		in.next(l);
		if (!flat) {
			o.append(word(l));
			if (do_bx_format) { o.append(" BX:Z:"); put_bc(); if (!is_haplotag) o.append("-1"); }
			o.push_back('\n');
		}
		in.next(l);
		o.append(l);
		if (flat) o.push_back(' '); else o.append("\n+\n");
		in.next(l); in.next(l);
		stale_len = l.size();
		o.append(l);
		o.push_back('\n');
		if (o.size() >= buffer_size) { ok = put(outs[(size_t)fidx].f, o.data(), o.size()); o.clear(); }
	}
	for (auto &x : outs) { ok = ok && put(x.f, x.buf.data(), x.buf.size()); }
	close_all();
	return ok ? EMAB_OK : fail(EMAB_ERR_IO, "fwrite failed");
}
