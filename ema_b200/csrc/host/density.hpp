// -d: the fragment read-density optimiser of `ema align` (reference: mark_optimal_alignments_in_cloud,
// src/split.c:38-338, constants include/split.h:8-17).  Included by ema_host.cpp after Rec / is_pair / Platform.
//
// What it does: inside one "bad" cloud (a read has several alignments there) choose one alignment per read so that the
// reads spread over the cloud's 1 kb bins the way the platform's density profile expects, by simulated annealing;
// the alignments not chosen get active = 0 and are ignored by the EM's best pick (src/samdict.c:173-176).
//
// How it is built here.  A cloud's records arrive sorted by (read name, mate) (name_cmp, src/align.c:71-82), so the
// alignments of one read are adjacent and so are the two mates of a pair.  The optimiser works on READ SLOTS — one
// per read: its surviving alignments, the one currently chosen, the slot of its mate if the mate is in the cloud —
// and a histogram of the chosen positions.  The reference keeps two arrays (uniquely and multiply mapped reads) with
// cross indices found by linear searches; on name-sorted input those searches can only ever find the neighbouring
// slot, which is what `mate` stores directly.
//
// What must be the reference's exactly is the use of the random stream: one process-wide libc rand() sequence seeded
// from time(), consumed in (bucket, barcode, cloud) order: per proposal one draw for the read, one for the target
// alignment, and one more for the Metropolis test only when the move is neither forced nor an improvement
// (src/split.c:229-306).  GlibcRand reproduces glibc's rand() (TYPE_3 additive feedback, stdlib/random_r.c) without
// its lock; emab_host_selftest checks it against the C library's own rand() on this machine.
#pragma once
#include <cmath>
#include <cstdint>
#include <ctime>
#include <vector>

namespace emab {

struct GlibcRand {
	uint32_t r[31];
	int f = 3, b = 0;
	void seed(unsigned s)
	{
		if (s == 0) s = 1;
		int32_t w = (int32_t)s;
		r[0] = (uint32_t)w;
		for (int i = 1; i < 31; ++i) {  // 16807 * w mod (2^31 - 1) without overflow
			const int32_t hi = w / 127773, lo = w % 127773;
			w = 16807 * lo - 2836 * hi;
			if (w < 0) w += 2147483647;
			r[i] = (uint32_t)w;
		}
		f = 3; b = 0;
		for (int i = 0; i < 310; ++i) next();
	}
	int next()
	{
		const uint32_t v = (r[f] += r[b]);
		if (++f == 31) f = 0;
		if (++b == 31) b = 0;
		return (int)(v >> 1);
	}
};

class DensityOptimiser {
public:
	static constexpr int kBin = 1000, kMaxBins = 1000, kScoreScale = 20, kMaxRejected = 500, kProposals = 50000, kExtraDepth = 5;
	static constexpr size_t kMaxRecords = 50000;

	explicit DensityOptimiser(const Platform *tech) : tech_(tech)
	{
		for (int d = 0; d < kLogTable; ++d) log_density_[d] = density_log(d);
		temperature_.resize(kProposals);
		for (int k = 0; k < kProposals; ++k) temperature_[k] = pow(10.0, 0.0 - ((0.0 - (-12.0)) * k) / kProposals);   // TMAX_LOG 0, TMIN_LOG -12
	}

	// `cloud`: indices into `recs`, sorted by (name, mate).  Only the records' `active` flags change.
	void run(std::vector<Rec> &recs, const std::vector<int> &cloud, GlibcRand &rng)
	{
		const size_t n_in = cloud.size();
		if (n_in >= kMaxRecords || n_in <= 5) return;
		build_slots(recs, cloud);
		if (order_.size() <= 5 || movable_.empty()) return;
		if ((size_t)(hi_ - lo_) / kBin + 1 >= (size_t)kMaxBins) return;
		// from here on the cloud's verdict is ours: everything off, the chosen ones back on at the end
		for (int i : order_) recs[i].active = 0;
		bins_.assign(kMaxBins, 0);
		for (const Slot &s : slots_) ++bins_[bin_of(recs[order_[s.first + s.chosen]].pos)];
		anneal(recs, rng);
		for (const Slot &s : slots_) recs[order_[s.first + s.chosen]].active = 1;
	}

private:
	struct Slot { int first, count, chosen, mate; };   // alignments order_[first .. first+count), mate = slot index or -1
	struct Move {
		int slot, to;            // the read and the alignment it would switch to
		int mate_slot, mate_to;  // the mate is dragged along to keep the pair proper (mate_to < 0: it stays)
		bool forced;             // the move turns an improper pair into a proper one: always taken
		double gain;             // change of the configuration's log probability
	};
	static constexpr int kLogTable = 64;

	const Platform *tech_;
	double log_density_[kLogTable];
	std::vector<double> temperature_;
	std::vector<int> order_;           // the cloud's records after pruning
	std::vector<Slot> slots_;
	std::vector<int> movable_;         // slots with more than one alignment, in cloud order
	std::vector<unsigned short> bins_;
	uint32_t lo_ = 0, hi_ = 0;

	double density_log(unsigned density) const
	{  // log_density_prob (src/split.c:15-35): the profile's entries, then halving per extra read
		const size_t size = tech_->n_density_probs;
		if (density < size) return log(tech_->density_probs[density]);
		return log(tech_->density_probs[size - 1]) - (density - size + 1) * log(2.0);
	}
	double ld(unsigned short density) const { return density < kLogTable ? log_density_[density] : density_log(density); }
	size_t bin_of(uint32_t pos) const { return (size_t)((pos - lo_) / kBin); }

	// src/split.c:84-196: per read keep the alignments within kExtraDepth of its best clip+edit distance (the rest are
	// switched off for good), start from its best-scoring one, link mates
	void build_slots(std::vector<Rec> &recs, const std::vector<int> &cloud)
	{
		order_.clear(); slots_.clear(); movable_.clear();
		lo_ = 0xffffffffu; hi_ = 0;
		const size_t n = cloud.size();
		auto same_read = [&](int a, int b) { return recs[a].mate == recs[b].mate && recs[a].ident == recs[b].ident; };
		for (size_t i = 0; i < n;) {
			size_t j = i + 1;
			while (j < n && same_read(cloud[j], cloud[i])) ++j;
			Slot s{(int)order_.size(), 0, 0, -1};
			if (j - i > 1) {
				int floor = recs[cloud[i]].clip_edit_dist;
				for (size_t k = i + 1; k < j; ++k) floor = std::min(floor, recs[cloud[k]].clip_edit_dist);
				for (size_t k = i; k < j; ++k) {
					if (recs[cloud[k]].clip_edit_dist <= floor + kExtraDepth) order_.push_back(cloud[k]);
					else recs[cloud[k]].active = 0;
				}
			} else order_.push_back(cloud[i]);
			s.count = (int)order_.size() - s.first;
			for (int k = 1; k < s.count; ++k)   // first of the highest scores
				if (recs[order_[s.first + k]].score > recs[order_[s.first + s.chosen]].score) s.chosen = k;
			for (int k = 0; k < s.count; ++k) {
				const uint32_t p = recs[order_[s.first + k]].pos;
				lo_ = std::min(lo_, p); hi_ = std::max(hi_, p);
			}
			if (!slots_.empty()) {  // the other mate of the same pair, if present, is the slot just before
				const Rec &prev = recs[order_[slots_.back().first]], &cur = recs[order_[s.first]];
				if (prev.ident == cur.ident && prev.mate != cur.mate) { s.mate = (int)slots_.size() - 1; slots_.back().mate = (int)slots_.size(); }
			}
			if (s.count > 1) movable_.push_back((int)slots_.size());
			slots_.push_back(s);
			i = j;
		}
	}

	const Rec &chosen(const std::vector<Rec> &recs, const Slot &s) const { return recs[order_[s.first + s.chosen]]; }

	// what switching `m.slot` to alignment `m.to` would change (src/split.c:236-300)
	void evaluate(const std::vector<Rec> &recs, Move &m) const
	{
		const Slot &s = slots_[m.slot];
		const Rec &from = chosen(recs, s), &to = recs[order_[s.first + m.to]];
		m.mate_slot = s.mate; m.mate_to = -1; m.forced = false;
		double score_gain = 0.0;
		size_t mate_from_bin = 0, mate_to_bin = 0;
		if (s.mate >= 0) {
			const Slot &ms = slots_[s.mate];
			const Rec &mate_now = chosen(recs, ms);
			const bool was_pair = is_pair(from, mate_now), will_pair = is_pair(to, mate_now);
			if (!was_pair && will_pair) m.forced = true;
			else if (was_pair && !will_pair && ms.count > 1) {
				for (int k = 0; k < ms.count; ++k) {   // the first alignment of the mate that pairs with the target
					const Rec &cand = recs[order_[ms.first + k]];
					if (!is_pair(to, cand)) continue;
					m.mate_to = k;
					mate_from_bin = bin_of(mate_now.pos); mate_to_bin = bin_of(cand.pos);
					score_gain += (cand.score - mate_now.score) / kScoreScale;
					break;
				}
			}
		}
		const size_t from_bin = bin_of(from.pos), to_bin = bin_of(to.pos);
		const bool drag = m.mate_to >= 0;
		const int leave = drag && from_bin == mate_from_bin ? 2 : 1;   // both mates leave / enter the same bin
		const int enter = drag && to_bin == mate_to_bin ? 2 : 1;
		double density_gain = (ld(bins_[from_bin] - leave) - ld(bins_[from_bin])) + (ld(bins_[to_bin] + enter) - ld(bins_[to_bin]));
		if (leave == 1 && drag) density_gain += ld(bins_[mate_from_bin] - 1) - ld(bins_[mate_from_bin]);
		if (enter == 1 && drag) density_gain += ld(bins_[mate_to_bin] + 1) - ld(bins_[mate_to_bin]);
		score_gain += (to.score - from.score) / kScoreScale;
		m.gain = density_gain + score_gain;
	}

	void apply(const std::vector<Rec> &recs, const Move &m)
	{
		Slot &s = slots_[m.slot];
		--bins_[bin_of(chosen(recs, s).pos)];
		s.chosen = m.to;
		++bins_[bin_of(chosen(recs, s).pos)];
		if (m.mate_to >= 0) {
			Slot &ms = slots_[m.mate_slot];
			--bins_[bin_of(chosen(recs, ms).pos)];
			ms.chosen = m.mate_to;
			++bins_[bin_of(chosen(recs, ms).pos)];
		}
	}

	void anneal(const std::vector<Rec> &recs, GlibcRand &rng)
	{
		int rejected = 0;
		const int n_movable = (int)movable_.size();
		for (int k = 0; k < kProposals && rejected < kMaxRejected; ++k) {
			Move m;
			m.slot = movable_[rng.next() % n_movable];
			const Slot &s = slots_[m.slot];
			m.to = rng.next() % (s.count - 1);
			if (m.to >= s.chosen) ++m.to;   // any alignment but the current one
			evaluate(recs, m);
			if (m.forced || m.gain > 0 || exp(m.gain / temperature_[k]) >= ((double)rng.next()) / RAND_MAX) apply(recs, m);
			else ++rejected;
		}
	}
};

}  // namespace emab
