// Chaining and chain filtering: mem_chain (bwa/bwamem.c:277-341) with test_and_merge (:216-237),
// mem_chain_weight (:239-258) and mem_chain_flt (:353-411).  One thread per read.
//
// The reference keeps chains in a klib B-tree keyed by the reference position of each chain's first
// seed and asks it for "the chain with the largest key <= rbeg".  Keys can repeat (two chains
// starting at the same reference position), and with repeated keys both the answer of that query
// and the final in-order traversal depend on the tree's shape, so the tree itself is restated here
// (bwa/kbtree.h:118-233, t = 5 for the 40-byte mem_chain_t: at most 9 keys per node) over a
// per-read node pool.  Seeds of a chain form a linked list in the read's seed pool and are
// compacted per chain once filtering is done.
#pragma once
#include "seed.cuh"
#include "sort.cuh"

#define BT_T 5
#define BT_MAXKEYS (2 * BT_T - 1)

struct BNode {
	int32_t n, is_internal;
	int32_t key[BT_MAXKEYS];      // chain indices; ordered by chains[key].pos
	int32_t ptr[BT_MAXKEYS + 1];  // child node indices
};

struct ChainWork {   // per-read working pools (capacity `cap` = number of SA occurrences enumerated for the read)
	Seed *seeds;     // [cap]
	Chain *chains;   // [cap]
	BNode *nodes;    // [cap/3 + 2]
	int32_t *ord;    // [3*cap] scratch: traversal order / kept list, then (weight,index) pairs
};

struct WIdx { int32_t w, idx; };  // chain weight + index: what ks_introsort(mem_flt) compares and moves

// __kb_getp_aux (bwa/kbtree.h:118-133): lower-bound probe inside one node
EMAB_HD int bt_probe(const BNode &x, const Chain *chains, int64_t pos, int *r)
{
	if (x.n == 0) return -1;
	int begin = 0, end = x.n;
	while (begin < end) {
		int mid = (begin + end) >> 1;
		if (chains[x.key[mid]].pos < pos) begin = mid + 1;
		else end = mid;
	}
	if (begin == x.n) { *r = 1; return x.n - 1; }
	int64_t kp = chains[x.key[begin]].pos;
	*r = (kp < pos) - (pos < kp);
	if (*r < 0) --begin;
	return begin;
}

// kb_intervalp (bwa/kbtree.h:152-170): only `lower` is consumed by mem_chain
EMAB_HD int bt_lower(const BNode *nodes, int root, const Chain *chains, int64_t pos)
{
	int lower = -1, x = root;
	while (x >= 0) {
		int r = 0;
		const BNode &nd = nodes[x];
		int i = bt_probe(nd, chains, pos, &r);
		if (i >= 0 && r == 0) return nd.key[i];
		if (i >= 0) lower = nd.key[i];
		if (!nd.is_internal) return lower;
		x = nd.ptr[i + 1];
	}
	return lower;
}

// __kb_split (bwa/kbtree.h:177-193): y = child i of x is full
EMAB_HD void bt_split(BNode *nodes, int *n_nodes, int xi, int i, int yi)
{
	int zi = (*n_nodes)++;
	BNode &x = nodes[xi], &y = nodes[yi], &z = nodes[zi];
	z.is_internal = y.is_internal;
	z.n = BT_T - 1;
	for (int k = 0; k < BT_T - 1; ++k) z.key[k] = y.key[BT_T + k];
	if (y.is_internal) for (int k = 0; k < BT_T; ++k) z.ptr[k] = y.ptr[BT_T + k];
	y.n = BT_T - 1;
	for (int k = x.n; k > i; --k) x.ptr[k + 1] = x.ptr[k];
	x.ptr[i + 1] = zi;
	for (int k = x.n - 1; k >= i; --k) x.key[k + 1] = x.key[k];
	x.key[i] = y.key[BT_T - 1];
	++x.n;
}

// kb_putp + __kb_putp_aux (bwa/kbtree.h:194-229), iterative
EMAB_HD void bt_put(BNode *nodes, int *n_nodes, int *root, const Chain *chains, int ci)
{
	const int64_t pos = chains[ci].pos;
	int r = *root;
	if (nodes[r].n == BT_MAXKEYS) {
		int s = (*n_nodes)++;
		nodes[s].is_internal = 1; nodes[s].n = 0; nodes[s].ptr[0] = r;
		*root = s;
		bt_split(nodes, n_nodes, s, 0, r);
		r = s;
	}
	int x = r;
	for (;;) {
		BNode &nd = nodes[x];
		int dummy = 0;
		if (!nd.is_internal) {
			int i = bt_probe(nd, chains, pos, &dummy);
			for (int k = nd.n - 1; k > i; --k) nd.key[k + 1] = nd.key[k];
			nd.key[i + 1] = ci;
			++nd.n;
			return;
		}
		int i = bt_probe(nd, chains, pos, &dummy) + 1;
		if (nodes[nd.ptr[i]].n == BT_MAXKEYS) {
			bt_split(nodes, n_nodes, x, i, nd.ptr[i]);
			int64_t kp = chains[nodes[x].key[i]].pos;
			if (pos > kp) ++i;
		}
		x = nodes[x].ptr[i];
	}
}

// __kb_traverse (bwa/kbtree.h:329-352): in-order walk into ord[]
EMAB_HD int bt_traverse(const BNode *nodes, int root, int32_t *ord)
{
	int sx[24], si[24], sp = 0, n = 0;  // height <= log_5(#chains) + 1
	sx[0] = root; si[0] = 0;
	for (;;) {
		while (sx[sp] >= 0 && si[sp] <= nodes[sx[sp]].n) {
			const BNode &nd = nodes[sx[sp]];
			sx[sp + 1] = nd.is_internal ? nd.ptr[si[sp]] : -1;
			si[sp + 1] = 0;
			++sp;
		}
		--sp;
		if (sp < 0) break;
		if (sx[sp] >= 0 && si[sp] < nodes[sx[sp]].n) ord[n++] = nodes[sx[sp]].key[si[sp]];
		++si[sp];
	}
	return n;
}

// test_and_merge (bwa/bwamem.c:216-237)
EMAB_HD int test_and_merge(int64_t l_pac, Chain &c, Seed *seeds, int si, int seed_rid)
{
	const Seed &p = seeds[si];
	const Seed &last = seeds[c.seed_last];
	const Seed &first = seeds[c.seed_beg];
	int64_t qend = last.qbeg + last.len, rend = last.rbeg + last.len;
	if (seed_rid != c.rid) return 0;
	if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend) return 1;  // contained
	if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return 0;  // different strand
	int64_t x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
	if (y >= 0 && x - y <= opt::w && y - x <= opt::w && x - last.len < opt::max_chain_gap && y - last.len < opt::max_chain_gap) {
		seeds[c.seed_last].next = si;
		c.seed_last = si;
		++c.n;
		return 1;
	}
	return 0;
}

// mem_chain_weight (bwa/bwamem.c:239-258)
EMAB_HD int chain_weight(const Chain &c, const Seed *seeds)
{
	int64_t end = 0;
	int w = 0, tmp;
	for (int s = c.seed_beg, j = 0; j < c.n; ++j, s = seeds[s].next) {
		const Seed &sd = seeds[s];
		if (sd.qbeg >= end) w += sd.len;
		else if (sd.qbeg + sd.len > end) w += (int)(sd.qbeg + sd.len - end);
		end = end > sd.qbeg + sd.len ? end : sd.qbeg + sd.len;
	}
	tmp = w; w = 0; end = 0;
	for (int s = c.seed_beg, j = 0; j < c.n; ++j, s = seeds[s].next) {
		const Seed &sd = seeds[s];
		if (sd.rbeg >= end) w += sd.len;
		else if (sd.rbeg + sd.len > end) w += (int)(sd.rbeg + sd.len - end);
		end = end > sd.rbeg + sd.len ? end : sd.rbeg + sd.len;
	}
	w = w < tmp ? w : tmp;
	return w < (1 << 30) ? w : (1 << 30) - 1;
}

struct WIdxLess { EMAB_HD bool operator()(const WIdx &a, const WIdx &b) const { return a.w > b.w; } };  // flt_lt, bwa/bwamem.c:350

// mem_chain + mem_chain_flt for one read.  Writes the surviving chains (in mem_chain_flt's output
// order) to out_chains and their seeds, contiguous per chain, to out_seeds; returns their number.
// sa_vals (optional): the SA values of this read's occurrences in enumeration order, gathered ahead (k_sa_gather): the
// dependent DRAM gathers of bwt_sa then become reads of a contiguous array
EMAB_HD int chain_read(const DevIndex &ix, int len, const Intv *intv, int n_intv, ChainWork wk, int cap,
                       Chain *out_chains, Seed *out_seeds, const int64_t *sa_vals = nullptr)
{
	if (len < opt::min_seed_len || n_intv == 0 || cap == 0) return 0;
	const int64_t l_pac = ix.l_pac;
	// fraction of the read covered by over-frequent seeds (bwa/bwamem.c:291-298)
	int b = 0, e = 0, l_rep = 0;
	for (int i = 0; i < n_intv; ++i) {
		int sb = (int)(intv[i].info >> 32), se = (int)(uint32_t)intv[i].info;
		if (intv[i].x2 <= (uint64_t)opt::max_occ) continue;
		if (sb > e) { l_rep += e - b; b = sb; e = se; }
		else e = e > se ? e : se;
	}
	l_rep += e - b;
	const float frac_rep = (float)l_rep / len;

	int n_seeds = 0, n_chains = 0, n_nodes = 1, root = 0, occ_idx = 0;
	wk.nodes[0].n = 0; wk.nodes[0].is_internal = 0;
	for (int i = 0; i < n_intv; ++i) {
		const Intv p = intv[i];
		const int slen = (int)(uint32_t)p.info - (int)(p.info >> 32);
		const int64_t step = p.x2 > (uint64_t)opt::max_occ ? (int64_t)(p.x2 / opt::max_occ) : 1;
		int count = 0;
		for (int64_t k = 0; k < (int64_t)p.x2 && count < opt::max_occ; k += step, ++count) {
			int si = n_seeds;  // tentative slot
			Seed &s = wk.seeds[si];
			s.rbeg = sa_vals ? sa_vals[occ_idx] : (int64_t)bwt_sa_dense(ix, p.x0 + (uint64_t)k);
			++occ_idx;
			s.qbeg = (int)(p.info >> 32);
			s.score = s.len = slen;
			s.next = -1;
			int rid = bns_intv2rid(ix, s.rbeg, s.rbeg + s.len);
			if (rid < 0) continue;  // bridging contigs or the forward-reverse boundary
			bool to_add = true;
			if (n_chains) {
				int lower = bt_lower(wk.nodes, root, wk.chains, s.rbeg);
				if (lower >= 0 && test_and_merge(l_pac, wk.chains[lower], wk.seeds, si, rid)) to_add = false;
				if (!to_add) {  // merged: keep the slot only if the seed was appended (not merely contained)
					if (wk.chains[lower].seed_last == si) ++n_seeds;
				}
			}
			if (to_add) {
				Chain &c = wk.chains[n_chains];
				c.pos = s.rbeg; c.n = 1; c.rid = rid; c.seed_beg = c.seed_last = si;
				c.w = 0; c.kept = 0; c.first = -1; c.frac_rep = frac_rep;
				++n_seeds;
				bt_put(wk.nodes, &n_nodes, &root, wk.chains, n_chains);
				++n_chains;
			}
		}
	}
	if (n_chains == 0) return 0;
	// chains in key order (bwa/bwamem.c:330-334)
	int32_t *ord = wk.ord;
	int n = bt_traverse(wk.nodes, root, ord);
	// ---- mem_chain_flt (bwa/bwamem.c:353-411); min_chain_weight = 0 keeps every chain
	WIdx *a = (WIdx *)(wk.ord + cap);  // [cap] pairs live in the second half of ord (2 x int32 each => needs 2*cap ints)
	for (int i = 0; i < n; ++i) {
		Chain &c = wk.chains[ord[i]];
		c.first = -1; c.kept = 0;
		c.w = chain_weight(c, wk.seeds);
		c.qbeg0 = wk.seeds[c.seed_beg].qbeg;
		c.qend_last = wk.seeds[c.seed_last].qbeg + wk.seeds[c.seed_last].len;
		a[i].w = c.w; a[i].idx = ord[i];
	}
	ks_introsort((size_t)n, a, WIdxLess());
	// pairwise comparisons; `kept list` reuses ord[]
	int n_kept = 0;
	wk.chains[a[0].idx].kept = 3;
	ord[n_kept++] = 0;
	for (int i = 1; i < n; ++i) {
		Chain &ci = wk.chains[a[i].idx];
		int large_ovlp = 0, k;
		for (k = 0; k < n_kept; ++k) {
			int j = ord[k];
			Chain &cj = wk.chains[a[j].idx];
			int b_max = cj.qbeg0 > ci.qbeg0 ? cj.qbeg0 : ci.qbeg0;
			int e_min = cj.qend_last < ci.qend_last ? cj.qend_last : ci.qend_last;
			if (e_min > b_max) {  // overlap (no ALT contigs on this path: is_alt = 0)
				int li = ci.qend_last - ci.qbeg0, lj = cj.qend_last - cj.qbeg0;
				int min_l = li < lj ? li : lj;
				if (e_min - b_max >= min_l * opt::mask_level && min_l < opt::max_chain_gap) {
					large_ovlp = 1;
					if (cj.first < 0) cj.first = i;
					if (ci.w < cj.w * opt::drop_ratio && cj.w - ci.w >= (opt::min_seed_len << 1)) break;
				}
			}
		}
		if (k == n_kept) {
			ord[n_kept++] = i;
			ci.kept = large_ovlp ? 2 : 3;
		}
	}
	for (int i = 0; i < n_kept; ++i) {
		Chain &c = wk.chains[a[ord[i]].idx];
		if (c.first >= 0) wk.chains[a[c.first].idx].kept = 1;
	}
	// max_chain_extend = 1<<30 never bites (bwa/bwamem.c:399-404)
	int n_out = 0, so = 0;
	for (int i = 0; i < n; ++i) {
		Chain &c = wk.chains[a[i].idx];
		if (c.kept == 0) continue;
		Chain o = c;
		o.seed_beg = so;
		for (int s = c.seed_beg, j = 0; j < c.n; ++j, s = wk.seeds[s].next) {
			out_seeds[so] = wk.seeds[s];
			out_seeds[so].next = -1;
			++so;
		}
		o.seed_last = so - 1;
		out_chains[n_out++] = o;
	}
	return n_out;
}
