// klib's ks_introsort, restated (bwa/ksort.h:176-226 plus ks_combsort :154-174 and the final
// insertion sort :144-152).  BWA-MEM sorts keys that tie often (chain weight, region end,
// score/rb/qb) with this *unstable but deterministic* routine, so the order of equal keys in the
// output — and therefore which chain / region survives the later filters — is a function of the
// exact sequence of comparisons and swaps.  Reproducing it is part of the output contract
// (SURVEY.md §7 hard part 2).  Works on any element type with a strict-weak `lt`.
#pragma once
#include "common.cuh"

template <class T, class LT>
EMAB_HD void ks_insertsort(T *s, T *t, LT lt)
{
	for (T *i = s + 1; i < t; ++i)
		for (T *j = i; j > s && lt(*j, *(j - 1)); --j) {
			T tmp = *j; *j = *(j - 1); *(j - 1) = tmp;
		}
}

template <class T, class LT>
EMAB_HD void ks_combsort(size_t n, T *a, LT lt)
{
	const double shrink = 1.2473309501039786540366528676643;
	bool swapped;
	size_t gap = n;
	do {
		if (gap > 2) {
			gap = (size_t)(gap / shrink);
			if (gap == 9 || gap == 10) gap = 11;
		}
		swapped = false;
		for (T *i = a; i < a + n - gap; ++i) {
			T *j = i + gap;
			if (lt(*j, *i)) { T tmp = *i; *i = *j; *j = tmp; swapped = true; }
		}
	} while (swapped || gap > 2);
	if (gap != 1) ks_insertsort(a, a + n, lt);
}

template <class T, class LT>
EMAB_HD void ks_introsort(size_t n, T *a, LT lt)
{
	struct Frame { T *left, *right; int depth; };
	if (n < 1) return;
	if (n == 2) {
		if (lt(a[1], a[0])) { T tmp = a[0]; a[0] = a[1]; a[1] = tmp; }
		return;
	}
	int d;
	for (d = 2; (1ul << d) < n; ++d) {}
	Frame stack[72];  // at most one frame per partition level, and levels are capped at 2*ceil(log2 n)
	Frame *top = stack;
	T *s = a, *t = a + (n - 1);
	d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) {  // too deep: comb sort this range
				ks_combsort((size_t)(t - s) + 1, s, lt);
				t = s;
				continue;
			}
			T *i = s, *j = t, *k = i + ((j - i) >> 1) + 1;
			// median of three; the pivot ends up at *t
			if (lt(*k, *i)) {
				if (lt(*k, *j)) k = j;
			} else k = lt(*j, *i) ? i : j;
			T rp = *k;
			if (k != t) { T tmp = *k; *k = *t; *t = tmp; }
			for (;;) {
				do ++i; while (lt(*i, rp));
				do --j; while (i <= j && lt(rp, *j));
				if (j <= i) break;
				T tmp = *i; *i = *j; *j = tmp;
			}
			{ T tmp = *i; *i = *t; *t = tmp; }
			if (i - s > t - i) {
				if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == stack) {
				ks_insertsort(a, a + n, lt);
				return;
			}
			--top;
			s = top->left; t = top->right; d = top->depth;
		}
	}
}
