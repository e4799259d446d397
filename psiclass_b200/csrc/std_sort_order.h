// std_sort_order.h -- the permutation libstdc++'s std::sort leaves equal elements in.
//
// Constraints::CoalesceSameConstraints (Constraints.cpp:74-113) sorts the constraints of a component with
// std::sort and then adds the weights 10/NH of equal bit vectors IN THAT ORDER.  For NH = 3 the addends
// are inexact, so the fp64 sum depends on where std::sort happened to leave each of the equal elements
// (quirk Q9) -- and a last-bit difference in a weight is enough to flip exact ties further down
// (set-cover arg-max, DP strict '>' folds).  std::sort's result is a deterministic function of the input
// sequence and of the comparator's answers, so it can be reproduced by making the same moves: this file
// restates the algorithm of libstdc++ (bits/stl_algo.h, unchanged since GCC 4.x through 14: introsort with
// a median-of-three pivot moved to the front, depth limit 2*floor(log2 n) with a heap-sort fallback, a
// final insertion sort with threshold 16) over an array of ints.  One thread runs it.
//
// tests/std_sort_check.cpp compares it with the real std::sort / std::partial_sort on the host.
#pragma once

#ifndef PSC_DEV
#define PSC_DEV inline
#endif

namespace pscdev {

template <class Less> PSC_DEV void sso_unguarded_linear_insert(int *a, int last, Less less)
{
	int val = a[last];
	int next = last - 1;
	while (less(val, a[next])) {
		a[last] = a[next];
		last = next;
		--next;
	}
	a[last] = val;
}

template <class Less> PSC_DEV void sso_insertion_sort(int *a, int first, int last, Less less)
{
	if (first == last) return;
	for (int i = first + 1; i != last; ++i) {
		if (less(a[i], a[first])) {
			int val = a[i];
			for (int q = i; q > first; --q) a[q] = a[q - 1];
			a[first] = val;
		} else
			sso_unguarded_linear_insert(a, i, less);
	}
}

template <class Less> PSC_DEV void sso_push_heap(int *a, int first, int hole, int top, int value, Less less)
{
	int parent = (hole - 1) / 2;
	while (hole > top && less(a[first + parent], value)) {
		a[first + hole] = a[first + parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	a[first + hole] = value;
}

template <class Less> PSC_DEV void sso_adjust_heap(int *a, int first, int hole, int len, int value, Less less)
{
	const int top = hole;
	int child = hole;
	while (child < (len - 1) / 2) {
		child = 2 * (child + 1);
		if (less(a[first + child], a[first + child - 1])) --child;
		a[first + hole] = a[first + child];
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2) {
		child = 2 * (child + 1);
		a[first + hole] = a[first + child - 1];
		hole = child - 1;
	}
	sso_push_heap(a, first, hole, top, value, less);
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
template <class Less> PSC_DEV void sso_heap_sort(int *a, int first, int last, Less less)
{
	int len = last - first;
	if (len >= 2) {
		int parent = (len - 2) / 2;
		for (;;) {
			int value = a[first + parent];
			sso_adjust_heap(a, first, parent, len, value, less);
			if (parent == 0) break;
			--parent;
		}
	}
	while (last - first > 1) {
		--last;
		int value = a[last];
		a[last] = a[first];
		sso_adjust_heap(a, first, 0, last - first, value, less);
	}
}

template <class Less> PSC_DEV int sso_partition_pivot(int *a, int first, int last, Less less)
{
	int mid = first + (last - first) / 2;
	{ // __move_median_to_first(first, first + 1, mid, last - 1)
		int ia = first + 1, ib = mid, ic = last - 1, pick;
		if (less(a[ia], a[ib])) {
			if (less(a[ib], a[ic])) pick = ib;
			else if (less(a[ia], a[ic])) pick = ic;
			else pick = ia;
		} else if (less(a[ia], a[ic])) pick = ia;
		else if (less(a[ib], a[ic])) pick = ic;
		else pick = ib;
		int t = a[first];
		a[first] = a[pick];
		a[pick] = t;
	}
	int lo = first + 1, hi = last;
	const int pivot = a[first];
	for (;;) {
		while (less(a[lo], pivot)) ++lo;
		--hi;
		while (less(pivot, a[hi])) --hi;
		if (!(lo < hi)) return lo;
		int t = a[lo];
		a[lo] = a[hi];
		a[hi] = t;
		++lo;
	}
}

// a[0, n) is rearranged exactly as std::sort(a, a + n, less) would
template <class Less> PSC_DEV void std_sort_order(int *a, int n, Less less)
{
	if (n <= 1) return;
	// __introsort_loop.  Its recursion on the right part becomes an explicit stack: the sub-ranges are disjoint,
	// so the order in which they are finished does not change where any element ends up.
	int stFirst[64], stLast[64], stDepth[64];
	int sp = 0;
	int lg = 0;
	for (int m = n; m > 1; m >>= 1) ++lg; // std::__lg(n)
	stFirst[0] = 0;
	stLast[0] = n;
	stDepth[0] = 2 * lg;
	sp = 1;
	while (sp > 0) {
		--sp;
		int first = stFirst[sp], last = stLast[sp], depth = stDepth[sp];
		while (last - first > 16) {
			if (depth == 0) {
				sso_heap_sort(a, first, last, less);
				break;
			}
			--depth;
			int cut = sso_partition_pivot(a, first, last, less);
			stFirst[sp] = cut;
			stLast[sp] = last;
			stDepth[sp] = depth;
			++sp;
			last = cut;
		}
	}
	// __final_insertion_sort
	if (n > 16) {
		sso_insertion_sort(a, 0, 16, less);
		for (int i = 16; i < n; ++i) sso_unguarded_linear_insert(a, i, less);
	} else
		sso_insertion_sort(a, 0, n, less);
}

} // namespace pscdev
