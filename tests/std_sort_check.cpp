// TEST: psiclass_b200/csrc/std_sort_order.h against the real std::sort of this toolchain's libstdc++.
// Elements carry a payload that the comparator ignores, so equal keys are distinguishable and the exact
// permutation can be compared.  Includes sequences that drive std::sort into its heap-sort fallback.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../psiclass_b200/csrc/std_sort_order.h"

static uint64_t rs = 88172645463325252ull;
static uint32_t rnd()
{
	rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;
	return (uint32_t)(rs >> 16);
}

int main(int argc, char **argv)
{
	int rounds = argc > 1 ? atoi(argv[1]) : 20000;
	long long bad = 0, cases = 0, heapCases = 0;
	auto less = [](int x, int y) { return (x >> 12) < (y >> 12); };
	for (int r = 0; r < rounds; ++r) {
		int n = 1 + (int)(rnd() % (r % 50 == 0 ? 40000 : (r % 7 == 0 ? 3000 : 200)));
		int keys = 1 + (int)(rnd() % (r % 3 == 0 ? 4 : (r % 3 == 1 ? 60 : 100000)));
		std::vector<int> a((size_t)n);
		int shape = (int)(rnd() % 6);
		for (int i = 0; i < n; ++i) {
			int k;
			if (shape == 0) k = (int)(rnd() % (uint32_t)keys);
			else if (shape == 1) k = i * keys / n;                 // sorted with runs
			else if (shape == 2) k = (n - 1 - i) * keys / n;       // reversed
			else if (shape == 3) k = (i & 1) ? i : n - i;          // organ pipe-ish
			else if (shape == 4) k = (int)((rnd() % (uint32_t)keys) * (rnd() & 1)); // half zeros
			else k = (i < n / 2) ? 2 * i : 2 * (n - i) + 1;
			a[(size_t)i] = ((k & 0x7ffff) << 12) | (i & 0xfff);
		}
		if (r % 97 == 0 && n > 64) {
			// median-of-three killer (Musser): forces quadratic partitions, hence the heap-sort fallback
			int m = n & ~1;
			for (int i = 0; i < m / 2; ++i) {
				a[(size_t)(2 * i)] = ((i + 1) << 12) | ((2 * i) & 0xfff);
				a[(size_t)(2 * i + 1)] = ((m / 2 + i + 1) << 12) | ((2 * i + 1) & 0xfff);
			}
			for (int i = 0; i < m / 2; i += 2) a[(size_t)i] = (a[(size_t)i] & 0xfff) | ((i + 1) << 12);
			++heapCases;
		}
		std::vector<int> b = a, c = a;
		std::sort(b.begin(), b.end(), less);
		pscdev::std_sort_order(c.data(), n, less);
		++cases;
		if (b != c) {
			++bad;
			if (bad < 5) fprintf(stderr, "mismatch: round %d n %d keys %d shape %d\n", r, n, keys, shape);
		}
		// the heap-sort fallback on its own
		if (r % 11 == 0) {
			std::vector<int> d = a, e = a;
			std::partial_sort(d.begin(), d.end(), d.end(), less);
			pscdev::sso_heap_sort(e.data(), 0, n, less);
			++cases;
			if (d != e) {
				++bad;
				if (bad < 5) fprintf(stderr, "heap mismatch: round %d n %d\n", r, n);
			}
		}
	}
	printf("cases %lld killer %lld mismatches %lld\n", cases, heapCases, bad);
	return bad ? 1 : 0;
}
