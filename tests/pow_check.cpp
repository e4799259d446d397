// tests/pow_check.cpp -- pow_quarter_libm (the device routine, compiled for the host) against the host libm.
// Built and run by tests/test_host_and_shim.py; prints "mismatches 0" when every result equals pow(x, 0.25).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../psiclass_b200/csrc/pow_quarter.h"
int main(int argc, char **argv)
{
	long long n = argc > 1 ? atoll(argv[1]) : 20000000, bad = 0;
	std::mt19937_64 g(20260);
	for (long long i = 0; i < n; ++i) {
		uint64_t u = g();
		double x = 0.1 + 0.9 * ((u >> 11) * (1.0 / 9007199254740992.0));
		if (i % 3 == 0) { // ratios like the abundances the score sees
			double a = 1 + (u % 100000) / 7.0, b = a + ((u >> 20) % 100000) / 3.0;
			x = a / b;
			if (!(x > 0.1)) continue;
		}
		if (i % 1000 == 1) x = 1.0 - (u % 1000) * 1.1102230246251565e-16;
		if (std::pow(x, 0.25) != pscdev::pow_quarter_libm(x)) ++bad;
	}
	printf("checked %lld mismatches %lld\n", n, bad);
	return bad ? 1 : 0;
}
