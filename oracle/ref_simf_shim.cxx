/* TEST INFRASTRUCTURE — part of oracle/_ref/libphylo_ref.so (see ref_shim.cxx).
 * Exposes the reference's genome simulator print_seq()
 * (/root/reference/test/simf.cxx:93-140, included unmodified; its main() is
 * renamed out of the way) so tests and bench.py can make the BASELINE.json
 * inputs in memory instead of through FASTA files. */
#include <cmath>
#include <cstring>
#include <sstream>
#include <string>

#define main po_ref_simf_main
#include "simf.cxx" // /root/reference/test/simf.cxx
#undef main

#include "po_api.h"

extern "C" void po_simf(uint32_t base_seed, uint32_t mut_seed, int64_t length, double divergence,
                        int raw, char *out)
{
	// simf.cxx:62-68 — JC distance to per-site substitution probability
	double p = raw ? divergence : 0.75 - 0.75 * exp(-(4.0 / 3.0) * divergence);
	std::ostringstream os;
	print_seq(os, base_seed, mut_seed, (size_t)length, (size_t)70, p);
	const std::string s = os.str();
	int64_t w = 0;
	for (char c : s)
		if (c != '\n' && w < length) out[w++] = c;
}
