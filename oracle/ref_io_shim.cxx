/* TEST INFRASTRUCTURE — part of oracle/_ref/libphylo_ref.so (see ref_shim.cxx).
 * Routes the reference's own PHYLIP printer just_print()
 * (/root/reference/src/io.cxx:141-163, compiled unmodified) into a caller buffer
 * by swapping std::cout's stream buffer for the duration of the call. */
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "global.h"
#include "po_api.h"

void just_print(const std::vector<std::string> &names, const std::vector<double> &dist_matrix);

extern "C" int64_t po_format_matrix(const char *const *names, const uint64_t *subst,
                                    const uint64_t *homologs, int64_t N, int kind, char *out,
                                    int64_t cap)
{
	std::vector<std::string> nm;
	for (int64_t i = 0; i < N; i++)
		nm.emplace_back(names[i]);
	std::vector<double> dist((size_t)(N * N), NAN);
	for (int64_t k = 0; k < N * N; k++)
		dist[(size_t)k] = po_estimate(subst[k], homologs[k], kind);

	int saved = FLAGS;
	FLAGS &= ~(flags::dist_ani | flags::dist_raw);
	if (kind == 2) FLAGS |= flags::dist_ani;
	if (kind == 0) FLAGS |= flags::dist_raw;

	std::ostringstream os;
	auto *old = std::cout.rdbuf(os.rdbuf());
	auto oldflags = std::cout.flags();
	auto oldprec = std::cout.precision();
	just_print(nm, dist);
	std::cout.flush();
	std::cout.rdbuf(old);
	std::cout.flags(oldflags);
	std::cout.precision(oldprec);
	FLAGS = saved;

	std::string s = os.str();
	if ((int64_t)s.size() < cap) std::memcpy(out, s.data(), s.size());
	return (int64_t)s.size();
}
