// Benchmark input generator of the B200 build: the same simulated genomes as the
// reference's test/simf.cxx (/root/reference/test/simf.cxx:93-140), produced in memory.
// BASELINE.json names simf as the source of the benchmark genomes; this is the
// product-side counterpart used by bench.py so that nothing on the measured path touches
// the oracle.  tests/test_simgen.py checks it byte for byte against the oracle's simf and
// against the files written by the unmodified simf binary.
#include <cmath>
#include <cstdint>
#include <random>

extern "C" {

// One genome of `length` bases.  base_seed selects the ancestral sequence, mut_seed the
// substitutions; divergence is a Jukes-Cantor distance unless raw != 0 (simf.cxx:62-68).
void phylo_simgen(uint32_t base_seed, uint32_t mut_seed, int64_t length, double divergence, int raw, char *out)
{
	const double p = raw ? divergence : 0.75 - 0.75 * std::exp(-(4.0 / 3.0) * divergence);
	std::default_random_engine base_rand{base_seed};
	std::uniform_int_distribution<int> base_dist{0, 3};
	// simf binds its mutation engine by value (simf.cxx:108): the coin flips and the choice
	// of the substituted base advance two separate copies seeded alike
	std::default_random_engine coin_rand{mut_seed}, pick_rand{mut_seed};
	std::uniform_real_distribution<double> coin{0, 1};
	std::uniform_int_distribution<int> pick{0, 2};
	static const char *const other[4] = {"CGT", "AGT", "ACT", "ACG"};
	double left = (double)length;
	double todo = left * p;
	for (int64_t k = 0; k < length; k++) {
		const int b = base_dist(base_rand);
		char c = "ACGT"[b];
		if (coin(coin_rand) < todo / left) {
			c = other[b][pick(pick_rand)];
			todo--;
		}
		out[k] = c;
		left--;
	}
}
}
