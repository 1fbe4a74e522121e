// phylonium-b200: the `phylonium` command line on top of libphylonium_b200.so.
//
// Written fresh for this repository; it keeps what BASELINE.json's north star leaves on the
// host — option parsing, FASTA loading, reference choice, the process()/evo_model
// interfaces and the PHYLIP printer (/root/reference/src/phylonium.cxx, io.cxx, sequence.cxx)
// — and calls the CUDA library for the hot path.  Output on stdout is byte-identical to the
// reference's for the options supported here.
#include "phylonium_host.h"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <err.h>
#include <fstream>
#include <functional>
#include <getopt.h>
#include <iostream>
#include <limits>
#include <atomic>
#include <numeric>
#include <random>
#include <sys/stat.h>
#include <thread>

#include "../../include/phylonium_b200.h"

int FLAGS = flags::none;
int RETURN_CODE = EXIT_SUCCESS;
size_t reference_index = 0;

// ------------------------------------------------------------------ sequence helpers

std::string reverse(const std::string &base)
{
	std::string r(base.size(), '\0');
	const size_t n = base.size();
	for (size_t k = 0; k < n; k++) {
		unsigned char c = (unsigned char)base[n - 1 - k];
		if (c >= 'A') c ^= (c & 2) ? 4 : 21; // C<->G differ in bit 2, A<->T in 0x15
		r[k] = (char)c;
	}
	return r;
}

std::string filter_nucl(const std::string &base)
{
	std::string r;
	r.reserve(base.size());
	for (char c : base) {
		switch (c) {
			case 'A': case 'C': case 'G': case 'T': r += c; break;
			case 'a': r += 'A'; break;
			case 'c': r += 'C'; break;
			case 'g': r += 'G'; break;
			case 't': r += 'T'; break;
			default: break;
		}
	}
	return r;
}

double gc_content(const std::string &seq) noexcept
{
	return phylo_gc_content(seq.data(), seq.size());
}

sequence join(const genome &gen)
{
	const auto &contigs = gen.get_contigs();
	if (contigs.empty()) return sequence();
	if (contigs.size() == 1) return sequence(gen.get_name(), contigs[0].get_nucl());
	std::string all = contigs[0].get_nucl();
	for (size_t k = 1; k < contigs.size(); k++) {
		all += '!';
		all += contigs[k].get_nucl();
	}
	return sequence(gen.get_name(), all);
}

// ------------------------------------------------------------------ evo_model

double evo_model::estimate_raw(bool zero_on_error) const noexcept
{
	if (homologs == 0) return zero_on_error ? 0.0 : NAN;
	return substitutions / (double)homologs;
}

double evo_model::estimate_ani(bool zero_on_error) const noexcept
{
	if (homologs == 0) return zero_on_error ? 0.0 : NAN;
	return (1.0 - substitutions / (double)homologs) * 100;
}

double evo_model::estimate_JC(bool zero_on_error) const noexcept
{
	double dist = estimate_raw(zero_on_error);
	dist = -0.75 * log(1.0 - (4.0 / 3.0) * dist);
	return dist <= 0.0 ? 0.0 : dist;
}

// ------------------------------------------------------------------ reference positions (-p)

// src/process.cxx:471-513: after complete deletion all sequences cover the same parts of the
// reference; for every part print its range, the positions at which some sequence differs
// from the first one, and the reference's bases.  The library hands back the core genome as
// bitmaps over the reference columns (phylo_core_sites).
static std::string REFPOS_FILE_NAME;
static size_t BOOTSTRAP = 0; // replicates to print after the matrix itself (-b N prints N matrices)

static void write_reference_positions(phylo_ctx *ctx, const sequence &subject)
{
	uint64_t words = 0;
	if (phylo_core_sites(ctx, nullptr, nullptr, nullptr, &words) != PHYLO_OK) errx(1, "%s", phylo_last_error(ctx));
	std::vector<uint32_t> core(words), border(words), seg(words);
	if (phylo_core_sites(ctx, core.data(), border.data(), seg.data(), &words) != PHYLO_OK)
		errx(1, "%s", phylo_last_error(ctx));
	auto bit = [](const std::vector<uint32_t> &v, size_t c) { return (v[c >> 5] >> (c & 31)) & 1u; };
	std::ofstream out(REFPOS_FILE_NAME);
	const size_t n = subject.size();
	size_t counter = 1;
	size_t c = 0;
	while (c < n) {
		if (!core[c >> 5]) { // nothing in this word
			c = ((c >> 5) + 1) << 5;
			continue;
		}
		if (!bit(core, c)) {
			c++;
			continue;
		}
		const size_t start = c;
		std::vector<size_t> pos;
		do {
			if (bit(seg, c)) pos.push_back(c - start);
			c++;
		} while (c < n && bit(core, c) && !bit(border, c));
		const size_t end = c;
		out << ">part" << counter++ << "\t(" << (start + 1) << ".." << (end + 1) << ")  " << pos.size();
		for (size_t p : pos)
			out << "  " << (p + 1);
		out << std::endl;
		out << std::string(subject.c_str() + start, subject.c_str() + end) << std::endl;
	}
}

// ------------------------------------------------------------------ the seam

// one context for the life of the program (the second pass of --2pass reuses it)
static phylo_ctx *context()
{
	static phylo_ctx *ctx = nullptr;
	if (!ctx && phylo_ctx_create(-1, &ctx) != PHYLO_OK) errx(1, "%s", phylo_last_error(nullptr));
	return ctx;
}

// the strings whose bytes the context already holds on the device (handed over while the
// files were being read, or by the previous process() call)
static std::vector<const char *> resident_ptr;
static std::vector<uint64_t> resident_len;

std::vector<evo_model> process(const sequence &subject, const std::vector<sequence> &queries)
{
	const size_t N = queries.size();
	size_t ref = reference_index;
	if (ref >= N || !(queries[ref] == subject)) {
		ref = std::find(queries.begin(), queries.end(), subject) - queries.begin();
		if (ref == N) errx(1, "process(): the subject is not one of the queries");
	}
	phylo_ctx *ctx = context();

	std::vector<const char *> ptr(N);
	std::vector<uint64_t> len(N);
	for (size_t k = 0; k < N; k++) {
		ptr[k] = queries[k].c_str();
		len[k] = queries[k].size();
	}
	if (FLAGS & flags::verbose) std::cerr << "ref: " << subject.get_name() << std::endl;
	std::vector<uint64_t> subst(N * N), homol(N * N);
	// Strings the device already holds — uploaded while the files were read (read_and_upload),
	// or by the first call of --2pass (src/phylonium.cxx:289-296) — are not sent again: only the
	// index and the mapping are done.
	int rc;
	if (ptr == resident_ptr && len == resident_len) {
		rc = phylo_process_again(ctx, ref, FLAGS & flags::complete_deletion, subst.data(), homol.data());
	} else {
		rc = phylo_process(ctx, ptr.data(), len.data(), N, ref, FLAGS & flags::complete_deletion, subst.data(), homol.data());
		resident_ptr = ptr;
		resident_len = len;
	}
	if (rc != PHYLO_OK) errx(1, "%s", phylo_last_error(ctx));
	if (FLAGS & flags::print_positions) write_reference_positions(ctx, subject);
	std::vector<evo_model> matrix(N * N);
	for (size_t k = 0; k < N * N; k++)
		matrix[k] = evo_model(subst[k], homol[k]);
	return matrix;
}

// ------------------------------------------------------------------ I/O

static std::string genome_name(const std::string &file)
{
	size_t left = file.rfind('/');
	left = left == std::string::npos ? 0 : left + 1;
	size_t right = file.rfind('.');
	if (right == std::string::npos || right < left) {
		right = file.size();
	} else {
		const std::string ext = file.substr(right);
		if (ext != ".fa" && ext != ".fas" && ext != ".fasta") right = file.size();
	}
	return file.substr(left, right - left);
}

genome read_genome(const std::string &file_name)
{
	std::ifstream in(file_name, std::ios::binary);
	if (!in) err(1, "%s", file_name.c_str());
	std::vector<sequence> contigs;
	std::string line, name, bases;
	bool have = false;
	auto flush = [&] {
		if (have) contigs.emplace_back(name, filter_nucl(bases));
		bases.clear();
	};
	while (std::getline(in, line)) {
		if (!line.empty() && line[0] == '>') {
			flush();
			have = true;
			size_t e = line.find_first_of(" \t\r", 1);
			name = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
		} else if (line.empty() || line[0] == ';') {
			continue;
		} else {
			if (!have) errx(1, "%s: expected '>' at the start of the file", file_name.c_str());
			bases += line;
		}
	}
	flush();
	if (contigs.empty()) errx(1, "%s: no sequence found", file_name.c_str());
	return genome(genome_name(file_name), std::move(contigs));
}

// Reads all files with a pool of threads; every genome goes to the device (packed, see
// phylo_ingest_put) the moment its file is parsed, while the other files are still being read
// — the reference parses everything first (src/phylonium.cxx:255-270, src/io.cxx:66-104).
static std::vector<sequence> read_and_upload(const std::vector<std::string> &files)
{
	const size_t N = files.size();
	std::vector<sequence> queries(N);
	std::vector<uint64_t> caps(N);
	for (size_t i = 0; i < N; i++) {
		struct stat st;
		if (stat(files[i].c_str(), &st) != 0) err(1, "%s", files[i].c_str());
		caps[i] = (uint64_t)st.st_size; // a sequence is never longer than its file
	}
	const unsigned hw = std::thread::hardware_concurrency();
	const size_t threads = std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 1), N, (size_t)16}));
	phylo_ctx *ctx = context();
	if (phylo_ingest_begin(ctx, N, caps.data(), (int)threads) != PHYLO_OK) errx(1, "%s", phylo_last_error(ctx));
	std::atomic<size_t> next{0};
	auto worker = [&] {
		for (size_t i = next++; i < N; i = next++) {
			queries[i] = join(read_genome(files[i]));
			phylo_ingest_put(ctx, i, queries[i].c_str(), queries[i].size()); // failures surface in phylo_ingest_end
		}
	};
	std::vector<std::thread> pool;
	for (size_t t = 1; t < threads; t++)
		pool.emplace_back(worker);
	worker();
	for (auto &t : pool)
		t.join();
	if (phylo_ingest_end(ctx) != PHYLO_OK) errx(1, "%s", phylo_last_error(ctx));
	resident_ptr.resize(N);
	resident_len.resize(N);
	for (size_t i = 0; i < N; i++) {
		resident_ptr[i] = queries[i].c_str();
		resident_len[i] = queries[i].size();
	}
	return queries;
}

static void soft_warnx(const char *fmt, const char *a, const char *b, double x = 0, double y = 0)
{
	RETURN_CODE |= EXIT_FAILURE;
	warnx(fmt, a, b, x, y);
}

static void just_print(const std::vector<std::string> &names, const std::vector<double> &dist)
{
	const size_t N = names.size();
	std::cout << N << std::endl;
	std::cout.precision(4);
	std::cout << (FLAGS & flags::dist_ani ? std::dec : std::scientific);
	for (size_t i = 0; i < N; i++) {
		std::cout << names[i];
		for (size_t j = 0; j < N; j++)
			std::cout << "  " << (i == j ? 0.0 : dist[i * N + j]);
		std::cout << std::endl;
	}
}

void print_matrix(const std::vector<sequence> &queries, const std::vector<evo_model> &matrix)
{
	const size_t N = queries.size();
	std::vector<std::string> names(N);
	for (size_t i = 0; i < N; i++)
		names[i] = queries[i].get_name();
	std::vector<double> dist(N * N, NAN);
	for (size_t k = 0; k < N * N; k++)
		dist[k] = FLAGS & flags::dist_raw   ? matrix[k].estimate_raw()
		          : FLAGS & flags::dist_ani ? matrix[k].estimate_ani()
		                                    : matrix[k].estimate_JC();
	double sum = 0;
	size_t counter = 0;
	for (size_t i = 0; i < N; i++) {
		for (size_t j = 0; j < i; j++) {
			const size_t k = i * N + j;
			if (std::isnan(dist[k])) {
				soft_warnx("For the two sequences '%s' and '%s' the distance computation failed and is reported as nan.",
				           names[i].c_str(), names[j].c_str());
				continue;
			}
			const double c1 = matrix[k].coverage(queries[i].size()), c2 = matrix[k].coverage(queries[j].size());
			if (c1 < 0.2 || c2 < 0.2)
				soft_warnx("For the two sequences '%s' and '%s' less than 20%% homology were found (%f and %f, "
				           "respectively).",
				           names[i].c_str(), names[j].c_str(), c1, c2);
			sum += c1 + c2;
			counter += 2;
		}
	}
	just_print(names, dist);
	// -b: more matrices from resampled substitution counts (src/io.cxx:188-200 and
	// evo_model::bootstrap, src/evo_model.cxx:136-147: binomial in the number of homologous
	// positions, every cell on its own).  Like the reference's, the generator is seeded from
	// std::random_device, so replicates differ from run to run.
	if (BOOTSTRAP) {
		static std::mt19937 prng{std::random_device{}()};
		std::vector<double> neu(N * N);
		for (size_t rep = 0; rep < BOOTSTRAP; rep++) {
			for (size_t k = 0; k < N * N; k++) {
				evo_model em = matrix[k];
				if (em.homologs) {
					const double rate = em.substitutions / (double)em.homologs;
					std::binomial_distribution<long long> d((long long)em.homologs, rate);
					em.substitutions = (uint64_t)d(prng);
				}
				neu[k] = FLAGS & flags::dist_raw ? em.estimate_raw() : FLAGS & flags::dist_ani ? em.estimate_ani() : em.estimate_JC();
			}
			just_print(names, neu);
		}
	}
	if (FLAGS & flags::verbose) {
		uint64_t aligned = 0, total = 0;
		for (size_t i = 0; i < N; i++) {
			if (i == reference_index) continue;
			aligned += matrix[reference_index * N + i].total();
			total += queries[i].size();
		}
		std::cerr << "avg coverage:\t" << sum / counter << std::endl;
		std::cerr << "alignment:\t" << aligned << "\t" << total << "\t" << aligned / (double)total << std::endl;
	}
}

// ------------------------------------------------------------------ reference choice

static size_t pick_first_pass(std::vector<sequence> &sequences)
{
	// medium length, same nth_element call as src/phylonium.cxx:360-382
	std::vector<std::reference_wrapper<sequence>> ret(sequences.begin(), sequences.end());
	std::nth_element(ret.begin(), ret.begin() + ret.size() / 2, ret.end(),
	                 [](const sequence &a, const sequence &b) { return a.size() < b.size(); });
	auto &reference = ret[ret.size() / 2].get();
	reference_index = std::find(sequences.begin(), sequences.end(), reference) - sequences.begin();
	if (FLAGS & flags::verbose) std::cerr << "chosen reference: " << reference.get_name() << std::endl;
	return reference_index;
}

static size_t pick_second_pass(const std::vector<sequence> &sequences, const std::vector<evo_model> &matrix)
{
	// most central sequence: smallest row sum of JC distances (src/phylonium.cxx:317-344)
	const size_t N = sequences.size();
	double best = std::numeric_limits<double>::max();
	size_t at = 0;
	for (size_t i = 0; i < N; i++) {
		double sum = 0.0;
		for (size_t j = 0; j < N; j++)
			sum += matrix[i * N + j].estimate_JC(true);
		if (sum < best) {
			best = sum;
			at = i;
		}
	}
	reference_index = at;
	return at;
}

static void usage(int status)
{
	static const char str[] =
		"Usage: phylonium-b200 [OPTIONS] FILES...\n"
		"\tFILES... can be any sequence of FASTA files, each file representing one genome.\n\n"
		"Options:\n"
		"  -2, --2pass          Enable two-pass algorithm\n"
		"  -b, --bootstrap=N    Print additional bootstrap matrices\n"
		"  --complete-deletion  Delete the whole aligned column in case of gaps\n"
		"  -p FILE              Print reference positions to FILE (implies complete deletion)\n"
		"  -r FILE              Set the reference genome\n"
		"  -t, --threads=N      Accepted for compatibility (the work runs on the GPU)\n"
		"  -v, --verbose        Print additional information\n"
		"      --distance=OPT   Choose between raw, jc corrected and ANI\n"
		"      --progress=WHEN  Accepted for compatibility\n"
		"  -h, --help           Display this help and exit\n"
		"      --version        Output version information\n";
	fprintf(status == EXIT_SUCCESS ? stdout : stderr, "%s", str);
	exit(status);
}

int main(int argc, char *argv[])
{
	bool two_pass = false;
	std::string reference_name;
	static struct option long_options[] = {{"2pass", no_argument, nullptr, '2'},
	                                       {"bootstrap", required_argument, nullptr, 'b'},
	                                       {"complete-deletion", no_argument, nullptr, 0},
	                                       {"distance", required_argument, nullptr, 0},
	                                       {"progress", required_argument, nullptr, 0},
	                                       {"help", no_argument, nullptr, 'h'},
	                                       {"threads", required_argument, nullptr, 't'},
	                                       {"verbose", no_argument, nullptr, 'v'},
	                                       {"version", no_argument, nullptr, 0},
	                                       {nullptr, 0, nullptr, 0}};
	for (;;) {
		int idx = 0;
		const int c = getopt_long(argc, argv, "2b:hp:r:t:v", long_options, &idx);
		if (c == -1) break;
		switch (c) {
			case 0: {
				const std::string opt = long_options[idx].name;
				if (opt == "complete-deletion") {
					FLAGS |= flags::complete_deletion;
				} else if (opt == "distance") {
					const std::string a = optarg;
					if (a == "raw")
						FLAGS |= flags::dist_raw;
					else if (a == "ani")
						FLAGS |= flags::dist_ani;
					else if (a != "jc") {
						RETURN_CODE |= EXIT_FAILURE;
						warnx("unknown distance '%s', falling back to jc", optarg);
					}
				} else if (opt == "version") {
					printf("phylonium-b200 (%s)\n", phylo_version());
					return 0;
				}
				break;
			}
			case '2': two_pass = true; break;
			case 'b': { // src/phylonium.cxx:165-180
				errno = 0;
				char *end;
				const unsigned long bootstrap = strtoul(optarg, &end, 10);
				if (errno || end == optarg || *end != '\0' || bootstrap == 0) {
					RETURN_CODE |= EXIT_FAILURE;
					warnx("Expected a positive number for -b argument, but '%s' was given. Ignoring -b argument.", optarg);
					break;
				}
				BOOTSTRAP = bootstrap - 1;
				break;
			}
			case 'h': usage(EXIT_SUCCESS); break;
			case 'p': // src/phylonium.cxx:183-188
				FLAGS |= flags::print_positions | flags::complete_deletion;
				REFPOS_FILE_NAME = optarg;
				break;
			case 'r': reference_name = optarg; break;
			case 't': break;
			case 'v': FLAGS |= (FLAGS & flags::verbose) ? flags::extra_verbose : flags::verbose; break;
			default: usage(EXIT_FAILURE);
		}
	}
	if (FLAGS & flags::print_positions) { // src/phylonium.cxx:233-240: avoid overwriting files
		std::ifstream file(REFPOS_FILE_NAME);
		if (file.good()) errx(1, "output file '%s' already exists", REFPOS_FILE_NAME.c_str());
	}
	std::vector<std::string> files(argv + optind, argv + argc);
	if (!reference_name.empty()) { // src/phylonium.cxx:384-391: the list gets sorted and made unique
		files.push_back(reference_name);
		std::sort(files.begin(), files.end());
		files.erase(std::unique(files.begin(), files.end()), files.end());
	}
	if (files.size() < 2) usage(EXIT_FAILURE);

	std::vector<sequence> queries = read_and_upload(files);

	if (reference_name.empty())
		pick_first_pass(queries);
	else
		reference_index = std::find(files.begin(), files.end(), reference_name) - files.begin();

	auto matrix = process(queries[reference_index], queries);
	if (two_pass) {
		const size_t again = pick_second_pass(queries, matrix);
		matrix = process(queries[again], queries);
	}
	print_matrix(queries, matrix);
	return RETURN_CODE;
}
