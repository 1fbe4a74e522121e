// Host-side C++ mirror of the reference interfaces around the hot path, written fresh
// for the B200 build.  Same names and meaning as the reference so that a phylonium user
// finds what they expect:
//   sequence / genome / join / reverse / filter_nucl / gc_content   src/sequence.h:18-162
//   evo_model (two counters + estimators)                          src/evo_model.h:13-50
//   process(subject, queries) -> N*N evo_model                     src/process.h:12
//   read_genome / print_matrix                                     src/io.h
// process() does no computing itself: it hands the sequences to libphylonium_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

struct phylo_ctx;

class sequence
{
	std::string name_, nucl_;

  public:
	sequence() = default;
	sequence(std::string name, std::string nucl) : name_(std::move(name)), nucl_(std::move(nucl)) {}
	size_t size() const noexcept { return nucl_.size(); }
	const std::string &get_name() const noexcept { return name_; }
	const std::string &get_nucl() const noexcept { return nucl_; }
	const char *c_str() const noexcept { return nucl_.c_str(); }
	bool operator==(const sequence &o) const { return name_ == o.name_ && nucl_ == o.nucl_; }
};

class genome
{
	std::string name_;
	std::vector<sequence> contigs_;

  public:
	genome() = default;
	genome(std::string name, std::vector<sequence> contigs) : name_(std::move(name)), contigs_(std::move(contigs)) {}
	const std::string &get_name() const noexcept { return name_; }
	const std::vector<sequence> &get_contigs() const noexcept { return contigs_; }
};

std::string reverse(const std::string &);       // reverse complement, '!' kept (sequence.cxx:73-103)
std::string filter_nucl(const std::string &);   // keep ACGTacgt, upper-cased (sequence.cxx:109-146)
double gc_content(const std::string &) noexcept; // sequence.cxx:152-165
sequence join(const genome &);                   // contigs glued with '!' (sequence.cxx:171-199)

class evo_model
{
  public:
	uint64_t substitutions = 0, homologs = 0;
	evo_model() = default;
	evo_model(uint64_t s, uint64_t h) : substitutions(s), homologs(h) {}
	uint64_t total() const noexcept { return homologs; }
	double estimate_raw(bool zero_on_error = false) const noexcept;
	double estimate_JC(bool zero_on_error = false) const noexcept;
	double estimate_ani(bool zero_on_error = false) const noexcept;
	double coverage(size_t length) const noexcept { return (double)homologs / length; }
};

enum flags { none = 0, verbose = 1, extra_verbose = 2, complete_deletion = 4, print_positions = 16, dist_ani = 32, dist_raw = 64 };
extern int FLAGS;
extern int RETURN_CODE;
extern size_t reference_index;

// The seam: src/process.cxx:408-556, served by the CUDA library.  Exits via errx() on
// failure, like the reference does for fatal conditions.
std::vector<evo_model> process(const sequence &subject, const std::vector<sequence> &queries);

genome read_genome(const std::string &file_name);
void print_matrix(const std::vector<sequence> &queries, const std::vector<evo_model> &matrix);
