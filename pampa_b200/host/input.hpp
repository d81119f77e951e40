// input.hpp -- the plain-text input tokenizer and value readers (grammar of SURVEY.md App. B;
// reference behaviour: src/input.cxx:4-60 tokenizer, :62-300 array readers, :367-427 BC and
// convergence readers).
#pragma once

#include "util.hpp"

namespace pampa {
namespace input {

// Next non-empty, non-comment line split on single spaces (tabs and runs of spaces collapsed,
// lines starting with '#' skipped).  Empty vector at end of file.
std::vector<std::string> get_next_line(std::ifstream& file);

int PAMPA_WARN_UNUSED read(int& x, int x1, int x2, const std::string& s);
int PAMPA_WARN_UNUSED read(double& x, double x1, double x2, const std::string& s);
int PAMPA_WARN_UNUSED read(bool& q, const std::string& s);

// n values spread over as many lines as needed
int PAMPA_WARN_UNUSED read(std::vector<double>& v, unsigned n, double x1, double x2, std::ifstream& file);
int PAMPA_WARN_UNUSED read(std::vector<int>& v, unsigned n, int x1, int x2, std::ifstream& file);
// n rows of exactly m values
int PAMPA_WARN_UNUSED read(std::vector<double>& v, unsigned n, unsigned m, double x1, double x2,
                           std::ifstream& file);
// n rows of any length, nt values in total (ragged: row pointer + values)
int PAMPA_WARN_UNUSED read(std::vector<int>& ptr, std::vector<int>& v, unsigned n, unsigned nt, int x1,
                           int x2, std::ifstream& file);

// "<n>" followed by n values, or "-<n>" followed by one value repeated n times
int PAMPA_WARN_UNUSED read_axis(std::vector<double>& d, int& n, const std::string& count, std::ifstream& file);

int PAMPA_WARN_UNUSED read(BoundaryCondition& bc, const std::vector<std::string>& line, unsigned& i,
                           std::ifstream& file);

}   // namespace input
}   // namespace pampa
