#include "input.hpp"

namespace pampa {

namespace output {
bool verbose = false, silent = false;
int padding = 0;

void print(const std::string& message, bool info) {
   if ((!info || verbose) && !silent) std::cout << std::string(3 * padding, ' ') << message << std::endl;
}

void print(const std::string& name, double x, bool scientific, int precision, bool info) {
   if ((!info || verbose) && !silent) {
      std::cout << (scientific ? std::scientific : std::fixed) << std::setprecision(precision)
                << std::string(3 * padding, ' ') << name << ": " << x << "." << std::endl;
   }
}

void indent(bool info) { if ((!info || verbose) && !silent) padding++; }
void outdent(bool info) { if ((!info || verbose) && !silent) padding--; }
}   // namespace output

namespace input {

static void clean(std::string& s) {
   for (char& c : s) if (c == '\t') c = ' ';
   std::string out;
   out.reserve(s.size());
   for (char c : s) if (!(c == ' ' && !out.empty() && out.back() == ' ')) out.push_back(c);
   size_t a = out.find_first_not_of(' ');
   if (a == std::string::npos) { s.clear(); return; }
   size_t b = out.find_last_not_of(' ');
   s = out.substr(a, b - a + 1);
   if (s[0] == '#') s.clear();
}

std::vector<std::string> get_next_line(std::ifstream& file) {
   std::string line;
   std::vector<std::string> words;
   while (std::getline(file, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      clean(line);
      if (line.empty()) continue;
      std::istringstream iss(line);
      std::string w;
      while (std::getline(iss, w, ' ')) words.push_back(w);
      break;
   }
   return words;
}

int read(int& x, int x1, int x2, const std::string& s) {
   try { x = std::stoi(s); } catch (...) { PAMPA_CHECK(true, "wrong int value '" + s + "'"); }
   PAMPA_CHECK(x < x1 || x > x2, "out-of-bounds int value");
   return 0;
}

int read(double& x, double x1, double x2, const std::string& s) {
   try { x = std::stod(s); } catch (...) { PAMPA_CHECK(true, "wrong double value '" + s + "'"); }
   PAMPA_CHECK(x < x1 || x > x2, "out-of-bounds double value");
   return 0;
}

int read(bool& q, const std::string& s) {
   int x = -1;
   PAMPA_CHECK(read(x, 0, 1, s), "wrong bool value");
   q = (x == 1);
   return 0;
}

template <typename T>
static int read_flat(std::vector<T>& v, unsigned n, T x1, T x2, std::ifstream& file) {
   v.assign(n, T());
   unsigned l = 0;
   while (l < n) {
      std::vector<std::string> line = get_next_line(file);
      PAMPA_CHECK(line.empty(), "missing data");
      for (const std::string& w : line) {
         PAMPA_CHECK(l >= n, "out-of-bounds data");
         PAMPA_CHECK(read(v[l++], x1, x2, w), "wrong data");
      }
   }
   return 0;
}

int read(std::vector<double>& v, unsigned n, double x1, double x2, std::ifstream& file) {
   return read_flat<double>(v, n, x1, x2, file);
}

int read(std::vector<int>& v, unsigned n, int x1, int x2, std::ifstream& file) {
   return read_flat<int>(v, n, x1, x2, file);
}

int read(std::vector<double>& v, unsigned n, unsigned m, double x1, double x2, std::ifstream& file) {
   v.assign((size_t)n * m, 0.0);
   for (unsigned r = 0; r < n; r++) {
      std::vector<std::string> line = get_next_line(file);
      PAMPA_CHECK(line.size() < m, "missing data");
      PAMPA_CHECK(line.size() > m, "out-of-bounds data");
      for (unsigned c = 0; c < m; c++) PAMPA_CHECK(read(v[(size_t)r * m + c], x1, x2, line[c]), "wrong data");
   }
   return 0;
}

int read(std::vector<int>& ptr, std::vector<int>& v, unsigned n, unsigned nt, int x1, int x2,
         std::ifstream& file) {
   ptr.assign(1, 0);
   v.clear();
   v.reserve(nt);
   for (unsigned r = 0; r < n; r++) {
      std::vector<std::string> line = get_next_line(file);
      PAMPA_CHECK(line.empty(), "missing data");
      PAMPA_CHECK(v.size() + line.size() > nt, "out-of-bounds data");
      for (const std::string& w : line) {
         int x;
         PAMPA_CHECK(read(x, x1, x2, w), "wrong data");
         v.push_back(x);
      }
      ptr.push_back((int)v.size());
   }
   return 0;
}

int read_axis(std::vector<double>& d, int& n, const std::string& count, std::ifstream& file) {
   PAMPA_CHECK(read(n, -INT_MAX, INT_MAX, count), "wrong number of intervals");
   if (n > 0) {
      PAMPA_CHECK(read(d, n, 0.0, DBL_MAX, file), "wrong interval data");
   } else {
      PAMPA_CHECK(read(d, 1, 0.0, DBL_MAX, file), "wrong interval data");
      n = -n;
      d.assign(n, d[0]);
   }
   return 0;
}

int read(BoundaryCondition& bc, const std::vector<std::string>& line, unsigned& i, std::ifstream&) {
   const std::string& t = line[i++];
   if (t == "vacuum") bc.type = BC::VACUUM;
   else if (t == "reflective") bc.type = BC::REFLECTIVE;
   else if (t == "robin") bc.type = BC::ROBIN;
   else if (t == "dirichlet") bc.type = BC::DIRICHLET;
   else if (t == "adiabatic") bc.type = BC::ADIABATIC;
   else if (t == "convection") bc.type = BC::CONVECTION;
   else PAMPA_CHECK(true, "wrong boundary-condition type");
   // parameters of the non-neutronic types are kept as plain numbers (time functions are not
   // needed by the SN path)
   bc.parameters.clear();
   while (i < line.size() && line[i] != "{") {
      double x;
      PAMPA_CHECK(read(x, -DBL_MAX, DBL_MAX, line[i++]), "wrong boundary-condition parameter");
      bc.parameters.push_back(x);
   }
   return 0;
}

}   // namespace input
}   // namespace pampa
