// util.hpp -- error convention, boundary conditions, output helpers of the host code.
//
// Error convention of the reference (src/utils.hxx:41-47): every function returns 0 on success and
// 1 on error, printing "file:line:function(): error: <message>." on the way up.
#pragma once

#include <cfloat>
#include <climits>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#define PAMPA_WARN_UNUSED __attribute__((warn_unused_result))

#define PAMPA_CHECK(condition, message) \
   do { \
      if (condition) { \
         std::cout << __FILE__ << ":" << __LINE__ << ":" << __FUNCTION__ << "(): " \
                   << "error: " << message << "." << std::endl; \
         return 1; \
      } \
   } while (0)

namespace pampa {

constexpr double DBL_TOL = 1.0e-6;     // src/utils.hxx:35

namespace BC {
enum Type { NONE, VACUUM, REFLECTIVE, ROBIN, DIRICHLET, ADIABATIC, CONVECTION };   // src/utils.hxx:113
}

struct BoundaryCondition {
   BC::Type type = BC::NONE;
   std::vector<double> parameters;     // Robin / Dirichlet / convection values (unused by SN)
};

// rank-0 style printing with indentation and info / verbose / silent levels (src/output.cxx:31-95)
namespace output {
extern bool verbose, silent;
extern int padding;
void print(const std::string& message, bool info = false);
void print(const std::string& name, double x, bool scientific, int precision, bool info = false);
void indent(bool info = false);
void outdent(bool info = false);
}   // namespace output

}   // namespace pampa
