// parser.hpp -- main input file (reference: src/Parser.cxx:4-260; grammar in SURVEY.md App. B).
#pragma once

#include "solver.hpp"

namespace pampa {

class Parser {
  public:
   int PAMPA_WARN_UNUSED read(const std::string& filename, Mesh** mesh, std::vector<Material*>& materials,
                              std::vector<Solver*>& solvers, std::vector<double>& dt);
};

}   // namespace pampa
