// driver.hpp -- calculation driver (reference: src/Driver.hxx:13-58, src/Driver.cxx:4-127).
#pragma once

#include "parser.hpp"

namespace pampa {

class Driver {
  public:
   int PAMPA_WARN_UNUSED initialize(int argc, char* argv[], std::vector<double>& dt);
   int PAMPA_WARN_UNUSED solve(int n = 0, double dt = 0.0, double t = 0.0);
   int PAMPA_WARN_UNUSED finalize();
   int PAMPA_WARN_UNUSED getField(double* v, const std::string& name) const;
   int PAMPA_WARN_UNUSED setField(const double* v, const std::string& name);
   long getFieldSize(const std::string& name) const { return solver ? solver->getFieldSize(name) : -1; }
   double getKeff() const;

  private:
   Mesh* mesh = nullptr;
   std::vector<Material*> materials;
   std::vector<Solver*> solvers;
   Solver* solver = nullptr;
};

}   // namespace pampa
