#include "driver.hpp"

#include <chrono>
#include <cstring>

namespace pampa {

int Driver::initialize(int argc, char* argv[], std::vector<double>& dt) {
   PAMPA_CHECK(argc < 2, "missing input file");
   const std::string filename(argv[1]);
   output::verbose = output::silent = false;
   output::padding = 0;
   for (int a = 2; a < argc; a++) {
      if (!std::strcmp(argv[a], "-verbose")) output::verbose = true;
      if (!std::strcmp(argv[a], "-silent")) output::silent = true;
   }
   output::print("\nInitialize...");
   output::indent();
   Parser parser;
   output::print("Parse the input file...", true);
   PAMPA_CHECK(parser.read(filename, &mesh, materials, solvers, dt), "unable to parse " + filename);
   output::print("Done.", true);
   PAMPA_CHECK(!dt.empty(), "transient calculations are outside the scope of this build");
   output::print("Initialize the solver...", true);
   PAMPA_CHECK(solvers.empty(), "no solvers defined");
   solver = solvers[0];
   if (solvers.size() > 1) {
      solver = nullptr;
      for (Solver* s : solvers) if (s->name == "main") solver = s;
      PAMPA_CHECK(solver == nullptr, "unable to find the main solver");
   }
   PAMPA_CHECK(solver->initialize(false), "unable to initialize the solver");
   output::print("Done.", true);
   output::outdent();
   output::print("Done.");
   return 0;
}

int Driver::solve(int n, double dt, double t) {
   PAMPA_CHECK(solver == nullptr, "driver not initialised");
   output::print("\n--------------------------------");
   if (n == 0) output::print("\nSolve steady state...\n");
   else output::print("\nSolve time step " + std::to_string(n) + "...\n");
   const auto t1 = std::chrono::steady_clock::now();
   PAMPA_CHECK(solver->solve(n, dt, t), "unable to get the solution");
   const auto t2 = std::chrono::steady_clock::now();
   output::print("", true);
   output::print("Solution time", std::chrono::duration<double>(t2 - t1).count(), true, 3, true);
   PAMPA_CHECK(solver->output(".", n), "unable to output the solution");
   output::print("\nDone.");
   return 0;
}

int Driver::finalize() {
   output::print("\n--------------------------------");
   output::print("\nFinalize...");
   output::indent();
   output::print("Finalize the solver...", true);
   if (solver) PAMPA_CHECK(solver->finalize(), "unable to finalize the solver");
   output::print("Done.", true);
   delete mesh; mesh = nullptr;
   for (Material* m : materials) delete m;
   for (Solver* s : solvers) delete s;
   materials.clear(); solvers.clear(); solver = nullptr;
   output::outdent();
   output::print("Done.\n");
   return 0;
}

int Driver::getField(double* v, const std::string& name) const {
   PAMPA_CHECK(solver == nullptr, "driver not initialised");
   return solver->getField(v, name);
}

int Driver::setField(const double* v, const std::string& name) {
   PAMPA_CHECK(solver == nullptr, "driver not initialised");
   return solver->setField(v, name);
}

double Driver::getKeff() const {
   const NeutronicSolver* ns = dynamic_cast<const NeutronicSolver*>(solver);
   return ns ? ns->getKeff() : -1.0;
}

}   // namespace pampa
