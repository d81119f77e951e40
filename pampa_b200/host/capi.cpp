// capi.cpp -- the reference's C API (src/pampa.cxx:8-113) over a global Driver.
#include <cstdio>
#include "vtk.hpp"
#include <cstdlib>

#include "../../include/pampa.h"
#include "driver.hpp"

extern "C" {

static pampa::Driver pampa_driver;

void pampa_initialize(int argc, char* argv[], double** dt, int* ndt, int* error) {
   std::vector<double> dt0;
   *error = pampa_driver.initialize(argc, argv, dt0);
   if (*error > 0) { printf("Error in pampa.initialize().\n"); return; }
   *ndt = (int)dt0.size();
   *dt = new double[*ndt > 0 ? *ndt : 1];
   for (int i = 0; i < *ndt; i++) (*dt)[i] = dt0[i];
}

void pampa_initialize_steady_state(int argc, char* argv[], int* error) {
   std::vector<double> dt0;
   *error = pampa_driver.initialize(argc, argv, dt0);
   if (*error > 0) { printf("Error in pampa.initialize().\n"); return; }
}

void pampa_solve(int n, double dt, double t, int* error) {
   *error = pampa_driver.solve(n, dt, t);
   if (*error > 0) { printf("Error in pampa.solve().\n"); return; }
}

void pampa_solve_steady_state(int* error) {
   pampa_solve(0, 0.0, 0.0, error);
   if (*error > 0) { printf("Error in pampa_solve().\n"); return; }
}

void pampa_finalize(double** dt, int* error) {
   *error = pampa_driver.finalize();
   if (*error > 0) { printf("Error in pampa.finalize().\n"); return; }
   delete[] *dt;
}

void pampa_finalize_steady_state(int* error) {
   *error = pampa_driver.finalize();
   if (*error > 0) { printf("Error in pampa.finalize().\n"); return; }
}

void pampa_get_field(double* v, const char name[], int* error) {
   *error = pampa_driver.getField(v, std::string(name));
   if (*error > 0) { printf("Error in pampa.getField().\n"); return; }
}

void pampa_set_field(const double* v, const char name[], int* error) {
   *error = pampa_driver.setField(v, std::string(name));
   if (*error > 0) { printf("Error in pampa.setField().\n"); return; }
}

long pampa_get_field_size(const char name[], int* error) {
   const long n = pampa_driver.getFieldSize(std::string(name));
   *error = n < 0 ? 1 : 0;
   return n;
}

/* Host-only digest of an input deck (no GPU): parses it with the same Parser / mesh / material
 * code the solver uses and returns sums the CPU test-suite compares with the oracle's parse of the
 * same deck.  out[16]: cells, dims, sum V, faces, sum A, boundary faces, sum of interior neighbour
 * indices, weighted sums of cell centroids / face centroids / normals, sum of cell materials,
 * sum sigma_t, weighted sum sigma_s, sum nu-sigma-f, sum kappa-sigma-f, sum chi_eff. */
int pampa_debug_describe(const char* deck, double* out) {
   pampa::Mesh* mesh = nullptr;
   std::vector<pampa::Material*> materials;
   std::vector<pampa::Solver*> solvers;
   std::vector<double> dt;
   pampa::Parser parser;
   int rc = parser.read(std::string(deck), &mesh, materials, solvers, dt);
   if (!rc && mesh) {
      const pampa::Cells& c = mesh->getCells();
      const pampa::Faces& f = mesh->getFaces();
      for (int i = 0; i < 16; i++) out[i] = 0.0;
      out[0] = mesh->getNumCells(); out[1] = mesh->getNumDimensions();
      for (double v : c.volumes) out[2] += v;
      out[3] = (double)f.areas.size();
      for (double a : f.areas) out[4] += a;
      for (int nb : f.neighbors) { if (nb < 0) out[5] += 1.0; else out[6] += nb; }
      for (size_t i = 0; i < c.centroids.size(); i++) out[7] += (1 + i % 3) * c.centroids[i];
      for (size_t i = 0; i < f.centroids.size(); i++) out[8] += (1 + i % 3) * f.centroids[i];
      for (size_t i = 0; i < f.normals.size(); i++) out[9] += (1 + i % 3) * f.normals[i];
      for (int m : c.materials) out[10] += m;
      for (const pampa::Material* mat : materials) {
         const int G = mat->numEnergyGroups();
         for (int g = 0; g < G; g++) {
            out[11] += mat->sigmaTotal(g, 0.0); out[13] += mat->sigmaNuFission(g, 0.0);
            out[14] += mat->sigmaKappaFission(g, 0.0); out[15] += mat->chiEffective(g, 0.0);
            for (int g2 = 0; g2 < G; g2++) out[12] += (1 + g + 2 * g2) * mat->sigmaScattering(g, g2, 0.0);
         }
      }
   }
   delete mesh;
   for (auto* m : materials) delete m;
   for (auto* s : solvers) delete s;
   return rc;
}

/* Test hook (no GPU needed): parse a deck and write its mesh to <prefix>.vtk in the reference's format
 * (src/vtk.cxx:28-121), whatever the deck's own `vtk` switch says. */
int pampa_debug_write_mesh_vtk(const char* deck, const char* prefix) {
   pampa::Mesh* mesh = nullptr;
   std::vector<pampa::Material*> materials;
   std::vector<pampa::Solver*> solvers;
   std::vector<double> dt;
   pampa::Parser parser;
   int rc = parser.read(std::string(deck), &mesh, materials, solvers, dt);
   if (!rc && mesh) {
      const bool on = pampa::vtk::on;
      pampa::vtk::on = true;
      rc = mesh->writeVTK(std::string(prefix), -1);
      pampa::vtk::on = on;
   }
   delete mesh;
   for (auto* m : materials) delete m;
   for (auto* s : solvers) delete s;
   return rc;
}

/* Test hook (no GPU needed): parse a deck and write every array of its mesh in the reference's plain-text format
 * (Mesh::writeData, src/Mesh.cxx:408-569) -- the file `mesh partitioned <file>` reads back. */
int pampa_debug_write_mesh_data(const char* deck, const char* filename, int digits) {
   pampa::Mesh* mesh = nullptr;
   std::vector<pampa::Material*> materials;
   std::vector<pampa::Solver*> solvers;
   std::vector<double> dt;
   pampa::Parser parser;
   int rc = parser.read(std::string(deck), &mesh, materials, solvers, dt);
   if (!rc && mesh) rc = mesh->writeData(std::string(filename), digits);
   delete mesh;
   for (auto* m : materials) delete m;
   for (auto* s : solvers) delete s;
   return rc;
}

/* Test hook: write `count` doubles as <prefix>_<n>.ptc whatever the `petsc dump` switch says. */
int pampa_debug_write_ptc(const char* prefix, int n, const double* v, long count) {
   const bool on = pampa::ptc::dump;
   pampa::ptc::dump = true;
   const int rc = pampa::ptc::write(std::string(prefix), n, v, count);
   pampa::ptc::dump = on;
   return rc;
}

double pampa_get_keff(int* error) {
   const double k = pampa_driver.getKeff();
   *error = k < 0.0 ? 1 : 0;
   return k;
}

}
