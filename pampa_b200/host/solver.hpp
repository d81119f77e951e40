// solver.hpp -- the plugin interface the SN hot path sits behind, mirrored from the reference:
//   Field / Solver            src/Solver.hxx:14-98      (lifecycle virtuals, field registry)
//   PhysicsSolver             src/PhysicsSolver.hxx:7-60 (initialize = checkMaterials + build, output)
//   NeutronicSolver           src/NeutronicSolver.hxx:6-75 (solve, k-eff log)
//   SNSolver                  src/SNSolver.hxx:7-87
// The PETSc Vec behind a Field becomes a plain host buffer; the monolithic matrix assembly and the
// SLEPc eigen-solve (src/SNSolver.cxx:344-657) become calls into the C-ABI CUDA layer
// (include/pampa_sn.h).  There is no CPU fallback: without that library / a GPU, solve() fails.
#pragma once

#include "material.hpp"
#include "mesh.hpp"
#include "quadrature.hpp"

struct pampa_sn_handle;

namespace pampa {

struct ConvergenceError {
   std::string name;
   int norm = 2;          // 1, 2 or 0 (= max)
   bool relative = true;
   double tol = 1.0;
};

struct Field {
   std::string name;
   std::vector<double>* vec = nullptr;    // host copy, reference layout
   bool input = false, output = false;
   ConvergenceError* delta = nullptr;
};

class Solver {
  public:
   const std::string name;
   Solver(const std::string& name, const Mesh* mesh)
       : name(name), mesh(mesh), num_cells(mesh->getNumCells()), num_cells_global(mesh->getNumCellsGlobal()),
         num_faces_max(mesh->getNumFacesMax()), cells(mesh->getCells()), faces(mesh->getFaces()) {}
   virtual ~Solver() {}
   std::vector<Field>& getFields() { return fields; }
   int PAMPA_WARN_UNUSED read(const std::string& filename, std::vector<Solver*>& solvers);
   virtual int PAMPA_WARN_UNUSED read(std::ifstream& file, std::vector<Solver*>& solvers) = 0;
   virtual int PAMPA_WARN_UNUSED initialize(bool transient = false) = 0;
   virtual int PAMPA_WARN_UNUSED solve(int n = 0, double dt = 0.0, double t = 0.0) = 0;
   virtual int PAMPA_WARN_UNUSED output(const std::string& path, int n = 0, bool write_mesh = true) const = 0;
   virtual int PAMPA_WARN_UNUSED finalize() = 0;
   virtual int PAMPA_WARN_UNUSED getField(double* v, const std::string& name) const;
   virtual int PAMPA_WARN_UNUSED setField(const double* v, const std::string& name);
   virtual long getFieldSize(const std::string& name) const;

  protected:
   const Mesh* mesh;
   const int num_cells, num_cells_global, num_faces_max;
   const Cells& cells;
   const Faces& faces;
   std::vector<Field> fields;
};

class PhysicsSolver : public Solver {
  public:
   PhysicsSolver(const std::string& name, const Mesh* mesh, const std::vector<Material*>& materials)
       : Solver(name, mesh), materials(materials) {}
   int PAMPA_WARN_UNUSED initialize(bool transient = false) override;
   int PAMPA_WARN_UNUSED output(const std::string& path, int n = 0, bool write_mesh = true) const override;

  protected:
   const std::vector<Material*>& materials;
   virtual int PAMPA_WARN_UNUSED checkMaterials(bool transient = false) = 0;
   virtual int PAMPA_WARN_UNUSED build() = 0;
   virtual int PAMPA_WARN_UNUSED printLog(int n = 0) const = 0;
   virtual int PAMPA_WARN_UNUSED writeVTK(const std::string& path, int n = 0) const = 0;
   virtual int PAMPA_WARN_UNUSED writePETSc(int n = 0) const = 0;
};

class NeutronicSolver : public PhysicsSolver {
  public:
   NeutronicSolver(const std::string& name, const Mesh* mesh, const std::vector<Material*>& materials)
       : PhysicsSolver(name, mesh, materials) {}
   int PAMPA_WARN_UNUSED solve(int n = 0, double dt = 0.0, double t = 0.0) override;
   double getKeff() const { return keff; }

  protected:
   int num_energy_groups = -1;
   std::vector<BoundaryCondition> bcs;                 // 1-based
   double power = 1.0;
   double keff = -1.0;
   std::vector<double> T, S, phi, q, P;                // temperature, delayed source, flux, power, production
   ConvergenceError dq{"power", 2, true, 1.0}, dP{"production-rate", 2, true, 1.0};
   // the two hooks of the reference: "build the matrices" = refresh the device cross sections,
   // "get the solution" = run the device eigen-solve and fetch the fields
   virtual int PAMPA_WARN_UNUSED buildMatrices(int n, double dt, double t) = 0;
   virtual int PAMPA_WARN_UNUSED getSolution(int n = 0) = 0;
   int PAMPA_WARN_UNUSED printLog(int n = 0) const override;
};

class SNSolver : public NeutronicSolver {
  public:
   SNSolver(const Mesh* mesh, const std::vector<Material*>& materials) : NeutronicSolver("sn", mesh, materials) {}
   ~SNSolver() override;
   int PAMPA_WARN_UNUSED read(std::ifstream& file, std::vector<Solver*>& solvers) override;
   int PAMPA_WARN_UNUSED finalize() override;
   int PAMPA_WARN_UNUSED getField(double* v, const std::string& name) const override;
   int PAMPA_WARN_UNUSED setField(const double* v, const std::string& name) override;
   long getFieldSize(const std::string& name) const override;
   long numDirections() const { return num_directions; }
   int iterations() const { return num_iterations; }

  private:
   int order = -1;
   double face_interpolation_delta = 0.1;              // reference default (src/SNSolver.hxx:16)
   bool boundary_interpolation_ls = false;
   std::string ls_mode = "auto";                       // auto | literal | reference-effective
   double tol_keff = 1.0e-10, tol_flux = 1.0e-9;       // tighter than SLEPc's 1e-8 eigen-residual
   int max_iterations = 100000;
   int num_directions = -1, num_iterations = 0;
   AngularQuadratureSet quadrature;
   std::vector<double> psi;                            // fetched lazily (large)
   pampa_sn_handle* device = nullptr;
   bool xs_dirty = false;

   int PAMPA_WARN_UNUSED checkMaterials(bool transient = false) override;
   int PAMPA_WARN_UNUSED build() override;
   int PAMPA_WARN_UNUSED buildMatrices(int n, double dt, double t) override;
   int PAMPA_WARN_UNUSED getSolution(int n = 0) override;
   int PAMPA_WARN_UNUSED writeVTK(const std::string& path, int n = 0) const override;
   int PAMPA_WARN_UNUSED writePETSc(int n = 0) const override;
   int PAMPA_WARN_UNUSED packCrossSections(std::vector<double>& st, std::vector<double>& ss, std::vector<double>& nsf,
                                           std::vector<double>& ksf, std::vector<double>& chi, std::vector<double>& beta,
                                           std::vector<int>& cell_material) const;
   int PAMPA_WARN_UNUSED buildLSCorrection(std::vector<int>& cell, std::vector<int>& ptr, std::vector<int>& nbr,
                                           std::vector<double>& omega, std::vector<double>& nvec) const;
};

}   // namespace pampa
