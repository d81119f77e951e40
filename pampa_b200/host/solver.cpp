#include "solver.hpp"
#include "vtk.hpp"

#include <algorithm>
#include <cstring>
#include <map>

#include "../../include/pampa_sn.h"

namespace pampa {

// ------------------------------------------------------------------------------ Solver
int Solver::read(const std::string& filename, std::vector<Solver*>& solvers) {
   std::ifstream file(filename, std::ios_base::in);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   PAMPA_CHECK(read(file, solvers), "unable to read the solver from " + filename);
   return 0;
}

long Solver::getFieldSize(const std::string& name) const {
   for (const Field& f : fields) if (f.name == name) return (long)f.vec->size();
   return -1;
}

int Solver::getField(double* v, const std::string& name) const {
   for (const Field& f : fields)
      if (f.name == name) {
         std::copy(f.vec->begin(), f.vec->end(), v);
         return 0;
      }
   PAMPA_CHECK(true, "unable to find field '" + name + "'");
   return 0;
}

int Solver::setField(const double* v, const std::string& name) {
   for (Field& f : fields)
      if (f.name == name) {
         std::copy(v, v + f.vec->size(), f.vec->begin());
         return 0;
      }
   PAMPA_CHECK(true, "unable to find field '" + name + "'");
   return 0;
}

// ------------------------------------------------------------------------------ PhysicsSolver
int PhysicsSolver::initialize(bool transient) {
   PAMPA_CHECK(checkMaterials(transient), "wrong material data");
   PAMPA_CHECK(build(), "unable to build the solver");
   return 0;
}

int PhysicsSolver::output(const std::string& path, int n, bool write_mesh) const {
   PAMPA_CHECK(printLog(n), "unable to print the solution summary to standard output");
   if (write_mesh) PAMPA_CHECK(mesh->writeVTK(path + "/output", n), "unable to write the mesh in .vtk format");
   PAMPA_CHECK(writeVTK(path, n), "unable to write the solution in .vtk format");
   PAMPA_CHECK(writePETSc(n), "unable to write the solution in PETSc format");
   return 0;
}

// ------------------------------------------------------------------------------ NeutronicSolver
int NeutronicSolver::solve(int n, double dt, double t) {
   output::print("Run " + name + " solver...", true);
   output::indent(true);
   PAMPA_CHECK(buildMatrices(n, dt, t), "unable to build the coefficient matrices");
   PAMPA_CHECK(getSolution(n), "unable to solve the linear system and get the solution");
   output::outdent(true);
   output::print("Done.", true);
   return 0;
}

int NeutronicSolver::printLog(int n) const {
   if (n == 0) output::print("Effective multiplication factor", keff, false, 6);
   double total = 0.0;
   for (double x : q) total += x;
   output::print("Power", total, true, 3);
   return 0;
}

// ------------------------------------------------------------------------------ SNSolver
SNSolver::~SNSolver() {
   if (device) pampa_sn_destroy(device);
}

int SNSolver::read(std::ifstream& file, std::vector<Solver*>&) {
   while (true) {
      std::vector<std::string> line = input::get_next_line(file);
      if (line.empty() || line[0] == "}") break;
      const std::string& k = line[0];
      if (k == "bc") {
         PAMPA_CHECK(line.size() < 3, "wrong number of arguments for keyword '" + k + "'");
         const std::vector<std::string>& boundaries = mesh->getBoundaries();
         if (bcs.empty()) bcs.resize(1 + boundaries.size());
         const int ibc = mesh->findBoundary(line[1]);
         PAMPA_CHECK(ibc < 0, "wrong boundary name");
         unsigned l = 2;
         PAMPA_CHECK(input::read(bcs[ibc + 1], line, l, file), "wrong boundary condition");
         continue;
      }
      if (k == "convergence") {
         PAMPA_CHECK(line.size() != 5, "wrong number of arguments for keyword '" + k + "'");
         ConvergenceError* d = line[1] == "power" ? &dq : (line[1] == "production-rate" ? &dP : nullptr);
         PAMPA_CHECK(d == nullptr, "wrong field");
         d->norm = line[2] == "1" ? 1 : (line[2] == "2" ? 2 : (line[2] == "max" ? 0 : -1));
         PAMPA_CHECK(d->norm < 0, "wrong convergence norm type");
         PAMPA_CHECK(input::read(d->relative, line[3]), "wrong switch for relative convergence");
         PAMPA_CHECK(input::read(d->tol, 0.0, DBL_MAX, line[4]), "wrong convergence tolerance");
         continue;
      }
      PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
      if (k == "energy-groups") PAMPA_CHECK(input::read(num_energy_groups, 1, INT_MAX, line[1]), "wrong number of energy groups");
      else if (k == "power") PAMPA_CHECK(input::read(power, 0.0, DBL_MAX, line[1]), "wrong power level");
      else if (k == "order") PAMPA_CHECK(input::read(order, 1, INT_MAX, line[1]), "wrong SN order");
      else if (k == "mixed-face-interpolation")
         PAMPA_CHECK(input::read(face_interpolation_delta, 0.0, 1.0, line[1]), "wrong weight between upwind and linear interpolation");
      else if (k == "least-squares-boundary-interpolation")
         PAMPA_CHECK(input::read(boundary_interpolation_ls, line[1]), "wrong switch for least-squares boundary interpolation");
      // extensions of this implementation (not in the reference grammar)
      else if (k == "least-squares-mode") ls_mode = line[1];
      else if (k == "keff-tolerance") PAMPA_CHECK(input::read(tol_keff, 0.0, 1.0, line[1]), "wrong k-eff tolerance");
      else if (k == "flux-tolerance") PAMPA_CHECK(input::read(tol_flux, 0.0, 1.0, line[1]), "wrong flux tolerance");
      else if (k == "max-iterations") PAMPA_CHECK(input::read(max_iterations, 1, INT_MAX, line[1]), "wrong iteration limit");
      else PAMPA_CHECK(true, "unrecognized keyword '" + k + "'");
   }
   return 0;
}

int SNSolver::checkMaterials(bool transient) {
   for (const Material* mat : materials) {
      PAMPA_CHECK(!mat->hasNuclearData(), "missing nuclear data");
      PAMPA_CHECK(mat->checkNuclearData(num_energy_groups, transient), "wrong nuclear data");
   }
   return 0;
}

// Device material = (material, temperature) pair actually present in the mesh: a standalone run
// (T = 0 everywhere) maps one to one onto the input materials (src/FeedbackNuclearData.hxx:62-73).
int SNSolver::packCrossSections(std::vector<double>& st, std::vector<double>& ss, std::vector<double>& nsf,
                                std::vector<double>& ksf, std::vector<double>& chi, std::vector<double>& beta,
                                std::vector<int>& cell_material) const {
   const int G = num_energy_groups;
   std::map<std::pair<int, double>, int> ids;
   cell_material.assign(num_cells, 0);
   for (int i = 0; i < num_cells; i++) {
      const int m = cells.materials[i];
      PAMPA_CHECK(m < 0 || m >= (int)materials.size(), "wrong material index");
      auto key = std::make_pair(m, T[i]);
      auto it = ids.find(key);
      if (it == ids.end()) {
         it = ids.emplace(key, (int)ids.size()).first;
         const Material* mat = materials[m];
         for (int g = 0; g < G; g++) {
            st.push_back(mat->sigmaTotal(g, T[i]));
            nsf.push_back(mat->sigmaNuFission(g, T[i]));
            ksf.push_back(mat->sigmaKappaFission(g, T[i]));
            chi.push_back(mat->chiEffective(g, T[i]));
            for (int g2 = 0; g2 < G; g2++) ss.push_back(mat->sigmaScattering(g, g2, T[i]));
         }
         beta.push_back(mat->beta());
      }
      cell_material[i] = it->second;
   }
   return 0;
}

// Least-squares boundary gradient of the reference (src/SNSolver.cxx:212-269) and the coupling
// terms it adds to vacuum faces (:485-518).  Two bugs of the reference decide the 6th decimal of
// its 2-D goldens: the d matrix is read transposed, and G is never
// zeroed.  "literal" = the code as written with G zero-initialised; "reference-effective" = the
// closed form that reproduces the reference's printed 2-D results (G(1,1) effectively infinite).
int SNSolver::buildLSCorrection(std::vector<int>& cell, std::vector<int>& ptr, std::vector<int>& nbr,
                                std::vector<double>& omega, std::vector<double>& nvec) const {
   const int nd = mesh->getNumDimensions();
   std::string mode = ls_mode;
   if (mode == "auto") mode = nd == 2 ? "reference-effective" : "literal";
   PAMPA_CHECK(mode != "literal" && mode != "reference-effective", "wrong least-squares mode");
   PAMPA_CHECK(mesh->hasZFaces(), "least-squares boundary interpolation is only supported on 1-D and 2-D meshes");
   ptr.assign(1, 0);
   for (int i = 0; i < num_cells; i++) {
      const int f0 = faces.ptr[i], nf = faces.num_faces(i);
      bool boundary = false;
      for (int f = 0; f < nf; f++) boundary |= faces.neighbors[f0 + f] < 0;
      if (!boundary) continue;
      const double* ci = &cells.centroids[3 * (size_t)i];
      std::vector<double> v((size_t)nf * nd);                      // row-major d(f, id)
      for (int f = 0; f < nf; f++) {
         const int i2 = faces.neighbors[f0 + f];
         const double* c2 = i2 < 0 ? &faces.centroids[3 * (size_t)(f0 + f)] : &cells.centroids[3 * (size_t)i2];
         for (int id = 0; id < nd; id++) v[(size_t)f * nd + id] = c2[id] - ci[id];
      }
      std::vector<double> coef((size_t)nf * 3, 0.0);
      if (mode == "reference-effective" && nd > 1) {
         double g00 = 0.0;
         for (int f = 0; f < nf; f++) g00 += v[f] * v[(size_t)f * nd];
         for (int f = 0; f < nf; f++) coef[3 * (size_t)f] = v[f] / g00;
      } else {
         double Gm[9] = {0}, Gi[9] = {0};
         for (int jg = 0; jg < nd; jg++)
            for (int ig = 0; ig < nd; ig++)
               for (int f = 0; f < nf; f++) Gm[jg * nd + ig] += v[(size_t)jg * nd + f] * v[(size_t)f * nd + ig];
         if (nd == 1) Gi[0] = 1.0 / Gm[0];
         else {
            const double det = Gm[0] * Gm[3] - Gm[1] * Gm[2];
            Gi[0] = Gm[3] / det; Gi[1] = -Gm[1] / det; Gi[2] = -Gm[2] / det; Gi[3] = Gm[0] / det;
         }
         for (int f = 0; f < nf; f++)
            for (int id = 0; id < nd; id++)
               for (int jd = 0; jd < nd; jd++) coef[3 * (size_t)f + id] += Gi[id * nd + jd] * v[(size_t)jd * nd + f];
      }
      const size_t before = nbr.size();
      for (int f = 0; f < nf; f++) {
         const int i2 = faces.neighbors[f0 + f];
         if (i2 >= 0 || bcs[-i2].type != BC::VACUUM) continue;
         const double* xf = &faces.centroids[3 * (size_t)(f0 + f)];
         const double dp[3] = {xf[0] - ci[0], xf[1] - ci[1], xf[2] - ci[2]};
         for (int f2 = 0; f2 < nf; f2++) {
            const int i3 = faces.neighbors[f0 + f2];
            if (i3 < 0) continue;
            nbr.push_back(i3);
            omega.push_back(dp[0] * coef[3 * (size_t)f2] + dp[1] * coef[3 * (size_t)f2 + 1] + dp[2] * coef[3 * (size_t)f2 + 2]);
            for (int c = 0; c < 3; c++)
               nvec.push_back(faces.normals[3 * (size_t)(f0 + f) + c] * faces.areas[f0 + f] / cells.volumes[i]);
         }
      }
      if (nbr.size() > before) { cell.push_back(i); ptr.push_back((int)nbr.size()); }
   }
   return 0;
}

int SNSolver::build() {
   PAMPA_CHECK(num_energy_groups < 1, "missing number of energy groups");
   // mixed-face-interpolation < 1 (the reference's default is 0.1, src/SNSolver.hxx:16) couples every face both
   // ways: the device layer sweeps the upwind part and treats the rest as a deferred correction (pampa_sn.h)
   PAMPA_CHECK(face_interpolation_delta <= 0.0,
               "mixed-face-interpolation 0 (pure linear interpolation) has no upwind part to sweep: use a value in (0, 1]");
   if (bcs.empty()) bcs = mesh->getBoundaryConditions();
   quadrature = AngularQuadratureSet(order);
   PAMPA_CHECK(quadrature.build(), "unable to build the angular quadrature set");
   num_directions = quadrature.getNumDirections();

   // fields, in the reference's order and layouts (src/SNSolver.cxx:721-747)
   // (build() also runs when a temperature update needs a new device plan: the input fields are kept then)
   if ((int)T.size() != num_cells) T.assign(num_cells, 0.0);
   if ((int)S.size() != num_cells) S.assign(num_cells, 0.0);
   phi.assign((size_t)num_cells * num_energy_groups, 0.0);
   q.assign(num_cells, 0.0);
   P.assign(num_cells, 0.0);
   fields.clear();
   fields.push_back(Field{"temperature", &T, true, false, nullptr});
   fields.push_back(Field{"delayed-source", &S, true, false, nullptr});
   fields.push_back(Field{"scalar-flux", &phi, false, false, nullptr});
   fields.push_back(Field{"angular-flux", &psi, false, false, nullptr});
   fields.push_back(Field{"power", &q, false, true, &dq});
   fields.push_back(Field{"production-rate", &P, false, true, &dP});

   // extruded description of the mesh for the device layer
   const int nxy = mesh->getNumXYCells(), nz = mesh->getNumLayers();
   PAMPA_CHECK(nxy * nz != num_cells, "the mesh is not an extruded mesh");
   int F = 0;
   std::vector<int> nlat(nxy, 0);
   for (int i = 0; i < nxy; i++) {
      for (int f = 0; f < faces.num_faces(i); f++)
         if (std::fabs(faces.normals[3 * (size_t)(faces.ptr[i] + f) + 2]) < 0.5) nlat[i]++;
      F = std::max(F, nlat[i]);
   }
   const double h0 = mesh->hasZFaces() ? mesh->getDz()[0] : 1.0;
   std::vector<int> nb((size_t)nxy * F, 0);
   std::vector<double> fx((size_t)nxy * F, 0.0), fy((size_t)nxy * F, 0.0), cf((size_t)nxy * F, 1.0);
   std::vector<double> kout((size_t)nxy * F, 0.0), kin((size_t)nxy * F, 0.0);   // delta < 1 (src/SNSolver.cxx:193-198)
   std::vector<double> area(nxy), cx(nxy), cy(nxy);
   auto dist = [](const double* a, const double* b) {
      return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
   };
   for (int i = 0; i < nxy; i++) {
      const double* ci = &cells.centroids[3 * (size_t)i];
      area[i] = cells.volumes[i] / h0; cx[i] = ci[0]; cy[i] = ci[1];
      int a = 0;
      for (int f = 0; f < faces.num_faces(i); f++) {
         const size_t gf = (size_t)faces.ptr[i] + f;
         if (std::fabs(faces.normals[3 * gf + 2]) >= 0.5) continue;
         const size_t s = (size_t)i * F + a++;
         const int i2 = faces.neighbors[gf];
         nb[s] = i2;
         fx[s] = faces.normals[3 * gf] * faces.areas[gf] / h0;
         fy[s] = faces.normals[3 * gf + 1] * faces.areas[gf] / h0;
         if (i2 >= 0) {          // upwind face weight for delta = 1 (src/SNSolver.cxx:188-198)
            const double* xf = &faces.centroids[3 * gf];
            const double* c2 = &cells.centroids[3 * (size_t)i2];
            cf[s] = (dist(xf, ci) + dist(xf, c2)) / dist(ci, c2);
            kout[s] = (1.0 - face_interpolation_delta) * dist(xf, ci) / dist(ci, c2);
            kin[s] = (1.0 - face_interpolation_delta) * dist(xf, c2) / dist(ci, c2);
         }
      }
   }
   std::vector<int> bc_types(bcs.size(), PAMPA_SN_BC_NONE);
   for (size_t b = 1; b < bcs.size(); b++)
      bc_types[b] = bcs[b].type == BC::VACUUM ? PAMPA_SN_BC_VACUUM
                    : (bcs[b].type == BC::REFLECTIVE ? PAMPA_SN_BC_REFLECTIVE : PAMPA_SN_BC_NONE);

   std::vector<double> st, ss, nsf, ksf, chi, beta;
   std::vector<int> cell_material;
   PAMPA_CHECK(packCrossSections(st, ss, nsf, ksf, chi, beta, cell_material), "unable to pack the cross sections");

   pampa_sn_mesh ms;
   std::memset(&ms, 0, sizeof(ms));
   ms.num_xy_cells = nxy; ms.num_layers = nz; ms.has_z_faces = mesh->hasZFaces() ? 1 : 0; ms.max_xy_faces = F;
   ms.xy_num_faces = nlat.data(); ms.xy_neighbor = nb.data(); ms.xy_face_fx = fx.data(); ms.xy_face_fy = fy.data();
   ms.xy_face_cf = cf.data(); ms.xy_area = area.data(); ms.xy_cx = cx.data(); ms.xy_cy = cy.data();
   ms.xy_ij = mesh->getXYij().empty() ? nullptr : mesh->getXYij().data();
   ms.dz = mesh->hasZFaces() ? mesh->getDz().data() : nullptr;
   ms.materials = cell_material.data();
   ms.bc_minus_z = mesh->hasZFaces() ? mesh->findBoundary("-z") + 1 : 0;
   ms.bc_plus_z = mesh->hasZFaces() ? mesh->findBoundary("+z") + 1 : 0;
   ms.num_bcs = (int)bc_types.size() - 1; ms.bc_types = bc_types.data();
   ms.face_interpolation_delta = face_interpolation_delta;
   ms.xy_face_kout = kout.data(); ms.xy_face_kin = kin.data();

   pampa_sn_xs xs;
   xs.num_materials = (int)beta.size(); xs.num_groups = num_energy_groups;
   xs.sigma_total = st.data(); xs.sigma_scattering = ss.data(); xs.nu_sigma_fission = nsf.data();
   xs.kappa_sigma_fission = ksf.data(); xs.chi_effective = chi.data(); xs.beta_total = beta.data();

   pampa_sn_quadrature qd;
   qd.num_directions = num_directions; qd.directions = quadrature.getDirections().data();
   qd.weights = quadrature.getWeights().data(); qd.reflected = quadrature.getReflectedDirections().data();

   std::vector<int> ls_cell, ls_ptr, ls_nbr;
   std::vector<double> ls_omega, ls_nvec;
   pampa_sn_ls ls;
   std::memset(&ls, 0, sizeof(ls));
   if (boundary_interpolation_ls) {
      PAMPA_CHECK(buildLSCorrection(ls_cell, ls_ptr, ls_nbr, ls_omega, ls_nvec),
                  "unable to build the least-squares gradient discretization for boundary cells");
      ls.num_cells = (int)ls_cell.size(); ls.cell = ls_cell.data(); ls.ptr = ls_ptr.data(); ls.nbr = ls_nbr.data();
      ls.omega = ls_omega.data(); ls.nvec = ls_nvec.data();
   }

   pampa_sn_options opts;
   pampa_sn_default_options(&opts);
   if (const char* dev = std::getenv("PAMPA_SN_DEVICE")) opts.device = std::atoi(dev);
   if (device) { pampa_sn_destroy(device); device = nullptr; }
   PAMPA_CHECK(pampa_sn_create(&device, &ms, &xs, &qd, ls.num_cells > 0 ? &ls : nullptr, &opts),
               std::string("unable to create the device transport solver: ") + pampa_sn_last_error(nullptr));
   return 0;
}

int SNSolver::buildMatrices(int n, double, double) {
   PAMPA_CHECK(n > 0, "the transient SN branch is not implemented");
   if (!xs_dirty) return 0;
   std::vector<double> st, ss, nsf, ksf, chi, beta;
   std::vector<int> cell_material;
   PAMPA_CHECK(packCrossSections(st, ss, nsf, ksf, chi, beta, cell_material), "unable to pack the cross sections");
   pampa_sn_xs xs;
   xs.num_materials = (int)beta.size(); xs.num_groups = num_energy_groups;
   xs.sigma_total = st.data(); xs.sigma_scattering = ss.data(); xs.nu_sigma_fission = nsf.data();
   xs.kappa_sigma_fission = ksf.data(); xs.chi_effective = chi.data(); xs.beta_total = beta.data();
   // the temperature field changes which (material, temperature) row each cell uses, and possibly how many rows
   // there are: new tables and a new cell -> row map; if the new table does not fit the device plan, re-create the
   // device problem (build() keeps the temperature and delayed-source fields)
   if (pampa_sn_update_materials(device, &xs, cell_material.data()))
      PAMPA_CHECK(build(), "unable to rebuild the device transport solver");
   xs_dirty = false;
   return 0;
}

int SNSolver::getSolution(int n) {
   PAMPA_CHECK(n > 0, "the transient SN branch is not implemented");
   int its = 0;
   PAMPA_CHECK(pampa_sn_solve_keff(device, tol_keff, tol_flux, max_iterations, power, &keff, &its),
               std::string("unable to solve the eigensystem: ") + pampa_sn_last_error(device));
   num_iterations = its;
   PAMPA_CHECK(pampa_sn_get(device, "scalar-flux", phi.data()), pampa_sn_last_error(device));
   PAMPA_CHECK(pampa_sn_get(device, "power", q.data()), pampa_sn_last_error(device));
   PAMPA_CHECK(pampa_sn_get(device, "production-rate", P.data()), pampa_sn_last_error(device));
   PAMPA_CHECK(pampa_sn_get(device, "delayed-source", S.data()), pampa_sn_last_error(device));
   psi.clear();                                          // fetched on demand (cells x groups x directions)
   for (double x : phi) PAMPA_CHECK(x < 0.0, "negative values in the scalar-flux solution");
   if (face_interpolation_delta < 1.0 || boundary_interpolation_ls) {
      // SNSolver::normalizeAngularFlux fails the solve on a negative angular flux (src/SNSolver.cxx:329); only the
      // non-monotone terms (linear face interpolation, LS boundary gradient) can produce one
      double psi_min = 0.0;
      PAMPA_CHECK(pampa_sn_get(device, "angular-flux-min", &psi_min), pampa_sn_last_error(device));
      PAMPA_CHECK(psi_min < 0.0, "negative values in the angular-flux solution");
   }
   power = 0.0;
   for (double x : q) power += x;
   return 0;
}

int SNSolver::getField(double* v, const std::string& fname) const {
   if (fname == "angular-flux") {
      PAMPA_CHECK(device == nullptr, "solver not initialised");
      PAMPA_CHECK(pampa_sn_get(device, "angular-flux", v), pampa_sn_last_error(device));
      return 0;
   }
   return Solver::getField(v, fname);
}

int SNSolver::setField(const double* v, const std::string& fname) {
   PAMPA_CHECK(Solver::setField(v, fname), "unable to set field '" + fname + "'");
   if (fname == "temperature") xs_dirty = true;        // cross sections are re-tabulated on the next solve
   return 0;
}

long SNSolver::getFieldSize(const std::string& fname) const {
   if (fname == "angular-flux") return (long)num_cells * num_energy_groups * num_directions;
   return Solver::getFieldSize(fname);
}

// scalar flux, angular flux and thermal power appended to the mesh file (src/SNSolver.cxx:754-770); the angular
// flux (cells x groups x directions) is fetched from the device only when .vtk output is switched on
int SNSolver::writeVTK(const std::string& path, int n) const {
   if (!vtk::on || (n % vtk::dn != 0)) return 0;
   PAMPA_CHECK(vtk::write(path + "/output", n, "flux", phi.data(), num_cells, num_energy_groups),
               "unable to write the scalar flux");
   std::vector<double> angular((size_t)num_cells * num_energy_groups * num_directions);
   PAMPA_CHECK(getField(angular.data(), "angular-flux"), "unable to get the angular flux");
   PAMPA_CHECK(vtk::write(path + "/output", n, "flux", angular.data(), num_cells, num_energy_groups, num_directions),
               "unable to write the angular flux");
   PAMPA_CHECK(vtk::write(path + "/output", n, "power", q.data(), num_cells), "unable to write the thermal power");
   return 0;
}

// angular flux in PETSc's binary Vec format (src/SNSolver.cxx:773-780), fetched only when `petsc dump 1` is set
int SNSolver::writePETSc(int n) const {
   if (!ptc::dump) return 0;
   std::vector<double> angular((size_t)num_cells * num_energy_groups * num_directions);
   PAMPA_CHECK(getField(angular.data(), "angular-flux"), "unable to get the angular flux");
   PAMPA_CHECK(ptc::write("angular_flux", n, angular.data(), (long)angular.size()), "unable to write the angular flux");
   return 0;
}

int SNSolver::finalize() {
   if (device) { PAMPA_CHECK(pampa_sn_destroy(device), "unable to destroy the device solver"); device = nullptr; }
   return 0;
}

}   // namespace pampa
