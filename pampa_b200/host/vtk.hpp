// vtk.hpp -- plain-text .vtk output of the host code, in the reference's format (src/vtk.cxx:4-170):
// the mesh (points, cells in gmsh point order, VTK cell types, materials) written by Mesh::writeVTK, then one
// SCALARS block per field / group / direction appended by the solver.  Switched by the `vtk <0|1> [interval]`
// line of the main input file (src/Parser.cxx:203-217).
#pragma once

#include "util.hpp"

namespace pampa {
namespace vtk {

extern bool on;      // switch for .vtk output
extern int dn;       // output interval in time steps

constexpr int PRECISION = 6;   // src/utils.hxx:32

// mesh: `cell_ptr` / `cell_points` = ragged point lists of the cells; materials 0-based (written 1-based)
int PAMPA_WARN_UNUSED write(const std::string& prefix, int n, const std::vector<double>& points, int num_points,
                            const std::vector<int>& cell_ptr, const std::vector<int>& cell_points, int num_cells,
                            const std::vector<int>& materials);
// solution vector in the reference layout v[(i*num_groups + g)*num_directions + m]
int PAMPA_WARN_UNUSED write(const std::string& prefix, int n, const std::string& name, const double* v,
                            int num_cells, int num_groups = 1, int num_directions = 1);

}   // namespace vtk

// `petsc dump 1` (src/Parser.cxx:231-234): solution vectors in PETSc's binary Vec format, <prefix>_<n>.ptc
// (src/petsc.cxx:491-511 -> VecView on a binary viewer): big-endian int32 class id 1211214, int32 length, then
// the values as big-endian float64 -- readable by PetscBinaryIO / VecLoad.  No PETSc is involved here.
namespace ptc {
extern bool dump;
int PAMPA_WARN_UNUSED write(const std::string& prefix, int n, const double* v, long count);
}   // namespace ptc
}   // namespace pampa
