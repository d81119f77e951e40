/* pampa.h -- the reference's own C API (src/pampa.hxx:1-31), implemented by libpampa.so of this
 * repository on top of the B200 transport layer (include/pampa_sn.h).  Same eight entry points,
 * same argument meaning and error convention (*error = 0 on success, > 0 on failure with a
 * message on stdout), so the reference's C / C++ / Fortran drivers (src/c/main.c:6-34,
 * src/cxx/main.cxx:4-28, src/f90/main.f90:1-81) link against it unchanged. */
#ifndef PAMPA_H
#define PAMPA_H
#ifdef __cplusplus
extern "C" {
#endif

void pampa_initialize(int argc, char* argv[], double** dt, int* ndt, int* error);
void pampa_initialize_steady_state(int argc, char* argv[], int* error);
void pampa_solve(int n, double dt, double t, int* error);
void pampa_solve_steady_state(int* error);
void pampa_finalize(double** dt, int* error);
void pampa_finalize_steady_state(int* error);
void pampa_get_field(double* v, const char name[], int* error);
void pampa_set_field(const double* v, const char name[], int* error);

/* Additions of this implementation (not in the reference): sizes and the solver's k-eff, so that a
 * host code can allocate field buffers without knowing the mesh. */
long pampa_get_field_size(const char name[], int* error);
double pampa_get_keff(int* error);
/* Host-only entry points used by the CPU test-suite (no device needed): a digest of a parsed deck, the mesh of
 * a deck in the reference's .vtk format (src/Mesh.cxx writeVTK), and a vector in PETSc's binary Vec format
 * (src/petsc.cxx:491-511). */
int pampa_debug_describe(const char* deck, double* out16);
int pampa_debug_write_mesh_vtk(const char* deck, const char* prefix);
int pampa_debug_write_ptc(const char* prefix, int n, const double* v, long count);
int pampa_debug_write_mesh_data(const char* deck, const char* filename, int digits);   /* src/Mesh.cxx:408-569 */

#ifdef __cplusplus
}
#endif
#endif
