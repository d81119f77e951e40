/* pampa_sn.h -- C ABI of the B200 (sm_100a) discrete-ordinates transport layer.
 *
 * This is the drop-in boundary for pampa's SN k-eigenvalue hot path.  The reference has no
 * such seam: SNSolver assembles a monolithic PETSc matrix and calls SLEPc.  The entry points
 * below are what a PETSc-free SNSolver binds instead; each one names the reference code it
 * replaces (paths relative to the reference tree).
 *
 *   pampa_sn_create        SNSolver::build                      src/SNSolver.cxx:687-751
 *                          + buildGaussGradientScheme           src/SNSolver.cxx:159-208 (via cf)
 *                          + AngularQuadratureSet tables         src/AngularQuadratureSet.cxx:4-209
 *   pampa_sn_update_xs     the XS reads of buildMatrices        src/SNSolver.cxx:383-439
 *   pampa_sn_update_materials   ... when T_data changes them   src/SNSolver.cxx:383-439, src/FeedbackNuclearData.hxx:62-140
 *   pampa_sn_source        scattering + fission rows of R / F   src/SNSolver.cxx:417-439
 *   pampa_sn_sweep         R^-1 applied by EPSSolve (LU solve)  src/petsc.cxx:428-433
 *   pampa_sn_reduce        calculateScalarFlux + production     src/SNSolver.cxx:272-299,
 *                                                               src/NeutronicSolver.cxx:81-116
 *   pampa_sn_solve_keff    NeutronicSolver::solve / getSolution src/NeutronicSolver.cxx:4-43,
 *                                                               src/SNSolver.cxx:634-657
 *   pampa_sn_get / _set    Solver::getField / setField          src/Solver.cxx:18-43
 *   pampa_sn_destroy       PhysicsSolver::finalize              src/PhysicsSolver.cxx:38-65
 *
 * Conventions (reference: src/utils.hxx:41-47): every function returns 0 on success and 1 on
 * error; the message is retrievable with pampa_sn_last_error().  Host arrays are borrowed for
 * the duration of the call only.  All reals are fp64, all indices int32.  A handle is not
 * thread-safe.  There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef PAMPA_SN_H
#define PAMPA_SN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pampa_sn_handle pampa_sn_handle;

enum { PAMPA_SN_BC_NONE = 0, PAMPA_SN_BC_VACUUM = 1, PAMPA_SN_BC_REFLECTIVE = 2 };

/* Extruded finite-volume mesh: a 2-D polygon mesh times num_layers prisms.  Cell index
 * i = k*num_xy_cells + ixy, the ordering of both CartesianMesh (src/CartesianMesh.cxx:146-414,
 * void cells removed) and UnstructuredExtrudedMesh (src/UnstructuredExtrudedMesh.cxx:153-364).
 * 1-D and 2-D meshes have num_layers = 1 and has_z_faces = 0. */
typedef struct {
   int32_t num_xy_cells;
   int32_t num_layers;
   int32_t has_z_faces;          /* 1: cells have -z/+z faces (3-D mesh) */
   int32_t max_xy_faces;         /* row stride of the face tables below */
   const int32_t* xy_num_faces;  /* [num_xy_cells] lateral faces of each xy cell */
   const int32_t* xy_neighbor;   /* [num_xy_cells*max_xy_faces] >= 0: xy cell, < 0: -(1-based bc) */
   const double*  xy_face_fx;    /* [..] outward normal x lateral face length (area / dz) */
   const double*  xy_face_fy;
   const double*  xy_face_cf;    /* [..] upwind face weight (r_if + r_i2f)/r_ii2, 1 on boundaries
                                    (src/SNSolver.cxx:193-198 with delta = 1) */
   const double*  xy_area;       /* [num_xy_cells] base area (volume / dz) */
   const double*  xy_cx;         /* [num_xy_cells] centroid, used only to build sweep patches */
   const double*  xy_cy;
   const int32_t* xy_ij;         /* optional [num_xy_cells*2] structured (i,j) of Cartesian meshes */
   const double*  dz;            /* [num_layers] (ignored when has_z_faces = 0) */
   const int32_t* materials;     /* [num_layers*num_xy_cells] 0-based */
   int32_t bc_minus_z;           /* 1-based bc index of the -z / +z boundaries */
   int32_t bc_plus_z;
   int32_t num_bcs;
   const int32_t* bc_types;      /* [1 + num_bcs], entry 0 unused; PAMPA_SN_BC_* */
   /* mixed-face-interpolation delta < 1 (src/SNSolver.hxx:16 default 0.1, weights src/SNSolver.cxx:193-198):
    * the face flux is a blend of the upwind value and the linear interpolation between the two cell centres.
    * The sweep inverts the delta = 1 (upwind) part; the remainder
    *    (T_delta - T_1) psi |_i = sum_f |Omega.n_f| A_f / V_i * kappa_f * (psi_nbr - psi_i),
    *    kappa_f = (1-delta) r_if / r_ii2 on outgoing faces, (1-delta) r_i2f / r_ii2 on incoming faces,
    * is applied to the angular flux of the previous sweep and moved to the right-hand side (a deferred
    * correction that converges with the source iteration to the eigenpair of the reference's matrix).
    * NULL / 1.0: pure upwind.  The z faces use the same formula with the layer thicknesses. */
   const double*  xy_face_kout;  /* [num_xy_cells*max_xy_faces] (1-delta) r_if / r_ii2, 0 on boundary faces */
   const double*  xy_face_kin;   /* [num_xy_cells*max_xy_faces] (1-delta) r_i2f / r_ii2, 0 on boundary faces */
   double face_interpolation_delta;   /* delta in (0, 1]; 0 is read as 1 (structs zeroed by older callers) */
} pampa_sn_mesh;

/* Multigroup cross sections per material (src/Material.hxx:130-169). */
typedef struct {
   int32_t num_materials;
   int32_t num_groups;
   const double* sigma_total;        /* [mat][g] */
   const double* sigma_scattering;   /* [mat][g_from][g_to]  (src/ConstantNuclearData.hxx:57) */
   const double* nu_sigma_fission;   /* [mat][g] */
   const double* kappa_sigma_fission;/* [mat][g] */
   const double* chi_effective;      /* [mat][g] */
   const double* beta_total;         /* [mat] (delayed source S_i = beta * P_i) */
} pampa_sn_xs;

/* Angular quadrature (src/AngularQuadratureSet.hxx:44-53). */
typedef struct {
   int32_t num_directions;
   const double*  directions;    /* [m][3] */
   const double*  weights;       /* [m], sum = 1 */
   const int32_t* reflected;     /* [m][3] mirror of m about the x, y, z planes */
} pampa_sn_quadrature;

/* Lagged least-squares boundary correction (src/SNSolver.cxx:485-518), CSR over boundary
 * cells; only meshes without z faces.  The correction of vacuum face f of cell i for
 * direction m is omega * max(0, Omega_m . nvec) * (psi_nbr - psi_i), nvec = n_f A_f / V_i. */
typedef struct {
   int32_t num_cells;            /* boundary cells carrying a correction (0 = none) */
   const int32_t* cell;          /* [num_cells] cell index */
   const int32_t* ptr;           /* [num_cells+1] */
   const int32_t* nbr;           /* [nnz] neighbouring cell i3 */
   const double*  omega;         /* [nnz] (x_f - x_i) . c_bc(i, f2) */
   const double*  nvec;          /* [nnz][3] */
} pampa_sn_ls;

typedef struct {
   int32_t device;               /* CUDA device ordinal */
   int32_t store_psi;            /* 1: keep the angular flux in HBM (needed for angular-flux) */
   int32_t patch_cells;          /* sweep patch size = CTA size (0: default 256) */
   int32_t tile_i, tile_j;       /* Cartesian tile shape when xy_ij is given (0: default 16x16) */
   int32_t z_chunk;              /* layers per sweep task (0: automatic) */
   int32_t rank, num_ranks;      /* angle/group sharding; (0,1) for one GPU */
   int32_t shard_mode;           /* 0: shard sweep chunks (angle sets), 1: shard energy groups */
   int32_t verbose;
   int32_t dt_max;               /* directions swept together per CTA, 1..10 (0: default) */
   int32_t generic_only;         /* 1: never use the staged tile kernel (A/B testing) */
   int32_t single_stream;        /* 1: launch all ordering classes on one stream (A/B testing) */
   int32_t anderson_depth;       /* Anderson acceleration of the k-eff iteration: history depth 1..7
                                    (0: default 7, < 0: plain power iteration) */
   int32_t wave_launch;          /* 1: sweep the tile classes with one launch per wavefront instead of the
                                    single dataflow launch (A/B testing) */
   int32_t group_merge;          /* energy groups a dataflow sweep task handles back to back (0: default 8);
                                    more groups = less pipeline fill/drain padding in the step-major arrays */
   int32_t inline_edges;         /* 1: structured tiles numbered perimeter first and the dataflow sweep reads the
                                    perimeter lanes of the neighbouring patch's psi rows directly instead of
                                    compact edge copies behind each row (6 GB less written per sweep at C4, but
                                    no faster: measured 22.35 vs 21.90 ms per iteration); 0: default */
   int32_t no_graph;             /* 1: never replay the sweep's launch sequence from a CUDA graph (plans with >= 32
                                    launches per sweep are captured once and replayed; A/B testing) */
   int32_t partition_fields;     /* sharded runs, 1: pampa_sn_get / _set / _field_size of the cell fields work on this
                                    rank's contiguous range of cells only -- the local-length vectors the reference
                                    hands its MPI ranks (src/Solver.cxx:18-43) -- so that every rank moves 1/N of the
                                    host bytes: cells [rank*c, min(N, (rank+1)*c)), c = ceil(N / num_ranks);
                                    set("flux-moments") completes the iterate with an allgather on the device */
} pampa_sn_options;

void pampa_sn_default_options(pampa_sn_options* opts);

int pampa_sn_create(pampa_sn_handle** h, const pampa_sn_mesh* mesh, const pampa_sn_xs* xs,
                    const pampa_sn_quadrature* quad, const pampa_sn_ls* ls,
                    const pampa_sn_options* opts);
int pampa_sn_destroy(pampa_sn_handle* h);
const char* pampa_sn_last_error(const pampa_sn_handle* h);   /* h may be NULL: create errors */

int pampa_sn_update_xs(pampa_sn_handle* h, const pampa_sn_xs* xs);
/* Temperature feedback (src/FeedbackNuclearData.hxx:62-140 through SNSolver::buildMatrices): a new table of
 * (material, temperature) rows AND a new cell -> row map [num_layers*num_xy_cells]; the number of rows may differ
 * from the one the handle was created with.  Returns 1 (handle unchanged but for the error text) when the new
 * table does not fit the handle's sweep plan; the caller then destroys and re-creates it. */
int pampa_sn_update_materials(pampa_sn_handle* h, const pampa_sn_xs* xs, const int32_t* materials);

/* One source iteration, split the way the three kernels are: q <- (S + F/keff) phi;
 * psi <- T^-1 q (all owned angle sets and groups); phi <- sum_m w_m psi and the production /
 * power integrals.  pampa_sn_reduce also rotates phi_new into phi. */
int pampa_sn_source(pampa_sn_handle* h, double keff);
int pampa_sn_sweep(pampa_sn_handle* h);
int pampa_sn_reduce(pampa_sn_handle* h, double* production, double* power, double* dphi_rel);

/* k-eigenvalue solve: source iterations (Anderson-accelerated by default) until |dk| < tol_k and the
 * relative L2 change of the flux moments over one source iteration is < tol_phi (or max_it). */
int pampa_sn_solve_keff(pampa_sn_handle* h, double tol_k, double tol_phi, int32_t max_it,
                        double power, double* keff, int32_t* iterations);

/* Run `iterations` source iterations with no convergence test (benchmark / warm-up). */
int pampa_sn_iterate(pampa_sn_handle* h, int32_t iterations, double* keff);

/* Same, timed on the device with CUDA events on the launching stream: total_ms spans the whole
 * call, sweep_ms is the sum over iterations of the sweep-kernel launches alone. */
int pampa_sn_iterate_timed(pampa_sn_handle* h, int32_t iterations, double* keff, double* total_ms,
                           double* sweep_ms);

/* Fields in the reference layouts (src/SNSolver.cxx:721-747):
 *   "scalar-flux" [i][g], "angular-flux" [i][g][m], "power" [i], "production-rate" [i],
 *   "delayed-source" [i]  (get);   "temperature" [i], "delayed-source" [i] (set, stored only);
 *   "flux-moments" [i][g] (get/set): the raw iteration state sum_m w_m psi, un-normalised;
 *   "keff" [1] (get/set): the current eigenvalue estimate / the one the next iteration starts from.
 *   "angular-flux-min" [1] (get): smallest stored angular-flux value (the reference fails the solve when it is
 *   negative, src/SNSolver.cxx:329; possible only with delta < 1 or the least-squares boundary term).
 * Normalised as the reference does after the eigen-solve (src/NeutronicSolver.cxx:46-78,
 * src/SNSolver.cxx:302-341) by the last pampa_sn_solve_keff. */
int pampa_sn_get(pampa_sn_handle* h, const char* name, double* out);
int pampa_sn_set(pampa_sn_handle* h, const char* name, const double* in);
int64_t pampa_sn_field_size(const pampa_sn_handle* h, const char* name);

/* Sharded runs: the flux-moment exchange.  The caller owns the transport (NCCL communicator
 * created from an id broadcast by the host code); these give it the device buffers. */
int pampa_sn_comm_init(pampa_sn_handle* h, const void* nccl_unique_id, int32_t id_bytes);
int pampa_sn_comm_unique_id(void* nccl_unique_id, int32_t id_bytes);
void* pampa_sn_device_ptr(pampa_sn_handle* h, const char* name, int64_t* count);

/* Introspection for tests and the benchmark. */
typedef struct {
   int64_t num_cells, num_groups, num_directions;
   int64_t updates_per_sweep;        /* owned cell*angle*group updates per pampa_sn_sweep */
   int64_t sweep_launches;           /* kernel launches per pampa_sn_sweep */
   int64_t sweep_tasks;              /* CTAs per pampa_sn_sweep */
   int64_t num_classes, num_chunks;
   int64_t tile_classes;             /* ordering classes swept on 2-D tiles (vs level chunks) */
   int64_t device_bytes;
   double  last_sweep_ms, last_source_ms, last_reduce_ms;
   int64_t kernel_launches;          /* total launches issued by this handle */
   double  timed_kernel_ms;          /* last pampa_sn_iterate_timed: device time of the sweep-kernel launches
                                        alone (between the shear and un-shear passes), summed over iterations */
   int64_t num_tilings;              /* shared patch partitions (1: structured tiles / k-d leaves; a lattice of
                                        congruent cells has one per pair of lattice directions) */
   int64_t flow_classes;             /* ordering classes the one-launch dataflow kernel can sweep */
   int64_t lattice;                  /* 1: unstructured mesh recognised as a lattice of congruent cells */
   double  last_solve_ms;            /* device time of the last pampa_sn_solve_keff (events on the launching stream) */
   /* last pampa_sn_iterate_timed, summed over its iterations: the source kernel; everything after the sweep
    * (reduction, scalar collective, k update, flux-moment exchange); the flux-moment exchange alone */
   double  timed_source_ms, timed_reduce_ms, timed_exchange_ms;
} pampa_sn_info;
int pampa_sn_get_info(pampa_sn_handle* h, pampa_sn_info* info);

/* Build and validate the sweep plan on the host only (no GPU needed): patch coverage, acyclic
 * patch order, upwind sources.  Used by the CPU test-suite. */
int pampa_sn_plan_check(const pampa_sn_mesh* mesh, const pampa_sn_quadrature* quad,
                        int32_t num_groups, const pampa_sn_options* opts, pampa_sn_info* info);

#ifdef __cplusplus
}
#endif
#endif
