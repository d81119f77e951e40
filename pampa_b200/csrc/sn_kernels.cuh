// sn_kernels.cuh -- device-side structures and kernel launchers of the SN transport layer.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "sn_plan.hpp"

namespace pampa_sn {

// Per ordering class, device pointers into the uploaded plan.
constexpr int PS = 256;       // patch slots = sweep CTA size
constexpr int PEDGE = 32;     // compact copies of the lanes other patches read, behind each psi row
constexpr int PSX = PS + PEDGE;
constexpr int PEER_MAX = 7;   // other ranks of a node whose iterate buffers a kernel stores into (8 GPUs)

// Step-major ("sheared") arrays.  The pipeline step of (lane, layer) is kp + lvl(lane), kp = position of
// the layer in sweep order: the lanes of a CTA work on different layers in the same step (wavefront skew),
// and storing by step instead of by layer makes every CTA-step one contiguous block.  The owned groups are
// stored in blocks of gm consecutive groups that share ONE run of rows per patch: group gi of a block
// starts nz rows after group gi - 1, so that the skewed tail of one group interleaves with the skewed head
// of the next and a task that sweeps the gm groups back to back (flow kernel) has no padding in between:
//    rows per (block, patch):  nsm = gm * nz + max local levels - 1
//    row of (gl, patch, step): ((gl / gm) * npatch + patch) * nsm + (gl % gm) * nz + step
// psi of a chunk: [block][patch][row][direction][lane | edge copies] (row = PSX doubles per direction).
__host__ __device__ inline int64_t block_row0(int gl, int patch, int npatch, int nsm, int gm, int nz) {
   return ((int64_t)(gl / gm) * npatch + patch) * nsm + (int64_t)(gl % gm) * nz;
}
__host__ __device__ inline int64_t psi_index(int gl, int64_t slot, int step, int d, int npatch,
                                             int nsm, int gm, int nz, int nd, int pstride) {
   return ((block_row0(gl, (int)(slot >> 8), npatch, nsm, gm, nz) + step) * nd + d) * pstride + (slot & (PS - 1));
}

struct ClassDev {
   int64_t S;                 // slots of this class (npatch * PS)
   int32_t zdir;              // +1 / -1 / 0
   int32_t ring;              // smem ring depth
   int32_t tiles;             // 1: class slot == base slot
   int32_t npatch;
   int32_t nsteps;            // pipeline steps of one group per patch: max local levels + nz - 1
   int32_t gm;                // owned groups per block of the step-major arrays
   int32_t nsm;               // rows per (block, patch) = gm * nz + max local levels - 1
   int32_t mat_bytes;         // element size of mats_c: 1 (uint8, <= 256 materials) or 4 (int32)
   int32_t inline_edges;      // 1: dataflow kernel, neighbouring patches read the first PERIM_MAX lanes of a psi
   int32_t pstride;           //    row directly (perimeter-first lane order) instead of edge copies;
                              // pstride: doubles per direction of a psi row (PSX with edge copies, PS inline)
   const int32_t* mats_s;     // [npatch][nsteps][PS] material of (lane, step), -1 outside
   const uint8_t* mats_c;     // [npatch][nz][PS] x mat_bytes, cyclic: row r holds the material of layer
                              // (r - lvl) mod nz (dataflow kernel; 0 in holes)
   const int32_t* cell_of;    // [S] base slot or -1
   const uint16_t* lvl;       // [S]
   const int32_t* patch_nlev; // [npatch]
   const double2* out_vec;    // [S]
   const int32_t* in_src;     // [FIN_MAX][S]
   const double2* in_vec;     // [FIN_MAX][S]
   const int32_t* rout;       // [ROUT_MAX][S]
   const int32_t* ls_of;      // [S] index into the LS cell list or -1 (nullptr: no LS)
   const uint16_t* in_hidx;   // [FIN_MAX][S] halo index of patch-boundary / reflective sources
   const uint8_t* eidx;       // [S] edge index of lanes read by other patches (255: none)
   double* q_sheared;         // [block][patch][nsm][PS] source in this class's step-major order
                              // (nullptr: class swept by the generic kernel)
};

// Per chunk of directions swept together by one CTA.
struct ChunkDev {
   int32_t cls;
   int32_t nd;
   int32_t m[DT_MAX];         // quadrature index
   int32_t mrefl[DT_MAX][3];  // mirrored direction about x, y, z
   double mux[DT_MAX], muy[DT_MAX], muz_abs[DT_MAX], w[DT_MAX];
   double* psi;               // [block][patch][nsm][nd][PSX]
   double* phi_part;          // [block][patch][nsm][PS] sum_d w_d psi_d of this chunk (tile / flow kernels)
   int32_t flow_slot;         // index of this chunk in the direction table of its flow launch
   int32_t pad_;
};

struct SweepGlobals {
   const ClassDev* classes;
   const ChunkDev* chunks;
   const int32_t* gloc;       // [G] local index of an owned group or -1
   const int32_t* gown;       // [Gown] group of a local index
   const double* q;           // [G][nz][Sb]
   double* phi_new;           // [G][nz][Sb] accumulated with atomics
   const int32_t* mats;       // [nz][Sb] (-1 in holes)
   const double* sigma_t;     // [mat][G]
   const double* inv_dz;      // [nz]
   // reflective boundary buffers (old: read, new: written)
   const double* bnd_old;     // [M][G][nz][nrf]
   double* bnd_new;
   const double* bndz_old;    // [2][M][G][Sb]  (0: -z face, 1: +z face)
   double* bndz_new;
   // least-squares lagged correction (2-D / 1-D meshes only)
   const double* ls_dD;       // [M][nls]
   const double* ls_rhs;      // [M][G][nls]
   int64_t Sb;
   int32_t G, Gown, M, nz, Kc, has_z, nrf, nls;
   int32_t gm;                // owned groups per block (ClassDev::gm, the same for every class)
   int32_t np_stride;         // patches per (chunk, block) in the progress-counter array: the largest patch count of
                              // any shared tiling (classes on different tilings have different patch counts)
   int32_t bcz_minus_refl, bcz_plus_refl;   // 1 if that z boundary is reflective
   int32_t store_psi;
   int32_t nmat;
   int32_t uniform_dz;        // 1: every layer has the same thickness (or the mesh has no z faces)
   int64_t corr_off;          // delta < 1: offset (doubles) from a psi entry to its deferred-correction entry (0: none)
   int32_t dbg;               // PAMPA_SN_DBG: (dbg >> 4) & 0xff = rows between two publishes of a dataflow task's
                              // progress counter (0: default 16; the stress test uses 1); no other bits are read
};

struct ReduceScalars {        // device-resident iteration state
   double keff;
   double production;         // sum V nu-sigma-f phi of the current phi
   double power;              // sum V kappa-sigma-f phi
   double dphi2, phi2;        // ||phi_new - phi||^2, ||phi_new||^2 of the last reduce
   double dk;                 // keff change of the last reduce
   double min_phi;            // min over cells/groups of phi
   double pad;
};

void launch_sweep(const SweepGlobals& gp, const Task* d_tasks, int ntasks, int dt, int fin,
                  int ring, bool extras, cudaStream_t st);
cudaError_t configure_sweep_kernels();
cudaError_t configure_tile_kernels();
cudaError_t configure_shear_kernels();
int shear_max_classes();      // fast classes per z direction the shear kernel handles
constexpr int SHEAR_MAX_PER_PASS = 32;   // fast classes / chunks per z direction the shear kernels handle
void launch_sweep_tile(const SweepGlobals& gp, const Task* d_tasks, int ntasks, int dt, bool extras,
                       cudaStream_t st);
// dataflow version of the tile kernel: one launch, tasks taken by ticket in topological order,
// patch-to-patch dependencies through progress counters (see sn_kernels.cu)
int launch_sweep_flow(const SweepGlobals& gp, const Task* d_tasks, int ntasks, int dt, int fin, bool extras, int* ticket,
                      int* progress, const double* mw_host, int nch, cudaStream_t st);   // 1: direction table overflow
int flow3_max_dt();           // widest chunk of the three-incoming-face variant
int flow_max_chunks(int dt);
cudaError_t configure_flow_kernels();
// base [g][k][slot] <-> step-major [g][patch][step][lane] transforms for the tile kernel
// (per shared tiling: npatch and cell_of are the tiling's; cell_of = nullptr for the base tiling)
void launch_shear_q(const SweepGlobals& gp, const ClassDev* d_classes, const int32_t* d_fast_classes,
                    int nfast, int npatch, const int32_t* cell_of, cudaStream_t st);
void launch_unshear_phi(const SweepGlobals& gp, const ChunkDev* d_chunks, const ClassDev* d_classes,
                        const int32_t* d_fast_chunks, int nfast, int nplus, int npatch, int overwrite_first,
                        const int32_t* cell_of, int zsplit, cudaStream_t st);
// slabs per column (grid y) of the un-shear passes for a rank that owns `active_columns` (patch, group) columns
constexpr int UNSHEAR_ZSPLIT_MAX = 8;
int unshear_zsplit(int active_columns, int nz, int num_sms);

// un-shear with the reduction pass and the delivery of the flux moments fused into its last sweep over a column
void launch_unshear_phi_fused(const SweepGlobals& gp, const ChunkDev* d_chunks, const ClassDev* d_classes,
                              const int32_t* d_fast_chunks, int nfast, int nplus, int npatch, int overwrite_first,
                              const double* phi_old, double* phi_out, double* const* peer_out, int npeers,
                              const int32_t* mats, const double* nusf, const double* kapsf, const double* area,
                              const double* dz, int has_z, double* partials, double* sums, int zsplit, cudaStream_t st);

void launch_source(const double* phi, double* q, const int32_t* mats, const double* sig_s,
                   const double* chi, const double* nusf, const ReduceScalars* sc, const int32_t* gloc,
                   int G, int nmat, int nz, int64_t Sb, cudaStream_t st);

void launch_reduce(double* phi, double* phi_new, const int32_t* mats, const double* nusf,
                   const double* kapsf, const double* area, const double* dz, int has_z, int G,
                   int nz, int64_t Sb, const int32_t* gloc, int owned_only, int rotate, double* partials,
                   int nblocks, double* sums, cudaStream_t st);
// group-sharded runs with peer access: reduction of the owned groups + delivery of the new flux moments into the
// other iterate buffer of this rank (phi_out) and of every peer (peer_out[0 .. npeers)), see sn_kernels.cu
void launch_reduce_push(const double* phi, double* phi_new, double* phi_out, double* const* peer_out, int npeers,
                        const int32_t* mats, const double* nusf, const double* kapsf, const double* area,
                        const double* dz, int has_z, int G, int nz, int64_t Sb, const int32_t* gloc, double* partials,
                        int nblocks, double* sums, cudaStream_t st);
// Anderson acceleration (history of up to 8 iterates, see sn_api.cu: pampa_sn_solve_keff)
constexpr int AA_SLOTS = 8;
// The whole bookkeeping of the accelerated iteration lives on the device, so that pampa_sn_solve_keff enqueues
// iteration after iteration without waiting for any of them (the host reads `converged` one iteration late):
// Gram matrix of the stored residuals, eigenvalue estimates, window ages, mixing weights.
struct AAState {
   double M[AA_SLOTS][AA_SLOTS];   // Gram matrix of the residuals in the history slots
   double kg[AA_SLOTS];            // k estimate that came with each slot
   double mix[AA_SLOTS];           // weights of the next iterate
   double kn;                      // k of the current iterate
   double prod_x;                  // production the iterates are normalised to
   double inv;                     // normalisation of the last sweep result (prod_x / its production)
   double best, res, dk;           // smallest residual so far, last residual, last k change
   double power_integral, min_phi, phi2;
   double tol_k, tol_phi;
   double ncells;                  // flux entries (cells x groups): scale of the negative-flux test
   int32_t age[AA_SLOTS];          // iteration that filled each slot (-1: empty)
   int32_t cur, slots, it, nvisit;
   int32_t converged, failed, aa_start;
   int32_t hold;                   // plain steps left after a sweep result with negative flux (see sn_aa_solve_kernel)
   int32_t negatives, pad_[3];     // such events so far
};
// (every launcher below takes the device-resident AAState: slot, window and weights are read on the device)
void launch_aa_begin(AAState* st_dev, const double* sums, cudaStream_t st);
void launch_aa_store(const double* phi, double* phi_new, const int32_t* gloc, int owned_only, int zero_new, int G, int64_t n,
                     double* const* hist_f, double* const* hist_g, const AAState* st_dev, double* partials,
                     int nblocks, double* dots, cudaStream_t st);
void launch_aa_solve(AAState* st_dev, const double* dots, ReduceScalars* sc, cudaStream_t st);
void launch_scale_copy_slot(double* const* hist, const double* src, const AAState* st_dev, int64_t n, cudaStream_t st);
void launch_vec_mix(double* out, double* const* hist, const AAState* st_dev, int64_t n, cudaStream_t st);
void launch_aa_mix(double* phi, const int32_t* mats, const int32_t* gloc, int owned_only, int G, int64_t n,
                   double* const* hist_g, const AAState* st_dev, int nblocks, double* const* peer_out, int npeers,
                   cudaStream_t st);   // peer_out: the same iterate buffer of the other ranks (nullptr: none)
void launch_update_k(const double* sums, ReduceScalars* sc, int update_k, cudaStream_t st);
void launch_combine_sums(const double* all, int nranks, double* sums, cudaStream_t st);   // [rank][5] -> [5]

void launch_ls_rhs(const SweepGlobals& gp, const int32_t* ls_ptr, const int32_t* ls_nbr_slot,
                   const double* ls_coef, int64_t nnz, const int32_t* dir_chunk,
                   const int32_t* dir_d, const int32_t* const* class_pos_of, cudaStream_t st);

// delta < 1: deferred correction (T_delta - T_1) psi of one chunk into `corr` (layout of the chunk's psi block)
void launch_delta_corr(const SweepGlobals& gp, int chunk, int npatch, const int32_t* pos_of, const int32_t* fnb,
                       const double* fvx, const double* fvy, const double* fkout, const double* fkin, int F,
                       double one_minus_delta, const double* dz, double* corr, cudaStream_t st);
void launch_min(const double* p, int64_t n, double* out, cudaStream_t st);   // *out = min(*out, min p), *out <= 0

// field export (reference layouts)
// i0, ni: window of cells (reference numbering) written to out[0 .. ni*G) / out[0 .. ni)
void launch_export_phi(const double* phi, const int32_t* slot_of_xy, double scale, int G, int nz,
                       int nxy, int64_t Sb, int64_t i0, int64_t ni, double* out, cudaStream_t st);
void launch_export_cell(const double* phi, const int32_t* slot_of_xy, const int32_t* mats,
                        const double* xs_g, const double* area, const double* dz, int has_z,
                        double scale, int G, int nz, int nxy, int64_t Sb, int64_t i0, int64_t ni, double* out,
                        cudaStream_t st);
void launch_export_psi(const double* psi_block, const ClassDev* cl, const int32_t* pos_of,
                       const int32_t* slot_of_xy, int d, int nd, int m, const int32_t* gloc,
                       double scale, int G, int M, int nz, int nxy, double* out, double* minval,
                       cudaStream_t st);
void launch_import_phi(double* phi, const int32_t* slot_of_xy, int G, int nz, int nxy, int64_t Sb,
                       const double* in, cudaStream_t st);
void launch_fill(double* p, double v, int64_t n, cudaStream_t st);
void launch_fill_phi(double* phi, const int32_t* mats, double v, int G, int64_t n, cudaStream_t st);

}  // namespace pampa_sn
