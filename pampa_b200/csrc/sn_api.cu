// sn_api.cu -- the C ABI declared in include/pampa_sn.h: handle, device memory, iteration.
#include <dlfcn.h>

#include <array>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "sn_kernels.cuh"

using namespace pampa_sn;

namespace {

std::string g_create_error;

// minimal NCCL surface, bound at run time so that one-GPU use has no NCCL dependency
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
   void* lib = nullptr;
   int (*GetUniqueId)(NcclUniqueId*) = nullptr;
   int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
   int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
   int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
   int (*CommDestroy)(NcclComm) = nullptr;
   const char* (*GetErrorString)(int) = nullptr;
   bool load(std::string& err) {
      if (lib) return true;
      lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!lib) { err = std::string("unable to load libnccl.so.2: ") + dlerror(); return false; }
      GetUniqueId = (int (*)(NcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
      CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
      AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllReduce");
      AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllGather");
      CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
      GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
      if (!GetUniqueId || !CommInitRank || !AllReduce || !AllGather || !CommDestroy) { err = "incomplete NCCL library"; return false; }
      return true;
   }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MIN = 3;

struct LaunchGroup { int wave, cls, kind, dt, fin, ring; bool extras; int64_t offset; int count; };
// one dataflow launch: every tile-class task of one chunk size, in ticket (topological) order
struct FlowLaunch { int dt, fin; int64_t offset; int count; std::vector<double> mw; int nch; };
// fast classes / chunks of one shared tiling (the shear passes run once per tiling)
struct TilingDev { int npatch = 0; int32_t* d_cell_of = nullptr; int32_t *d_classes = nullptr, *d_chunks = nullptr; int nclasses = 0, nchunks = 0;
                   int nplus = 0;       // +z chunks, at the head of d_chunks
                   int zsplit = 1; };   // slabs per column of this tiling's un-shear pass (unshear_zsplit)

}  // namespace

struct pampa_sn_handle {
   std::string err;
   pampa_sn_options opts;
   Plan plan;
   int device = 0;
   cudaStream_t stream = nullptr;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   // the ordering classes sweep independently: one stream each, so that the thin first / last
   // wavefronts of one class overlap the wide ones of another
   static constexpr int NSTREAMS = 8;
   cudaStream_t cls_stream[NSTREAMS] = {};
   cudaEvent_t ev_fork = nullptr, ev_join[NSTREAMS] = {};
   std::vector<void*> allocs;
   std::vector<void*> ipc_mapped;        // peer buffers opened with cudaIpcOpenMemHandle
   int64_t device_bytes = 0;
   int64_t launches = 0;

   int G = 0, M = 0, Gown = 0, nmat = 0;
   std::vector<int32_t> gloc;
   std::vector<double> h_beta;

   // device data
   int32_t *d_slot_of_xy = nullptr, *d_mats = nullptr, *d_gloc = nullptr;
   double *d_area = nullptr, *d_dz = nullptr, *d_inv_dz = nullptr;
   double *d_sig_t = nullptr, *d_sig_s = nullptr, *d_chi = nullptr, *d_nusf = nullptr, *d_kapsf = nullptr;
   double *d_phi = nullptr, *d_phi_new = nullptr, *d_q = nullptr, *d_psi = nullptr;
   double *d_bnd[2] = {nullptr, nullptr}, *d_bndz[2] = {nullptr, nullptr};
   int bnd_cur = 0;
   int64_t bnd_count = 0, bndz_count = 0;
   double *d_partials = nullptr, *d_sums = nullptr, *d_sums_all = nullptr;   // d_sums_all: [rank][5] (sharded runs)
   // Anderson acceleration state (allocated by the first accelerated solve)
   double* aa_f[AA_SLOTS] = {};
   double* aa_g[AA_SLOTS] = {};
   double* aa_b[AA_SLOTS] = {};         // boundary-flux part of the state (reflective problems)
   double* aa_bz[AA_SLOTS] = {};
   double *d_aa_partials = nullptr, *d_aa_dots = nullptr;
   AAState* d_aa_state = nullptr;       // device-resident bookkeeping of the accelerated iteration
   AAState* h_aa_ring = nullptr;        // pinned: snapshots of it, read one iteration late
   int aa_slots = 0;
   double psi_scale_factor = 1.0;       // psi normalisation relative to phi (1 unless accelerated)
   double* d_stage = nullptr;           // device staging buffer of the field import / export calls
   int64_t stage_count = 0;
   bool group_gather = false;           // group-sharded run with the in-place allgather of phi
   // peer-to-peer delivery of the flux moments (group-sharded runs, all ranks on one node): the iterate is
   // double-buffered and the reduction pass stores the new slabs straight into the other buffer of every peer
   bool p2p = false;
   int fuse_next = 0;                    // the next sweep may fuse the reduction into its last un-shear pass: 1 = and
                                         // rotate / deliver the iterate, 2 = in place (phi_new keeps the result)
   int fused_done = 0;                   // ... and did (same codes): d_sums holds this rank's sums
   double* d_fuse_partials = nullptr;    // [5][npatch_b * G]
   double* d_phi_buf[2] = {nullptr, nullptr};   // d_phi is d_phi_buf[phi_cur]
   int phi_cur = 0;
   double* peer_phi[2][PEER_MAX] = {};   // the two buffers of the other ranks (IPC mappings)
   int npeers = 0;
   int nblocks_reduce = 0;
   ReduceScalars* d_sc = nullptr;
   ClassDev* d_classes = nullptr;
   ChunkDev* d_chunks = nullptr;
   Task* d_tasks = nullptr;
   std::vector<LaunchGroup> groups;
   std::vector<FlowLaunch> flows;
   int* d_flow_ctl = nullptr;            // [16 ticket counters | progress[chunk][owned group][patch]]
   int64_t flow_ctl_count = 0;
   std::vector<int32_t*> d_pos_of;       // per class
   int32_t** d_class_pos_of = nullptr;
   // LS
   int nls = 0; int64_t ls_nnz = 0;
   int32_t *d_ls_ptr = nullptr, *d_ls_nbr = nullptr, *d_dir_chunk = nullptr, *d_dir_d = nullptr;
   double *d_ls_coef = nullptr, *d_ls_dD = nullptr, *d_ls_rhs = nullptr;
   std::vector<int> dir_chunk, dir_d;
   std::vector<double*> chunk_psi;       // host copy of ChunkDev::psi (field export)
   // mixed-face-interpolation delta < 1: deferred correction (sn_delta_corr_kernel)
   double delta = 1.0;
   double* d_corr = nullptr;             // same size and layout as d_psi
   int64_t psi_count = 0;
   int32_t* d_fnb = nullptr;             // [Sb][F] base slot of the neighbour across lateral face f, -1: boundary
   double *d_fvx = nullptr, *d_fvy = nullptr, *d_fkout = nullptr, *d_fkin = nullptr;   // [Sb][F]
   int F = 0;
   bool extras = false;
   bool multi_stream = false;
   // staged tile kernel
   std::vector<char> class_fast;
   std::vector<TilingDev> tilings;       // shear / un-shear work lists per shared tiling
   std::vector<int32_t*> class_mats_s;   // per class: device material maps (nullptr: not used by its kernel)
   std::vector<uint8_t*> class_mats_c;
   int mat_bytes = 4;                    // element size of the dataflow kernel's material rows
   int nmat_cap = 0;                     // materials the cross-section tables were allocated for
   int nfast_classes = 0, nfast_chunks = 0;
   bool use_graph = false, graph_failed = false;
   cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};   // one per parity of the alternating boundary buffers
   int64_t launches_per_sweep = 0;
   int groups_generic = 0;               // launch groups of the generic kernel (it accumulates into phi_new with atomics)
   int gm = 1;                           // owned groups per block of the step-major arrays (sn_kernels.cuh)
   int32_t* d_gown = nullptr;            // [Gown] group of a local index

   // iteration state
   ReduceScalars sc{};
   // Normalisation of the last solve: fields = scale * device values.  The reference forms
   // phi = 4 pi sum_m w_m psi and then rescales phi and psi separately to the requested power
   // (src/NeutronicSolver.cxx:46-78, src/SNSolver.cxx:302-341), so the 4 pi cancels and both
   // end up as scale * (device value), scale = power / sum_i V_i sum_g kappa-sigma-f phi~.
   double scale = 1.0;
   double keff = 1.0;
   bool solved = false;
   double last_sweep_ms = 0, last_source_ms = 0, last_reduce_ms = 0, last_solve_ms = 0;
   // pampa_sn_iterate_timed: events around the sweep-kernel launches alone (after the shear pass,
   // before the un-shear pass), one pair per iteration
   std::vector<cudaEvent_t>* kernel_events = nullptr;
   std::vector<cudaEvent_t>* exchange_events = nullptr;   // around the flux-moment allgather of a sharded run
   double timed_kernel_ms = 0, timed_source_ms = 0, timed_reduce_ms = 0, timed_exchange_ms = 0;
   std::vector<double> h_temperature, h_delayed;
   NcclComm comm = nullptr;

   SweepGlobals globals() const {
      SweepGlobals gp{};
      gp.classes = d_classes; gp.chunks = d_chunks; gp.gloc = d_gloc; gp.gown = d_gown; gp.gm = gm; gp.q = d_q;
      gp.phi_new = d_phi_new; gp.mats = d_mats; gp.sigma_t = d_sig_t; gp.inv_dz = d_inv_dz;
      gp.bnd_old = d_bnd[bnd_cur]; gp.bnd_new = d_bnd[1 - bnd_cur];
      gp.bndz_old = d_bndz[bnd_cur]; gp.bndz_new = d_bndz[1 - bnd_cur];
      gp.ls_dD = d_ls_dD; gp.ls_rhs = d_ls_rhs;
      gp.Sb = plan.Sb; gp.G = G; gp.Gown = Gown; gp.M = M; gp.nz = plan.nz; gp.Kc = plan.Kc;
      gp.has_z = plan.has_z; gp.nrf = plan.num_rfaces; gp.nls = nls;
      gp.bcz_minus_refl = bcz_refl[0]; gp.bcz_plus_refl = bcz_refl[1];
      gp.store_psi = opts.store_psi ? 1 : 0;
      gp.nmat = nmat;
      gp.uniform_dz = uniform_dz;
      gp.corr_off = d_corr ? (int64_t)(d_corr - d_psi) : 0;
      gp.np_stride = np_stride;
      { const char* e = std::getenv("PAMPA_SN_DBG"); gp.dbg = e ? std::atoi(e) : 0; }
      return gp;
   }
   int bcz_refl[2] = {0, 0};
   int uniform_dz = 1;
   int np_stride = 0;                    // patches per (chunk, block) in the dataflow progress counters
};

#define SN_FAIL(h, msg) do { (h)->err = (msg); return 1; } while (0)
#define SN_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
   (h)->err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call; return 1; } } while (0)

namespace {

template <typename T>
int dev_alloc(pampa_sn_handle* h, T** p, int64_t count) {
   *p = nullptr;
   if (count <= 0) return 0;
   void* v = nullptr;
   cudaError_t e = cudaMalloc(&v, (size_t)count * sizeof(T));
   if (e != cudaSuccess) {
      h->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " allocating " +
               std::to_string((size_t)count * sizeof(T)) + " bytes";
      return 1;
   }
   h->allocs.push_back(v);
   h->device_bytes += count * (int64_t)sizeof(T);
   *p = (T*)v;
   return 0;
}

template <typename T>
int dev_upload(pampa_sn_handle* h, T** p, const T* src, int64_t count) {
   if (dev_alloc(h, p, count)) return 1;
   if (count > 0) SN_CUDA(h, cudaMemcpy(*p, src, (size_t)count * sizeof(T), cudaMemcpyHostToDevice));
   return 0;
}

template <typename T>
int dev_upload(pampa_sn_handle* h, T** p, const std::vector<T>& v) {
   return dev_upload(h, p, v.data(), (int64_t)v.size());
}

int dt_template(int nd) { return nd; }      // one instantiation per chunk size 1..DT_MAX

int upload_xs(pampa_sn_handle* h, const pampa_sn_xs* xs, bool first, bool allow_resize = false) {
   const int G = xs->num_groups, nm = xs->num_materials;
   if (!first && (G != h->G || (nm != h->nmat && !allow_resize))) SN_FAIL(h, "cross-section table shape changed");
   const int64_t n = (int64_t)nm * G;
   if (first || nm > h->nmat_cap) {      // (the old tables of a grown set stay allocated until the handle goes)
      if (dev_alloc(h, &h->d_sig_t, n) || dev_alloc(h, &h->d_sig_s, n * G) || dev_alloc(h, &h->d_chi, n) ||
          dev_alloc(h, &h->d_nusf, n) || dev_alloc(h, &h->d_kapsf, n)) return 1;
      h->nmat_cap = nm;
   }
   h->nmat = nm;
   SN_CUDA(h, cudaMemcpy(h->d_sig_t, xs->sigma_total, n * sizeof(double), cudaMemcpyHostToDevice));
   SN_CUDA(h, cudaMemcpy(h->d_sig_s, xs->sigma_scattering, n * G * sizeof(double), cudaMemcpyHostToDevice));
   SN_CUDA(h, cudaMemcpy(h->d_chi, xs->chi_effective, n * sizeof(double), cudaMemcpyHostToDevice));
   SN_CUDA(h, cudaMemcpy(h->d_nusf, xs->nu_sigma_fission, n * sizeof(double), cudaMemcpyHostToDevice));
   SN_CUDA(h, cudaMemcpy(h->d_kapsf, xs->kappa_sigma_fission, n * sizeof(double), cudaMemcpyHostToDevice));
   h->h_beta.assign(nm, 0.0);
   if (xs->beta_total) h->h_beta.assign(xs->beta_total, xs->beta_total + nm);
   return 0;
}

// material map in the padded base numbering [nz][Sb] (-1 in holes)
std::vector<int32_t> base_material_map(const Plan& pl, const int32_t* materials) {
   std::vector<int32_t> mats((size_t)pl.nz * pl.Sb, -1);
   for (int k = 0; k < pl.nz; k++)
      for (int c = 0; c < pl.nxy; c++) mats[(size_t)k * pl.Sb + pl.slot_of_xy[c]] = materials[(size_t)k * pl.nxy + c];
   return mats;
}
// ... in a class's (patch, pipeline step, lane) order (generic and tile kernels)
std::vector<int32_t> class_step_map(const Plan& pl, const ClassPlan& cp, const std::vector<int32_t>& mats) {
   const int nz = pl.nz;
   std::vector<int32_t> ms((size_t)cp.npatch * cp.nsteps * PS, -1);
   for (int64_t sl = 0; sl < cp.S; sl++) {
      if (cp.cell_of[sl] < 0) continue;
      const int64_t p = sl / PS, lane = sl % PS;
      for (int kp = 0; kp < nz; kp++) {
         const int k = cp.zdir >= 0 ? kp : nz - 1 - kp;
         ms[((size_t)p * cp.nsteps + kp + cp.lvl[sl]) * PS + lane] = mats[(size_t)k * pl.Sb + cp.cell_of[sl]];
      }
   }
   return ms;
}
// ... of the dataflow kernel, cyclic in the pipeline step: the lane at level l is at layer (step - l) mod nz of
// some group of its block
std::vector<uint8_t> class_cyclic_map(const Plan& pl, const ClassPlan& cp, const std::vector<int32_t>& mats, int mb) {
   const int nz = pl.nz;
   std::vector<uint8_t> mc((size_t)cp.npatch * nz * PS * mb, 0);
   for (int64_t sl = 0; sl < cp.S; sl++) {
      if (cp.cell_of[sl] < 0) continue;
      const int64_t p = sl / PS, lane = sl % PS;
      for (int kp = 0; kp < nz; kp++) {
         const int k = cp.zdir >= 0 ? kp : nz - 1 - kp;
         const int32_t m = mats[(size_t)k * pl.Sb + cp.cell_of[sl]];
         const size_t e = ((size_t)p * nz + (kp + cp.lvl[sl]) % nz) * PS + lane;
         if (mb == 1) mc[e] = (uint8_t)m; else std::memcpy(&mc[e * 4], &m, 4);
      }
   }
   return mc;
}

int sync_scalars(pampa_sn_handle* h) {
   SN_CUDA(h, cudaMemcpyAsync(&h->sc, h->d_sc, sizeof(ReduceScalars), cudaMemcpyDeviceToHost, h->stream));
   SN_CUDA(h, cudaStreamSynchronize(h->stream));
   return 0;
}

int do_source(pampa_sn_handle* h) {
   launch_source(h->d_phi, h->d_q, h->d_mats, h->d_sig_s, h->d_chi, h->d_nusf, h->d_sc, h->d_gloc, h->G,
                 h->nmat, h->plan.nz, h->plan.Sb, h->stream);
   h->launches++;
   return 0;
}

// the launch sequence of one sweep (issued directly, or captured once into a CUDA graph)
int sweep_launches(pampa_sn_handle* h) {
   SweepGlobals gp = h->globals();
   if (h->nls > 0) {
      launch_ls_rhs(gp, h->d_ls_ptr, h->d_ls_nbr, h->d_ls_coef, h->ls_nnz, h->d_dir_chunk, h->d_dir_d,
                    (const int32_t* const*)h->d_class_pos_of, h->stream);
      h->launches++;
   }
   if (h->opts.num_ranks > 1) {
      // sharded: the boundary buffers are summed over the ranks afterwards, so the entries this rank
      // does not write must be zero rather than two sweeps old
      if (gp.bnd_new) cudaMemsetAsync(gp.bnd_new, 0, (size_t)h->bnd_count * sizeof(double), h->stream);
      if (gp.bndz_new) cudaMemsetAsync(gp.bndz_new, 0, (size_t)h->bndz_count * sizeof(double), h->stream);
   }
   if (h->d_corr) {
      // delta < 1: (T_delta - T_1) applied to the angular flux of the previous sweep, chunk by chunk
      for (size_t c = 0; c < h->plan.chunks.size(); c++) {
         if (!h->chunk_psi[c]) continue;
         const int cls = h->plan.chunks[c].cls;
         launch_delta_corr(gp, (int)c, h->plan.classes[cls].npatch, h->d_pos_of[cls], h->d_fnb, h->d_fvx, h->d_fvy,
                           h->d_fkout, h->d_fkin, h->F, 1.0 - h->delta, h->d_dz, h->d_corr + (h->chunk_psi[c] - h->d_psi),
                           h->stream);
         h->launches++;
      }
   }
   for (const TilingDev& tg : h->tilings)
      if (tg.nclasses > 0) {
         launch_shear_q(gp, h->d_classes, tg.d_classes, tg.nclasses, tg.npatch, tg.d_cell_of, h->stream);
         h->launches++;
      }
   const int ns = h->multi_stream ? pampa_sn_handle::NSTREAMS : 0;
   if (!h->flows.empty()) {
      // tickets and progress counters of the dataflow launches start from zero every sweep
      cudaMemsetAsync(h->d_flow_ctl, 0, (size_t)h->flow_ctl_count * sizeof(int), h->stream);
   }
   if (h->kernel_events) {
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->stream); h->kernel_events->push_back(e);
   }
   if (ns) {
      cudaEventRecord(h->ev_fork, h->stream);
      for (int i = 0; i < ns; i++) cudaStreamWaitEvent(h->cls_stream[i], h->ev_fork, 0);
   }
   for (size_t f = 0; f < h->flows.size(); f++) {
      const FlowLaunch& fl = h->flows[f];
      cudaStream_t st = ns ? h->cls_stream[f % ns] : h->stream;
      if (launch_sweep_flow(gp, h->d_tasks + fl.offset, fl.count, fl.dt, fl.fin, h->extras, h->d_flow_ctl + f, h->d_flow_ctl + 16,
                            fl.mw.data(), fl.nch, st))
         SN_FAIL(h, "internal: more chunks in a dataflow launch than its direction table holds");
      h->launches++;
   }
   for (const LaunchGroup& lg : h->groups) {
      cudaStream_t st = ns ? h->cls_stream[lg.cls % ns] : h->stream;
      if (lg.kind == 1) launch_sweep_tile(gp, h->d_tasks + lg.offset, lg.count, lg.dt, lg.extras, st);
      else launch_sweep(gp, h->d_tasks + lg.offset, lg.count, lg.dt, lg.fin, lg.ring, lg.extras, st);
      h->launches++;
   }
   for (int i = 0; i < ns; i++) {
      cudaEventRecord(h->ev_join[i], h->cls_stream[i]);
      cudaStreamWaitEvent(h->stream, h->ev_join[i], 0);
   }
   if (h->kernel_events) {
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->stream); h->kernel_events->push_back(e);
   }
   // Fused tail of the iteration (every class on a dataflow kernel, one GPU or a group-sharded run): the last un-shear
   // pass -- the base tiling's, after those of the other tilings -- holds the final flux moments of a column in
   // registers and also reduces them.  Mode 1 (plain source iterations; sharded: needs peer access) writes them into
   // the other iterate buffer, here and on every peer; mode 2 (accelerated iterations, whose next iterate is a mix)
   // leaves them in phi_new.
   const int fuse = h->fuse_next;
   bool can_fuse = fuse && h->groups_generic == 0 && h->tilings[0].nchunks > 0 && h->d_fuse_partials &&
                   !std::getenv("PAMPA_SN_NO_FUSE");
   if (fuse == 1) can_fuse = can_fuse && h->d_phi_buf[1] && (h->opts.num_ranks == 1 || (h->comm && h->group_gather && h->p2p));
   if (fuse == 2) can_fuse = can_fuse && (!h->comm || h->group_gather);
   if (can_fuse) {
      bool overwrite = true;
      for (size_t t = 1; t < h->tilings.size(); t++) {
         const TilingDev& tg = h->tilings[t];
         if (tg.nchunks <= 0) continue;
         launch_unshear_phi(gp, h->d_chunks, h->d_classes, tg.d_chunks, tg.nchunks, tg.nplus, tg.npatch, overwrite ? 1 : 0,
                            tg.d_cell_of, tg.zsplit, h->stream);
         overwrite = false;
         h->launches++;
      }
      const TilingDev& tg = h->tilings[0];
      const int out = 1 - h->phi_cur;
      double* phi_out = fuse == 1 ? h->d_phi_buf[out] : h->d_phi_new;
      launch_unshear_phi_fused(gp, h->d_chunks, h->d_classes, tg.d_chunks, tg.nchunks, tg.nplus, tg.npatch, overwrite ? 1 : 0,
                               h->d_phi, phi_out, fuse == 1 ? h->peer_phi[out] : nullptr,
                               fuse == 1 ? h->npeers : 0, h->d_mats, h->d_nusf, h->d_kapsf,
                               h->d_area, h->d_dz, h->plan.has_z, h->d_fuse_partials, h->d_sums, tg.zsplit, h->stream);
      h->launches += 2;
      if (fuse == 1) {
         h->phi_cur = out;
         h->d_phi = h->d_phi_buf[out];
      }
      h->fused_done = fuse;
      return 0;
   }
   // every owned chunk on the tile kernels: nothing else adds to phi_new, the first pass may overwrite it
   bool overwrite = h->groups_generic == 0;
   for (const TilingDev& tg : h->tilings)
      if (tg.nchunks > 0) {
         launch_unshear_phi(gp, h->d_chunks, h->d_classes, tg.d_chunks, tg.nchunks, tg.nplus, tg.npatch, overwrite ? 1 : 0,
                            tg.d_cell_of, tg.zsplit, h->stream);
         overwrite = false;
         h->launches++;
      }
   return 0;
}

// One sweep.  Plans with many small launches (unstructured meshes: one launch per wavefront, ordering class
// and kernel variant, thousands per sweep) are launch-bound, so their sequence -- fixed for the life of the
// handle but for the two boundary buffers that alternate -- is captured once per buffer parity into a CUDA
// graph, fork / join over the class streams included, and replayed.
int do_sweep(pampa_sn_handle* h, int fuse = 0) {
   const bool graphed = h->use_graph && !h->graph_failed;
   h->fuse_next = graphed ? 0 : fuse;
   h->fused_done = 0;
   if (!graphed) {
      if (sweep_launches(h)) return 1;
   } else {
      const int par = h->bnd_cur;
      if (!h->graph_exec[par]) {
         const int64_t l0 = h->launches;
         cudaGraph_t graph = nullptr;
         cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
         int rc = 0;
         if (e == cudaSuccess) {
            std::vector<cudaEvent_t>* kev = h->kernel_events;
            h->kernel_events = nullptr;                  // timing events cannot be recorded inside a capture
            rc = sweep_launches(h);
            h->kernel_events = kev;
            e = cudaStreamEndCapture(h->stream, &graph);
         }
         if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&h->graph_exec[par], graph, 0);
         if (graph) cudaGraphDestroy(graph);
         h->launches_per_sweep = h->launches - l0;
         h->launches = l0;
         if (rc) return 1;
         if (e != cudaSuccess) {                         // fall back to direct launches, loudly in verbose mode
            cudaGetLastError();
            h->graph_failed = true; h->graph_exec[par] = nullptr;
            if (h->opts.verbose) std::printf("pampa_sn: sweep graph capture failed (%s), launching directly\n", cudaGetErrorString(e));
            if (sweep_launches(h)) return 1;
            h->bnd_cur = 1 - h->bnd_cur;
            return 0;
         }
      }
      if (h->kernel_events) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, h->stream); h->kernel_events->push_back(ev); }
      SN_CUDA(h, cudaGraphLaunch(h->graph_exec[par], h->stream));
      if (h->kernel_events) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, h->stream); h->kernel_events->push_back(ev); }
      h->launches += h->launches_per_sweep;
   }
   h->bnd_cur = 1 - h->bnd_cur;      // what this sweep wrote is what the next one reads
   return 0;
}

int nccl_fail(pampa_sn_handle* h, int r, const char* what) {
   h->err = std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
   return 1;
}

// mirrored-direction boundary fluxes may live on another rank: sum the (zero-filled) buffers
int exchange_boundaries(pampa_sn_handle* h) {
   double* bufs[2] = {h->d_bnd[h->bnd_cur], h->d_bndz[h->bnd_cur]};
   int64_t cnt[2] = {h->bnd_count, h->bndz_count};
   for (int b = 0; b < 2; b++)
      if (bufs[b] && cnt[b] > 0) {
         int r = g_nccl.AllReduce(bufs[b], bufs[b], (size_t)cnt[b], NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
         if (r != 0) return nccl_fail(h, r, "the boundary-flux allreduce");
      }
   return 0;
}

// One exchange + reduction step after a sweep.
//  * one GPU:            reduce (production, norms, phi <- phi_new), k update.
//  * angle-set sharding: allreduce(phi_new) over the ranks, then the same on every rank.
//  * group sharding:     every rank reduces the groups it swept, the five scalars are allreduced,
//                        and the new phi is completed with an in-place allgather of the group slabs
//                        (half the wire bytes of the allreduce, and source / reduce are sharded too).
// exchange (sharded runs) + block reduction of the sweep result into d_sums[5]
// after_sweep = false: phi_new holds a field every rank already has in full (pampa_sn_set), not this rank's share of
// a sweep, so the angle-set allreduce of phi_new and of the boundary fluxes must not run (they would multiply by
// the number of ranks)
int reduce_sums(pampa_sn_handle* h, int rotate, bool* pushed = nullptr, bool after_sweep = true) {
   const Plan& pl = h->plan;
   const int64_t slab = (int64_t)pl.nz * pl.Sb;
   if (h->comm && !h->group_gather && after_sweep) {
      int r = g_nccl.AllReduce(h->d_phi_new, h->d_phi_new, (size_t)(h->G * slab), NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
      if (r != 0) return nccl_fail(h, r, "the flux-moment allreduce");
   }
   if (h->comm && after_sweep && exchange_boundaries(h)) return 1;
   const int owned_only = (h->comm && h->group_gather) ? 1 : 0;
   if (pushed) *pushed = false;
   if (h->fused_done == 1) {
      // the sweep's un-shear pass already reduced (d_sums) and delivered the flux moments (buffers swapped)
      h->fused_done = 0;
      if (!rotate || !pushed) SN_FAIL(h, "internal: fused sweep followed by a non-rotating reduction");
      *pushed = true;
   } else if (h->fused_done == 2) {
      // ... reduced only: phi_new holds the sweep result, d_sums this rank's sums
      h->fused_done = 0;
      if (rotate) SN_FAIL(h, "internal: in-place fused sweep followed by a rotating reduction");
   } else if (owned_only && rotate && h->p2p && pushed) {
      // reduction of the owned groups that also delivers them to every rank's other iterate buffer
      const int out = 1 - h->phi_cur;
      launch_reduce_push(h->d_phi, h->d_phi_new, h->d_phi_buf[out], h->peer_phi[out], h->npeers, h->d_mats, h->d_nusf,
                         h->d_kapsf, h->d_area, h->d_dz, pl.has_z, h->G, pl.nz, pl.Sb, h->d_gloc, h->d_partials,
                         h->nblocks_reduce, h->d_sums, h->stream);
      h->phi_cur = out;
      h->d_phi = h->d_phi_buf[out];
      *pushed = true;
      h->launches += 2;
   } else {
      launch_reduce(h->d_phi, h->d_phi_new, h->d_mats, h->d_nusf, h->d_kapsf, h->d_area, h->d_dz, pl.has_z, h->G,
                    pl.nz, pl.Sb, h->d_gloc, owned_only, rotate, h->d_partials, h->nblocks_reduce, h->d_sums, h->stream);
      h->launches += 2;
   }
   if (owned_only) {
      // one collective for the five scalars (four sums and a minimum): allgather, combined in rank order
      int r = g_nccl.AllGather(h->d_sums, h->d_sums_all, 5, NCCL_FLOAT64, h->comm, h->stream);
      if (r != 0) return nccl_fail(h, r, "the scalar allgather");
      launch_combine_sums(h->d_sums_all, h->opts.num_ranks, h->d_sums, h->stream);
      h->launches++;
   }
   return 0;
}

// group-sharded runs: complete phi with an in-place allgather of the group slabs
int gather_phi(pampa_sn_handle* h) {
   if (!(h->comm && h->group_gather)) return 0;
   const int64_t slab = (int64_t)h->plan.nz * h->plan.Sb;
   const int nr = h->opts.num_ranks;
   for (int j = 0; j < h->G / nr; j++) {                // groups j*nr .. j*nr+nr-1: rank r owns j*nr + r
      double* base = h->d_phi + (int64_t)j * nr * slab;
      int r = g_nccl.AllGather(base + (int64_t)h->opts.rank * slab, base, (size_t)slab, NCCL_FLOAT64, h->comm, h->stream);
      if (r != 0) return nccl_fail(h, r, "the flux-moment allgather");
   }
   return 0;
}

// One exchange + reduction step after a sweep.
//  * one GPU:            reduce (production, norms, phi <- phi_new), k update.
//  * angle-set sharding: allreduce(phi_new) over the ranks, then the same on every rank.
//  * group sharding:     every rank reduces the groups it swept, the five scalars are allreduced,
//                        and the new phi is completed with an in-place allgather of the group slabs
//                        (half the wire bytes of the allreduce, and source / reduce are sharded too).
int do_reduce(pampa_sn_handle* h, int update_k, bool after_sweep = true) {
   bool pushed = false;
   if (reduce_sums(h, 1, &pushed, after_sweep)) return 1;
   launch_update_k(h->d_sums, h->d_sc, update_k, h->stream);
   h->launches++;
   if (h->exchange_events) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->stream); h->exchange_events->push_back(e); }
   const int rc = pushed ? 0 : gather_phi(h);            // (pushed: the scalar collective was the barrier)
   if (h->exchange_events) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->stream); h->exchange_events->push_back(e); }
   return rc;
}

int check_async(pampa_sn_handle* h, const char* what) {
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) SN_FAIL(h, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
   return 0;
}

}  // namespace

extern "C" {

void pampa_sn_default_options(pampa_sn_options* o) {
   std::memset(o, 0, sizeof(*o));
   o->store_psi = 1;
   o->group_merge = 0;
   o->inline_edges = 0;
   o->no_graph = 0;
   o->num_ranks = 1;
}

const char* pampa_sn_last_error(const pampa_sn_handle* h) {
   return h ? h->err.c_str() : g_create_error.c_str();
}

int pampa_sn_plan_check(const pampa_sn_mesh* mesh, const pampa_sn_quadrature* quad, int32_t num_groups,
                        const pampa_sn_options* opts, pampa_sn_info* info) {
   try {
      pampa_sn_options o;
      if (opts) o = *opts; else pampa_sn_default_options(&o);
      Plan pl;
      build_plan(PlanInput{mesh, quad, num_groups, o}, pl);
      // invariants: every xy cell appears exactly once per class; sources point upwind
      for (auto& cp : pl.classes) {
         std::vector<int> seen(pl.Sb, 0);
         for (int64_t s = 0; s < cp.S; s++) {
            if ((cp.cell_of[s] >= 0) != (cp.lvl[s] != LVL_EMPTY)) throw std::runtime_error("slot/level mismatch");
            if (cp.cell_of[s] >= 0) seen[cp.cell_of[s]]++;
         }
         for (int c = 0; c < pl.nxy; c++) if (seen[pl.slot_of_xy[c]] != 1) throw std::runtime_error("cell coverage");
         for (int f = 0; f < FIN_MAX; f++)
            for (int64_t s = 0; s < cp.S; s++) {
               int32_t code = cp.in_src[(size_t)f * cp.S + s];
               if (code < 0) continue;
               int kind = code >> SRC_KIND_SHIFT, pay = code & SRC_PAYLOAD;
               int64_t p = s / pl.P;
               if (kind == SRC_LOCAL) {
                  int64_t u = p * pl.P + pay;
                  int diff = (int)cp.lvl[s] - (int)cp.lvl[u];
                  if (cp.lvl[u] == LVL_EMPTY || diff < 1 || diff >= cp.ring) throw std::runtime_error("local source level");
                  // the dataflow kernel reads an in-patch source `diff` pipeline steps back (1 or 2)
                  if (cp.in_hidx[(size_t)f * cp.S + s] != diff) throw std::runtime_error("local source delay");
                  if (cp.fast_flow && diff > 2) throw std::runtime_error("dataflow class with a source more than two steps back");
               } else if (kind == SRC_GLOBAL) {
                  int64_t up = pay / pl.P;
                  if (cp.lvl[pay] == LVL_EMPTY) throw std::runtime_error("global source hole");
                  if (up != p && cp.patch_level[up] >= cp.patch_level[p]) throw std::runtime_error("patch order");
                  if (cp.fast_flow && (cp.in_hidx[(size_t)f * cp.S + s] >= FLOW_HALO || cp.eidx[pay] >= FLOW_EXPORT || f >= FLOW_FIN))
                     throw std::runtime_error("dataflow class outside the kernel's halo / export / face limits");
               }
            }
      }
      if (info) {
         std::memset(info, 0, sizeof(*info));
         info->num_cells = (int64_t)pl.nxy * pl.nz; info->num_groups = pl.G; info->num_directions = pl.M;
         info->updates_per_sweep = pl.owned_updates;
         info->sweep_launches = (int64_t)pl.waves.size();
         for (auto& w : pl.waves) info->sweep_tasks += (int64_t)w.size();
         info->num_classes = (int64_t)pl.classes.size(); info->num_chunks = (int64_t)pl.chunks.size();
         info->tile_classes = pl.tile_classes;
         info->num_tilings = (int64_t)pl.tilings.size(); info->lattice = pl.lattice;
         for (auto& cp : pl.classes) info->flow_classes += cp.fast_flow ? 1 : 0;
      }
   } catch (const std::exception& e) {
      g_create_error = e.what();
      return 1;
   }
   return 0;
}

int pampa_sn_create(pampa_sn_handle** out, const pampa_sn_mesh* mesh, const pampa_sn_xs* xs,
                    const pampa_sn_quadrature* quad, const pampa_sn_ls* ls, const pampa_sn_options* opts) {
   *out = nullptr;
   std::unique_ptr<pampa_sn_handle> hp(new pampa_sn_handle());
   pampa_sn_handle* h = hp.get();
   auto fail = [&](int) { g_create_error = h->err; pampa_sn_destroy(hp.release()); return 1; };
   if (opts) h->opts = *opts; else pampa_sn_default_options(&h->opts);
   if (!mesh || !xs || !quad) { h->err = "null input"; return fail(1); }
   if (h->opts.num_ranks < 1 || h->opts.rank < 0 || h->opts.rank >= h->opts.num_ranks) { h->err = "wrong rank"; return fail(1); }
   if (h->opts.patch_cells > PS) { h->err = "patch_cells must be <= 256"; return fail(1); }
   h->G = xs->num_groups; h->M = quad->num_directions; h->nmat = xs->num_materials;
   for (int64_t i = 0; i < (int64_t)mesh->num_layers * mesh->num_xy_cells; i++)
      if (mesh->materials[i] < 0 || mesh->materials[i] >= xs->num_materials) { h->err = "wrong material index"; return fail(1); }
   if (ls && ls->num_cells > 0 && (mesh->has_z_faces || mesh->num_layers != 1)) {
      h->err = "least-squares boundary interpolation is only supported on 1-D and 2-D meshes"; return fail(1);
   }

   h->delta = mesh->face_interpolation_delta > 0.0 ? mesh->face_interpolation_delta : 1.0;
   if (h->delta > 1.0) { h->err = "wrong weight between upwind and linear interpolation"; return fail(1); }
   if (h->delta < 1.0) {
      if (!mesh->xy_face_kout || !mesh->xy_face_kin) { h->err = "mixed-face-interpolation < 1 needs the xy_face_kout / xy_face_kin weights"; return fail(1); }
      if (!h->opts.store_psi) { h->err = "mixed-face-interpolation < 1 needs the angular flux in memory (store_psi = 1)"; return fail(1); }
   }

   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0) {
      h->err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this layer has no CPU fallback";
      return fail(1);
   }
   h->device = h->opts.device;
   if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "unable to select the CUDA device"; return fail(1); }

   try {
      build_plan(PlanInput{mesh, quad, h->G, h->opts}, h->plan);
   } catch (const std::exception& ex) { h->err = ex.what(); return fail(1); }
   Plan& pl = h->plan;
   if (pl.P != PS) { h->err = "internal: patch stride must be 256"; return fail(1); }

   auto body = [&]() -> int {
      SN_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
      SN_CUDA(h, cudaEventCreate(&h->ev0));
      SN_CUDA(h, cudaEventCreate(&h->ev1));
      SN_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      for (int i = 0; i < pampa_sn_handle::NSTREAMS; i++) {
         SN_CUDA(h, cudaStreamCreateWithFlags(&h->cls_stream[i], cudaStreamNonBlocking));
         SN_CUDA(h, cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
      }
      SN_CUDA(h, configure_sweep_kernels());
      SN_CUDA(h, configure_tile_kernels());
      SN_CUDA(h, configure_flow_kernels());
      SN_CUDA(h, configure_shear_kernels());

      const int nr = h->opts.num_ranks, rank = h->opts.rank;
      h->gloc.assign(h->G, -1); h->Gown = 0;
      for (int g = 0; g < h->G; g++)
         if (h->opts.shard_mode != 1 || g % nr == rank) h->gloc[g] = h->Gown++;
      {
         std::vector<int32_t> gown;
         for (int g = 0; g < h->G; g++) if (h->gloc[g] >= 0) gown.push_back(g);
         if (dev_upload(h, &h->d_gown, gown)) return 1;
      }
      // Groups per block of the step-major arrays = groups a dataflow task sweeps back to back.  More
      // groups: less pipeline fill / drain padding (max local levels - 1 rows per block); fewer, longer
      // tasks: a longer tail at the end of the launch.  The sigma_t table of a block lives in shared memory.
      h->gm = 1;
      if (!h->opts.wave_launch && pl.nzc == 1 && pl.tile_classes > 0) {
         int gm = h->opts.group_merge > 0 ? h->opts.group_merge : 8;
         gm = std::min(gm, std::max(1, 2048 / std::max(1, h->nmat)));
         h->gm = std::max(1, std::min(gm, h->Gown));
      }
      const int gm = h->gm, nblk = (h->Gown + gm - 1) / gm;
      auto nsm_of = [&](const ClassPlan& cp) { return cp.nsteps + (gm - 1) * pl.nz; };
      std::vector<char> chunk_owned(pl.chunks.size(), 1);
      if (h->opts.shard_mode != 1)
         for (size_t c = 0; c < pl.chunks.size(); c++) chunk_owned[c] = ((int)(c % nr) == rank);

      // mesh arrays in the padded slot numbering
      const int64_t Sb = pl.Sb; const int nz = pl.nz;
      std::vector<int32_t> mats = base_material_map(pl, mesh->materials);
      std::vector<double> area(Sb, 0.0), dz(nz, 1.0), idz(nz, 0.0);
      for (int c = 0; c < pl.nxy; c++) area[pl.slot_of_xy[c]] = mesh->xy_area[c];
      for (int k = 0; k < nz; k++)
         if (pl.has_z) { dz[k] = mesh->dz[k]; idz[k] = 1.0 / mesh->dz[k]; }
      h->uniform_dz = 1;
      for (int k = 1; k < nz; k++) if (pl.has_z && mesh->dz[k] != mesh->dz[0]) h->uniform_dz = 0;
      if (dev_upload(h, &h->d_slot_of_xy, pl.slot_of_xy) || dev_upload(h, &h->d_mats, mats) ||
          dev_upload(h, &h->d_area, area) || dev_upload(h, &h->d_dz, dz) || dev_upload(h, &h->d_inv_dz, idz) ||
          dev_upload(h, &h->d_gloc, h->gloc)) return 1;
      if (upload_xs(h, xs, true)) return 1;
      if (pl.has_z) {
         h->bcz_refl[0] = mesh->bc_types[mesh->bc_minus_z] == PAMPA_SN_BC_REFLECTIVE;
         h->bcz_refl[1] = mesh->bc_types[mesh->bc_plus_z] == PAMPA_SN_BC_REFLECTIVE;
      }

      // flux arrays
      const int64_t nphi = (int64_t)h->G * nz * Sb;
      if (dev_alloc(h, &h->d_phi, nphi) || dev_alloc(h, &h->d_phi_new, nphi) || dev_alloc(h, &h->d_q, nphi)) return 1;
      h->d_phi_buf[0] = h->d_phi; h->phi_cur = 0;
      int64_t psi_doubles = 0;
      std::vector<int64_t> psi_off(pl.chunks.size(), -1);
      for (size_t c = 0; c < pl.chunks.size(); c++)
         if (chunk_owned[c]) {
            psi_off[c] = psi_doubles;
            const ClassPlan& cpc = pl.classes[pl.chunks[c].cls];
            psi_doubles += (int64_t)pl.chunks[c].nd * nblk * cpc.npatch * nsm_of(cpc) * (h->opts.store_psi ? PSX : PERIM_MAX);
         }
      if (dev_alloc(h, &h->d_psi, psi_doubles)) return 1;
      SN_CUDA(h, cudaMemsetAsync(h->d_psi, 0, (size_t)psi_doubles * sizeof(double), h->stream));
      h->psi_count = psi_doubles;
      if (h->delta < 1.0) {
         if (dev_alloc(h, &h->d_corr, psi_doubles)) return 1;
         SN_CUDA(h, cudaMemsetAsync(h->d_corr, 0, (size_t)psi_doubles * sizeof(double), h->stream));
         const int F = mesh->max_xy_faces;
         h->F = F;
         std::vector<int32_t> fnb((size_t)Sb * F, -1);
         std::vector<double> fvx((size_t)Sb * F, 0.0), fvy((size_t)Sb * F, 0.0), fko((size_t)Sb * F, 0.0), fki((size_t)Sb * F, 0.0);
         for (int c = 0; c < pl.nxy; c++)
            for (int f = 0; f < mesh->xy_num_faces[c]; f++) {
               const size_t a = (size_t)c * F + f, b = (size_t)pl.slot_of_xy[c] * F + f;
               const int nb = mesh->xy_neighbor[a];
               if (nb < 0) continue;
               fnb[b] = pl.slot_of_xy[nb];
               fvx[b] = mesh->xy_face_fx[a] / mesh->xy_area[c]; fvy[b] = mesh->xy_face_fy[a] / mesh->xy_area[c];
               fko[b] = mesh->xy_face_kout[a]; fki[b] = mesh->xy_face_kin[a];
            }
         if (dev_upload(h, &h->d_fnb, fnb) || dev_upload(h, &h->d_fvx, fvx) || dev_upload(h, &h->d_fvy, fvy) ||
             dev_upload(h, &h->d_fkout, fko) || dev_upload(h, &h->d_fkin, fki)) return 1;
      }

      // reflective boundary buffers
      h->extras = h->delta < 1.0;       // the deferred correction is read on the kernels' EXTRAS path
      if (pl.num_rfaces > 0) {
         h->bnd_count = (int64_t)h->M * h->G * nz * pl.num_rfaces;
         for (int b = 0; b < 2; b++) {
            if (dev_alloc(h, &h->d_bnd[b], h->bnd_count)) return 1;
            SN_CUDA(h, cudaMemsetAsync(h->d_bnd[b], 0, (size_t)h->bnd_count * sizeof(double), h->stream));
         }
         h->extras = true;
      }
      if (h->bcz_refl[0] || h->bcz_refl[1]) {
         h->bndz_count = 2LL * h->M * h->G * Sb;
         for (int b = 0; b < 2; b++) {
            if (dev_alloc(h, &h->d_bndz[b], h->bndz_count)) return 1;
            SN_CUDA(h, cudaMemsetAsync(h->d_bndz[b], 0, (size_t)h->bndz_count * sizeof(double), h->stream));
         }
         h->extras = true;
      }

      // LS correction tables
      std::vector<std::vector<int32_t>> ls_of(pl.classes.size());
      if (ls && ls->num_cells > 0) {
         h->nls = ls->num_cells; h->ls_nnz = ls->ptr[ls->num_cells];
         std::vector<int32_t> nbr_slot(h->ls_nnz);
         for (int64_t a = 0; a < h->ls_nnz; a++) nbr_slot[a] = pl.slot_of_xy[ls->nbr[a]];
         std::vector<double> coef((size_t)h->M * h->ls_nnz), dD((size_t)h->M * h->nls, 0.0);
         for (int m = 0; m < h->M; m++)
            for (int b = 0; b < h->nls; b++)
               for (int a = ls->ptr[b]; a < ls->ptr[b + 1]; a++) {
                  double w = quad->directions[3*m] * ls->nvec[3*a] + quad->directions[3*m+1] * ls->nvec[3*a+1] +
                             quad->directions[3*m+2] * ls->nvec[3*a+2];
                  double c = w > 0.0 ? ls->omega[a] * w : 0.0;
                  coef[(size_t)m * h->ls_nnz + a] = c;
                  dD[(size_t)m * h->nls + b] -= c;
               }
         std::vector<int32_t> ptr(ls->ptr, ls->ptr + h->nls + 1);
         if (dev_upload(h, &h->d_ls_ptr, ptr) || dev_upload(h, &h->d_ls_nbr, nbr_slot) ||
             dev_upload(h, &h->d_ls_coef, coef) || dev_upload(h, &h->d_ls_dD, dD) ||
             dev_alloc(h, &h->d_ls_rhs, (int64_t)h->M * h->G * h->nls)) return 1;
         SN_CUDA(h, cudaMemsetAsync(h->d_ls_rhs, 0, (size_t)h->M * h->G * h->nls * sizeof(double), h->stream));
         for (size_t ci = 0; ci < pl.classes.size(); ci++) {
            ls_of[ci].assign(pl.classes[ci].S, -1);
            for (int b = 0; b < h->nls; b++) ls_of[ci][pl.classes[ci].pos_of[pl.slot_of_xy[ls->cell[b]]]] = b;
         }
         h->extras = true;
      }

      // Which kernel sweeps a class.  The dataflow kernel (one launch per sweep) takes the classes the plan marked
      // fast_flow; with wave_launch = 1 or z chunks the wavefront-launched tile kernel takes the narrower set
      // `fast`; everything else goes to the generic kernel.  The lagged LS / delta terms live on the generic path.
      const bool base_ok = h->nls == 0 && h->delta == 1.0 && h->nmat <= 4096 && !h->opts.generic_only;
      bool want_flow = !h->opts.wave_launch && pl.nzc == 1;
      auto needs_fin3 = [&](const ClassPlan& cp) { return cp.fin > 2 || cp.ring > 2 || cp.max_halo > 32; };
      if (want_flow) {   // every chunk of a flow launch needs a slot in its kernel-parameter direction table
         int per[DT_MAX + 1][2] = {};
         for (size_t c = 0; c < pl.chunks.size(); c++) {
            const ClassPlan& cp = pl.classes[pl.chunks[c].cls];
            if (chunk_owned[c] && base_ok && cp.fast_flow) per[pl.chunks[c].nd][needs_fin3(cp) ? 1 : 0]++;
         }
         for (int dt = 1; dt <= DT_MAX; dt++)
            for (int f3 = 0; f3 < 2; f3++)
               if (per[dt][f3] > flow_max_chunks(dt) || (f3 && per[dt][f3] > 0 && dt > flow3_max_dt())) want_flow = false;
      }
      const bool use_flow = want_flow;
      std::vector<ClassDev> cdev(pl.classes.size());
      h->class_fast.assign(pl.classes.size(), 0);
      // the streaming shear kernels take a bounded number of classes / chunks per tiling and z direction
      std::vector<std::array<int, 2>> fast_chunk_count(pl.tilings.size(), std::array<int, 2>{0, 0});
      std::vector<std::array<int, 2>> fast_class_count(pl.tilings.size(), std::array<int, 2>{0, 0});
      h->d_pos_of.assign(pl.classes.size(), nullptr);
      h->class_mats_s.assign(pl.classes.size(), nullptr);
      h->class_mats_c.assign(pl.classes.size(), nullptr);
      h->mat_bytes = h->nmat <= 256 ? 1 : 4;
      for (size_t ci = 0; ci < pl.classes.size(); ci++) {
         const ClassPlan& cp = pl.classes[ci];
         ClassDev& cd = cdev[ci];
         cd.S = cp.S; cd.zdir = cp.zdir; cd.ring = cp.ring; cd.tiles = cp.tiles ? 1 : 0; cd.npatch = cp.npatch;
         cd.nsteps = cp.nsteps; cd.gm = gm; cd.nsm = nsm_of(cp); cd.mat_bytes = 4; cd.mats_c = nullptr;
         h->class_fast[ci] = base_ok && (use_flow ? cp.fast_flow : cp.fast);
         if (h->class_fast[ci]) {
            bool any_owned = false;
            for (size_t c = 0; c < pl.chunks.size(); c++) any_owned |= (pl.chunks[c].cls == (int)ci && chunk_owned[c]);
            if (!any_owned) h->class_fast[ci] = 0;
         }
         if (h->class_fast[ci]) {
            int& cnt = fast_chunk_count[cp.tiling][cp.zdir >= 0 ? 0 : 1];
            int mine = 0;
            for (size_t c = 0; c < pl.chunks.size(); c++) mine += (pl.chunks[c].cls == (int)ci && chunk_owned[c]);
            int& ccnt = fast_class_count[cp.tiling][cp.zdir >= 0 ? 0 : 1];
            if (cnt + mine > SHEAR_MAX_PER_PASS || ccnt + 1 > shear_max_classes()) h->class_fast[ci] = 0;
            else { cnt += mine; ccnt++; }
         }
         cd.mats_s = nullptr;
         if (!(use_flow && h->class_fast[ci])) {
            int32_t* d_ms;
            if (dev_upload(h, &d_ms, class_step_map(pl, cp, mats))) return 1;
            cd.mats_s = d_ms;
            h->class_mats_s[ci] = d_ms;
         }
         int32_t *d_cell_of, *d_patch_nlev, *d_in_src, *d_rout, *d_ls_of = nullptr;
         uint16_t* d_lvl; Vec2 *d_out_vec, *d_in_vec;
         if (dev_upload(h, &d_cell_of, cp.cell_of) || dev_upload(h, &d_lvl, cp.lvl) ||
             dev_upload(h, &d_patch_nlev, cp.patch_nlev) || dev_upload(h, &d_out_vec, cp.out_vec) ||
             dev_upload(h, &d_in_src, cp.in_src) || dev_upload(h, &d_in_vec, cp.in_vec) ||
             dev_upload(h, &d_rout, cp.rout) || dev_upload(h, &h->d_pos_of[ci], cp.pos_of)) return 1;
         if (h->nls > 0 && dev_upload(h, &d_ls_of, ls_of[ci])) return 1;
         uint16_t* d_hidx;
         if (dev_upload(h, &d_hidx, cp.in_hidx)) return 1;
         cd.in_hidx = d_hidx;
         uint8_t* d_eidx;
         if (dev_upload(h, &d_eidx, cp.eidx)) return 1;
         cd.eidx = d_eidx;
         cd.q_sheared = nullptr;
         if (h->class_fast[ci]) {
            double* d_qs;
            const int64_t nq = (int64_t)nblk * cp.npatch * nsm_of(cp) * PS;
            if (dev_alloc(h, &d_qs, nq)) return 1;
            SN_CUDA(h, cudaMemsetAsync(d_qs, 0, (size_t)nq * sizeof(double), h->stream));
            cd.q_sheared = d_qs;
            const int mb = h->mat_bytes;
            uint8_t* d_mc;
            if (dev_upload(h, &d_mc, class_cyclic_map(pl, cp, mats, mb))) return 1;
            cd.mat_bytes = mb;
            cd.mats_c = d_mc;
            h->class_mats_c[ci] = d_mc;
         }
         cd.cell_of = d_cell_of; cd.lvl = d_lvl; cd.patch_nlev = d_patch_nlev;
         cd.out_vec = (const double2*)d_out_vec; cd.in_src = d_in_src; cd.in_vec = (const double2*)d_in_vec;
         cd.rout = d_rout; cd.ls_of = d_ls_of;
      }
      if (dev_upload(h, &h->d_class_pos_of, h->d_pos_of.data(), (int64_t)h->d_pos_of.size())) return 1;

      // chunks
      std::vector<ChunkDev> chdev(pl.chunks.size());
      std::vector<int32_t> fast_chunks;
      h->dir_chunk.assign(h->M, -1); h->dir_d.assign(h->M, -1);
      for (size_t c = 0; c < pl.chunks.size(); c++) {
         const Chunk& ch = pl.chunks[c];
         ChunkDev& cd = chdev[c];
         std::memset(&cd, 0, sizeof(cd));
         cd.cls = ch.cls; cd.nd = ch.nd;
         for (int d = 0; d < DT_MAX; d++) {
            const int m = d < ch.nd ? ch.m[d] : -1;
            cd.m[d] = m;
            if (m >= 0) {
               cd.mux[d] = quad->directions[3*m]; cd.muy[d] = quad->directions[3*m+1];
               cd.muz_abs[d] = std::fabs(quad->directions[3*m+2]); cd.w[d] = quad->weights[m];
               for (int ax = 0; ax < 3; ax++) cd.mrefl[d][ax] = quad->reflected[3*m+ax];
               if (chunk_owned[c]) { h->dir_chunk[m] = (int)c; h->dir_d[m] = d; }
            }
         }
         cd.psi = chunk_owned[c] ? h->d_psi + psi_off[c] : nullptr;
         h->chunk_psi.push_back(cd.psi);
         cd.phi_part = nullptr;
         if (chunk_owned[c] && h->class_fast[ch.cls]) {
            const ClassPlan& cpc = pl.classes[ch.cls];
            if (dev_alloc(h, &cd.phi_part, (int64_t)nblk * cpc.npatch * nsm_of(cpc) * PS)) return 1;
            fast_chunks.push_back((int32_t)c);
         }
      }
      h->tilings.assign(pl.tilings.size(), TilingDev{});
      h->nfast_classes = 0; h->nfast_chunks = (int)fast_chunks.size();
      int num_sms = 148;
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, h->device);
      for (size_t tg = 0; tg < pl.tilings.size(); tg++) {
         TilingDev& td = h->tilings[tg];
         td.npatch = pl.tilings[tg].npatch;
         std::vector<int32_t> cls_list, chunk_list;
         for (size_t ci = 0; ci < pl.classes.size(); ci++)
            if (h->class_fast[ci] && pl.classes[ci].tiling == (int)tg) cls_list.push_back((int32_t)ci);
         for (int32_t c : fast_chunks) if (pl.classes[pl.chunks[c].cls].tiling == (int)tg) chunk_list.push_back(c);
         // +z chunks first: the un-shear kernel processes one z direction per pass, 8 chunks at a time
         std::stable_sort(chunk_list.begin(), chunk_list.end(), [&](int32_t a, int32_t b) {
            return (pl.classes[pl.chunks[a].cls].zdir < 0) < (pl.classes[pl.chunks[b].cls].zdir < 0); });
         td.nclasses = (int)cls_list.size(); td.nchunks = (int)chunk_list.size();
         td.nplus = 0;
         for (int32_t c : chunk_list) if (pl.classes[pl.chunks[c].cls].zdir >= 0) td.nplus++;
         td.zsplit = unshear_zsplit(td.npatch * h->Gown, pl.nz, num_sms);
         h->nfast_classes += td.nclasses;
         if (dev_upload(h, &td.d_classes, cls_list) || dev_upload(h, &td.d_chunks, chunk_list)) return 1;
         if (tg > 0 && td.nclasses > 0) {     // slot of this tiling -> base slot (the classes on it share the map)
            if (dev_upload(h, &td.d_cell_of, pl.classes[cls_list[0]].cell_of)) return 1;
         }
      }
      {
         std::vector<int32_t> a(h->dir_chunk.begin(), h->dir_chunk.end()), b(h->dir_d.begin(), h->dir_d.end());
         if (dev_upload(h, &h->d_dir_chunk, a) || dev_upload(h, &h->d_dir_d, b)) return 1;
      }

      // launch schedule.  Tile classes: one dataflow launch per chunk size, tasks in ticket order =
      // wavefront number first, so that every class (octant) advances together and a task's upwind
      // tasks always hold lower tickets.  Other classes (and wave_launch = 1): per wave, tasks grouped
      // by kernel variant, one launch each.
      std::vector<Task> all;
      h->groups.clear(); h->flows.clear();
      if (use_flow) {
         for (int dtf = 2; dtf <= 2 * DT_MAX + 1; dtf++) {
            const int dt = dtf / 2, fin = (dtf & 1) ? 3 : 2;
            auto in_launch = [&](int chunk) {
               const ClassPlan& cpc = pl.classes[pl.chunks[chunk].cls];
               return h->class_fast[pl.chunks[chunk].cls] && pl.chunks[chunk].nd == dt && (needs_fin3(cpc) ? 3 : 2) == fin;
            };
            const size_t first = all.size();
            for (size_t w = 0; w < pl.waves.size(); w++)
               for (const Task& t : pl.waves[w])
                  if (in_launch(t.chunk) && h->gloc[t.group] % gm == 0)
                     all.push_back(Task{t.chunk, h->gloc[t.group] / gm, t.patch, 0});   // group field = block
            if (all.size() > first) {
               if (h->flows.size() >= 16) SN_FAIL(h, "internal: too many dataflow launches");
               FlowLaunch fl{dt, fin, (int64_t)first, (int)(all.size() - first), {}, 0};
               for (size_t c = 0; c < pl.chunks.size(); c++) {
                  if (!(chunk_owned[c] && in_launch((int)c))) continue;
                  chdev[c].flow_slot = fl.nch++;
                  for (int d = 0; d < DT_MAX; d++) {
                     fl.mw.push_back(pl.has_z ? chdev[c].muz_abs[d] * (h->uniform_dz ? 1.0 / mesh->dz[0] : 1.0) : 0.0);
                     fl.mw.push_back(chdev[c].w[d]);
                  }
               }
               h->flows.push_back(fl);
            }
         }
         if (!h->flows.empty()) {
            int np_max = pl.npatch_b;
            for (const Tiling& tg : pl.tilings) np_max = std::max(np_max, tg.npatch);
            h->np_stride = np_max;
            h->flow_ctl_count = 16 + (int64_t)pl.chunks.size() * nblk * np_max;
            if (dev_alloc(h, &h->d_flow_ctl, h->flow_ctl_count)) return 1;
         }
      }
      for (size_t w = 0; w < pl.waves.size(); w++) {
         std::vector<Task> tasks;
         for (const Task& t : pl.waves[w])
            if (!(use_flow && h->class_fast[pl.chunks[t.chunk].cls])) tasks.push_back(t);
         auto variant = [&](const Task& t) {
            const Chunk& ch = pl.chunks[t.chunk]; const ClassPlan& cp = pl.classes[ch.cls];
            return std::make_tuple(ch.cls, h->class_fast[ch.cls] ? 1 : 0, dt_template(ch.nd), cp.fin <= 2 ? 2 : FIN_MAX, cp.ring);
         };
         std::stable_sort(tasks.begin(), tasks.end(), [&](const Task& a, const Task& b) { return variant(a) < variant(b); });
         size_t i = 0;
         while (i < tasks.size()) {
            size_t j = i;
            while (j < tasks.size() && variant(tasks[j]) == variant(tasks[i])) j++;
            auto v = variant(tasks[i]);
            h->groups.push_back(LaunchGroup{(int)w, std::get<0>(v), std::get<1>(v), std::get<2>(v), std::get<3>(v), std::get<4>(v), h->extras,
                                            (int64_t)all.size() + (int64_t)i, (int)(j - i)});
            i = j;
         }
         all.insert(all.end(), tasks.begin(), tasks.end());
      }
      h->groups_generic = 0;
      for (const LaunchGroup& lg : h->groups) if (lg.kind != 1) h->groups_generic++;
      h->use_graph = h->groups.size() >= 32 && !h->opts.no_graph;
      if (!h->opts.store_psi && !h->groups.empty())
         SN_FAIL(h, "store_psi = 0 needs every ordering class on the dataflow tile kernel (Cartesian or lattice mesh, "
                    "no least-squares / delta < 1 term, wave_launch = 0)");
      // dataflow classes on structured tiles: neighbouring patches read the perimeter lanes of the psi rows
      // themselves (no edge copies); the wavefront-launched kernels keep the copies
      for (size_t ci = 0; ci < pl.classes.size(); ci++) {
         cdev[ci].inline_edges = (use_flow && h->class_fast[ci] && pl.classes[ci].inline_ok && h->opts.inline_edges) ? 1 : 0;
         cdev[ci].pstride = cdev[ci].inline_edges ? PS : PSX;
      }
      if (dev_upload(h, &h->d_classes, cdev)) return 1;
      if (dev_upload(h, &h->d_chunks, chdev)) return 1;
      if (dev_upload(h, &h->d_tasks, all)) return 1;
      // reflective / LS problems read what another class wrote in the previous sweep only, so classes
      // stay independent within a sweep; streams are used unless the option turns them off
      h->multi_stream = (pl.classes.size() > 1 || h->flows.size() > 1) && !h->opts.single_stream;

      // fused tail of the iteration: second iterate buffer (one GPU: here; sharded: pampa_sn_comm_init) and the
      // partials of the un-shear CTAs
      if (h->groups_generic == 0 && h->tilings[0].nchunks > 0) {
         if (dev_alloc(h, &h->d_fuse_partials, 5LL * pl.npatch_b * h->G * UNSHEAR_ZSPLIT_MAX)) return 1;
         if (h->opts.num_ranks == 1) {
            if (dev_alloc(h, &h->d_phi_buf[1], nphi)) return 1;
            SN_CUDA(h, cudaMemsetAsync(h->d_phi_buf[1], 0, (size_t)nphi * sizeof(double), h->stream));
         }
      }

      // reduction scratch and iteration state
      h->nblocks_reduce = (int)std::min<int64_t>(((int64_t)nz * Sb + 255) / 256, 148 * 8);
      if (dev_alloc(h, &h->d_partials, 5LL * h->nblocks_reduce) || dev_alloc(h, &h->d_sc, 1) ||
          dev_alloc(h, &h->d_sums, 8) || dev_alloc(h, &h->d_sums_all, 5LL * std::max(1, h->opts.num_ranks))) return 1;
      ReduceScalars sc0{}; sc0.keff = 1.0;
      SN_CUDA(h, cudaMemcpyAsync(h->d_sc, &sc0, sizeof(sc0), cudaMemcpyHostToDevice, h->stream));
      SN_CUDA(h, cudaMemsetAsync(h->d_phi, 0, (size_t)nphi * sizeof(double), h->stream));
      launch_fill_phi(h->d_phi_new, h->d_mats, 1.0, h->G, (int64_t)nz * Sb, h->stream);
      if (do_reduce(h, 0)) return 1;
      if (sync_scalars(h)) return 1;
      return check_async(h, "initialisation");
   };
   if (body()) return fail(1);
   if (h->opts.verbose)
      std::printf("pampa_sn: %d xy cells x %d layers, %d groups, %d directions, %zu classes (%d tiled), %zu chunks, "
                  "%zu sweep launches, %.1f MB on device %d\n", pl.nxy, pl.nz, h->G, h->M, pl.classes.size(),
                  pl.tile_classes, pl.chunks.size(), h->groups.size() + h->flows.size(), h->device_bytes / 1.0e6, h->device);
   *out = hp.release();
   return 0;
}

int pampa_sn_destroy(pampa_sn_handle* h) {
   if (!h) return 0;
   cudaSetDevice(h->device);
   if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
   if (h->stream) cudaStreamSynchronize(h->stream);
   for (void* p : h->ipc_mapped) cudaIpcCloseMemHandle(p);
   for (void* p : h->allocs) cudaFree(p);
   if (h->d_stage) cudaFree(h->d_stage);
   if (h->h_aa_ring) cudaFreeHost(h->h_aa_ring);
   for (int i = 0; i < pampa_sn_handle::NSTREAMS; i++) {
      if (h->cls_stream[i]) { cudaStreamSynchronize(h->cls_stream[i]); cudaStreamDestroy(h->cls_stream[i]); }
      if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
   }
   for (int i = 0; i < 2; i++) if (h->graph_exec[i]) cudaGraphExecDestroy(h->graph_exec[i]);
   if (h->ev_fork) cudaEventDestroy(h->ev_fork);
   if (h->ev0) cudaEventDestroy(h->ev0);
   if (h->ev1) cudaEventDestroy(h->ev1);
   if (h->stream) cudaStreamDestroy(h->stream);
   delete h;
   return 0;
}

int pampa_sn_update_xs(pampa_sn_handle* h, const pampa_sn_xs* xs) {
   SN_CUDA(h, cudaSetDevice(h->device));
   SN_CUDA(h, cudaStreamSynchronize(h->stream));
   return upload_xs(h, xs, false);
}

// New cross-section tables together with a new cell -> table-row map (temperature feedback changes which
// (material, temperature) pair a cell uses): the material maps of every kernel are rebuilt in place.  Fails
// when the new number of rows does not fit what the handle was planned for -- the caller then re-creates it.
int pampa_sn_update_materials(pampa_sn_handle* h, const pampa_sn_xs* xs, const int32_t* materials) {
   SN_CUDA(h, cudaSetDevice(h->device));
   const Plan& pl = h->plan;
   const int nm = xs->num_materials;
   if (xs->num_groups != h->G) SN_FAIL(h, "cross-section table shape changed");
   for (int64_t i = 0; i < (int64_t)pl.nz * pl.nxy; i++)
      if (materials[i] < 0 || materials[i] >= nm) SN_FAIL(h, "wrong material index");
   bool any_fast = false;
   for (char f : h->class_fast) any_fast |= f != 0;
   if ((h->mat_bytes == 1 && nm > 256) || (any_fast && ((int64_t)h->gm * nm > 2048 || nm > 4096)))
      SN_FAIL(h, "the new material table does not fit the sweep plan of this handle: re-create it");
   SN_CUDA(h, cudaStreamSynchronize(h->stream));
   if (upload_xs(h, xs, false, true)) return 1;
   const std::vector<int32_t> mats = base_material_map(pl, materials);
   SN_CUDA(h, cudaMemcpy(h->d_mats, mats.data(), mats.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
   for (size_t ci = 0; ci < pl.classes.size(); ci++) {
      if (h->class_mats_s[ci]) {
         const std::vector<int32_t> ms = class_step_map(pl, pl.classes[ci], mats);
         SN_CUDA(h, cudaMemcpy(h->class_mats_s[ci], ms.data(), ms.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      }
      if (h->class_mats_c[ci]) {
         const std::vector<uint8_t> mc = class_cyclic_map(pl, pl.classes[ci], mats, h->mat_bytes);
         SN_CUDA(h, cudaMemcpy(h->class_mats_c[ci], mc.data(), mc.size(), cudaMemcpyHostToDevice));
      }
   }
   return 0;
}

int pampa_sn_source(pampa_sn_handle* h, double keff) {
   SN_CUDA(h, cudaSetDevice(h->device));
   SN_CUDA(h, cudaMemcpyAsync(&h->d_sc->keff, &keff, sizeof(double), cudaMemcpyHostToDevice, h->stream));
   SN_CUDA(h, cudaEventRecord(h->ev0, h->stream));
   if (do_source(h)) return 1;
   SN_CUDA(h, cudaEventRecord(h->ev1, h->stream));
   SN_CUDA(h, cudaStreamSynchronize(h->stream));
   float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->last_source_ms = ms;
   return check_async(h, "the source kernel");
}

int pampa_sn_sweep(pampa_sn_handle* h) {
   SN_CUDA(h, cudaSetDevice(h->device));
   SN_CUDA(h, cudaEventRecord(h->ev0, h->stream));
   if (do_sweep(h)) return 1;
   SN_CUDA(h, cudaEventRecord(h->ev1, h->stream));
   SN_CUDA(h, cudaStreamSynchronize(h->stream));
   float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->last_sweep_ms = ms;
   return check_async(h, "the sweep kernel");
}

int pampa_sn_reduce(pampa_sn_handle* h, double* production, double* power, double* dphi_rel) {
   SN_CUDA(h, cudaSetDevice(h->device));
   SN_CUDA(h, cudaEventRecord(h->ev0, h->stream));
   if (do_reduce(h, 0)) return 1;
   SN_CUDA(h, cudaEventRecord(h->ev1, h->stream));
   if (sync_scalars(h)) return 1;
   float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->last_reduce_ms = ms;
   if (production) *production = h->sc.production;
   if (power) *power = h->sc.power;
   if (dphi_rel) *dphi_rel = h->sc.phi2 > 0 ? std::sqrt(h->sc.dphi2 / h->sc.phi2) : 0.0;
   return check_async(h, "the reduction kernel");
}

int pampa_sn_iterate(pampa_sn_handle* h, int32_t iterations, double* keff) {
   SN_CUDA(h, cudaSetDevice(h->device));
   for (int it = 0; it < iterations; it++) {
      if (do_source(h) || do_sweep(h, 1) || do_reduce(h, 1)) return 1;
   }
   if (sync_scalars(h)) return 1;
   h->keff = h->sc.keff;
   if (keff) *keff = h->sc.keff;
   return check_async(h, "the source iteration");
}

int pampa_sn_iterate_timed(pampa_sn_handle* h, int32_t iterations, double* keff, double* total_ms,
                           double* sweep_ms) {
   SN_CUDA(h, cudaSetDevice(h->device));
   // per iteration: [start] source [a] sweep [b] reduce + exchange [c]
   std::vector<cudaEvent_t> ev(3 * (size_t)iterations + 2);
   for (auto& e : ev) SN_CUDA(h, cudaEventCreate(&e));
   int rc = 0;
   std::vector<cudaEvent_t> kev, xev;
   SN_CUDA(h, cudaStreamSynchronize(h->stream));
   h->kernel_events = &kev;
   h->exchange_events = &xev;
   cudaEventRecord(ev[0], h->stream);
   for (int it = 0; it < iterations && !rc; it++) {
      rc = do_source(h);
      cudaEventRecord(ev[2 + 3 * it], h->stream);
      if (!rc) rc = do_sweep(h, 1);
      cudaEventRecord(ev[3 + 3 * it], h->stream);
      if (!rc) rc = do_reduce(h, 1);
      cudaEventRecord(ev[4 + 3 * it], h->stream);
   }
   cudaEventRecord(ev[1], h->stream);
   h->kernel_events = nullptr;
   h->exchange_events = nullptr;
   if (!rc) rc = sync_scalars(h);
   h->timed_kernel_ms = 0; h->timed_exchange_ms = 0; h->timed_source_ms = 0; h->timed_reduce_ms = 0;
   for (size_t i = 0; i + 1 < kev.size() && !rc; i += 2) {
      float t = 0; cudaEventElapsedTime(&t, kev[i], kev[i + 1]); h->timed_kernel_ms += t;
   }
   for (size_t i = 0; i + 1 < xev.size() && !rc; i += 2) {
      float t = 0; cudaEventElapsedTime(&t, xev[i], xev[i + 1]); h->timed_exchange_ms += t;
   }
   for (auto& e : kev) cudaEventDestroy(e);
   for (auto& e : xev) cudaEventDestroy(e);
   if (!rc) {
      float ms = 0, sw = 0, t = 0;
      cudaEventElapsedTime(&ms, ev[0], ev[1]);
      for (int it = 0; it < iterations; it++) {
         cudaEventElapsedTime(&t, ev[2 + 3 * it], ev[3 + 3 * it]); sw += t;
         cudaEventElapsedTime(&t, it == 0 ? ev[0] : ev[4 + 3 * (it - 1)], ev[2 + 3 * it]); h->timed_source_ms += t;
         cudaEventElapsedTime(&t, ev[3 + 3 * it], ev[4 + 3 * it]); h->timed_reduce_ms += t;
      }
      if (total_ms) *total_ms = ms;
      if (sweep_ms) *sweep_ms = sw;
      h->keff = h->sc.keff;
      if (keff) *keff = h->sc.keff;
   }
   for (auto& e : ev) cudaEventDestroy(e);
   if (rc) return 1;
   return check_async(h, "the source iteration");
}

int pampa_sn_solve_keff(pampa_sn_handle* h, double tol_k, double tol_phi, int32_t max_it, double power,
                        double* keff, int32_t* iterations) {
   SN_CUDA(h, cudaSetDevice(h->device));
   const Plan& pl = h->plan;
   const int64_t nslab = (int64_t)pl.nz * pl.Sb, nphi = h->G * nslab;
   int depth = h->opts.anderson_depth == 0 ? AA_SLOTS - 1 : h->opts.anderson_depth;     // < 0: plain power iteration
   depth = std::min(depth, AA_SLOTS - 1);
   // the lagged LS boundary term depends on the angular flux of the previous sweep, which is not
   // part of the mixed state: keep the history short there (1-D / 2-D problems only)
   if (h->nls > 0 && depth > 3) depth = 3;
   // The deferred correction of delta < 1 is formed from the previous sweep's angular flux too, and it is not a
   // small term (the iteration matrix of the lagged part has a spectral radius of ~0.9 at delta = 0.1): mixing
   // iterates whose hidden state differs can diverge (seen on a slab between reflective boundaries), so those
   // problems run the plain iteration unless an acceleration depth is asked for explicitly
   if (h->d_corr) depth = h->opts.anderson_depth > 0 ? std::min(depth, 3) : -1;
   int it = 0;
   bool converged = false;
   h->psi_scale_factor = 1.0;
   double power_integral = 0.0, min_phi = 0.0;
   SN_CUDA(h, cudaEventRecord(h->ev0, h->stream));

   // plain power iteration; the convergence test reads the scalars of iteration i while i + 1 is in the queue
   auto plain_iteration = [&]() -> int {
      while (it < max_it) {
         if (do_source(h) || do_sweep(h, 1) || do_reduce(h, 1)) return 1;
         it++;
         if (sync_scalars(h)) return 1;
         const double dphi = h->sc.phi2 > 0 ? std::sqrt(h->sc.dphi2 / h->sc.phi2) : 0.0;
         if (!(h->sc.keff == h->sc.keff)) SN_FAIL(h, "the power iteration diverged (NaN)");
         if (it > 1 && std::fabs(h->sc.dk) < tol_k && dphi < tol_phi) { converged = true; break; }
      }
      power_integral = h->sc.power; min_phi = h->sc.min_phi;
      h->psi_scale_factor = 1.0;
      return 0;
   };

   if (depth < 1) {
      if (plain_iteration()) return 1;
   } else {
      // Anderson-accelerated fixed-point iteration on x = phi (constant production) and k.  All the bookkeeping
      // (Gram matrix, window, weights, k) is in the device-resident AAState: the host enqueues iteration after
      // iteration and reads the convergence flag of iteration i while iteration i + 1 is already in the queue, so
      // the device never waits for the host (one iteration may run past convergence; it is a valid iterate).
      const int slots = depth + 1;
      if (h->aa_slots < slots) {
         for (int j = h->aa_slots; j < slots; j++)
            if (dev_alloc(h, &h->aa_f[j], nphi) || dev_alloc(h, &h->aa_g[j], nphi) ||
                dev_alloc(h, &h->aa_b[j], h->bnd_count) || dev_alloc(h, &h->aa_bz[j], h->bndz_count)) return 1;
         if (!h->d_aa_partials && (dev_alloc(h, &h->d_aa_partials, (int64_t)AA_SLOTS * h->nblocks_reduce) ||
                                   dev_alloc(h, &h->d_aa_dots, AA_SLOTS) || dev_alloc(h, &h->d_aa_state, 1))) return 1;
         h->aa_slots = slots;
      }
      if (!h->h_aa_ring) SN_CUDA(h, cudaHostAlloc((void**)&h->h_aa_ring, 2 * sizeof(AAState), cudaHostAllocDefault));
      const int owned_only = (h->comm && h->group_gather) ? 1 : 0;
      // phi_new must be cleared for the next sweep only where something accumulates into it (generic kernels) or a
      // rank of an angle-sharded run may own no chunk at all; otherwise the first un-shear pass overwrites it
      const int zero_new = (h->groups_generic == 0 && h->nfast_chunks > 0 && (!h->comm || h->group_gather)) ? 0 : 1;
      if (sync_scalars(h)) return 1;
      if (!(h->sc.production > 0.0)) SN_FAIL(h, "zero fission production: no fissile material in the mesh");
      AAState st0;
      std::memset(&st0, 0, sizeof(st0));
      for (int j = 0; j < AA_SLOTS; j++) st0.age[j] = -1;
      st0.kn = h->sc.keff; st0.prod_x = h->sc.production; st0.best = 1.0e300; st0.inv = 1.0;
      st0.tol_k = tol_k; st0.tol_phi = tol_phi; st0.slots = slots;
      st0.ncells = (double)pl.nxy * pl.nz * h->G;
      // the first iterations only settle k and the gross flux shape: mixing them in could slow the acceleration
      // down, so the history may start after `aa_start` plain steps (measured: no benefit at depth 7, default 0)
      { const char* e = std::getenv("PAMPA_SN_AA_START"); st0.aa_start = e ? std::atoi(e) : 0; }
      SN_CUDA(h, cudaMemcpyAsync(h->d_aa_state, &st0, sizeof(st0), cudaMemcpyHostToDevice, h->stream));
      SN_CUDA(h, cudaStreamSynchronize(h->stream));     // st0 is on the stack
      double* fptr[AA_SLOTS]; double* gptr[AA_SLOTS]; double* bptr[AA_SLOTS]; double* bzptr[AA_SLOTS];
      for (int j = 0; j < AA_SLOTS; j++) {
         fptr[j] = h->aa_f[j < slots ? j : 0]; gptr[j] = h->aa_g[j < slots ? j : 0];
         bptr[j] = h->aa_b[j < slots ? j : 0]; bzptr[j] = h->aa_bz[j < slots ? j : 0];
      }
      cudaEvent_t ev[2] = {nullptr, nullptr};
      for (auto& e : ev) SN_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      int rc = 0;
      bool failed = false;
      while (it < max_it && !rc) {
         if ((rc = do_source(h) || do_sweep(h, 2) || reduce_sums(h, 0))) break;
         it++;
         launch_aa_begin(h->d_aa_state, h->d_sums, h->stream);
         launch_aa_store(h->d_phi, h->d_phi_new, h->d_gloc, owned_only, zero_new, h->G, nslab, fptr, gptr, h->d_aa_state,
                         h->d_aa_partials, h->nblocks_reduce, h->d_aa_dots, h->stream);
         h->launches += 3;
         // the lagged boundary fluxes are part of the fixed-point state: same normalisation, same mixing
         launch_scale_copy_slot(bptr, h->d_bnd[h->bnd_cur], h->d_aa_state, h->bnd_count, h->stream);
         launch_scale_copy_slot(bzptr, h->d_bndz[h->bnd_cur], h->d_aa_state, h->bndz_count, h->stream);
         if (owned_only) {
            int r = g_nccl.AllReduce(h->d_aa_dots, h->d_aa_dots, AA_SLOTS, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
            if (r != 0) { rc = nccl_fail(h, r, "the Anderson allreduce"); break; }
         }
         launch_aa_solve(h->d_aa_state, h->d_aa_dots, h->d_sc, h->stream);
         // group-sharded with peer access: the next iterate goes into the other buffer here and on every peer
         // (the mix pass stores it), and a one-value collective is the barrier before anybody reads it
         const bool push = owned_only && h->p2p;
         const int out = push ? 1 - h->phi_cur : h->phi_cur;
         launch_aa_mix(h->d_phi_buf[out], h->d_mats, h->d_gloc, owned_only, h->G, nslab, gptr, h->d_aa_state,
                       h->nblocks_reduce, push ? h->peer_phi[out] : nullptr, h->npeers, h->stream);
         if (push) { h->phi_cur = out; h->d_phi = h->d_phi_buf[out]; }
         h->launches += 2;
         if (h->bnd_count > 0) launch_vec_mix(h->d_bnd[h->bnd_cur], bptr, h->d_aa_state, h->bnd_count, h->stream);
         if (h->bndz_count > 0) launch_vec_mix(h->d_bndz[h->bnd_cur], bzptr, h->d_aa_state, h->bndz_count, h->stream);
         if (push) {
            int r = g_nccl.AllReduce(h->d_aa_dots, h->d_aa_dots, 1, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);   // barrier
            if (r != 0) { rc = nccl_fail(h, r, "the barrier after the peer-to-peer exchange"); break; }
         } else if ((rc = gather_phi(h))) break;
         // snapshot of the state after this iteration; looked at one iteration later
         cudaMemcpyAsync(&h->h_aa_ring[it & 1], h->d_aa_state, sizeof(AAState), cudaMemcpyDeviceToHost, h->stream);
         cudaEventRecord(ev[it & 1], h->stream);
         if (it >= 2) {
            cudaEventSynchronize(ev[(it - 1) & 1]);
            const AAState& prev = h->h_aa_ring[(it - 1) & 1];
            if (h->opts.verbose > 1)
               std::printf("pampa_sn: it %d k %.12f dk %.3e res %.3e min phi %.3e negatives %d hold %d\n", prev.it, prev.kn,
                           prev.dk, prev.res, prev.min_phi, prev.negatives, prev.hold);
            if (prev.failed) { failed = true; break; }
            if (prev.converged) { converged = true; break; }
         }
      }
      cudaError_t es = cudaStreamSynchronize(h->stream);
      for (auto& e : ev) if (e) cudaEventDestroy(e);
      if (rc) return 1;
      if (es != cudaSuccess) SN_FAIL(h, std::string("CUDA error in the k-eff iteration: ") + cudaGetErrorString(es));
      if (it > 0 && (h->h_aa_ring[it & 1].failed || failed)) {
         // the accelerated iteration broke down (NaN or a non-positive production after mixing): start again from a
         // flat flux with the plain iteration, which cannot
         if (h->opts.verbose) std::printf("pampa_sn: accelerated iteration failed after %d iterations, restarting plain\n", it);
         SN_CUDA(h, cudaMemsetAsync(h->d_phi, 0, (size_t)nphi * sizeof(double), h->stream));
         launch_fill_phi(h->d_phi_new, h->d_mats, 1.0, h->G, nslab, h->stream);
         if (h->d_psi) SN_CUDA(h, cudaMemsetAsync(h->d_psi, 0, (size_t)h->psi_count * sizeof(double), h->stream));
         for (int b = 0; b < 2; b++) {
            if (h->d_bnd[b]) SN_CUDA(h, cudaMemsetAsync(h->d_bnd[b], 0, (size_t)h->bnd_count * sizeof(double), h->stream));
            if (h->d_bndz[b]) SN_CUDA(h, cudaMemsetAsync(h->d_bndz[b], 0, (size_t)h->bndz_count * sizeof(double), h->stream));
         }
         if (h->d_ls_rhs) SN_CUDA(h, cudaMemsetAsync(h->d_ls_rhs, 0, (size_t)h->M * h->G * h->nls * sizeof(double), h->stream));
         const double one = 1.0;
         SN_CUDA(h, cudaMemcpyAsync(&h->d_sc->keff, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream));
         SN_CUDA(h, cudaStreamSynchronize(h->stream));
         if (do_reduce(h, 0, false)) return 1;
         if (plain_iteration()) return 1;
      } else if (it > 0) {
         const AAState& last = h->h_aa_ring[it & 1];     // the state the device fields are in
         converged = converged || last.converged != 0;
         h->sc.keff = last.kn;
         h->psi_scale_factor = last.inv;
         power_integral = last.power_integral; min_phi = last.min_phi;
         // A mixed iterate is a combination with weights of both signs: where the flux is many decades below its
         // maximum (deep in a reflector, far corners) it can be negative by less than the tolerance, and so is the
         // sweep result there.  The eigenvector is positive, and the plain iteration keeps what is positive positive:
         // a few unaccelerated iterations remove the undershoot (the reference rejects any negative flux,
         // src/NeutronicSolver.cxx:67).
         for (int polish = 0; converged && min_phi < 0.0 && polish < 200 && it < max_it; polish++) {
            if (do_source(h) || do_sweep(h, 1) || do_reduce(h, 1)) return 1;
            it++;
            if (sync_scalars(h)) return 1;
            power_integral = h->sc.power; min_phi = h->sc.min_phi;
            h->psi_scale_factor = 1.0;
         }
      }
   }
   SN_CUDA(h, cudaEventRecord(h->ev1, h->stream));
   SN_CUDA(h, cudaStreamSynchronize(h->stream));
   { float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->last_solve_ms = ms; }
   if (check_async(h, "the k-eff iteration")) return 1;
   h->keff = h->sc.keff;
   if (keff) *keff = h->sc.keff;
   if (iterations) *iterations = it;
   if (power_integral == 0.0) SN_FAIL(h, "zero fission power: no fissile material in the mesh");
   h->scale = power / power_integral;
   h->solved = true;
   if (min_phi * h->scale < 0.0) SN_FAIL(h, "negative values in the scalar-flux solution");
   if (!converged) SN_FAIL(h, "the power iteration did not converge in " + std::to_string(max_it) + " iterations");
   return 0;
}

// Partitioned fields (opts.partition_fields in a sharded run): this rank's window of cells
static void cell_window(const pampa_sn_handle* h, int64_t* i0, int64_t* ni, int64_t* per_rank) {
   const int64_t N = (int64_t)h->plan.nxy * h->plan.nz;
   if (h->opts.partition_fields && h->opts.num_ranks > 1) {
      const int64_t c = (N + h->opts.num_ranks - 1) / h->opts.num_ranks;
      *i0 = std::min(N, c * h->opts.rank);
      *ni = std::min(N, c * (h->opts.rank + 1)) - *i0;
      if (per_rank) *per_rank = c;
   } else { *i0 = 0; *ni = N; if (per_rank) *per_rank = N; }
}

int64_t pampa_sn_field_size(const pampa_sn_handle* h, const char* name) {
   const int64_t N = (int64_t)h->plan.nxy * h->plan.nz;
   const std::string s(name);
   int64_t i0, ni;
   cell_window(h, &i0, &ni, nullptr);
   if (s == "scalar-flux" || s == "flux-moments") return ni * h->G;
   if (s == "angular-flux") return N * h->G * h->M;
   if (s == "power" || s == "production-rate") return ni;
   if (s == "temperature" || s == "delayed-source") return N;
   if (s == "keff" || s == "angular-flux-min") return 1;
   return -1;
}

int pampa_sn_get(pampa_sn_handle* h, const char* name, double* out) {
   SN_CUDA(h, cudaSetDevice(h->device));
   const std::string s(name);
   const Plan& pl = h->plan;
   const int64_t N = (int64_t)pl.nxy * pl.nz;
   const int64_t count = pampa_sn_field_size(h, name);
   if (count < 0) SN_FAIL(h, "unable to find field '" + s + "'");
   if (s == "angular-flux" && !h->opts.store_psi) SN_FAIL(h, "the angular flux is not kept in memory (store_psi = 0)");
   if (s == "keff") {
      if (sync_scalars(h)) return 1;
      out[0] = h->sc.keff;
      return 0;
   }
   if (s == "angular-flux-min") {
      double* d_min = nullptr;
      SN_CUDA(h, cudaMalloc(&d_min, sizeof(double)));
      cudaMemsetAsync(d_min, 0, sizeof(double), h->stream);
      launch_min(h->d_psi, h->psi_count, d_min, h->stream);
      cudaError_t e = cudaMemcpyAsync(out, d_min, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      cudaFree(d_min);
      if (e != cudaSuccess) SN_FAIL(h, std::string("CUDA error in the angular-flux minimum: ") + cudaGetErrorString(e));
      out[0] *= h->scale * h->psi_scale_factor;
      return 0;
   }
   if (s == "temperature" || s == "delayed-source") {
      std::vector<double>& v = s == "temperature" ? h->h_temperature : h->h_delayed;
      if (s == "delayed-source" && h->solved) {
         // S_i = beta * P_i (reference src/NeutronicSolver.cxx:103)
         std::vector<double> P(N);
         const int32_t part = h->opts.partition_fields;
         h->opts.partition_fields = 0;                   // whole field
         const int rcp = pampa_sn_get(h, "production-rate", P.data());
         h->opts.partition_fields = part;
         if (rcp) return 1;
         std::vector<int32_t> mats((size_t)pl.nz * pl.Sb);
         SN_CUDA(h, cudaMemcpy(mats.data(), h->d_mats, mats.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
         for (int64_t i = 0; i < N; i++) {
            int k = (int)(i / pl.nxy), c = (int)(i % pl.nxy);
            out[i] = h->h_beta[mats[(size_t)k * pl.Sb + pl.slot_of_xy[c]]] * P[i];
         }
         return 0;
      }
      for (int64_t i = 0; i < N; i++) out[i] = (int64_t)v.size() == N ? v[i] : 0.0;
      return 0;
   }
   // staging buffer: kept between calls up to 256 M doubles (2 GB), temporary above that
   double* d_out = nullptr;
   bool temp = false;
   if (count <= h->stage_count) d_out = h->d_stage;
   else if (count <= (int64_t)1 << 28) {
      if (h->d_stage) cudaFree(h->d_stage);
      h->d_stage = nullptr; h->stage_count = 0;
      SN_CUDA(h, cudaMalloc(&h->d_stage, (size_t)count * sizeof(double)));
      h->stage_count = count; d_out = h->d_stage;
   } else {
      SN_CUDA(h, cudaMalloc(&d_out, (size_t)count * sizeof(double)));
      temp = true;
   }
   int rc = 0;
   int64_t i0, ni;
   cell_window(h, &i0, &ni, nullptr);
   if (s == "scalar-flux") {
      launch_export_phi(h->d_phi, h->d_slot_of_xy, h->scale, h->G, pl.nz, pl.nxy, pl.Sb, i0, ni, d_out, h->stream);
   } else if (s == "flux-moments") {      // raw device moments sum_m w_m psi, no normalisation
      launch_export_phi(h->d_phi, h->d_slot_of_xy, 1.0, h->G, pl.nz, pl.nxy, pl.Sb, i0, ni, d_out, h->stream);
   } else if (s == "power") {
      launch_export_cell(h->d_phi, h->d_slot_of_xy, h->d_mats, h->d_kapsf, h->d_area, h->d_dz, pl.has_z,
                         h->scale, h->G, pl.nz, pl.nxy, pl.Sb, i0, ni, d_out, h->stream);
   } else if (s == "production-rate") {
      launch_export_cell(h->d_phi, h->d_slot_of_xy, h->d_mats, h->d_nusf, h->d_area, h->d_dz, pl.has_z,
                         h->scale / h->keff, h->G, pl.nz, pl.nxy, pl.Sb, i0, ni, d_out, h->stream);
   } else {   // angular-flux
      cudaMemsetAsync(d_out, 0, (size_t)count * sizeof(double), h->stream);
      double* d_min = nullptr;
      SN_CUDA(h, cudaMalloc(&d_min, sizeof(double)));
      cudaMemsetAsync(d_min, 0, sizeof(double), h->stream);
      for (int m = 0; m < h->M; m++) {
         const int c = h->dir_chunk[m];
         if (c < 0) continue;
         const Chunk& ch = pl.chunks[c];
         const ClassPlan& cp = pl.classes[ch.cls];
         launch_export_psi(h->chunk_psi[c], h->d_classes + ch.cls, h->d_pos_of[ch.cls], h->d_slot_of_xy, h->dir_d[m],
                           ch.nd, m, h->d_gloc, h->scale * h->psi_scale_factor, h->G, h->M, pl.nz, pl.nxy, d_out, d_min,
                           h->stream);
      }
      double mn = 0.0;
      cudaMemcpyAsync(&mn, d_min, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
      cudaStreamSynchronize(h->stream);
      cudaFree(d_min);
      if (mn < 0.0) { h->err = "negative values in the angular-flux solution"; rc = 1; }
   }
   cudaError_t e = cudaMemcpyAsync(out, d_out, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
   if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
   if (temp) cudaFree(d_out);
   if (e != cudaSuccess) SN_FAIL(h, std::string("CUDA error exporting field: ") + cudaGetErrorString(e));
   return rc;
}

int pampa_sn_set(pampa_sn_handle* h, const char* name, const double* in) {
   const std::string s(name);
   const int64_t N = (int64_t)h->plan.nxy * h->plan.nz;
   if (s == "temperature") { h->h_temperature.assign(in, in + N); return 0; }
   if (s == "delayed-source") { h->h_delayed.assign(in, in + N); return 0; }
   if (s == "keff") {              // eigenvalue estimate the next iteration starts from (1 value)
      SN_CUDA(h, cudaSetDevice(h->device));
      if (!(in[0] > 0.0)) SN_FAIL(h, "the eigenvalue estimate must be positive");
      SN_CUDA(h, cudaMemcpyAsync(&h->d_sc->keff, in, sizeof(double), cudaMemcpyHostToDevice, h->stream));
      SN_CUDA(h, cudaStreamSynchronize(h->stream));
      h->sc.keff = in[0]; h->keff = in[0];
      return 0;
   }
   if (s == "flux-moments") {      // iteration state / initial guess, layout [i][g]
      SN_CUDA(h, cudaSetDevice(h->device));
      int64_t i0, ni, per_rank;
      cell_window(h, &i0, &ni, &per_rank);
      const bool part = h->opts.partition_fields && h->opts.num_ranks > 1;
      const int64_t count = part ? per_rank * h->opts.num_ranks * h->G : N * h->G;
      if (count > h->stage_count) {
         if (h->d_stage) cudaFree(h->d_stage);
         h->d_stage = nullptr; h->stage_count = 0;
         SN_CUDA(h, cudaMalloc(&h->d_stage, (size_t)count * sizeof(double)));
         h->stage_count = count;
      }
      double* d_in = h->d_stage;
      if (part) {
         // this rank uploads its window of cells; the windows are exchanged on the device (in-place allgather)
         if (!h->comm) SN_FAIL(h, "partitioned fields need the communicator (pampa_sn_comm_init)");
         cudaMemcpyAsync(d_in + i0 * h->G, in, (size_t)(ni * h->G) * sizeof(double), cudaMemcpyHostToDevice, h->stream);
         int r = g_nccl.AllGather(d_in + (int64_t)h->opts.rank * per_rank * h->G, d_in, (size_t)(per_rank * h->G), NCCL_FLOAT64,
                                  h->comm, h->stream);
         if (r != 0) return nccl_fail(h, r, "the field allgather");
      } else
         cudaMemcpyAsync(d_in, in, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, h->stream);
      cudaMemsetAsync(h->d_phi, 0, (size_t)h->G * h->plan.nz * h->plan.Sb * sizeof(double), h->stream);
      launch_import_phi(h->d_phi_new, h->d_slot_of_xy, h->G, h->plan.nz, h->plan.nxy, h->plan.Sb, d_in, h->stream);
      if (do_reduce(h, 0, false)) return 1;
      if (sync_scalars(h)) return 1;
      return check_async(h, "field import");
   }
   SN_FAIL(h, "unable to find field '" + s + "'");
}

int pampa_sn_comm_unique_id(void* id, int32_t id_bytes) {
   if (id_bytes != (int32_t)sizeof(NcclUniqueId)) { g_create_error = "NCCL unique id must be 128 bytes"; return 1; }
   if (!g_nccl.load(g_create_error)) return 1;
   return g_nccl.GetUniqueId((NcclUniqueId*)id) == 0 ? 0 : 1;
}

int pampa_sn_comm_init(pampa_sn_handle* h, const void* id, int32_t id_bytes) {
   if (id_bytes != (int32_t)sizeof(NcclUniqueId)) SN_FAIL(h, "NCCL unique id must be 128 bytes");
   if (!g_nccl.load(h->err)) return 1;
   SN_CUDA(h, cudaSetDevice(h->device));
   NcclUniqueId uid;
   std::memcpy(&uid, id, sizeof(uid));
   int r = g_nccl.CommInitRank(&h->comm, h->opts.num_ranks, uid, h->opts.rank);
   if (r != 0) SN_FAIL(h, std::string("ncclCommInitRank failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
   h->group_gather = h->opts.shard_mode == 1 && h->G % h->opts.num_ranks == 0;
   // Peer-to-peer delivery of the flux moments: map the two iterate buffers of every other rank (CUDA IPC; the
   // handles travel through an allgather on the new communicator).  Any failure -- ranks on different nodes, IPC
   // not permitted, PAMPA_SN_NO_P2P=1 -- leaves the NCCL allgather in place, on every rank alike.
   h->p2p = false;
   if (h->group_gather && h->opts.num_ranks - 1 <= PEER_MAX && !std::getenv("PAMPA_SN_NO_P2P")) {
      const int R = h->opts.num_ranks;
      const int64_t nphi = (int64_t)h->G * h->plan.nz * h->plan.Sb;
      if (!h->d_phi_buf[1]) {
         if (dev_alloc(h, &h->d_phi_buf[1], nphi)) return 1;
         SN_CUDA(h, cudaMemsetAsync(h->d_phi_buf[1], 0, (size_t)nphi * sizeof(double), h->stream));
      }
      struct Pack { cudaIpcMemHandle_t hnd[2]; int32_t ok; int32_t pad[15]; };
      static_assert(sizeof(Pack) % 8 == 0, "pack size");
      Pack mine;
      std::memset(&mine, 0, sizeof(mine));
      mine.ok = cudaIpcGetMemHandle(&mine.hnd[0], h->d_phi_buf[0]) == cudaSuccess &&
                cudaIpcGetMemHandle(&mine.hnd[1], h->d_phi_buf[1]) == cudaSuccess;
      cudaGetLastError();
      Pack* d_all = nullptr;
      std::vector<Pack> all(R);
      SN_CUDA(h, cudaMalloc((void**)&d_all, sizeof(Pack) * R));
      SN_CUDA(h, cudaMemcpyAsync(d_all + h->opts.rank, &mine, sizeof(Pack), cudaMemcpyHostToDevice, h->stream));
      int rr = g_nccl.AllGather(d_all + h->opts.rank, d_all, sizeof(Pack) / 8, NCCL_FLOAT64, h->comm, h->stream);
      if (rr != 0) { cudaFree(d_all); return nccl_fail(h, rr, "the allgather of the IPC handles"); }
      SN_CUDA(h, cudaMemcpyAsync(all.data(), d_all, sizeof(Pack) * R, cudaMemcpyDeviceToHost, h->stream));
      SN_CUDA(h, cudaStreamSynchronize(h->stream));
      cudaFree(d_all);
      bool ok = true;
      for (int p = 0; p < R; p++) ok = ok && all[p].ok;
      int np = 0;
      for (int p = 0; p < R && ok; p++) {
         if (p == h->opts.rank) continue;
         for (int b = 0; b < 2 && ok; b++) {
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[p].hnd[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); }
            else { h->peer_phi[b][np] = (double*)ptr; h->ipc_mapped.push_back(ptr); }
         }
         np++;
      }
      // every rank must take the same path: agree on success with a one-value collective
      double flag = ok ? 1.0 : 0.0, *d_flag = nullptr;
      SN_CUDA(h, cudaMalloc((void**)&d_flag, sizeof(double)));
      SN_CUDA(h, cudaMemcpyAsync(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice, h->stream));
      rr = g_nccl.AllReduce(d_flag, d_flag, 1, NCCL_FLOAT64, NCCL_MIN, h->comm, h->stream);
      if (rr == 0) { cudaMemcpyAsync(&flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost, h->stream); cudaStreamSynchronize(h->stream); }
      cudaFree(d_flag);
      if (rr != 0) return nccl_fail(h, rr, "the peer-access agreement");
      h->p2p = flag > 0.5;
      h->npeers = h->p2p ? R - 1 : 0;
      if (h->opts.verbose) std::printf("pampa_sn: rank %d: flux-moment exchange by %s\n", h->opts.rank,
                                       h->p2p ? "peer-to-peer stores fused into the reduction pass" : "NCCL allgather");
   }
   return 0;
}

void* pampa_sn_device_ptr(pampa_sn_handle* h, const char* name, int64_t* count) {
   const std::string s(name);
   const int64_t nphi = (int64_t)h->G * h->plan.nz * h->plan.Sb;
   if (s == "phi") { if (count) *count = nphi; return h->d_phi; }
   if (s == "phi-new") { if (count) *count = nphi; return h->d_phi_new; }
   if (s == "q") { if (count) *count = nphi; return h->d_q; }
   if (count) *count = 0;
   return nullptr;
}

int pampa_sn_get_info(pampa_sn_handle* h, pampa_sn_info* info) {
   std::memset(info, 0, sizeof(*info));
   const Plan& pl = h->plan;
   info->num_cells = (int64_t)pl.nxy * pl.nz; info->num_groups = h->G; info->num_directions = h->M;
   info->updates_per_sweep = pl.owned_updates;
   info->sweep_launches = (int64_t)(h->groups.size() + h->flows.size());
   for (auto& g : h->groups) info->sweep_tasks += g.count;
   for (auto& f : h->flows) info->sweep_tasks += f.count;
   info->num_classes = (int64_t)pl.classes.size(); info->num_chunks = (int64_t)pl.chunks.size();
   info->tile_classes = pl.tile_classes;
   info->num_tilings = (int64_t)pl.tilings.size(); info->lattice = pl.lattice;
   for (size_t ci = 0; ci < pl.classes.size(); ci++) info->flow_classes += (!h->flows.empty() && h->class_fast[ci]) ? 1 : 0;
   info->device_bytes = h->device_bytes;
   info->timed_kernel_ms = h->timed_kernel_ms;
   info->last_sweep_ms = h->last_sweep_ms; info->last_source_ms = h->last_source_ms;
   info->last_reduce_ms = h->last_reduce_ms; info->last_solve_ms = h->last_solve_ms;
   info->timed_source_ms = h->timed_source_ms; info->timed_reduce_ms = h->timed_reduce_ms;
   info->timed_exchange_ms = h->timed_exchange_ms;
   info->kernel_launches = h->launches;
   return 0;
}

}  // extern "C"
