// sn_plan.hpp -- host-side sweep planning for the B200 SN transport layer.
//
// Turns the extruded mesh + quadrature into what the sweep kernel consumes:
//   * a padded "slot" numbering of the xy cells (patch-major: slot = patch*P + lane), shared by
//     q, phi and the material map, so that a CTA's accesses are contiguous;
//   * ordering classes: sets of directions with the same upwind face pattern (one per octant on
//     Cartesian meshes; azimuthal sectors on hexagonal / triangular meshes);
//   * per class, a partition of the xy cells into patches of <= P cells whose patch graph is
//     acyclic: the shared 2-D tiles when they are acyclic for the class, otherwise chunks of the
//     class's own topological (level, lateral) order;
//   * per slot, the local pipeline level and the upwind sources (in-patch lane, other-patch
//     slot, reflective boundary face), with the face vectors that give the streaming
//     coefficients of Appendix A of SURVEY.md (reference: src/SNSolver.cxx:574-597);
//   * the launch schedule: tasks (chunk of directions, group, patch, z chunk) grouped by
//     wavefront number = patch level + z chunk.
//
// Nothing here touches the GPU; tests exercise it through pampa_sn_plan_check().
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pampa_sn.h"

namespace pampa_sn {

constexpr int FIN_MAX = 4;         // incoming lateral faces per cell the kernel is instantiated for
constexpr int ROUT_MAX = 2;        // outgoing reflective lateral faces per cell
constexpr int DT_MAX = 10;         // directions per sweep chunk (register-resident chains)
constexpr int DT_DEFAULT = 5;      // measured best on B200 (registers -> 2 CTAs/SM without spills)
constexpr int RING_MAX = 4;        // smem ring depth for in-patch upwind values
constexpr int FLOW_FIN = 3;        // incoming lateral faces the dataflow kernel handles (hexagonal lattices: 3)
constexpr int FLOW_HALO = 64;      // patch-boundary / reflective sources per patch the dataflow kernel stages
constexpr int FLOW3_DT_CAP = 6;    // widest chunk the three-face variant is instantiated for
constexpr int FLOW3_DT_DEFAULT = 5;  // S8 on a hexagonal lattice: classes of 6 / 7 directions -> 3+3 / 4+3 either way; S12 (14 per
                                    // class): 5+5+4 instead of 4+4+4+2, measured 2.5 % faster (19 441 hexagons x 100 layers, 16 groups)
constexpr int FLOW_EXPORT = 32;    // lanes of a patch other patches read (compact edge copies per psi row)
constexpr uint16_t LVL_EMPTY = 0xFFFF;
constexpr int PERIM_MAX = 64;           // perimeter lanes of a structured tile (16x16: 60)

// upwind source codes
constexpr int32_t SRC_NONE = -1;
constexpr int32_t SRC_KIND_SHIFT = 28;
constexpr int32_t SRC_GLOBAL = 0, SRC_LOCAL = 1, SRC_REFL = 2;
constexpr int32_t SRC_PAYLOAD = 0x0FFFFFFF;
constexpr int32_t SRC_AXIS_SHIFT = 26;

struct Vec2 { double x, y; };

// A partition of the xy cells into patches of <= P cells shared by several ordering classes: the structured tiles
// of a Cartesian mesh, the rhombic tiles of a lattice of congruent cells in one of its bases, or k-d leaves.
struct Tiling {
   int npatch = 0;
   std::vector<int32_t> slot_of_xy;     // [nxy] xy cell -> slot of this tiling (patch * P + lane)
};

struct ClassPlan {
   int zdir = 0;                        // +1: sweep k upward, -1: downward, 0: no z faces
   std::vector<int> dirs;               // quadrature directions of this class
   bool tiles = false;                  // true: swept on the base tiling (class slot == base slot)
   int tiling = -1;                     // shared tiling this class is swept on (-1: its own level-chunk patches)
   int npatch = 0;
   int64_t S = 0;                       // slots = npatch * P
   int fin = 0;                         // max incoming lateral faces
   int ring = 2;                        // smem ring depth
   int nsteps = 1;                      // pipeline steps stored per patch: max local levels + nz - 1
   std::vector<int32_t> cell_of;        // [S] base slot (or -1)
   std::vector<int32_t> pos_of;         // [Sb] base slot -> class slot (or -1)
   std::vector<uint16_t> lvl;           // [S] local level, LVL_EMPTY for holes
   std::vector<int32_t> patch_nlev;     // [npatch]
   std::vector<int32_t> patch_level;    // [npatch] level in the patch graph
   std::vector<Vec2> out_vec;           // [S] sum over outgoing faces of cf*f/area
   std::vector<int32_t> in_src;         // [FIN_MAX][S]
   std::vector<Vec2> in_vec;            // [FIN_MAX][S] cf*f/area of the incoming face
   std::vector<int32_t> rout;           // [ROUT_MAX][S] reflective face id written by this slot
   std::vector<uint16_t> in_hidx;       // [FIN_MAX][S] halo index of a patch-boundary / reflective source
   int max_halo = 0;                    // largest number of such sources in one patch
   std::vector<uint8_t> eidx;           // [S] compact index of a lane other patches read from (255: none)
   int max_export = 0;
   bool fast = false;                   // eligible for the staged tile kernel (and the dataflow kernel)
   bool fast_flow = false;              // eligible for the dataflow kernel: shared tiling, <= 3 incoming faces,
                                        // in-patch level differences <= 2, <= 64 staged sources per patch
   bool inline_ok = false;              // every lane another patch reads is one of the first PERIM_MAX lanes
};

struct Chunk {
   int cls = 0;
   int nd = 0;
   int m[DT_MAX];                       // quadrature direction indices
   int64_t psi_offset = 0;              // offset (doubles) of this chunk's psi block
};

struct Task { int32_t chunk, group, patch, zc; };

struct Plan {
   int P = 256;
   int nxy = 0, nz = 1, has_z = 0;
   int G = 0, M = 0;
   int npatch_b = 0;
   int64_t Sb = 0;                      // base slots
   std::vector<int32_t> slot_of_xy;     // [nxy] xy cell -> base slot
   std::vector<int32_t> xy_of_slot;     // [Sb] base slot -> xy cell or -1
   std::vector<ClassPlan> classes;
   std::vector<int> class_of_dir;       // [M]
   std::vector<Chunk> chunks;
   int Kc = 1, nzc = 1;                 // layers per task, z chunks
   std::vector<std::vector<Task>> waves;  // launch schedule (owned tasks only)
   int64_t psi_doubles = 0;
   int num_rfaces = 0;                  // lateral reflective faces (xy cell, face) pairs
   std::vector<int32_t> rface_axis;     // [num_rfaces]
   int64_t owned_updates = 0;
   int tile_classes = 0;
   int nperim = 0;                      // structured tiles: lanes 0 .. nperim-1 are the tile perimeter (0: row-major)
   std::vector<Tiling> tilings;         // [0] = the base partition (slot numbering of q, phi, materials)
   int lattice = 0;                     // 1: unstructured mesh recognised as a lattice of congruent cells
};

namespace detail {

inline void kd_split(std::vector<int>& ids, int lo, int hi, const double* cx, const double* cy,
                     int P, std::vector<std::pair<int, int>>& leaves) {
   if (hi - lo <= P) { leaves.push_back({lo, hi}); return; }
   double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
   for (int a = lo; a < hi; a++) {
      x0 = std::min(x0, cx[ids[a]]); x1 = std::max(x1, cx[ids[a]]);
      y0 = std::min(y0, cy[ids[a]]); y1 = std::max(y1, cy[ids[a]]);
   }
   const bool by_x = (x1 - x0) >= (y1 - y0);
   // split into the smallest number of <= P leaves: left part gets a multiple of P when possible
   int n = hi - lo;
   int nleaf = (n + P - 1) / P;
   int mid = lo + (nleaf / 2) * ((n + nleaf - 1) / nleaf);
   if (mid <= lo || mid >= hi) mid = lo + n / 2;
   auto key = [&](int c) { return by_x ? cx[c] : cy[c]; };
   std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi,
                    [&](int a, int b) { return key(a) < key(b) || (key(a) == key(b) && a < b); });
   kd_split(ids, lo, mid, cx, cy, P, leaves);
   kd_split(ids, mid, hi, cx, cy, P, leaves);
}

// longest-path levels of a DAG given as upwind adjacency; returns false on a cycle
inline bool dag_levels(int n, const std::vector<std::vector<int>>& up, std::vector<int>& level) {
   std::vector<int> indeg(n, 0);
   std::vector<std::vector<int>> down(n);
   for (int v = 0; v < n; v++)
      for (int u : up[v]) { down[u].push_back(v); indeg[v]++; }
   level.assign(n, 0);
   std::vector<int> stack;
   for (int v = 0; v < n; v++) if (indeg[v] == 0) stack.push_back(v);
   int done = 0;
   while (!stack.empty()) {
      int u = stack.back(); stack.pop_back(); done++;
      for (int v : down[u]) {
         level[v] = std::max(level[v], level[u] + 1);
         if (--indeg[v] == 0) stack.push_back(v);
      }
   }
   return done == n;
}

// Unstructured meshes whose cells are translates of one another (hexagonal assemblies, quadrilateral grids written
// as polygon lists) are lattices: every centroid is c0 + q a + r b with integer (q, r).  Returns the integer
// coordinates and the distinct neighbour offsets (up to sign) in those coordinates, or false.
inline bool detect_lattice(const pampa_sn_mesh& ms, std::vector<int>& qr, std::vector<std::pair<int, int>>& offs) {
   const int nxy = ms.num_xy_cells, F = ms.max_xy_faces;
   if (nxy < 4) return false;
   int c0 = -1, best = 0;
   for (int c = 0; c < nxy && best < F; c++) {          // a cell with as many interior neighbours as possible
      int n = 0;
      for (int f = 0; f < ms.xy_num_faces[c]; f++) n += ms.xy_neighbor[(size_t)c * F + f] >= 0;
      if (n > best) { best = n; c0 = c; }
   }
   if (c0 < 0 || best < 2) return false;
   std::vector<Vec2> o;
   for (int f = 0; f < ms.xy_num_faces[c0]; f++) {
      const int nb = ms.xy_neighbor[(size_t)c0 * F + f];
      if (nb >= 0) o.push_back(Vec2{ms.xy_cx[nb] - ms.xy_cx[c0], ms.xy_cy[nb] - ms.xy_cy[c0]});
   }
   const Vec2 a = o[0];
   const double la = std::sqrt(a.x * a.x + a.y * a.y);
   Vec2 b{0, 0};
   double bdet = 0.0;
   for (const Vec2& v : o) {                            // the neighbour offset that spans the smallest cell with a
      const double det = a.x * v.y - a.y * v.x;
      const double lv = std::sqrt(v.x * v.x + v.y * v.y);
      if (std::fabs(det) > 0.1 * la * lv &&              // not (anti)parallel to a, rounded coordinates included
          (bdet == 0.0 || std::fabs(det) < std::fabs(bdet) * (1.0 - 1.0e-3))) { b = v; bdet = det; }
   }
   if (bdet == 0.0) return false;
   // Integer coordinates by walking the neighbour graph: every step is the offset to a face neighbour, a small
   // lattice vector that rounds safely even when the file carries three decimals (the reference's own mesh writer
   // does); rounding absolute positions instead would accumulate the error of a and b over the width of the mesh.
   qr.assign((size_t)2 * nxy, 0);
   std::vector<char> visited(nxy, 0);
   std::vector<int> queue;
   queue.reserve(nxy);
   queue.push_back(c0); visited[c0] = 1;
   for (size_t head = 0; head < queue.size(); head++) {
      const int c = queue[head];
      for (int f = 0; f < ms.xy_num_faces[c]; f++) {
         const int nb = ms.xy_neighbor[(size_t)c * F + f];
         if (nb < 0) continue;
         const double dx = ms.xy_cx[nb] - ms.xy_cx[c], dy = ms.xy_cy[nb] - ms.xy_cy[c];
         const double q = (dx * b.y - dy * b.x) / bdet, r = (a.x * dy - a.y * dx) / bdet;
         const double qi = std::nearbyint(q), ri = std::nearbyint(r);
         if (std::fabs(q - qi) > 0.05 || std::fabs(r - ri) > 0.05 || std::fabs(qi) > 1.0 || std::fabs(ri) > 1.0) return false;
         const int nq = qr[2 * c] + (int)qi, nr = qr[2 * c + 1] + (int)ri;
         if (!visited[nb]) { visited[nb] = 1; qr[2 * nb] = nq; qr[2 * nb + 1] = nr; queue.push_back(nb); }
         else if (qr[2 * nb] != nq || qr[2 * nb + 1] != nr) return false;
      }
   }
   if ((int)queue.size() != nxy) return false;           // disconnected mesh
   std::map<std::pair<int, int>, int> seen;
   for (int c = 0; c < nxy; c++)
      if (!seen.emplace(std::make_pair(qr[2 * c], qr[2 * c + 1]), c).second) return false;
   std::map<std::pair<int, int>, int> dirs;
   for (int c = 0; c < nxy; c++)
      for (int f = 0; f < ms.xy_num_faces[c]; f++) {
         const int nb = ms.xy_neighbor[(size_t)c * F + f];
         if (nb < 0) continue;
         int dq = qr[2 * nb] - qr[2 * c], dr = qr[2 * nb + 1] - qr[2 * c + 1];
         if (std::abs(dq) > 1 || std::abs(dr) > 1) return false;       // neighbours are lattice neighbours
         if (dq < 0 || (dq == 0 && dr < 0)) { dq = -dq; dr = -dr; }
         dirs[{dq, dr}]++;
      }
   offs.clear();
   for (auto& d : dirs) offs.push_back(d.first);
   return offs.size() >= 2 && offs.size() <= 4;
}

}  // namespace detail

struct PlanInput {
   const pampa_sn_mesh* mesh;
   const pampa_sn_quadrature* quad;
   int G;
   pampa_sn_options opts;
};

inline void build_plan(const PlanInput& in, Plan& pl) {
   const pampa_sn_mesh& ms = *in.mesh;
   const pampa_sn_quadrature& qd = *in.quad;
   const int nxy = ms.num_xy_cells, F = ms.max_xy_faces;
   if (nxy <= 0 || ms.num_layers <= 0 || F <= 0) throw std::runtime_error("empty mesh");
   pl.nxy = nxy; pl.nz = ms.num_layers; pl.has_z = ms.has_z_faces ? 1 : 0;
   pl.G = in.G; pl.M = qd.num_directions;
   const int P = 256;                 // slot stride of a patch = sweep CTA size (fixed)
   int cap = in.opts.patch_cells > 0 ? in.opts.patch_cells : P;   // cells per patch (<= P)
   int ti = in.opts.tile_i > 0 ? in.opts.tile_i : 16;
   int tj = in.opts.tile_j > 0 ? in.opts.tile_j : 16;
   if (cap > P || cap < 1) throw std::runtime_error("patch_cells must be in 1..256");
   if (ms.xy_ij) {
      // thin meshes (1-D slabs, narrow strips): stretch the tile along i instead of wasting lanes
      int jext = 0;
      for (int c = 0; c < nxy; c++) jext = std::max(jext, ms.xy_ij[2*c+1] + 1);
      while (tj > 1 && tj / 2 >= jext) { tj /= 2; ti *= 2; }
      if (ti * tj > P) throw std::runtime_error("tile_i * tile_j must be <= 256");
      cap = ti * tj;
   }
   pl.P = P;

   auto bc_type = [&](int nb) -> int {
      int b = -nb;
      if (b < 1 || b > ms.num_bcs) return PAMPA_SN_BC_NONE;
      return ms.bc_types[b];
   };

   // ---- shared partitions (class independent) ------------------------------------------------
   // tilings[0] is the base partition (numbering of q, phi and the material map).  Structured meshes have one
   // tiling; a lattice of congruent cells (hexagonal assemblies) gets one rhombic tiling per pair of lattice
   // directions, because the straight tile edges of a basis are zig-zag lines for the directions that run along
   // them (the two neighbouring tiles would depend on each other): every ordering class picks a tiling whose
   // patch graph is acyclic for it.  Anything else: k-d leaves.
   pl.tilings.clear();
   pl.nperim = 0;
   auto tile_from_coords = [&](const int* ij, int ti_, int tj_, bool perimeter_first) {
      Tiling tg;
      tg.slot_of_xy.assign(nxy, -1);
      int imax = 0, jmax = 0;
      for (int c = 0; c < nxy; c++) { imax = std::max(imax, ij[2*c]); jmax = std::max(jmax, ij[2*c+1]); }
      int ntx = imax / ti_ + 1, nty = jmax / tj_ + 1;
      std::vector<int> tile_id((size_t)ntx * nty, -1);
      int np = 0;
      for (int c = 0; c < nxy; c++) {           // number the non-empty tiles in first-touch order
         int t = (ij[2*c+1] / tj_) * ntx + ij[2*c] / ti_;
         if (tile_id[t] < 0) tile_id[t] = 0;
      }
      for (int t = 0; t < ntx * nty; t++) if (tile_id[t] == 0) tile_id[t] = np++;
      // Lane order inside a tile: row-major, or (inline_edges) the perimeter first, walked as a ring (bottom row,
      // right column, top row backwards, left column downwards), then the interior row by row.  Only perimeter
      // lanes are read by other patches, and the two sides a sweep direction exports are adjacent on the ring, so
      // a neighbouring patch can read them straight from the psi rows as one or two contiguous segments of the
      // first PERIM_MAX lanes (no edge copies).  Thin tiles (a side < 3) keep the row-major order.
      std::vector<int> lane_of_pos(ti_ * tj_);
      std::iota(lane_of_pos.begin(), lane_of_pos.end(), 0);
      if (perimeter_first && ti_ >= 3 && tj_ >= 3 && 2 * (ti_ + tj_) - 4 <= PERIM_MAX) {
         int n = 0;
         for (int i = 0; i < ti_; i++) lane_of_pos[i] = n++;                                   // j = 0
         for (int j = 1; j < tj_; j++) lane_of_pos[j * ti_ + ti_ - 1] = n++;                   // i = ti - 1
         for (int i = ti_ - 2; i >= 0; i--) lane_of_pos[(tj_ - 1) * ti_ + i] = n++;            // j = tj - 1
         for (int j = tj_ - 2; j >= 1; j--) lane_of_pos[j * ti_] = n++;                        // i = 0
         pl.nperim = n;
         for (int j = 1; j < tj_ - 1; j++)
            for (int i = 1; i < ti_ - 1; i++) lane_of_pos[j * ti_ + i] = n++;
      }
      for (int c = 0; c < nxy; c++) {
         int i = ij[2*c], j = ij[2*c+1];
         int t = (j / tj_) * ntx + i / ti_;
         tg.slot_of_xy[c] = tile_id[t] * P + lane_of_pos[(j % tj_) * ti_ + (i % ti_)];   // ti*tj <= P lanes used
      }
      tg.npatch = np;
      return tg;
   };
   pl.lattice = 0;
   if (ms.xy_ij) {
      pl.tilings.push_back(tile_from_coords(ms.xy_ij, ti, tj, in.opts.inline_edges != 0));
   } else {
      std::vector<int> qr;
      std::vector<std::pair<int, int>> offs;
      if (cap == P && detail::detect_lattice(ms, qr, offs)) {
         pl.lattice = 1;
         for (size_t a = 0; a < offs.size(); a++)
            for (size_t b = a + 1; b < offs.size(); b++) {
               const int det = offs[a].first * offs[b].second - offs[a].second * offs[b].first;
               if (det != 1 && det != -1) continue;
               std::vector<int> ab((size_t)2 * nxy);
               int amin = INT32_MAX, bmin = INT32_MAX;
               for (int c = 0; c < nxy; c++) {
                  const int q = qr[2*c], r = qr[2*c+1];
                  ab[2*c] = (q * offs[b].second - r * offs[b].first) / det;
                  ab[2*c+1] = (offs[a].first * r - offs[a].second * q) / det;
                  amin = std::min(amin, ab[2*c]); bmin = std::min(bmin, ab[2*c+1]);
               }
               for (int c = 0; c < nxy; c++) { ab[2*c] -= amin; ab[2*c+1] -= bmin; }
               pl.tilings.push_back(tile_from_coords(ab.data(), ti, tj, false));
            }
         // Lanes of the other tilings in the order of the base slots: the shear passes move q and phi between the
         // base numbering and a tiling's by gather / scatter, and a rhombic tile cuts every row of a base tile in
         // one contiguous run, so sorted lanes turn most of those accesses into whole sectors (the sweep itself
         // does not care how the lanes of a patch are numbered).
         for (size_t tg = 1; tg < pl.tilings.size(); tg++) {
            Tiling& t = pl.tilings[tg];
            std::vector<std::vector<std::pair<int32_t, int>>> members(t.npatch);
            for (int c = 0; c < nxy; c++) members[t.slot_of_xy[c] / P].push_back({pl.tilings[0].slot_of_xy[c], c});
            for (int p = 0; p < t.npatch; p++) {
               std::sort(members[p].begin(), members[p].end());
               for (size_t l = 0; l < members[p].size(); l++) t.slot_of_xy[members[p][l].second] = p * P + (int)l;
            }
         }
      }
      if (pl.tilings.empty()) {
         pl.lattice = 0;
         std::vector<int> ids(nxy);
         std::iota(ids.begin(), ids.end(), 0);
         std::vector<std::pair<int, int>> leaves;
         detail::kd_split(ids, 0, nxy, ms.xy_cx, ms.xy_cy, cap, leaves);
         Tiling tg;
         tg.slot_of_xy.assign(nxy, -1);
         int np = 0;
         for (auto& lf : leaves) {
            std::sort(ids.begin() + lf.first, ids.begin() + lf.second);
            for (int a = lf.first; a < lf.second; a++) tg.slot_of_xy[ids[a]] = np * P + (a - lf.first);
            np++;
         }
         tg.npatch = np;
         pl.tilings.push_back(tg);
      }
   }
   pl.slot_of_xy = pl.tilings[0].slot_of_xy;
   pl.npatch_b = pl.tilings[0].npatch;
   pl.Sb = (int64_t)pl.npatch_b * P;
   pl.xy_of_slot.assign(pl.Sb, -1);
   for (int c = 0; c < nxy; c++) {
      if (pl.xy_of_slot[pl.slot_of_xy[c]] != -1) throw std::runtime_error("duplicate structured (i,j) in xy_ij");
      pl.xy_of_slot[pl.slot_of_xy[c]] = c;
   }

   // ---- reflective lateral faces -----------------------------------------------------------
   std::vector<int32_t> rface_id((size_t)nxy * F, -1);
   pl.num_rfaces = 0; pl.rface_axis.clear();
   for (int c = 0; c < nxy; c++)
      for (int f = 0; f < ms.xy_num_faces[c]; f++) {
         int nb = ms.xy_neighbor[(size_t)c * F + f];
         if (nb >= 0) continue;
         int t = bc_type(nb);
         if (t == PAMPA_SN_BC_REFLECTIVE) {
            double fx = ms.xy_face_fx[(size_t)c * F + f], fy = ms.xy_face_fy[(size_t)c * F + f];
            double len = std::sqrt(fx * fx + fy * fy);
            int axis = -1;                         // intended per-cell test of SNSolver.cxx:539-543
            if (std::fabs(fx / len) > 1.0 - 1.0e-6) axis = 0;
            if (std::fabs(fy / len) > 1.0 - 1.0e-6) axis = 1;
            if (axis < 0) throw std::runtime_error("reflected direction not found");
            rface_id[(size_t)c * F + f] = pl.num_rfaces++;
            pl.rface_axis.push_back(axis);
         } else if (t != PAMPA_SN_BC_VACUUM) {
            throw std::runtime_error("boundary condition not implemented");
         }
      }
   if (pl.has_z) {
      for (int b : {ms.bc_minus_z, ms.bc_plus_z}) {
         int t = (b >= 1 && b <= ms.num_bcs) ? ms.bc_types[b] : PAMPA_SN_BC_NONE;
         if (t != PAMPA_SN_BC_VACUUM && t != PAMPA_SN_BC_REFLECTIVE)
            throw std::runtime_error("boundary condition not implemented");
      }
   }

   // ---- ordering classes -------------------------------------------------------------------
   const double eps = 1.0e-14;
   std::map<std::vector<uint8_t>, int> class_key;
   pl.class_of_dir.assign(pl.M, -1);
   pl.classes.clear();
   std::vector<std::vector<uint8_t>> class_flags;   // incoming flag per (xy cell, face)
   for (int m = 0; m < pl.M; m++) {
      const double ox = qd.directions[3*m], oy = qd.directions[3*m+1], oz = qd.directions[3*m+2];
      std::vector<uint8_t> key((size_t)nxy * F + 1, 0);
      for (int c = 0; c < nxy; c++)
         for (int f = 0; f < ms.xy_num_faces[c]; f++) {
            size_t a = (size_t)c * F + f;
            double w = ox * ms.xy_face_fx[a] + oy * ms.xy_face_fy[a];
            key[a] = (w > eps) ? 1 : (w < -eps ? 2 : 0);
         }
      key[(size_t)nxy * F] = pl.has_z ? (oz > 0 ? 1 : 2) : 0;
      auto it = class_key.find(key);
      int ci;
      if (it == class_key.end()) {
         ci = (int)pl.classes.size();
         class_key.emplace(key, ci);
         ClassPlan cp; cp.zdir = pl.has_z ? (oz > 0 ? 1 : -1) : 0;
         pl.classes.push_back(cp);
         class_flags.push_back(key);
      } else ci = it->second;
      pl.classes[ci].dirs.push_back(m);
      pl.class_of_dir[m] = ci;
   }
   class_key.clear();

   // ---- per-class patches, levels, sources ---------------------------------------------------
   pl.tile_classes = 0;
   for (size_t ci = 0; ci < pl.classes.size(); ci++) {
      ClassPlan& cp = pl.classes[ci];
      const std::vector<uint8_t>& flg = class_flags[ci];
      // xy upwind graph
      std::vector<std::vector<int>> up(nxy);
      int fin = 0;
      for (int c = 0; c < nxy; c++) {
         int cnt = 0;
         for (int f = 0; f < ms.xy_num_faces[c]; f++) {
            size_t a = (size_t)c * F + f;
            if (flg[a] != 2) continue;
            int nb = ms.xy_neighbor[a];
            if (nb >= 0) { up[c].push_back(nb); cnt++; }
            else if (bc_type(nb) == PAMPA_SN_BC_REFLECTIVE) cnt++;
         }
         fin = std::max(fin, cnt);
      }
      if (fin > FIN_MAX) throw std::runtime_error("cell with more than 4 incoming lateral faces");
      cp.fin = fin;
      std::vector<int> glevel;
      if (!detail::dag_levels(nxy, up, glevel)) throw std::runtime_error("cyclic upwind dependency in the xy mesh");

      std::vector<int32_t> patch_of(nxy), lane_of(nxy);
      auto try_partition = [&](ClassPlan& cq, int np) -> bool {
         // patch graph
         std::vector<std::vector<int>> pup(np);
         for (int c = 0; c < nxy; c++)
            for (int u : up[c]) if (patch_of[u] != patch_of[c]) pup[patch_of[c]].push_back(patch_of[u]);
         for (auto& v : pup) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
         std::vector<int> plevel;
         if (!detail::dag_levels(np, pup, plevel)) return false;
         cq.npatch = np; cq.S = (int64_t)np * P;
         cq.patch_level.assign(plevel.begin(), plevel.end());
         return true;
      };
      // slot tables, local levels, upwind sources of the partition in patch_of / lane_of
      auto build_tables = [&](ClassPlan& cq) {
         const int64_t S = cq.S;
         cq.cell_of.assign(S, -1); cq.pos_of.assign(pl.Sb, -1);
         for (int c = 0; c < nxy; c++) {
            int64_t sl = (int64_t)patch_of[c] * P + lane_of[c];
            cq.cell_of[sl] = pl.slot_of_xy[c];
            cq.pos_of[pl.slot_of_xy[c]] = (int32_t)sl;
         }
         // local levels: longest path over in-patch edges
         std::vector<std::vector<int>> lup(nxy);
         for (int c = 0; c < nxy; c++)
            for (int u : up[c]) if (patch_of[u] == patch_of[c]) lup[c].push_back(u);
         std::vector<int> ll;
         detail::dag_levels(nxy, lup, ll);
         cq.lvl.assign(S, LVL_EMPTY);
         cq.patch_nlev.assign(cq.npatch, 1);
         int maxdiff = 1;
         for (int c = 0; c < nxy; c++) {
            if (ll[c] >= LVL_EMPTY) throw std::runtime_error("patch pipeline too deep");
            cq.lvl[(int64_t)patch_of[c] * P + lane_of[c]] = (uint16_t)ll[c];
            cq.patch_nlev[patch_of[c]] = std::max(cq.patch_nlev[patch_of[c]], ll[c] + 1);
            for (int u : lup[c]) maxdiff = std::max(maxdiff, ll[c] - ll[u]);
         }
         cq.ring = std::min(RING_MAX, maxdiff + 1);
         cq.nsteps = *std::max_element(cq.patch_nlev.begin(), cq.patch_nlev.end()) + pl.nz - 1;
         // sources and vectors
         cq.out_vec.assign(S, Vec2{0, 0});
         cq.in_src.assign((size_t)FIN_MAX * S, SRC_NONE);
         cq.in_vec.assign((size_t)FIN_MAX * S, Vec2{0, 0});
         cq.rout.assign((size_t)ROUT_MAX * S, -1);
         cq.in_hidx.assign((size_t)FIN_MAX * S, 0);
         std::vector<int> halo_count(cq.npatch, 0);
         for (int c = 0; c < nxy; c++) {
            const int64_t sl = (int64_t)patch_of[c] * P + lane_of[c];
            int nin = 0, nro = 0;
            const double ia = 1.0 / ms.xy_area[c];
            for (int f = 0; f < ms.xy_num_faces[c]; f++) {
               size_t a = (size_t)c * F + f;
               const double vx = ms.xy_face_cf[a] * ms.xy_face_fx[a] * ia;
               const double vy = ms.xy_face_cf[a] * ms.xy_face_fy[a] * ia;
               const int nb = ms.xy_neighbor[a];
               if (flg[a] == 1) {
                  cq.out_vec[sl].x += vx; cq.out_vec[sl].y += vy;
                  if (nb < 0 && rface_id[a] >= 0) {
                     if (nro >= ROUT_MAX) throw std::runtime_error("cell with more than 2 outgoing reflective faces");
                     cq.rout[(size_t)nro * S + sl] = rface_id[a]; nro++;
                  }
               } else if (flg[a] == 2) {
                  int32_t code = SRC_NONE;
                  int dly = 0;
                  if (nb >= 0) {
                     if (patch_of[nb] == patch_of[c] && ll[c] - ll[nb] < cq.ring) {
                        code = (SRC_LOCAL << SRC_KIND_SHIFT) | lane_of[nb];
                        dly = ll[c] - ll[nb];
                     } else
                        code = (SRC_GLOBAL << SRC_KIND_SHIFT) | (int32_t)((int64_t)patch_of[nb] * P + lane_of[nb]);
                  } else if (rface_id[a] >= 0) {
                     code = (SRC_REFL << SRC_KIND_SHIFT) | (pl.rface_axis[rface_id[a]] << SRC_AXIS_SHIFT) | rface_id[a];
                  }
                  if (code != SRC_NONE) {
                     cq.in_src[(size_t)nin * S + sl] = code;
                     cq.in_vec[(size_t)nin * S + sl] = Vec2{vx, vy};
                     // in-patch sources: pipeline steps between the writer and the reader (1 when the ring has two
                     // entries); other sources: index of the staged ("halo") entry of the patch
                     if ((code >> SRC_KIND_SHIFT) != SRC_LOCAL)
                        cq.in_hidx[(size_t)nin * S + sl] = (uint16_t)std::min(65535, halo_count[patch_of[c]]++);
                     else
                        cq.in_hidx[(size_t)nin * S + sl] = (uint16_t)dly;
                     nin++;
                  }
               }
            }
         }
         cq.max_halo = *std::max_element(halo_count.begin(), halo_count.end());
         // lanes whose flux another patch reads get a compact "edge" index: the tile kernel stores a
         // contiguous copy of them behind each psi row so that the importing patch reads whole sectors
         cq.eidx.assign(S, 255);
         std::vector<char> exported(S, 0);
         for (int f = 0; f < FIN_MAX; f++)
            for (int64_t sl = 0; sl < S; sl++) {
               const int32_t code = cq.in_src[(size_t)f * S + sl];
               if (code >= 0 && (code >> SRC_KIND_SHIFT) == SRC_GLOBAL) exported[code & SRC_PAYLOAD] = 1;
            }
         cq.max_export = 0;
         for (int p = 0; p < cq.npatch; p++) {
            int n = 0;
            for (int l = 0; l < P; l++)
               if (exported[(int64_t)p * P + l]) { cq.eidx[(int64_t)p * P + l] = (uint8_t)std::min(254, n); n++; }
            cq.max_export = std::max(cq.max_export, n);
         }
         cq.inline_ok = pl.nperim > 0 && cq.tiling == 0;
         for (int64_t sl = 0; sl < S; sl++) if (exported[sl] && (sl % P) >= PERIM_MAX) cq.inline_ok = false;
         const int maxlev = *std::max_element(cq.patch_nlev.begin(), cq.patch_nlev.end());
         // the staged tile kernel: base tiling, <= 2 incoming faces, double-buffered ring, a halo that fits the
         // 256-wide staging rows; the dataflow kernel also takes other shared tilings, a third incoming face,
         // in-patch sources two steps back and 64 staged sources
         cq.fast = cq.tiling == 0 && cq.fin <= 2 && cq.ring == 2 && cq.max_halo <= 32 && cq.max_export <= 32 && maxlev <= 31;
         cq.fast_flow = cq.tiling >= 0 && cq.fin <= FLOW_FIN && cq.ring <= 3 && cq.max_halo <= FLOW_HALO &&
                        cq.max_export <= FLOW_EXPORT && maxlev <= 31;
      };
      // candidates: the shared tilings in order (the first one the dataflow kernel can sweep, else the first acyclic
      // one), then chunks of the class's own (level, lateral) order, which are acyclic by construction
      bool placed = false;
      ClassPlan fallback;
      bool have_fallback = false;
      for (size_t tgi = 0; tgi < pl.tilings.size() && !placed; tgi++) {
         const Tiling& tg = pl.tilings[tgi];
         for (int c = 0; c < nxy; c++) { patch_of[c] = tg.slot_of_xy[c] / P; lane_of[c] = tg.slot_of_xy[c] % P; }
         ClassPlan cq = cp;
         cq.tiling = (int)tgi; cq.tiles = (tgi == 0);
         if (!try_partition(cq, tg.npatch)) continue;
         build_tables(cq);
         if (cq.fast_flow) { cp = std::move(cq); placed = true; }
         else if (!have_fallback) { fallback = std::move(cq); have_fallback = true; }
      }
      if (!placed && have_fallback) { cp = std::move(fallback); placed = true; }
      if (placed) pl.tile_classes++;
      else {
         double ox = 0, oy = 0;
         for (int m : cp.dirs) { ox += qd.directions[3*m]; oy += qd.directions[3*m+1]; }
         std::vector<int> ord(nxy);
         std::iota(ord.begin(), ord.end(), 0);
         std::vector<double> lat(nxy);
         for (int c = 0; c < nxy; c++) lat[c] = -oy * ms.xy_cx[c] + ox * ms.xy_cy[c];
         std::sort(ord.begin(), ord.end(), [&](int a, int b) {
            if (glevel[a] != glevel[b]) return glevel[a] < glevel[b];
            if (lat[a] != lat[b]) return lat[a] < lat[b];
            return a < b; });
         for (int a = 0; a < nxy; a++) { patch_of[ord[a]] = a / cap; lane_of[ord[a]] = a % cap; }
         cp.tiling = -1; cp.tiles = false;
         if (!try_partition(cp, (nxy + cap - 1) / cap)) throw std::runtime_error("internal: level-chunk partition is cyclic");
         build_tables(cp);
      }
   }
   class_flags.clear();

   // ---- chunks of directions ---------------------------------------------------------------
   pl.chunks.clear();
   pl.psi_doubles = 0;
   for (size_t ci = 0; ci < pl.classes.size(); ci++) {
      const ClassPlan& cp = pl.classes[ci];
      int n = (int)cp.dirs.size();
      int dtm = (in.opts.dt_max >= 1 && in.opts.dt_max <= DT_MAX) ? in.opts.dt_max : DT_DEFAULT;
      // classes only the three-face variant of the dataflow kernel can sweep (a third coefficient set per lane):
      // narrower chunks keep it at two CTAs per SM without spills
      if (cp.fast_flow && !cp.fast)
         dtm = (in.opts.dt_max >= 1 && in.opts.dt_max <= DT_MAX) ? std::min(in.opts.dt_max, FLOW3_DT_CAP) : FLOW3_DT_DEFAULT;
      // the fewest chunks of <= dtm directions, sizes balanced to within one (S12 on a Cartesian mesh: 21 directions
      // per octant -> 5+4+4+4+4, not 5+5+5+5+1: a one-direction chunk pays the whole per-step overhead for it)
      const int nch = (n + dtm - 1) / dtm;
      for (int c = 0, a = 0; c < nch; c++) {
         Chunk ch; ch.cls = (int)ci; ch.nd = n / nch + (c < n % nch ? 1 : 0);
         for (int d = 0; d < DT_MAX; d++) ch.m[d] = d < ch.nd ? cp.dirs[a + d] : -1;
         a += ch.nd;
         pl.chunks.push_back(ch);
      }
   }

   // ---- z chunks and the launch schedule ---------------------------------------------------
   int Kc = pl.nz;
   if (in.opts.z_chunk > 0) Kc = std::min(pl.nz, in.opts.z_chunk);
   // level-chunk patches (unstructured meshes) form a chain per ordering class: the sweep time goes like
   // (patches + z chunks) x (layers per chunk + local levels), and the local levels are few -- short z chunks
   // win (measured on a 43 561-hexagon x 100-layer core, S8: 63 / 47 / 44 / 47 ms per sweep at 32 / 16 / 8 / 4)
   else if (pl.has_z && pl.tile_classes < (int)pl.classes.size()) Kc = std::min(pl.nz, 8);
   else if (pl.has_z && in.opts.num_ranks > 2 && in.opts.wave_launch) {
      // sharded over many GPUs a rank owns too few sweeps to fill its SMs wavefront by wavefront:
      // pipeline in z as well (tasks x nzc, critical path ~ (patch levels + nzc) x (levels + nz/nzc))
      const int nzc = 2;      // measured on 4 and 8 B200s: 2-3 chunks beat 1 and 4
      if (pl.nz / nzc >= 32) Kc = (pl.nz + nzc - 1) / nzc;
   }
   pl.Kc = Kc; pl.nzc = (pl.nz + Kc - 1) / Kc;

   const int rank = in.opts.rank, nr = std::max(1, in.opts.num_ranks);
   int maxw = 0;
   for (auto& cp : pl.classes)
      for (int p = 0; p < cp.npatch; p++) maxw = std::max(maxw, cp.patch_level[p] + pl.nzc);
   pl.waves.assign(maxw, {});
   pl.owned_updates = 0;
   for (size_t ch = 0; ch < pl.chunks.size(); ch++) {
      const Chunk& c = pl.chunks[ch];
      const ClassPlan& cp = pl.classes[c.cls];
      for (int g = 0; g < pl.G; g++) {
         bool mine = in.opts.shard_mode == 1 ? (g % nr == rank) : ((int)(ch % nr) == rank);
         if (!mine) continue;
         pl.owned_updates += (int64_t)c.nd * nxy * pl.nz;
         for (int p = 0; p < cp.npatch; p++)
            for (int zc = 0; zc < pl.nzc; zc++)
               pl.waves[cp.patch_level[p] + zc].push_back(Task{(int32_t)ch, g, p, zc});
      }
   }
   // drop empty waves (ranks that own nothing at some wavefront number)
   pl.waves.erase(std::remove_if(pl.waves.begin(), pl.waves.end(),
                                 [](const std::vector<Task>& w) { return w.empty(); }), pl.waves.end());
   // psi storage: per chunk [nd][G][nz][S_class]
   for (auto& c : pl.chunks) {
      c.psi_offset = pl.psi_doubles;
      pl.psi_doubles += (int64_t)c.nd * pl.G * pl.nz * pl.classes[c.cls].S;
   }
}

}  // namespace pampa_sn
