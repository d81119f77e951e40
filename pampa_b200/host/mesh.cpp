#include "mesh.hpp"
#include "vtk.hpp"

#include <algorithm>
#include <cmath>
#include <iomanip>

namespace pampa {

int Mesh::writeVTK(const std::string& prefix, int n) const {
   PAMPA_CHECK(vtk::write(prefix, n, points, getNumPoints(), cell_point_ptr, cell_points, num_cells, cells.materials),
               "unable to write the mesh");
   return 0;
}

int Mesh::findBoundary(const std::string& name) const {
   for (size_t i = 0; i < boundaries.size(); i++) if (boundaries[i] == name) return (int)i;
   return -1;
}

int Mesh::readBC(const std::vector<std::string>& line, std::ifstream& file) {
   PAMPA_CHECK(line.size() < 3, "wrong number of arguments for keyword 'bc'");
   if (bcs.empty()) bcs.resize(1 + boundaries.size());
   int ibc = findBoundary(line[1]);
   PAMPA_CHECK(ibc < 0 || ibc + 1 >= (int)bcs.size(), "wrong boundary name");
   unsigned l = 2;
   PAMPA_CHECK(input::read(bcs[ibc + 1], line, l, file), "wrong boundary condition");
   return 0;
}

// ------------------------------------------------------------------------------ Cartesian
int CartesianMesh::read(const std::string& filename) {
   std::ifstream file(filename, std::ios_base::in);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   while (true) {
      std::vector<std::string> line = input::get_next_line(file);
      if (line.empty()) break;
      const std::string& k = line[0];
      if (k == "dx" || k == "dy" || k == "dz") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         std::vector<double>& d = (k == "dx") ? dx : (k == "dy") ? dy : dz;
         int& n = (k == "dx") ? nx : (k == "dy") ? ny : nz;
         PAMPA_CHECK(input::read_axis(d, n, line[1], file), "wrong " + k + " data");
         if (k == "dx" || n > 1) num_dims++;
         boundaries.push_back("-" + k.substr(1));
         boundaries.push_back("+" + k.substr(1));
      } else if (k == "bc") {
         PAMPA_CHECK(readBC(line, file), "wrong boundary condition");
      } else if (k == "materials" || k == "nodal-indices") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         int n, total = nx * std::max(ny, 1) * std::max(nz, 1);
         PAMPA_CHECK(input::read(n, total, total, line[1]), "wrong number of " + k);
         std::vector<int> v;
         PAMPA_CHECK(input::read(v, total, k == "materials" ? 0 : -INT_MAX, INT_MAX, file), "wrong " + k + " data");
         if (k == "materials") {
            file_materials = v;
            for (int& m : file_materials) m--;            // 1-based -> 0-based, void -> -1
         }
      } else {
         PAMPA_CHECK(true, "unrecognized keyword '" + k + "'");
      }
   }
   return 0;
}

void CartesianMesh::set(const std::vector<double>& dx_, const std::vector<double>& dy_,
                        const std::vector<double>& dz_, const std::vector<int>& mats,
                        const std::vector<BC::Type>& bc_types) {
   dx = dx_; nx = (int)dx.size(); dy = dy_; ny = (int)dy.size(); dz = dz_; nz = (int)dz.size();
   num_dims = 1 + (ny > 1) + (nz > 1);
   boundaries = {"-x", "+x"};
   if (ny) { boundaries.push_back("-y"); boundaries.push_back("+y"); }
   if (nz) { boundaries.push_back("-z"); boundaries.push_back("+z"); }
   bcs.assign(1 + boundaries.size(), BoundaryCondition());
   for (size_t b = 0; b < boundaries.size() && b < bc_types.size(); b++) bcs[b + 1].type = bc_types[b];
   file_materials = mats;
   for (int& m : file_materials) m--;
}

int CartesianMesh::build() {
   const int nyy = std::max(ny, 1), nzz = std::max(nz, 1);
   PAMPA_CHECK((int)file_materials.size() != nx * nyy * nzz, "missing material data");
   std::vector<double> x(nx + 1, 0.0), y(nyy + 1, 0.0), z(nzz + 1, 0.0);
   for (int i = 0; i < nx; i++) x[i + 1] = x[i] + dx[i];
   for (int j = 0; j < ny; j++) y[j + 1] = y[j] + dy[j];
   for (int k = 0; k < nz; k++) z[k + 1] = z[k] + dz[k];
   auto at = [&](int k, int j, int i) { return file_materials[((size_t)k * nyy + j) * nx + i]; };

   // physical cells of the xy plane, identical in every layer
   std::vector<int> xy_id((size_t)nyy * nx, -1);
   num_xy_cells = 0;
   xy_ij.clear();
   for (int j = 0; j < nyy; j++)
      for (int i = 0; i < nx; i++)
         if (at(0, j, i) != -1) { xy_id[(size_t)j * nx + i] = num_xy_cells++; xy_ij.push_back(i); xy_ij.push_back(j); }
   for (int k = 1; k < nzz; k++)
      for (int j = 0; j < nyy; j++)
         for (int i = 0; i < nx; i++)
            PAMPA_CHECK((at(k, j, i) == -1) != (at(0, j, i) == -1), "wrong material definition");
   num_layers = nzz;
   has_z_faces = nz > 0;
   if (!has_z_faces) dz.clear();
   num_cells = num_xy_cells * nzz;
   num_faces_max = nz > 0 ? 6 : (ny > 0 ? 4 : 2);

   int bi[6];
   const char* names[6] = {"-x", "+x", "-y", "+y", "-z", "+z"};
   for (int b = 0; b < 6; b++) bi[b] = findBoundary(names[b]);
   PAMPA_CHECK(bi[0] < 0 || bi[1] < 0 || (ny > 0 && (bi[2] < 0 || bi[3] < 0)) || (nz > 0 && (bi[4] < 0 || bi[5] < 0)),
               "wrong boundary name");
   auto neighbor = [&](int k, int j, int i, int b) -> int {
      if (k < 0 || k >= nzz || j < 0 || j >= nyy || i < 0 || i >= nx || at(k, j, i) == -1) return -bi[b] - 1;
      return k * num_xy_cells + xy_id[(size_t)j * nx + i];
   };

   // mesh points (x fastest, then y, then z) and cell point lists in the gmsh order of the reference
   points.clear(); cell_point_ptr.assign(1, 0); cell_points.clear();
   for (int k = 0; k < nz + 1; k++)
      for (int j = 0; j < ny + 1; j++)
         for (int i = 0; i < nx + 1; i++) points.insert(points.end(), {x[i], y[j], z[k]});
   for (int k = 0; k < nzz; k++)
      for (int j = 0; j < nyy; j++)
         for (int i = 0; i < nx; i++) {
            if (at(k, j, i) == -1) continue;
            const int sx = nx + 1, sxy = (nx + 1) * (ny + 1);
            cell_points.push_back(i + j * sx + k * sxy);
            cell_points.push_back((i + 1) + j * sx + k * sxy);
            if (ny > 0) {
               cell_points.push_back((i + 1) + (j + 1) * sx + k * sxy);
               cell_points.push_back(i + (j + 1) * sx + k * sxy);
               if (nz > 0) {
                  cell_points.push_back(i + j * sx + (k + 1) * sxy);
                  cell_points.push_back((i + 1) + j * sx + (k + 1) * sxy);
                  cell_points.push_back((i + 1) + (j + 1) * sx + (k + 1) * sxy);
                  cell_points.push_back(i + (j + 1) * sx + (k + 1) * sxy);
               }
            }
            cell_point_ptr.push_back((int)cell_points.size());
         }

   cells.volumes.clear(); cells.centroids.clear(); cells.materials.clear(); cells.global_indices.clear();
   faces = Faces();
   faces.ptr.push_back(0);
   auto add_face = [&](double area, double cx, double cy, double cz, double n0, double n1, double n2, int nb) {
      faces.areas.push_back(area);
      faces.centroids.insert(faces.centroids.end(), {cx, cy, cz});
      faces.normals.insert(faces.normals.end(), {n0, n1, n2});
      faces.neighbors.push_back(nb);
   };
   for (int k = 0; k < nzz; k++)
      for (int j = 0; j < nyy; j++)
         for (int i = 0; i < nx; i++) {
            if (at(k, j, i) == -1) continue;
            const double dxi = dx[i], dyj = ny ? dy[j] : 0.0, dzk = nz ? dz[k] : 0.0;
            const double cx = x[i] + 0.5 * dxi, cy = y[j] + 0.5 * dyj, cz = z[k] + 0.5 * dzk;
            cells.volumes.push_back(nz ? dxi * dyj * dzk : (ny ? dxi * dyj : dxi));
            cells.centroids.insert(cells.centroids.end(), {cx, cy, cz});
            cells.materials.push_back(at(k, j, i));
            cells.global_indices.push_back((int)cells.volumes.size() - 1);
            const double ax = nz ? dyj * dzk : (ny ? dyj : 1.0), ay = nz ? dxi * dzk : dxi;
            if (ny) add_face(ay, cx, y[j], cz, 0, -1, 0, neighbor(k, j - 1, i, 2));
            add_face(ax, x[i] + dxi, cy, cz, 1, 0, 0, neighbor(k, j, i + 1, 1));
            if (ny) add_face(ay, cx, y[j] + dyj, cz, 0, 1, 0, neighbor(k, j + 1, i, 3));
            add_face(ax, x[i], cy, cz, -1, 0, 0, neighbor(k, j, i - 1, 0));
            if (nz) {
               add_face(dxi * dyj, cx, cy, z[k], 0, 0, -1, neighbor(k - 1, j, i, 4));
               add_face(dxi * dyj, cx, cy, z[k] + dzk, 0, 0, 1, neighbor(k + 1, j, i, 5));
            }
            faces.ptr.push_back((int)faces.areas.size());
         }
   return 0;
}

// ------------------------------------------------------------------------------ unstructured
int UnstructuredExtrudedMesh::read(const std::string& filename) {
   std::ifstream file(filename, std::ios_base::in);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   int num_xy = 0;
   while (true) {
      std::vector<std::string> line = input::get_next_line(file);
      if (line.empty()) break;
      const std::string& k = line[0];
      if (k == "points") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read(num_xy_points, 1, INT_MAX, line[1]), "wrong number of points");
         PAMPA_CHECK(input::read(xy_points, num_xy_points, 2, -DBL_MAX, DBL_MAX, file), "wrong point data");
      } else if (k == "cells") {
         PAMPA_CHECK(line.size() != 3, "wrong number of arguments for keyword '" + k + "'");
         int total;
         PAMPA_CHECK(input::read(num_xy, 1, INT_MAX, line[1]), "wrong number of cells");
         PAMPA_CHECK(input::read(total, 1, INT_MAX, line[2]), "wrong number of cell points");
         PAMPA_CHECK(input::read(xy_cell_ptr, xy_cell_points, num_xy, total, 0, INT_MAX, file), "wrong cell data");
         num_dims += 2;
      } else if (k == "dz") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read_axis(dz, nz, line[1], file), "wrong dz data");
         if (nz > 1) num_dims++;
         boundaries.push_back("-z");
         boundaries.push_back("+z");
      } else if (k == "boundary") {
         PAMPA_CHECK(line.size() != 3, "wrong number of arguments for keyword '" + k + "'");
         boundaries.push_back(line[1]);
         xy_boundary_names.push_back(line[1]);
         int npts;
         PAMPA_CHECK(input::read(npts, 0, INT_MAX, line[2]), "wrong number of boundary points");
         std::vector<int> pts;
         if (npts > 0) PAMPA_CHECK(input::read(pts, npts, 0, INT_MAX, file), "wrong boundary data");
         else xy_default_boundary = (int)xy_boundary_points.size();
         xy_boundary_points.push_back(pts);
      } else if (k == "bc") {
         PAMPA_CHECK(readBC(line, file), "wrong boundary condition");
      } else if (k == "materials" || k == "nodal-indices") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         int n, total = num_xy * std::max(nz, 1);
         PAMPA_CHECK(input::read(n, total, total, line[1]), "wrong number of " + k);
         std::vector<int> v;
         PAMPA_CHECK(input::read(v, total, k == "materials" ? 0 : -INT_MAX, INT_MAX, file), "wrong " + k + " data");
         if (k == "materials") {
            cells.materials = v;
            for (int& m : cells.materials) m--;
         }
      } else {
         PAMPA_CHECK(true, "unrecognized keyword '" + k + "'");
      }
   }
   num_xy_cells = num_xy;
   return 0;
}

int UnstructuredExtrudedMesh::build() {
   const int nxy = num_xy_cells, nzz = std::max(nz, 1);
   PAMPA_CHECK((int)cells.materials.size() != nxy * nzz, "missing material data");
   std::vector<double> z(nzz + 1, 0.0);
   for (int k = 0; k < nz; k++) z[k + 1] = z[k] + dz[k];
   num_layers = nzz;
   has_z_faces = nz > 0;
   num_cells = nxy * nzz;
   auto px = [&](int p) { return xy_points[2 * (size_t)p]; };
   auto py = [&](int p) { return xy_points[2 * (size_t)p + 1]; };

   // mesh points (the xy points repeated for every z level) and cell point lists: bottom polygon, then top
   points.clear(); cell_point_ptr.assign(1, 0); cell_points.clear();
   for (int k = 0; k < nz + 1; k++)
      for (int p = 0; p < num_xy_points; p++) points.insert(points.end(), {px(p), py(p), z[k]});
   for (int k = 0; k < nzz; k++)
      for (int i = 0; i < nxy; i++) {
         for (int a = xy_cell_ptr[i]; a < xy_cell_ptr[i + 1]; a++) cell_points.push_back(xy_cell_points[a] + k * num_xy_points);
         if (nz > 0)
            for (int a = xy_cell_ptr[i]; a < xy_cell_ptr[i + 1]; a++)
               cell_points.push_back(xy_cell_points[a] + (k + 1) * num_xy_points);
         cell_point_ptr.push_back((int)cell_points.size());
      }

   // polygon areas and centroids (shoelace)
   std::vector<double> area(nxy), ccx(nxy), ccy(nxy);
   int nf_max = 0;
   for (int i = 0; i < nxy; i++) {
      const int* c = &xy_cell_points[xy_cell_ptr[i]];
      const int n = xy_cell_ptr[i + 1] - xy_cell_ptr[i];
      nf_max = std::max(nf_max, n);
      double a = 0.0, sx = 0.0, sy = 0.0;
      for (int f = 0; f < n; f++) {
         const int p1 = c[f], p2 = c[(f + 1) % n];
         const double da = px(p1) * py(p2) - px(p2) * py(p1);
         a += da; sx += (px(p1) + px(p2)) * da; sy += (py(p1) + py(p2)) * da;
      }
      a *= 0.5;
      area[i] = a; ccx[i] = sx * (1.0 / (6.0 * a)); ccy[i] = sy * (1.0 / (6.0 * a));
   }
   num_faces_max = nz > 0 ? nf_max + 2 : nf_max;

   // what touches each point: cells, then listed boundaries (tagged -(global index) - 1)
   std::vector<std::vector<int>> touch(num_xy_points);
   for (int i = 0; i < nxy; i++)
      for (int a = xy_cell_ptr[i]; a < xy_cell_ptr[i + 1]; a++) touch[xy_cell_points[a]].push_back(i);
   for (size_t b = 0; b < xy_boundary_names.size(); b++) {
      const int gi = findBoundary(xy_boundary_names[b]);
      PAMPA_CHECK(gi < 0, "wrong boundary name");
      for (int p : xy_boundary_points[b]) {
         PAMPA_CHECK(p < 0 || p >= num_xy_points, "wrong boundary point");
         touch[p].push_back(-gi - 1);
      }
   }
   // neighbour across each edge: the other entity sharing both end points; later matches win and
   // unmatched edges take the default boundary by its xy ordinal (reference behaviour,
   // src/UnstructuredExtrudedMesh.cxx:249-278)
   std::vector<int> xy_neighbors(xy_cell_points.size(), 0);
   for (int i = 0; i < nxy; i++) {
      const int n = xy_cell_ptr[i + 1] - xy_cell_ptr[i];
      for (int f = 0; f < n; f++) {
         const int p1 = xy_cell_points[xy_cell_ptr[i] + f], p2 = xy_cell_points[xy_cell_ptr[i] + (f + 1) % n];
         bool found = false;
         int value = 0;
         for (int e1 : touch[p1])
            for (int e2 : touch[p2])
               if (e1 == e2 && e1 != i) { value = e1; found = true; }
         if (!found && xy_default_boundary >= 0) { value = -xy_default_boundary - 1; found = true; }
         PAMPA_CHECK(!found, "wrong mesh connectivity");
         xy_neighbors[xy_cell_ptr[i] + f] = value;
      }
   }
   const int iz0 = nz > 0 ? findBoundary("-z") : -1, iz1 = nz > 0 ? findBoundary("+z") : -1;

   cells.volumes.clear(); cells.centroids.clear(); cells.global_indices.clear();
   faces = Faces();
   faces.ptr.push_back(0);
   for (int k = 0; k < nzz; k++)
      for (int i = 0; i < nxy; i++) {
         const int n = xy_cell_ptr[i + 1] - xy_cell_ptr[i];
         const double h = nz > 0 ? dz[k] : 0.0, zc = z[k] + 0.5 * h;
         cells.volumes.push_back(nz > 0 ? area[i] * h : area[i]);
         cells.centroids.insert(cells.centroids.end(), {ccx[i], ccy[i], zc});
         cells.global_indices.push_back(k * nxy + i);
         for (int f = 0; f < n; f++) {
            const int p1 = xy_cell_points[xy_cell_ptr[i] + f], p2 = xy_cell_points[xy_cell_ptr[i] + (f + 1) % n];
            const double ex = px(p2) - px(p1), ey = py(p2) - py(p1);
            const double len = std::sqrt(ex * ex + ey * ey);
            faces.areas.push_back(nz > 0 ? len * h : len);
            faces.centroids.insert(faces.centroids.end(), {0.5 * (px(p1) + px(p2)), 0.5 * (py(p1) + py(p2)), zc});
            const double n0 = ey, n1 = -ex, nn = std::sqrt(n0 * n0 + n1 * n1);     // (dy, -dx): outward for CCW
            faces.normals.insert(faces.normals.end(), {n0 / nn, n1 / nn, 0.0});
            const int nb = xy_neighbors[xy_cell_ptr[i] + f];
            faces.neighbors.push_back(nb >= 0 ? nb + k * nxy : nb);
         }
         if (nz > 0) {
            const int ic = k * nxy + i;
            faces.areas.push_back(area[i]);
            faces.centroids.insert(faces.centroids.end(), {ccx[i], ccy[i], z[k]});
            faces.normals.insert(faces.normals.end(), {0.0, 0.0, -1.0});
            faces.neighbors.push_back(k == 0 ? -iz0 - 1 : ic - nxy);
            faces.areas.push_back(area[i]);
            faces.centroids.insert(faces.centroids.end(), {ccx[i], ccy[i], z[k] + h});
            faces.normals.insert(faces.normals.end(), {0.0, 0.0, 1.0});
            faces.neighbors.push_back(k == nz - 1 ? -iz1 - 1 : ic + nxy);
         }
         faces.ptr.push_back((int)faces.areas.size());
      }
   if (!has_z_faces) dz.clear();
   return 0;
}

// ------------------------------------------------------------------------------ mesh data / partitioned
int Mesh::writeData(const std::string& filename, int digits) const {
   std::ofstream file(filename, std::ios_base::out);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   if (digits < 0) file << std::fixed << std::setprecision(3);       // the reference's own precision
   else file << std::setprecision(digits);
   const int np = getNumPoints();
   file << "points " << np << std::endl;
   for (int i = 0; i < np; i++) file << points[3 * i] << " " << points[3 * i + 1] << " " << points[3 * i + 2] << std::endl;
   file << std::endl;
   file << "cells " << num_cells << " " << 0 << " " << num_cells << std::endl;
   file << "cell-points " << num_cells << " " << cell_points.size() << std::endl;
   for (int i = 0; i < num_cells; i++) {
      for (int j = cell_point_ptr[i]; j < cell_point_ptr[i + 1]; j++) file << (j > cell_point_ptr[i] ? " " : "") << cell_points[j];
      file << std::endl;
   }
   file << std::endl;
   file << "cell-volumes " << num_cells << std::endl;
   for (int i = 0; i < num_cells; i++) file << cells.volumes[i] << std::endl;
   file << std::endl;
   file << "cell-centroids " << num_cells << std::endl;
   for (int i = 0; i < num_cells; i++)
      file << cells.centroids[3 * (size_t)i] << " " << cells.centroids[3 * (size_t)i + 1] << " " << cells.centroids[3 * (size_t)i + 2] << std::endl;
   file << std::endl;
   file << "cell-materials " << num_cells << std::endl;
   for (int i = 0; i < num_cells; i++) file << cells.materials[i] + 1 << std::endl;
   file << std::endl;
   file << "cell-global-indices " << num_cells << std::endl;
   for (int i = 0; i < num_cells; i++) file << i << std::endl;
   file << std::endl;
   file << "faces " << num_cells << std::endl;
   for (int i = 0; i < num_cells; i++) file << faces.num_faces(i) << std::endl;
   file << std::endl;
   const size_t nf = faces.areas.size();
   file << "face-areas " << num_cells << " " << nf << std::endl;
   for (int i = 0; i < num_cells; i++) {
      for (int f = faces.ptr[i]; f < faces.ptr[i + 1]; f++) file << (f > faces.ptr[i] ? " " : "") << faces.areas[f];
      file << std::endl;
   }
   file << std::endl;
   file << "face-centroids " << num_cells << " " << 3 * nf << std::endl;
   for (size_t f = 0; f < nf; f++) file << faces.centroids[3 * f] << " " << faces.centroids[3 * f + 1] << " " << faces.centroids[3 * f + 2] << std::endl;
   file << std::endl;
   file << "face-normals " << num_cells << " " << 3 * nf << std::endl;
   for (size_t f = 0; f < nf; f++) file << faces.normals[3 * f] << " " << faces.normals[3 * f + 1] << " " << faces.normals[3 * f + 2] << std::endl;
   file << std::endl;
   file << "face-neighbors " << num_cells << " " << nf << std::endl;
   for (int i = 0; i < num_cells; i++) {
      for (int f = faces.ptr[i]; f < faces.ptr[i + 1]; f++) file << (f > faces.ptr[i] ? " " : "") << faces.neighbors[f];
      file << std::endl;
   }
   file << std::endl;
   for (const std::string& b : boundaries) file << "boundary " << b << std::endl;
   for (size_t i = 1; i < bcs.size(); i++) {
      file << "bc " << boundaries[i - 1];
      switch (bcs[i].type) {
         case BC::VACUUM: file << " vacuum"; break;
         case BC::REFLECTIVE: file << " reflective"; break;
         case BC::ROBIN: file << " robin"; break;
         case BC::DIRICHLET: file << " dirichlet"; break;
         case BC::ADIABATIC: file << " adiabatic"; break;
         case BC::CONVECTION: file << " convection"; break;
         default: file << " none"; break;
      }
      for (double x : bcs[i].parameters) file << " " << x;
      file << std::endl;
   }
   file << std::endl;
   return 0;
}

int PartitionedMesh::read(const std::string& filename) {
   std::ifstream file(filename, std::ios_base::in);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   int nf_total = 0;
   std::vector<int> nfaces;
   while (true) {
      std::vector<std::string> line = input::get_next_line(file);
      if (line.empty()) break;
      const std::string& k = line[0];
      auto rows = [&](int want) -> int {                 // "<keyword> <rows> [<total>]": check the row count
         int n;
         PAMPA_CHECK(line.size() < 2 || input::read(n, want, want, line[1]), "wrong number of rows for keyword '" + k + "'");
         return 0;
      };
      if (k == "points") {
         int np;
         PAMPA_CHECK(line.size() != 2 || input::read(np, 1, INT_MAX, line[1]), "wrong number of points");
         PAMPA_CHECK(input::read(points, np, 3, -DBL_MAX, DBL_MAX, file), "wrong point data");
      } else if (k == "cells") {
         PAMPA_CHECK(line.size() != 4, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read(num_cells, 1, INT_MAX, line[1]), "wrong number of cells");
         PAMPA_CHECK(input::read(num_ghost_cells, 0, INT_MAX, line[2]), "wrong number of ghost cells");
         PAMPA_CHECK(input::read(num_cells_global, 1, INT_MAX, line[3]), "wrong global number of cells");
         PAMPA_CHECK(num_ghost_cells != 0 || num_cells_global != num_cells,
                     "this file is one rank's part of a domain decomposition (it has ghost cells): the sweeps of this "
                     "build are sharded by angle set and energy group over the whole mesh, give the original mesh");
      } else if (k == "cell-points") {
         int total;
         PAMPA_CHECK(line.size() != 3 || rows(num_cells) || input::read(total, 1, INT_MAX, line[2]), "wrong number of cell points");
         PAMPA_CHECK(input::read(cell_point_ptr, cell_points, num_cells, total, 0, INT_MAX, file), "wrong cell-point data");
      } else if (k == "cell-volumes") {
         PAMPA_CHECK(line.size() != 2 || rows(num_cells), "wrong number of cell volumes");
         PAMPA_CHECK(input::read(cells.volumes, num_cells, 0.0, DBL_MAX, file), "wrong cell-volume data");
      } else if (k == "cell-centroids") {
         PAMPA_CHECK(line.size() != 2 || rows(num_cells), "wrong number of cell centroids");
         PAMPA_CHECK(input::read(cells.centroids, num_cells, 3, -DBL_MAX, DBL_MAX, file), "wrong cell-centroid data");
      } else if (k == "cell-materials") {
         PAMPA_CHECK(line.size() != 2 || rows(num_cells), "wrong number of cell materials");
         PAMPA_CHECK(input::read(cells.materials, num_cells, 1, INT_MAX, file), "wrong cell-material data");
         for (int& m : cells.materials) m--;
      } else if (k == "cell-nodal-indices" || k == "cell-global-indices") {
         std::vector<int> v;
         PAMPA_CHECK(line.size() != 2 || rows(num_cells), "wrong number of cell indices");
         PAMPA_CHECK(input::read(v, num_cells, 0, INT_MAX, file), "wrong cell-index data");
         if (k == "cell-global-indices") cells.global_indices = v;
      } else if (k == "faces") {
         PAMPA_CHECK(line.size() != 2 || rows(num_cells), "wrong number of cells");
         PAMPA_CHECK(input::read(nfaces, num_cells, 1, INT_MAX, file), "wrong face data");
         faces.ptr.assign(num_cells + 1, 0);
         for (int i = 0; i < num_cells; i++) { faces.ptr[i + 1] = faces.ptr[i] + nfaces[i]; num_faces_max = std::max(num_faces_max, nfaces[i]); }
         nf_total = faces.ptr[num_cells];
      } else if (k == "face-areas") {
         PAMPA_CHECK(line.size() != 3 || rows(num_cells) || nf_total == 0, "wrong number of face areas");
         PAMPA_CHECK(input::read(faces.areas, nf_total, 0.0, DBL_MAX, file), "wrong face-area data");
      } else if (k == "face-centroids" || k == "face-normals") {
         PAMPA_CHECK(line.size() != 3 || rows(num_cells) || nf_total == 0, "wrong number of face vectors");
         PAMPA_CHECK(input::read(k == "face-centroids" ? faces.centroids : faces.normals, nf_total, 3, -DBL_MAX, DBL_MAX, file),
                     "wrong face-vector data");
      } else if (k == "face-neighbors") {
         PAMPA_CHECK(line.size() != 3 || rows(num_cells) || nf_total == 0, "wrong number of face neighbors");
         PAMPA_CHECK(input::read(faces.neighbors, nf_total, -INT_MAX, INT_MAX, file), "wrong face-neighbor data");
      } else if (k == "boundary") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         boundaries.push_back(line[1]);
      } else if (k == "bc") {
         PAMPA_CHECK(readBC(line, file), "wrong boundary condition");
      } else {
         PAMPA_CHECK(true, "unrecognized keyword '" + k + "'");
      }
   }
   return 0;
}

int PartitionedMesh::build() {
   const size_t nf = faces.ptr.empty() ? 0 : (size_t)faces.ptr[num_cells];
   PAMPA_CHECK(num_cells < 1 || (int)cells.volumes.size() != num_cells || cells.centroids.size() != 3 * (size_t)num_cells ||
               (int)cells.materials.size() != num_cells || nf == 0 || faces.areas.size() != nf ||
               faces.centroids.size() != 3 * nf || faces.normals.size() != 3 * nf || faces.neighbors.size() != nf,
               "missing mesh data");
   if (bcs.empty()) bcs.resize(1 + boundaries.size());
   if (cells.global_indices.empty()) { cells.global_indices.resize(num_cells); for (int i = 0; i < num_cells; i++) cells.global_indices[i] = i; }
   // extruded structure: cells whose -z face is a boundary (or that have no z faces at all) form the first layer,
   // and cell i + num_xy_cells sits on top of cell i
   auto zface = [&](int f) { return std::fabs(faces.normals[3 * (size_t)f + 2]) > 0.5; };
   has_z_faces = false;
   bool has_y = false;
   for (size_t f = 0; f < nf; f++) { has_z_faces |= zface((int)f); has_y |= std::fabs(faces.normals[3 * f + 1]) > 0.5; }
   num_dims = has_z_faces ? 3 : (has_y ? 2 : 1);
   if (!has_z_faces) { num_xy_cells = num_cells; num_layers = 1; dz.clear(); return 0; }
   num_xy_cells = 0;
   for (int i = 0; i < num_cells; i++) {
      bool bottom = false;
      for (int f = faces.ptr[i]; f < faces.ptr[i + 1]; f++)
         if (zface(f) && faces.normals[3 * (size_t)f + 2] < 0.0 && faces.neighbors[f] < 0) bottom = true;
      if (!bottom) break;
      num_xy_cells++;
   }
   PAMPA_CHECK(num_xy_cells < 1 || num_cells % num_xy_cells != 0, "the mesh is not an extruded mesh");
   num_layers = num_cells / num_xy_cells;
   dz.assign(num_layers, 0.0);
   for (int k = 0; k < num_layers; k++)
      for (int c = 0; c < num_xy_cells; c++) {
         const int i = k * num_xy_cells + c;
         double area_z = 0.0;
         for (int f = faces.ptr[i]; f < faces.ptr[i + 1]; f++) {
            if (!zface(f)) continue;
            const bool up = faces.normals[3 * (size_t)f + 2] > 0.0;
            const int want = up ? (k + 1 < num_layers ? i + num_xy_cells : -1) : (k > 0 ? i - num_xy_cells : -1);
            PAMPA_CHECK(want >= 0 ? faces.neighbors[f] != want : faces.neighbors[f] >= 0, "the mesh is not an extruded mesh");
            if (up) area_z = faces.areas[f];
         }
         PAMPA_CHECK(!(area_z > 0.0), "the mesh is not an extruded mesh");
         const double h = cells.volumes[i] / area_z;
         if (c == 0) dz[k] = h;
         PAMPA_CHECK(std::fabs(h - dz[k]) > 1.0e-6 * dz[k], "the mesh is not an extruded mesh");
      }
   return 0;
}

}   // namespace pampa
