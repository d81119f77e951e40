// mesh.hpp -- finite-volume meshes of the host code: per-cell face tables in the reference's
// conventions (src/Mesh.hxx:12-52) plus the extruded description the device layer consumes.
//
//   CartesianMesh               src/CartesianMesh.cxx:19-414  (face order -y,+x,+y,-x,-z,+z;
//                               material 0 = void cell; boundary names -x,+x,-y,+y,-z,+z)
//   UnstructuredExtrudedMesh    src/UnstructuredExtrudedMesh.cxx:19-364 (CCW polygons x layers)
//   PartitionedMesh             src/PartitionedMesh.cxx:4-263 (the per-cell dump Mesh::writeData produces,
//                               src/Mesh.cxx:408-569); whole-domain files only, see the class
#pragma once

#include "input.hpp"

namespace pampa {

struct Cells {
   std::vector<double> volumes;
   std::vector<double> centroids;      // [N][3]
   std::vector<int> materials;         // 0-based
   std::vector<int> global_indices;
};

// ragged per-cell face arrays: face f of cell i is entry ptr[i] + f
struct Faces {
   std::vector<int> ptr;               // [N+1]
   std::vector<double> areas;
   std::vector<double> centroids;      // [nf][3]
   std::vector<double> normals;        // [nf][3]
   std::vector<int> neighbors;         // >= 0 cell, < 0 -(1-based boundary index)
   int num_faces(int i) const { return ptr[i + 1] - ptr[i]; }
};

class Mesh {
  public:
   virtual ~Mesh() {}
   virtual int PAMPA_WARN_UNUSED read(const std::string& filename) = 0;
   virtual int PAMPA_WARN_UNUSED build() = 0;

   int getNumDimensions() const { return num_dims; }
   int getNumCells() const { return num_cells; }
   int getNumCellsGlobal() const { return num_cells; }
   int getNumFacesMax() const { return num_faces_max; }
   const Cells& getCells() const { return cells; }
   const Faces& getFaces() const { return faces; }
   const std::vector<std::string>& getBoundaries() const { return boundaries; }
   const std::vector<BoundaryCondition>& getBoundaryConditions() const { return bcs; }
   int findBoundary(const std::string& name) const;
   // points and per-cell point lists in the reference's numbering (src/CartesianMesh.cxx:157-231,
   // src/UnstructuredExtrudedMesh.cxx:160-207), kept for the .vtk output (src/Mesh.cxx:397-405)
   int getNumPoints() const { return (int)(points.size() / 3); }
   const std::vector<double>& getPoints() const { return points; }
   const std::vector<int>& getCellPointPtr() const { return cell_point_ptr; }
   const std::vector<int>& getCellPoints() const { return cell_points; }
   int PAMPA_WARN_UNUSED writeVTK(const std::string& prefix, int n) const;
   // every array of the mesh in the reference's plain-text format (src/Mesh.cxx:408-569, what `mesh partitioned`
   // reads back); `digits` < 0 writes the reference's fixed 3 decimals, otherwise that many significant digits
   int PAMPA_WARN_UNUSED writeData(const std::string& filename, int digits = 17) const;
   void addBoundary(const std::string& name) { boundaries.push_back(name); }

   // extruded structure (read-only accessors the reference keeps private:
   // CartesianMesh.hxx:11-14, UnstructuredExtrudedMesh.hxx:11-29)
   int getNumXYCells() const { return num_xy_cells; }
   int getNumLayers() const { return num_layers; }
   bool hasZFaces() const { return has_z_faces; }
   const std::vector<double>& getDz() const { return dz; }
   const std::vector<int>& getXYij() const { return xy_ij; }       // structured (i,j) or empty

  protected:
   int num_dims = 0, num_cells = 0, num_faces_max = 0;
   Cells cells;
   Faces faces;
   std::vector<std::string> boundaries;
   std::vector<BoundaryCondition> bcs;   // 1-based
   int num_xy_cells = 0, num_layers = 1;
   bool has_z_faces = false;
   std::vector<double> dz;
   std::vector<int> xy_ij;
   std::vector<double> points;            // [np][3]
   std::vector<int> cell_point_ptr, cell_points;
   int PAMPA_WARN_UNUSED readBC(const std::vector<std::string>& line, std::ifstream& file);
};

class CartesianMesh : public Mesh {
  public:
   int PAMPA_WARN_UNUSED read(const std::string& filename) override;
   int PAMPA_WARN_UNUSED build() override;
   // programmatic construction (synthetic cores): materials 1-based as in the file, 0 = void
   void set(const std::vector<double>& dx, const std::vector<double>& dy, const std::vector<double>& dz,
            const std::vector<int>& materials_1based, const std::vector<BC::Type>& bc_types);

  private:
   int nx = 1, ny = 0, nz = 0;
   std::vector<double> dx{1.0}, dy{0.0};
   std::vector<int> file_materials;      // [nz][ny][nx], -1 = void
};

class UnstructuredExtrudedMesh : public Mesh {
  public:
   int PAMPA_WARN_UNUSED read(const std::string& filename) override;
   int PAMPA_WARN_UNUSED build() override;

  private:
   int num_xy_points = 0, nz = 0, xy_default_boundary = -1;
   std::vector<double> xy_points;                       // [np][2]
   std::vector<int> xy_cell_ptr, xy_cell_points;        // CCW point lists
   std::vector<std::string> xy_boundary_names;
   std::vector<std::vector<int>> xy_boundary_points;
};

// A mesh given cell by cell (volumes, centroids, face areas / centroids / normals / neighbours): the format the
// reference writes for every rank of a domain-decomposed run and as mesh_data.pmp of any run.  This build does not
// decompose the domain (the sweeps are sharded by angle set and energy group instead), so only whole-domain files
// (no ghost cells) are accepted; the extruded structure the device layer needs -- cells ordered layer by layer,
// z faces between consecutive layers -- is recovered from the face tables and checked.
class PartitionedMesh : public Mesh {
  public:
   int PAMPA_WARN_UNUSED read(const std::string& filename) override;
   int PAMPA_WARN_UNUSED build() override;

  private:
   int num_ghost_cells = 0, num_cells_global = 0;
};

}   // namespace pampa
