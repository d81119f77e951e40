#include "vtk.hpp"

#include <algorithm>
#include <cstdint>
#include <cstring>

namespace pampa {
namespace vtk {

bool on = false;
int dn = 1;

int write(const std::string& prefix, int n, const std::vector<double>& points, int num_points,
          const std::vector<int>& cell_ptr, const std::vector<int>& cell_points, int num_cells,
          const std::vector<int>& materials) {
   if (!on || (n % dn != 0)) return 0;
   const std::string filename = n < 0 ? prefix + ".vtk" : prefix + "_" + std::to_string(n / dn) + ".vtk";
   std::ofstream file(filename, std::ios_base::out);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   file << std::scientific << std::setprecision(PRECISION);
   file << "# vtk DataFile Version 3.0\n" << "FVM mesh\n" << "ASCII\n" << "DATASET UNSTRUCTURED_GRID\n\n";
   file << "POINTS " << num_points << " double\n";
   for (int i = 0; i < num_points; i++)
      file << points[3 * (size_t)i] << " " << points[3 * (size_t)i + 1] << " " << points[3 * (size_t)i + 2] << "\n";
   file << "\n";
   file << "CELLS " << num_cells << " " << num_cells + (int)cell_points.size() << "\n";
   for (int i = 0; i < num_cells; i++) {
      file << cell_ptr[i + 1] - cell_ptr[i];
      for (int a = cell_ptr[i]; a < cell_ptr[i + 1]; a++) file << " " << cell_points[a];
      file << "\n";
   }
   file << "\n";
   file << "CELL_TYPES " << num_cells << "\n";
   for (int i = 0; i < num_cells; i++) {
      // line, triangle, quad, wedge, hexahedron, hexagonal prism (src/vtk.cxx:80-98)
      int type = 0;
      switch (cell_ptr[i + 1] - cell_ptr[i]) {
         case 2: type = 3; break;
         case 3: type = 5; break;
         case 4: type = 9; break;
         case 6: type = 13; break;
         case 8: type = 12; break;
         case 12: type = 16; break;
         default: PAMPA_CHECK(true, "wrong cell type");
      }
      file << type << "\n";
   }
   file << "\n";
   file << "CELL_DATA " << num_cells << "\n\n";
   file << "SCALARS materials double 1\n" << "LOOKUP_TABLE default\n";
   for (int i = 0; i < num_cells; i++) file << materials[i] + 1 << "\n";
   file << "\n";
   PAMPA_CHECK(!file.good(), "unable to write " + filename);
   return 0;
}

int write(const std::string& prefix, int n, const std::string& name, const double* v, int num_cells, int num_groups,
          int num_directions) {
   if (!on || (n % dn != 0)) return 0;
   const std::string filename = prefix + "_" + std::to_string(n / dn) + ".vtk";
   std::ofstream file(filename, std::ios_base::app);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   file << std::scientific << std::setprecision(PRECISION);
   const size_t div = (size_t)num_directions * num_groups;
   for (int g = 0; g < num_groups; g++)
      for (int m = 0; m < num_directions; m++) {
         file << "SCALARS " << name;
         if (num_groups > 1) file << "_" << (g + 1);
         if (num_directions > 1) file << "_" << (m + 1);
         file << " double 1\n" << "LOOKUP_TABLE default\n";
         size_t iv = (size_t)g * num_directions + m;
         for (int i = 0; i < num_cells; i++, iv += div) file << v[iv] << "\n";
         file << "\n";
      }
   PAMPA_CHECK(!file.good(), "unable to write " + filename);
   return 0;
}

}   // namespace vtk

namespace ptc {

bool dump = false;

namespace {
void put_be32(std::ofstream& f, uint32_t x) {
   const unsigned char b[4] = {(unsigned char)(x >> 24), (unsigned char)(x >> 16), (unsigned char)(x >> 8), (unsigned char)x};
   f.write((const char*)b, 4);
}
}   // namespace

int write(const std::string& prefix, int n, const double* v, long count) {
   if (!dump) return 0;
   PAMPA_CHECK(count < 0 || count > INT_MAX, "vector too long for the PETSc binary format with 32-bit indices");
   const std::string filename = prefix + "_" + std::to_string(n) + ".ptc";
   std::ofstream file(filename, std::ios_base::out | std::ios_base::binary);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   put_be32(file, 1211214u);                  // VEC_FILE_CLASSID
   put_be32(file, (uint32_t)count);
   std::vector<unsigned char> buf((size_t)std::min<long>(count, 1 << 16) * 8);
   for (long i0 = 0; i0 < count; i0 += 1 << 16) {
      const long m = std::min<long>(1 << 16, count - i0);
      for (long i = 0; i < m; i++) {
         uint64_t bits;
         std::memcpy(&bits, v + i0 + i, 8);
         for (int b = 0; b < 8; b++) buf[(size_t)i * 8 + b] = (unsigned char)(bits >> (56 - 8 * b));
      }
      file.write((const char*)buf.data(), m * 8);
   }
   PAMPA_CHECK(!file.good(), "unable to write " + filename);
   return 0;
}

}   // namespace ptc
}   // namespace pampa
