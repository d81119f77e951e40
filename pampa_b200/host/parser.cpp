#include "parser.hpp"
#include "vtk.hpp"

namespace pampa {

int Parser::read(const std::string& filename, Mesh** mesh, std::vector<Material*>& materials,
                 std::vector<Solver*>& solvers, std::vector<double>& dt) {
   std::ifstream file(filename, std::ios_base::in);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   while (true) {
      std::vector<std::string> line = input::get_next_line(file);
      if (line.empty()) break;
      const std::string& k = line[0];
      if (k == "mesh") {
         PAMPA_CHECK(line.size() != 3, "wrong number of arguments for keyword '" + k + "'");
         if (line[1] == "cartesian") *mesh = new CartesianMesh();
         else if (line[1] == "unstructured") *mesh = new UnstructuredExtrudedMesh();
         else if (line[1] == "partitioned") *mesh = new PartitionedMesh();   // whole-domain files only (mesh.hpp)
         else PAMPA_CHECK(true, "wrong mesh type");
         PAMPA_CHECK((*mesh)->read(line[2]), "unable to read the mesh from " + line[2]);
         PAMPA_CHECK((*mesh)->build(), "unable to build the mesh");
      } else if (k == "material") {
         PAMPA_CHECK(line.size() != 3, "wrong number of arguments for keyword '" + k + "'");
         Material* mat = new Material(line[1]);
         materials.push_back(mat);
         if (line[2] == "{") PAMPA_CHECK(mat->read(file), "unable to read the material from " + filename);
         else PAMPA_CHECK(mat->read(line[2]), "unable to read the material from " + line[2]);
         PAMPA_CHECK(mat->isBC() || mat->isSplit(),
                     "boundary-condition and split materials are heat-conduction features outside the SN path");
      } else if (k == "solver") {
         PAMPA_CHECK(line.size() < 3, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(*mesh == nullptr, "the mesh has to be defined before the solvers");
         PAMPA_CHECK(line[1] != "sn", "solver '" + line[1] + "' is outside the scope of this build: only the "
                                      "discrete-ordinates solver ('solver sn') is implemented");
         Solver* solver = new SNSolver(*mesh, materials);
         solvers.push_back(solver);
         if (line[2] == "{") PAMPA_CHECK(solver->read(file, solvers), "unable to read the solver from " + filename);
         else PAMPA_CHECK(solver->read(line[2], solvers), "unable to read the solver from " + line[2]);
      } else if (k == "dt") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         int nt;
         PAMPA_CHECK(input::read_axis(dt, nt, line[1], file), "wrong dt data");
      } else if (k == "vtk") {
         PAMPA_CHECK(line.size() < 2 || line.size() > 3, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read(vtk::on, line[1]), "wrong switch for .vtk output");
         if (line.size() == 3) PAMPA_CHECK(input::read(vtk::dn, 1, INT_MAX, line[2]), "wrong .vtk output interval");
      } else if (k == "petsc") {
         // `petsc dump <0|1>` switches the .ptc output of the solution; the other options address the
         // reference's PETSc / SLEPc linear algebra, of which there is none here: accepted and ignored
         PAMPA_CHECK(line.size() != 3, "wrong number of arguments for keyword '" + k + "'");
         if (line[1] == "dump")
            PAMPA_CHECK(input::read(ptc::dump, line[2]), "wrong switch to write the solution in PETSc format");
      } else if (k == "include") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(read(line[1], mesh, materials, solvers, dt), "unable to parse " + line[1]);
      } else {
         PAMPA_CHECK(true, "unrecognized keyword '" + k + "'");
      }
   }
   return 0;
}

}   // namespace pampa
