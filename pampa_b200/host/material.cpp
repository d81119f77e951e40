#include "material.hpp"

namespace pampa {

int NuclearDataTable::read(std::ifstream& file) {
   auto groups = [&](std::vector<double>& v, const char* what) -> int {
      PAMPA_CHECK(num_energy_groups < 1, "energy-groups must come first");
      PAMPA_CHECK(input::read(v, num_energy_groups, 0.0, DBL_MAX, file), std::string("wrong ") + what);
      return 0;
   };
   while (true) {
      std::vector<std::string> line = input::get_next_line(file);
      if (line.empty() || line[0] == "}") break;
      const std::string& k = line[0];
      if (k == "energy-groups") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read(num_energy_groups, 1, INT_MAX, line[1]), "wrong number of energy groups");
         continue;
      }
      PAMPA_CHECK(line.size() != 1, "wrong number of arguments for keyword '" + k + "'");
      if (k == "sigma-total") { PAMPA_CHECK(groups(sigma_total, "total cross sections"), "wrong nuclear data"); }
      else if (k == "nu-sigma-fission") { PAMPA_CHECK(groups(nu_sigma_fission, "nu-fission cross sections"), "wrong nuclear data"); }
      else if (k == "kappa-sigma-fission") { PAMPA_CHECK(groups(kappa_sigma_fission, "kappa-fission cross sections"), "wrong nuclear data"); }
      else if (k == "sigma-transport") { PAMPA_CHECK(groups(sigma_transport, "transport cross sections"), "wrong nuclear data"); }
      else if (k == "sigma-scattering") {
         PAMPA_CHECK(num_energy_groups < 1, "energy-groups must come first");
         PAMPA_CHECK(input::read(sigma_scattering, num_energy_groups, num_energy_groups, 0.0, DBL_MAX, file),
                     "wrong scattering cross sections");
      }
      else if (k == "diffusion-coefficient") { PAMPA_CHECK(groups(diffusion_coefficient, "diffusion coefficients"), "wrong nuclear data"); }
      else if (k == "fission-spectrum" || k == "fission-spectrum-prompt") { PAMPA_CHECK(groups(chi_prompt, "prompt fission spectrum"), "wrong nuclear data"); }
      else if (k == "fission-spectrum-delayed") { PAMPA_CHECK(groups(chi_delayed, "delayed fission spectrum"), "wrong nuclear data"); }
      else if (k == "neutron-velocity") { PAMPA_CHECK(groups(velocity, "neutron velocity"), "wrong nuclear data"); }
      else PAMPA_CHECK(true, "unrecognized keyword '" + k + "'");
   }
   return 0;
}

int NuclearDataTable::finish(double beta_total) {
   const int G = num_energy_groups;
   if (!nu_sigma_fission.empty() || !chi_prompt.empty()) {
      PAMPA_CHECK(nu_sigma_fission.empty(), "missing nu-fission cross sections");
      PAMPA_CHECK(chi_prompt.empty(), "missing fission spectrum");
   }
   if (nu_sigma_fission.empty()) nu_sigma_fission.assign(G, 0.0);
   if (kappa_sigma_fission.empty()) {
      kappa_sigma_fission = nu_sigma_fission;
      const double f = kappa / nu;
      for (double& x : kappa_sigma_fission) x *= f;
   }
   if (chi_prompt.empty()) chi_prompt.assign(G, 0.0);
   if (chi_delayed.empty()) chi_delayed = chi_prompt;
   if (beta_total > 0.0) {
      chi_effective.assign(G, 0.0);
      for (int g = 0; g < G; g++) chi_effective[g] = (1.0 - beta_total) * chi_prompt[g] + beta_total * chi_delayed[g];
   } else {
      chi_effective = chi_prompt;
   }
   if (!sigma_transport.empty()) {
      PAMPA_CHECK(!diffusion_coefficient.empty(),
                  "transport cross sections can only be defined if diffusion coefficients are not");
      diffusion_coefficient.assign(G, 0.0);
      for (int g = 0; g < G; g++) diffusion_coefficient[g] = 1.0 / (3.0 * sigma_transport[g]);
   }
   return 0;
}

int NuclearDataTable::check(int G, bool transient) const {
   PAMPA_CHECK(num_energy_groups != G, "wrong number of energy groups");
   PAMPA_CHECK(sigma_total.empty(), "missing total cross sections");
   PAMPA_CHECK(nu_sigma_fission.empty(), "missing nu-fission cross sections");
   PAMPA_CHECK(kappa_sigma_fission.empty(), "missing kappa-fission cross sections");
   PAMPA_CHECK(sigma_scattering.empty(), "missing scattering cross sections");
   PAMPA_CHECK(chi_effective.empty(), "missing effective fission spectrum");
   if (transient) {
      PAMPA_CHECK(chi_prompt.empty(), "missing prompt fission spectrum");
      PAMPA_CHECK(chi_delayed.empty(), "missing delayed fission spectrum");
      PAMPA_CHECK(velocity.empty(), "missing neutron velocities");
   }
   return 0;
}

int Material::read(const std::string& filename) {
   std::ifstream file(filename, std::ios_base::in);
   PAMPA_CHECK(!file.is_open(), "unable to open " + filename);
   PAMPA_CHECK(read(file), "unable to read the material from " + filename);
   return 0;
}

int Material::read(std::ifstream& file) {
   while (true) {
      std::vector<std::string> line = input::get_next_line(file);
      if (line.empty() || line[0] == "}") break;
      const std::string& k = line[0];
      if (k == "nuclear-data") {
         PAMPA_CHECK(line.size() != 2 || line[1] != "{", "missing opening '{' for constant nuclear data");
         tables.assign(1, NuclearDataTable());
         temperatures.clear();
         PAMPA_CHECK(tables[0].read(file), "unable to read the constant nuclear data");
      } else if (k == "nuclear-data-set") {
         PAMPA_CHECK(line.size() != 2 || line[1] != "{", "missing opening '{' for feedback nuclear data");
         tables.clear();
         while (true) {
            std::vector<std::string> sub = input::get_next_line(file);
            if (sub.empty() || sub[0] == "}") break;
            if (sub[0] == "temperature") {
               int n;
               PAMPA_CHECK(sub.size() != 2, "wrong number of arguments for keyword 'temperature'");
               PAMPA_CHECK(input::read(n, 1, INT_MAX, sub[1]), "wrong number of temperatures");
               PAMPA_CHECK(input::read(temperatures, n, 0.0, DBL_MAX, file), "wrong temperature data");
            } else if (sub[0] == "nuclear-data") {
               PAMPA_CHECK(sub.size() != 2 || sub[1] != "{", "missing opening '{' for nuclear data");
               tables.emplace_back();
               PAMPA_CHECK(tables.back().read(file), "unable to read the nuclear data");
            } else {
               PAMPA_CHECK(true, "unrecognized keyword '" + sub[0] + "'");
            }
         }
         PAMPA_CHECK(tables.size() != temperatures.size(), "wrong number of nuclear-data blocks");
      } else if (k == "precursor-data") {
         PAMPA_CHECK(line.size() != 2 || line[1] != "{", "missing opening '{' for precursor data");
         int npg = -1;
         while (true) {
            std::vector<std::string> sub = input::get_next_line(file);
            if (sub.empty() || sub[0] == "}") break;
            std::vector<double> v;
            if (sub[0] == "precursor-groups") {
               PAMPA_CHECK(sub.size() != 2, "wrong number of arguments for keyword 'precursor-groups'");
               PAMPA_CHECK(input::read(npg, 1, INT_MAX, sub[1]), "wrong number of precursor groups");
            } else if (sub[0] == "lambda") {
               PAMPA_CHECK(input::read(v, npg, 0.0, DBL_MAX, file), "wrong precursor decay constants");
            } else if (sub[0] == "beta") {
               PAMPA_CHECK(input::read(v, npg, 0.0, DBL_MAX, file), "wrong precursor fractions");
               for (double b : v) beta_total += b;
            } else {
               PAMPA_CHECK(true, "unrecognized keyword '" + sub[0] + "'");
            }
         }
      } else if (k == "thermal-properties") {
         PAMPA_CHECK(line.size() < 2, "wrong number of arguments for keyword '" + k + "'");
         // heat-conduction data: parsed for grammar compatibility, unused by the SN path
      } else if (k == "fuel") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read(fuel, line[1]), "wrong switch for fuel materials");
      } else if (k == "bc") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read(bc, line[1]), "wrong switch for boundary-condition materials");
      } else if (k == "split") {
         PAMPA_CHECK(line.size() != 2, "wrong number of arguments for keyword '" + k + "'");
         PAMPA_CHECK(input::read(split, line[1]), "wrong switch for split materials");
      } else {
         PAMPA_CHECK(true, "unrecognized keyword '" + k + "'");
      }
   }
   for (NuclearDataTable& t : tables) PAMPA_CHECK(t.finish(beta_total), "wrong nuclear data");
   return 0;
}

int Material::checkNuclearData(int G, bool transient) const {
   for (const NuclearDataTable& t : tables) PAMPA_CHECK(t.check(G, transient), "wrong nuclear data");
   return 0;
}

void Material::interpolate(double T, int& i1, int& i2, double& f) const {
   i1 = i2 = 0; f = 0.0;
   if (temperatures.size() < 2 || T < temperatures.front()) return;
   if (T > temperatures.back()) { i1 = i2 = (int)temperatures.size() - 1; return; }
   i2 = 1;
   while (i2 < (int)temperatures.size() - 1 && temperatures[i2] < T) i2++;
   i1 = i2 - 1;
   f = (T - temperatures[i1]) / (temperatures[i2] - temperatures[i1]);
}

double Material::mix(std::vector<double> NuclearDataTable::*field, int g, double T) const {
   int i1, i2; double f;
   interpolate(T, i1, i2, f);
   return (1.0 - f) * (tables[i1].*field)[g] + f * (tables[i2].*field)[g];
}

double Material::sigmaScattering(int g, int g2, double T) const {
   int i1, i2; double f;
   interpolate(T, i1, i2, f);
   const int G = tables[0].num_energy_groups;
   return (1.0 - f) * tables[i1].sigma_scattering[(size_t)g * G + g2] + f * tables[i2].sigma_scattering[(size_t)g * G + g2];
}

}   // namespace pampa
