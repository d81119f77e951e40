// material.hpp -- materials and multigroup nuclear data of the host code.
//
// Same input grammar and derived data as the reference (src/Material.cxx:18-134,
// src/ConstantNuclearData.cxx:4-177, src/FeedbackNuclearData.hxx:62-140, src/PrecursorData.cxx),
// stored as flat per-group tables ready to be packed into pampa_sn_xs.
#pragma once

#include "input.hpp"

namespace pampa {

// One temperature point of a material's cross sections.
struct NuclearDataTable {
   int num_energy_groups = -1;
   std::vector<double> sigma_total, nu_sigma_fission, kappa_sigma_fission, sigma_transport;
   std::vector<double> sigma_scattering;          // [from][to]
   std::vector<double> diffusion_coefficient, chi_prompt, chi_delayed, chi_effective, velocity;
   double nu = 2.4355, kappa = 3.2e-11;           // defaults for kappa-sigma-f (ConstantNuclearData.hxx:27)

   int PAMPA_WARN_UNUSED read(std::ifstream& file);
   int PAMPA_WARN_UNUSED finish(double beta_total);                        // derived data
   int PAMPA_WARN_UNUSED check(int num_energy_groups, bool transient) const;
};

class Material {
  public:
   const std::string name;
   explicit Material(const std::string& name) : name(name) {}

   int PAMPA_WARN_UNUSED read(const std::string& filename);
   int PAMPA_WARN_UNUSED read(std::ifstream& file);

   bool hasNuclearData() const { return !tables.empty(); }
   bool isFuel() const { return fuel; }
   bool isBC() const { return bc; }
   bool isSplit() const { return split; }
   double beta() const { return beta_total; }
   int PAMPA_WARN_UNUSED checkNuclearData(int num_energy_groups, bool transient) const;

   // temperature interpolation of the reference: clamp below / above the table, linear inside
   void interpolate(double T, int& i1, int& i2, double& f) const;
   double sigmaTotal(int g, double T) const { return mix(&NuclearDataTable::sigma_total, g, T); }
   double sigmaNuFission(int g, double T) const { return mix(&NuclearDataTable::nu_sigma_fission, g, T); }
   double sigmaKappaFission(int g, double T) const { return mix(&NuclearDataTable::kappa_sigma_fission, g, T); }
   double chiEffective(int g, double T) const { return mix(&NuclearDataTable::chi_effective, g, T); }
   double sigmaScattering(int g, int g2, double T) const;
   int numEnergyGroups() const { return tables.empty() ? -1 : tables[0].num_energy_groups; }

  private:
   std::vector<double> temperatures;              // reference temperatures of a nuclear-data-set
   std::vector<NuclearDataTable> tables;
   double beta_total = 0.0;
   bool fuel = false, bc = false, split = false;
   double mix(std::vector<double> NuclearDataTable::*field, int g, double T) const;
};

}   // namespace pampa
