// quadrature.hpp -- level-symmetric angular quadrature sets (reference:
// src/AngularQuadratureSet.cxx:4-209: S2/S4/S6/S8 first-octant tables mirrored to the eight
// octants by the bit pattern of the octant index, weights normalised to 1, reflection map).
// S12 and S16 are additions (the reference stops at S8, :153) needed by the synthetic configs.
#pragma once

#include "util.hpp"

namespace pampa {

class AngularQuadratureSet {
  public:
   AngularQuadratureSet() {}
   explicit AngularQuadratureSet(int order) : order(order) {}
   int PAMPA_WARN_UNUSED build();
   int getNumDirections() const { return num_directions; }
   const std::vector<double>& getDirections() const { return directions; }          // [M][3]
   const std::vector<double>& getWeights() const { return weights; }
   const std::vector<int>& getReflectedDirections() const { return reflected_directions; }   // [M][3]

  private:
   int order = -1, num_directions = -1;
   std::vector<double> directions, weights;
   std::vector<int> reflected_directions;
};

}   // namespace pampa
