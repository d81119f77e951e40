#include "quadrature.hpp"

#include <algorithm>
#include <array>
#include <map>

namespace pampa {

namespace {
struct Table { std::vector<double> mu; std::map<std::array<int, 3>, double> weight_of_class; bool renormalise; };

// S2..S8: the 7-digit constants of src/AngularQuadratureSet.cxx:15-140.  S12 (the reference stops at S8,
// :153; BASELINE config 5 asks for it): the standard LQ12 set, 7 digits like the others and used the same way
// (no renormalisation).  The constants are pinned by their defining equations, not by memory: mu_i^2 =
// mu_1^2 + (i-1) * 2 (1 - 3 mu_1^2) / (N-2) to 1e-7, and every even moment sum w mu^n = 1/(n+1), n = 2..12,
// to 6e-8 (tests/test_oracle.py::test_quadrature_tables).
// weight classes are keyed by the sorted triplet of direction-cosine indices (i <= j <= k,
// i + j + k = N/2 - 1); the point order inside an octant follows the reference's tables
const std::map<int, Table>& tables() {
   static const std::map<int, Table> t = {
      {2, {{1.0 / std::sqrt(3.0)}, {{{0, 0, 0}, 1.0}}, false}},
      {4, {{0.3500212, 0.8688903}, {{{0, 0, 1}, 1.0 / 3.0}}, false}},
      {6, {{0.2666355, 0.6815076, 0.9261808}, {{{0, 0, 2}, 0.1761263}, {{0, 1, 1}, 0.1572071}}, false}},
      {8, {{0.2182179, 0.5773503, 0.7867958, 0.9511897},
           {{{0, 0, 3}, 0.1209877}, {{0, 1, 2}, 0.0907407}, {{1, 1, 1}, 0.0925926}}, false}},
      {12, {{0.1672126, 0.4595476, 0.6280191, 0.7600210, 0.8722706, 0.9716377},
            {{{0, 0, 5}, 0.0707626}, {{0, 1, 4}, 0.0558811}, {{0, 2, 3}, 0.0373377}, {{1, 1, 3}, 0.0502819},
             {{1, 2, 2}, 0.0258513}}, false}},
   };
   return t;
}

// point order of the reference inside the first octant (src/AngularQuadratureSet.cxx:30-140)
std::vector<std::array<int, 3>> octant_points(int order) {
   switch (order) {
      case 2: return {{0, 0, 0}};
      case 4: return {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
      case 6: return {{0, 0, 2}, {0, 2, 0}, {2, 0, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
      case 8: return {{0, 0, 3}, {0, 3, 0}, {3, 0, 0}, {0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1},
                      {2, 1, 0}, {1, 1, 1}};
      default: {
         std::vector<std::array<int, 3>> pts;
         const int n = order / 2;
         for (int i = 0; i < n; i++)
            for (int j = 0; j < n - i; j++) pts.push_back({i, j, n - 1 - i - j});
         return pts;
      }
   }
}
}   // namespace

int AngularQuadratureSet::build() {
   auto it = tables().find(order);
   PAMPA_CHECK(it == tables().end(), "SN order not implemented");
   const Table& tb = it->second;
   const std::vector<std::array<int, 3>> pts = octant_points(order);
   const int per = (int)pts.size();
   num_directions = order * (order + 2);
   PAMPA_CHECK(num_directions != 8 * per, "inconsistent quadrature table");
   directions.assign((size_t)num_directions * 3, 0.0);
   weights.assign(num_directions, 0.0);
   double wsum = 0.0;
   for (int m = 0; m < per; m++) {
      std::array<int, 3> key = pts[m];
      std::sort(key.begin(), key.end());
      wsum += tb.weight_of_class.at(key);
   }
   for (int o = 0; o < 8; o++)
      for (int m = 0; m < per; m++) {
         double v[3] = {tb.mu[pts[m][0]], tb.mu[pts[m][1]], tb.mu[pts[m][2]]};
         if (tb.renormalise) {
            const double nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            for (double& c : v) c /= nrm;
         }
         const int l = o * per + m;
         directions[3 * l + 0] = (o & 1) ? -v[0] : v[0];
         directions[3 * l + 1] = (o & 2) ? -v[1] : v[1];
         directions[3 * l + 2] = (o & 4) ? -v[2] : v[2];
         std::array<int, 3> key = pts[m];
         std::sort(key.begin(), key.end());
         const double w = tb.weight_of_class.at(key);
         weights[l] = (tb.renormalise ? w / wsum : w) / 8.0;
      }
   // mirror of m about the x, y, z planes: same point in the octant with that bit flipped;
   // verified against the reference's dot-product search (AngularQuadratureSet.cxx:185-205)
   reflected_directions.assign((size_t)num_directions * 3, -1);
   for (int l = 0; l < num_directions; l++)
      for (int ax = 0; ax < 3; ax++) {
         const int r = ((l / per) ^ (1 << ax)) * per + l % per;
         double dot = 0.0;
         for (int c = 0; c < 3; c++) dot += directions[3 * r + c] * (c == ax ? -directions[3 * l + c] : directions[3 * l + c]);
         PAMPA_CHECK(dot <= 1.0 - DBL_TOL, "reflected direction not found");
         reflected_directions[3 * l + ax] = r;
      }
   return 0;
}

}   // namespace pampa
