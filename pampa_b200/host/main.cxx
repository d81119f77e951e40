// main.cxx -- stand-alone driver, the analogue of the reference's src/cxx/main.cxx:4-28:
//    pampa <input.pmp> [-verbose] [-silent]
#include <cstdio>

#include "../../include/pampa.h"

int main(int argc, char* argv[]) {
   int error = 0;
   pampa_initialize_steady_state(argc, argv, &error);
   if (error) { printf("Error in pampa_initialize().\n"); return 1; }
   pampa_solve_steady_state(&error);
   if (error) { printf("Error in pampa_solve().\n"); return 1; }
   pampa_finalize_steady_state(&error);
   if (error) { printf("Error in pampa_finalize().\n"); return 1; }
   return 0;
}
