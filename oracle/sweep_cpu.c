/* sweep_cpu.c -- CPU restatement (C, OpenMP) of the SN source iteration on a Cartesian core.
 * TEST / BASELINE INFRASTRUCTURE ONLY: nothing in the product path links or calls this file.
 *
 * It applies the same discrete operator the reference assembles (src/SNSolver.cxx:344-631,
 * steady branch, mixed-face-interpolation 1.0, vacuum boundaries, no LS term)
 * matrix-free: one transport sweep per direction and group in upwind order, the scattering +
 * fission source of src/SNSolver.cxx:417-439, the scalar-flux quadrature of :272-299 and the
 * production integral of src/NeutronicSolver.cxx:81-116, with the power-iteration update of k.
 * The reference itself LU-factorises the monolithic matrix (src/petsc.cxx:193-197), which cannot
 * be formed at the benchmark sizes; this is the "same operator,
 * matrix-free" CPU baseline of BASELINE.md section 4, kind = "port".
 *
 * Parity: checked against oracle/pampa_oracle.py (itself pinned to the reference's goldens) by
 * tests/test_oracle.py::test_c_port_matches_oracle.
 *
 * Layouts: phi, q [g][k][j][i]; materials [k][j][i]; psi [g][m][k][j][i] is stored only when the caller
 * passes a buffer (the bench does, so that the CPU arm moves the bytes the GPU arm's store_psi = 1 moves).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
   int nx, ny, nz, G, M, nmat;
   const double *dx, *dy, *dz;
   const int32_t* mats;                 /* [nz][ny][nx] */
   const double *sigma_t, *sigma_s, *nusf, *chi;   /* [mat][g], [mat][g_from][g_to], ... */
   const double *dirs, *w;              /* [M][3], [M] */
   double* psi;                         /* optional [G][M][nz][ny][nx] angular-flux store (NULL: not kept) */
} sweep_problem;

/* Thread count used by the parallel regions.  The bench sets it explicitly: under torch.distributed.run the
 * environment carries OMP_NUM_THREADS=1, which would silently time one core. */
void sweep_cpu_set_threads(int n) {
#ifdef _OPENMP
   if (n > 0) omp_set_num_threads(n);
#else
   (void)n;
#endif
}

int sweep_cpu_threads(void) {
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

/* q = (S + F/k) phi */
static void source(const sweep_problem* p, const double* phi, double keff, double* q) {
   const int64_t n = (int64_t)p->nx * p->ny * p->nz;
   const int G = p->G;
#pragma omp parallel for schedule(static)
   for (int64_t c = 0; c < n; c++) {
      const int m = p->mats[c];
      double fis = 0.0;
      for (int g2 = 0; g2 < G; g2++) fis += p->nusf[m * G + g2] * phi[g2 * n + c];
      fis /= keff;
      for (int g = 0; g < G; g++) {
         double acc = p->chi[m * G + g] * fis;
         for (int g2 = 0; g2 < G; g2++) acc += p->sigma_s[(m * G + g2) * G + g] * phi[g2 * n + c];
         q[g * n + c] = acc;
      }
   }
}

/* phi_new[g] = sum_m w_m psi_m[g].  The (group, direction) sweeps are independent given q.  Groups are taken in
 * batches of gb so that the gb * M sweeps of a batch divide evenly over the threads (80 directions over 32 threads
 * would leave the last round half empty); each thread accumulates into a private buffer that is then summed in a
 * fixed order (deterministic, no atomics). */
static int group_batch(int G, int M, int nt) {
   int best = 1;
   double best_eff = 0.0;
   for (int gb = 1; gb <= G && gb <= 4; gb++) {
      const int tasks = gb * M, rounds = (tasks + nt - 1) / nt;
      const double eff = (double)tasks / ((double)rounds * nt);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = gb; }
   }
   return best;
}

static void sweep_all(const sweep_problem* p, const double* q, double* phi_new) {
   const int nx = p->nx, ny = p->ny, nz = p->nz, G = p->G, M = p->M;
   const int64_t n = (int64_t)nx * ny * nz;
   const int nt = sweep_cpu_threads();
   const int gb = group_batch(G, M, nt);
   double* priv = (double*)malloc(sizeof(double) * n * nt * gb);
   for (int g0 = 0; g0 < G; g0 += gb) {
      const int ng = (G - g0 < gb) ? G - g0 : gb;
#pragma omp parallel
      {
#ifdef _OPENMP
         const int tid = omp_get_thread_num();
#else
         const int tid = 0;
#endif
         double* acc0 = priv + (int64_t)tid * n * gb;
         memset(acc0, 0, sizeof(double) * n * ng);
         double* plane = (double*)malloc(sizeof(double) * nx * ny);   /* psi of the previous layer */
         double* row = (double*)malloc(sizeof(double) * nx);          /* psi of the previous row */
#pragma omp for schedule(static, 1)
         for (int task = 0; task < ng * M; task++) {
            const int g = g0 + task / M, m = task % M;
            const double* qg = q + (int64_t)g * n;
            double* acc = acc0 + (int64_t)(g - g0) * n;
            double* psi_out = p->psi ? p->psi + ((int64_t)g * M + m) * n : NULL;
            const double ox = p->dirs[3 * m], oy = p->dirs[3 * m + 1], oz = p->dirs[3 * m + 2];
            const double ax = fabs(ox), ay = fabs(oy), az = fabs(oz), wm = p->w[m];
            const int sx = ox > 0 ? 1 : -1, sy = oy > 0 ? 1 : -1, sz = oz > 0 ? 1 : -1;
            memset(plane, 0, sizeof(double) * nx * ny);
            for (int kk = 0; kk < nz; kk++) {
               const int k = sz > 0 ? kk : nz - 1 - kk;
               const double cz = az / p->dz[k];
               for (int jj = 0; jj < ny; jj++) {
                  const int j = sy > 0 ? jj : ny - 1 - jj;
                  const double cy = ay / p->dy[j];
                  double up_x = 0.0;
                  for (int ii = 0; ii < nx; ii++) {
                     const int i = sx > 0 ? ii : nx - 1 - ii;
                     const int64_t c = ((int64_t)k * ny + j) * nx + i;
                     const double cx = ax / p->dx[i];
                     const double st = p->sigma_t[p->mats[c] * G + g];
                     const double up_y = jj > 0 ? row[i] : 0.0;
                     const double up_z = plane[j * nx + i];
                     const double psi = (qg[c] + cx * up_x + cy * up_y + cz * up_z) / (st + cx + cy + cz);
                     up_x = psi; row[i] = psi; plane[j * nx + i] = psi;
                     acc[c] += wm * psi;
                     if (psi_out) psi_out[c] = psi;
                  }
               }
            }
         }
         free(plane); free(row);
#pragma omp barrier
         for (int gi = 0; gi < ng; gi++) {
#pragma omp for schedule(static)
            for (int64_t c = 0; c < n; c++) {
               double sum = 0.0;
               for (int t2 = 0; t2 < nt; t2++) sum += priv[((int64_t)t2 * gb + gi) * n + c];
               phi_new[(int64_t)(g0 + gi) * n + c] = sum;
            }
         }
      }
   }
   free(priv);
}

static double production(const sweep_problem* p, const double* phi) {
   const int64_t n = (int64_t)p->nx * p->ny * p->nz;
   double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
   for (int64_t c = 0; c < n; c++) {
      const int i = (int)(c % p->nx), j = (int)((c / p->nx) % p->ny), k = (int)(c / ((int64_t)p->nx * p->ny));
      const double vol = p->dx[i] * p->dy[j] * p->dz[k];
      const int m = p->mats[c];
      for (int g = 0; g < p->G; g++) s += vol * p->nusf[m * p->G + g] * phi[g * n + c];
   }
   return s;
}

/* Run `iters` source iterations from phi (in/out, [g][k][j][i]); returns keff. */
double sweep_cpu_iterate(const sweep_problem* p, double* phi, double keff, int iters) {
   const int64_t n = (int64_t)p->nx * p->ny * p->nz;
   double* q = (double*)malloc(sizeof(double) * n * p->G);
   double* phin = (double*)malloc(sizeof(double) * n * p->G);
   double prod = production(p, phi);
   for (int it = 0; it < iters; it++) {
      source(p, phi, keff, q);
      sweep_all(p, q, phin);
      const double pn = production(p, phin);
      keff *= pn / prod;
      prod = pn;
      memcpy(phi, phin, sizeof(double) * n * p->G);
   }
   free(q); free(phin);
   return keff;
}

/* Power iteration to |dk| < tol_k and relative L2 flux change < tol_phi. */
double sweep_cpu_solve(const sweep_problem* p, double* phi, double tol_k, double tol_phi, int max_it,
                       int* iterations) {
   const int64_t n = (int64_t)p->nx * p->ny * p->nz * p->G;
   double* old = (double*)malloc(sizeof(double) * n);
   double keff = 1.0;
   int it = 0;
   for (; it < max_it;) {
      memcpy(old, phi, sizeof(double) * n);
      const double k2 = sweep_cpu_iterate(p, phi, keff, 1);
      it++;
      double d2 = 0.0, p2 = 0.0;
      for (int64_t a = 0; a < n; a++) { d2 += (phi[a] - old[a]) * (phi[a] - old[a]); p2 += phi[a] * phi[a]; }
      const double dk = fabs(k2 - keff);
      keff = k2;
      if (it > 1 && dk < tol_k && sqrt(d2 / p2) < tol_phi) break;
   }
   free(old);
   if (iterations) *iterations = it;
   return keff;
}
