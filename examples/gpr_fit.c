/* Plain-C use of the C ABI (include/gpb200.h): no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/gpr_fit.c -o gpr_fit -Lgptest_b200 -lgpb200 -Wl,-rpath,$PWD/gptest_b200 -lm
 *   ./gpr_fit data.bin
 *
 * data.bin (little endian): int64 n, int64 d, int64 m, then X (n*d doubles, row-major), y (n), Z (m*d),
 * khyp (d+2 doubles = [l_1..l_d, sf2, sn2], what SquaredExponential.__init__ derives, GPr.py:93-97).
 * Prints the negative log marginal likelihood (GPr.py:57-69), its gradient w.r.t. the log hyper-parameters and
 * the prediction (GPr.py:45-54) - the numbers tests/test_gpu_c_example.py compares with the oracle. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "gpb200.h"

static void die(gpb_handle* h, const char* what) {
  fprintf(stderr, "%s: %s\n", what, gpb_last_error(h));
  exit(1);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s data.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int64_t n, d, m;
  if (fread(&n, 8, 1, f) != 1 || fread(&d, 8, 1, f) != 1 || fread(&m, 8, 1, f) != 1) return 2;
  double* X = malloc(sizeof(double) * n * d);
  double* y = malloc(sizeof(double) * n);
  double* Z = malloc(sizeof(double) * m * d);
  double* khyp = malloc(sizeof(double) * (d + 2));
  double* grad = malloc(sizeof(double) * (d + 2));
  double* fz = malloc(sizeof(double) * m);
  double* cov = malloc(sizeof(double) * m);
  if (fread(X, 8, n * d, f) != (size_t)(n * d) || fread(y, 8, n, f) != (size_t)n ||
      fread(Z, 8, m * d, f) != (size_t)(m * d) || fread(khyp, 8, d + 2, f) != (size_t)(d + 2)) return 2;
  fclose(f);

  gpb_handle* h = NULL;
  if (gpb_create(0, &h)) die(NULL, "gpb_create");                 /* no GPU: fails loudly, there is no fallback */
  if (gpb_set_train(h, X, n, (int32_t)d, y)) die(h, "gpb_set_train");
  double nlml = 0.0;
  int32_t info = 0;
  if (gpb_gpr_nlml(h, khyp, 0.0, &nlml, grad, &info)) die(h, "gpb_gpr_nlml");
  if (info > 0) { fprintf(stderr, "not positive definite at pivot %d\n", info); return 1; }
  printf("nlml %.17g\n", nlml);
  for (int64_t k = 0; k < d + 2; ++k) printf("grad %lld %.17g\n", (long long)k, grad[k]);
  if (gpb_gpr_predict(h, khyp, 0.0, Z, m, fz, cov, &info)) die(h, "gpb_gpr_predict");
  for (int64_t i = 0; i < m; ++i) printf("pred %lld %.17g %.17g\n", (long long)i, fz[i], cov[i]);
  printf("launches %lld\n", (long long)gpb_launch_count(h));
  gpb_destroy(h);
  free(X); free(y); free(Z); free(khyp); free(grad); free(fz); free(cov);
  return 0;
}
