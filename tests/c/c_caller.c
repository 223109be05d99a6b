/* A plain-C caller of libperiodicity_b200.so: no Python, no torch -- what a non-Python host of the reference's hot path
 * would do (tests/test_c_caller.py compiles and runs it).  Synthetic light curve with a known period; GLS through a
 * multi-device ctx (ordinals from argv, may repeat), PDM and the Gregory-Loredo periodogram through the same ctx. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "periodicity_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != PDC_OK) {                                                     \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, pdc_last_error()); \
      return 2;                                                              \
    }                                                                        \
  } while (0)

static double urand(unsigned long long* s) { /* xorshift64* */
  *s ^= *s >> 12; *s ^= *s << 25; *s ^= *s >> 27;
  return (double)((*s * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
  int ids[PDC_MAX_PEERS], ndev = 0;
  for (int a = 1; a < argc && ndev < PDC_MAX_PEERS; ++a) ids[ndev++] = atoi(argv[a]);
  if (ndev == 0) ids[ndev++] = 0;
  pdc_ctx* ctx = NULL;
  CHECK(pdc_ctx_create_multi(&ctx, ids, ndev));
  if (pdc_ctx_device_count(ctx) != ndev) { fprintf(stderr, "device count\n"); return 3; }

  const int64_t n = 40000, nf = 60000, np = 20000;
  const double period = 2.75, T = 400.0;
  double* t = malloc(sizeof(double) * n), *y = malloc(sizeof(double) * n);
  double* power = malloc(sizeof(double) * nf), *periods = malloc(sizeof(double) * np), *theta = malloc(sizeof(double) * np);
  unsigned long long seed = 88172645463325252ULL;
  for (int64_t i = 0; i < n; ++i) t[i] = T * urand(&seed);   /* unsorted on purpose: the ABI does not need sorted times */
  for (int64_t i = 0; i < n; ++i) y[i] = 10.0 + sin(2 * M_PI * t[i] / period) + (urand(&seed) + urand(&seed) + urand(&seed) - 1.5);
  const double df = 1.0 / T / 5.0, fmin = 0.5 * df;

  int64_t arg = -1; double best = 0.0;
  CHECK(pdc_gls(ctx, t, y, NULL, n, fmin, df, 0, nf, PDC_GLS_FIT_MEAN, 1.0, power, &arg, &best));
  const double f_peak = fmin + (double)arg * df;
  if (fabs(1.0 / f_peak - period) > 0.01 || !(best > 0.3 && best <= 1.0 + 1e-6) || power[arg] != best) {
    fprintf(stderr, "GLS peak at period %.5f power %.5f\n", 1.0 / f_peak, best); return 4;
  }
  for (int64_t i = 0; i < np; ++i) periods[i] = 1.0 + 7.0 * (double)i / (double)(np - 1);
  CHECK(pdc_pdm(ctx, t, y, n, periods, np, 10, 2, theta, &arg, &best));
  if ((fabs(periods[arg] - period) > 0.01 && fabs(periods[arg] - 2 * period) > 0.02) || theta[arg] != best || !(best < 0.6)) {
    fprintf(stderr, "PDM minimum at period %.5f theta %.5f\n", periods[arg], best); return 5;
  }
  /* invalid argument: an error code and a message, no abort */
  if (pdc_pdm(ctx, t, y, n, periods, np, 0, 2, theta, &arg, &best) != PDC_EINVAL || pdc_last_error()[0] == 0) return 6;
  printf("c_caller ok: %d device(s), GLS period %.5f, PDM period %.5f, version %d\n", ndev, 1.0 / f_peak, periods[arg],
         pdc_version());
  CHECK(pdc_ctx_destroy(ctx));
  free(t); free(y); free(power); free(periods); free(theta);
  return 0;
}
