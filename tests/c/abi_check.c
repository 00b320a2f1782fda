/* TEST: include/pe_b200.h is a C header (C99, no C++), libpe_b200.so links from plain C, the POD layouts are the ones
 * the ctypes and the reference-side bindings assume, and without a GPU the library says so instead of computing anything. */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "pe_b200.h"

int main(void) {
  peb_icp_params p;
  peb_sac_params s;
  size_t lo = 0, hi = 0;
  peb_ctx* ctx = NULL;
  peb_multi* many = NULL;
  int devices[2] = {0, 1};
  int rc;
  if (sizeof(peb_icp_result) != 96 || sizeof(peb_icp_params) != 72 || sizeof(peb_cvicp_params) != 16) {
    printf("layout %zu %zu %zu\n", sizeof(peb_icp_result), sizeof(peb_icp_params), sizeof(peb_cvicp_params));
    return 1;
  }
  peb_icp_params_default(&p);
  if (p.max_iterations != 10 || p.min_correspondences != 3 || p.estimator != PEB_ESTIMATOR_SVD ||
      p.max_corr_dist != sqrt(DBL_MAX) || p.abs_mse_threshold != 1e-12 || p.euclidean_fitness_epsilon != -DBL_MAX) {
    printf("PCL 1.10 defaults wrong\n");
    return 2;
  }
  peb_sac_params_default(&s);
  if (s.max_iterations != 50 || s.probability != 0.99 || s.optimize_coefficients != 1) return 3;
  peb_multi_shard_range(13, 3, 2, &lo, &hi);
  if (lo != 10 || hi != 13) return 4;
  printf("version %s\n", peb_version());
  rc = peb_ctx_create(0, &ctx);
  if (rc != PEB_OK) {
    printf("ctx_create %d: %s\n", rc, peb_last_error(NULL));
    rc = peb_multi_create(2, devices, &many);
    printf("multi_create %d: %s\n", rc, peb_multi_last_error(NULL));
    return (ctx == NULL && many == NULL && rc != PEB_OK) ? 10 : 5;
  }
  peb_ctx_destroy(ctx);
  printf("ctx ok\n");
  return 0;
}
