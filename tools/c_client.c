/* Plain-C client of the drop-in boundary: proves include/gffm.h is valid C (no C++ types in any signature) and that
 * libgffm.so links from a C toolchain.  On a host without a B200 it reports the loud failure of gffm_create; on a B200 it
 * multiplies two small matrices mod 11 and prints the result.
 *   gcc -std=c11 -Wall -Iinclude tools/c_client.c -Lgpufinitefieldmatrices.jl_b200/lib -lgffm -Wl,-rpath,$PWD/gpufinitefieldmatrices.jl_b200/lib -o /tmp/c_client */
#include <stdio.h>
#include <stdlib.h>
#include "gffm.h"

int main(void) {
  int32_t ndev = -1;
  printf("%s\n", gffm_version());
  if (gffm_device_count(&ndev) != GFFM_OK) ndev = 0;
  printf("devices: %d\n", (int)ndev);
  gffm_ctx* ctx = NULL;
  int32_t st = gffm_create(0, &ctx);
  if (st != GFFM_OK) {
    printf("gffm_create failed (status %d): %s\n", (int)st, gffm_last_error());
    return st == GFFM_ERR_NO_DEVICE ? 0 : 1; /* no CPU fallback: this IS the expected outcome without a GPU */
  }
  const int64_t a[6] = {1, 2, 3, 4, 5, 6};  /* 2 x 3, column-major */
  const int64_t b[6] = {7, 8, 9, 10, 11, 12}; /* 3 x 2 */
  int64_t c[4] = {0, 0, 0, 0};
  gffm_mat *A = NULL, *B = NULL, *C = NULL;
  if (gffm_mat_create(ctx, 2, 3, 11, -1, &A) || gffm_mat_create(ctx, 3, 2, 11, -1, &B) || gffm_mat_create(ctx, 2, 2, 11, -1, &C) ||
      gffm_mat_upload(A, a, GFFM_I64, 2, 1) || gffm_mat_upload(B, b, GFFM_I64, 3, 1) ||
      gffm_gemm(C, A, B, 0, 0, GFFM_GEMM_STORE, GFFM_ALGO_AUTO) || gffm_mat_download(C, c, GFFM_I64, 2, 0)) {
    printf("error: %s\n", gffm_last_error());
    return 1;
  }
  printf("C = [%lld %lld; %lld %lld] mod 11\n", (long long)c[0], (long long)c[2], (long long)c[1], (long long)c[3]);
  gffm_mat_destroy(A); gffm_mat_destroy(B); gffm_mat_destroy(C);
  gffm_destroy(ctx);
  return 0;
}
