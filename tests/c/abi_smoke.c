/* abi_smoke.c — the C ABI exercised from plain C (gcc, no C++ and no Python in between): the header must compile as C,
 * the library must link, errors must come back as codes.  With a device it assembles the 1-form mass matrix of a
 * 2x2x2 Kuhn cube through fq_assemble and checks the closed-form row count; without one it checks that
 * fq_ctx_create fails with FQ_ERR_CUDA (there is no CPU fallback).  Prints "ok gpu" / "ok nogpu". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "formoniq_b200.h"

int main(void) {
  fq_ctx* ctx = NULL;
  const int rc = fq_ctx_create(0, &ctx);
  if (rc != FQ_OK) {
    if (rc != FQ_ERR_CUDA || strlen(fq_last_error()) == 0) {
      printf("unexpected error %d: %s\n", rc, fq_last_error());
      return 1;
    }
    printf("ok nogpu\n");
    return 0;
  }
  const size_t shape[3] = {2, 2, 2};
  fq_mesh* mesh = NULL;
  if (fq_mesh_create_kuhn(ctx, 3, shape, NULL, NULL, NULL, 0.0, 0, 2, &mesh) != FQ_OK) {
    printf("mesh: %s\n", fq_last_error());
    return 1;
  }
  fq_csr* m1 = NULL;
  if (fq_assemble(ctx, mesh, FQ_MASS, 1, 1, &m1) != FQ_OK) {
    printf("assemble: %s\n", fq_last_error());
    return 1;
  }
  size_t nrows = 0, ncols = 0, nnz = 0;
  fq_csr_shape(m1, &nrows, &ncols, &nnz);
  /* E = 7 N^3 + 9 N^2 + 3 N edges on an N^3 Kuhn cube */
  if (nrows != 98 || ncols != 98 || nnz == 0 || nnz > 98 * 98) {
    printf("shape %zu x %zu, nnz %zu\n", nrows, ncols, nnz);
    return 1;
  }
  size_t* rp = (size_t*)malloc((nrows + 1) * sizeof(size_t));
  size_t* ci = (size_t*)malloc(nnz * sizeof(size_t));
  double* va = (double*)malloc(nnz * sizeof(double));
  if (fq_csr_download(ctx, m1, rp, ci, va) != FQ_OK || rp[0] != 0 || rp[nrows] != nnz) {
    printf("download: %s\n", fq_last_error());
    return 1;
  }
  /* contract violations come back as codes, not as aborts */
  fq_csr* bad = NULL;
  if (fq_assemble(ctx, mesh, 99, 1, 1, &bad) != FQ_ERR_INVALID) {
    printf("an unknown kind was accepted\n");
    return 1;
  }
  free(rp), free(ci), free(va);
  fq_csr_destroy(m1);
  fq_mesh_destroy(mesh);
  fq_ctx_destroy(ctx);
  printf("ok gpu\n");
  return 0;
}
