/* TEST INFRASTRUCTURE — plain-C restatement of two integer/fp kernels of the path, used by the
 * tests as an independent checker (never by the product):
 *   csr_spmm_f32    : Y = A·X in CSR order, fp32 accumulate — what torch.spmm(adj, support)
 *                     computes at reference scgnn2.py:500 (sequential per-row accumulation).
 *   knn_rank_f64    : for one query row, fp64 euclidean distances to all rows computed like
 *                     scipy's cdist (sequential sum of squared differences, sqrt) and the k
 *                     smallest under (distance, index) order after dropping rank 0 —
 *                     calculateKNNgraphDistanceMatrixStatsSingleThread, scgnn2.py:675-689.
 * Built by oracle/Makefile into oracle/_build/liboracle_c.so.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void csr_spmm_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, const float* X, int64_t ldx,
                  float* Y, int64_t ldy, int32_t n_rows, int32_t F) {
  for (int32_t r = 0; r < n_rows; ++r) {
    for (int32_t f = 0; f < F; ++f) Y[(int64_t)r * ldy + f] = 0.f;
    for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const float w = vals ? vals[e] : 1.f;
      const float* x = X + (int64_t)colidx[e] * ldx;
      for (int32_t f = 0; f < F; ++f) Y[(int64_t)r * ldy + f] += w * x[f];
    }
  }
}

typedef struct { double d; int32_t i; } pair_t;
static int cmp_pair(const void* a, const void* b) {
  const pair_t *p = (const pair_t*)a, *q = (const pair_t*)b;
  if (p->d < q->d) return -1;
  if (p->d > q->d) return 1;
  return (p->i > q->i) - (p->i < q->i);
}

int knn_rank_f64(const float* X, int64_t ldx, int32_t n, int32_t d, int32_t query, int32_t k, int32_t* idx_out,
                 double* dist_out) {
  pair_t* buf = (pair_t*)malloc(sizeof(pair_t) * (size_t)n);
  if (!buf) return -1;
  const float* q = X + (int64_t)query * ldx;
  for (int32_t j = 0; j < n; ++j) {
    const float* x = X + (int64_t)j * ldx;
    double s = 0.0;
    for (int32_t c = 0; c < d; ++c) {
      const double diff = (double)q[c] - (double)x[c];
      s += diff * diff;
    }
    buf[j].d = sqrt(s);
    buf[j].i = j;
  }
  qsort(buf, (size_t)n, sizeof(pair_t), cmp_pair);
  for (int32_t t = 0; t < k; ++t) {
    idx_out[t] = buf[t + 1].i;
    if (dist_out) dist_out[t] = buf[t + 1].d;
  }
  free(buf);
  return 0;
}
