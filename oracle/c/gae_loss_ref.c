/* TEST INFRASTRUCTURE — plain-C restatement of the Graph-AE decoder loss and its gradient, used by the tests as an
 * independent checker of oracle/port.py and (through it) of the CUDA decoder; never linked by the product.
 *
 *   logits x_ij = z_i · z_j                                   InnerProductDecoder, scgnn2.py:423-426
 *   cost = norm · mean_ij BCEwithLogits(x_ij, L_ij, pos_weight = L_ij · pw)       gae_loss_function, scgnn2.py:603-609
 *        = norm / n² · Σ_ij [ (1 - L_ij) x_ij + (1 + (w_ij - 1) L_ij) softplus(-x_ij) ],  w_ij = L_ij · pw
 *   (torch's binary_cross_entropy_with_logits with pos_weight; for L_ij = 0 the weight w_ij is 0 but multiplies L_ij = 0.)
 * L = (A + I) given as a CSR pattern.  Everything in double; O(n²·d).  dz = ∂cost/∂z (z appears on both sides of z·zᵀ).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static double softplus(double x) { return x > 0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
static double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

int gae_loss_grad_f64(const float* z, int64_t ldz, int32_t n, int32_t d, const int32_t* lab_rowptr, const int32_t* lab_colidx,
                      double norm, double pos_weight, double* loss_out, double* dz /* [n, d] */) {
  unsigned char* row = (unsigned char*)calloc((size_t)n, 1);
  if (!row) return -1;
  for (int64_t t = 0; t < (int64_t)n * d; ++t) dz[t] = 0.0;
  const double scale = norm / ((double)n * (double)n);
  double loss = 0.0;
  for (int32_t i = 0; i < n; ++i) {
    for (int32_t e = lab_rowptr[i]; e < lab_rowptr[i + 1]; ++e) row[lab_colidx[e]] = 1;
    for (int32_t j = 0; j < n; ++j) {
      double x = 0.0;
      for (int32_t c = 0; c < d; ++c) x += (double)z[(int64_t)i * ldz + c] * (double)z[(int64_t)j * ldz + c];
      const double L = row[j] ? 1.0 : 0.0;
      const double lw = 1.0 + (pos_weight * L - 1.0) * L;          /* log-weight of torch's formula */
      loss += (1.0 - L) * x + lw * softplus(-x);
      const double g = scale * ((1.0 - L) - lw * sigmoid(-x));     /* d/dx */
      for (int32_t c = 0; c < d; ++c) {
        dz[(int64_t)i * d + c] += g * (double)z[(int64_t)j * ldz + c];
        dz[(int64_t)j * d + c] += g * (double)z[(int64_t)i * ldz + c];
      }
    }
    for (int32_t e = lab_rowptr[i]; e < lab_rowptr[i + 1]; ++e) row[lab_colidx[e]] = 0;
  }
  *loss_out = scale * loss;
  free(row);
  return 0;
}
