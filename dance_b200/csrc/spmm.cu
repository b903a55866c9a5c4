// CSR SpMM  Y = act(reduce(A · X))  — the GCN / message-passing aggregate.
//
// Replaces torch.spmm(adj, support) (reference scgnn2.py:500, spagcn.py:359,
// scdsc.py:498) and DGL update_all(u_mul_e, sum|mean) (gnn.py:90,
// graphsc.py:463-465).
//
// Layout: one sub-warp group of G lanes owns one output row; each lane owns
// VPL float4 slices of the feature row.  The group streams its (col, val)
// pairs G at a time with one coalesced load, then broadcasts them with
// shuffles while every lane issues G independent 16-byte gathers of X — the
// gathers are the traffic that matters (nnz · F · 4 bytes through L2), so the
// loop is organised to keep G of them in flight per lane.  Accumulation is in
// fp32 registers in CSR order (deterministic; same order as a sequential CPU
// CSR loop, which is what the oracle does).
#include "common.cuh"

namespace b2 {

template <int G, int VPL>
__global__ void __launch_bounds__(256)
spmm_csr_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const float* __restrict__ vals, const float4* __restrict__ X4, int64_t ldx4,
                float4* __restrict__ Y4, int64_t ldy4, int32_t n_rows, int32_t F4, int reduce, int act,
                const float4* __restrict__ bias4) {
  constexpr int RPW = 32 / G;  // rows per warp
  // gathers issued back-to-back before the FMAs consume them (register budget: TCH*VPL float4)
  constexpr int TCH = VPL >= 4 ? 2 : (VPL == 2 ? 4 : (G < 8 ? G : 8));
  const int lane = threadIdx.x & 31;
  const int sub = lane / G;
  const int gl = lane % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

  for (int64_t row = warp0 * RPW + sub; row < n_rows; row += nwarps * RPW) {
    const int32_t start = __ldg(rowptr + row);
    const int32_t end = __ldg(rowptr + row + 1);
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

    // (col,val) pairs of the first PRE*G entries are fetched in ONE batch of independent loads (most kNN-graph rows
    // are shorter than that), so the dependent chain per row is rowptr → pairs → gathers instead of one round trip
    // per G-entry chunk.
    constexpr int PRE = (G <= 8) ? 4 : ((G == 16) ? 2 : 1);
    int32_t pc[PRE];
    float pw[PRE];
#pragma unroll
    for (int q = 0; q < PRE; ++q) {
      const int32_t e = start + q * G + gl;
      pc[q] = 0;
      pw[q] = 0.f;
      if (e < end) {
        pc[q] = __ldg(colidx + e);
        pw[q] = vals ? __ldg(vals + e) : 1.f;
      }
    }
    int32_t base = start;
#pragma unroll
    for (int q = 0; q < PRE; ++q) {
      if (base >= end) break;
      const int32_t c = pc[q];
      const float w = pw[q];
      const int cnt = min(G, end - base);
      if (cnt == G) {
        float4 x[TCH][VPL];
#pragma unroll
        for (int t0 = 0; t0 < G; t0 += TCH) {
#pragma unroll
          for (int t = 0; t < TCH; ++t) {
            const int32_t cc = __shfl_sync(gmask, c, t0 + t, G);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              const int j = gl + v * G;
              x[t][v] = (j < F4) ? __ldg(X4 + (int64_t)cc * ldx4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int t = 0; t < TCH; ++t) {
            const float ww = __shfl_sync(gmask, w, t0 + t, G);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              acc[v].x = fmaf(ww, x[t][v].x, acc[v].x);
              acc[v].y = fmaf(ww, x[t][v].y, acc[v].y);
              acc[v].z = fmaf(ww, x[t][v].z, acc[v].z);
              acc[v].w = fmaf(ww, x[t][v].w, acc[v].w);
            }
          }
        }
      } else {
        for (int t = 0; t < cnt; ++t) {
          const int32_t cc = __shfl_sync(gmask, c, t, G);
          const float ww = __shfl_sync(gmask, w, t, G);
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int j = gl + v * G;
            if (j < F4) {
              const float4 xv = __ldg(X4 + (int64_t)cc * ldx4 + j);
              acc[v].x = fmaf(ww, xv.x, acc[v].x);
              acc[v].y = fmaf(ww, xv.y, acc[v].y);
              acc[v].z = fmaf(ww, xv.z, acc[v].z);
              acc[v].w = fmaf(ww, xv.w, acc[v].w);
            }
          }
        }
      }
      base += G;
    }
    for (; base < end; base += G) {
      const int32_t e = base + gl;
      int32_t c = 0;
      float w = 0.f;
      if (e < end) {
        c = __ldg(colidx + e);
        w = vals ? __ldg(vals + e) : 1.f;
      }
      const int cnt = min(G, end - base);
      if (cnt == G) {
        // full chunk: G independent gathers in flight
        float4 x[TCH][VPL];
#pragma unroll
        for (int t0 = 0; t0 < G; t0 += TCH) {
#pragma unroll
          for (int t = 0; t < TCH; ++t) {
            const int32_t cc = __shfl_sync(gmask, c, t0 + t, G);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              const int j = gl + v * G;
              x[t][v] = (j < F4) ? __ldg(X4 + (int64_t)cc * ldx4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int t = 0; t < TCH; ++t) {
            const float ww = __shfl_sync(gmask, w, t0 + t, G);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              acc[v].x = fmaf(ww, x[t][v].x, acc[v].x);
              acc[v].y = fmaf(ww, x[t][v].y, acc[v].y);
              acc[v].z = fmaf(ww, x[t][v].z, acc[v].z);
              acc[v].w = fmaf(ww, x[t][v].w, acc[v].w);
            }
          }
        }
      } else {
        for (int t = 0; t < cnt; ++t) {
          const int32_t cc = __shfl_sync(gmask, c, t, G);
          const float ww = __shfl_sync(gmask, w, t, G);
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int j = gl + v * G;
            if (j < F4) {
              const float4 xv = __ldg(X4 + (int64_t)cc * ldx4 + j);
              acc[v].x = fmaf(ww, xv.x, acc[v].x);
              acc[v].y = fmaf(ww, xv.y, acc[v].y);
              acc[v].z = fmaf(ww, xv.z, acc[v].z);
              acc[v].w = fmaf(ww, xv.w, acc[v].w);
            }
          }
        }
      }
    }

    const float scale = (reduce == 1 && end > start) ? 1.f / (float)(end - start) : 1.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int j = gl + v * G;
      if (j < F4) {
        float4 o = acc[v];
        if (reduce == 1) {
          // DGL fn.mean divides the sum by the in-degree
          o.x = o.x * scale; o.y = o.y * scale; o.z = o.z * scale; o.w = o.w * scale;
        }
        if (bias4) { const float4 bb = __ldg(bias4 + j); o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
        o.x = apply_act(o.x, act); o.y = apply_act(o.y, act);
        o.z = apply_act(o.z, act); o.w = apply_act(o.w, act);
        stg_stream_f4(Y4 + row * ldy4 + j, o);
      }
    }
  }
}

template <int G, int VPL>
static int launch_spmm(const int32_t* rowptr, const int32_t* colidx, const float* vals, const float* X,
                       int64_t ldx, float* Y, int64_t ldy, int32_t n_rows, int32_t F, int reduce, int act,
                       const float* bias, cudaStream_t st) {
  constexpr int RPW = 32 / G;
  const int threads = 256;
  const int64_t warps_needed = ceil_div<int64_t>(n_rows, RPW);
  int64_t blocks = ceil_div<int64_t>(warps_needed, threads / 32);
  const int64_t max_blocks = (int64_t)sm_count() * 64;  // grid-stride beyond this
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  spmm_csr_kernel<G, VPL><<<(unsigned)blocks, threads, 0, st>>>(
      rowptr, colidx, vals, reinterpret_cast<const float4*>(X), ldx / 4, reinterpret_cast<float4*>(Y),
      ldy / 4, n_rows, F / 4, reduce, act, reinterpret_cast<const float4*>(bias));
  B2_CHECK_LAUNCH("spmm_csr_kernel");
  return B2_OK;
}

// spmm_stream.cu: nnz-stream kernel for operand rows of 32 / 64 / 128 bytes; returns 1 when it does not take the shape
int spmm_stream_dispatch(int dtype, const int32_t* rowptr, const int32_t* colidx, const float* vals, const void* X, int64_t ldx, float* Y,
                         int64_t ldy, void* Y16, int64_t ldy16, int32_t n_rows, int32_t F, int reduce, int act, const float* bias,
                         cudaStream_t st);

}  // namespace b2

extern "C" int b2_spmm_csr_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                               const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t n_rows,
                               int32_t n_cols, int32_t F, int reduce, int act, const float* bias, void* stream) {
  using namespace b2;
  B2_REQUIRE(rowptr && colidx && X && Y, "b2_spmm_csr_f32: null pointer");
  B2_REQUIRE(n_rows >= 0 && n_cols >= 0, "b2_spmm_csr_f32: negative shape");
  B2_REQUIRE(F > 0 && F % 4 == 0, "b2_spmm_csr_f32: F=%d must be a positive multiple of 4", F);
  B2_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldx >= F && ldy >= F,
             "b2_spmm_csr_f32: ldx=%lld ldy=%lld must be multiples of 4 and >= F", (long long)ldx,
             (long long)ldy);
  B2_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0,
             "b2_spmm_csr_f32: X/Y must be 16-byte aligned");
  B2_REQUIRE(reduce == 0 || reduce == 1, "b2_spmm_csr_f32: reduce must be 0 (sum) or 1 (mean)");
  B2_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "b2_spmm_csr_f32: bias must be 16-byte aligned");
  if (n_rows == 0) return B2_OK;
  cudaStream_t st = as_stream(stream);
  {
    const int rc = spmm_stream_dispatch(2, rowptr, colidx, vals, X, ldx, Y, ldy, nullptr, 0, n_rows, F, reduce, act, bias, st);
    if (rc != 1) return rc;
  }
  const int F4 = F / 4;
#define B2_SPMM_CASE(G, VPL) return launch_spmm<G, VPL>(rowptr, colidx, vals, X, ldx, Y, ldy, n_rows, F, reduce, act, bias, st)
  if (F4 <= 2) B2_SPMM_CASE(2, 1);
  if (F4 <= 4) B2_SPMM_CASE(4, 1);
  if (F4 <= 8) B2_SPMM_CASE(8, 1);
  if (F4 <= 16) B2_SPMM_CASE(16, 1);
  if (F4 <= 32) B2_SPMM_CASE(32, 1);
  if (F4 <= 64) B2_SPMM_CASE(32, 2);
  if (F4 <= 128) B2_SPMM_CASE(32, 4);
#undef B2_SPMM_CASE
  set_error("b2_spmm_csr_f32: F=%d > 512 unsupported (split the feature dimension)", F);
  return B2_ERR_UNSUPPORTED;
}
