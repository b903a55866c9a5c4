// Tensor-core version of the matrix-free Graph-AE decoder loss (all-pairs part):
//   Σ_{i,j} softplus(z_i·z_j)   and   dz_i = 2c·Σ_j σ(z_i·z_j) z_j
// (reference: InnerProductDecoder + BCE-with-logits, scgnn2.py:423-426, 603-619).
//
// Flash-attention-shaped, all on tcgen05:
//   S  = Z_I · Z_Jᵀ            tcgen05.mma kind::tf32, 3-product split (Z pre-split into hi/lo), S in TMEM
//   G  = σ(S), loss += softplus(S)   eight "elementwise" warps: tcgen05.ld S → SFU math → tcgen05.st G_hi/G_lo into TMEM
//   dZ_I += G · Z_J             tcgen05.mma with the A operand read from TMEM (G), B = Z_J from shared memory
// The N×N logits never leave the SM.  One CTA owns a 128-row block I and streams 64-column tiles J through a
// TMA ring; S and G are double-buffered in TMEM so the two MMAs of tile t overlap the SFU work of tile t±1.
// Per logit: 2 SFU ops (ex2, rcp) — the log of softplus is taken once per 32 logits on a running product.
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace b2 {
namespace gtc {

using namespace tc;

constexpr int BI = 128;          // rows per CTA (UMMA M)
constexpr int BJ = 64;           // columns per tile (UMMA N of the S product, K of the dZ product)
constexpr int DP = 32;           // padded embedding width: 32 tf32 = one 128-byte swizzle span
constexpr int STAGES = 3;
constexpr int THREADS = 384;     // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 elementwise (2 per TMEM lane quarter)
constexpr int ZI_BYTES = BI * DP * 4;            // 16 KB
constexpr int ZJ_BYTES = BJ * DP * 4;            // 8 KB
constexpr int STAGE_BYTES = 4 * ZJ_BYTES;        // hi/lo × {K-major, MN-major}
constexpr uint32_t TM_S = 0, TM_GHI = 128, TM_GLO = 256, TM_DBIG = 384, TM_DSMALL = 416, TM_COLS = 512;

struct Params {
  CUtensorMap mI_hi, mI_lo;      // [n,32] box {32,128} SWIZZLE_128B         (A of the S product)
  CUtensorMap mJk_hi, mJk_lo;    // [n,32] box {32,64}  SWIZZLE_128B         (B of the S product, K-major)
  CUtensorMap mJm_hi, mJm_lo;    // [n,32] box {32,64}  SWIZZLE_128B_ATOM_32B (B of the dZ product, MN-major)
  float* dz;                     // [n_rows, d]
  double* loss_acc;
  int n, d, row_begin, n_rows, j_chunk, j_splits;
  float coef;
};

__global__ void __launch_bounds__(256)
split_pad_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, float* __restrict__ zh, float* __restrict__ zl) {
  const int64_t total = (int64_t)n * DP;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / DP;
    const int c = (int)(t % DP);
    const float v = c < d ? z[i * ldz + c] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);   // exactly representable in tf32
    zh[t] = h;
    zl[t] = v - h;
  }
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(THREADS, 1)
gae_allpairs_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t s_zi_hi = smem_u32(smem), s_zi_lo = s_zi_hi + ZI_BYTES;
  const uint32_t s_ring = s_zi_lo + ZI_BYTES;
  uint8_t* bar_area = smem + 2 * ZI_BYTES + STAGES * STAGE_BYTES;
  const uint32_t bars = smem_u32(bar_area);
  const uint32_t zi_bar = bars;                      // 1
  const uint32_t full_bar = bars + 8;                // [STAGES] TMA → MMA
  const uint32_t stage_free = full_bar + 8 * STAGES; // [STAGES] dZ-MMA commit → TMA
  const uint32_t s_full = stage_free + 8 * STAGES;   // [2] S-MMA commit → elementwise
  const uint32_t s_empty = s_full + 16;              // [2] elementwise → S-MMA   (256 arrivals)
  const uint32_t g_full = s_empty + 16;              // [2] elementwise → dZ-MMA  (256 arrivals)
  const uint32_t g_empty = g_full + 16;              // [2] dZ-MMA commit → elementwise
  const uint32_t d_full = g_empty + 16;              // 1  last commit → epilogue
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bar_area + 8 * (2 + 2 * STAGES + 8) + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ib = blockIdx.x;                          // row block
  const int j_begin = blockIdx.y * p.j_chunk;
  const int j_end = min(p.n, j_begin + p.j_chunk);
  const int n_tiles = (j_end - j_begin + BJ - 1) / BJ;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mI_hi); tma_prefetch_desc(&p.mI_lo); tma_prefetch_desc(&p.mJk_hi);
    tma_prefetch_desc(&p.mJk_lo); tma_prefetch_desc(&p.mJm_hi); tma_prefetch_desc(&p.mJm_lo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(zi_bar, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(stage_free + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full + 8 * b, 1);
      mbar_init(s_empty + 8 * b, 256);
      mbar_init(g_full + 8 * b, 256);
      mbar_init(g_empty + 8 * b, 1);
    }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(zi_bar, 2 * ZI_BYTES);
      tma_load_2d(s_zi_hi, &p.mI_hi, zi_bar, 0, p.row_begin + ib * BI);
      tma_load_2d(s_zi_lo, &p.mI_lo, zi_bar, 0, p.row_begin + ib * BI);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(stage_free + 8 * stage, phase ^ 1);
        const uint32_t fb = full_bar + 8 * stage, st = s_ring + stage * STAGE_BYTES;
        const int j0 = j_begin + t * BJ;
        mbar_expect_tx(fb, STAGE_BYTES);
        tma_load_2d(st, &p.mJk_hi, fb, 0, j0);
        tma_load_2d(st + ZJ_BYTES, &p.mJk_lo, fb, 0, j0);
        tma_load_2d(st + 2 * ZJ_BYTES, &p.mJm_hi, fb, 0, j0);
        tma_load_2d(st + 3 * ZJ_BYTES, &p.mJm_lo, fb, 0, j0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(BI, BJ, 0, 0);   // S = Z_I (K-major) · Z_J (K-major)
      const uint32_t idesc_d = umma_idesc(BI, DP, 0, 1);   // dZ = G (TMEM) · Z_J (MN-major, N = 32)
      mbar_wait(zi_bar, 0);
      tc_fence_after();
      auto issue_s = [&](int t, int stage, uint32_t phase) {
        const int b = t & 1;
        mbar_wait(s_empty + 8 * b, ((t >> 1) & 1) ^ 1);
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        const uint32_t st = s_ring + stage * STAGE_BYTES;
        const uint32_t d_s = tmem + TM_S + (uint32_t)(b * BJ);
#pragma unroll
        for (int k = 0; k < DP / 8; ++k) {
          const uint64_t a_hi = umma_desc(s_zi_hi + k * 32, 16, 1024, 2), a_lo = umma_desc(s_zi_lo + k * 32, 16, 1024, 2);
          const uint64_t b_hi = umma_desc(st + k * 32, 16, 1024, 2), b_lo = umma_desc(st + ZJ_BYTES + k * 32, 16, 1024, 2);
          umma_tf32(d_s, a_lo, b_hi, idesc_s, k > 0);
          umma_tf32(d_s, a_hi, b_lo, idesc_s, 1);
          umma_tf32(d_s, a_hi, b_hi, idesc_s, 1);
        }
        umma_commit(s_full + 8 * b);
      };
      int stage_s = 0, stage_d = 0;
      uint32_t phase_s = 0;
      if (n_tiles > 0) { issue_s(0, 0, 0); if (++stage_s == STAGES) { stage_s = 0; phase_s ^= 1; } }
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) { issue_s(t + 1, stage_s, phase_s); if (++stage_s == STAGES) { stage_s = 0; phase_s ^= 1; } }
        const int b = t & 1;
        mbar_wait(g_full + 8 * b, (t >> 1) & 1);
        tc_fence_after();
        const uint32_t st = s_ring + stage_d * STAGE_BYTES;
        const uint32_t g_hi = tmem + TM_GHI + (uint32_t)(b * BJ), g_lo = tmem + TM_GLO + (uint32_t)(b * BJ);
#pragma unroll
        for (int k = 0; k < BJ / 8; ++k) {
          const uint64_t b_hi = umma_desc(st + 2 * ZJ_BYTES + k * 1024, 4096, 512, 1);
          const uint64_t b_lo = umma_desc(st + 3 * ZJ_BYTES + k * 1024, 4096, 512, 1);
          const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
          umma_tf32_ts(tmem + TM_DSMALL, g_lo + k * 8, b_hi, idesc_d, acc);
          umma_tf32_ts(tmem + TM_DSMALL, g_hi + k * 8, b_lo, idesc_d, 1);
          umma_tf32_ts(tmem + TM_DBIG, g_hi + k * 8, b_hi, idesc_d, acc);
        }
        umma_commit(g_empty + 8 * b);
        umma_commit(stage_free + 8 * stage_d);
        if (++stage_d == STAGES) stage_d = 0;
      }
      umma_commit(d_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== elementwise warps =====================
    const int sub = warp & 3;             // TMEM lane quarter
    const int half = (warp - 4) >> 2;     // which 32 of the tile's 64 columns
    const int row_local = ib * BI + sub * 32 + lane;
    const bool live = row_local < p.n_rows;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    float relu_sum = 0.f, lg_sum = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int b = t & 1;
      const uint32_t par = (t >> 1) & 1;
      mbar_wait(s_full + 8 * b, par);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + lane_off + TM_S + (uint32_t)(b * BJ + half * 32), v);
      tc_fence_before();
      mbar_arrive(s_empty + 8 * b);       // S[b] has been copied to registers
      const int col0 = j_begin + t * BJ + half * 32;
      uint32_t hi[32], lo[32];
      float prod = 1.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float x = __uint_as_float(v[c]);
        const float e = ex2a(-fabsf(x) * LOG2E);
        const float inv = rcpa(1.f + e);
        float s = x >= 0.f ? inv : e * inv;                     // sigmoid(x)
        if (col0 + c < j_end) { relu_sum += fmaxf(x, 0.f); prod *= (1.f + e); } else { s = 0.f; }
        const uint32_t h = __float_as_uint(s) & 0xFFFFE000u;
        hi[c] = h;
        lo[c] = __float_as_uint(s - __uint_as_float(h));
      }
      lg_sum += lg2a(prod);                                      // Σ log2(1+e) over the 32 logits of this tile
      mbar_wait(g_empty + 8 * b, par ^ 1);                       // the dZ-MMA of tile t-2 has finished reading G[b]
      tc_fence_after();
      tmem_st_32x32b_x32(tmem + lane_off + TM_GHI + (uint32_t)(b * BJ + half * 32), hi);
      tmem_st_32x32b_x32(tmem + lane_off + TM_GLO + (uint32_t)(b * BJ + half * 32), lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(g_full + 8 * b);
    }
    double loss = live ? (double)(relu_sum + LN2 * lg_sum) : 0.0;
    loss = warp_sum(loss);
    if (lane == 0 && loss != 0.0) atomicAdd(p.loss_acc, loss * (double)p.coef);
    if (half == 0) {
      mbar_wait(d_full, 0);
      tc_fence_after();
      uint32_t big[32], small[32];
      tmem_ld_32x32b_x32(tmem + lane_off + TM_DBIG, big);
      tmem_ld_32x32b_x32(tmem + lane_off + TM_DSMALL, small);
      if (live) {
        const float c2 = 2.f * p.coef;
        float* dst = p.dz + (size_t)row_local * p.d;
        for (int c = 0; c < p.d; ++c) {
          const float g = c2 * (__uint_as_float(big[c]) + __uint_as_float(small[c]));
          if (p.j_splits == 1) dst[c] += g; else atomicAdd(dst + c, g);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
}

size_t workspace_bytes(int32_t n) { return 2 * align_up((size_t)n * DP * sizeof(float), 256); }

bool eligible(int32_t n, int32_t d, int32_t n_rows) {
  if (getenv("B2_GAE_NO_TC")) return false;
  return d >= 4 && d <= DP && (int64_t)n * n_rows >= (1ll << 22);
}

// all-pairs part on the tensor cores; returns B2_ERR_UNSUPPORTED if tensor maps cannot be built
int launch(const float* z, int64_t ldz, int32_t n, int32_t d, int32_t row_begin, int32_t n_rows, float coef, float* dz,
           double* loss_acc, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < workspace_bytes(n)) return B2_ERR_UNSUPPORTED;
  float* zh = reinterpret_cast<float*>(ws);
  float* zl = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up((size_t)n * DP * sizeof(float), 256));
  {
    int64_t blocks = ceil_div<int64_t>((int64_t)n * DP, 256 * 4);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    split_pad_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ldz, n, d, zh, zl);
    B2_CHECK_LAUNCH("split_pad_kernel");
  }
  Params p;
  memset(&p, 0, sizeof(p));
  bool ok = make_tensor_map_f32(&p.mI_hi, zh, DP, (uint64_t)n, DP, BI, false) && make_tensor_map_f32(&p.mI_lo, zl, DP, (uint64_t)n, DP, BI, false) &&
            make_tensor_map_f32(&p.mJk_hi, zh, DP, (uint64_t)n, DP, BJ, false) && make_tensor_map_f32(&p.mJk_lo, zl, DP, (uint64_t)n, DP, BJ, false) &&
            make_tensor_map_f32(&p.mJm_hi, zh, DP, (uint64_t)n, DP, BJ, true) && make_tensor_map_f32(&p.mJm_lo, zl, DP, (uint64_t)n, DP, BJ, true);
  if (!ok) return B2_ERR_UNSUPPORTED;
  const int row_blocks = ceil_div(n_rows, BI);
  int j_splits = 1;
  const int target = sm_count();
  if (row_blocks < target) j_splits = min(ceil_div(target, row_blocks), ceil_div(n, 8 * BJ));
  if (j_splits < 1) j_splits = 1;
  int j_chunk = ceil_div(ceil_div(n, j_splits), BJ) * BJ;
  j_splits = ceil_div(n, j_chunk);
  p.dz = dz; p.loss_acc = loss_acc; p.n = n; p.d = d; p.row_begin = row_begin; p.n_rows = n_rows;
  p.j_chunk = j_chunk; p.j_splits = j_splits; p.coef = coef;
  const size_t smem = 2 * ZI_BYTES + STAGES * STAGE_BYTES + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(gae_allpairs_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(row_blocks, j_splits);
  gae_allpairs_tc_kernel<<<grid, THREADS, smem, st>>>(p);
  B2_CHECK_LAUNCH("gae_allpairs_tc_kernel");
  return B2_OK;
}

}  // namespace gtc
}  // namespace b2
