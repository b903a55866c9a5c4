// Tensor-core version of the matrix-free Graph-AE decoder loss (all-pairs part):
//   Σ_{i,j} softplus(z_i·z_j)   and   dz_i = 2c·Σ_j σ(z_i·z_j) z_j
// (reference: InnerProductDecoder + BCE-with-logits, scgnn2.py:423-426, 603-619).
//
// Flash-attention-shaped, all on tcgen05:
//   S  = Z_I · Z_Jᵀ            tcgen05.mma kind::tf32, 3-product split (Z pre-split into hi/lo), S in TMEM
//   G  = σ(S), loss += softplus(S)   sixteen "elementwise" warps: tcgen05.ld S → SFU math → tcgen05.st G_hi/G_lo into TMEM
//   dZ_I += G · Z_J             tcgen05.mma with the A operand read from TMEM (G), B = Z_Jᵀ tile from shared memory
// The N×N logits never leave the SM.  One CTA owns a 128-row block I and streams 64-column tiles J through a
// TMA ring; S and G are double-buffered in TMEM so the two MMAs of tile t overlap the SFU work of tile t±1.
// Per logit: 2 SFU ops (ex2, rcp) — the log of softplus is taken once per 16 logits on a running product.
//
// Operand layouts (embedding width d ≤ 16, zero-padded to 16):
//   Z16  [n,16]  rows of 64 B  → K-major SWIZZLE_64B tiles: A (128 rows) and B (64 rows) of the S product, K = 16
//   ZT   [16,n'] transposed    → K-major SWIZZLE_128B tiles [16 rows(d) × 32 j]: B of the dZ product (N = 16, K = j)
// Round-1 measurements that shaped this: each CTA streams every row of Z, so the TMA ring is L2-bandwidth bound
// unless the per-tile bytes are small (16 KB here), and a tcgen05.mma of this size costs ~60 cycles regardless of N,
// so consecutive k-steps alternate between independent accumulators.
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace b2 {
namespace gtc {

using namespace tc;

constexpr int BI = 128;          // rows per CTA (UMMA M)
constexpr int BJ = 64;           // columns per tile (UMMA N of the S product, K of the dZ product)
constexpr int DW = 16;           // embedding width handled (d ≤ 16, zero-padded)
constexpr int STAGES = 4;
constexpr int EW_WARPS = 16;     // elementwise warps: 4 per TMEM lane quarter (= per SM sub-partition), 16 columns each
constexpr int EW_THREADS = EW_WARPS * 32;
constexpr int EW_COLS = BJ / (EW_WARPS / 4);
constexpr int THREADS = 128 + EW_THREADS;   // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. elementwise
constexpr int ZI_BYTES = BI * DW * 4;            // 8 KB  (hi or lo)
constexpr int ZJ_BYTES = BJ * DW * 4;            // 4 KB  (hi or lo), K-major rows of 64 B
constexpr int ZT_BYTES = DW * BJ * 4;            // 4 KB  (hi or lo): two boxes of [16 rows x 128 B]
constexpr int STAGE_BYTES = 2 * ZJ_BYTES + 2 * ZT_BYTES;   // 16 KB
constexpr uint32_t TM_S = 0, TM_GHI = 128, TM_GLO = 256, TM_D = 384, TM_COLS = 512;
// dZ accumulators at TM_D + 16·{0: big A, 1: small A, 2: big B, 3: small B}

struct Params {
  CUtensorMap mI_hi, mI_lo;      // Z16 [n,16] box {16,128} SWIZZLE_64B   (A of the S product)
  CUtensorMap mJ_hi, mJ_lo;      // Z16 [n,16] box {16,64}  SWIZZLE_64B   (B of the S product)
  CUtensorMap mT_hi, mT_lo;      // ZT  [16,n'] box {32,16} SWIZZLE_128B  (B of the dZ product)
  float* dz;                     // [n_rows, d]
  double* loss_acc;
  int n, d, row_begin, n_rows, j_chunk, j_splits;
  float coef;
};

// z [n,d] → hi/lo tf32 split, as Z16 (row-major, padded to 16) and ZT (transposed, row pitch npad)
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, int64_t npad, float* __restrict__ z16h,
             float* __restrict__ z16l, float* __restrict__ zth, float* __restrict__ ztl) {
  const int64_t total = npad * DW;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / DW;
    const int c = (int)(t % DW);
    const float v = (c < d && i < n) ? z[i * ldz + c] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);   // exactly representable in tf32
    if (i < n) { z16h[t] = h; z16l[t] = v - h; }
    zth[(int64_t)c * npad + i] = h;
    ztl[(int64_t)c * npad + i] = v - h;
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
  const uint32_t s_empty = s_full + 16;              // [2] elementwise → S-MMA
  const uint32_t g_full = s_empty + 16;              // [2] elementwise → dZ-MMA
  const uint32_t g_empty = g_full + 16;              // [2] dZ-MMA commit → elementwise
  const uint32_t d_full = g_empty + 16;              // 1  last commit → epilogue
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bar_area + 8 * (2 + 2 * STAGES + 8) + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ib = blockIdx.x;                          // row block
  const int j_begin = blockIdx.y * p.j_chunk;
  const int j_end = min(p.n, j_begin + p.j_chunk);
  const int n_tiles = (j_end - j_begin + BJ - 1) / BJ;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mI_hi); tma_prefetch_desc(&p.mI_lo); tma_prefetch_desc(&p.mJ_hi);
    tma_prefetch_desc(&p.mJ_lo); tma_prefetch_desc(&p.mT_hi); tma_prefetch_desc(&p.mT_lo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(zi_bar, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(stage_free + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full + 8 * b, 1);
      mbar_init(s_empty + 8 * b, EW_WARPS);   // one arrival per elementwise warp (512 per-thread arrivals on one
      mbar_init(g_full + 8 * b, EW_WARPS);    // mbarrier cost ~2000 cycles per tile — measured)
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
        tma_load_2d(st, &p.mJ_hi, fb, 0, j0);
        tma_load_2d(st + ZJ_BYTES, &p.mJ_lo, fb, 0, j0);
        tma_load_2d(st + 2 * ZJ_BYTES, &p.mT_hi, fb, j0, 0);
        tma_load_2d(st + 2 * ZJ_BYTES + 2048, &p.mT_hi, fb, j0 + 32, 0);
        tma_load_2d(st + 2 * ZJ_BYTES + ZT_BYTES, &p.mT_lo, fb, j0, 0);
        tma_load_2d(st + 2 * ZJ_BYTES + ZT_BYTES + 2048, &p.mT_lo, fb, j0 + 32, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(BI, BJ, 0, 0);   // S  = Z_I (K-major, K = 16) · Z_J (K-major)
      const uint32_t idesc_d = umma_idesc(BI, DW, 0, 0);   // dZ = G (TMEM, K = j) · Z_Jᵀ tile (K-major, N = 16)
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
        for (int k = 0; k < DW / 8; ++k) {
          // SWIZZLE_64B K-major: 64-byte rows, 8-row groups 512 B apart
          const uint64_t a_hi = umma_desc(s_zi_hi + k * 32, 16, 512, 4), a_lo = umma_desc(s_zi_lo + k * 32, 16, 512, 4);
          const uint64_t b_hi = umma_desc(st + k * 32, 16, 512, 4), b_lo = umma_desc(st + ZJ_BYTES + k * 32, 16, 512, 4);
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
        const uint32_t zt = s_ring + stage_d * STAGE_BYTES + 2 * ZJ_BYTES;
        const uint32_t g_hi = tmem + TM_GHI + (uint32_t)(b * BJ), g_lo = tmem + TM_GLO + (uint32_t)(b * BJ);
#pragma unroll
        for (int k = 0; k < BJ / 8; ++k) {
          // ZT tile: two boxes of [16 rows(d) x 128 B (32 j)], SWIZZLE_128B K-major; k-step = 8 j = 32 B
          const uint32_t boff = (uint32_t)(k >> 2) * 2048u + (uint32_t)(k & 3) * 32u;
          const uint64_t b_hi = umma_desc(zt + boff, 16, 1024, 2);
          const uint64_t b_lo = umma_desc(zt + ZT_BYTES + boff, 16, 1024, 2);
          // even / odd k-steps feed independent accumulator pairs (A / B) so consecutive MMAs never wait on each other
          const uint32_t dbase = tmem + TM_D + (uint32_t)((k & 1) * 32);
          const uint32_t acc = (t > 0 || k > 1) ? 1u : 0u;
          umma_tf32_ts(dbase + 16, g_lo + k * 8, b_hi, idesc_d, acc);   // small: lo·hi
          umma_tf32_ts(dbase, g_hi + k * 8, b_hi, idesc_d, acc);        // big:   hi·hi
          umma_tf32_ts(dbase + 16, g_hi + k * 8, b_lo, idesc_d, 1);     // small: hi·lo
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
    const int part = (warp - 4) >> 2;     // which EW_COLS-wide slice of the tile's 64 columns
    const int row_local = ib * BI + sub * 32 + lane;
    const bool live = row_local < p.n_rows;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    static_assert(EW_COLS == 16, "elementwise slice is written for 16 columns");
    float relu_sum = 0.f, lg_sum = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int b = t & 1;
      const uint32_t par = (t >> 1) & 1;
      mbar_wait(s_full + 8 * b, par);
      tc_fence_after();
      uint32_t v[EW_COLS];
      tmem_ld_32x32b_x16(tmem + lane_off + TM_S + (uint32_t)(b * BJ + part * EW_COLS), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty + 8 * b);       // S[b] has been copied to registers
      const int col0 = j_begin + t * BJ + part * EW_COLS;
      // staged so that the 16 independent SFU chains are issued back to back (ILP instead of one long dependent chain)
      float e[EW_COLS], sg[EW_COLS];
#pragma unroll
      for (int c = 0; c < EW_COLS; ++c) e[c] = ex2a(-fabsf(__uint_as_float(v[c])) * LOG2E);
#pragma unroll
      for (int c = 0; c < EW_COLS; ++c) sg[c] = rcpa(1.f + e[c]);
      float prod0 = 1.f, prod1 = 1.f;
      uint32_t hi[EW_COLS], lo[EW_COLS];
#pragma unroll
      for (int c = 0; c < EW_COLS; ++c) {
        const float x = __uint_as_float(v[c]);
        float s = x >= 0.f ? sg[c] : e[c] * sg[c];               // sigmoid(x)
        const bool valid = col0 + c < j_end;
        relu_sum += valid ? fmaxf(x, 0.f) : 0.f;
        const float f = valid ? 1.f + e[c] : 1.f;
        if (c & 1) prod1 *= f; else prod0 *= f;
        s = valid ? s : 0.f;
        const uint32_t h = __float_as_uint(s) & 0xFFFFE000u;
        hi[c] = h;
        lo[c] = __float_as_uint(s - __uint_as_float(h));
      }
      lg_sum += lg2a(prod0 * prod1);                             // Σ log2(1+e) over this thread's logits of the tile
      mbar_wait(g_empty + 8 * b, par ^ 1);                       // the dZ-MMA of tile t-2 has finished reading G[b]
      tc_fence_after();
      tmem_st_32x32b_x16(tmem + lane_off + TM_GHI + (uint32_t)(b * BJ + part * EW_COLS), hi);
      tmem_st_32x32b_x16(tmem + lane_off + TM_GLO + (uint32_t)(b * BJ + part * EW_COLS), lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_full + 8 * b);
    }
    double loss = live ? (double)(relu_sum + LN2 * lg_sum) : 0.0;
    loss = warp_sum(loss);
    if (lane == 0 && loss != 0.0) atomicAdd(p.loss_acc, loss * (double)p.coef);
    if (part == 0) {
      mbar_wait(d_full, 0);
      tc_fence_after();
      uint32_t a0[16], a1[16], a2[16], a3[16];
      tmem_ld_32x32b_x16(tmem + lane_off + TM_D, a0);
      tmem_ld_32x32b_x16(tmem + lane_off + TM_D + 16, a1);
      tmem_ld_32x32b_x16(tmem + lane_off + TM_D + 32, a2);
      tmem_ld_32x32b_x16(tmem + lane_off + TM_D + 48, a3);
      if (live && n_tiles > 0) {
        const float c2 = 2.f * p.coef;
        float* dst = p.dz + (size_t)row_local * p.d;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < p.d) {
            const float big = __uint_as_float(a0[c]) + __uint_as_float(a2[c]);
            const float small = __uint_as_float(a1[c]) + __uint_as_float(a3[c]);
            const float g = c2 * (big + small);
            if (p.j_splits == 1) dst[c] += g; else atomicAdd(dst + c, g);
          }
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

static int64_t padded_n(int32_t n) { return ((int64_t)n + 63) / 64 * 64; }

size_t workspace_bytes(int32_t n) {
  return 2 * align_up((size_t)n * DW * sizeof(float), 256) + 2 * align_up((size_t)padded_n(n) * DW * sizeof(float), 256);
}

bool eligible(int32_t n, int32_t d, int32_t n_rows) {
  const int mode = path_mode(B2_PATH_GAE_DECODER);           // 0 auto · 2 this kernel
  if (mode != 0 && mode != 2) return false;
  return d >= 1 && d <= DW && (int64_t)n * n_rows >= (1ll << 22);
}

// all-pairs part on the tensor cores; returns B2_ERR_UNSUPPORTED if tensor maps cannot be built
int launch(const float* z, int64_t ldz, int32_t n, int32_t d, int32_t row_begin, int32_t n_rows, float coef, float* dz,
           double* loss_acc, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < workspace_bytes(n)) return B2_ERR_UNSUPPORTED;
  const int64_t npad = padded_n(n);
  char* w = reinterpret_cast<char*>(ws);
  const size_t a16 = align_up((size_t)n * DW * sizeof(float), 256), at = align_up((size_t)npad * DW * sizeof(float), 256);
  float* z16h = reinterpret_cast<float*>(w);
  float* z16l = reinterpret_cast<float*>(w + a16);
  float* zth = reinterpret_cast<float*>(w + 2 * a16);
  float* ztl = reinterpret_cast<float*>(w + 2 * a16 + at);
  {
    int64_t blocks = ceil_div<int64_t>(npad * DW, 256 * 4);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    split_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ldz, n, d, npad, z16h, z16l, zth, ztl);
    B2_CHECK_LAUNCH("split_kernel");
  }
  Params p;
  memset(&p, 0, sizeof(p));
  const int SW64 = (int)CU_TENSOR_MAP_SWIZZLE_64B, SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  bool ok = make_tensor_map_f32_ex(&p.mI_hi, z16h, DW, (uint64_t)n, DW, DW, BI, SW64) &&
            make_tensor_map_f32_ex(&p.mI_lo, z16l, DW, (uint64_t)n, DW, DW, BI, SW64) &&
            make_tensor_map_f32_ex(&p.mJ_hi, z16h, DW, (uint64_t)n, DW, DW, BJ, SW64) &&
            make_tensor_map_f32_ex(&p.mJ_lo, z16l, DW, (uint64_t)n, DW, DW, BJ, SW64) &&
            make_tensor_map_f32_ex(&p.mT_hi, zth, (uint64_t)npad, DW, (uint64_t)npad, 32, DW, SW128) &&
            make_tensor_map_f32_ex(&p.mT_lo, ztl, (uint64_t)npad, DW, (uint64_t)npad, 32, DW, SW128);
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
