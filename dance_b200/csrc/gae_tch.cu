// fp16-split tensor-core version of the matrix-free Graph-AE decoder loss (all-pairs part) — same pipeline as gae_tc.cu
// (S = Z_I·Z_Jᵀ → σ/softplus on sixteen elementwise warps → dZ_I += G·Z_J with G read from TMEM), but every operand is
// carried as an fp16 (hi, lo) pair and the products run as tcgen05.mma kind::f16 with K = 16 per instruction:
//   x ≈ hi + lo with hi = fp16(x), lo = fp16(x − hi)  → 22 significant bits, products accumulated in fp32 in TMEM;
//   the three products hi·hi, lo·hi, hi·lo replace the tf32 three-product split at HALF the instruction count
//   (15 instead of 30 tcgen05.mma per 128×64 tile).  Round-1 profiling showed this kernel is bound by the ~60-cycle
//   issue cost of each small tcgen05.mma, not by the SFU or the tensor pipe's flops.
// Scaling: the A operand of the S product is z·log2(e), so the accumulator holds v = log2(e)·x and the exponential is a
// bare MUFU.EX2(-|v|); the B operand of the dZ product (ZT) is multiplied by a power of two 2^e chosen from max|z|
// (device-side, no host sync) so that max|z|·2^e ∈ [256, 512), and σ ∈ (0,1) by 2^11 — both keep the fp16 lo halves out
// of the subnormals and are undone exactly.  Σ_j max(x,0) = ½(Σ_j x + Σ_j |x|) with Σ_j x_ij = z_i·Σ_j z_j in closed form.
//
// Operand layouts (embedding width d ≤ 16, zero-padded to 16):
//   Z16  [n,16] halves, rows of 32 B  → K-major SWIZZLE_32B tiles: A (128 rows) and B (64 rows) of the S product, K = 16
//   ZT   [16,n'] halves, transposed   → K-major SWIZZLE_128B tile [16 rows(d) × 64 j]: B of the dZ product (N = 16, K = j)
//   G    TMEM, 16-bit elements packed two per 32-bit column (even k in the low half): A of the dZ product
#include "tc_common.cuh"

#include <cuda_fp16.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

namespace b2 {
namespace gtch {

using namespace tc;

constexpr int BI = 128;          // rows per CTA (UMMA M)
constexpr int BJ = 64;           // columns per tile (UMMA N of the S product, K of the dZ product)
constexpr int DW = 16;           // embedding width handled (d ≤ 16, zero-padded)
constexpr int STAGES = 4;
constexpr int EW_WARPS = 16;     // elementwise warps: 4 per TMEM lane quarter (= per SM sub-partition), 16 columns each
constexpr int EW_THREADS = EW_WARPS * 32;
constexpr int EW_COLS = BJ / (EW_WARPS / 4);
constexpr int THREADS = 128 + EW_THREADS;   // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. elementwise
constexpr int ZI_BYTES = BI * DW * 2;            // 4 KB  (hi or lo)
constexpr int ZJ_BYTES = BJ * DW * 2;            // 2 KB  (hi or lo), K-major rows of 32 B
constexpr int ZT_BYTES = DW * BJ * 2;            // 2 KB  (hi or lo): one box of [16 rows x 128 B]
constexpr int STAGE_BYTES = 2 * ZJ_BYTES + 2 * ZT_BYTES;   // 8 KB
constexpr int GCOLS = BJ / 2;                    // TMEM columns of one packed G tile
constexpr uint32_t TM_S = 0, TM_GHI = 128, TM_GLO = 192, TM_D = 256, TM_COLS = 512;
constexpr float G_SCALE = 2048.f;                // σ·2^11 keeps the hi/lo halves of small σ out of the fp16 subnormals
// dZ accumulators at TM_D + 16·{0: big A, 1: small A, 2: big B, 3: small B}   (packed variant: 48 columns per parity)

struct Params {
  CUtensorMap mI_hi, mI_lo;      // Z16 [n,16] halves box {16,128} SWIZZLE_32B   (A of the S product)
  CUtensorMap mJ_hi, mJ_lo;      // Z16 [n,16] halves box {16,64}  SWIZZLE_32B   (B of the S product)
  CUtensorMap mT_hi, mT_lo;      // ZT  [16,n'] halves box {64,16} SWIZZLE_128B  (B of the dZ product)
  const float* scale;            // scale[0] = 2^e applied to ZT, scale[2] = 2^-e
  const float* z; int64_t ldz;   // original embeddings (closed-form Σ_j x_ij in the epilogue)
  const double* zsum;            // [j_splits, 16] column sums of z over each CTA column range
  float* dz;                     // [n_rows, d]
  double* loss_acc;
  int n, d, row_begin, n_rows, j_chunk, j_splits;
  float coef;
};
constexpr int STAGGER_CYCLES = 800;   // initial delay of the second elementwise group (half a tile period)

// max |z| → power-of-two scale (scale[0] = 2^e, scale[1] = 2^-2e, scale[2] = 2^-e)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, uint32_t* __restrict__ maxbits) {
  float m = 0.f;
  const int64_t total = (int64_t)n * d;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(z[(t / d) * ldz + t % d]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(m));     // non-negative floats order like their bits
}

__global__ void scale_kernel(const uint32_t* __restrict__ maxbits, float* __restrict__ scale) {
  const float m = __uint_as_float(maxbits[0]);
  int e = 0;
  if (m > 0.f && isfinite(m)) { int ex; frexpf(m, &ex); e = 9 - ex; }      // m·2^e ∈ [256, 512)
  e = e > 40 ? 40 : (e < -40 ? -40 : e);
  scale[0] = ldexpf(1.f, e);
  scale[1] = ldexpf(1.f, -2 * e);
  scale[2] = ldexpf(1.f, -e);
  // the S-product operands are UNSCALED fp16 pairs of log2(e)·z: beyond 2^15 they overflow — this kernel then steps aside
  // (scale[3] = 1) and b2_gae_loss_grad_f32 runs the fp32 CUDA-core kernel instead (device-side predicate, no host sync)
  scale[3] = (m * 1.4426950408889634f >= 32768.f || !isfinite(m)) ? 1.f : 0.f;
}

// z [n,d] → fp16 hi/lo split of 2^e·z, as Z16 (row-major, padded to 16) and ZT (transposed, row pitch npad)
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, int64_t npad, const float* __restrict__ scale,
             __half* __restrict__ z16h, __half* __restrict__ z16l, __half* __restrict__ za16h, __half* __restrict__ za16l,
             __half* __restrict__ zth, __half* __restrict__ ztl) {
  const int64_t total = npad * DW;
  const float s = scale[0];
  constexpr float LOG2E = 1.4426950408889634f;
  auto split = [](float v, __half& h, __half& l) { h = __float2half_rn(v); l = __float2half_rn(v - __half2float(h)); };
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / DW;
    const int c = (int)(t % DW);
    const float v = (c < d && i < n) ? z[i * ldz + c] : 0.f;
    __half h, l;
    if (i < n) {
      split(v, h, l);          z16h[t] = h;  z16l[t] = l;      // B of the S product
      split(v * LOG2E, h, l);  za16h[t] = h; za16l[t] = l;     // A of the S product: logits come out in log2 units
    }
    split(v * s, h, l);                                         // B of the dZ product
    zth[(int64_t)c * npad + i] = h;
    ztl[(int64_t)c * npad + i] = l;
  }
}

// zsum[y, c] += Σ_{j in column range y, slice blockIdx.y} z[j, c]   (fp64; zsum zeroed by the caller)
__global__ void __launch_bounds__(256)
colrange_sum_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, int32_t j_chunk, double* __restrict__ zsum) {
  __shared__ double sh[256];
  const int y = blockIdx.x, c = threadIdx.x & 15, part = threadIdx.x >> 4;
  const int j0 = y * j_chunk, j1 = min(n, j0 + j_chunk);
  const int per = (j1 - j0 + gridDim.y - 1) / gridDim.y;
  const int a0 = j0 + blockIdx.y * per, a1 = min(j1, a0 + per);
  double a = 0.0;
  if (c < d) for (int j = a0 + part; j < a1; j += 16) a += (double)z[(int64_t)j * ldz + c];
  sh[threadIdx.x] = a;
  __syncthreads();
  if (part == 0) {
    for (int q = 1; q < 16; ++q) a += sh[q * 16 + c];
    atomicAdd(zsum + y * 16 + c, a);
  }
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(THREADS, 1)
gae_allpairs_tch_kernel(const __grid_constant__ Params p) {
  if (p.scale[3] != 0.f) return;                 // embedding too large for the fp16 operand format (see scale_kernel)
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
      mbar_init(s_empty + 8 * b, EW_WARPS / 2);   // one arrival per elementwise warp of the group that owns buffer b
      mbar_init(g_full + 8 * b, EW_WARPS / 2);
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
        tma_load_2d(st + 2 * ZJ_BYTES + ZT_BYTES, &p.mT_lo, fb, j0, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(BI, BJ, 0, 0);   // S  = Z_I (K-major, K = 16) · Z_J (K-major)
      const uint32_t idesc_d = umma_idesc_f16(BI, DW, 0, 0);   // dZ = G (TMEM, K = j) · Z_Jᵀ tile (K-major, N = 16)
      mbar_wait(zi_bar, 0);
      tc_fence_after();
      auto issue_s = [&](int t, int stage, uint32_t phase) {
        const int b = t & 1;
        mbar_wait(s_empty + 8 * b, ((t >> 1) & 1) ^ 1);
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        const uint32_t st = s_ring + stage * STAGE_BYTES;
        const uint32_t d_s = tmem + TM_S + (uint32_t)(b * BJ);
        // SWIZZLE_32B K-major: 32-byte rows (the whole K = 16), 8-row groups 256 B apart — one k-step
        const uint64_t a_hi = umma_desc(s_zi_hi, 16, 256, 6), a_lo = umma_desc(s_zi_lo, 16, 256, 6);
        const uint64_t b_hi = umma_desc(st, 16, 256, 6), b_lo = umma_desc(st + ZJ_BYTES, 16, 256, 6);
        umma_f16(d_s, a_lo, b_hi, idesc_s, 0);
        umma_f16(d_s, a_hi, b_lo, idesc_s, 1);
        umma_f16(d_s, a_hi, b_hi, idesc_s, 1);
        umma_commit(s_full + 8 * b);
      };
      int stage_s = 0, stage_d = 0;
      uint32_t phase_s = 0;
      // S runs TWO tiles ahead of dZ: the group that finishes tile t needs S(t+2) next, and the tensor pipe executes in order,
      // so S(t+2) must be queued BEFORE dZ(t) (it only needs the group to have copied S(t) out of the shared buffer, which
      // happens at the start of its work on tile t) — otherwise every group idles for a dZ + S round trip per tile.
      for (int t0 = 0; t0 < 2 && t0 < n_tiles; ++t0) { issue_s(t0, stage_s, phase_s); if (++stage_s == STAGES) { stage_s = 0; phase_s ^= 1; } }
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 2 < n_tiles) { issue_s(t + 2, stage_s, phase_s); if (++stage_s == STAGES) { stage_s = 0; phase_s ^= 1; } }
        const int b = t & 1;
        mbar_wait(g_full + 8 * b, (t >> 1) & 1);
        tc_fence_after();
        const uint32_t zt = s_ring + stage_d * STAGE_BYTES + 2 * ZJ_BYTES;
        const uint32_t g_hi = tmem + TM_GHI + (uint32_t)(b * GCOLS), g_lo = tmem + TM_GLO + (uint32_t)(b * GCOLS);
#pragma unroll
        for (int k = 0; k < BJ / 16; ++k) {
          // ZT tile: [16 rows(d) x 128 B (64 j)], SWIZZLE_128B K-major; k-step = 16 j = 32 B; G: 8 packed columns per k-step
          const uint64_t b_hi = umma_desc(zt + (uint32_t)k * 32u, 16, 1024, 2);
          const uint64_t b_lo = umma_desc(zt + ZT_BYTES + (uint32_t)k * 32u, 16, 1024, 2);
          // even / odd k-steps feed independent accumulator pairs (A / B) so consecutive MMAs never wait on each other
          const uint32_t acc = (t > 0 || k > 1) ? 1u : 0u;
          const uint32_t dbase = tmem + TM_D + (uint32_t)((k & 1) * 32);
          umma_f16_ts(dbase + 16, g_lo + k * 8, b_hi, idesc_d, acc);   // small: lo·hi
          umma_f16_ts(dbase, g_hi + k * 8, b_hi, idesc_d, acc);        // big:   hi·hi
          umma_f16_ts(dbase + 16, g_hi + k * 8, b_lo, idesc_d, 1);     // small: hi·lo
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
    static_assert(EW_COLS == 16 && EW_WARPS == 16 && BJ == 64, "elementwise slice is written for 16 columns");
    float abs_sum = 0.f, lg_sum = 0.f;
    int tiles_done = 0;                                   // 16-logit chunks processed by this thread
    // Two groups of eight warps (two per SM sub-partition) take alternate tiles: group g owns S/G buffer g.  All warps of
    // one tile run the same MUFU-burst / ALU-tail phases in lock-step, so with a single group the XU and the ALU pipes
    // idle in turn (ncu: XU 58 %, issue 58 %); two groups half a period apart keep both busy.
    const int group = part >> 1;                          // tile parity handled (= buffer index)
    const int half = part & 1;                            // 32-column half of the tile
    if (group == 1) { const long long t0 = clock64(); while (clock64() - t0 < STAGGER_CYCLES) { } }
    for (int t = group; t < n_tiles; t += 2) {
      const int b = group;
      const uint32_t par = (t >> 1) & 1;
      mbar_wait(s_full + 8 * b, par);
      tc_fence_after();
      // both 16-column chunks are fetched up front and processed in one basic block (32 independent chains): with two warps
      // per sub-partition per tile the kernel is otherwise bound by the latency of each warp's ld → MUFU → st chain
      const int cofs = half * 32;
      const int col0 = j_begin + t * BJ + cofs;
      uint32_t v0[EW_COLS], v1[EW_COLS];
      tmem_ld_32x32b_x16_nowait(tmem + lane_off + TM_S + (uint32_t)(b * BJ + cofs), v0);
      tmem_ld_32x32b_x16_nowait(tmem + lane_off + TM_S + (uint32_t)(b * BJ + cofs + EW_COLS), v1);
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(s_empty + 8 * b);         // S[b] has been copied to registers (tcgen05.wait::ld is warp-collective)
      uint32_t hi0[EW_COLS / 2], lo0[EW_COLS / 2], hi1[EW_COLS / 2], lo1[EW_COLS / 2];
      // Lean per-logit sequence (v = log2(e)·x straight out of the accumulator):
      //   e = 2^(-|v|)                     MUFU.EX2
      //   q = (1+e)/2048, r = 1/q          FFMA + MUFU.RCP      (r = 2048·σ(|x|): the G scale comes for free)
      //   g = v >= 0 ? r : e·r             FMUL + FSETP + FSEL  (2048·σ(x))
      //   Σ |v|, Π q                       FADD + FMUL          (softplus = ½(x+|x|) + ln(1+e); Σx in closed form; one LG2 per 8 logits)
      //   g → fp16 hi (mantissa mask) / lo LOP3 + FADD + ½·2 F2FP
      auto chunk_math = [&](auto full_tag, const uint32_t (&v)[EW_COLS], int c0, uint32_t (&hi)[EW_COLS / 2], uint32_t (&lo)[EW_COLS / 2]) {
        constexpr bool FULL = decltype(full_tag)::value;
        float prod0 = 1.f, prod1 = 1.f;
        float g[EW_COLS];
#pragma unroll
        for (int c = 0; c < EW_COLS; ++c) {
          const float x = __uint_as_float(v[c]);
          const float e = ex2a(-fabsf(x));
          const float q = fmaf(e, 1.f / G_SCALE, 1.f / G_SCALE);
          const float r = rcpa(q);
          const float er = e * r;
          float gc = x >= 0.f ? r : er;
          if (FULL) {
            abs_sum += fabsf(x);
            if (c & 1) prod1 *= q; else prod0 *= q;
          } else {
            const bool valid = c0 + c < j_end;
            abs_sum += valid ? fabsf(x) : 0.f;
            const float f = valid ? q : 1.f / G_SCALE;      // invalid columns carry the same 2^-11 factor as valid ones (undone below)
            if (c & 1) prod1 *= f; else prod0 *= f;
            gc = valid ? gc : 0.f;
          }
          g[c] = gc;
        }
#pragma unroll
        for (int c = 0; c < EW_COLS; c += 2) {   // two fp16 per TMEM column: element k in the low half, k+1 in the high half
          const float h0 = __uint_as_float(__float_as_uint(g[c]) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(g[c + 1]) & 0xFFFFE000u);
          const __half2 h2 = __floats2half2_rn(h0, h1);                 // exact: 10 mantissa bits survive the mask
          const __half2 l2 = __floats2half2_rn(g[c] - h0, g[c + 1] - h1);
          hi[c >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[c >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        lg_sum += lg2a(prod0) + lg2a(prod1);                            // Σ log2((1+e)/2048): 8 factors each, ≥ 2^-88
      };
      if (col0 + 2 * EW_COLS <= j_end) {
        chunk_math(std::true_type{}, v0, col0, hi0, lo0);
        chunk_math(std::true_type{}, v1, col0 + EW_COLS, hi1, lo1);
      } else {
        chunk_math(std::false_type{}, v0, col0, hi0, lo0);
        chunk_math(std::false_type{}, v1, col0 + EW_COLS, hi1, lo1);
      }
      tiles_done += 2;
      mbar_wait(g_empty + 8 * b, par ^ 1);                              // the dZ-MMA of tile t-2 has finished reading G[b]
      tc_fence_after();
      tmem_st_32x32b_x8(tmem + lane_off + TM_GHI + (uint32_t)(b * GCOLS + cofs / 2), hi0);
      tmem_st_32x32b_x8(tmem + lane_off + TM_GLO + (uint32_t)(b * GCOLS + cofs / 2), lo0);
      tmem_st_32x32b_x8(tmem + lane_off + TM_GHI + (uint32_t)(b * GCOLS + cofs / 2 + EW_COLS / 2), hi1);
      tmem_st_32x32b_x8(tmem + lane_off + TM_GLO + (uint32_t)(b * GCOLS + cofs / 2 + EW_COLS / 2), lo1);
      tmem_st_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(g_full + 8 * b);
    }
    // Σ_j softplus(x_ij) over this thread's logits = ln2·[½(Σv + Σ|v|) + Σlog2(1+e)], v in log2 units; every logit carried a
    // 2^-11 factor inside the products (11 per logit, 16 logits per tile); Σ_j v_ij = log2(e)·z_i·Σ_j z_j is added once per
    // row by the part-0 warp.
    double lin = 0.0;
    if (live && part == 0 && n_tiles > 0) {
      const float* zi = p.z + (size_t)(p.row_begin + row_local) * p.ldz;
      const double* zs = p.zsum + blockIdx.y * 16;
      double dot = 0.0;
      for (int c = 0; c < p.d; ++c) dot += (double)zi[c] * zs[c];
      lin = dot * (double)LOG2E;
    }
    double loss = live ? (double)LN2 * (0.5 * ((double)abs_sum + lin) + (double)lg_sum + 11.0 * EW_COLS * (double)tiles_done) : 0.0;
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
        const float c2 = 2.f * p.coef * p.scale[2] * (1.f / G_SCALE);   // undo the ZT and G scales
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

const float* overflow_flag(const void* ws) { return reinterpret_cast<const float*>(reinterpret_cast<const char*>(ws) + 16) + 3; }

size_t workspace_bytes(int32_t n) {
  return 256 + 32768 + 4 * align_up((size_t)n * DW * sizeof(__half), 256) + 2 * align_up((size_t)padded_n(n) * DW * sizeof(__half), 256);
}

bool eligible(int32_t n, int32_t d, int32_t n_rows) {
  const int mode = path_mode(B2_PATH_GAE_DECODER);           // 0 auto · 3 this kernel · 4 symmetric (row shards fall back to this one)
  if (mode == 1 || mode == 2) return false;
  return d >= 1 && d <= DW && (int64_t)n * n_rows >= (1ll << 22);
}

// all-pairs part on the tensor cores; returns B2_ERR_UNSUPPORTED if tensor maps cannot be built
int launch(const float* z, int64_t ldz, int32_t n, int32_t d, int32_t row_begin, int32_t n_rows, float coef, float* dz,
           double* loss_acc, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < workspace_bytes(n)) return B2_ERR_UNSUPPORTED;
  const int64_t npad = padded_n(n);
  char* w = reinterpret_cast<char*>(ws);
  uint32_t* maxbits = reinterpret_cast<uint32_t*>(w);
  float* scale = reinterpret_cast<float*>(w + 16);
  double* zsum = reinterpret_cast<double*>(w + 256);        // up to 256 column ranges × 16
  w += 256 + 32768;
  const size_t a16 = align_up((size_t)n * DW * sizeof(__half), 256), at = align_up((size_t)npad * DW * sizeof(__half), 256);
  __half* z16h = reinterpret_cast<__half*>(w);
  __half* z16l = reinterpret_cast<__half*>(w + a16);
  __half* za16h = reinterpret_cast<__half*>(w + 2 * a16);
  __half* za16l = reinterpret_cast<__half*>(w + 3 * a16);
  __half* zth = reinterpret_cast<__half*>(w + 4 * a16);
  __half* ztl = reinterpret_cast<__half*>(w + 4 * a16 + at);
  {
    B2_CHECK_CUDA(cudaMemsetAsync(maxbits, 0, 4, st));
    int64_t blocks = ceil_div<int64_t>((int64_t)n * d, 256 * 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ldz, n, d, maxbits);
    B2_CHECK_LAUNCH("absmax_kernel");
    scale_kernel<<<1, 1, 0, st>>>(maxbits, scale);
    B2_CHECK_LAUNCH("scale_kernel");
    blocks = ceil_div<int64_t>(npad * DW, 256 * 4);
    const int64_t cap2 = (int64_t)sm_count() * 16;
    if (blocks > cap2) blocks = cap2;
    split_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ldz, n, d, npad, scale, z16h, z16l, za16h, za16l, zth, ztl);
    B2_CHECK_LAUNCH("split_kernel");
  }
  Params p;
  memset(&p, 0, sizeof(p));
  const int SW32 = (int)CU_TENSOR_MAP_SWIZZLE_32B, SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  bool ok = make_tensor_map_f16_ex(&p.mI_hi, za16h, DW, (uint64_t)n, DW, DW, BI, SW32) &&
            make_tensor_map_f16_ex(&p.mI_lo, za16l, DW, (uint64_t)n, DW, DW, BI, SW32) &&
            make_tensor_map_f16_ex(&p.mJ_hi, z16h, DW, (uint64_t)n, DW, DW, BJ, SW32) &&
            make_tensor_map_f16_ex(&p.mJ_lo, z16l, DW, (uint64_t)n, DW, DW, BJ, SW32) &&
            make_tensor_map_f16_ex(&p.mT_hi, zth, (uint64_t)npad, DW, (uint64_t)npad, BJ, DW, SW128) &&
            make_tensor_map_f16_ex(&p.mT_lo, ztl, (uint64_t)npad, DW, (uint64_t)npad, BJ, DW, SW128);
  if (!ok) return B2_ERR_UNSUPPORTED;
  p.scale = scale;
  const int row_blocks = ceil_div(n_rows, BI);
  int j_splits = 1;
  const int target = sm_count();
  if (row_blocks < target) j_splits = min(ceil_div(target, row_blocks), ceil_div(n, 8 * BJ));
  if (j_splits < 1) j_splits = 1;
  int j_chunk = ceil_div(ceil_div(n, j_splits), BJ) * BJ;
  j_splits = ceil_div(n, j_chunk);
  p.dz = dz; p.loss_acc = loss_acc; p.n = n; p.d = d; p.row_begin = row_begin; p.n_rows = n_rows;
  p.j_chunk = j_chunk; p.j_splits = j_splits; p.coef = coef;
  if (j_splits > 256) return B2_ERR_UNSUPPORTED;
  B2_CHECK_CUDA(cudaMemsetAsync(zsum, 0, sizeof(double) * 16 * (size_t)j_splits, st));
  {
    int slices = ceil_div(sm_count() * 2, j_splits);
    const int max_slices = ceil_div(j_chunk, 4096);
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    colrange_sum_kernel<<<dim3(j_splits, slices), 256, 0, st>>>(z, ldz, n, d, j_chunk, zsum);
  }
  B2_CHECK_LAUNCH("colrange_sum_kernel");
  p.z = z; p.ldz = ldz; p.zsum = zsum;
  const size_t smem = 2 * ZI_BYTES + STAGES * STAGE_BYTES + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(gae_allpairs_tch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(row_blocks, j_splits);
  gae_allpairs_tch_kernel<<<grid, THREADS, smem, st>>>(p);
  B2_CHECK_LAUNCH("gae_allpairs_tch_kernel");
  return B2_OK;
}

}  // namespace gtch
}  // namespace b2
