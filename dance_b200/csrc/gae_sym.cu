// Symmetric tcgen05 decoder: the all-pairs part of the Graph-AE loss (scgnn2.py:423-426, 603-619) evaluated over UNORDERED
// block pairs.  S = Z·Zᵀ and the label-free part of the cost are symmetric, so every 128×128 logit tile (I, J), I ≠ J, is
// computed ONCE and used twice:
//     S_IJ = Z_I·Z_Jᵀ  →  G = 2^11·σ(S)  (elementwise warps)  →  dZ_I += G·Z_J   and   dZ_J += Gᵀ·Z_I
// which halves the exponentials / reciprocals / packing work that bounds gae_tch.cu (profiles/r01_ncu_gae_tch.md: 85 % of that
// kernel is the elementwise chain, the tensor pipe is 10 % busy).
//
// Schedule (cyclic, balanced): with nb row blocks of 128 and h = ⌊nb/2⌋, block I owns the pairs {I, I+o mod nb}, o = 0..h (for
// even nb the antipodal pair o = h belongs to the smaller index) — every unordered pair exactly once, every block the same
// amount of work.  One CTA owns TWO adjacent blocks (I0, I1 = I0+1) and sweeps J = I0, I0+1, …: per J-step the two tiles
// (I0,J), (I1,J) share the staged Z_J and accumulate their Gᵀ·Z_I into ONE dZ_J accumulator, which is flushed to global memory
// with vector reductions once per step (1 024 fp32 atomics per tile instead of 2 048); dZ_I0 / dZ_I1 stay in TMEM for the
// whole sweep.  A rank-sharded run gives each rank a contiguous range of super-blocks and all-reduces the [N, d] gradient.
//
// G goes through SHARED memory (fp16 hi / lo planes, canonical SWIZZLE_128B tiles of 64 columns): the same bytes are the
// K-major A operand of G·Z_J (M = i, K = j) and the MN-major A operand of Gᵀ·Z_I (M = j, K = i) — cute's Layout_K_SW128 and
// Layout_MN_SW128 atoms coincide for a [rows × 64 halves] tile with 128-byte rows.
//
// Warp roles (768 threads, 1 CTA / SM):  0 TMA · 1 MMA issue · 2 TMEM alloc · 3 idle | 4-19 elementwise (group g = tiles of
// block I_g, two warps per TMEM lane quarter per group, 64 columns each) | 20-23 dZ_J flush.  setmaxnreg moves registers from
// the service warps to the elementwise warps.
//
// Arithmetic is that of gae_tch.cu: fp16 (hi, lo) operand pairs, three-product split accumulated in fp32 in TMEM, logits in
// log2 units (A operand pre-multiplied by log2 e), Σ softplus = ½(Σx + Σ|x|) + Σ ln(1+e^-|x|) with Σ_ij x_ij = ‖Σ_i z_i‖² in
// closed form, one LG2 per 8 logits.  Off-diagonal tiles count twice in the loss.
#include "tc_common.cuh"

#include <cuda_fp16.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

namespace b2 {
namespace gsym {

using namespace tc;

constexpr int BT = 128;          // tile edge
constexpr int DW = 16;           // embedding width handled (d ≤ 16, zero-padded)
constexpr int STAGES = 3;
constexpr int EW_WARPS = 16, FL_WARPS = 4;
constexpr int THREADS = 128 + 32 * EW_WARPS + 32 * FL_WARPS;   // 768
constexpr int ZA_BYTES = BT * DW * 2;                 // 4 KB: one fp16 plane of a 128-row block, K-major rows of 32 B (SWIZZLE_32B)
constexpr int ZT_BOX = DW * 64 * 2;                   // 2 KB: [16 rows(d) × 64 j] fp16, SWIZZLE_128B
constexpr int ZT_BYTES = 4 * ZT_BOX;                  // 8 KB: two 64-column blocks × (hi | lo) — hi and lo adjacent = one N = 32 operand
constexpr int ZI_BYTES = 2 * ZA_BYTES + ZT_BYTES;     // 16 KB per owned block: A of S (hi, lo) + B of Gᵀ·Z_I
constexpr int STAGE_BYTES = 2 * ZA_BYTES + ZT_BYTES;  // 16 KB per J-step: B of S (hi, lo) + B of G·Z_J
constexpr int G_PLANE = BT * BT * 2;                  // 32 KB: fp16 plane of one G tile (two 16 KB blocks of 64 columns)
constexpr int G_BYTES = 2 * G_PLANE;                  // hi + lo
constexpr int SMEM_BYTES = 2 * ZI_BYTES + STAGES * STAGE_BYTES + 2 * G_BYTES;   // 208 KB
constexpr uint32_t TM_S = 0, TM_D1 = 256, TM_D2 = 384, TM_COLS = 512;   // S: 2×128, dZ_I: 2 segments × 2 blocks × 32, dZ_J: 3×32
// The tensor core adds into its fp32 accumulator with truncation, so a chain of k accumulations drifts by up to k·2^-24 (measured:
// 6e-4 after the 125 000 accumulations of a 1 M-cell sweep).  dZ_I is therefore accumulated in SEGMENTS of SEG_STEPS J-steps
// (≤ 128 accumulations), double-buffered in TMEM; the flush warps add each finished segment to global memory (round-to-nearest
// atomics) while the next one accumulates.  dZ_J is flushed every step anyway.
constexpr int SEG_STEPS = 8;
constexpr float G_SCALE = 2048.f;
constexpr int STAGGER_CYCLES = 1500;

struct Params {
  CUtensorMap mA_hi, mA_lo;      // log2(e)·z [npad,16] halves, box {16,128} SWIZZLE_32B   (A of S)
  CUtensorMap mB_hi, mB_lo;      // z         [npad,16] halves, box {16,128} SWIZZLE_32B   (B of S)
  CUtensorMap mT_hi, mT_lo;      // 2^e·zᵀ    [16,npad] halves, box {64,16}  SWIZZLE_128B  (B of G·Z_J and of Gᵀ·Z_I)
  const float* scale;            // [0] 2^e (ZT), [2] 2^-e, [3] 2^2f, [4] 2^-f, [5] f > 0 (see scale_kernel)
  float* dz;                     // [n, d], zero-initialised by the caller; every contribution is an atomic add
  double* loss_acc;
  int n, d, nb, sb_begin;
  float coef;
  int stagger, late_gempty;      // b2_set_tuning knobs
  int splits;                    // CTAs per super-block (step ranges)
#ifdef B2_GAE_TRACE
  unsigned long long* trace;     // lab build only (scripts/lab/build_trace.sh): clock64 stamps of CTA 0's roles
  int trace_tiles;
#endif
};

#ifdef B2_GAE_TRACE
// agent: 0/1 elementwise group, 2 S issue, 3 D issue, 4 dZ_J flush;  8 stamps per (agent, index)
#define B2_TRACE(agent, idx, field)                                                                          \
  do {                                                                                                       \
    if (p.trace && blockIdx.x == 0 && (idx) < p.trace_tiles) p.trace[((agent) * p.trace_tiles + (idx)) * 8 + (field)] = clock64(); \
  } while (0)
unsigned long long* g_trace = nullptr;
int g_trace_tiles = 0;
#else
#define B2_TRACE(agent, idx, field) do { } while (0)
#endif

// ---- sweep bookkeeping shared by all roles ----------------------------------------------------------------------------------
struct Sweep {
  int nb, h, I0, I1, n_steps;
  int s0, s1;                                  // this CTA's share [s0, s1) of the super-block's J-steps (part `part` of `splits`)
  bool even;
  __device__ Sweep(int nb_, int sb, int part, int splits)
      : nb(nb_), h(nb_ / 2), I0(2 * sb), I1(2 * sb + 1 < nb_ ? 2 * sb + 1 : -1), even((nb_ & 1) == 0) {
    n_steps = 0;
    for (int s = h + 1; s >= 0; --s) if (active(0, s) || active(1, s)) { n_steps = s + 1; break; }   // dead steps form a suffix
    s0 = (int)((long long)n_steps * part / splits);
    s1 = (int)((long long)n_steps * (part + 1) / splits);
  }
  __device__ int J(int s) const { return (I0 + s) % nb; }
  __device__ int block(int g) const { return g ? I1 : I0; }
  __device__ bool active(int g, int s) const {
    const int I = g ? I1 : I0;
    if (I < 0) return false;
    const int o = s - g;                       // offset of J from this block
    if (o < 0 || o > h) return false;
    if (o == h && o > 0 && even && I + h >= nb) return false;   // antipodal pair belongs to the smaller index
    return true;
  }
  __device__ bool diag(int g, int s) const { return s == g; }
  __device__ bool has_d2(int s) const { return (active(0, s) && !diag(0, s)) || (active(1, s) && !diag(1, s)); }
  __device__ bool last_of_step(int g, int s) const { return g == 1 || !active(1, s); }
};

__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, uint32_t* __restrict__ maxbits) {
  float m = 0.f;
  const int64_t total = (int64_t)n * d;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(z[(t / d) * ldz + t % d]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(m));
}

__global__ void scale_kernel(const uint32_t* __restrict__ maxbits, float* __restrict__ scale) {
  const float m = __uint_as_float(maxbits[0]);
  int e = 0;
  if (m > 0.f && isfinite(m)) { int ex; frexpf(m, &ex); e = 9 - ex; }      // m·2^e ∈ [256, 512)
  e = e > 40 ? 40 : (e < -40 ? -40 : e);
  scale[0] = ldexpf(1.f, e);
  scale[1] = ldexpf(1.f, -2 * e);
  scale[2] = ldexpf(1.f, -e);
  // Operands of the S product are fp16 pairs: |log2(e)·z| must stay below 2^15.  Larger embeddings (an untrained Graph-AE at 1 M cells
  // draws z = mu + eps·exp(logvar) with logvar ≈ 14) are scaled down by 2^-f on BOTH sides; the accumulator then holds 2^-2f·v and
  // the elementwise warps multiply it back (SCALED kernel variant, one extra FMUL per logit; the unscaled variant runs otherwise).
  int f = 0;
  if (m > 0.f && isfinite(m)) { int ex; frexpf(m * 1.4426950408889634f, &ex); f = ex > 15 ? ex - 15 : 0; }
  f = f > 60 ? 60 : f;
  scale[3] = ldexpf(1.f, 2 * f);
  scale[4] = ldexpf(1.f, -f);
  scale[5] = f > 0 ? 1.f : 0.f;
}

// z [n,d] → fp16 hi/lo planes: za = log2(e)·z, zb = z (row-major, 16 wide, npad rows), zt = 2^e·zᵀ (row pitch npad); zero padding
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, int64_t npad, const float* __restrict__ scale,
             __half* __restrict__ zah, __half* __restrict__ zal, __half* __restrict__ zbh, __half* __restrict__ zbl,
             __half* __restrict__ zth, __half* __restrict__ ztl) {
  const int64_t total = npad * DW;
  const float s = scale[0], sd = scale[4];      // sd = 2^-f (1 unless the embedding is too large for fp16 operands)
  constexpr float LOG2E = 1.4426950408889634f;
  auto split = [](float v, __half& h, __half& l) { h = __float2half_rn(v); l = __float2half_rn(v - __half2float(h)); };
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / DW;
    const int c = (int)(t % DW);
    const float v = (c < d && i < n) ? z[i * ldz + c] : 0.f;
    __half h, l;
    split(v * sd, h, l);          zbh[t] = h; zbl[t] = l;
    split(v * LOG2E * sd, h, l);  zah[t] = h; zal[t] = l;
    split(v * s, h, l);
    zth[(int64_t)c * npad + i] = h;
    ztl[(int64_t)c * npad + i] = l;
  }
}

// zsum[c] += Σ_j z[j, c]  (fp64; zeroed by the caller)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ z, int64_t ldz, int32_t n, int32_t d, double* __restrict__ zsum) {
  __shared__ double sh[256];
  const int c = threadIdx.x & 15, part = threadIdx.x >> 4;
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int a0 = blockIdx.x * per, a1 = min(n, a0 + per);
  double a = 0.0;
  if (c < d) for (int j = a0 + part; j < a1; j += 16) a += (double)z[(int64_t)j * ldz + c];
  sh[threadIdx.x] = a;
  __syncthreads();
  if (part == 0) {
    for (int q = 1; q < 16; ++q) a += sh[q * 16 + c];
    atomicAdd(zsum + c, a);
  }
}

// Σ_ij x_ij = ‖Σ_i z_i‖²:  loss += coef·½·that   (the linear half of Σ max(x, 0) = ½(Σx + Σ|x|))
__global__ void linear_term_kernel(const double* __restrict__ zsum, int d, float coef, double* __restrict__ loss_acc) {
  double s = 0.0;
  for (int c = 0; c < d; ++c) s += zsum[c] * zsum[c];
  atomicAdd(loss_acc, 0.5 * s * (double)coef);
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <bool SCALED>
__global__ void __launch_bounds__(THREADS, 1)
gae_sym_kernel(const __grid_constant__ Params p) {
  // both variants are launched back to back; the one that does not match the device-side scale flag returns at once (the host never
  // reads the embedding's magnitude, so there is no synchronisation)
  if ((p.scale[5] != 0.f) != SCALED) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t s_g = smem_u32(smem);                               // G planes first: 1024-byte aligned swizzle atoms
  const uint32_t s_zi = s_g + 2 * G_BYTES;
  const uint32_t s_ring = s_zi + 2 * ZI_BYTES;
  uint8_t* bar_area = smem + SMEM_BYTES;
  const uint32_t bars = smem_u32(bar_area);
  const uint32_t zi_bar = bars;                          // 1
  const uint32_t full_bar = bars + 8;                    // [STAGES] TMA → MMA
  const uint32_t stage_free = full_bar + 8 * STAGES;     // [STAGES] last D-MMA of the step → TMA
  const uint32_t s_full = stage_free + 8 * STAGES;       // [2] S-MMA commit → elementwise group
  const uint32_t s_empty = s_full + 16;                  // [2]
  const uint32_t g_full = s_empty + 16;                  // [2] elementwise group → D-MMAs
  const uint32_t g_empty = g_full + 16;                  // [2] D-MMA commit → elementwise group
  const uint32_t d2_full = g_empty + 16;                 // [3] D-MMA commit → flush warps
  const uint32_t d2_empty = d2_full + 24;                // [3] flush warps → MMA
  const uint32_t d1_full = d2_empty + 24;                // [2] last D-MMA of a segment → flush warps
  const uint32_t d1_empty = d1_full + 16;                // [2] flush warps → MMA
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bar_area + 224);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a super-block's sweep can be cut into `splits` step ranges (one CTA each) so that the grid fills whole waves of 148 SMs under
  // sharding; every accumulator is flushed with atomic adds, so the parts are independent
  const Sweep sw(p.nb, p.sb_begin + (int)blockIdx.x / p.splits, (int)blockIdx.x % p.splits, p.splits);
  const int n_last = p.n - (p.nb - 1) * BT;              // valid rows of the last block (1..128)
  const bool ragged = n_last < BT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.mA_hi); tma_prefetch_desc(&p.mA_lo); tma_prefetch_desc(&p.mB_hi);
    tma_prefetch_desc(&p.mB_lo); tma_prefetch_desc(&p.mT_hi); tma_prefetch_desc(&p.mT_lo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(zi_bar, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(stage_free + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full + 8 * b, 1);
      mbar_init(s_empty + 8 * b, EW_WARPS / 2);
      mbar_init(g_full + 8 * b, EW_WARPS / 2);
      mbar_init(g_empty + 8 * b, 1);
    }
    for (int b = 0; b < 3; ++b) { mbar_init(d2_full + 8 * b, 1); mbar_init(d2_empty + 8 * b, FL_WARPS); }
    for (int b = 0; b < 2; ++b) { mbar_init(d1_full + 8 * b, 1); mbar_init(d1_empty + 8 * b, FL_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<48>();   // the register pool is the CTA's LAUNCH allocation (80 x 768): 128x48 + 512x96 + 128x48 = 61440 exactly
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      const int blocks_owned = sw.I1 >= 0 ? 2 : 1;
      mbar_expect_tx(zi_bar, blocks_owned * ZI_BYTES);
      for (int g = 0; g < blocks_owned; ++g) {
        const uint32_t zi = s_zi + g * ZI_BYTES;
        const int r0 = sw.block(g) * BT;
        tma_load_2d(zi, &p.mA_hi, zi_bar, 0, r0);
        tma_load_2d(zi + ZA_BYTES, &p.mA_lo, zi_bar, 0, r0);
        for (int jb = 0; jb < 2; ++jb) {
          tma_load_2d(zi + 2 * ZA_BYTES + jb * 2 * ZT_BOX, &p.mT_hi, zi_bar, r0 + jb * 64, 0);
          tma_load_2d(zi + 2 * ZA_BYTES + jb * 2 * ZT_BOX + ZT_BOX, &p.mT_lo, zi_bar, r0 + jb * 64, 0);
        }
      }
      for (int s = sw.s0; s < sw.s1; ++s) {
        const int ls = s - sw.s0, stage = ls % STAGES;
        mbar_wait(stage_free + 8 * stage, ((ls / STAGES) & 1) ^ 1);
        const uint32_t fb = full_bar + 8 * stage, st = s_ring + stage * STAGE_BYTES;
        const int c0 = sw.J(s) * BT;
        mbar_expect_tx(fb, STAGE_BYTES);
        tma_load_2d(st, &p.mB_hi, fb, 0, c0);
        tma_load_2d(st + ZA_BYTES, &p.mB_lo, fb, 0, c0);
        for (int jb = 0; jb < 2; ++jb) {
          tma_load_2d(st + 2 * ZA_BYTES + jb * 2 * ZT_BOX, &p.mT_hi, fb, c0 + jb * 64, 0);
          tma_load_2d(st + 2 * ZA_BYTES + jb * 2 * ZT_BOX + ZT_BOX, &p.mT_lo, fb, c0 + jb * 64, 0);
        }
      }
    } else if (warp == 1 || warp == 3) {
      // ===================== MMA issuers: warp 1 issues the S products, warp 3 the gradient products =====================
      // Each WHOLE warp runs its role (waits and bookkeeping are warp-uniform); tcgen05.mma / commit are issued by one elected lane
      // inside the asm.  History (clock64 traces, profiles/r02_sym_trace.md): one thread issuing everything from a divergent
      // `lane == 0` branch spent 2 500 cycles ISSUING the 32 gradient MMAs of a tile (~14 SASS instructions per MMA: an ELECT /
      // BRA.U.ANY loop plus descriptor arithmetic), the elementwise warps waited 48 % of the time; warp-uniform issue: 1 850 cycles —
      // the issuer shares its scheduler with four busy elementwise warps and gets roughly every fifth issue slot — and it was still
      // the busiest role (S + gradient products + bookkeeping = the whole tile period).  The two queues are independent (different
      // TMEM regions and shared-memory buffers, every hand-over is an mbarrier), so they run on two warps of two schedulers.
      mbar_wait(zi_bar, 0);
      tc_fence_after();
      auto next_tile = [&](int& s, int& g) {              // advance to the next active tile (s == n_steps: end)
        do { if (++g == 2) { g = 0; ++s; } } while (s < sw.s1 && !sw.active(g, s));
      };
      // A descriptor's start-address field counts 16-byte units in its low 14 bits and shared memory is < 256 KB, so a byte offset is
      // added to a base descriptor as (offset >> 4) without touching the other fields.
      auto adv = [](uint64_t desc, uint32_t byte_off) { return desc + (uint64_t)(byte_off >> 4); };
      // g = which owned block the tile belongs to (operands, dZ_I accumulator); q = parity of the tile in the CTA's tile sequence =
      // elementwise group / S buffer / G buffer (alternating by SEQUENCE, not by block: consecutive tiles never share a group).
      if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_f16(BT, BT, 0, 0);        // S = Z_I (K-major, K = 16) · Z_J (K-major)
        uint32_t cnt_s = 0;                                           // bit q = phase of S buffer q
        int s = sw.s0, g = -1, k = 0;
        for (next_tile(s, g); s < sw.s1; next_tile(s, g), ++k) {
          const int ls = s - sw.s0;
          const int q = k & 1, stage = ls % STAGES;
          if (lane == 0) B2_TRACE(2, k, 0);
          mbar_wait(s_empty + 8 * q, ((cnt_s >> q) & 1u) ^ 1u);       // the group has pulled S(k-2) into registers
          if (lane == 0) B2_TRACE(2, k, 1);
          mbar_wait(full_bar + 8 * stage, (ls / STAGES) & 1);
          tc_fence_after();
          if (lane == 0) B2_TRACE(2, k, 2);
          const uint32_t st = s_ring + stage * STAGE_BYTES, zi = s_zi + g * ZI_BYTES;
          const uint32_t d_s = tmem + TM_S + (uint32_t)(q * BT);
          if (elect_one_pred()) {
            // SWIZZLE_32B K-major: 32-byte rows (the whole K = 16), 8-row groups 256 B apart — one k-step
            constexpr uint32_t HI = umma_desc_hi(256, 6);
            const uint32_t a_hi = umma_desc_lo(zi, 16), a_lo = a_hi + (ZA_BYTES >> 4);
            const uint32_t b_hi = umma_desc_lo(st, 16), b_lo = b_hi + (ZA_BYTES >> 4);
            umma_f16_lo<HI, HI>(d_s, a_lo, b_hi, idesc_s, 0);
            umma_f16_lo<HI, HI>(d_s, a_hi, b_lo, idesc_s, 1);
            umma_f16_lo<HI, HI>(d_s, a_hi, b_hi, idesc_s, 1);
            umma_commit(s_full + 8 * q);
          }
          __syncwarp();
          if (lane == 0) B2_TRACE(2, k, 3);
          cnt_s ^= 1u << q;
        }
      } else {
        const uint32_t idesc_d1 = umma_idesc_f16(BT, 2 * DW, 0, 0);   // dZ_I = G  (K-major A,  K = j) · [Z_hi | Z_lo]_J
        const uint32_t idesc_d2 = umma_idesc_f16(BT, 2 * DW, 1, 0);   // dZ_J = Gᵀ (MN-major A, K = i) · [Z_hi | Z_lo]_I
        // per-buffer phase counters packed into scalars (dynamic indexing of local arrays would put them on the stack)
        uint32_t cnt_d = 0, use_d2 = 0, use_d1 = 0, d1_fresh = 0;     // bit q / b3 / segment parity = phase (or flag) of that slot
        int cur_seg = -1;
        int s = sw.s0, g = -1, k = 0;
        for (next_tile(s, g); s < sw.s1; next_tile(s, g), ++k) {
          const int ls = s - sw.s0;
          const int q = k & 1, stage = ls % STAGES;
          if (lane == 0) B2_TRACE(3, k, 0);
          mbar_wait(g_full + 8 * q, (cnt_d >> q) & 1u);               // G(k) written
          mbar_wait(full_bar + 8 * stage, (ls / STAGES) & 1);         // (complete long ago: the TMA writes of Z_J made visible to this thread)
          tc_fence_after();
          if (lane == 0) B2_TRACE(3, k, 1);
          const uint32_t zt_j = s_ring + stage * STAGE_BYTES + 2 * ZA_BYTES;
          const int seg = ls / SEG_STEPS, sp = seg & 1;
          if (seg != cur_seg) {                    // first dZ_I product of a new segment: its TMEM buffers must have been drained
            mbar_wait(d1_empty + 8 * sp, ((use_d1 >> sp) & 1u) ^ 1u);
            tc_fence_after();
            cur_seg = seg;
            d1_fresh = 3u;                         // both blocks start the segment with accumulate = 0
          }
          const uint32_t d1 = tmem + TM_D1 + (uint32_t)((sp * 2 + g) * 2 * DW);
          const bool off_diag = !sw.diag(g, s);
          const int b3 = ls % 3;
          const bool first = (g == 0) || !(sw.active(0, s) && !sw.diag(0, s));     // first tile of this step that feeds dZ_J
          if (off_diag && first) {
            mbar_wait(d2_empty + 8 * b3, ((use_d2 >> b3) & 1u) ^ 1u);
            tc_fence_after();
          }
          const bool last = sw.last_of_step(g, s);
          const bool seg_done = (ls % SEG_STEPS == SEG_STEPS - 1 || s == sw.s1 - 1);
          if (elect_one_pred()) {
            constexpr uint32_t HI = umma_desc_hi(1024, 2);           // SWIZZLE_128B, 8-row groups 1 KB apart (all four operand views)
            // K-major SWIZZLE_128B view of G: 64 j per 128-byte row; k-step = 32 B inside the row, 64-column blocks 16 KB apart.
            // [Z_hi | Z_lo]_J (N = 32): same layout, 64-column blocks 2·ZT_BOX apart.
            const uint32_t ga_hi = umma_desc_lo(s_g + q * G_BYTES, 16), ga_lo = ga_hi + (G_PLANE >> 4);
            const uint32_t zb_j = umma_desc_lo(zt_j, 16);
            const uint32_t acc0 = ((d1_fresh >> g) & 1u) ? 0u : 1u;
#pragma unroll
            for (int ks = 0; ks < BT / 16; ++ks) {
              const uint32_t koff = ((uint32_t)(ks >> 2) * (BT * 128) + (uint32_t)(ks & 3) * 32u) >> 4;
              const uint32_t bd = zb_j + (((uint32_t)(ks >> 2) * (2 * ZT_BOX) + (uint32_t)(ks & 3) * 32u) >> 4);
              umma_f16_lo<HI, HI>(d1, ga_hi + koff, bd, idesc_d1, ks > 0 ? 1u : acc0);
              umma_f16_lo<HI, HI>(d1, ga_lo + koff, bd, idesc_d1, 1);
            }
            if (off_diag) {
              const uint32_t d2 = tmem + TM_D2 + (uint32_t)(b3 * 2 * DW);
              // MN-major SWIZZLE_128B over the SAME bytes of G: 64 j (M) contiguous per 128-byte row, k-step = 16 rows (i) = 2 KB,
              // the second 64-j block (LBO) 16 KB further
              const uint32_t gt_hi = umma_desc_lo(s_g + q * G_BYTES, BT * 128), gt_lo = gt_hi + (G_PLANE >> 4);
              const uint32_t zb_i = umma_desc_lo(s_zi + g * ZI_BYTES + 2 * ZA_BYTES, 16);
              const uint32_t acc2 = first ? 0u : 1u;
#pragma unroll
              for (int ks = 0; ks < BT / 16; ++ks) {
                const uint32_t bd = zb_i + (((uint32_t)(ks >> 2) * (2 * ZT_BOX) + (uint32_t)(ks & 3) * 32u) >> 4);
                umma_f16_lo<HI, HI>(d2, gt_hi + (((uint32_t)ks * 2048u) >> 4), bd, idesc_d2, ks > 0 ? 1u : acc2);
                umma_f16_lo<HI, HI>(d2, gt_lo + (((uint32_t)ks * 2048u) >> 4), bd, idesc_d2, 1);
              }
            }
            umma_commit(g_empty + 8 * q);
            if (last) {
              // the S products of this step finished before its elementwise passes started, so the gradient products are the last readers
              umma_commit(stage_free + 8 * stage);
              if (sw.has_d2(s)) umma_commit(d2_full + 8 * b3);
              if (seg_done) umma_commit(d1_full + 8 * sp);
            }
          }
          __syncwarp();
          if (lane == 0) B2_TRACE(3, k, 2);
          d1_fresh &= ~(1u << g);
          cnt_d ^= 1u << q;
          if (last) {
            if (sw.has_d2(s)) use_d2 ^= 1u << b3;
            if (seg_done) use_d1 ^= 1u << sp;
          }
          if (lane == 0) B2_TRACE(3, k, 3);
        }
      }
    }
    __syncwarp();
  } else if (warp < 4 + EW_WARPS) {
    // ===================== elementwise warps =====================
    setmaxnreg_inc<96>();
    const int sub = warp & 3;                     // TMEM lane quarter
    const int part = (warp - 4) >> 2;             // 0..3
    const int q = part >> 1;                      // elementwise group = parity of the tile in the CTA's tile sequence
    const int half = part & 1;                    // 64-column half of the tile
    const int row = sub * 32 + lane;              // row inside the tile
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    constexpr float LN2 = 0.6931471805599453f;
    const float vs = SCALED ? p.scale[3] : 1.f;   // 2^2f: undoes the operand scaling of the S product
    float abs_w = 0.f, lg_w = 0.f;
    int chunks_w = 0;                             // 16-logit chunks processed, weighted like the sums
    if (q == 1) { const long long t0 = clock64(); while (clock64() - t0 < p.stagger) { } }
    int ng = 0;                                   // tiles this group has processed (phase of its S / G buffers)
    const uint32_t g_hi = s_g + q * G_BYTES + (uint32_t)half * (BT * 128) + (uint32_t)row * 128u, g_lo = g_hi + G_PLANE;
    const uint32_t xr = (uint32_t)(row & 7);
    {
      int seq = -1;                               // index of the tile in the CTA's tile sequence (same enumeration as the MMA issuer)
      for (int s = sw.s0; s < sw.s1; ++s) {
       for (int g = 0; g < 2; ++g) {
        if (!sw.active(g, s)) continue;
        ++seq;
        if ((seq & 1) != q) continue;
        const int I = sw.block(g);
        const int J = sw.J(s);
        const int wgt = sw.diag(g, s) ? 1 : 2;
        const bool masked = ragged && (I == p.nb - 1 || J == p.nb - 1);
        const bool row_ok = !(ragged && I == p.nb - 1 && row >= n_last);
        const int col_end = (ragged && J == p.nb - 1) ? n_last : BT;       // valid columns of this tile
        const bool tr = (sub == 0 && half == 0 && lane == 0);
        if (tr) B2_TRACE(q, ng, 0);
        mbar_wait(s_full + 8 * q, ng & 1);
        tc_fence_after();
        if (tr) B2_TRACE(q, ng, 1);
        if (!p.late_gempty) mbar_wait(g_empty + 8 * q, (ng & 1) ^ 1);
        if (tr) B2_TRACE(q, ng, 2);
        float abs_t = 0.f, lg_t = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cofs = half * 64 + hh * 32;
          uint32_t v0[16], v1[16];
          tmem_ld_32x32b_x16_nowait(tmem + lane_off + TM_S + (uint32_t)(q * BT + cofs), v0);
          tmem_ld_32x32b_x16_nowait(tmem + lane_off + TM_S + (uint32_t)(q * BT + cofs + 16), v1);
          tmem_ld_wait();
          if (hh == 1) {
            tc_fence_before();
            if (lane == 0) mbar_arrive(s_empty + 8 * q);                   // S[q] is in registers
          }
          uint32_t hi0[8], lo0[8], hi1[8], lo1[8];
          // Two variants of this math were measured and dropped (both correct): (a) the fp16 planes by bit extraction from a 2^31-scaled σ
          // — no F2FP, 4 fewer XU cycles per warp-logit — was 11 % SLOWER: the shifts, permutes and clamps cost more issue slots than
          // the conversions; (b) the four modifier-free FMA-pipe operations per logit as packed FFMA2 / FMUL2 pairs was 5 % slower
          // (packed fp32 only issues to one half of the FMA pipe).
          auto chunk_math = [&](auto full_tag, const uint32_t (&v)[16], int c0, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
            constexpr bool FULL = decltype(full_tag)::value;
            float prod0 = 1.f, prod1 = 1.f;
            float gg[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float x = SCALED ? __uint_as_float(v[c]) * vs : __uint_as_float(v[c]);
              const float e = ex2a(-fabsf(x));
              const float q = fmaf(e, 1.f / G_SCALE, 1.f / G_SCALE);
              const float r = rcpa(q);
              const float er = e * r;
              float gc = x >= 0.f ? r : er;
              if (FULL) {
                abs_t += fabsf(x);
                if (c & 1) prod1 *= q; else prod0 *= q;
              } else {
                const bool valid = row_ok && (c0 + c < col_end);
                abs_t += valid ? fabsf(x) : 0.f;
                const float f = valid ? q : 1.f / G_SCALE;
                if (c & 1) prod1 *= f; else prod0 *= f;
                gc = valid ? gc : 0.f;
              }
              gg[c] = gc;
            }
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              const float h0 = __uint_as_float(__float_as_uint(gg[c]) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(gg[c + 1]) & 0xFFFFE000u);
              const __half2 h2 = __floats2half2_rn(h0, h1);
              const __half2 l2 = __floats2half2_rn(gg[c] - h0, gg[c + 1] - h1);
              hi[c >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
              lo[c >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            lg_t += lg2a(prod0) + lg2a(prod1);
          };
          if (!masked) {
            chunk_math(std::true_type{}, v0, cofs, hi0, lo0);
            chunk_math(std::true_type{}, v1, cofs + 16, hi1, lo1);
          } else {
            chunk_math(std::false_type{}, v0, cofs, hi0, lo0);
            chunk_math(std::false_type{}, v1, cofs + 16, hi1, lo1);
          }
          // G[q] may be overwritten once the D-MMAs of this group's previous tile have read it: waited for as late as possible
          // (after the first half's math), so the group is not idle while the tensor pipe drains the previous tile
          if (hh == 0 && p.late_gempty) mbar_wait(g_empty + 8 * q, (ng & 1) ^ 1);
          if (tr) B2_TRACE(q, ng, 3 + hh);
          // 16-byte chunk c of this thread's 128-byte row holds columns 8c..8c+7; swizzle: chunk ^= row % 8
          const uint32_t cb = (uint32_t)hh * 4u;
          sts_v4(g_hi + (((cb + 0) ^ xr) << 4), hi0[0], hi0[1], hi0[2], hi0[3]);
          sts_v4(g_hi + (((cb + 1) ^ xr) << 4), hi0[4], hi0[5], hi0[6], hi0[7]);
          sts_v4(g_hi + (((cb + 2) ^ xr) << 4), hi1[0], hi1[1], hi1[2], hi1[3]);
          sts_v4(g_hi + (((cb + 3) ^ xr) << 4), hi1[4], hi1[5], hi1[6], hi1[7]);
          sts_v4(g_lo + (((cb + 0) ^ xr) << 4), lo0[0], lo0[1], lo0[2], lo0[3]);
          sts_v4(g_lo + (((cb + 1) ^ xr) << 4), lo0[4], lo0[5], lo0[6], lo0[7]);
          sts_v4(g_lo + (((cb + 2) ^ xr) << 4), lo1[0], lo1[1], lo1[2], lo1[3]);
          sts_v4(g_lo + (((cb + 3) ^ xr) << 4), lo1[4], lo1[5], lo1[6], lo1[7]);
        }
        fence_proxy_async();                                               // generic-proxy stores → visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(g_full + 8 * q);
        if (tr) B2_TRACE(q, ng, 5);
        abs_w += (float)wgt * abs_t;
        lg_w += (float)wgt * lg_t;
        chunks_w += wgt * 4;
        ++ng;
       }
      }
    }
    // Σ softplus over this thread's logits (both orientations of off-diagonal tiles) = ln2·[½Σ|v| + Σlog2(1+e)], the ½Σv half is
    // added in closed form by linear_term_kernel; every logit carried a 2^-11 factor inside the products
    double loss = (double)LN2 * (0.5 * (double)abs_w + (double)lg_w + 11.0 * 16.0 * (double)chunks_w);
    loss = warp_sum(loss);
    if (lane == 0 && loss != 0.0) atomicAdd(p.loss_acc, loss * (double)p.coef);
  } else {
    // ===================== dZ_J flush warps =====================
    setmaxnreg_dec<48>();
    const int sub = warp & 3;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    uint32_t fcnt = 0;                            // bit b3 = phase parity of dZ_J buffer b3
    const float c2 = 2.f * p.coef * p.scale[2] * (1.f / G_SCALE);
    uint32_t f1cnt = 0;                           // bit = phase parity of dZ_I segment buffer
    auto add_rows = [&](int block, const uint32_t (&a0)[16], const uint32_t (&a1)[16]) {
      const int gr = block * BT + sub * 32 + lane;
      if (gr < p.n) {
        float* dst = p.dz + (size_t)gr * p.d;
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = c2 * (__uint_as_float(a0[c]) + __uint_as_float(a1[c]));
        if (p.d == 16) {
#pragma unroll
          for (int c = 0; c < 16; c += 4) red_add_v4(dst + c, o[c], o[c + 1], o[c + 2], o[c + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) if (c < p.d) atomicAdd(dst + c, o[c]);
        }
      }
    };
    for (int s = sw.s0; s < sw.s1; ++s) {
      const int ls = s - sw.s0;
      if (sw.has_d2(s)) {
        const int b3 = ls % 3;
        if (sub == 0 && lane == 0) B2_TRACE(4, ls, 0);
        mbar_wait(d2_full + 8 * b3, (fcnt >> b3) & 1u);
        tc_fence_after();
        if (sub == 0 && lane == 0) B2_TRACE(4, ls, 1);
        uint32_t a0[16], a1[16];
        tmem_ld_32x32b_x16(tmem + lane_off + TM_D2 + (uint32_t)(b3 * 2 * DW), a0);
        tmem_ld_32x32b_x16(tmem + lane_off + TM_D2 + (uint32_t)(b3 * 2 * DW + DW), a1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty + 8 * b3);
        fcnt ^= 1u << b3;
        add_rows(sw.J(s), a0, a1);
        if (sub == 0 && lane == 0) B2_TRACE(4, ls, 2);
      }
      if (ls % SEG_STEPS == SEG_STEPS - 1 || s == sw.s1 - 1) {
        // a dZ_I segment is complete: [G·Z_hi | G·Z_lo] of both owned blocks → global (atomic: other CTAs add to the same rows)
        const int seg = ls / SEG_STEPS, sp = seg & 1;
        mbar_wait(d1_full + 8 * sp, (f1cnt >> sp) & 1u);
        tc_fence_after();
        bool wrote0 = false, wrote1 = false;
        for (int s2 = sw.s0 + seg * SEG_STEPS; s2 <= s; ++s2) { wrote0 |= sw.active(0, s2); wrote1 |= sw.active(1, s2); }
        {
          uint32_t a0[16], a1[16];
          if (wrote0) {
            tmem_ld_32x32b_x16(tmem + lane_off + TM_D1 + (uint32_t)((sp * 2 + 0) * 2 * DW), a0);
            tmem_ld_32x32b_x16(tmem + lane_off + TM_D1 + (uint32_t)((sp * 2 + 0) * 2 * DW + DW), a1);
            add_rows(sw.I0, a0, a1);
          }
          if (wrote1) {
            tmem_ld_32x32b_x16(tmem + lane_off + TM_D1 + (uint32_t)((sp * 2 + 1) * 2 * DW), a0);
            tmem_ld_32x32b_x16(tmem + lane_off + TM_D1 + (uint32_t)((sp * 2 + 1) * 2 * DW + DW), a1);
            add_rows(sw.I1, a0, a1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1_empty + 8 * sp);
        f1cnt ^= 1u << sp;
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

static int64_t padded_n(int32_t n) { return ((int64_t)n + BT - 1) / BT * BT; }

size_t workspace_bytes(int32_t n) {
  return 512 + 6 * align_up((size_t)padded_n(n) * DW * sizeof(__half), 256);
}

int super_blocks(int32_t n) { return (int)((padded_n(n) / BT + 1) / 2); }

bool eligible(int32_t n, int32_t d) {
  const int mode = path_mode(B2_PATH_GAE_DECODER);
  if (d < 1 || d > DW) return false;
  if (mode == 4) return true;                                  // forced (tests drive tiny graphs through it)
  return mode == 0 && (int64_t)n * n >= (1ll << 24);
}

// All-pairs part for the super-block range [sb_begin, sb_end) of the cyclic pair schedule: adds into dz[n, d] (ALL rows — the
// caller zero-initialises it and, when the range is split over ranks, sums the per-rank results) and into loss_acc.
int launch(const float* z, int64_t ldz, int32_t n, int32_t d, int32_t sb_begin, int32_t sb_end, float coef, float* dz,
           double* loss_acc, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < workspace_bytes(n)) return B2_ERR_UNSUPPORTED;
  const int64_t npad = padded_n(n);
  char* w = reinterpret_cast<char*>(ws);
  uint32_t* maxbits = reinterpret_cast<uint32_t*>(w);
  float* scale = reinterpret_cast<float*>(w + 16);
  double* zsum = reinterpret_cast<double*>(w + 256);
  w += 512;
  const size_t pl = align_up((size_t)npad * DW * sizeof(__half), 256);
  __half* zah = reinterpret_cast<__half*>(w);
  __half* zal = reinterpret_cast<__half*>(w + pl);
  __half* zbh = reinterpret_cast<__half*>(w + 2 * pl);
  __half* zbl = reinterpret_cast<__half*>(w + 3 * pl);
  __half* zth = reinterpret_cast<__half*>(w + 4 * pl);
  __half* ztl = reinterpret_cast<__half*>(w + 5 * pl);
  B2_CHECK_CUDA(cudaMemsetAsync(maxbits, 0, 4, st));
  B2_CHECK_CUDA(cudaMemsetAsync(zsum, 0, 16 * sizeof(double), st));
  {
    int64_t blocks = ceil_div<int64_t>((int64_t)n * d, 256 * 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ldz, n, d, maxbits);
    B2_CHECK_LAUNCH("gsym::absmax_kernel");
    scale_kernel<<<1, 1, 0, st>>>(maxbits, scale);
    B2_CHECK_LAUNCH("gsym::scale_kernel");
    blocks = ceil_div<int64_t>(npad * DW, 256 * 4);
    const int64_t cap2 = (int64_t)sm_count() * 16;
    if (blocks > cap2) blocks = cap2;
    split_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ldz, n, d, npad, scale, zah, zal, zbh, zbl, zth, ztl);
    B2_CHECK_LAUNCH("gsym::split_kernel");
    if (sb_begin == 0) {     // the closed-form linear term is added once (by the rank that owns super-block 0)
      int cb = ceil_div(n, 4096);
      if (cb > sm_count() * 2) cb = sm_count() * 2;
      colsum_kernel<<<cb, 256, 0, st>>>(z, ldz, n, d, zsum);
      B2_CHECK_LAUNCH("gsym::colsum_kernel");
      linear_term_kernel<<<1, 1, 0, st>>>(zsum, d, coef, loss_acc);
      B2_CHECK_LAUNCH("gsym::linear_term_kernel");
    }
  }
  Params p;
  memset(&p, 0, sizeof(p));
  const int SW32 = (int)CU_TENSOR_MAP_SWIZZLE_32B, SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  bool ok = make_tensor_map_f16_ex(&p.mA_hi, zah, DW, (uint64_t)npad, DW, DW, BT, SW32) &&
            make_tensor_map_f16_ex(&p.mA_lo, zal, DW, (uint64_t)npad, DW, DW, BT, SW32) &&
            make_tensor_map_f16_ex(&p.mB_hi, zbh, DW, (uint64_t)npad, DW, DW, BT, SW32) &&
            make_tensor_map_f16_ex(&p.mB_lo, zbl, DW, (uint64_t)npad, DW, DW, BT, SW32) &&
            make_tensor_map_f16_ex(&p.mT_hi, zth, (uint64_t)npad, DW, (uint64_t)npad, 64, DW, SW128) &&
            make_tensor_map_f16_ex(&p.mT_lo, ztl, (uint64_t)npad, DW, (uint64_t)npad, 64, DW, SW128);
  if (!ok) return B2_ERR_UNSUPPORTED;
  p.stagger = tuning(B2_TUNE_GAE_STAGGER); p.late_gempty = tuning(B2_TUNE_GAE_LATE_GEMPTY);
  p.scale = scale; p.dz = dz; p.loss_acc = loss_acc; p.n = n; p.d = d; p.nb = (int)(npad / BT); p.sb_begin = sb_begin; p.coef = coef;
#ifdef B2_GAE_TRACE
  p.trace = g_trace; p.trace_tiles = g_trace_tiles;
#endif
  if (sb_end <= sb_begin) return B2_OK;
  const size_t smem = SMEM_BYTES + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(gae_sym_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CHECK_CUDA(cudaFuncSetAttribute(gae_sym_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  // CTAs per super-block: whole waves of SMs.  A rank's share of the super-blocks (e.g. 489 of 3 907 at 8 GPUs = 3.3 waves of 148)
  // would idle a large part of the last wave; cutting every sweep into 2–8 step ranges makes the grid many small waves.
  const int n_sb = sb_end - sb_begin, n_steps_min = p.nb / 2;
  int splits = tuning(B2_TUNE_GAE_SPLITS);
  if (splits <= 0) {
    splits = 1;
    double best = 0.0;
    const int cand[6] = {1, 2, 3, 4, 6, 8};
    for (int c = 0; c < 6; ++c) {
      if (cand[c] > 1 && n_steps_min / cand[c] < 4 * SEG_STEPS) break;
      const double waves = (double)n_sb * cand[c] / sm_count();
      const double eff = waves / ceil(waves);
      if (eff > best + 0.03) { best = eff; splits = cand[c]; }       // measured: 125 k cells 7.5 → 6.6 ms with 2–8 parts, 1 M cells neutral
    }
  }
  if (splits > 1 && n_steps_min / splits < 1) splits = 1;
  p.splits = splits;
  gae_sym_kernel<false><<<n_sb * splits, THREADS, smem, st>>>(p);
  B2_CHECK_LAUNCH("gae_sym_kernel");
  gae_sym_kernel<true><<<n_sb * splits, THREADS, smem, st>>>(p);      // no-op unless the embedding needed operand scaling
  B2_CHECK_LAUNCH("gae_sym_kernel<scaled>");
  return B2_OK;
}

}  // namespace gsym
}  // namespace b2

#ifdef B2_GAE_TRACE
extern "C" int b2_debug_gae_sym_trace(unsigned long long* buf, int tiles) {
  b2::gsym::g_trace = buf;
  b2::gsym::g_trace_tiles = tiles;
  return 0;
}
#endif
