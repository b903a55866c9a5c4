// tcgen05 GEMM for the dense feature projections (fp32 in / fp32 out):
//   C[M,N] = epilogue( op(A)[M,K] · op(B)[K,N] )
// Blackwell-native structure: persistent CTAs (one per SM), TMA (cp.async.bulk.tensor)
// stages the fp32 operand tiles into 128B-swizzled shared memory, one elected thread
// issues tcgen05.mma.kind::tf32 with the accumulator in TMEM (double-buffered), four
// epilogue warps read the accumulator back with tcgen05.ld and apply bias / activation /
// ReLU-mask / accumulate before storing.  All operand layouts (row-major A or Aᵀ, B or Bᵀ)
// are handled by choosing K-major or MN-major UMMA descriptors — nothing is transposed
// in memory.
//
// Precision modes
//   TF32   : one tcgen05.mma per k-step (inputs truncated to 10-bit mantissa by the MMA)
//   TF32X3 : fp32-accurate. Four "splitter" warps rewrite each staged tile in shared memory
//            as hi = x & 0xFFFFE000 (exactly representable in tf32) and lo = x - hi, and the
//            issuing thread accumulates lo·hi + hi·lo + hi·hi  (error ~2^-21 relative).
//
// Replaces torch.mm / nn.Linear on the reference path (scgnn2.py:352-370, 499).
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace b2 {
namespace tc {

constexpr int BM = 128;            // UMMA M (cta_group::1)
constexpr int BK = 32;             // k-block: 32 tf32 = 128 B = one swizzle span
constexpr int UK = 8;              // UMMA K for kind::tf32 (32 bytes)
constexpr int THREADS = 384;       // 12 warps: TMA, MMA, TMEM-alloc, idle, 4 epilogue, 4 splitter
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int MAX_STAGES = 8;
constexpr int ATOM_BYTES = BK * 128;  // one MN-major atom: BK k-rows x 128 B

struct Params {
  CUtensorMap tmA;
  CUtensorMap tmB;
  float* C;
  const float* bias;
  const float* mask;
  float* partial;          // split-K partial tiles [splits][M][N] (dense) or nullptr
  long long ldc, ldmask;
  int M, N, K;
  int BN;                  // 32 | 64 | 128
  int a_mn, b_mn;          // operand major-ness: 0 = K-major, 1 = MN-major
  int x3;                  // 3xTF32 split
  int act;
  float beta;
  int tiles_m, tiles_n, splits, kb_per_split, kb_total, stages;
  unsigned mn_lbo, mn_sbo, mn_layout;   // MN-major descriptor fields (bring-up overridable, see gemm_tc())
};

struct StageLayout {
  uint32_t a_hi, a_lo, b_hi, b_lo;  // byte offsets inside one stage
  uint32_t stage_bytes, b_bytes;
};
__host__ __device__ inline StageLayout stage_layout(int BN, int x3) {
  StageLayout L;
  L.b_bytes = (uint32_t)BN * BK * 4;
  L.a_hi = 0;
  L.a_lo = A_TILE_BYTES;
  L.b_hi = x3 ? 2 * A_TILE_BYTES : A_TILE_BYTES;
  L.b_lo = L.b_hi + L.b_bytes;
  L.stage_bytes = (x3 ? 2u : 1u) * (A_TILE_BYTES + L.b_bytes);
  return L;
}

__device__ __forceinline__ void split_tile(uint32_t hi_addr, uint32_t lo_addr, uint32_t bytes, int tid) {
  // elementwise, so the 128B swizzle pattern written by TMA is preserved in both copies.  The "hi" operand is the staged fp32 tile
  // itself: kind::tf32 reads the top 19 bits of each 32-bit container and ignores the low 13 mantissa bits — the truncation the
  // residual below is computed against — so only the lo plane is written (one third less shared-memory traffic for the splitters).
  for (uint32_t off = (uint32_t)tid * 16; off < bytes; off += 128 * 16) {
    uint32_t x0, x1, x2, x3;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(hi_addr + off));
    const uint32_t h0 = x0 & 0xFFFFE000u, h1 = x1 & 0xFFFFE000u, h2 = x2 & 0xFFFFE000u, h3 = x3 & 0xFFFFE000u;
    const float l0 = __uint_as_float(x0) - __uint_as_float(h0);
    const float l1 = __uint_as_float(x1) - __uint_as_float(h1);
    const float l2 = __uint_as_float(x2) - __uint_as_float(h2);
    const float l3 = __uint_as_float(x3) - __uint_as_float(h3);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo_addr + off), "f"(l0), "f"(l1), "f"(l2), "f"(l3) : "memory");
  }
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const StageLayout L = stage_layout(p.BN, p.x3);
  const uint32_t smem_base = smem_u32(smem);
  uint8_t* bar_area = smem + (size_t)p.stages * L.stage_bytes;
  const uint32_t bars = smem_u32(bar_area);
  // barrier slots (8 bytes each)
  const uint32_t full_bar = bars;                               // [stages]
  const uint32_t split_bar = bars + 8 * MAX_STAGES;             // [stages]
  const uint32_t empty_bar = bars + 16 * MAX_STAGES;            // [stages]
  const uint32_t tfull_bar = bars + 24 * MAX_STAGES;            // [2]
  const uint32_t tempty_bar = tfull_bar + 16;                   // [2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bar_area + 24 * MAX_STAGES + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // accumulator ring: 2 stages; in 3xTF32 mode each stage holds TWO accumulators (big = hi·hi, small = lo·hi + hi·lo)
  const uint32_t acc_stride = (uint32_t)(p.x3 ? 2 * p.BN : p.BN);
  const uint32_t need_cols = 2 * acc_stride;
  const uint32_t tmem_cols = need_cols <= 32 ? 32 : (need_cols <= 64 ? 64 : (need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(split_bar + 8 * s, 4);     // one arrival per splitter warp
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, 4);    // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = A_TILE_BYTES + L.b_bytes;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int ks = u % p.splits;
        const int t = u / p.splits;
        const int tn = t % p.tiles_n, tm = t / p.tiles_n;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          mbar_expect_tx(fb, tx_bytes);
          const uint32_t sa = smem_base + stage * L.stage_bytes + L.a_hi;
          const uint32_t sb = smem_base + stage * L.stage_bytes + L.b_hi;
          if (!p.a_mn) {
            tma_load_2d(sa, &p.tmA, fb, kb * BK, tm * BM);
          } else {
#pragma unroll
            for (int at = 0; at < BM / 32; ++at) tma_load_2d(sa + at * ATOM_BYTES, &p.tmA, fb, tm * BM + at * 32, kb * BK);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &p.tmB, fb, kb * BK, tn * p.BN);
          } else {
            for (int at = 0; at < p.BN / 32; ++at) tma_load_2d(sb + at * ATOM_BYTES, &p.tmB, fb, tn * p.BN + at * 32, kb * BK);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp runs the role; each k-block's MMAs and commits are one single-thread issue region (tc_common.cuh: issued from a
    // divergent `lane == 0` branch each MMA costs an ELECT / BRA.U.ANY loop plus per-instruction descriptor arithmetic)
    {
      const uint32_t idesc = umma_idesc(BM, p.BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? p.mn_lbo : 16, b_lbo = p.b_mn ? p.mn_lbo : 16;
      const uint32_t a_kstep = p.a_mn ? 1024u : (uint32_t)(UK * 4), b_kstep = p.b_mn ? 1024u : (uint32_t)(UK * 4);
      const uint32_t a_sbo = p.a_mn ? p.mn_sbo : 1024u, b_sbo = p.b_mn ? p.mn_sbo : 1024u;
      const uint32_t a_lay = p.a_mn ? p.mn_layout : 2u, b_lay = p.b_mn ? p.mn_layout : 2u;
      const uint32_t a_hiw = umma_desc_hi(a_sbo, a_lay), b_hiw = umma_desc_hi(b_sbo, b_lay);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int ks = u % p.splits;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        const uint32_t d_small = d_tmem + (uint32_t)p.BN;
        uint32_t accumulate = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait((p.x3 ? split_bar : full_bar) + 8 * stage, phase);
          tc_fence_after();
          const uint32_t st = smem_base + stage * L.stage_bytes;
          if (elect_one_pred()) {
            // single-thread issue region (tc_common.cuh): 32-bit descriptor words; the start-address field counts 16-byte units in
            // the low 14 bits (shared memory < 256 KB), so a k-step advances the low word by kstep / 16
            const uint32_t a_hi0 = umma_desc_lo(st + L.a_hi, a_lbo), b_hi0 = umma_desc_lo(st + L.b_hi, b_lbo);
            const uint32_t a_lo0 = umma_desc_lo(st + L.a_lo, a_lbo), b_lo0 = umma_desc_lo(st + L.b_lo, b_lbo);
            uint32_t accf = accumulate;
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint32_t ak = (k * a_kstep) >> 4, bk = (k * b_kstep) >> 4;
              if (p.x3) {
                // The TMEM accumulate truncates (measured: error grows linearly with the number of accumulations),
                // so the two small cross terms go to their own accumulator: the big one sees 1/3 of the additions
                // and the small one's truncation is ~2^-11 smaller in absolute terms.  Summed (RN) in the epilogue.
                umma_tf32_lo(d_small, a_lo0 + ak, a_hiw, b_hi0 + bk, b_hiw, idesc, accf);
                umma_tf32_lo(d_small, a_hi0 + ak, a_hiw, b_lo0 + bk, b_hiw, idesc, 1);
                umma_tf32_lo(d_tmem, a_hi0 + ak, a_hiw, b_hi0 + bk, b_hiw, idesc, accf);
              } else {
                umma_tf32_lo(d_tmem, a_hi0 + ak, a_hiw, b_hi0 + bk, b_hiw, idesc, accf);
              }
              accf = 1;
            }
            umma_commit(empty_bar + 8 * stage);                        // frees the smem stage once these MMAs have read it
            if (kb == kb1 - 1) umma_commit(tfull_bar + 8 * acc);       // accumulator complete → epilogue
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0) {                                // empty K range: nothing was issued, the epilogue still expects the hand-over
          if (elect_one_pred()) umma_commit(tfull_bar + 8 * acc);
          __syncwarp();
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ===================== splitter warps (3xTF32) =====================
    if (p.x3) {
      const int tid = threadIdx.x - 8 * 32;
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int ks = u % p.splits;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          const uint32_t st = smem_base + stage * L.stage_bytes;
          split_tile(st + L.a_hi, st + L.a_lo, A_TILE_BYTES, tid);
          split_tile(st + L.b_hi, st + L.b_lo, L.b_bytes, tid);
          fence_proxy_async();                   // generic-proxy writes → visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(split_bar + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int sub = warp & 3;                    // TMEM sub-partition this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int ks = u % p.splits;
      const int t = u / p.splits;
      const int tn = t % p.tiles_n, tm = t / p.tiles_n;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const int row = tm * BM + sub * 32 + lane;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)acc * acc_stride + (uint32_t)c0;
        tmem_ld_32x32b_x32(taddr, v);
        if (p.x3) {
          uint32_t w[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)p.BN, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
        const int col0 = tn * p.BN + c0;
        if (row < p.M && col0 < p.N) {
          if (p.partial) {
            float* dst = p.partial + ((size_t)ks * p.M + row) * p.N + col0;
            const bool pv = ((p.N & 3) == 0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (pv && col0 + j + 3 < p.N) {
                *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) if (col0 + j + jj < p.N) dst[j + jj] = __uint_as_float(v[j + jj]);
              }
            }
          } else {
            float* dst = p.C + (size_t)row * p.ldc + col0;
            const float* mrow = p.mask ? p.mask + (size_t)row * p.ldmask + col0 : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float o[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int col = col0 + j + jj;
                float x = __uint_as_float(v[j + jj]);
                if (col < p.N) {
                  if (p.bias) x += __ldg(p.bias + col);
                  x = apply_act(x, p.act);
                  if (mrow && !(__ldg(mrow + j + jj) > 0.f)) x = 0.f;
                  if (p.beta != 0.f) x = fmaf(p.beta, dst[j + jj], x);
                }
                o[jj] = x;
              }
              if (vec_ok && col0 + j + 3 < p.N) {
                *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) if (col0 + j + jj < p.N) dst[j + jj] = o[jj];
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// deterministic split-K reduction + epilogue
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ C, long long ldc,
                     const float* __restrict__ bias, int act, const float* __restrict__ mask, long long ldmask,
                     float beta) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / N), c = (int)(i % N);
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(size_t)k * total + i];
    if (bias) s += bias[c];
    s = apply_act(s, act);
    if (mask && !(mask[(size_t)r * ldmask + c] > 0.f)) s = 0.f;
    float* dst = C + (size_t)r * ldc + c;
    if (beta != 0.f) s = fmaf(beta, *dst, s);
    *dst = s;
  }
}

// ---- host side -----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows `ld` elements apart; box = {32, box_outer}
// general 2-D fp32 tensor map: box = {box_inner, box_outer}; swizzle: 0 none, 1 32B, 2 64B, 3 128B, 4 128B_ATOM_32B
bool make_tensor_map_f32_ex(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                            uint32_t box_outer, int swizzle) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_tensor_map_f16_ex(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                            uint32_t box_outer, int swizzle) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_tensor_map_f32(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer,
                         bool mn_major) {
  const int mn_swz = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  return make_tensor_map_f32_ex(map, ptr, inner, outer, ld, 32, box_outer, mn_major ? mn_swz : (int)CU_TENSOR_MAP_SWIZZLE_128B);
}

struct Plan {
  int BN, stages, tiles_m, tiles_n, splits, kb_per_split, kb_total;
  size_t smem;
};

static Plan make_plan(int M, int N, int K, int x3) {
  Plan pl;
  pl.BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  const StageLayout L = stage_layout(pl.BN, x3);
  const size_t budget = 200 * 1024;
  int stages = (int)(budget / L.stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  pl.stages = stages;
  pl.smem = (size_t)stages * L.stage_bytes + 1024 /*align slack*/ + 24 * MAX_STAGES + 64;
  pl.tiles_m = ceil_div(M, BM);
  pl.tiles_n = ceil_div(N, pl.BN);
  pl.kb_total = ceil_div(K, BK);
  const int tiles = pl.tiles_m * pl.tiles_n;
  const int sms = sm_count();
  int splits = 1;
  if (tiles * 10 < sms * 6 && pl.kb_total >= 8) {
    splits = ceil_div(sms, tiles);
    const int max_splits = pl.kb_total / 4;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  pl.kb_per_split = ceil_div(pl.kb_total, splits);
  pl.splits = ceil_div(pl.kb_total, pl.kb_per_split);
  return pl;
}

}  // namespace tc

size_t gemm_tc_workspace_bytes(int M, int N, int K, int transA, int transB, int precision) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const tc::Plan pl = tc::make_plan(M, N, K, precision == B2_PREC_TF32X3);
  return pl.splits > 1 ? (size_t)pl.splits * M * N * sizeof(float) : 0;
}

int gemm_tc(const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, float* C, int64_t ldc,
            int M, int N, int K, const float* bias, int act, const float* mask, int64_t ldmask, float beta,
            int precision, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  using namespace tc;
  // qualification: TMA needs 16-byte aligned bases and row pitches; tiny problems stay on the CUDA-core kernel
  if (K < 8 || (int64_t)M * N * K < (1ll << 18)) return B2_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda & 3) || (ldb & 3))
    return B2_ERR_UNSUPPORTED;
  if (!get_encode()) return B2_ERR_UNSUPPORTED;

  const int x3 = precision == B2_PREC_TF32X3;
  const Plan pl = make_plan(M, N, K, x3);
  Params p;
  memset(&p, 0, sizeof(p));
  // A operand: K-major when A is [M,K] row-major, MN-major when stored [K,M]
  const bool okA = transA ? make_tensor_map_f32(&p.tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK, true)
                          : make_tensor_map_f32(&p.tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BM, false);
  // B operand (UMMA B is N x K): K-major when B is stored [N,K] (transB), MN-major when [K,N]
  const bool okB = transB ? make_tensor_map_f32(&p.tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)pl.BN, false)
                          : make_tensor_map_f32(&p.tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK, true);
  if (!okA || !okB) return B2_ERR_UNSUPPORTED;
  p.C = C; p.bias = bias; p.mask = mask; p.ldc = ldc; p.ldmask = ldmask;
  p.M = M; p.N = N; p.K = K; p.BN = pl.BN;
  p.a_mn = transA ? 1 : 0;
  p.b_mn = transB ? 0 : 1;
  p.x3 = x3; p.act = act; p.beta = beta;
  p.tiles_m = pl.tiles_m; p.tiles_n = pl.tiles_n; p.splits = pl.splits; p.kb_per_split = pl.kb_per_split;
  p.kb_total = pl.kb_total; p.stages = pl.stages;
  p.mn_lbo = ATOM_BYTES; p.mn_sbo = 512; p.mn_layout = 1;
  p.partial = nullptr;
  if (pl.splits > 1) {
    const size_t need = (size_t)pl.splits * M * N * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("b2_gemm_f32: split-K workspace too small (%zu < %zu)", workspace_bytes, need);
      return B2_ERR_WORKSPACE;
    }
    p.partial = reinterpret_cast<float*>(workspace);
  }
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int units = pl.tiles_m * pl.tiles_n * pl.splits;
  const int grid = units < sm_count() ? units : sm_count();
  gemm_tc_kernel<<<grid, THREADS, pl.smem, st>>>(p);
  B2_CHECK_LAUNCH("gemm_tc_kernel");
  if (pl.splits > 1) {
    size_t total = (size_t)M * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(p.partial, pl.splits, M, N, C, ldc, bias, act, mask, ldmask, beta);
    B2_CHECK_LAUNCH("splitk_reduce_kernel");
  }
  return B2_OK;
}

}  // namespace b2
