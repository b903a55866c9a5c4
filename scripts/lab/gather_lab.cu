// gather_lab — experimental gather mechanisms for the CSR aggregate on a real graph (lab tool, NOT part of the product library).
//
//   gather_lab <csr.bin | synth:N> <variant> [reps]
//
// csr.bin: int64 n, int64 nnz, int32 rowptr[n+1], int32 colidx[nnz], float vals[nnz]   (scripts/spmm_dump.py writes it)
// Every variant computes the full Y = A·X (F = 32) and is checked against a naive kernel; timing = CUDA events, median of `reps`
// launches, each after a 256 MB L2 flush.
//
// variants
//   rg:<dt>:<tch>                      row-per-lane-group register gathers, warp-uniform control flow (dt = f32 | bf16)
//   st:<dt>:<mech>:<cb>:<ng>           nnz-stream kernel: rows gathered into a per-warp shared-memory ring, consumed in CSR order
//        mech = l (cp.async 16 B, LDGSTS) | b (cp.async.bulk, one row per lane) | g (TMA tile::gather4, box rows 1) | h (gather4, box rows 4)
//        cb   = bytes per lane at consumption (4, 8, 16);  ng = 32-nnz blocks in flight per warp
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      fprintf(stderr, "CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                    \
    }                                                                                             \
  } while (0)

static constexpr int F = 32;
static constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------ reference
__global__ void ref_kernel(const int32_t* rowptr, const int32_t* colidx, const float* vals, const float* X, float* Y, int n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * F) return;
  const int r = (int)(t / F), f = (int)(t % F);
  float acc = 0.f;
  for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) acc = fmaf(vals[e], X[(int64_t)colidx[e] * F + f], acc);
  Y[t] = acc;
}

__global__ void to_bf16_kernel(const float* X, __nv_bfloat16* Xb, float* Xr, int64_t total) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const __nv_bfloat16 b = __float2bfloat16_rn(X[t]);
  Xb[t] = b;
  Xr[t] = __bfloat162float(b);
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async_16_hint(uint32_t dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_4_hint(uint32_t dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void stg_cs(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void stg_cs2(float* p, float a, float b) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void stg_cs4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// smallest r in [0, n_rows] with rowptr[r] >= t (rowptr non-decreasing, rowptr[n_rows] >= t); warp-cooperative 32-ary search
__device__ __forceinline__ int warp_lower_bound(const int32_t* __restrict__ rowptr, int n_rows, int64_t t, int lane) {
  int lo = 0, hi = n_rows;
  while (hi > lo) {
    const int span = hi - lo;
    const int step = (span + 31) >> 5;
    const int seg_lo = lo + lane * step;
    int q = seg_lo + step - 1;
    if (q > hi - 1) q = hi - 1;
    const bool in = seg_lo < hi;
    const bool pred = in ? ((int64_t)__ldg(rowptr + q) >= t) : true;
    const unsigned m = __ballot_sync(FULL, pred);
    const int f = __ffs(m) - 1;
    const int flo = lo + f * step;
    if (f < 0 || flo >= hi) {
      lo = hi;
    } else {
      int fhi = flo + step - 1;
      if (fhi > hi - 1) fhi = hi - 1;
      lo = flo;
      hi = fhi;
    }
  }
  return lo;
}

template <int DT> __device__ __forceinline__ void unpack2(uint32_t u, float& a, float& b) {
  if (DT == 0) {
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xffff0000u);
  }
}

// ------------------------------------------------------------------------------------------------ rowgroup kernel (register gathers)
// G lanes own one output row (16 bytes of the operand row per lane); RPW = 32 / G rows per warp.  Control flow is warp-uniform
// (loop bounds = the longest row of the group), so every shuffle uses the full mask and compiles to one SHFL.
template <int DT, int TCH>
__global__ void __launch_bounds__(256)
rowgroup_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                const uint8_t* __restrict__ X, int64_t ldxb, float* __restrict__ Y, int n_rows) {
  constexpr int RB = DT == 2 ? F * 4 : F * 2;
  constexpr int G = RB / 16;
  constexpr int RPW = 32 / G;
  constexpr int NV = DT == 2 ? 4 : 8;
  constexpr int WIN = 32;
  constexpr int PRE = WIN / G;
  const int lane = threadIdx.x & 31, sub = lane / G, gl = lane % G;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ngroups = ((int64_t)n_rows + RPW - 1) / RPW;
  for (int64_t grp = warp0; grp < ngroups; grp += nwarps) {
    const int64_t row = grp * RPW + sub;
    const bool valid = row < n_rows;
    const int start = valid ? __ldg(rowptr + row) : 0;
    const int end = valid ? __ldg(rowptr + row + 1) : 0;
    const int maxlen = __reduce_max_sync(FULL, end - start);
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    for (int base = 0; base < maxlen; base += WIN) {
      int pc[PRE];
      float pw[PRE];
#pragma unroll
      for (int q = 0; q < PRE; ++q) {
        const int e = start + base + q * G + gl;
        pc[q] = -1;
        pw[q] = 0.f;
        if (e < end) {
          pc[q] = __ldg(colidx + e);
          pw[q] = __ldg(vals + e);
        }
      }
#pragma unroll
      for (int bt = 0; bt < WIN / TCH; ++bt) {
        if (base + bt * TCH >= maxlen) break;
        uint4 x[TCH];
        float w[TCH];
#pragma unroll
        for (int t = 0; t < TCH; ++t) {
          const int idx = bt * TCH + t;
          const int cc = __shfl_sync(FULL, pc[idx / G], sub * G + idx % G);
          w[t] = __shfl_sync(FULL, pw[idx / G], sub * G + idx % G);
          x[t] = make_uint4(0u, 0u, 0u, 0u);
          if (cc >= 0) x[t] = ldg_nc16(X + (int64_t)cc * ldxb + gl * 16);
        }
#pragma unroll
        for (int t = 0; t < TCH; ++t) {
          if (DT == 2) {
            acc[0] = fmaf(w[t], __uint_as_float(x[t].x), acc[0]);
            acc[1] = fmaf(w[t], __uint_as_float(x[t].y), acc[1]);
            acc[2] = fmaf(w[t], __uint_as_float(x[t].z), acc[2]);
            acc[3] = fmaf(w[t], __uint_as_float(x[t].w), acc[3]);
          } else {
            float a, b;
            unpack2<0>(x[t].x, a, b); acc[0 % NV] = fmaf(w[t], a, acc[0 % NV]); acc[1 % NV] = fmaf(w[t], b, acc[1 % NV]);
            unpack2<0>(x[t].y, a, b); acc[2 % NV] = fmaf(w[t], a, acc[2 % NV]); acc[3 % NV] = fmaf(w[t], b, acc[3 % NV]);
            unpack2<0>(x[t].z, a, b); acc[4 % NV] = fmaf(w[t], a, acc[4 % NV]); acc[5 % NV] = fmaf(w[t], b, acc[5 % NV]);
            unpack2<0>(x[t].w, a, b); acc[6 % NV] = fmaf(w[t], a, acc[6 % NV]); acc[7 % NV] = fmaf(w[t], b, acc[7 % NV]);
          }
        }
      }
    }
    if (valid) {
      float* y = Y + row * F + gl * NV;
      stg_cs4(y, acc[0], acc[1], acc[2], acc[3]);
      if (NV == 8) stg_cs4(y + 4, acc[4 % NV], acc[5 % NV], acc[6 % NV], acc[7 % NV]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ nnz-stream kernel (shared-memory ring)
enum { MECH_LDGSTS = 0, MECH_BULK = 1, MECH_GATHER4 = 2 };

template <int DT, int CB, int NG, int WARPS, int MECH>
__global__ void __launch_bounds__(WARPS * 32)
stream_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const float* __restrict__ vals,
              const uint8_t* __restrict__ X, int64_t ldxb, float* __restrict__ Y, int n_rows, int64_t nnz,
              const __grid_constant__ CUtensorMap tmap, int* __restrict__ err) {
  constexpr int ESZ = DT == 2 ? 4 : 2;
  constexpr int RB = F * ESZ;          // bytes of one gathered row
  constexpr int BLK = 32;              // nnz per block (one per lane)
  constexpr int LPR = RB / 16;         // lanes per row for 16-byte copies
  constexpr int RPI = 32 / LPR;        // rows per cp.async instruction
  constexpr int LPRC = RB / CB;        // lanes per row at consumption
  constexpr int NPI = 32 / LPRC;       // nnz per consumption step
  constexpr int NV = CB / ESZ;         // values per lane
  static_assert(LPRC <= 32 && NPI >= 1, "row too wide for this lab kernel");
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + (size_t)warp * (NG * BLK * RB);
  float* vring = reinterpret_cast<float*>(smem + (size_t)WARPS * NG * BLK * RB) + warp * (NG * BLK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * NG * BLK * (RB + 4)) + warp * NG;
  const uint32_t ring_s = smem_u32(ring);
  const uint32_t bars_s = smem_u32(bars);
  if (MECH != MECH_LDGSTS) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NG; ++s) mbar_init(bars_s + 8 * s, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  const int W = gridDim.x * WARPS, w = blockIdx.x * WARPS + warp;
  const int64_t t0 = (int64_t)w * nnz / W, t1 = (int64_t)(w + 1) * nnz / W;
  const int R0 = (w == 0) ? 0 : warp_lower_bound(rowptr, n_rows, t0, lane);
  const int R1 = (w == W - 1) ? n_rows : warp_lower_bound(rowptr, n_rows, t1, lane);
  if (R0 >= R1) return;
  const int E0 = __ldg(rowptr + R0), E1 = __ldg(rowptr + R1);
  const int nblk = (E1 - E0 + BLK - 1) / BLK;

  // row-pointer window: lane l holds rowptr[rb + 1 + l]; the next window is prefetched
  int rb = R0;
  int rpv = __ldg(rowptr + min(rb + 1 + lane, n_rows));
  int rpn = __ldg(rowptr + min(rb + 33 + lane, n_rows));
  int r = R0, rbeg = E0, rend = __shfl_sync(FULL, rpv, 0);
  (void)rbeg;
  const int subc = lane / LPRC, glc = lane % LPRC;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  int c_n = -1;
  float v_n = 0.f;
  {
    const int e = E0 + lane;
    if (e < E1) {
      c_n = __ldg(colidx + e);
      v_n = __ldg(vals + e);
    }
  }
  int st_i = 0;                // ring stage of the block being issued
  int st_c = 0, ph_c = 0;      // ring stage / phase parity of the block being consumed
  for (int b = 0; b < nblk + NG - 1; ++b) {
    if (b < nblk) {
      const int c = c_n;
      const float v = v_n;
      c_n = -1;
      v_n = 0.f;
      {
        const int e = E0 + (b + 1) * BLK + lane;
        if (b + 1 < nblk && e < E1) {
          c_n = __ldg(colidx + e);
          v_n = __ldg(vals + e);
        }
      }
      __syncwarp();
      vring[st_i * BLK + lane] = v;
      const int nvalid = min(BLK, E1 - (E0 + b * BLK));
      if (MECH == MECH_LDGSTS) {
#pragma unroll
        for (int i = 0; i < LPR; ++i) {
          const int idx = i * RPI + lane / LPR;
          const int cc = __shfl_sync(FULL, c, idx);
          if (cc >= 0) cp_async_16(ring_s + (st_i * BLK + idx) * RB + (lane % LPR) * 16, X + (int64_t)cc * ldxb + (lane % LPR) * 16);
        }
      } else if (MECH == MECH_BULK) {
        if (lane == 0) mbar_expect_tx(bars_s + 8 * st_i, (uint32_t)nvalid * RB);
        __syncwarp();
        if (c >= 0) bulk_copy(ring_s + (st_i * BLK + lane) * RB, X + (int64_t)c * ldxb, RB, bars_s + 8 * st_i);
      } else {
        const int ng4 = (nvalid + 3) >> 2;
        if (lane == 0) mbar_expect_tx(bars_s + 8 * st_i, (uint32_t)ng4 * 4 * RB);
        const int j = (lane & 7) * 4;
        int r0 = __shfl_sync(FULL, c, j), r1 = __shfl_sync(FULL, c, j + 1), r2 = __shfl_sync(FULL, c, j + 2), r3 = __shfl_sync(FULL, c, j + 3);
        r0 = max(r0, 0); r1 = max(r1, 0); r2 = max(r2, 0); r3 = max(r3, 0);
        __syncwarp();
        if (lane < ng4) tma_gather4(ring_s + (st_i * BLK + j) * RB, &tmap, bars_s + 8 * st_i, 0, r0, r1, r2, r3);
      }
      st_i = (st_i + 1 == NG) ? 0 : st_i + 1;
    }
    if (MECH == MECH_LDGSTS) cp_async_commit();
    if (b >= NG - 1) {
      const int bc = b - (NG - 1);
      if (MECH == MECH_LDGSTS) {
        cp_async_wait<NG - 1>();
      } else {
        bool ok = false;
        for (int it = 0; it < (1 << 24); ++it) {
          if (mbar_test(bars_s + 8 * st_c, ph_c)) { ok = true; break; }
        }
        if (!ok) {
          if (lane == 0) atomicExch(err, 1);
          return;
        }
      }
      __syncwarp();
      const int eb = E0 + bc * BLK;
      const int eend = min(E1, eb + BLK);
      const uint8_t* blk = ring + (size_t)st_c * BLK * RB;
      const float* vb = vring + st_c * BLK;
      int e = eb;
      while (true) {
        while (r < R1 && rend <= e) {
          // ---- row r complete: combine the NPI lane groups, write the row ----
#pragma unroll
          for (int o = LPRC; o < 32; o <<= 1) {
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
          }
          if (lane < LPRC) {
            float* y = Y + (int64_t)r * F + lane * NV;
            if (NV == 1) stg_cs(y, acc[0]);
            else if (NV == 2) stg_cs2(y, acc[0], acc[1 % NV]);
            else if (NV == 4) stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]);
            else { stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]); stg_cs4(y + 4, acc[4 % NV], acc[5 % NV], acc[6 % NV], acc[7 % NV]); }
          }
#pragma unroll
          for (int i = 0; i < NV; ++i) acc[i] = 0.f;
          ++r;
          rbeg = rend;
          int j = r - rb;
          if (j == 32) {
            rb += 32;
            rpv = rpn;
            rpn = __ldg(rowptr + min(rb + 33 + lane, n_rows));
            j = 0;
          }
          rend = __shfl_sync(FULL, rpv, j);
        }
        if (e >= eend || r >= R1) break;
        const int run_end = min(rend, eend);
#pragma unroll 4
        for (int k = e + subc; k < run_end; k += NPI) {
          const int slot = k - eb;
          const float wv = vb[slot];
          const uint8_t* src = blk + slot * RB + glc * CB;
          if (DT == 2) {
            if (NV == 1) {
              acc[0] = fmaf(wv, *reinterpret_cast<const float*>(src), acc[0]);
            } else if (NV == 2) {
              const float2 x = *reinterpret_cast<const float2*>(src);
              acc[0] = fmaf(wv, x.x, acc[0]); acc[1 % NV] = fmaf(wv, x.y, acc[1 % NV]);
            } else {
              const float4 x = *reinterpret_cast<const float4*>(src);
              acc[0] = fmaf(wv, x.x, acc[0]); acc[1 % NV] = fmaf(wv, x.y, acc[1 % NV]);
              acc[2 % NV] = fmaf(wv, x.z, acc[2 % NV]); acc[3 % NV] = fmaf(wv, x.w, acc[3 % NV]);
            }
          } else {
            float a, bb;
            if (NV == 2) {
              unpack2<0>(*reinterpret_cast<const uint32_t*>(src), a, bb);
              acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
            } else if (NV == 4) {
              const uint2 x = *reinterpret_cast<const uint2*>(src);
              unpack2<0>(x.x, a, bb); acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
              unpack2<0>(x.y, a, bb); acc[2 % NV] = fmaf(wv, a, acc[2 % NV]); acc[3 % NV] = fmaf(wv, bb, acc[3 % NV]);
            } else {
              const uint4 x = *reinterpret_cast<const uint4*>(src);
              unpack2<0>(x.x, a, bb); acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
              unpack2<0>(x.y, a, bb); acc[2 % NV] = fmaf(wv, a, acc[2 % NV]); acc[3 % NV] = fmaf(wv, bb, acc[3 % NV]);
              unpack2<0>(x.z, a, bb); acc[4 % NV] = fmaf(wv, a, acc[4 % NV]); acc[5 % NV] = fmaf(wv, bb, acc[5 % NV]);
              unpack2<0>(x.w, a, bb); acc[6 % NV] = fmaf(wv, a, acc[6 % NV]); acc[7 % NV] = fmaf(wv, bb, acc[7 % NV]);
            }
          }
        }
        e = run_end;
      }
      if (st_c + 1 == NG) { st_c = 0; ph_c ^= 1; } else { ++st_c; }
    }
  }
  // rows that end exactly at E1 (and trailing empty rows of this warp's range)
  while (r < R1) {
#pragma unroll
    for (int o = LPRC; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
    }
    if (lane < LPRC) {
      float* y = Y + (int64_t)r * F + lane * NV;
      if (NV == 1) stg_cs(y, acc[0]);
      else if (NV == 2) stg_cs2(y, acc[0], acc[1 % NV]);
      else if (NV == 4) stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]);
      else { stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]); stg_cs4(y + 4, acc[4 % NV], acc[5 % NV], acc[6 % NV], acc[7 % NV]); }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    ++r;
  }
}

// ------------------------------------------------------------------------------------------------ nnz-stream kernel, generation 2
// As stream_kernel<…, MECH_LDGSTS>, plus: the (col, val) stream itself is staged through shared memory with cp.async, NG blocks
// ahead, inside the same commit groups as the row gathers — generation 1 loaded the next block's pairs into registers one
// iteration ahead, and that load's latency (≈ the gather latency under load) paced every warp: throughput scaled with the number
// of warps and not with the ring depth.
template <int DT, int CB, int NG, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
stream2_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const float* __restrict__ vals,
               const uint8_t* __restrict__ X, int64_t ldxb, float* __restrict__ Y, int n_rows, int64_t nnz) {
  constexpr int ESZ = DT == 2 ? 4 : 2;
  constexpr int RB = F * ESZ;
  constexpr int BLK = 32;
  constexpr int LPR = RB / 16;
  constexpr int RPI = 32 / LPR;
  constexpr int LPRC = RB / CB;
  constexpr int NPI = 32 / LPRC;
  constexpr int NV = CB / ESZ;
  constexpr int NC = 2 * NG;           // blocks of (col, val) pairs resident per warp
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + (size_t)warp * (NG * BLK * RB);
  int32_t* cring = reinterpret_cast<int32_t*>(smem + (size_t)WARPS * NG * BLK * RB) + warp * (NC * BLK);
  float* vring = reinterpret_cast<float*>(smem + (size_t)WARPS * NG * BLK * RB + (size_t)WARPS * NC * BLK * 4) + warp * (NC * BLK);
  const uint32_t ring_s = smem_u32(ring), cring_s = smem_u32(cring), vring_s = smem_u32(vring);
  const int W = gridDim.x * WARPS, w = blockIdx.x * WARPS + warp;
  const int64_t t0 = (int64_t)w * nnz / W, t1 = (int64_t)(w + 1) * nnz / W;
  const int R0 = (w == 0) ? 0 : warp_lower_bound(rowptr, n_rows, t0, lane);
  const int R1 = (w == W - 1) ? n_rows : warp_lower_bound(rowptr, n_rows, t1, lane);
  if (R0 >= R1) return;
  const int E0 = __ldg(rowptr + R0), E1 = __ldg(rowptr + R1);
  const int nblk = (E1 - E0 + BLK - 1) / BLK;

  int rb = R0;
  int rpv = __ldg(rowptr + min(rb + 1 + lane, n_rows));
  int rpn = __ldg(rowptr + min(rb + 33 + lane, n_rows));
  int r = R0, rend = __shfl_sync(FULL, rpv, 0);
  const int subc = lane / LPRC, glc = lane % LPRC;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  auto issue_pairs = [&](int blk_i, int cs) {     // (col, val) of block blk_i → pair-ring slot cs
    const int e = E0 + blk_i * BLK + lane;
    if (blk_i < nblk && e < E1) {
      cp_async_4(cring_s + (cs * BLK + lane) * 4, colidx + e);
      cp_async_4(vring_s + (cs * BLK + lane) * 4, vals + e);
    } else {
      cring[cs * BLK + lane] = -1;
    }
  };
  // prologue: pairs of blocks 0 .. NG-1
#pragma unroll
  for (int j = 0; j < NG; ++j) issue_pairs(j, j);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

  int st_i = 0, cs_i = 0;      // row-ring stage / pair-ring slot of the block being issued
  int st_c = 0, cs_c = 0;      // … of the block being consumed
  int cs_p = NG;               // pair-ring slot receiving block b + NG
  for (int b = 0; b < nblk + NG - 1; ++b) {
    if (b < nblk) {
      const int c = cring[cs_i * BLK + lane];
#pragma unroll
      for (int i = 0; i < LPR; ++i) {
        const int idx = i * RPI + lane / LPR;
        const int cc = __shfl_sync(FULL, c, idx);
        if (cc >= 0) cp_async_16(ring_s + (st_i * BLK + idx) * RB + (lane % LPR) * 16, X + (int64_t)cc * ldxb + (lane % LPR) * 16);
      }
      issue_pairs(b + NG, cs_p);
      st_i = (st_i + 1 == NG) ? 0 : st_i + 1;
      cs_i = (cs_i + 1 == NC) ? 0 : cs_i + 1;
      cs_p = (cs_p + 1 == NC) ? 0 : cs_p + 1;
    }
    cp_async_commit();
    if (b >= NG - 1) {
      const int bc = b - (NG - 1);
      cp_async_wait<NG - 1>();
      __syncwarp();
      const int eb = E0 + bc * BLK;
      const int eend = min(E1, eb + BLK);
      const uint8_t* blk = ring + (size_t)st_c * BLK * RB;
      const float* vb = vring + cs_c * BLK;
      int e = eb;
      while (true) {
        while (r < R1 && rend <= e) {
#pragma unroll
          for (int o = LPRC; o < 32; o <<= 1) {
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
          }
          if (lane < LPRC) {
            float* y = Y + (int64_t)r * F + lane * NV;
            if (NV == 1) stg_cs(y, acc[0]);
            else if (NV == 2) stg_cs2(y, acc[0], acc[1 % NV]);
            else if (NV == 4) stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]);
            else { stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]); stg_cs4(y + 4, acc[4 % NV], acc[5 % NV], acc[6 % NV], acc[7 % NV]); }
          }
#pragma unroll
          for (int i = 0; i < NV; ++i) acc[i] = 0.f;
          ++r;
          int j = r - rb;
          if (j == 32) {
            rb += 32;
            rpv = rpn;
            rpn = __ldg(rowptr + min(rb + 33 + lane, n_rows));
            j = 0;
          }
          rend = __shfl_sync(FULL, rpv, j);
        }
        if (e >= eend || r >= R1) break;
        const int run_end = min(rend, eend);
#pragma unroll 4
        for (int k = e + subc; k < run_end; k += NPI) {
          const int slot = k - eb;
          const float wv = vb[slot];
          const uint8_t* src = blk + slot * RB + glc * CB;
          if (DT == 2) {
            if (NV == 1) {
              acc[0] = fmaf(wv, *reinterpret_cast<const float*>(src), acc[0]);
            } else if (NV == 2) {
              const float2 x = *reinterpret_cast<const float2*>(src);
              acc[0] = fmaf(wv, x.x, acc[0]); acc[1 % NV] = fmaf(wv, x.y, acc[1 % NV]);
            } else {
              const float4 x = *reinterpret_cast<const float4*>(src);
              acc[0] = fmaf(wv, x.x, acc[0]); acc[1 % NV] = fmaf(wv, x.y, acc[1 % NV]);
              acc[2 % NV] = fmaf(wv, x.z, acc[2 % NV]); acc[3 % NV] = fmaf(wv, x.w, acc[3 % NV]);
            }
          } else {
            float a, bb;
            if (NV == 2) {
              unpack2<0>(*reinterpret_cast<const uint32_t*>(src), a, bb);
              acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
            } else if (NV == 4) {
              const uint2 x = *reinterpret_cast<const uint2*>(src);
              unpack2<0>(x.x, a, bb); acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
              unpack2<0>(x.y, a, bb); acc[2 % NV] = fmaf(wv, a, acc[2 % NV]); acc[3 % NV] = fmaf(wv, bb, acc[3 % NV]);
            } else {
              const uint4 x = *reinterpret_cast<const uint4*>(src);
              unpack2<0>(x.x, a, bb); acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
              unpack2<0>(x.y, a, bb); acc[2 % NV] = fmaf(wv, a, acc[2 % NV]); acc[3 % NV] = fmaf(wv, bb, acc[3 % NV]);
              unpack2<0>(x.z, a, bb); acc[4 % NV] = fmaf(wv, a, acc[4 % NV]); acc[5 % NV] = fmaf(wv, bb, acc[5 % NV]);
              unpack2<0>(x.w, a, bb); acc[6 % NV] = fmaf(wv, a, acc[6 % NV]); acc[7 % NV] = fmaf(wv, bb, acc[7 % NV]);
            }
          }
        }
        e = run_end;
      }
      st_c = (st_c + 1 == NG) ? 0 : st_c + 1;
      cs_c = (cs_c + 1 == NC) ? 0 : cs_c + 1;
    }
  }
  while (r < R1) {
#pragma unroll
    for (int o = LPRC; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
    }
    if (lane < LPRC) {
      float* y = Y + (int64_t)r * F + lane * NV;
      if (NV == 1) stg_cs(y, acc[0]);
      else if (NV == 2) stg_cs2(y, acc[0], acc[1 % NV]);
      else if (NV == 4) stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]);
      else { stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]); stg_cs4(y + 4, acc[4 % NV], acc[5 % NV], acc[6 % NV], acc[7 % NV]); }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    ++r;
  }
}

template <int DT, int CB, int NG, int WARPS, int HINT>
__global__ void __launch_bounds__(WARPS * 32)
stream3_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const float* __restrict__ vals,
               const uint8_t* __restrict__ X, int64_t ldxb, float* __restrict__ Y, int n_rows, int row_begin, int row_end) {
  constexpr int ESZ = DT == 2 ? 4 : 2;
  constexpr int RB = F * ESZ;
  constexpr int BLK = 32;
  constexpr int LPR = RB / 16;
  constexpr int RPI = 32 / LPR;
  constexpr int LPRC = RB / CB;
  constexpr int NPI = 32 / LPRC;
  constexpr int NV = CB / ESZ;
  constexpr int NC = 2 * NG;           // blocks of (col, val) pairs resident per warp
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + (size_t)warp * (NG * BLK * RB);
  int32_t* cring = reinterpret_cast<int32_t*>(smem + (size_t)WARPS * NG * BLK * RB) + warp * (NC * BLK);
  float* vring = reinterpret_cast<float*>(smem + (size_t)WARPS * NG * BLK * RB + (size_t)WARPS * NC * BLK * 4) + warp * (NC * BLK);
  const uint32_t ring_s = smem_u32(ring), cring_s = smem_u32(cring), vring_s = smem_u32(vring);
  const int W = gridDim.x * WARPS, w = blockIdx.x * WARPS + warp;
  // this launch covers rows [row_begin, row_end) (one slab); the slab's non-zeros are split evenly over the warps
  const int64_t s0 = __ldg(rowptr + row_begin), s1 = __ldg(rowptr + row_end);
  const int64_t t0 = s0 + (int64_t)w * (s1 - s0) / W, t1 = s0 + (int64_t)(w + 1) * (s1 - s0) / W;
  const int R0 = (w == 0) ? row_begin : max(row_begin, min(row_end, warp_lower_bound(rowptr, n_rows, t0, lane)));
  const int R1 = (w == W - 1) ? row_end : max(row_begin, min(row_end, warp_lower_bound(rowptr, n_rows, t1, lane)));
  const uint64_t pol_x = policy_evict_last(), pol_s = policy_evict_first();
  if (R0 >= R1) return;
  const int E0 = __ldg(rowptr + R0), E1 = __ldg(rowptr + R1);
  const int nblk = (E1 - E0 + BLK - 1) / BLK;

  int rb = R0;
  int rpv = __ldg(rowptr + min(rb + 1 + lane, n_rows));
  int rpn = __ldg(rowptr + min(rb + 33 + lane, n_rows));
  int r = R0, rend = __shfl_sync(FULL, rpv, 0);
  const int subc = lane / LPRC, glc = lane % LPRC;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  auto issue_pairs = [&](int blk_i, int cs) {     // (col, val) of block blk_i → pair-ring slot cs
    const int e = E0 + blk_i * BLK + lane;
    if (blk_i < nblk && e < E1) {
      if (HINT) {
        cp_async_4_hint(cring_s + (cs * BLK + lane) * 4, colidx + e, pol_s);
        cp_async_4_hint(vring_s + (cs * BLK + lane) * 4, vals + e, pol_s);
      } else {
        cp_async_4(cring_s + (cs * BLK + lane) * 4, colidx + e);
        cp_async_4(vring_s + (cs * BLK + lane) * 4, vals + e);
      }
    } else {
      cring[cs * BLK + lane] = -1;
    }
  };
  // prologue: pairs of blocks 0 .. NG-1
#pragma unroll
  for (int j = 0; j < NG; ++j) issue_pairs(j, j);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

  int st_i = 0, cs_i = 0;      // row-ring stage / pair-ring slot of the block being issued
  int st_c = 0, cs_c = 0;      // … of the block being consumed
  int cs_p = NG;               // pair-ring slot receiving block b + NG
  for (int b = 0; b < nblk + NG - 1; ++b) {
    if (b < nblk) {
      const int c = cring[cs_i * BLK + lane];
#pragma unroll
      for (int i = 0; i < LPR; ++i) {
        const int idx = i * RPI + lane / LPR;
        const int cc = __shfl_sync(FULL, c, idx);
        if (cc >= 0) {
          if (HINT) cp_async_16_hint(ring_s + (st_i * BLK + idx) * RB + (lane % LPR) * 16, X + (int64_t)cc * ldxb + (lane % LPR) * 16, pol_x);
          else cp_async_16(ring_s + (st_i * BLK + idx) * RB + (lane % LPR) * 16, X + (int64_t)cc * ldxb + (lane % LPR) * 16);
        }
      }
      issue_pairs(b + NG, cs_p);
      st_i = (st_i + 1 == NG) ? 0 : st_i + 1;
      cs_i = (cs_i + 1 == NC) ? 0 : cs_i + 1;
      cs_p = (cs_p + 1 == NC) ? 0 : cs_p + 1;
    }
    cp_async_commit();
    if (b >= NG - 1) {
      const int bc = b - (NG - 1);
      cp_async_wait<NG - 1>();
      __syncwarp();
      const int eb = E0 + bc * BLK;
      const int eend = min(E1, eb + BLK);
      const uint8_t* blk = ring + (size_t)st_c * BLK * RB;
      const float* vb = vring + cs_c * BLK;
      int e = eb;
      while (true) {
        while (r < R1 && rend <= e) {
#pragma unroll
          for (int o = LPRC; o < 32; o <<= 1) {
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
          }
          if (lane < LPRC) {
            float* y = Y + (int64_t)r * F + lane * NV;
            if (NV == 1) stg_cs(y, acc[0]);
            else if (NV == 2) stg_cs2(y, acc[0], acc[1 % NV]);
            else if (NV == 4) stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]);
            else { stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]); stg_cs4(y + 4, acc[4 % NV], acc[5 % NV], acc[6 % NV], acc[7 % NV]); }
          }
#pragma unroll
          for (int i = 0; i < NV; ++i) acc[i] = 0.f;
          ++r;
          int j = r - rb;
          if (j == 32) {
            rb += 32;
            rpv = rpn;
            rpn = __ldg(rowptr + min(rb + 33 + lane, n_rows));
            j = 0;
          }
          rend = __shfl_sync(FULL, rpv, j);
        }
        if (e >= eend || r >= R1) break;
        const int run_end = min(rend, eend);
#pragma unroll 4
        for (int k = e + subc; k < run_end; k += NPI) {
          const int slot = k - eb;
          const float wv = vb[slot];
          const uint8_t* src = blk + slot * RB + glc * CB;
          if (DT == 2) {
            if (NV == 1) {
              acc[0] = fmaf(wv, *reinterpret_cast<const float*>(src), acc[0]);
            } else if (NV == 2) {
              const float2 x = *reinterpret_cast<const float2*>(src);
              acc[0] = fmaf(wv, x.x, acc[0]); acc[1 % NV] = fmaf(wv, x.y, acc[1 % NV]);
            } else {
              const float4 x = *reinterpret_cast<const float4*>(src);
              acc[0] = fmaf(wv, x.x, acc[0]); acc[1 % NV] = fmaf(wv, x.y, acc[1 % NV]);
              acc[2 % NV] = fmaf(wv, x.z, acc[2 % NV]); acc[3 % NV] = fmaf(wv, x.w, acc[3 % NV]);
            }
          } else {
            float a, bb;
            if (NV == 2) {
              unpack2<0>(*reinterpret_cast<const uint32_t*>(src), a, bb);
              acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
            } else if (NV == 4) {
              const uint2 x = *reinterpret_cast<const uint2*>(src);
              unpack2<0>(x.x, a, bb); acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
              unpack2<0>(x.y, a, bb); acc[2 % NV] = fmaf(wv, a, acc[2 % NV]); acc[3 % NV] = fmaf(wv, bb, acc[3 % NV]);
            } else {
              const uint4 x = *reinterpret_cast<const uint4*>(src);
              unpack2<0>(x.x, a, bb); acc[0] = fmaf(wv, a, acc[0]); acc[1 % NV] = fmaf(wv, bb, acc[1 % NV]);
              unpack2<0>(x.y, a, bb); acc[2 % NV] = fmaf(wv, a, acc[2 % NV]); acc[3 % NV] = fmaf(wv, bb, acc[3 % NV]);
              unpack2<0>(x.z, a, bb); acc[4 % NV] = fmaf(wv, a, acc[4 % NV]); acc[5 % NV] = fmaf(wv, bb, acc[5 % NV]);
              unpack2<0>(x.w, a, bb); acc[6 % NV] = fmaf(wv, a, acc[6 % NV]); acc[7 % NV] = fmaf(wv, bb, acc[7 % NV]);
            }
          }
        }
        e = run_end;
      }
      st_c = (st_c + 1 == NG) ? 0 : st_c + 1;
      cs_c = (cs_c + 1 == NC) ? 0 : cs_c + 1;
    }
  }
  while (r < R1) {
#pragma unroll
    for (int o = LPRC; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
    }
    if (lane < LPRC) {
      float* y = Y + (int64_t)r * F + lane * NV;
      if (NV == 1) stg_cs(y, acc[0]);
      else if (NV == 2) stg_cs2(y, acc[0], acc[1 % NV]);
      else if (NV == 4) stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]);
      else { stg_cs4(y, acc[0], acc[1 % NV], acc[2 % NV], acc[3 % NV]); stg_cs4(y + 4, acc[4 % NV], acc[5 % NV], acc[6 % NV], acc[7 % NV]); }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    ++r;
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* map, void* ptr, int dt, uint64_t rows, uint32_t box_rows) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  const int esz = dt == 2 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)F, rows};
  cuuint64_t strides[1] = {(cuuint64_t)F * esz};
  cuuint32_t box[2] = {(cuuint32_t)F, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult rc = enc(map, dt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) fprintf(stderr, "cuTensorMapEncodeTiled(box rows %u) failed: %d\n", box_rows, (int)rc);
  return rc == CUDA_SUCCESS;
}

struct Ctx {
  int n = 0;
  int64_t nnz = 0;
  int32_t *rowptr = nullptr, *colidx = nullptr;
  float* vals = nullptr;
  float *X = nullptr, *Xr = nullptr, *Y = nullptr, *Yref = nullptr, *Yref16 = nullptr;
  __nv_bfloat16* Xb = nullptr;
  int* err = nullptr;
  int sms = 148;
};

template <int DT, int CB, int NG, int MECH>
static void launch_stream(const Ctx& c, const CUtensorMap& map, cudaStream_t st) {
  constexpr int WARPS = 8;
  constexpr int RB = F * (DT == 2 ? 4 : 2);
  const size_t smem = (size_t)WARPS * NG * 32 * (RB + 4) + WARPS * NG * 8 + 128;
  static bool once = false;
  auto kern = stream_kernel<DT, CB, NG, WARPS, MECH>;
  if (!once) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  int per_sm = (int)std::min<size_t>((227 * 1024) / (smem + 1024), 2048 / (WARPS * 32));
  if (per_sm < 1) per_sm = 1;
  const uint8_t* X = DT == 2 ? reinterpret_cast<const uint8_t*>(c.X) : reinterpret_cast<const uint8_t*>(c.Xb);
  kern<<<c.sms * per_sm, WARPS * 32, smem, st>>>(c.rowptr, c.colidx, c.vals, X, (int64_t)RB, c.Y, c.n, c.nnz, map, c.err);
}

template <int DT, int CB, int NG, int WARPS>
static void launch_stream2(const Ctx& c, const CUtensorMap&, cudaStream_t st) {
  constexpr int RB = F * (DT == 2 ? 4 : 2);
  const size_t smem = (size_t)WARPS * NG * 32 * RB + (size_t)WARPS * 2 * NG * 32 * 8 + 128;
  static bool once = false;
  auto kern = stream2_kernel<DT, CB, NG, WARPS>;
  if (!once) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  int per_sm = (int)std::min<size_t>((227 * 1024) / (smem + 1024), 2048 / (WARPS * 32));
  if (per_sm < 1) per_sm = 1;
  const uint8_t* X = DT == 2 ? reinterpret_cast<const uint8_t*>(c.X) : reinterpret_cast<const uint8_t*>(c.Xb);
  kern<<<c.sms * per_sm, WARPS * 32, smem, st>>>(c.rowptr, c.colidx, c.vals, X, (int64_t)RB, c.Y, c.n, c.nnz);
}

template <int DT, int CB, int NG, int WARPS, int HINT, int SLABS>
static void launch_stream3(const Ctx& c, const CUtensorMap&, cudaStream_t st) {
  constexpr int RB = F * (DT == 2 ? 4 : 2);
  const size_t smem = (size_t)WARPS * NG * 32 * RB + (size_t)WARPS * 2 * NG * 32 * 8 + 128;
  static bool once = false;
  auto kern = stream3_kernel<DT, CB, NG, WARPS, HINT>;
  if (!once) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  int per_sm = (int)std::min<size_t>((227 * 1024) / (smem + 1024), 2048 / (WARPS * 32));
  if (per_sm < 1) per_sm = 1;
  const uint8_t* X = DT == 2 ? reinterpret_cast<const uint8_t*>(c.X) : reinterpret_cast<const uint8_t*>(c.Xb);
  for (int sl = 0; sl < SLABS; ++sl) {
    const int r0 = (int)((int64_t)c.n * sl / SLABS), r1 = (int)((int64_t)c.n * (sl + 1) / SLABS);
    kern<<<c.sms * per_sm, WARPS * 32, smem, st>>>(c.rowptr, c.colidx, c.vals, X, (int64_t)RB, c.Y, c.n, r0, r1);
  }
}

template <int DT, int TCH>
static void launch_rg(const Ctx& c, cudaStream_t st) {
  constexpr int G = (DT == 2 ? F * 4 : F * 2) / 16;
  const int64_t groups = ((int64_t)c.n + 32 / G - 1) / (32 / G);
  int64_t blocks = (groups + 7) / 8;
  blocks = std::min<int64_t>(blocks, (int64_t)c.sms * 64);
  const uint8_t* X = DT == 2 ? reinterpret_cast<const uint8_t*>(c.X) : reinterpret_cast<const uint8_t*>(c.Xb);
  rowgroup_kernel<DT, TCH><<<(unsigned)blocks, 256, 0, st>>>(c.rowptr, c.colidx, c.vals, X, (int64_t)(DT == 2 ? F * 4 : F * 2), c.Y, c.n);
}

typedef void (*LaunchFn)(const Ctx&, const CUtensorMap&, cudaStream_t);
struct Variant { std::string name; int dt; int box_rows; LaunchFn fn; };

template <int DT, int TCH> static void rg_thunk(const Ctx& c, const CUtensorMap&, cudaStream_t st) { launch_rg<DT, TCH>(c, st); }

#define ST(DTN, DT, MN, MECH, CB, NG, BOX) \
  v.push_back({std::string("st:") + DTN + ":" + MN + ":" #CB ":" #NG, DT, BOX, launch_stream<DT, CB, NG, MECH>})

static std::vector<Variant> variants() {
  std::vector<Variant> v;
  v.push_back({"rg:f32:8", 2, 0, rg_thunk<2, 8>});
  v.push_back({"rg:f32:16", 2, 0, rg_thunk<2, 16>});
  v.push_back({"rg:bf16:8", 0, 0, rg_thunk<0, 8>});
  v.push_back({"rg:bf16:16", 0, 0, rg_thunk<0, 16>});
  ST("f32", 2, "l", MECH_LDGSTS, 4, 3, 0);   ST("f32", 2, "l", MECH_LDGSTS, 16, 3, 0);  ST("f32", 2, "l", MECH_LDGSTS, 16, 2, 0);
  ST("f32", 2, "l", MECH_LDGSTS, 16, 4, 0);  ST("f32", 2, "l", MECH_LDGSTS, 8, 3, 0);
  ST("f32", 2, "b", MECH_BULK, 16, 3, 0);    ST("f32", 2, "b", MECH_BULK, 16, 4, 0);     ST("f32", 2, "b", MECH_BULK, 4, 3, 0);
  ST("f32", 2, "g", MECH_GATHER4, 16, 3, 1); ST("f32", 2, "g", MECH_GATHER4, 16, 4, 1);
  ST("f32", 2, "h", MECH_GATHER4, 16, 3, 4);
  ST("bf16", 0, "l", MECH_LDGSTS, 4, 4, 0);  ST("bf16", 0, "l", MECH_LDGSTS, 8, 4, 0);   ST("bf16", 0, "l", MECH_LDGSTS, 4, 6, 0);
  ST("bf16", 0, "l", MECH_LDGSTS, 4, 3, 0);  ST("bf16", 0, "l", MECH_LDGSTS, 16, 4, 0);
  ST("bf16", 0, "b", MECH_BULK, 4, 4, 0);    ST("bf16", 0, "b", MECH_BULK, 4, 6, 0);     ST("bf16", 0, "b", MECH_BULK, 8, 6, 0);
  ST("bf16", 0, "g", MECH_GATHER4, 4, 4, 1); ST("bf16", 0, "g", MECH_GATHER4, 4, 6, 1);
  ST("bf16", 0, "h", MECH_GATHER4, 4, 4, 4);
#define S2(DTN, DT, CB, NG, WP) v.push_back({std::string("s2:") + DTN + ":" #CB ":" #NG ":" #WP, DT, 0, launch_stream2<DT, CB, NG, WP>})
  S2("f32", 2, 16, 2, 8);  S2("f32", 2, 16, 3, 8);  S2("f32", 2, 16, 4, 8);  S2("f32", 2, 16, 6, 8);  S2("f32", 2, 4, 3, 8);  S2("f32", 2, 8, 3, 8);
  S2("f32", 2, 16, 2, 4);  S2("f32", 2, 16, 3, 4);  S2("f32", 2, 16, 2, 16);
  S2("bf16", 0, 4, 2, 8);  S2("bf16", 0, 4, 3, 8);  S2("bf16", 0, 4, 4, 8);  S2("bf16", 0, 4, 6, 8);  S2("bf16", 0, 4, 8, 8);  S2("bf16", 0, 8, 4, 8);
  S2("bf16", 0, 16, 4, 8); S2("bf16", 0, 4, 4, 16);
#define S3(DTN, DT, CB, NG, WP, HINT, SLABS) \
  v.push_back({std::string("s3:") + DTN + ":" #CB ":" #NG ":" #WP ":h" #HINT ":s" #SLABS, DT, 0, launch_stream3<DT, CB, NG, WP, HINT, SLABS>})
  S3("f32", 2, 16, 2, 8, 0, 1);  S3("f32", 2, 16, 2, 8, 1, 1);  S3("f32", 2, 16, 2, 8, 0, 10); S3("f32", 2, 16, 2, 8, 1, 10);
  S3("f32", 2, 16, 2, 8, 1, 5);  S3("f32", 2, 16, 2, 8, 1, 20); S3("f32", 2, 16, 2, 8, 1, 40); S3("f32", 2, 16, 3, 8, 1, 10);
  S3("f32", 2, 16, 3, 4, 1, 10); S3("f32", 2, 16, 4, 8, 1, 10);
  S3("bf16", 0, 4, 2, 8, 0, 1);  S3("bf16", 0, 4, 2, 8, 1, 1);  S3("bf16", 0, 4, 2, 8, 1, 10); S3("bf16", 0, 4, 2, 8, 1, 20);
  S3("bf16", 0, 4, 3, 8, 1, 10); S3("bf16", 0, 4, 4, 8, 1, 10); S3("bf16", 0, 4, 2, 8, 0, 10);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: gather_lab <csr.bin|synth:N> <variant|list> [reps]\n");
    for (auto& v : variants()) fprintf(stderr, "  %s\n", v.name.c_str());
    return 1;
  }
  const std::string src = argv[1], want = argv[2];
  if (want == "list") {
    for (auto& v : variants()) printf("%s\n", v.name.c_str());
    return 0;
  }
  const int reps = argc > 3 ? atoi(argv[3]) : 7;
  Ctx c;
  std::vector<int32_t> rp, ci;
  std::vector<float> va;
  if (src.rfind("synth:", 0) == 0) {
    c.n = atoi(src.c_str() + 6);
    rp.resize(c.n + 1);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    rp[0] = 0;
    const int ncl = 10, per = (c.n + ncl - 1) / ncl;
    for (int i = 0; i < c.n; ++i) {
      const int deg = 16 + (int)(rnd() % 25);
      const int lo = (i / per) * per, hi = std::min(c.n, lo + per);
      std::vector<int> cols(deg);
      for (int& x : cols) x = lo + (int)(rnd() % (uint64_t)(hi - lo));
      std::sort(cols.begin(), cols.end());
      cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
      for (int x : cols) { ci.push_back(x); va.push_back(1.0f / (float)cols.size()); }
      rp[i + 1] = (int32_t)ci.size();
    }
    c.nnz = (int64_t)ci.size();
  } else {
    FILE* f = fopen(src.c_str(), "rb");
    if (!f) { perror("open csr"); return 1; }
    int64_t hdr[2];
    if (fread(hdr, 8, 2, f) != 2) return 1;
    c.n = (int)hdr[0];
    c.nnz = hdr[1];
    rp.resize(c.n + 1); ci.resize(c.nnz); va.resize(c.nnz);
    if (fread(rp.data(), 4, rp.size(), f) != rp.size() || fread(ci.data(), 4, ci.size(), f) != ci.size() ||
        fread(va.data(), 4, va.size(), f) != va.size()) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  c.sms = prop.multiProcessorCount;
  const int64_t NF = (int64_t)c.n * F;
  CK(cudaMalloc(&c.rowptr, (c.n + 1) * 4)); CK(cudaMalloc(&c.colidx, c.nnz * 4)); CK(cudaMalloc(&c.vals, c.nnz * 4));
  CK(cudaMemcpy(c.rowptr, rp.data(), (c.n + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.colidx, ci.data(), c.nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.vals, va.data(), c.nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c.X, NF * 4)); CK(cudaMalloc(&c.Xr, NF * 4)); CK(cudaMalloc(&c.Xb, NF * 2));
  CK(cudaMalloc(&c.Y, NF * 4)); CK(cudaMalloc(&c.Yref, NF * 4)); CK(cudaMalloc(&c.Yref16, NF * 4));
  CK(cudaMalloc(&c.err, 4)); CK(cudaMemset(c.err, 0, 4));
  {
    std::vector<float> hx(NF);
    uint64_t s = 1234567ull;
    for (auto& x : hx) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (float)((s >> 40) & 0xffff) / 65536.0f - 0.5f; }
    CK(cudaMemcpy(c.X, hx.data(), NF * 4, cudaMemcpyHostToDevice));
  }
  to_bf16_kernel<<<(unsigned)((NF + 255) / 256), 256>>>(c.X, c.Xb, c.Xr, NF);
  ref_kernel<<<(unsigned)((NF + 255) / 256), 256>>>(c.rowptr, c.colidx, c.vals, c.X, c.Yref, c.n);
  ref_kernel<<<(unsigned)((NF + 255) / 256), 256>>>(c.rowptr, c.colidx, c.vals, c.Xr, c.Yref16, c.n);
  CK(cudaDeviceSynchronize());
  std::vector<float> href(NF), href16(NF), hy(NF);
  CK(cudaMemcpy(href.data(), c.Yref, NF * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(href16.data(), c.Yref16, NF * 4, cudaMemcpyDeviceToHost));
  void* flush = nullptr;
  CK(cudaMalloc(&flush, 256 << 20));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  printf("graph: n=%d nnz=%lld (%.1f per row), F=%d, SMs=%d\n", c.n, (long long)c.nnz, (double)c.nnz / c.n, F, c.sms);
  for (auto& v : variants()) {
    if (want != "all" && want != v.name) continue;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (v.box_rows) {
      void* base = v.dt == 2 ? (void*)c.X : (void*)c.Xb;
      if (!make_map(&map, base, v.dt, (uint64_t)c.n, v.box_rows)) { printf("%-16s tensor map failed\n", v.name.c_str()); continue; }
    }
    CK(cudaMemsetAsync(c.Y, 0xff, NF * 4, st));
    v.fn(c, map, st);
    cudaError_t le = cudaStreamSynchronize(st);
    if (le != cudaSuccess) { printf("%-16s launch failed: %s\n", v.name.c_str(), cudaGetErrorString(le)); return 3; }
    int herr = 0;
    CK(cudaMemcpy(&herr, c.err, 4, cudaMemcpyDeviceToHost));
    if (herr) { printf("%-16s TIMEOUT waiting on an mbarrier (mechanism not delivering)\n", v.name.c_str()); CK(cudaMemset(c.err, 0, 4)); continue; }
    CK(cudaMemcpy(hy.data(), c.Y, NF * 4, cudaMemcpyDeviceToHost));
    const std::vector<float>& ref = v.dt == 2 ? href : href16;
    double num = 0, den = 0;
    int64_t bad = 0;
    for (int64_t i = 0; i < NF; ++i) {
      const double d = (double)hy[i] - ref[i];
      if (!(std::fabs(d) <= 1e30)) ++bad;
      else num += d * d;
      den += (double)ref[i] * ref[i];
    }
    const double rel = std::sqrt(num / den);
    std::vector<float> ts;
    for (int it = 0; it < reps; ++it) {
      CK(cudaMemsetAsync(flush, it, 256 << 20, st));
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      CK(cudaEventRecord(a, st));
      v.fn(c, map, st);
      CK(cudaEventRecord(b, st));
      CK(cudaEventSynchronize(b));
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      ts.push_back(ms);
      CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
    }
    std::sort(ts.begin(), ts.end());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a, st));
    for (int it = 0; it < 10; ++it) v.fn(c, map, st);
    CK(cudaEventRecord(b, st));
    CK(cudaEventSynchronize(b));
    float warm;
    CK(cudaEventElapsedTime(&warm, a, b));
    warm /= 10;
    const double esz = v.dt == 2 ? 4 : 2;
    const double alg = (double)c.nnz * 8 + (c.n + 1) * 4.0 + (double)NF * esz + (double)NF * 4;
    const double med = ts[ts.size() / 2];
    printf("%-16s cold %.3f ms (min %.3f)  %6.0f GB/s = %4.1f %% of 6566   back-to-back %.3f ms   rel err %.2e  bad %lld\n", v.name.c_str(), med,
           ts[0], alg / med / 1e6, alg / med / 1e6 / 6566.1 * 100, warm, rel, (long long)bad);
    fflush(stdout);
  }
  return 0;
}
