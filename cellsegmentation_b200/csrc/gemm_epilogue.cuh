// Shared epilogue of the tcgen05 conv kernels: TMEM accumulator -> (+bias, +residual hi/lo,
// ReLU) -> bf16 hi / lo (and optional raw fp32) rows in global memory.
//
// One warp owns 32 accumulator rows (its TMEM lane quadrant) and kChunks * 32 consecutive
// columns.  The residual of chunk c+1 is fetched with 256-bit loads while chunk c is
// converted and stored, and the first fetch is issued before the accumulator-ready wait.
#pragma once
#include "tc_ptx.cuh"

namespace cs {

struct EpiArgs {
  const float* bias;            // indexed by output column
  const __nv_bfloat16* res_hi;  // nullable
  const __nv_bfloat16* res_lo;  // nullable
  __nv_bfloat16* out_hi;        // nullable
  __nv_bfloat16* out_lo;        // nullable
  float* out_f32;               // nullable
  int relu;
};

// row_ok : this lane's row exists          off0  : element offset of (row, first column)
// col0   : first output column (bias index) taddr : TMEM address of (lane quadrant, first column)
template <int kChunks>
__device__ __forceinline__ void epilogue_warp(const EpiArgs& e, bool row_ok, int64_t off0, int col0,
                                              uint32_t taddr, uint32_t tfull_bar, uint32_t phase) {
  const bool has_res_hi = e.res_hi != nullptr, has_res_lo = e.res_lo != nullptr;
  U32x8 rh[2][2], rl[2][2];
  auto load_res = [&](int c, int slot) {
    if (row_ok && has_res_hi) {
      rh[slot][0] = ldg256(e.res_hi + off0 + c * 32);
      rh[slot][1] = ldg256(e.res_hi + off0 + c * 32 + 16);
    }
    if (row_ok && has_res_lo) {
      rl[slot][0] = ldg256(e.res_lo + off0 + c * 32);
      rl[slot][1] = ldg256(e.res_lo + off0 + c * 32 + 16);
    }
  };
  load_res(0, 0);
  mbar_wait(tfull_bar, phase);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int slot = c & 1;
    if (c + 1 < kChunks) load_res(c + 1, slot ^ 1);
    uint32_t r[32];
    tmem_ld32(taddr + (uint32_t)(c * 32), r);
    tmem_ld_wait();
    if (row_ok) {
      const int64_t off = off0 + c * 32;
      float v[32];
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0 + c * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 bb = __ldg(b4 + j);
        v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bb.x;
        v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
        v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z;
        v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
      }
      if (has_res_hi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t q = rh[slot][j >> 3].v[j & 7];
          v[2 * j] += bf16_lo_f(q);
          v[2 * j + 1] += bf16_hi_f(q);
        }
      }
      if (has_res_lo) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t q = rl[slot][j >> 3].v[j & 7];
          v[2 * j] += bf16_lo_f(q);
          v[2 * j + 1] += bf16_hi_f(q);
        }
      }
      if (e.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (e.out_f32) {
        float4* of = reinterpret_cast<float4*>(e.out_f32 + off);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          of[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      U32x8 hi[2];
#pragma unroll
      for (int j = 0; j < 16; ++j) hi[j >> 3].v[j & 7] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      if (e.out_hi) {
        stg256(e.out_hi + off, hi[0]);
        stg256(e.out_hi + off + 16, hi[1]);
      }
      if (e.out_lo) {
        U32x8 lo[2];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t h = hi[j >> 3].v[j & 7];
          lo[j >> 3].v[j & 7] = pack_bf16x2(v[2 * j] - bf16_lo_f(h), v[2 * j + 1] - bf16_hi_f(h));
        }
        stg256(e.out_lo + off, lo[0]);
        stg256(e.out_lo + off + 16, lo[1]);
      }
    }
  }
}

}  // namespace cs
