// Shared epilogue of the tcgen05 conv kernels: TMEM accumulator -> (+bias, +residual hi/lo,
// ReLU) -> bf16 hi / lo tiles, moved by TMA.
//
// Why TMA: accumulator rows are one per lane, so direct global loads/stores touch 32 different
// 128-byte lines per warp instruction = 32 LSU wavefronts of 32 B.  The ncu capture of the
// first version (profiles/r01_c) shows the residual layers pinned at 75 % of the LSU wavefront
// pipe with DRAM at 55 %.  Here the threads only touch shared memory (conflict-free 16-byte
// pieces of a 128-byte-swizzled [128 rows][64 col] tile, 4 wavefronts per 512 B) and a DMA
// thread moves whole tiles:
//     residual tile  --TMA load-->  staging set s  --threads: +acc,+bias,ReLU, in place-->
//     hi / lo tile   --TMA store--> global
// Two staging sets (hi + lo, 16 KB each) alternate per 64-column chunk; the residual of chunk
// q+2 is requested as soon as the store of chunk q has finished reading its set.
//
// Barriers per set: res_full (count 1: the DMA thread, plus TMA bytes) and out_ready (count
// 256: every epilogue thread after its shared-memory writes + proxy fence).
#pragma once
#include "tc_ptx.cuh"

namespace cs {

constexpr int kEpiWarps = 8;                 // warp % 4 = TMEM lane quadrant, 2 column halves
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr uint32_t kEpiTileBytes = 128 * 128;           // [128 rows][64 bf16], swizzle 128B
constexpr uint32_t kEpiSetBytes = 2 * kEpiTileBytes;    // hi + lo
constexpr uint32_t kEpiStagingBytes = 2 * kEpiSetBytes; // two sets

struct EpiArgs {
  const float* bias;  // indexed by output column
  bool res_hi, res_lo, out_hi, out_lo;
  float* out_f32;     // nullable: raw fp32 result written directly (diagnostics only)
  int relu;
};

// One 64-column chunk, this warp's 32 rows x 32 columns (column half h of the chunk).
//   stg      : shared address of the staging set (hi tile, lo tile right after it)
//   r        : row inside the 128-row tile (= TMEM lane)
//   taddr    : TMEM address of (lane quadrant, first of the 32 columns)
//   bias_col : output column of the first of the 32 columns
//   f32_off  : element offset for the diagnostic fp32 store (row-major), row_ok gates it
__device__ __forceinline__ void epi_chunk(const EpiArgs& e, uint32_t stg, int r, int h, uint32_t taddr,
                                          int bias_col, bool row_ok, int64_t f32_off) {
  const uint32_t row_addr = stg + (uint32_t)r * 128u;
  const uint32_t sw = (uint32_t)(r & 7);
  uint4 rh[4], rl[4];
  if (e.res_hi) {
#pragma unroll
    for (int j = 0; j < 4; ++j) rh[j] = lds128(row_addr + ((((uint32_t)(h * 4 + j)) ^ sw) << 4));
  }
  if (e.res_lo) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      rl[j] = lds128(row_addr + kEpiTileBytes + ((((uint32_t)(h * 4 + j)) ^ sw) << 4));
  }
  uint32_t acc[32];
  tmem_ld32(taddr, acc);
  tmem_ld_wait();
  float v[32];
  const float4* b4 = reinterpret_cast<const float4*>(e.bias + bias_col);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bb = __ldg(b4 + j);
    v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + bb.x;
    v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + bb.y;
    v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + bb.z;
    v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + bb.w;
  }
  if (e.res_hi) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[8 * j + 0] += bf16_lo_f(rh[j].x); v[8 * j + 1] += bf16_hi_f(rh[j].x);
      v[8 * j + 2] += bf16_lo_f(rh[j].y); v[8 * j + 3] += bf16_hi_f(rh[j].y);
      v[8 * j + 4] += bf16_lo_f(rh[j].z); v[8 * j + 5] += bf16_hi_f(rh[j].z);
      v[8 * j + 6] += bf16_lo_f(rh[j].w); v[8 * j + 7] += bf16_hi_f(rh[j].w);
    }
  }
  if (e.res_lo) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[8 * j + 0] += bf16_lo_f(rl[j].x); v[8 * j + 1] += bf16_hi_f(rl[j].x);
      v[8 * j + 2] += bf16_lo_f(rl[j].y); v[8 * j + 3] += bf16_hi_f(rl[j].y);
      v[8 * j + 4] += bf16_lo_f(rl[j].z); v[8 * j + 5] += bf16_hi_f(rl[j].z);
      v[8 * j + 6] += bf16_lo_f(rl[j].w); v[8 * j + 7] += bf16_hi_f(rl[j].w);
    }
  }
  if (e.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (e.out_f32 != nullptr && row_ok) {
    float4* of = reinterpret_cast<float4*>(e.out_f32 + f32_off);
#pragma unroll
    for (int j = 0; j < 8; ++j) of[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  uint32_t hi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  if (e.out_hi) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128v(row_addr + ((((uint32_t)(h * 4 + j)) ^ sw) << 4),
              make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]));
  }
  if (e.out_lo) {
    uint32_t lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      lo[j] = pack_bf16x2(v[2 * j] - bf16_lo_f(hi[j]), v[2 * j + 1] - bf16_hi_f(hi[j]));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128v(row_addr + kEpiTileBytes + ((((uint32_t)(h * 4 + j)) ^ sw) << 4),
              make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]));
  }
}

// Barrier addresses of the staging protocol (two sets; four in the generic GEMM kernel when a
// residual is loaded and the bf16-only stream leaves room, see epi_dma_loop).
constexpr int kEpiMaxSets = 4;
struct EpiBars {
  uint32_t res_full[kEpiMaxSets];
  uint32_t out_ready[kEpiMaxSets];
};

__device__ __forceinline__ void epi_bars_init(const EpiBars& b, uint32_t epi_threads = kEpiThreads,
                                              int nsets = 2) {
  for (int s = 0; s < nsets; ++s) {
    mbar_init(b.res_full[s], 1);
    mbar_init(b.out_ready[s], epi_threads);
  }
}

// DMA-thread loop over Q chunks.  load(q, set_addr, bar) issues the residual TMA loads of
// chunk q (or nothing), store(q, set_addr) the TMA stores.  res_bytes = bytes the loads of one
// chunk deliver (0: no residual, the set is handed over with a plain arrive).
//
// nsets = 4 (a ring): with two sets the residual of chunk q+2 can only be requested once the
// store of chunk q has read its set, so every chunk waits for a full TMA round trip (ncu r02w:
// the epilogue warps of the residual GEMMs spend 25 % of their samples on res_full, 78-80 %
// tensor pipe against 88-91 % without a residual).  With four sets the DMA thread lets the
// latest store stay in flight and hands over the set of the chunk before it: three residual
// tiles are on their way while one chunk is being worked on.
template <class LoadFn, class StoreFn>
__device__ __forceinline__ void epi_dma_loop(int64_t Q, uint32_t staging, uint32_t set_bytes,
                                             const EpiBars& bars,
                                             uint32_t res_bytes, bool any_store, LoadFn load,
                                             StoreFn store, int nsets = 2) {
  if (nsets == 4) {
    auto hand_over = [&](int64_t q) {
      const int s = (int)(q & 3);
      if (res_bytes) {
        mbar_expect_tx(bars.res_full[s], res_bytes);
        load(q, staging + s * set_bytes, bars.res_full[s]);
      } else {
        mbar_arrive(bars.res_full[s]);
      }
    };
    for (int64_t q = 0; q < 4 && q < Q; ++q) hand_over(q);
    for (int64_t q = 0; q < Q; ++q) {
      const int s = (int)(q & 3);
      mbar_wait(bars.out_ready[s], (uint32_t)((q >> 2) & 1));
      if (any_store) {
        store(q, staging + s * set_bytes);
        bulk_commit_group();
        bulk_wait_read_1();            // every store but the latest has read its set
        if (q >= 1 && q + 3 < Q) hand_over(q + 3);
      } else if (q + 4 < Q) {
        hand_over(q + 4);
      }
    }
    if (any_store) bulk_wait_all();
    return;
  }
  auto hand_over = [&](int64_t q) {
    const int s = (int)(q & 1);
    if (res_bytes) {
      mbar_expect_tx(bars.res_full[s], res_bytes);
      load(q, staging + s * set_bytes, bars.res_full[s]);
    } else {
      mbar_arrive(bars.res_full[s]);
    }
  };
  if (Q > 0) hand_over(0);
  if (Q > 1) hand_over(1);
  for (int64_t q = 0; q < Q; ++q) {
    const int s = (int)(q & 1);
    mbar_wait(bars.out_ready[s], (uint32_t)((q >> 1) & 1));
    if (any_store) {
      store(q, staging + s * set_bytes);
      bulk_commit_group();
    }
    if (q + 2 < Q) {
      if (any_store) bulk_wait_read_all();   // the set is free once the store has read it
      hand_over(q + 2);
    }
  }
  if (any_store) bulk_wait_all();
}

}  // namespace cs
