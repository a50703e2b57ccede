// K2c: the 8x8 x 64-channel stride-1 3x3 convolutions of layer 1 with the vertical taps summed
// in the epilogue ("y-sum" form).
//
// Reference: BasicBlock conv1 / conv2 of layer1 (model/resnet.py:19-23, 28-43).
//
// With 64 output channels an MMA is N = 64 wide and the stage is bound by operand reads from
// shared memory (~66 B/clk measured): every 128 x 16 A slice (4 KB) is fetched for 2 KB of B,
// once per tap -- 36 MMAs and 216 KB of operand reads per M tile in the halo kernel.  Here the
// three vertical taps of a kernel column share ONE MMA: the B tile of column dx stacks
// W[dy = 0..2][dx] along N (N = 192), so an A slice is read once for three taps:
//     D_dy[r][co] = sum_dx sum_ci  x[r shifted by dx][ci] * W[co][ci][dy][dx]      (12 MMAs, 120 KB)
//     out[y][x]   = D_0[y-1][x] + D_1[y][x] + D_2[y+1][x]                          (epilogue)
// The horizontal shift is still a TMA box fetched at x + dx - 1 (zero fill = padding).  Rows are
// in natural order r = 64*img + 8*y + x, so tcgen05.ld.16x256b hands a thread the four image
// rows of its TMEM lane quarter at one x: the vertical sum is register arithmetic, and only the
// rows y = 3 | 4 between the two lane quarters of an image are exchanged through shared memory
// (8 KB per tile).  Bias, residual, ReLU and the bf16 (hi / lo) split follow in place in the
// swizzled staging tile that the DMA warp moves with TMA, as in gemm_epilogue.cuh.
//
// Warp roles (one CTA per SM, persistent over M tiles of two instances):
//   warp 0 TMA producer   warp 1 TMEM alloc + MMA issue   warps 2 .. 2+EW-1 epilogue (TMEM lane
//   quarter = warp % 4, channel group = (warp - 2) / 4)   last warp: epilogue DMA
// EW = 16 epilogue warps (four per scheduler, 16 channels each) is the default: with EW = 8 a
// tile's epilogue (~450 dependent instructions per thread, two warps per scheduler) took ~2 400
// cycles against 1 152 cycles of MMA and bounded the kernel (46-48 % tensor pipe, ncu r01_n).
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "fwd.cuh"
#include "gemm_epilogue.cuh"

namespace cs {
namespace {

constexpr int kMaxStages = 6;                      // 6 when only bf16 tiles are staged, else 4
constexpr uint32_t kABytes = 128 * 128;            // one shifted box: 128 rows x 64 ch bf16
constexpr uint32_t kABoxBytes = 160 * 128;         // one-box form: 2 images x 8 rows x 10 pixels
constexpr uint32_t kBTileFull = 192 * 128;         // [dy*64 + co][64 ci] of one dx
constexpr uint32_t kBBytes = 3 * kBTileFull;       // 72 KB resident (36 KB used by a pair member)
constexpr uint32_t kXchBytes = 2 * 8 * 1024;       // 2 tile parities x 8 warps x 1 KB
// layout: [stages x A][B resident][2 staging sets][exchange][barriers]; stages * 16 KB + 2 sets
// is 128 KB either way (4 + 2 x 32 KB with hi/lo tiles, 6 + 2 x 16 KB with bf16 only)
constexpr uint32_t kOffB = 4 * kABytes + kEpiStagingBytes;        // start of the resident weights
constexpr uint32_t kOffXch = kOffB + kBBytes;
constexpr uint32_t kOffBars = kOffXch + kXchBytes;
constexpr uint32_t kSmemBytes = kOffBars + 256 + 1024;
constexpr uint32_t kAccCols = 256;                 // accumulator stage pitch in TMEM columns
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void named_bar(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <int NI>
__device__ __forceinline__ void tmem_ld_16x256b(uint32_t taddr, uint32_t (&r)[4 * NI]);
template <>
__device__ __forceinline__ void tmem_ld_16x256b<4>(uint32_t taddr, uint32_t (&r)[16]) {
  tmem_ld_16x256b_x4(taddr, r);
}
template <>
__device__ __forceinline__ void tmem_ld_16x256b<2>(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr));
}

// CL = 2: CTA pairs (tcgen05 cta_group::2) as in conv_gemm.cu -- neighbouring M tiles, each CTA
// holds its own A boxes and 96 of the 192 rows of every weight tile, the leader issues M = 256
// MMAs for both.
// OB = true ("one box"): instead of three x-shifted 128-row boxes per M tile (48 KB written into
// shared memory by TMA), ONE box of 10 pixels per image row {64 ch, x = -1..8, 8 rows, 2 images}
// = 160 rows = 20 KB; the three horizontal taps are the same tile read through descriptors whose
// start address moves by one 128-byte row and whose 8-row groups are 10 rows (1280 B) apart.
// That is legal because tcgen05 (like TMA) applies the 128-byte swizzle XOR on absolute
// shared-memory address bits (tools/umma_probe.cu, "sw128 start128 sbo1280").
template <int CL, int EW, bool OB>
__global__ void __launch_bounds__((EW + 3) * 32, 1)
conv_ysum_kernel(const __grid_constant__ YsumParams p) {
  constexpr int NI = 32 / EW;                     // 8-channel groups per epilogue thread: 4 | 2
  constexpr uint32_t kStageBytes = OB ? kABoxBytes : kABytes;
  constexpr uint32_t kLoadsPerTile = OB ? 1 : 3;
  constexpr int CW = 8 * NI;                      // channels per epilogue warp: 32 | 16
  constexpr uint32_t kBTile = kBTileFull / CL;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const bool lo_tiles = p.res_lo != nullptr || p.out_lo != nullptr;
  const int kStages = OB ? (lo_tiles ? 3 : 4) : (lo_tiles ? 4 : kMaxStages);
  const uint32_t set_bytes = lo_tiles ? kEpiSetBytes : kEpiTileBytes;
  const uint32_t bres = base + kOffB;
  const uint32_t staging = base + kStages * kStageBytes;
  const uint32_t bar_base = base + kOffBars;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t bres_bar = bar_base + 8u * (2 * kMaxStages + 4);
  EpiBars ebars;
  for (int s = 0; s < 2; ++s) {
    ebars.res_full[s] = bar_base + 8u * (2 * kMaxStages + 5 + s);
    ebars.out_ready[s] = bar_base + 8u * (2 * kMaxStages + 7 + s);
  }
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(base_ptr + kOffBars + 8 * (2 * kMaxStages + 9));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  // a cluster walks groups of CL neighbouring M tiles; CTA `rank` owns tile group * CL + rank
  const int m_groups = (p.num_m_tiles + CL - 1) / CL;
  const int first = blockIdx.x / CL, step_g = gridDim.x / CL;
  auto tile_of = [&](int gi) { return p.tile_base + (p.reverse ? m_groups - 1 - gi : gi) * CL + rank; };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CL == 1 ? EW * 32 : CL * EW);
    }
    mbar_init(bres_bar, 1);
    epi_bars_init(ebars, EW * 32);
    fence_barrier_init();
    prefetch_tmap(OB ? &p.a_box_map : &p.a_map);
    prefetch_tmap(&p.b_map);
  }
  if (warp == 1) {
    if (CL == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    else tmem_alloc_pair(smem_u32((const void*)tmem_slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // the leader's barriers exist before the peer's loads signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // weights are constants: fetch them while the previous layer still drains
      if (rank == 0) mbar_expect_tx(bres_bar, CL * 3 * kBTile);
      for (int dx = 0; dx < 3; ++dx) {
        if (CL == 1) tma_load_2d(bres + dx * kBTile, &p.b_map, bres_bar, 0, dx * 192);
        else tma_load_2d_pair(bres + dx * kBTile, &p.b_map, bres_bar, 0, dx * 192 + rank * (192 / CL));
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int gi = first; gi < m_groups; gi += step_g) {
        const int m_tile = tile_of(gi);
        for (int dx = 0; dx < (int)kLoadsPerTile; ++dx) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(stage), CL * kStageBytes);
          const CUtensorMap* amap = OB ? &p.a_box_map : &p.a_map;
          if (CL == 1)
            tma_load_4d(base + stage * kStageBytes, amap, full_bar(stage), 0, dx - 1, 0, m_tile * 2);
          else
            tma_load_4d_pair(base + stage * kStageBytes, amap, full_bar(stage), 0, dx - 1, 0, m_tile * 2);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop and waits; one elected lane issues (tc_ptx.cuh: elect_one_sync
    // keeps the issue sequence on the uniform datapath).
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CL, 192);
      mbar_wait(bres_bar, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int gi = first; gi < m_groups; gi += step_g) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          if (!OB || dx == 0) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
          }
          if (elect_one_sync()) {
            const uint64_t a_desc = OB ? umma_desc_sw128_sbo(base + stage * kStageBytes + dx * 128u, 1280u)
                                       : umma_desc_sw128(base + stage * kStageBytes);
            const uint64_t b_desc = umma_desc_sw128(bres + dx * kBTile);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CL == 1)
                umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                          (dx > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                               (dx > 0 || k > 0) ? 1u : 0u);
            }
            if (!OB || dx == 2) {
              if (CL == 1) umma_commit(empty_bar(stage));
              else umma_commit_pair(empty_bar(stage), kMask);
            }
            if (dx == 2) {
              if (CL == 1) umma_commit(tfull_bar(acc));
              else umma_commit_pair(tfull_bar(acc), kMask);
            }
          }
          __syncwarp();
          if (!OB || dx == 2) {
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 2 + EW) {
    pdl_wait();
    const int quad = warp & 3;                 // TMEM lanes 32*quad ..: image quad/2, rows y0 .. y0+3
    const int cg = (warp - 2) >> 2;            // channels CW*cg .. CW*cg + CW-1
    const int g = lane >> 2, t = lane & 3;     // x, channel-pair slot
    const bool upper = (quad & 1) == 0;        // owns image rows 0..3 (else 4..7)
    const int pair_bar = 1 + (quad >> 1) * (EW / 4) + cg;     // named barrier of the two quarters
    const int partner = (warp - 2) ^ 1;                        // same image, same channel group
    float bias_r[NI][2];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      bias_r[i][0] = p.bias[cg * CW + 8 * i + 2 * t];
      bias_r[i][1] = p.bias[cg * CW + 8 * i + 2 * t + 1];
    }
    const bool rh = p.res_hi != nullptr, rl = p.res_lo != nullptr;
    const bool oh = p.out_hi != nullptr, ol = p.out_lo != nullptr;
    // Everything that does not depend on the tile is worked out once: the epilogue issues ~440
    // instructions per tile and warp, 16 warps deep -- more issue slots than the tile's 1 152
    // cycles of MMA offer (ncu r02: 49-58 % tensor pipe), and 160 of them were address arithmetic.
    uint32_t off[4][NI];                       // byte offset of (row yy, channel group i) in a staging tile
#pragma unroll
    for (int yy = 0; yy < 4; ++yy)
#pragma unroll
      for (int i = 0; i < NI; ++i)
        off[yy][i] = (uint32_t)(quad * 32 + yy * 8 + g) * 128u + 4u * t + ((((uint32_t)(cg * NI + i)) ^ (uint32_t)g) << 4);
    constexpr int kXw = 64 * NI;               // floats per warp slot (kXchBytes covers 2 x EW slots)
    float* const xch = reinterpret_cast<float*>(base_ptr + kOffXch);
    float* const xw0 = xch + (warp - 2) * kXw + lane * 2 * NI;
    const float* const xr0 = xch + partner * kXw + lane * 2 * NI;
    float2 bias2[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) bias2[i] = make_float2(bias_r[i][0], bias_r[i][1]);
    const bool fast_cfg = !rl && !ol && oh && p.relu != 0 && p.out_f32 == nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    int64_t q = 0;
    for (int gi = first; gi < m_groups; gi += step_g, ++q) {
      const int m_tile = tile_of(gi);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      // R[dy][w][4i + 2h + e]: image row y0 + 2w + h, column x = g, channel CW*cg + 8i + 2t + e
      uint32_t R[3][2][4 * NI];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int w = 0; w < 2; ++w)
          tmem_ld_16x256b<NI>(tmem_base + ((uint32_t)(quad * 32 + w * 16) << 16) +
                                  (uint32_t)(acc * kAccCols + dy * 64 + cg * CW),
                              R[dy][w]);
      tmem_ld_wait();
      tc_fence_before();
      if (CL == 1) {
        mbar_arrive(tempty_bar(acc));
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      // the channel pair (e = 0, 1) of one (tap row, image row, channel group): packed fp32 adds
      auto D2 = [&](int dy, int yy, int i) -> float2 {
        return make_float2(__uint_as_float(R[dy][yy >> 1][4 * i + 2 * (yy & 1)]),
                           __uint_as_float(R[dy][yy >> 1][4 * i + 2 * (yy & 1) + 1]));
      };
      // rows 3 | 4 of an image live in different lane quarters: exchange them (fp32, 2*NI per thread)
      const int par = (int)(q & 1);
      float* xw = xw0 + par * EW * kXw;
      const float* xr = xr0 + par * EW * kXw;
#pragma unroll
      for (int i = 0; i < NI; i += 2) {
        // D_0 of row 3 feeds row 4 of the lower quarter; D_2 of row 4 feeds row 3 of the upper one
        const float2 a = upper ? D2(0, 3, i) : D2(2, 0, i), b = upper ? D2(0, 3, i + 1) : D2(2, 0, i + 1);
        reinterpret_cast<float4*>(xw)[i >> 1] = make_float4(a.x, a.y, b.x, b.y);
      }
      named_bar(pair_bar, 64);
      float2 imp[NI];
#pragma unroll
      for (int i = 0; i < NI; i += 2) {
        const float4 v4 = reinterpret_cast<const float4*>(xr)[i >> 1];
        imp[i] = make_float2(v4.x, v4.y);
        imp[i + 1] = make_float2(v4.z, v4.w);
      }

      mbar_wait(ebars.res_full[par], (uint32_t)((q >> 1) & 1));
      const uint32_t stg = staging + par * set_bytes;
      const float2 zero2 = make_float2(0.f, 0.f);
      // kFast: the two configurations the forward runs (bf16 stream, ReLU, with or without a
      // residual) with the flags as compile-time constants -- tested per element, the five
      // runtime flags cost ~10 instructions each time they were re-materialised
      auto emit = [&](auto rh_c, auto fast_c) {
        constexpr bool kRH = decltype(rh_c)::value, kFast = decltype(fast_c)::value;
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            // D_0 of row y-1 + D_1 of row y + D_2 of row y+1 + bias, in that order
            const float2 above = yy > 0 ? D2(0, yy - 1, i) : (upper ? zero2 : imp[i]);
            const float2 below = yy < 3 ? D2(2, yy + 1, i) : (upper ? imp[i] : zero2);
            float2 v = __fadd2_rn(__fadd2_rn(__fadd2_rn(above, D2(1, yy, i)), below), bias2[i]);
            const uint32_t addr = stg + off[yy][i];
            if (kFast) {
              if (kRH) { const uint32_t r2 = lds32(addr); v = __fadd2_rn(v, make_float2(bf16_lo_f(r2), bf16_hi_f(r2))); }
              sts32(addr, pack_bf16x2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)));
            } else {
              if (rh) { const uint32_t r2 = lds32(addr); v = __fadd2_rn(v, make_float2(bf16_lo_f(r2), bf16_hi_f(r2))); }
              if (rl) { const uint32_t r2 = lds32(addr + kEpiTileBytes); v = __fadd2_rn(v, make_float2(bf16_lo_f(r2), bf16_hi_f(r2))); }
              if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
              if (p.out_f32 != nullptr) {
                const int row = quad * 32 + yy * 8 + g;      // row of the 128-row tile
                const int64_t grow = (int64_t)m_tile * 128 + row;
                if (grow < p.n_inst * 64)
                  *reinterpret_cast<float2*>(p.out_f32 + grow * 64 + cg * CW + 8 * i + 2 * t) = v;
              }
              const uint32_t hi = pack_bf16x2(v.x, v.y);
              if (oh) sts32(addr, hi);
              if (ol) sts32(addr + kEpiTileBytes, pack_bf16x2(v.x - bf16_lo_f(hi), v.y - bf16_hi_f(hi)));
            }
          }
        }
      };
      if (fast_cfg) {
        if (rh) emit(std::true_type{}, std::true_type{});
        else emit(std::false_type{}, std::true_type{});
      } else {
        emit(std::false_type{}, std::false_type{});
      }
      fence_async_shared();
      mbar_arrive(ebars.out_ready[par]);
    }
  } else if (lane == 0) {
    pdl_wait();
    int n_items = 0;
    for (int gi = first; gi < m_groups; gi += step_g) ++n_items;
    auto row_of = [&](int64_t q) { return tile_of(first + (int)q * step_g) * 128; };
    const bool rh = p.res_hi != nullptr, rl = p.res_lo != nullptr;
    const bool oh = p.out_hi != nullptr, ol = p.out_lo != nullptr;
    epi_dma_loop(
        (int64_t)n_items, staging, set_bytes, ebars,
        (rh ? kEpiTileBytes : 0u) + (rl ? kEpiTileBytes : 0u), oh || ol,
        [&](int64_t q, uint32_t set, uint32_t bar) {
          if (rh) tma_load_2d(set, &p.res_hi_map, bar, 0, row_of(q));
          if (rl) tma_load_2d(set + kEpiTileBytes, &p.res_lo_map, bar, 0, row_of(q));
        },
        [&](int64_t q, uint32_t set) {
          if (oh) tma_store_2d(&p.out_hi_map, set, 0, row_of(q));
          if (ol) tma_store_2d(&p.out_lo_map, set + kEpiTileBytes, 0, row_of(q));
        });
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while the pair's MMAs may still read its tiles
  if (warp == 1) {
    tc_fence_after();
    if (CL == 1) tmem_dealloc(tmem_base, 512);
    else tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

bool ysum_supported(int W, int Cin, int Cout) { return W == 8 && Cin == 64 && Cout == 64; }

// Host: [dx*192 + dy*64 + co][ci] bf16 from folded OIHW fp32 weights [64][64][3][3].
void pack_ysum_weights(const float* w_oihw, uint16_t* out) {
  auto rn = [](float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
  };
  for (int dx = 0; dx < 3; ++dx)
    for (int dy = 0; dy < 3; ++dy)
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 64; ++ci)
          out[((dx * 192) + dy * 64 + co) * 64 + ci] = rn(w_oihw[((co * 64 + ci) * 3 + dy) * 3 + dx]);
}

template <int CL, int EW, bool OB>
int launch_ysum(const YsumParams& p, cudaStream_t st) {
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(conv_ysum_kernel<CL, EW, OB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kSmemBytes));
    if (dev < 64) attr_done[dev] = true;
  }
  const int sms = num_sms();
  const int groups = (p.num_m_tiles + CL - 1) / CL;
  const int clusters = groups < sms / CL ? groups : sms / CL;
  CS_CUDA(launch_pdl(conv_ysum_kernel<CL, EW, OB>, dim3((unsigned)(clusters * CL)), dim3((EW + 3) * 32), kSmemBytes,
                     st, CL, p));
  return CS_OK;
}

// CELLSEG_YSUM_EPI=8 restores the two-warps-per-scheduler epilogue of round 1.
const bool g_epi8 = []() {
  const char* e = getenv("CELLSEG_YSUM_EPI");
  return e != nullptr && strcmp(e, "8") == 0;
}();

// One 10-pixel-wide box per M tile (default) instead of three shifted boxes: 20 instead of 48 KB
// written into shared memory per tile and 58 % less L2 -> SM traffic for the activations;
// 10.25 -> 10.38 M instances/s on one box (gpurun r2k, two runs each).  CELLSEG_YSUM_BOX=0
// restores the three-box form.
const bool g_one_box = []() {
  const char* e = getenv("CELLSEG_YSUM_BOX");
  return !(e != nullptr && strcmp(e, "0") == 0);
}();

int launch_conv_ysum(const YsumParams& p, cudaStream_t st) {
  if (p.num_m_tiles <= 0) return CS_OK;
  if (g_epi8) return p.cluster > 1 ? launch_ysum<2, 8, false>(p, st) : launch_ysum<1, 8, false>(p, st);
  if (g_one_box) return p.cluster > 1 ? launch_ysum<2, 16, true>(p, st) : launch_ysum<1, 16, true>(p, st);
  return p.cluster > 1 ? launch_ysum<2, 16, false>(p, st) : launch_ysum<1, 16, false>(p, st);
}

}  // namespace cs
