// K2b: stride-1 3x3 convolutions of the 8x8 (64 ch) and 4x4 (128 ch) stages with y-halo tiles.
//
// Reference: BasicBlock conv1 / conv2 of layer1 and layer2 (model/resnet.py:19-23, 28-43).
//
// The generic kernel (conv_gemm.cu) fetches one shifted 128-row box per tap: 9 L2 reads of the
// same activations and 9 reads of the weights per M tile, and the launch list shows those
// layers pinned at the L2->SM bandwidth.  Here a K step fetches ONE box per horizontal shift
// dx that also carries a one-pixel halo in y:
//     tensor map dims {C, W, T, H}  (instances before rows!)   box {64, W, IMG, H + 2} at
//     (c0, dx - 1, t0, -1)  ->  shared rows ordered  r = x + W * (img + IMG * y'),  y' = y + 1
// so the three vertical taps dy are the SAME tile read from a start address moved by
// dy * W * IMG rows = dy * 2 KB (8x8) or dy * 4 KB (4x4): whole 1024-byte swizzle atoms,
// i.e. just another UMMA descriptor.  TMA's out-of-bounds zero fill provides the padding in
// x (shifted start) and y (halo rows -1 and H).  A traffic per M tile drops from 9 to 3.75
// (8x8) / 4.5 (4x4) tiles; with 64 channels the 72 KB of weights stay resident in shared
// memory for the whole kernel.  M row r of the accumulator is pixel (oy = r / (W*IMG),
// img = (r / W) % IMG, ox = r % W); the epilogue maps it back to [instance][y][x][c].
//
// CTA pairs (CL = 2, tcgen05 cta_group::2) as in conv_gemm.cu: neighbouring M tiles, each CTA
// holds its own halo tile and HALF of every weight tile (the resident set shrinks to 36 KB),
// the leader issues M = 256 MMAs for both.
//
// DS = true adds the BasicBlock's 1x1 stride-2 downsample of the block input as one more K step
// (layer-2 entry, model/resnet.py:36-40 with downsample): a {64 ch, W, IMG, H} box of the
// block input's even pixels, already in accumulator row order, meets the weight columns behind
// the nine taps.  That convolution ran in the generic shifted-box kernel before (nine boxes per
// M tile): 220 us against 138 us for the same 3x3 without the shortcut (ncu r2a).
//
// Warp roles as in conv_gemm.cu: warp 0 TMA, warp 1 MMA issue, warps 2-9 epilogue, warp 10 DMA.
#include "fwd.cuh"
#include "gemm_epilogue.cuh"

namespace cs {
namespace {

constexpr int kHaloThreads = 352;  // TMA warp, MMA warp, 8 epilogue warps, epilogue DMA warp

template <int BN, int W, int CCH, bool BRES, int CL>
struct HaloCfg {
  static constexpr int kImg = 128 / (W * W);            // instances per M tile
  static constexpr int kRowsY = W * kImg;               // accumulator rows per output y
  static constexpr int kHaloRows = kRowsY * (W + 2);
  static constexpr uint32_t kABytes = kHaloRows * 128;  // 20 KB (8x8) / 24 KB (4x4)
  static constexpr uint32_t kBTile = (BN / CL) * 128;
  static constexpr uint32_t kStageBytes = kABytes + (BRES ? 0 : 3 * kBTile);
  static constexpr int kStages = BRES ? (CL == 2 ? 6 : 4) : (CL == 2 ? 3 : 2);   // what fits 227 KB
  static constexpr uint32_t kBResBytes = BRES ? 9 * CCH * kBTile : 0;
  static constexpr uint32_t kStagingOffset = kStages * kStageBytes + kBResBytes;
  static constexpr uint32_t kBarOffset = kStagingOffset + kEpiStagingBytes;
  static constexpr int kChunks = BN / 64;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256 + 1024;
  static constexpr int kSteps = 3 * CCH;                // (dx, channel chunk)
  static constexpr uint32_t kTmemCols = 2 * BN;
  static_assert(kABytes % 1024 == 0 && kStageBytes % 1024 == 0, "stages must keep 1 KB alignment");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert((kRowsY * 128) % 1024 == 0, "a dy shift must move whole swizzle atoms");
};

template <int BN, int W, int CCH, bool BRES, int CL, bool DS>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ HaloParams p) {
  using Cfg = HaloCfg<BN, W, CCH, BRES, CL>;
  constexpr int Cin = CCH * 64;
  static_assert(!DS || !BRES, "the downsample step loads its weight tile with the stage");
  constexpr uint32_t kDsBytes = 128 * 128 + Cfg::kBTile;   // one 128-row box + one weight tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bres = base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = base + Cfg::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t bres_bar = bar_base + 8u * (2 * Cfg::kStages + 4);
  EpiBars ebars;
  for (int s = 0; s < 2; ++s) {
    ebars.res_full[s] = bar_base + 8u * (2 * Cfg::kStages + 5 + s);
    ebars.out_ready[s] = bar_base + 8u * (2 * Cfg::kStages + 7 + s);
  }
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::kBarOffset + 8 * (2 * Cfg::kStages + 9));
  const uint32_t staging = base + Cfg::kStagingOffset;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  // a cluster walks groups of CL neighbouring M tiles; CTA `rank` owns tile group * CL + rank
  const int m_groups = (p.num_m_tiles + CL - 1) / CL;
  const int first = blockIdx.x / CL, step_g = gridDim.x / CL;
  auto tile_of = [&](int gi) { return (p.reverse ? m_groups - 1 - gi : gi) * CL + rank; };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CL == 1 ? kEpiThreads : CL * kEpiWarps);
    }
    mbar_init(bres_bar, 1);
    epi_bars_init(ebars);
    fence_barrier_init();
    prefetch_tmap(&p.a_map);
    prefetch_tmap(&p.b_map);
    if (DS) prefetch_tmap(&p.ds_map);
  }
  if (warp == 1) {
    if (CL == 1) tmem_alloc(smem_u32((const void*)tmem_slot), Cfg::kTmemCols);
    else tmem_alloc_pair(smem_u32((const void*)tmem_slot), Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // the leader's barriers exist before the peer's loads signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      const int b_row = rank * (BN / CL);   // this CTA's rows of every weight tile
      if (BRES) {   // weights are constants: fetch them while the previous layer still drains
        if (rank == 0) mbar_expect_tx(bres_bar, CL * Cfg::kBResBytes);
        for (int t = 0; t < 9 * CCH; ++t) {   // tile index = tap * CCH + chunk = K column / 64
          if (CL == 1) tma_load_2d(bres + t * Cfg::kBTile, &p.b_map, bres_bar, t * 64, 0);
          else tma_load_2d_pair(bres + t * Cfg::kBTile, &p.b_map, bres_bar, t * 64, b_row);
        }
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int gi = first; gi < m_groups; gi += step_g) {
        const int m_tile = tile_of(gi);
        for (int s = 0; s < Cfg::kSteps; ++s) {
          const int dx = s / CCH, ch = s - dx * CCH;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          if (rank == 0) mbar_expect_tx(full_bar(stage), CL * Cfg::kStageBytes);
          if (CL == 1)
            tma_load_4d(a_dst, &p.a_map, full_bar(stage), ch * 64, dx - 1, m_tile * Cfg::kImg, -1);
          else
            tma_load_4d_pair(a_dst, &p.a_map, full_bar(stage), ch * 64, dx - 1, m_tile * Cfg::kImg, -1);
          if (!BRES) {
            for (int dy = 0; dy < 3; ++dy) {
              if (CL == 1)
                tma_load_2d(a_dst + Cfg::kABytes + dy * Cfg::kBTile, &p.b_map, full_bar(stage),
                            (dy * 3 + dx) * Cin + ch * 64, 0);
              else
                tma_load_2d_pair(a_dst + Cfg::kABytes + dy * Cfg::kBTile, &p.b_map, full_bar(stage),
                                 (dy * 3 + dx) * Cin + ch * 64, b_row);
            }
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        if (DS) {   // shortcut: even pixels of the block input + the weight columns behind the taps
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          if (rank == 0) mbar_expect_tx(full_bar(stage), CL * kDsBytes);
          if (CL == 1) {
            tma_load_4d(a_dst, &p.ds_map, full_bar(stage), 0, 0, m_tile * Cfg::kImg, 0);
            tma_load_2d(a_dst + Cfg::kABytes, &p.b_map, full_bar(stage), 9 * Cin, 0);
          } else {
            tma_load_4d_pair(a_dst, &p.ds_map, full_bar(stage), 0, 0, m_tile * Cfg::kImg, 0);
            tma_load_2d_pair(a_dst + Cfg::kABytes, &p.b_map, full_bar(stage), 9 * Cin, b_row);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CL, BN);
      if (BRES) mbar_wait(bres_bar, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int gi = first; gi < m_groups; gi += step_g) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int s = 0; s < Cfg::kSteps; ++s) {
          const int dx = s / CCH, ch = s - dx * CCH;
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = base + stage * Cfg::kStageBytes;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t a_desc = umma_desc_sw128(a_src + dy * (Cfg::kRowsY * 128));
            const uint32_t b_src = BRES ? bres + ((dy * 3 + dx) * CCH + ch) * Cfg::kBTile
                                        : a_src + Cfg::kABytes + dy * Cfg::kBTile;
            const uint64_t b_desc = umma_desc_sw128(b_src);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CL == 1)
                umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                          (s > 0 || dy > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                               (s > 0 || dy > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (CL == 1) umma_commit(empty_bar(stage));
          else umma_commit_pair(empty_bar(stage), kMask);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        if (DS) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = base + stage * Cfg::kStageBytes;
          const uint64_t a_desc = umma_desc_sw128(a_src);
          const uint64_t b_desc = umma_desc_sw128(a_src + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (CL == 1) umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, 1u);
            else umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, 1u);
          }
          if (CL == 1) umma_commit(empty_bar(stage));
          else umma_commit_pair(empty_bar(stage), kMask);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        if (CL == 1) umma_commit(tfull_bar(acc));
        else umma_commit_pair(tfull_bar(acc), kMask);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    pdl_wait();
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int oy = r / Cfg::kRowsY, img = (r % Cfg::kRowsY) / W, ox = r % W;
    const EpiArgs ea{p.bias, p.res_hi != nullptr, p.res_lo != nullptr, p.out_hi != nullptr,
                     p.out_lo != nullptr, p.out_f32, p.relu};
    int acc = 0;
    uint32_t acc_phase = 0;
    int64_t q = 0;
    for (int gi = first; gi < m_groups; gi += step_g) {
      const int m_tile = tile_of(gi);
      const int64_t inst = (int64_t)m_tile * Cfg::kImg + img;
      const int64_t row = inst * (W * W) + oy * W + ox;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < Cfg::kChunks; ++c, ++q) {
        const int s = (int)(q & 1);
        mbar_wait(ebars.res_full[s], (uint32_t)((q >> 1) & 1));
        const int col = c * 64 + half * 32;
        epi_chunk(ea, staging + s * kEpiSetBytes, r, half,
                  tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col), col,
                  inst < p.n_inst, row * (int64_t)BN + col);
        fence_async_shared();
        mbar_arrive(ebars.out_ready[s]);
      }
      tc_fence_before();
      if (CL == 1) {
        mbar_arrive(tempty_bar(acc));
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (lane == 0) {
    pdl_wait();
    // epilogue DMA thread: tiles move as {64 ch, W, IMG, H} boxes of the {C, W, T, H} maps, i.e.
    // in accumulator row order
    int n_items = 0;
    for (int gi = first; gi < m_groups; gi += step_g) ++n_items;
    auto coords = [&](int64_t q, int* c0, int* t0) {
      const int item = (int)(q / Cfg::kChunks), c = (int)(q % Cfg::kChunks);
      const int m_tile = tile_of(first + item * step_g);
      *c0 = c * 64;
      *t0 = m_tile * Cfg::kImg;
    };
    const bool rh = p.res_hi != nullptr, rl = p.res_lo != nullptr;
    const bool oh = p.out_hi != nullptr, ol = p.out_lo != nullptr;
    epi_dma_loop(
        (int64_t)n_items * Cfg::kChunks, staging, kEpiSetBytes, ebars,
        (rh ? kEpiTileBytes : 0u) + (rl ? kEpiTileBytes : 0u), oh || ol,
        [&](int64_t q, uint32_t set, uint32_t bar) {
          int c0, t0;
          coords(q, &c0, &t0);
          if (rh) tma_load_4d(set, &p.res_hi_map, bar, c0, 0, t0, 0);
          if (rl) tma_load_4d(set + kEpiTileBytes, &p.res_lo_map, bar, c0, 0, t0, 0);
        },
        [&](int64_t q, uint32_t set) {
          int c0, t0;
          coords(q, &c0, &t0);
          if (oh) tma_store_4d(&p.out_hi_map, set, c0, 0, t0, 0);
          if (ol) tma_store_4d(&p.out_lo_map, set + kEpiTileBytes, c0, 0, t0, 0);
        });
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while the pair's MMAs may still read its tiles
  if (warp == 1) {
    tc_fence_after();
    if (CL == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int W, int CCH, bool BRES, int CL, bool DS = false>
int launch_halo(const HaloParams& p, cudaStream_t st) {
  using Cfg = HaloCfg<BN, W, CCH, BRES, CL>;
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, W, CCH, BRES, CL, DS>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
    if (dev < 64) attr_done[dev] = true;
  }
  const int groups = (p.num_m_tiles + CL - 1) / CL;
  const int clusters = groups < num_sms() / CL ? groups : num_sms() / CL;
  CS_CUDA(launch_pdl(conv_halo_kernel<BN, W, CCH, BRES, CL, DS>, dim3((unsigned)(clusters * CL)),
                     dim3(kHaloThreads), Cfg::kSmemBytes, st, CL, p));
  return CS_OK;
}

}  // namespace

bool halo_supported(int W, int Cin, int Cout) {
  return (W == 8 && Cin == 64 && Cout == 64) || (W == 4 && Cin == 128 && Cout == 128);
}

int launch_conv_halo(const HaloParams& p, int W, int Cin, cudaStream_t st) {
  if (p.num_m_tiles <= 0) return CS_OK;
  if (p.has_ds) {
    if (W == 4 && Cin == 128)
      return p.cluster > 1 ? launch_halo<128, 4, 2, false, 2, true>(p, st) : launch_halo<128, 4, 2, false, 1, true>(p, st);
    set_error("launch_conv_halo: fused downsample only for the 4x4 x 128 stage");
    return CS_ERR_UNSUPPORTED;
  }
  if (p.cluster > 1) {
    if (W == 8 && Cin == 64) return launch_halo<64, 8, 1, true, 2>(p, st);
    if (W == 4 && Cin == 128) return launch_halo<128, 4, 2, false, 2>(p, st);
  } else {
    if (W == 8 && Cin == 64) return launch_halo<64, 8, 1, true, 1>(p, st);
    if (W == 4 && Cin == 128) return launch_halo<128, 4, 2, false, 1>(p, st);
  }
  set_error("launch_conv_halo: unsupported geometry W=%d Cin=%d", W, Cin);
  return CS_ERR_UNSUPPORTED;
}

// View of the even pixels of a [T][2H][2W][C] tensor as {C, W, T, H} with a {64, W, IMG, H} box:
// the input of a 1x1 stride-2 downsample in accumulator row order.
int make_act_map_halo_ds(CUtensorMap* map, const void* base, int C, int W, int H, int64_t T) {
  // same encoder call as make_act_map_halo; pixel strides doubled, no halo
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !ptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return CS_ERR_CUDA;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ptr);
  const int img = 128 / (W * H);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)T, (cuuint64_t)H};
  // byte strides of dims 1..3: x (2 pixels), instance, y (2 rows of 2W pixels)
  cuuint64_t strides[3] = {(cuuint64_t)2 * C * 2, (cuuint64_t)(2 * H) * (2 * W) * C * 2, (cuuint64_t)2 * (2 * W) * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)img, (cuuint32_t)H};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(halo ds C=%d W=%d H=%d T=%lld) failed: %d", C, W, H, (long long)T, (int)r);
    return CS_ERR_CUDA;
  }
  return CS_OK;
}

// {C, W, T, H} map with a {64, W, IMG, H + 2} box: instances sit between x and y in the box
// order so that a vertical tap is a whole-atom shift of the shared-memory tile.
int make_act_map_halo(CUtensorMap* map, const void* base, int C, int W, int H, int64_t T,
                      int halo) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !ptr) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return CS_ERR_CUDA;
    }
    enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  const int img = 128 / (W * H);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)T, (cuuint64_t)H};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)img, (cuuint32_t)(H + 2 * halo)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(halo C=%d W=%d H=%d T=%lld) failed: %d", C, W, H, (long long)T,
              (int)r);
    return CS_ERR_CUDA;
  }
  return CS_OK;
}

}  // namespace cs
