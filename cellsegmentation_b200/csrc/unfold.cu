// K1: tile unfolding + ToTensor + Normalize, materialised as fp32 NCHW tiles.
//
// Reference: LystoTestset.__getitem__ mode "tile" (dataset/dataset.py:409-416),
// LystoDataset.__getitem__ modes 1 and 3 (:206-214, :244-251), transform :78-83:
//   tile = images[bag][x:x+S, y:y+S]                      (u8, HWC)
//   ToTensor  : CHW, float32(u) / 255
//   Normalize : (t - mean_c) / std_c, mean/std float32
// Every channel value is one of 256 inputs, so the transform is a 3x256 fp32
// look-up table built once on the host with the same fp32 operations
// (bit-exact against torchvision; checked in tests/).
//
// Algorithmic bytes per instance: 3*S*S read (L2-resident: tiles of one bag
// overlap) + 12*S*S written.  The fused bf16 forward (stem_win.cu) never
// materialises this tensor; this kernel is the drop-in for __getitem__ batches and
// the input of the fp32 parity path.
#include "common.cuh"

namespace {

__constant__ float c_norm_lut[3 * 256];
bool g_lut_ready[64] = {false};

struct TileSrc {
  const uint8_t* img;  // [n_bags][H][W][3]
  int H, W, tile, interval;
  int grid_w;          // grid positions along W
  int64_t tiles_per_bag;
};

// One warp owns `rows` consecutive rows of one tile (work item); lane = x (stepping by 32).  A row
// is 3*S contiguous source bytes and becomes three S-float rows of the NCHW tile: coalesced byte
// loads, coalesced 4-byte stores.  The 3x256 LUT is replicated once per shared-memory bank
// (lut32[v][lane], 96 KB) so the 32 data-dependent look-ups of a warp never conflict: the first
// version (one float4 per thread, unreplicated LUT, per-element index divisions) was bound by
// bank conflicts and integer math at 19 % of the HBM write roofline (ncu r01_p).
constexpr int kLutCopies = 32;
constexpr int kUnfoldSmem = 3 * 256 * kLutCopies * 4;

template <bool kGather>
__global__ void __launch_bounds__(256)
unfold_kernel(TileSrc src, int64_t inst_begin, int64_t inst_count, const int32_t* __restrict__ gbag,
              const int32_t* __restrict__ gx, const int32_t* __restrict__ gy,
              float* __restrict__ out, int rows) {
  extern __shared__ float lut32[];
  for (int i = threadIdx.x; i < 3 * 256 * kLutCopies; i += blockDim.x) lut32[i] = c_norm_lut[i / kLutCopies];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int S = src.tile;
  const int items_per_inst = (S + rows - 1) / rows;
  const int64_t n_items = inst_count * items_per_inst;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t plane = (int64_t)S * S;
  const float* l0 = lut32 + lane;
  for (int64_t item = warp0; item < n_items; item += n_warps) {
    const int64_t j = item / items_per_inst;
    const int r0 = (int)(item - j * items_per_inst) * rows;
    const int r1 = r0 + rows < S ? r0 + rows : S;
    int64_t bag;
    int row0, col0;
    if (kGather) {
      bag = gbag[j];
      row0 = gx[j];
      col0 = gy[j];
    } else {
      const int64_t inst = inst_begin + j;
      bag = inst / src.tiles_per_bag;
      const int t = (int)(inst - bag * src.tiles_per_bag);
      const int gyi = t / src.grid_w, gxi = t - gyi * src.grid_w;
      row0 = cs::grid_coord(gyi, src.H, S, src.interval);
      col0 = cs::grid_coord(gxi, src.W, S, src.interval);
    }
    const uint8_t* p = src.img + ((bag * src.H + (row0 + r0)) * (int64_t)src.W + col0) * 3;
    float* o = out + j * 3 * plane + (int64_t)r0 * S;          // out[j][c][y][x]
    const int64_t pitch = (int64_t)src.W * 3;
    for (int x = lane; x < S; x += 32) {
      // eight rows per step: 24 byte loads in flight per lane before the first look-up (one
      // row at a time left the kernel bound by the L2 round trip, 23 % of the write roofline)
      for (int y = r0; y < r1; y += 8) {
        int u[8][3];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint8_t* pr = p + (int64_t)(y - r0 + q) * pitch + 3 * x;
          const bool ok = y + q < r1;
          u[q][0] = ok ? pr[0] : 0; u[q][1] = ok ? pr[1] : 0; u[q][2] = ok ? pr[2] : 0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (y + q >= r1) break;
          float* oq = o + (int64_t)(y - r0 + q) * S + x;
          oq[0] = l0[u[q][0] * kLutCopies];
          oq[plane] = l0[(256 + u[q][1]) * kLutCopies];
          oq[2 * plane] = l0[(512 + u[q][2]) * kLutCopies];
        }
      }
    }
  }
}

int ensure_lut() {
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && g_lut_ready[dev]) return CS_OK;
  // torchvision: mean/std lists become float32 tensors; arithmetic is fp32.
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  static float lut[3 * 256];
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) {
      volatile float t = (float)u / 255.0f;
      volatile float d = t - mean[c];
      lut[c * 256 + u] = d / stdv[c];
    }
  CS_CUDA(cudaMemcpyToSymbol(c_norm_lut, lut, sizeof(lut)));
  if (dev < 64) g_lut_ready[dev] = true;
  return CS_OK;
}

// Rows per work item: whole 32-pixel tiles split in four so small batches still fill the GPU.
int rows_per_item(int tile) { return tile <= 64 ? 8 : 16; }

template <bool kGather>
int launch_unfold(const TileSrc& src, int64_t inst_begin, int64_t inst_count, const int32_t* gbag,
                  const int32_t* gx, const int32_t* gy, float* out, cudaStream_t st) {
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(unfold_kernel<kGather>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUnfoldSmem));
    if (dev < 64) attr_done[dev] = true;
  }
  const int rows = rows_per_item(src.tile);
  const int64_t items = inst_count * ((src.tile + rows - 1) / rows);
  const int64_t want = cs::ceil_div<int64_t>(items, 8);          // 8 warps per CTA
  const int64_t cap = (int64_t)cs::num_sms() * 2;                // 96 KB of LUT copies: two CTAs per SM
  unfold_kernel<kGather><<<(int)(want < cap ? want : cap), 256, kUnfoldSmem, st>>>(src, inst_begin, inst_count, gbag,
                                                                                  gx, gy, out, rows);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // namespace

namespace cs {
// Used by the forward kernels to share the same LUT values.
int get_norm_lut_host(float* dst768) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) {
      volatile float t = (float)u / 255.0f;
      volatile float d = t - mean[c];
      dst768[c * 256 + u] = d / stdv[c];
    }
  return CS_OK;
}
}  // namespace cs

extern "C" {

int cs_unfold_normalize(const uint8_t* img, int n_bags, int H, int W, int tile, int interval,
                        int64_t inst_begin, int64_t inst_count, float* out, void* stream) {
  CS_REQUIRE(img && out, "cs_unfold_normalize: NULL pointer");
  CS_REQUIRE(tile > 0, "cs_unfold_normalize: tile %d must be positive", tile);
  int gh = cs::grid_count(H, tile, interval), gw = cs::grid_count(W, tile, interval);
  CS_REQUIRE(gh > 0 && gw > 0, "cs_unfold_normalize: bad geometry H=%d W=%d tile=%d interval=%d", H,
             W, tile, interval);
  int64_t T = (int64_t)gh * gw;
  CS_REQUIRE(inst_begin >= 0 && inst_count >= 0 && inst_begin + inst_count <= (int64_t)n_bags * T,
             "cs_unfold_normalize: instance range [%lld,+%lld) outside %d bags x %lld tiles",
             (long long)inst_begin, (long long)inst_count, n_bags, (long long)T);
  if (inst_count == 0) return CS_OK;
  int rc = ensure_lut();
  if (rc != CS_OK) return rc;
  TileSrc src{img, H, W, tile, interval, gw, T};
  return launch_unfold<false>(src, inst_begin, inst_count, nullptr, nullptr, nullptr, out, cs::as_stream(stream));
}

int cs_gather_normalize(const uint8_t* img, int n_bags, int H, int W, int tile,
                        const int32_t* bag, const int32_t* x, const int32_t* y, int64_t n_tiles,
                        float* out, void* stream) {
  CS_REQUIRE(img && out && bag && x && y, "cs_gather_normalize: NULL pointer");
  CS_REQUIRE(tile > 0 && tile <= H && tile <= W, "cs_gather_normalize: tile %d must fit %dx%d", tile, H, W);
  CS_REQUIRE(n_tiles >= 0 && n_bags > 0, "cs_gather_normalize: bad counts");
  if (n_tiles == 0) return CS_OK;
  int rc = ensure_lut();
  if (rc != CS_OK) return rc;
  TileSrc src{img, H, W, tile, 1, 1, 1};
  return launch_unfold<true>(src, 0, n_tiles, bag, x, y, out, cs::as_stream(stream));
}

}  // extern "C"
