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

// 4 consecutive x of one (instance, channel, row).
__device__ __forceinline__ float4 load4(const uint8_t* __restrict__ row_px, int c,
                                        const float* __restrict__ lut) {
  float4 v;
  v.x = lut[c * 256 + row_px[c]];
  v.y = lut[c * 256 + row_px[3 + c]];
  v.z = lut[c * 256 + row_px[6 + c]];
  v.w = lut[c * 256 + row_px[9 + c]];
  return v;
}

template <bool kGather>
__global__ void __launch_bounds__(256)
unfold_kernel(TileSrc src, int64_t inst_begin, int64_t inst_count, const int32_t* __restrict__ gbag,
              const int32_t* __restrict__ gx, const int32_t* __restrict__ gy,
              float* __restrict__ out) {
  __shared__ float lut[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) lut[i] = c_norm_lut[i];
  __syncthreads();

  const int S = src.tile;
  const int qx = S / 4;                       // float4 groups per row
  const int64_t per_inst = (int64_t)3 * S * qx;
  const int64_t total = inst_count * per_inst;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t j = e / per_inst;
    int r = (int)(e - j * per_inst);
    int c = r / (S * qx);
    int r2 = r - c * (S * qx);
    int y = r2 / qx;
    int x4 = (r2 - y * qx) * 4;
    int64_t bag;
    int row0, col0;
    if (kGather) {
      bag = gbag[j];
      row0 = gx[j];
      col0 = gy[j];
    } else {
      int64_t inst = inst_begin + j;
      bag = inst / src.tiles_per_bag;
      int t = (int)(inst - bag * src.tiles_per_bag);
      int gyi = t / src.grid_w, gxi = t - gyi * src.grid_w;
      row0 = cs::grid_coord(gyi, src.H, S, src.interval);
      col0 = cs::grid_coord(gxi, src.W, S, src.interval);
    }
    const uint8_t* p = src.img + ((bag * src.H + (row0 + y)) * (int64_t)src.W + (col0 + x4)) * 3;
    float4 v = load4(p, c, lut);
    // out[j][c][y][x4..x4+3]
    *reinterpret_cast<float4*>(out + ((j * 3 + c) * S + y) * (int64_t)S + x4) = v;
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

int launch_cfg(int64_t total) {
  int64_t want = cs::ceil_div<int64_t>(total, 256);
  int64_t cap = (int64_t)cs::num_sms() * 8 * 8;
  return (int)(want < cap ? want : cap);
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
  CS_REQUIRE(tile > 0 && tile % 4 == 0, "cs_unfold_normalize: tile %d must be a multiple of 4", tile);
  int gh = cs::grid_count(H, tile, interval), gw = cs::grid_count(W, tile, interval);
  CS_REQUIRE(gh > 0 && gw > 0, "cs_unfold_normalize: bad geometry H=%d W=%d tile=%d interval=%d", H,
             W, tile, interval);
  int64_t T = (int64_t)gh * gw;
  CS_REQUIRE(inst_begin >= 0 && inst_count >= 0 && inst_begin + inst_count <= (int64_t)n_bags * T,
             "cs_unfold_normalize: instance range [%lld,+%lld) outside %d bags x %lld tiles",
             (long long)inst_begin, (long long)inst_count, n_bags, (long long)T);
  CS_REQUIRE(((uintptr_t)out & 15u) == 0, "cs_unfold_normalize: out must be 16-byte aligned");
  if (inst_count == 0) return CS_OK;
  int rc = ensure_lut();
  if (rc != CS_OK) return rc;
  TileSrc src{img, H, W, tile, interval, gw, T};
  int64_t total = inst_count * 3 * tile * (tile / 4);
  unfold_kernel<false><<<launch_cfg(total), 256, 0, cs::as_stream(stream)>>>(
      src, inst_begin, inst_count, nullptr, nullptr, nullptr, out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int cs_gather_normalize(const uint8_t* img, int n_bags, int H, int W, int tile,
                        const int32_t* bag, const int32_t* x, const int32_t* y, int64_t n_tiles,
                        float* out, void* stream) {
  CS_REQUIRE(img && out && bag && x && y, "cs_gather_normalize: NULL pointer");
  CS_REQUIRE(tile > 0 && tile % 4 == 0 && tile <= H && tile <= W,
             "cs_gather_normalize: tile %d must be a multiple of 4 and fit %dx%d", tile, H, W);
  CS_REQUIRE(n_tiles >= 0 && n_bags > 0, "cs_gather_normalize: bad counts");
  CS_REQUIRE(((uintptr_t)out & 15u) == 0, "cs_gather_normalize: out must be 16-byte aligned");
  if (n_tiles == 0) return CS_OK;
  int rc = ensure_lut();
  if (rc != CS_OK) return rc;
  TileSrc src{img, H, W, tile, 1, 1, 1};
  int64_t total = n_tiles * 3 * tile * (tile / 4);
  unfold_kernel<true><<<launch_cfg(total), 256, 0, cs::as_stream(stream)>>>(src, 0, n_tiles, bag, x,
                                                                           y, out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // extern "C"
