// K4a: painting kept tiles into per-bag maps.
//
// Reference:
//   generate_masks  utils/image_processing.py:91-98   mask[g][x:x+S, y:y+S] = 1
//   heatmap         utils/image_processing.py:153-158 map[g][x:x+S, y:y+S] = prob,
//                   visited in ascending (bag, prob) order, so the last write is the
//                   largest kept covering probability: per-pixel max.
//   heatmap         :165   255 - np.uint8(255 * map)   (float64 product, truncation)
//
// Scatter form: one warp paints one tile row-by-row; masks are idempotent byte stores, heatmaps
// use atomicMax on the int view of non-negative floats (the caller zero-fills the maps).
// Gather form (heatmaps of tiles on the regular grid, cs_paint_heatmap_gather): every pixel is
// written once as the maximum over the kept tiles covering it.
#include "common.cuh"

namespace {

struct Grid {
  int H, W, tile, interval, grid_w;
  int64_t tiles_per_bag;
  int bag_base, n_bags;
};

__device__ __forceinline__ bool tile_origin(const Grid& g, int32_t inst, int64_t* bag, int* row0,
                                            int* col0) {
  int64_t b = inst / g.tiles_per_bag;
  int t = (int)(inst - b * g.tiles_per_bag);
  int gy = t / g.grid_w, gx = t - gy * g.grid_w;
  *bag = g.bag_base + b;
  *row0 = cs::grid_coord(gy, g.H, g.tile, g.interval);
  *col0 = cs::grid_coord(gx, g.W, g.tile, g.interval);
  return inst >= 0 && *bag < g.n_bags;
}

// Explicit coordinates (the reference API hands over `tiles` and `groups` arrays).
struct XY {
  const int32_t* bag;
  const int32_t* x;
  const int32_t* y;
};

__device__ __forceinline__ bool tile_origin_xy(const Grid& g, const XY& c, int64_t j, int64_t* bag,
                                               int* row0, int* col0) {
  *bag = c.bag[j];
  *row0 = c.x[j];
  *col0 = c.y[j];
  return *bag >= 0 && *bag < g.n_bags && *row0 >= 0 && *col0 >= 0 && *row0 + g.tile <= g.H &&
         *col0 + g.tile <= g.W;
}

__global__ void __launch_bounds__(256)
paint_mask_kernel(Grid g, const int32_t* __restrict__ sel, XY xy, int64_t n_sel,
                  uint8_t* __restrict__ out) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int S = g.tile;
  for (int64_t j = warp; j < n_sel; j += n_warps) {
    int64_t bag; int r0, c0;
    if (!(sel ? tile_origin(g, sel[j], &bag, &r0, &c0) : tile_origin_xy(g, xy, j, &bag, &r0, &c0)))
      continue;
    uint8_t* base = out + (bag * g.H + r0) * (int64_t)g.W + c0;
    for (int e = lane; e < S * S; e += 32) {
      int y = e / S, x = e - y * S;
      base[(int64_t)y * g.W + x] = 1;
    }
  }
}

__global__ void __launch_bounds__(256)
paint_heat_kernel(Grid g, const int32_t* __restrict__ sel, XY xy, const float* __restrict__ prob,
                  int64_t n_sel, float* __restrict__ out) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int S = g.tile;
  for (int64_t j = warp; j < n_sel; j += n_warps) {
    int64_t bag; int r0, c0;
    if (!(sel ? tile_origin(g, sel[j], &bag, &r0, &c0) : tile_origin_xy(g, xy, j, &bag, &r0, &c0)))
      continue;
    float p = prob[j];
    if (!(p >= 0.0f)) continue;  // negative or NaN never passes `prob > thr >= 0`
    int pi = __float_as_int(p);
    int* base = reinterpret_cast<int*>(out) + (bag * g.H + r0) * (int64_t)g.W + c0;
    for (int e = lane; e < S * S; e += 32) {
      int y = e / S, x = e - y * S;
      atomicMax(base + (int64_t)y * g.W + x, pi);
    }
  }
}

__global__ void __launch_bounds__(256)
heat_to_gray_kernel(const float* __restrict__ heat, int64_t n, uint8_t* __restrict__ gray) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = 255.0 * (double)heat[i];       // numpy: float64 array times python int
    gray[i] = (uint8_t)(255 - (int)(uint8_t)(int)v);  // np.uint8() truncates toward zero
  }
}

// N3: heatmap rendering tail of heatmap() (utils/image_processing.py:164-166) in one pass:
//   gray = 255 - uint8(255 * heat)                      (float64 product, truncation)
//   cm   = applyColorMap(gray, JET)                      (256 x 3 LUT, channel order as cv2 returns)
//   out  = addWeighted(img, 0.5, cm, 0.5, 0)             (a/2 + b/2 exact in fp32, cvRound =
//                                                          round-half-to-even, no saturation needed)
// 10 B/pixel (4 heat + 3 image in, 3 out); four pixels per thread with 128/96-bit accesses.
__device__ __forceinline__ uint32_t blend_half_even(uint32_t a, uint32_t b) {
  const uint32_t s = a + b, r = s >> 1;
  return r + (s & r & 1u);
}

__global__ void __launch_bounds__(256)
heat_blend_kernel(const float* __restrict__ heat, const uint8_t* __restrict__ img,
                  const uint8_t* __restrict__ lut, int64_t n_px, uint8_t* __restrict__ out) {
  __shared__ uint8_t lut_s[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut_s[i] = lut[i];
  __syncthreads();
  auto gray_of = [](float h) {
    double v = 255.0 * (double)h;
    return (uint32_t)(uint8_t)(255 - (int)(uint8_t)(int)v);
  };
  const int64_t n4 = n_px >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    const float4 h = __ldcs(reinterpret_cast<const float4*>(heat) + q);
    const uint32_t* ip = reinterpret_cast<const uint32_t*>(img) + 3 * q;
    const uint32_t w[3] = {__ldcs(ip), __ldcs(ip + 1), __ldcs(ip + 2)};
    const uint32_t g[4] = {gray_of(h.x), gray_of(h.y), gray_of(h.z), gray_of(h.w)};
    uint32_t o[3] = {0, 0, 0};
#pragma unroll
    for (int b = 0; b < 12; ++b) {          // byte b = pixel b/3, channel b%3
      const uint32_t a = (w[b >> 2] >> (8 * (b & 3))) & 0xffu;
      const uint32_t c = lut_s[g[b / 3] * 3 + b % 3];
      o[b >> 2] |= blend_half_even(a, c) << (8 * (b & 3));
    }
    uint32_t* op = reinterpret_cast<uint32_t*>(out) + 3 * q;
    __stcs(op, o[0]); __stcs(op + 1, o[1]); __stcs(op + 2, o[2]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n_px & 3)) {   // tail pixels
    const int64_t px = (n4 << 2) + threadIdx.x;
    const uint32_t g = gray_of(heat[px]);
    for (int c = 0; c < 3; ++c)
      out[px * 3 + c] = (uint8_t)blend_half_even(img[px * 3 + c], lut_s[g * 3 + c]);
  }
}

// ---- heatmap, gather form -------------------------------------------------------------------
// The kept tiles of the regular grid are first written into a dense per-bag table of
// probabilities (T floats per bag, 3 % of a map); a CTA then builds one bag's map by a
// separable maximum -- over the covering grid columns for every (grid row, x), then over the
// covering grid rows for every pixel -- and writes every pixel exactly once: no zero-fill of the
// maps, no atomics on them, no read of them.  The covering ranges of every x and y are tabulated
// once per CTA, so the per-pixel work has no division.  Grid positions covering coordinate c along an axis
// are a contiguous range, because the position coordinates min(g * I, dim - S) are monotonic
// (cs::grid_cover, common.cuh).
__global__ void __launch_bounds__(256)
heat_table_kernel(Grid g, const int32_t* __restrict__ sel, const float* __restrict__ prob, int64_t n_sel,
                  int* __restrict__ table) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_sel; j += stride) {
    const int32_t inst = sel[j];
    if (inst < 0) continue;
    const int64_t b = inst / g.tiles_per_bag;
    if (g.bag_base + b >= g.n_bags) continue;
    const float p = prob[j];
    if (!(p >= 0.0f)) continue;                            // negative or NaN never passes `prob > thr >= 0`
    // a tile listed twice keeps its larger probability, like the last write of the reference loop
    atomicMax(table + (g.bag_base + b) * g.tiles_per_bag + (inst - b * g.tiles_per_bag), __float_as_int(p));
  }
}

constexpr int kGatherThreads = 640;

// Shared memory: colmax [grid_h][W] | bag table [grid_h][grid_w] | xlo, xhi [W] | ylo, yhi [H]
__host__ __device__ inline size_t gather_smem_bytes(int grid_h, int grid_w, int H, int W) {
  return ((size_t)grid_h * W + (size_t)grid_h * grid_w + 2 * (size_t)W + 2 * (size_t)H) * 4;
}

__global__ void __launch_bounds__(kGatherThreads, 2)
heat_gather_kernel(Grid g, int grid_h, const float* __restrict__ table, float* __restrict__ out) {
  extern __shared__ float gsm[];
  const int W = g.W, H = g.H, S = g.tile, I = g.interval, gw = g.grid_w;
  const int T = grid_h * gw;
  float* colmax = gsm;                                     // maximum over the covering grid columns
  float* tab = colmax + grid_h * W;                        // this bag's table
  int* xlo = reinterpret_cast<int*>(tab + T);
  int* xhi = xlo + W;
  int* ylo = xhi + W;
  int* yhi = ylo + H;
  // the covering ranges depend on the geometry only: once per CTA
  for (int i = threadIdx.x; i < W; i += kGatherThreads) cs::grid_cover(i, W, S, I, gw, xlo + i, xhi + i);
  for (int i = threadIdx.x; i < H; i += kGatherThreads) cs::grid_cover(i, H, S, I, grid_h, ylo + i, yhi + i);
  const int step_gy = kGatherThreads / W, step_x = kGatherThreads - step_gy * W;
  const int segs = max(1, kGatherThreads / W);
  for (int bag = blockIdx.x; bag < g.n_bags; bag += gridDim.x) {
    const float* tb = table + (int64_t)bag * g.tiles_per_bag;
    for (int i = threadIdx.x; i < T; i += kGatherThreads) tab[i] = __ldg(tb + i);
    __syncthreads();                                       // (also orders the range tables on the first bag)
    // phase A: colmax[gy][x]; consecutive threads take consecutive x of a grid row
    {
      int gy = threadIdx.x / W, x = threadIdx.x - gy * W;
      while (gy < grid_h) {
        float m = 0.0f;
        const float* row = tab + gy * gw;
        for (int gx = xlo[x]; gx <= xhi[x]; ++gx) m = fmaxf(m, row[gx]);
        colmax[gy * W + x] = m;
        gy += step_gy;
        x += step_x;
        if (x >= W) { x -= W; ++gy; }
      }
    }
    __syncthreads();
    // phase B: thread (segment, x) walks its rows of column x; consecutive threads write consecutive
    // x; the maximum over the covering grid rows is recomputed only when that range moves
    for (int col = threadIdx.x; col < segs * W; col += kGatherThreads) {
      const int seg = col / W, x = col - seg * W;
      const int y0 = (int)((int64_t)H * seg / segs), y1 = (int)((int64_t)H * (seg + 1) / segs);
      float* o = out + ((int64_t)bag * H + y0) * W + x;
      int clo = -1, chi = -2;
      float m = 0.0f;
      for (int y = y0; y < y1; ++y, o += W) {
        const int lo = ylo[y], hi = yhi[y];
        if (lo != clo || hi != chi) {
          m = 0.0f;
          for (int gy = lo; gy <= hi; ++gy) m = fmaxf(m, colmax[gy * W + x]);
          clo = lo; chi = hi;
        }
        __stcs(o, m);
      }
    }
    __syncthreads();                                       // colmax / tab are rewritten for the next bag
  }
}

int make_grid(const char* fn, int H, int W, int tile, int interval, int bag_base, int n_bags,
              Grid* g) {
  int gh = cs::grid_count(H, tile, interval), gw = cs::grid_count(W, tile, interval);
  CS_REQUIRE(gh > 0 && gw > 0, "%s: bad geometry H=%d W=%d tile=%d interval=%d", fn, H, W, tile,
             interval);
  CS_REQUIRE(bag_base >= 0 && n_bags > 0, "%s: bad bag_base/n_bags", fn);
  *g = Grid{H, W, tile, interval, gw, (int64_t)gh * gw, bag_base, n_bags};
  return CS_OK;
}

int warp_grid(int64_t n_items) {
  int64_t want = cs::ceil_div<int64_t>(n_items * 32, 256);
  int64_t cap = (int64_t)cs::num_sms() * 8 * 4;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

extern "C" {

int cs_paint_mask(const int32_t* sel_idx, int64_t n_sel, int H, int W, int tile, int interval,
                  int bag_base, int n_bags, uint8_t* mask_out, void* stream) {
  CS_REQUIRE(mask_out != nullptr && (sel_idx != nullptr || n_sel == 0), "cs_paint_mask: NULL pointer");
  CS_REQUIRE(n_sel >= 0, "cs_paint_mask: n_sel < 0");
  Grid g;
  int rc = make_grid("cs_paint_mask", H, W, tile, interval, bag_base, n_bags, &g);
  if (rc != CS_OK) return rc;
  if (n_sel == 0) return CS_OK;
  paint_mask_kernel<<<warp_grid(n_sel), 256, 0, cs::as_stream(stream)>>>(g, sel_idx, XY{}, n_sel, mask_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int cs_paint_heatmap(const int32_t* sel_idx, const float* sel_prob, int64_t n_sel, int H, int W,
                     int tile, int interval, int bag_base, int n_bags, float* heat_out,
                     void* stream) {
  CS_REQUIRE(heat_out != nullptr && ((sel_idx != nullptr && sel_prob != nullptr) || n_sel == 0),
             "cs_paint_heatmap: NULL pointer");
  CS_REQUIRE(n_sel >= 0, "cs_paint_heatmap: n_sel < 0");
  Grid g;
  int rc = make_grid("cs_paint_heatmap", H, W, tile, interval, bag_base, n_bags, &g);
  if (rc != CS_OK) return rc;
  if (n_sel == 0) return CS_OK;
  paint_heat_kernel<<<warp_grid(n_sel), 256, 0, cs::as_stream(stream)>>>(g, sel_idx, XY{}, sel_prob,
                                                                        n_sel, heat_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int64_t cs_paint_heatmap_gather_workspace_bytes(int H, int W, int tile, int interval, int n_bags) {
  const int gh = cs::grid_count(H, tile, interval), gw = cs::grid_count(W, tile, interval);
  if (gh <= 0 || gw <= 0 || n_bags <= 0) return 0;
  return (int64_t)n_bags * gh * gw * (int64_t)sizeof(float);
}

int cs_paint_heatmap_gather(const int32_t* sel_idx, const float* sel_prob, int64_t n_sel, int H, int W,
                            int tile, int interval, int bag_base, int n_bags, float* heat_out,
                            void* workspace, int64_t workspace_bytes, void* stream) {
  CS_REQUIRE(heat_out != nullptr && ((sel_idx != nullptr && sel_prob != nullptr) || n_sel == 0),
             "cs_paint_heatmap_gather: NULL pointer");
  CS_REQUIRE(n_sel >= 0, "cs_paint_heatmap_gather: n_sel < 0");
  Grid g;
  int rc = make_grid("cs_paint_heatmap_gather", H, W, tile, interval, bag_base, n_bags, &g);
  if (rc != CS_OK) return rc;
  const int gh = cs::grid_count(H, tile, interval);
  const int64_t need = cs_paint_heatmap_gather_workspace_bytes(H, W, tile, interval, n_bags);
  CS_REQUIRE(workspace != nullptr && workspace_bytes >= need,
             "cs_paint_heatmap_gather: workspace too small (need %lld bytes)", (long long)need);
  const size_t smem = gather_smem_bytes(gh, g.grid_w, H, W);
  if (smem > 200 * 1024) {
    cs::set_error("cs_paint_heatmap_gather: %d grid rows x %d pixels do not fit shared memory; use cs_paint_heatmap",
                  gh, W);
    return CS_ERR_UNSUPPORTED;
  }
  cudaStream_t st = cs::as_stream(stream);
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(heat_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev < 64) attr_done[dev] = true;
  }
  CS_CUDA(cudaMemsetAsync(workspace, 0, (size_t)need, st));
  if (n_sel > 0) {
    int64_t want = cs::ceil_div<int64_t>(n_sel, 256);
    int grid = (int)(want < (int64_t)cs::num_sms() * 8 ? want : (int64_t)cs::num_sms() * 8);
    heat_table_kernel<<<grid, 256, 0, st>>>(g, sel_idx, sel_prob, n_sel, static_cast<int*>(workspace));
    CS_LAUNCH_CHECK();
  }
  const int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
  const int resident = cs::num_sms() * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm));
  heat_gather_kernel<<<n_bags < resident ? n_bags : resident, kGatherThreads, smem, st>>>(
      g, gh, static_cast<const float*>(workspace), heat_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int cs_paint_mask_xy(const int32_t* bag, const int32_t* x, const int32_t* y, int64_t n_sel, int H,
                     int W, int tile, int n_bags, uint8_t* mask_out, void* stream) {
  CS_REQUIRE(mask_out != nullptr && ((bag && x && y) || n_sel == 0), "cs_paint_mask_xy: NULL pointer");
  CS_REQUIRE(n_sel >= 0 && tile > 0 && tile <= H && tile <= W && n_bags > 0,
             "cs_paint_mask_xy: bad arguments");
  if (n_sel == 0) return CS_OK;
  Grid g{H, W, tile, 1, 1, 1, 0, n_bags};
  paint_mask_kernel<<<warp_grid(n_sel), 256, 0, cs::as_stream(stream)>>>(g, nullptr, XY{bag, x, y},
                                                                        n_sel, mask_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int cs_paint_heatmap_xy(const int32_t* bag, const int32_t* x, const int32_t* y, const float* prob,
                        int64_t n_sel, int H, int W, int tile, int n_bags, float* heat_out,
                        void* stream) {
  CS_REQUIRE(heat_out != nullptr && ((bag && x && y && prob) || n_sel == 0),
             "cs_paint_heatmap_xy: NULL pointer");
  CS_REQUIRE(n_sel >= 0 && tile > 0 && tile <= H && tile <= W && n_bags > 0,
             "cs_paint_heatmap_xy: bad arguments");
  if (n_sel == 0) return CS_OK;
  Grid g{H, W, tile, 1, 1, 1, 0, n_bags};
  paint_heat_kernel<<<warp_grid(n_sel), 256, 0, cs::as_stream(stream)>>>(g, nullptr, XY{bag, x, y},
                                                                        prob, n_sel, heat_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int cs_heatmap_to_gray(const float* heat, int64_t n, uint8_t* gray_out, void* stream) {
  CS_REQUIRE(heat && gray_out, "cs_heatmap_to_gray: NULL pointer");
  CS_REQUIRE(n >= 0, "cs_heatmap_to_gray: n < 0");
  if (n == 0) return CS_OK;
  int64_t want = cs::ceil_div<int64_t>(n, 256);
  int grid = (int)(want < cs::num_sms() * 32 ? want : cs::num_sms() * 32);
  heat_to_gray_kernel<<<grid, 256, 0, cs::as_stream(stream)>>>(heat, n, gray_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int cs_heatmap_blend(const float* heat, const uint8_t* img, const uint8_t* lut768, int64_t n_px,
                     uint8_t* out, void* stream) {
  CS_REQUIRE(heat && img && lut768 && out, "cs_heatmap_blend: NULL pointer");
  CS_REQUIRE(n_px >= 0, "cs_heatmap_blend: n_px < 0");
  CS_REQUIRE((reinterpret_cast<uintptr_t>(heat) & 15) == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 3) == 0,
             "cs_heatmap_blend: heat must be 16-byte and img/out 4-byte aligned");
  if (n_px == 0) return CS_OK;
  int64_t want = cs::ceil_div<int64_t>(cs::ceil_div<int64_t>(n_px, 4), 256);
  int grid = (int)(want < cs::num_sms() * 16 ? (want > 0 ? want : 1) : cs::num_sms() * 16);
  heat_blend_kernel<<<grid, 256, 0, cs::as_stream(stream)>>>(heat, img, lut768, n_px, out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // extern "C"
