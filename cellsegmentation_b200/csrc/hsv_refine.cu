// K4b: HSV value threshold AND instance mask (bandwidth-bound elementwise kernel).
//
// Reference: preprocess_masks, utils/image_processing.py:114-120
//   img_split = cv2.split(cv2.cvtColor(img, cv2.COLOR_BGR2HSV))
//   _, mask_hsv = cv2.threshold(img_split[2], thresh=170, maxval=255, THRESH_BINARY)
//   mask = np.logical_and(mask, (1 - mask_hsv / 255).astype(bool))
// V of 8-bit HSV is max(B,G,R) and THRESH_BINARY is V > thresh, so per pixel
//   out = (mask != 0) & (max(c0,c1,c2) <= thresh).
//
// Algorithmic bytes: 3 (img) + 1 (mask) read + 1 written = 5 B / pixel.
// Layout: one thread owns 32 consecutive pixels = three 32-byte image loads, one
// 32-byte mask load, one 32-byte store; all streaming (evict-first) accesses.
#include <stdlib.h>

#include "common.cuh"

namespace {

// 256-bit global accesses (sm_100 LDG.256 / STG.256), streaming cache policy.
struct __align__(32) u32x8 { uint32_t v[8]; };

__device__ __forceinline__ u32x8 ld_stream(const u32x8* p) {
  u32x8 r;
  asm volatile(
      "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7])
      : "l"(p));
  return r;
}
// Coherent variant for the mask operand, which the caller may alias with `out`.
__device__ __forceinline__ u32x8 ld_stream_rw(const u32x8* p) {
  u32x8 r;
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                 "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_stream(u32x8* p, const u32x8& r) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
               "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]),
               "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}

// Three words = 12 bytes = 4 interleaved 3-channel pixels -> one word of 4 flags
// (0x01 where the refined mask is set).
__device__ __forceinline__ uint32_t refine4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t m,
                                            uint32_t thr4) {
  // channel-0 bytes 0,3,6,9 / channel-1 bytes 1,4,7,10 / channel-2 bytes 2,5,8,11
  uint32_t c0 = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
  uint32_t c1 = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
  uint32_t c2 = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
  uint32_t v = __vmaxu4(__vmaxu4(c0, c1), c2);
  uint32_t bright = __vcmpgtu4(v, thr4);  // 0xff where V > thresh
  uint32_t inst = __vcmpne4(m, 0u);       // 0xff where mask != 0
  return inst & ~bright & 0x01010101u;
}

constexpr int kPxPerThread = 32;
constexpr int kPxPerWarp = 32 * kPxPerThread;   // 1024 pixels = 3072 image bytes per warp block
constexpr int kWarpsPerCta = 8;

// One warp = one block of 1024 pixels.  The three image loads are fully coalesced (lane l
// takes bytes [k*1024 + 32 l, +32) of the 3072-byte block: 8 full lines per instruction);
// letting every lane read its own 96 contiguous bytes instead makes each instruction touch all
// 24 lines of the block (3x the L2 tag look-ups) and capped the kernel at 67 % of copy speed.
// The bytes are redistributed through a per-warp 3 KB shared-memory slab so that lane l ends up
// with the 96 bytes of its 32 pixels.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
hsv_refine_vec_kernel(const u32x8* __restrict__ img, const u32x8* mask, u32x8* out,
                      int64_t n_blocks, uint32_t thr4) {
  __shared__ __align__(32) uint32_t slab[kWarpsPerCta][768];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* my = slab[warp];
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t bi = (int64_t)blockIdx.x * kWarpsPerCta + warp; bi < n_blocks; bi += stride) {
    const u32x8* src = img + bi * 96;
    u32x8 a = ld_stream(src + lane);
    u32x8 b = ld_stream(src + 32 + lane);
    u32x8 c = ld_stream(src + 64 + lane);
    u32x8 m = ld_stream_rw(mask + bi * 32 + lane);
    uint4* s4 = reinterpret_cast<uint4*>(my);
    s4[2 * lane] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    s4[2 * lane + 1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
    s4[64 + 2 * lane] = make_uint4(b.v[0], b.v[1], b.v[2], b.v[3]);
    s4[64 + 2 * lane + 1] = make_uint4(b.v[4], b.v[5], b.v[6], b.v[7]);
    s4[128 + 2 * lane] = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
    s4[128 + 2 * lane + 1] = make_uint4(c.v[4], c.v[5], c.v[6], c.v[7]);
    __syncwarp();
    uint32_t w[24];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const uint4 q = s4[6 * lane + j];   // this lane's 96 bytes
      w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w;
    }
    __syncwarp();
    u32x8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = refine4(w[3 * j], w[3 * j + 1], w[3 * j + 2], m.v[j], thr4);
    st_stream(out + bi * 32 + lane, o);
  }
}

// Scalar path: tails and unaligned pointers.
__global__ void hsv_refine_scalar_kernel(const uint8_t* __restrict__ img, const uint8_t* mask,
                                         uint8_t* out, int64_t begin, int64_t end, int thr) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
    int v = max(max((int)img[3 * i], (int)img[3 * i + 1]), (int)img[3 * i + 2]);
    out[i] = (mask[i] != 0 && v <= thr) ? 1 : 0;
  }
}

// OpenCV 8-bit BGR->HSV, integer restatement of cv::hal::cvtBGRtoHSV (RGB2HSV_b,
// hsv_shift = 12, hrange = 180).  Tables: sdiv[i] = cvRound((255<<12)/i),
// hdiv[i] = cvRound((180<<12)/(6 i)), entry 0 = 0.
__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];

__global__ void bgr2hsv_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ hsv,
                               int64_t n_px) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
    int b = img[3 * i], g = img[3 * i + 1], r = img[3 * i + 2];
    int v = max(max(b, g), r);
    int vmin = min(min(b, g), r);
    int diff = v - vmin;
    int vr = (v == r) ? -1 : 0;
    int vg = (v == g) ? -1 : 0;
    int s = (diff * c_sdiv[v] + (1 << 11)) >> 12;
    int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    h = (h * c_hdiv[diff] + (1 << 11)) >> 12;
    h += (h < 0) ? 180 : 0;
    hsv[3 * i] = (uint8_t)h;
    hsv[3 * i + 1] = (uint8_t)s;
    hsv[3 * i + 2] = (uint8_t)v;
  }
}

bool g_tables_ready[64] = {false};

int ensure_hsv_tables() {
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && g_tables_ready[dev]) return CS_OK;
  int sdiv[256], hdiv[256];
  sdiv[0] = hdiv[0] = 0;
  for (int i = 1; i < 256; ++i) {
    sdiv[i] = (int)nearbyint((255 << 12) / (1.0 * i));
    hdiv[i] = (int)nearbyint((180 << 12) / (6.0 * i));
  }
  CS_CUDA(cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv)));
  CS_CUDA(cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv)));
  if (dev < 64) g_tables_ready[dev] = true;
  return CS_OK;
}

}  // namespace

extern "C" {

int cs_hsv_refine(const uint8_t* img, const uint8_t* mask, int64_t n_px, int v_thresh,
                  uint8_t* out, void* stream) {
  CS_REQUIRE(img && mask && out, "cs_hsv_refine: NULL pointer");
  CS_REQUIRE(n_px >= 0, "cs_hsv_refine: n_px < 0");
  CS_REQUIRE(v_thresh >= 0 && v_thresh <= 255, "cs_hsv_refine: v_thresh %d outside [0,255]",
             v_thresh);
  if (n_px == 0) return CS_OK;
  cudaStream_t st = cs::as_stream(stream);
  bool aligned = (((uintptr_t)img | (uintptr_t)mask | (uintptr_t)out) & 31u) == 0;
  int64_t n_vec = aligned ? n_px / kPxPerWarp : 0;   // full 1024-pixel warp blocks
  if (n_vec > 0) {
    uint32_t thr4 = (uint32_t)v_thresh * 0x01010101u;
    // 8 resident CTAs of 256 threads per SM; grid-stride over the rest.
    int64_t want = cs::ceil_div<int64_t>(n_vec, kWarpsPerCta);
    static const int ctas_per_sm = []() {
      const char* e = getenv("CELLSEG_HSV_CTAS_PER_SM");   // tuning knob (profiles/tune_hsv.py)
      int v = e ? atoi(e) : 0;
      return v > 0 ? v : 32;
    }();
    int64_t cap = (int64_t)cs::num_sms() * ctas_per_sm;
    int grid = (int)(want < cap ? want : cap);
    hsv_refine_vec_kernel<<<grid, 256, 0, st>>>((const u32x8*)img, (const u32x8*)mask,
                                               (u32x8*)out, n_vec, thr4);
    CS_LAUNCH_CHECK();
  }
  int64_t done = n_vec * kPxPerWarp;
  if (done < n_px) {
    int64_t rem = n_px - done;
    int grid = (int)(cs::ceil_div<int64_t>(rem, 256) < cs::num_sms() * 16 ? cs::ceil_div<int64_t>(rem, 256)
                                                                 : cs::num_sms() * 16);
    hsv_refine_scalar_kernel<<<grid, 256, 0, st>>>(img, mask, out, done, n_px, v_thresh);
    CS_LAUNCH_CHECK();
  }
  return CS_OK;
}

int cs_bgr2hsv_u8(const uint8_t* img, int64_t n_px, uint8_t* hsv_out, void* stream) {
  CS_REQUIRE(img && hsv_out, "cs_bgr2hsv_u8: NULL pointer");
  CS_REQUIRE(n_px >= 0, "cs_bgr2hsv_u8: n_px < 0");
  if (n_px == 0) return CS_OK;
  int rc = ensure_hsv_tables();
  if (rc != CS_OK) return rc;
  int64_t want = cs::ceil_div<int64_t>(n_px, 256);
  int grid = (int)(want < cs::num_sms() * 32 ? want : cs::num_sms() * 32);
  bgr2hsv_kernel<<<grid, 256, 0, cs::as_stream(stream)>>>(img, hsv_out, n_px);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // extern "C"
