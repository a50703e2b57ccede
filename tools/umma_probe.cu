// Probe: which shared-memory 16-byte chunks does tcgen05.mma fetch for an A descriptor with a
// given layout type / start offset / SBO / LBO?  B is a 16x16 identity, so D[r][k] = A[r][k] and
// every fetched chunk reveals its id.  Used to validate the overlapping-window (no im2col)
// operand layout of the stem kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
//   -I cellsegmentation_b200/csrc -o tools/umma_probe tools/umma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include "tc_ptx.cuh"
using namespace cs;

constexpr int kABytes = 32768;

__global__ void probe(int run, uint32_t start_off, uint32_t sbo, uint32_t lbo, uint32_t layout,
                      uint32_t base_off, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  uint16_t* A = reinterpret_cast<uint16_t*>(bp);
  uint16_t* B = reinterpret_cast<uint16_t*>(bp + kABytes);          // SW128 rows of 128 B
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kABytes / 2; i += blockDim.x) {
    int id = i / 8;
    float v = run == 0 ? (float)(id & 0xff) : (float)(id >> 8);
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    A[i] = *reinterpret_cast<uint16_t*>(&h);
  }
  for (int i = tid; i < 16 * 64; i += blockDim.x) B[i] = 0;
  __syncthreads();
  if (tid < 16) {   // B[n][k] = delta(n,k): row n at 128n, chunk (k/8)^(n&7)
    int n = tid, k = tid;
    B[n * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = 0x3f80;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    uint64_t ad = (uint64_t)(((base + start_off) >> 4) & 0x3fffu);
    ad |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    ad |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    ad |= (uint64_t)1 << 46;
    ad |= (uint64_t)(base_off & 7) << 49;
    ad |= (uint64_t)layout << 61;
    uint64_t bd = umma_desc_sw128(base + kABytes);
    umma_bf16(tmem, ad, bd, umma_idesc_bf16(128, 16), 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int k = 0; k < 16; ++k) out[tid * 16 + k] = __uint_as_float(r[k]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  struct Cfg { const char* name; uint32_t start, sbo, lbo, layout, boff; };
  std::vector<Cfg> cfgs = {
      {"sw32 canonical", 0, 256, 16, 6, 0},      {"sw32 start16", 16, 256, 16, 6, 0},
      {"sw32 start32", 32, 256, 16, 6, 0},       {"sw32 start128", 128, 256, 16, 6, 0},
      {"sw32 sbo608", 0, 608, 16, 6, 0},         {"sw32 start304 sbo608", 304, 608, 16, 6, 0},
      {"sw32 start352 sbo608", 352, 608, 16, 6, 0}, {"sw32 start128 boff1", 128, 256, 16, 6, 1},
      {"none lbo16 sbo608", 0, 608, 16, 0, 0},   {"none lbo128 sbo256", 0, 256, 128, 0, 0},
      {"none lbo16 sbo128 start48", 48, 128, 16, 0, 0},
      {"sw64 start0 sbo512", 0, 512, 16, 4, 0},  {"sw64 start32 sbo512", 32, 512, 16, 4, 0},
      {"sw128 start0 sbo1024", 0, 1024, 16, 2, 0}, {"sw128 start64 sbo1200", 64, 1200, 16, 2, 0},
      {"sw128 start128 sbo1280", 128, 1280, 16, 2, 0}, {"sw128 start256 sbo1280", 256, 1280, 16, 2, 0},
      {"sw128 start288 sbo1280", 288, 1280, 16, 2, 0}, {"sw128 start128 sbo1280 boff1", 128, 1280, 16, 2, 1},
  };
  float* d_out;
  cudaMalloc(&d_out, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kABytes + 4096 + 1024);
  std::vector<float> h0(2048), h1(2048);
  for (const Cfg& c : cfgs) {
    for (int run = 0; run < 2; ++run) {
      probe<<<1, 128, kABytes + 4096 + 1024>>>(run, c.start, c.sbo, c.lbo, c.layout, c.boff, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(run ? h1.data() : h0.data(), d_out, 2048 * 4, cudaMemcpyDeviceToHost);
    }
    // models: A = XOR from absolute address bits, B = XOR from row index
    int okA = 0, okB = 0, okN = 0, tot = 0;
    printf("== %s (start %u sbo %u lbo %u layout %u boff %u)\n", c.name, c.start, c.sbo, c.lbo, c.layout, c.boff);
    const int rowb = c.layout == 6 ? 32 : c.layout == 4 ? 64 : c.layout == 2 ? 128 : 16;
    const int sh = c.layout == 6 ? 7 : c.layout == 4 ? 7 : 7;
    const int mask = c.layout == 6 ? 1 : c.layout == 4 ? 3 : c.layout == 2 ? 7 : 0;
    for (int r = 0; r < 128; ++r) {
      for (int ch = 0; ch < 2; ++ch) {
        int uniform = 1;
        for (int k = 1; k < 8; ++k)
          if (h0[r * 16 + ch * 8 + k] != h0[r * 16 + ch * 8] || h1[r * 16 + ch * 8 + k] != h1[r * 16 + ch * 8]) uniform = 0;
        int id = (int)h0[r * 16 + ch * 8] + 256 * (int)h1[r * 16 + ch * 8];
        uint32_t L;
        if (c.layout == 0) L = c.start + c.sbo * (r / 8) + 16 * (r % 8) + c.lbo * ch;
        else L = c.start + c.sbo * (r / 8) + rowb * (r % 8) + 16 * ch;
        uint32_t pa = L ^ (((L >> sh) & mask) << 4);
        uint32_t pb = L ^ ((((uint32_t)(r % 8) * rowb >> 7) & mask) << 4);
        okA += (int)(pa / 16) == id; okB += (int)(pb / 16) == id; okN += (int)(L / 16) == id; ++tot;
        if (r < 10 || (r >= 16 && r < 18)) printf("  r%3d c%d id %4d%s  A:%4u B:%4u plain:%4u\n", r, ch, id, uniform ? "" : " (mixed)", pa / 16, pb / 16, L / 16);
      }
    }
    printf("  match: absolute-address %d/%d, row-index %d/%d, no-swizzle %d/%d\n", okA, tot, okB, tot, okN, tot);
  }
  return 0;
}
