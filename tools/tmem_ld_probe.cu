// Probe: register <-> (TMEM lane, column) mapping of tcgen05.ld.16x256b.  D[r][k] = A[r][k] with
// an identity B; run 0 stores A[r][k] = r, run 1 stores A[r][k] = k.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I cellsegmentation_b200/csrc
//          -o tools/tmem_ld_probe tools/tmem_ld_probe.cu -lcuda
#include <cstdio>
#include <vector>
#include <cuda.h>
#include "tc_ptx.cuh"
using namespace cs;

__global__ void probe(int run, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  uint16_t* A = reinterpret_cast<uint16_t*>(bp);               // 128 rows x 128 B (SW128)
  uint16_t* B = reinterpret_cast<uint16_t*>(bp + 16384);       // 16 rows x 128 B (SW128)
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8192 + 1024; i += blockDim.x) A[i] = 0;
  __syncthreads();
  for (int i = tid; i < 128 * 16; i += blockDim.x) {
    int r = i / 16, k = i % 16;
    __nv_bfloat16 h = __float2bfloat16_rn(run == 0 ? (float)r : (float)k);
    A[r * 64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7)] = *reinterpret_cast<uint16_t*>(&h);
  }
  if (tid < 16) B[tid * 64 + (((tid >> 3) ^ (tid & 7)) << 3) + (tid & 7)] = 0x3f80;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    umma_bf16(tmem, umma_desc_sw128(base), umma_desc_sw128(base + 16384), umma_idesc_bf16(128, 16), 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  // two 16-lane windows per warp, 16 columns each: .16x256b.x2 -> 8 registers
  for (int win = 0; win < 2; ++win) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32 + win * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    tmem_ld_wait();
    for (int k = 0; k < 8; ++k) out[(tid * 2 + win) * 8 + k] = __uint_as_float(r[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> h0(2048), h1(2048);
  for (int run = 0; run < 2; ++run) {
    probe<<<1, 128, 32768>>>(run, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(run ? h1.data() : h0.data(), d_out, 2048 * 4, cudaMemcpyDeviceToHost);
  }
  int bad = 0;
  for (int t = 0; t < 128; ++t)
    for (int win = 0; win < 2; ++win) {
      if (t < 8 || t == 33 || t == 127) printf("thread %3d win %d:", t, win);
      for (int k = 0; k < 8; ++k) {
        int row = (int)h0[(t * 2 + win) * 8 + k], col = (int)h1[(t * 2 + win) * 8 + k];
        if (t < 8 || t == 33 || t == 127) printf(" r%d=(%d,%d)", k, row, col);
        // expected: lane = 32*warp + 16*win + (T%32)/4 + 8*((k>>1)&1), col = 8*(k>>2) + 2*(T%4) + (k&1)
        int T = t % 32;
        int erow = 32 * (t / 32) + 16 * win + T / 4 + 8 * ((k >> 1) & 1);
        int ecol = 8 * (k >> 2) + 2 * (T % 4) + (k & 1);
        bad += (row != erow) || (col != ecol);
      }
      if (t < 8 || t == 33 || t == 127) printf("\n");
    }
  printf("mismatches vs expected mapping: %d\n", bad);
  return 0;
}
