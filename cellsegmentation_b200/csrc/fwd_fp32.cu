// K2 (fp32 parity mode): eval-mode ResNet tile forward on CUDA cores, fp32 end to end.
//
// Reference: MILResNet.resnet_forward / forward tile branch (model/resnet.py:234-269),
// BasicBlock.forward (:28-43), softmax(dim=1)[:,1] (inference.py:24-27).
// BatchNorm is folded into conv weight/bias by the caller (eval mode: running
// statistics, model/resnet.py:254-258 + inference.py:12).
//
// TF32 tensor cores miss the 1e-4 gate ("Precision gates"), so this mode is
// plain FFMA: one implicit-GEMM kernel  out[m][co] = act(bias[co] + sum_k A[m][k] W[k][co]
// (+ residual)) with m = (instance, oy, ox), k = (dy, dx, ci), 64x64x16 CTA tiles and
// 4x4 register tiles.  It exists for parity and as the on-device cross-check of the
// tcgen05 path, not for throughput.
#include "fwd.cuh"

namespace cs {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
conv_fp32_kernel(ConvF32Args a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int P = a.Ho * a.Wo;
  const int K = a.k * a.k * a.Cin;

  // A-load mapping: 64 rows x 16 k = 1024 elements, 4 per thread
  // thread -> row (tid % 64), k lanes (tid / 64) * 4 .. +3
  const int a_row = tid & 63;
  const int a_k0 = (tid >> 6) * 4;
  const int64_t m = m0 + a_row;
  const bool m_ok = m < a.M;
  int64_t n_img = 0;
  int oy = 0, ox = 0;
  if (m_ok) {
    n_img = m / P;
    int p = (int)(m - n_img * P);
    oy = p / a.Wo;
    ox = p - oy * a.Wo;
  }
  // B-load mapping: 16 k x 64 n = 1024, 4 per thread: k = tid / 16, n = (tid % 16) * 4
  const int b_k = tid >> 4;
  const int b_n = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, each 4 rows x 4 cols

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk = k0 + a_k0 + i;
      float v = 0.f;
      if (m_ok && kk < K) {
        int tap = kk / a.Cin, ci = kk - tap * a.Cin;
        int dy = tap / a.k, dx = tap - dy * a.k;
        int iy = oy * a.stride - a.pad + dy, ix = ox * a.stride - a.pad + dx;
        if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi)
          v = a.in[n_img * a.in_sn + (int64_t)ci * a.in_sc + (int64_t)iy * a.in_sy +
                   (int64_t)ix * a.in_sx];
      }
      As[a_k0 + i][a_row] = v;
    }
    {
      int kk = k0 + b_k;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < K && n0 + b_n < a.Cout)
        w = *reinterpret_cast<const float4*>(a.w + (int64_t)kk * a.Cout + n0 + b_n);
      Bs[b_k][b_n] = w.x; Bs[b_k][b_n + 1] = w.y; Bs[b_k][b_n + 2] = w.z; Bs[b_k][b_n + 3] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t mm = m0 + ty * 4 + i;
    if (mm >= a.M) continue;
    int n = n0 + tx * 4;
    if (n >= a.Cout) continue;
    float4 o;
    float* po = &o.x;
    const float4 bb = *reinterpret_cast<const float4*>(a.bias + n);
    const float* pb = &bb.x;
    float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.residual) rr = *reinterpret_cast<const float4*>(a.residual + mm * a.Cout + n);
    const float* pr = &rr.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + pb[j];
      if (a.residual) v += pr[j];
      if (a.relu) v = fmaxf(v, 0.f);
      po[j] = v;
    }
    *reinterpret_cast<float4*>(a.out + mm * a.Cout + n) = o;
  }
}

// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC fp32 (model/resnet.py:114, :239).
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, int Hi,
                    int Wi, int Ho, int Wo, int C) {
  int64_t total = n * Ho * Wo * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int c = (int)(e % C);
    int64_t r = e / C;
    int ox = (int)(r % Wo); r /= Wo;
    int oy = (int)(r % Ho);
    int64_t img = r / Ho;
    float v = -INFINITY;
    for (int dy = 0; dy < 3; ++dy) {
      int iy = oy * 2 - 1 + dy;
      if (iy < 0 || iy >= Hi) continue;
      for (int dx = 0; dx < 3; ++dx) {
        int ix = ox * 2 - 1 + dx;
        if (ix < 0 || ix >= Wi) continue;
        v = fmaxf(v, in[((img * Hi + iy) * Wi + ix) * (int64_t)C + c]);
      }
    }
    out[e] = v;
  }
}

// Tile head: avgpool(1)+maxpool(1) -> Linear(C,2) -> softmax[:,1]
// (model/resnet.py:266-267, inference.py:24-27).  One warp per instance.
__global__ void __launch_bounds__(256)
head_fp32_kernel(const float* __restrict__ x4, int64_t n, int P, int C,
                 const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                 float* __restrict__ prob_out, float* __restrict__ logits_out,
                 float* __restrict__ feat_out) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* x = x4 + warp * (int64_t)P * C;
  float z0 = 0.f, z1 = 0.f;
  for (int c = lane; c < C; c += 32) {
    float s = 0.f, mx = -INFINITY;
    for (int p = 0; p < P; ++p) {
      float v = x[(int64_t)p * C + c];
      s += v;
      mx = fmaxf(mx, v);
    }
    float f = s / (float)P + mx;
    if (feat_out) feat_out[warp * C + c] = f;
    z0 = fmaf(f, fc_w[c], z0);
    z1 = fmaf(f, fc_w[C + c], z1);
  }
  for (int o = 16; o > 0; o >>= 1) {
    z0 += __shfl_xor_sync(0xffffffffu, z0, o);
    z1 += __shfl_xor_sync(0xffffffffu, z1, o);
  }
  if (lane == 0) {
    z0 += fc_b[0];
    z1 += fc_b[1];
    if (logits_out) { logits_out[warp * 2] = z0; logits_out[warp * 2 + 1] = z1; }
    if (prob_out) {
      float mx = fmaxf(z0, z1);
      float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
      prob_out[warp] = e1 / (e0 + e1);
    }
  }
}

// F.interpolate(mode="bilinear", align_corners=True) on NHWC fp32 (model/resnet.py:282-300):
// source coordinate = dst * (in - 1) / (out - 1), the two-tap lerp ATen performs in fp32
// (upsample_bilinear2d: h1lambda = src - floor(src), out = h0l*(w0l*a + w1l*b) + h1l*(...)).
__global__ void __launch_bounds__(256)
bilinear_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, int Hi, int Wi,
                   int Ho, int Wo, int C, float sh, float sw) {
  const int64_t total = n * Ho * Wo * C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C);
    int64_t r = e / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int64_t img = r / Ho;
    const float fy = sh * (float)oy, fx = sw * (float)ox;
    int y0 = (int)fy, x0 = (int)fx;
    y0 = y0 < Hi - 1 ? y0 : Hi - 1;
    x0 = x0 < Wi - 1 ? x0 : Wi - 1;
    const int y1 = y0 < Hi - 1 ? y0 + 1 : y0, x1 = x0 < Wi - 1 ? x0 + 1 : x0;
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* base = in + img * (int64_t)Hi * Wi * C + c;
    const float a = base[((int64_t)y0 * Wi + x0) * C], b = base[((int64_t)y0 * Wi + x1) * C];
    const float cc = base[((int64_t)y1 * Wi + x0) * C], d = base[((int64_t)y1 * Wi + x1) * C];
    out[e] = hy * (hx * a + lx * b) + ly * (hx * cc + lx * d);
  }
}

}  // namespace

int launch_bilinear_fp32(const float* in, float* out, int64_t n, int Hi, int Wi, int Ho, int Wo, int C,
                         cudaStream_t st) {
  const int64_t total = n * Ho * Wo * C;
  if (total <= 0) return CS_OK;
  // area_pixel_compute_scale(align_corners=True): (in - 1) / (out - 1), 0 when out == 1
  const float sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const int64_t want = ceil_div<int64_t>(total, 256);
  const int grid = (int)(want < (int64_t)num_sms() * 32 ? want : (int64_t)num_sms() * 32);
  bilinear_ac_kernel<<<grid, 256, 0, st>>>(in, out, n, Hi, Wi, Ho, Wo, C, sh, sw);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int launch_conv_fp32(const ConvF32Args& a, cudaStream_t st) {
  if (a.Cout % 4 != 0) {
    set_error("conv_fp32: Cout %d not a multiple of 4", a.Cout);
    return CS_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)ceil_div<int64_t>(a.M, BM), (unsigned)ceil_div(a.Cout, BN));
  conv_fp32_kernel<<<grid, 256, 0, st>>>(a);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int launch_maxpool_fp32(const float* in, float* out, int64_t n, int Hi, int Wi, int C,
                        cudaStream_t st) {
  int Ho = (Hi + 2 - 3) / 2 + 1, Wo = (Wi + 2 - 3) / 2 + 1;
  int64_t total = n * Ho * Wo * C;
  int64_t want = ceil_div<int64_t>(total, 256);
  int grid = (int)(want < (int64_t)num_sms() * 32 ? want : (int64_t)num_sms() * 32);
  maxpool3x3s2_kernel<<<grid, 256, 0, st>>>(in, out, n, Hi, Wi, Ho, Wo, C);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int launch_head_fp32(const float* x4, int64_t n, int P, int C, const float* fc_w,
                     const float* fc_b, float* prob_out, float* logits_out, float* feat_out,
                     cudaStream_t st) {
  int64_t blocks = ceil_div<int64_t>(n * 32, 256);
  head_fp32_kernel<<<(unsigned)blocks, 256, 0, st>>>(x4, n, P, C, fc_w, fc_b, prob_out, logits_out,
                                                    feat_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // namespace cs
