// conv_fp32.cu -- IEEE fp32 (no TF32) CUDA-core path of the discriminator forward: the 1e-5 parity mode.
//
// Replaces F.conv2d / F.avg_pool2d / F.relu / F.leaky_relu / torch.sum / F.linear as called by the
// SNGAN blocks (torch-mimicry resblocks.py) and MNIST_DCGAN_Discriminator.forward
// (diagan-pkg/diagan/models/mnist.py:161-223) under trainer.py:145-150.  Activations are NHWC fp32.
// This path exists for exactness, not speed: the throughput mode is conv_tc.cu (tcgen05).
#include "kernels.cuh"

namespace sdg {

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == ACT_LRELU_SQRT2) return (v > 0.f ? v : 0.2f * v) * 1.4142135623730951f;   // fused_leaky_relu (op/fused_act.py:104-116)
  return v;
}

// implicit GEMM: M = n*Ho*Wo pixels, N = Cout, K = ks*ks*Cin; 64x64x16 tiles, 4x4 outputs per thread
constexpr int BM = 64, BN = 64, BK = 16, APAD = 2;

__global__ void __launch_bounds__(256)
conv_fp32_kernel(const float* __restrict__ in, const float* __restrict__ wp, const float* __restrict__ bias,
                 float* __restrict__ out, int64_t M, int H, int W, int Cin, int Cout, int Ho, int Wo, int ks,
                 int stride, int pad, int pre_act, int post_act) {
  __shared__ float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN];
  const int K = ks * ks * Cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // A-load assignment: k fastest across lanes (channels are contiguous in NHWC)
  const int a_k = tid & 15;
  const int a_m = tid >> 4;                 // + 16*j
  int64_t pix_n[4]; int pix_y[4], pix_x[4]; bool pix_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int64_t m = m0 + a_m + 16 * j;
    pix_ok[j] = m < M;
    int64_t mm = pix_ok[j] ? m : 0;
    pix_x[j] = (int)(mm % Wo);
    int64_t r = mm / Wo;
    pix_y[j] = (int)(r % Ho);
    pix_n[j] = r / Ho;
  }
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      int k = k0 + a_k;
      int tap = 0, c = 0, ky = 0, kx = 0;
      bool kok = k < K;
      if (kok) { tap = k / Cin; c = k - tap * Cin; ky = tap / ks; kx = tap - ky * ks; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (kok && pix_ok[j]) {
          int iy = pix_y[j] * stride + ky - pad, ix = pix_x[j] * stride + kx - pad;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            v = in[((pix_n[j] * H + iy) * W + ix) * Cin + c];
            v = apply_act(v, pre_act);
          }
        }
        As[a_k][a_m + 16 * j] = v;
      }
      int kb = k0 + b_k;
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kb < K && n0 + b_n < Cout) w4 = *reinterpret_cast<const float4*>(wp + (int64_t)kb * Cout + n0 + b_n);
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = w4;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int n = n0 + tx * 4;
    if (n >= Cout) continue;
    float4 o;
    o.x = apply_act(acc[i][0] + (bias ? bias[n + 0] : 0.f), post_act);
    o.y = apply_act(acc[i][1] + (bias ? bias[n + 1] : 0.f), post_act);
    o.z = apply_act(acc[i][2] + (bias ? bias[n + 2] : 0.f), post_act);
    o.w = apply_act(acc[i][3] + (bias ? bias[n + 3] : 0.f), post_act);
    *reinterpret_cast<float4*>(out + m * Cout + n) = o;
  }
}

int conv_fp32(const float* in, const float* wp, const float* bias, float* out, int64_t n, int H, int W, int Cin,
              int Cout, int ks, int stride, int pre_act, int post_act, cudaStream_t s, int pad) {
  SDG_REQUIRE(Cout % 4 == 0, SDG_E_UNSUPPORTED, "conv_fp32: Cout=%d not a multiple of 4", Cout);
  if (pad < 0) pad = ks / 2;
  int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  int64_t M = n * Ho * Wo;
  if (M == 0) return 0;
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(Cout, BN));
  SDG_LAUNCH(conv_fp32_kernel, grid, 256, 0, s, in, wp, bias, out, M, H, W, Cin, Cout, Ho, Wo, ks, stride,
             pad, pre_act, post_act);
  return 0;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pooled(const float* p, int W2, int C, int relu) {
  // F.avg_pool2d(x, 2): (x00 + x01 + x10 + x11) / 4
  float a = p[0], b = p[C], c = p[(int64_t)W2 * C], d = p[(int64_t)W2 * C + C];
  if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); c = fmaxf(c, 0.f); d = fmaxf(d, 0.f); }
  return (a + b + c + d) * 0.25f;
}

__global__ void __launch_bounds__(256)
combine_fp32_kernel(const float* __restrict__ a, int pool_a, const float* __restrict__ b, int pool_b, int relu_b,
                    float* __restrict__ out, int64_t total, int Ho, int Wo, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    int64_t r = i / C;
    int x = (int)(r % Wo);
    r /= Wo;
    int y = (int)(r % Ho);
    int64_t n = r / Ho;
    float va, vb = 0.f;
    if (pool_a) va = pooled(a + ((n * 2 * Ho + 2 * y) * (2 * Wo) + 2 * x) * (int64_t)C + c, 2 * Wo, C, 0);
    else va = a[i];
    if (b) {
      if (pool_b) vb = pooled(b + ((n * 2 * Ho + 2 * y) * (2 * Wo) + 2 * x) * (int64_t)C + c, 2 * Wo, C, relu_b);
      else { vb = b[i]; if (relu_b) vb = fmaxf(vb, 0.f); }
    }
    out[i] = b ? va + vb : va;
  }
}

int combine_fp32(const float* a, int pool_a, const float* b, int pool_b, int relu_b, float* out, int64_t n, int Ho,
                 int Wo, int C, cudaStream_t s) {
  int64_t total = n * Ho * Wo * C;
  if (total == 0) return 0;
  SDG_LAUNCH(combine_fp32_kernel, stream_grid(total, 256), 256, 0, s, a, pool_a, b, pool_b, relu_b, out, total, Ho,
             Wo, C);
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_input_kernel(const void* __restrict__ x, int layout, float* __restrict__ out, int64_t total, int H, int W) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    if (layout == SDG_LAYOUT_U8_NHWC) {
      // ToTensor (u8 -> float / 255) then Normalize((x - 0.5) / 0.5)   transform.py:3-11
      float v = __fdiv_rn((float)reinterpret_cast<const uint8_t*>(x)[i], 255.0f);
      out[i] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
    } else {
      int c = (int)(i % 3);
      int64_t r = i / 3;
      int xx = (int)(r % W);
      r /= W;
      int y = (int)(r % H);
      int64_t n = r / H;
      out[i] = reinterpret_cast<const float*>(x)[((n * 3 + c) * H + y) * (int64_t)W + xx];
    }
  }
}

int prep_input_fp32(const void* x, int layout, float* out_nhwc, int64_t n, int H, int W, cudaStream_t s) {
  int64_t total = n * H * W * 3;
  if (total == 0) return 0;
  SDG_LAUNCH(prep_input_kernel, stream_grid(total, 256), 256, 0, s, x, layout, out_nhwc, total, H, W);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// one CTA per sample; thread c owns channel c (strided over C), sums the HW positions in order
__global__ void __launch_bounds__(256)
head_sumpool_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ logits, int HW, int C, int relu) {
  const int64_t i = blockIdx.x;
  const float* hp = h + i * HW * (int64_t)C;
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) {
      float v = hp[(int64_t)p * C + c];
      s += relu ? fmaxf(v, 0.f) : v;
    }
    part = fmaf(s, w[c], part);
  }
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    logits[i] = t + bias[0];
  }
}

int head_sumpool_fp32(const float* h, const float* w, const float* bias, float* logits, int64_t n, int HW, int C,
                      int relu, cudaStream_t s) {
  if (n == 0) return 0;
  int threads = C >= 256 ? 256 : (C < 32 ? 32 : ((C + 31) / 32) * 32);
  SDG_LAUNCH(head_sumpool_kernel, (unsigned)n, threads, 0, s, h, w, bias, logits, HW, C, relu);
  return 0;
}

__global__ void __launch_bounds__(256)
head_dot_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ logits, int L) {
  const int64_t i = blockIdx.x;
  const float* hp = h + i * (int64_t)L;
  float part = 0.f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) part = fmaf(hp[j], w[j], part);
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    logits[i] = t + bias[0];
  }
}

int head_dot_fp32(const float* h, const float* w, const float* bias, float* logits, int64_t n, int L,
                  cudaStream_t s) {
  if (n == 0) return 0;
  SDG_LAUNCH(head_dot_kernel, (unsigned)n, 256, 0, s, h, w, bias, logits, L);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// DCGAN conv1 for the tensor-core path (mnist.py:163-164): Conv2d(3, 16, 3, stride 2, pad 1, bias=False) + LeakyReLU(0.2)
// straight from the image, fp32 accumulation (0.36 % of the network's MACs: CUDA cores), stored as ONE 128-byte row per
// output pixel = 16 real + 48 zero 16-bit channels, i.e. the 64-channel K chunk the tcgen05 kernels stream by TMA.
template <bool F16>
__global__ void __launch_bounds__(256, 4)
dcgan_first_conv_kernel(const void* __restrict__ x, int layout, const float* __restrict__ wp, h16* __restrict__ out,
                        int64_t total, int S, int* ovf) {
  // thread = (output pixel, half): 8 of the 16 real channels + half of the pixel row's zero padding (<= 64 registers, four
  // CTAs per SM: the kernel is latency-bound, the first version with 16 accumulators per thread ran at one CTA per SM)
  __shared__ float sw[27 * 16];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = wp[i];
  __syncthreads();
  const int So = S / 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * total; t += stride) {
    const int64_t i = t >> 1;
    const int half = (int)(t & 1);
    const int ox = (int)(i % So);
    const int64_t r = i / So;
    const int oy = (int)(r % So);
    const int64_t n = r / So;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        const bool ok = iy >= 0 && iy < S && ix >= 0 && ix < S;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (ok) {
            if (layout == SDG_LAYOUT_U8_NHWC) {
              v = __fdiv_rn((float)reinterpret_cast<const uint8_t*>(x)[((n * S + iy) * S + ix) * 3 + c], 255.0f);
              v = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
            } else {
              v = reinterpret_cast<const float*>(x)[((n * 3 + c) * S + iy) * (int64_t)S + ix];
            }
          }
          const float4* w = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + c) * 16 + half * 8);
          const float4 w0 = w[0], w1 = w[1];
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
    uint32_t pk[4];
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = acc[2 * j], b = acc[2 * j + 1];
      pk[j] = pack_h2<F16>(a > 0.f ? a : 0.2f * a, b > 0.f ? b : 0.2f * b);
      if (F16) bad |= f16x2_nonfinite_bits(pk[j]);
    }
    if (F16 && bad) range_flag_set(ovf, SDG_RANGE_ACT);
    uint4* o4 = reinterpret_cast<uint4*>(out + i * 64);
    o4[half] = make_uint4(pk[0], pk[1], pk[2], pk[3]);          // channels 8*half .. 8*half+7
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < 3; ++j) o4[2 + 3 * half + j] = z;      // this half's share of the 48 zero channels
  }
}

int dcgan_first_conv_h16(const void* x, int layout, const float* wp, h16* out, int64_t n, int S, int f16, cudaStream_t s) {
  const int64_t total = n * (S / 2) * (S / 2);
  if (total == 0) return 0;
  if (f16) { SDG_LAUNCH(dcgan_first_conv_kernel<true>, stream_grid(2 * total, 256, 16), 256, 0, s, x, layout, wp, out, total, S, t_range_flag); }
  else { SDG_LAUNCH(dcgan_first_conv_kernel<false>, stream_grid(2 * total, 256, 16), 256, 0, s, x, layout, wp, out, total, S, t_range_flag); }
  return 0;
}

}  // namespace sdg
