// sg2_fp32.cu -- fp32 pieces of the StyleGAN2 discriminator that are not plain convolutions.
//
// Replaces, from diagan-pkg/diagan/models/stylegan2.py (twin of stylegan2/model.py):
//   Blur.forward -> upfirdn2d(x, k, pad)        :75-90 with op/upfirdn2d.py (native kernel op/upfirdn2d_kernel.cu:107-207)
//   ResBlock's (out + skip) / sqrt(2)           :611-614
//   minibatch-stddev + concat                   :662-670
//   EqualLinear on the NCHW-flattened map       :672-673 (weight re-ordered once so activations stay NHWC)
// Correctness-first CUDA-core kernels (activations NHWC fp32); the tensor-core form of this network is future work.
#include "kernels.cuh"

namespace sdg {

// 4x4 FIR = outer([1,3,3,1]) / 64, symmetric (so the flip in upfirdn2d is a no-op); zero padding `pad` on every side
__global__ void __launch_bounds__(256)
blur_fp32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total, int H, int W, int C, int Ho, int Wo,
                 int pad) {
  const float k1[4] = {1.f, 3.f, 3.f, 1.f};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int x = (int)(r % Wo);
    r /= Wo;
    const int y = (int)(r % Ho);
    const int64_t n = r / Ho;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int iy = y + a - pad;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ix = x + b - pad;
        if (ix < 0 || ix >= W) continue;
        acc = fmaf(k1[a] * k1[b] * (1.f / 64.f), in[((n * H + iy) * W + ix) * (int64_t)C + c], acc);
      }
    }
    out[i] = acc;
  }
}

int blur_fp32(const float* in, float* out, int64_t n, int H, int W, int C, int pad, cudaStream_t s) {
  const int Ho = H + 2 * pad - 3, Wo = W + 2 * pad - 3;
  const int64_t total = n * Ho * Wo * C;
  if (total == 0) return 0;
  SDG_LAUNCH(blur_fp32_kernel, stream_grid(total, 256), 256, 0, s, in, out, total, H, W, C, Ho, Wo, pad);
  return 0;
}

__global__ void __launch_bounds__(256)
add_div_sqrt2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    out[i] = __fdiv_rn(a[i] + b[i], 1.4142135623730951f);
}

int add_div_sqrt2_fp32(const float* a, const float* b, float* out, int64_t total, cudaStream_t s) {
  if (total == 0) return 0;
  SDG_LAUNCH(add_div_sqrt2_kernel, stream_grid(total, 256), 256, 0, s, a, b, out, total);
  return 0;
}

// one CTA per (column m of the group view, batch): sd = mean_{p,c} sqrt(var_g(h[g*M+m][p][c]) + 1e-8), group = min(B,4)
__global__ void __launch_bounds__(256)
stddev_kernel(const float* __restrict__ in, float* __restrict__ sd, int batch, int group, int HW, int C) {
  const int M = batch / group;
  const int m = blockIdx.x;
  const int64_t b0 = (int64_t)blockIdx.y * batch;
  const int per = HW * C;
  float part = 0.f;
  for (int e = threadIdx.x; e < per; e += blockDim.x) {
    float mean = 0.f;
    for (int g = 0; g < group; ++g) mean += in[(b0 + (int64_t)g * M + m) * per + e];
    mean /= (float)group;
    float var = 0.f;
    for (int g = 0; g < group; ++g) {
      const float d = in[(b0 + (int64_t)g * M + m) * per + e] - mean;
      var = fmaf(d, d, var);
    }
    part += sqrtf(var / (float)group + 1e-8f);
  }
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    sd[blockIdx.y * M + m] = t / (float)per;
  }
}

__global__ void __launch_bounds__(256)
cat_stddev_kernel(const float* __restrict__ in, const float* __restrict__ sd, float* __restrict__ out, int64_t total,
                  int batch, int group, int HW, int C) {
  const int M = batch / group;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % (C + 1));
    const int64_t pix = i / (C + 1);                   // n*HW + p
    const int64_t n = pix / HW;
    if (c < C) out[i] = in[pix * C + c];
    else out[i] = sd[(n / batch) * M + (n % batch) % M];   // stddev.repeat(group, 1, H, W): sample b gets column b % M
  }
}

int minibatch_stddev_cat_fp32(const float* in, float* out, float* sd_scratch, int64_t n, int batch, int HW, int C,
                              cudaStream_t s) {
  SDG_REQUIRE(batch >= 1 && n % batch == 0, SDG_E_INVALID, "minibatch_stddev: n=%lld is not a multiple of the batch %d",
              (long long)n, batch);
  const int group = batch < 4 ? batch : 4;
  SDG_REQUIRE(batch % group == 0, SDG_E_INVALID, "minibatch_stddev: batch %d not divisible by the group %d", batch, group);
  if (n == 0) return 0;
  SDG_LAUNCH(stddev_kernel, dim3(batch / group, (unsigned)(n / batch)), 256, 0, s, in, sd_scratch, batch, group, HW, C);
  const int64_t total = n * HW * (C + 1);
  SDG_LAUNCH(cat_stddev_kernel, stream_grid(total, 256), 256, 0, s, in, sd_scratch, out, total, batch, group, HW, C);
  return 0;
}

__global__ void __launch_bounds__(256)
pack_linear_nchw_kernel(const float* __restrict__ W, float mul, float* __restrict__ wp, int O, int C, int HW) {
  const int64_t total = (int64_t)O * C * HW;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int o = (int)(i % O);
    const int64_t k = i / O;                 // p*C + c
    const int c = (int)(k % C), p = (int)(k / C);
    wp[i] = W[(int64_t)o * C * HW + (int64_t)c * HW + p] * mul;
  }
}

int pack_linear_nchw_fp32(const float* W, float mul, float* wp, int O, int C, int HW, cudaStream_t s) {
  SDG_LAUNCH(pack_linear_nchw_kernel, stream_grid((int64_t)O * C * HW, 256), 256, 0, s, W, mul, wp, O, C, HW);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// 16-bit (tensor-core path) helpers: first 1x1 conv from the image, blur, widening of the last feature map
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sg2_norm_px(const void* x, int layout, int64_t pix, int c, int64_t HW) {
  if (layout == SDG_LAYOUT_U8_NHWC) {
    float v = __fdiv_rn((float)reinterpret_cast<const uint8_t*>(x)[pix * 3 + c], 255.0f);
    return __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
  }
  const int64_t n = pix / HW, r = pix - n * HW;
  return reinterpret_cast<const float*>(x)[(n * 3 + c) * HW + r];
}

// ConvLayer(3, C, 1): out[pix][o] = flrelu(sum_c w[c][o] * x[pix][c] + b[o]); w3 = [3][C] (pack_conv_fp32 layout), already
// carrying 1/sqrt(3).
// one thread per (pixel, 8 output channels): HBM-write-bound (2*C bytes per pixel)
template <bool F16>
__global__ void __launch_bounds__(256)
sg2_first_conv_kernel(const void* __restrict__ x, int layout, const float* __restrict__ w3, const float* __restrict__ bias,
                      h16* __restrict__ out, int64_t n_pix, int64_t HW, int C) {
  const int C8 = C / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix * C8; i += stride) {
    const int g = (int)(i % C8);
    const int64_t pix = i / C8;
    const float x0 = sg2_norm_px(x, layout, pix, 0, HW), x1 = sg2_norm_px(x, layout, pix, 1, HW),
                x2 = sg2_norm_px(x, layout, pix, 2, HW);
    uint4 pk;
    uint32_t* h = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int o = g * 8 + 2 * j + e;
        float a = fmaf(w3[2 * C + o], x2, fmaf(w3[C + o], x1, fmaf(w3[o], x0, bias[o])));
        v[e] = (a > 0.f ? a : 0.2f * a) * 1.4142135623730951f;
      }
      h[j] = pack_h2<F16>(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(out + pix * C + g * 8) = pk;
  }
}

int sg2_first_conv_h16(const void* x, int layout, const float* w3, const float* bias, h16* out, int64_t n, int S, int C, int f16,
                       cudaStream_t s) {
  const int64_t n_pix = n * S * S;
  if (n_pix == 0) return 0;
  if (f16) { SDG_LAUNCH(sg2_first_conv_kernel<true>, stream_grid(n_pix * (C / 8), 256), 256, 0, s, x, layout, w3, bias, out, n_pix, (int64_t)S * S, C); }
  else { SDG_LAUNCH(sg2_first_conv_kernel<false>, stream_grid(n_pix * (C / 8), 256), 256, 0, s, x, layout, w3, bias, out, n_pix, (int64_t)S * S, C); }
  return 0;
}

// Blur on 16-bit NHWC: one thread per (output pixel, 8 channels), fp32 accumulation of the 16 taps.  `st` = 2 keeps only
// the even blur outputs (all that the stride-2 1x1 skip conv reads, stylegan2.py:575-590)
template <bool F16>
__global__ void __launch_bounds__(256)
blur_h16_kernel(const h16* __restrict__ in, h16* __restrict__ out, int64_t total8, int H, int W, int C, int Ho, int Wo, int pad,
                int st) {
  const float k1[4] = {1.f, 3.f, 3.f, 1.f};
  const int C8 = C / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
    const int c = (int)(i % C8) * 8;
    int64_t r = i / C8;
    const int x = (int)(r % Wo);
    r /= Wo;
    const int y = (int)(r % Ho);
    const int64_t n = r / Ho;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int iy = y * st + a - pad;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ix = x * st + b - pad;
        if (ix < 0 || ix >= W) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(in + ((n * H + iy) * W + ix) * (int64_t)C + c);
        const uint32_t* wv = reinterpret_cast<const uint32_t*>(&raw);
        const float kw = k1[a] * k1[b] * (1.f / 64.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_h2<F16>(wv[j]);
          acc[2 * j] = fmaf(kw, t.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(kw, t.y, acc[2 * j + 1]);
        }
      }
    }
    uint4 pk;
    uint32_t* h = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = pack_h2<F16>(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + i * 8) = pk;
  }
}

int blur_h16(const h16* in, h16* out, int64_t n, int H, int W, int C, int pad, int stride, int f16, cudaStream_t s) {
  const int Ho = (H + 2 * pad - 4) / stride + 1, Wo = (W + 2 * pad - 4) / stride + 1;
  const int64_t total8 = n * Ho * Wo * (C / 8);
  if (total8 == 0) return 0;
  if (f16) { SDG_LAUNCH(blur_h16_kernel<true>, stream_grid(total8, 256), 256, 0, s, in, out, total8, H, W, C, Ho, Wo, pad, stride); }
  else { SDG_LAUNCH(blur_h16_kernel<false>, stream_grid(total8, 256), 256, 0, s, in, out, total8, H, W, C, Ho, Wo, pad, stride); }
  return 0;
}

template <bool F16>
__global__ void widen_h16_kernel(const h16* __restrict__ in, float* __restrict__ out, int64_t total2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += stride) {
    const float2 t = unpack_h2<F16>(reinterpret_cast<const uint32_t*>(in)[i]);
    reinterpret_cast<float2*>(out)[i] = t;
  }
}

int widen_h16(const h16* in, float* out, int64_t total, int f16, cudaStream_t s) {
  if (total == 0) return 0;
  if (f16) { SDG_LAUNCH(widen_h16_kernel<true>, stream_grid(total / 2, 256), 256, 0, s, in, out, total / 2); }
  else { SDG_LAUNCH(widen_h16_kernel<false>, stream_grid(total / 2, 256), 256, 0, s, in, out, total / 2); }
  return 0;
}

}  // namespace sdg
