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

}  // namespace sdg
