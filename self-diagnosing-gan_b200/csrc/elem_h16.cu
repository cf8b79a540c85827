// elem_h16.cu -- HBM-streaming helpers of the 16-bit (tcgen05) discriminator path: input staging,
// pool/residual combine, and the sum-pool + SNLinear head.  16-byte (8 x 16-bit) accesses, NHWC.
// The storage type is fp16 or bf16 (template flag F16); all arithmetic is fp32.
//
// Replaces, around the tensor-core convs: transform.py:3-11 (ToTensor + Normalize on uint8 input),
// F.avg_pool2d / residual add / F.relu of torch-mimicry DBlock / DBlockOptimized, and the
// relu -> torch.sum(dim=(2,3)) -> SNLinear head of SNGANDiscriminator32/64 (SURVEY 8(a) a3,a4).
#include "kernels.cuh"

namespace sdg {

__device__ __forceinline__ float load_norm(const void* x, int layout, int64_t n, int y, int xx, int c, int H, int W) {
  if (layout == SDG_LAYOUT_U8_NHWC) {
    float v = __fdiv_rn((float)reinterpret_cast<const uint8_t*>(x)[((n * H + y) * W + xx) * 3 + c], 255.0f);
    return __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
  }
  return reinterpret_cast<const float*>(x)[((n * 3 + c) * H + y) * (int64_t)W + xx];
}

template <bool F16>
__device__ __forceinline__ void load8(const h16* p, float* f) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 t = unpack_h2<F16>(w[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}

template <bool F16>
__device__ __forceinline__ void store8(h16* p, const float* f, bool relu) {
  uint4 raw;
  uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = f[2 * j], y = f[2 * j + 1];
    if (relu) { x = fmaxf(x, 0.f); y = fmaxf(y, 0.f); }
    w[j] = pack_h2<F16>(x, y);
  }
  *reinterpret_cast<uint4*>(p) = raw;
}

// patches[n,y,x,0:27] = x[n, y+ky-1, x+kx-1, c] at k = (ky*3+kx)*3 + c (zero outside the image), 27..63 = 0
// one thread per (pixel, 8-channel group): 8 groups of 16 B per pixel
template <bool F16>
__global__ void __launch_bounds__(256)
stage_first_conv_kernel(const void* __restrict__ x, int layout, h16* __restrict__ patches, int64_t n_pix, int H, int W) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix * 8; i += stride) {
    int g = (int)(i & 7);
    int64_t pix = i >> 3;
    int xx = (int)(pix % W);
    int64_t r = pix / W;
    int y = (int)(r % H);
    int64_t n = r / H;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = g * 8 + j;
      float f = 0.f;
      if (k < 27) {
        int tap = k / 3, c = k - tap * 3;
        int iy = y + tap / 3 - 1, ix = xx + tap % 3 - 1;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) f = load_norm(x, layout, n, iy, ix, c, H, W);
      }
      v[j] = f;
    }
    store8<F16>(patches + pix * 64 + g * 8, v, false);
  }
}

// pooled[n,y,x,0:3] = avg_pool2d(x_norm, 2), 3..63 = 0       (input of DBlockOptimized.c_sc)
template <bool F16>
__global__ void __launch_bounds__(256)
stage_pooled_input_kernel(const void* __restrict__ x, int layout, h16* __restrict__ pooled, int64_t n_pix, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix * 8; i += stride) {
    int g = (int)(i & 7);
    int64_t pix = i >> 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (g == 0) {
      int xx = (int)(pix % Wo);
      int64_t r = pix / Wo;
      int y = (int)(r % Ho);
      int64_t n = r / Ho;
      for (int c = 0; c < 3; ++c) {
        float s = load_norm(x, layout, n, 2 * y, 2 * xx, c, H, W) + load_norm(x, layout, n, 2 * y, 2 * xx + 1, c, H, W) +
                  load_norm(x, layout, n, 2 * y + 1, 2 * xx, c, H, W) + load_norm(x, layout, n, 2 * y + 1, 2 * xx + 1, c, H, W);
        v[c] = s * 0.25f;
      }
    }
    store8<F16>(pooled + pix * 64 + g * 8, v, false);
  }
}

int stage_first_conv(const void* x, int layout, h16* patches, h16* pooled, int64_t n, int H, int W, int f16,
                     cudaStream_t s) {
  if (n == 0) return 0;
  int64_t np = n * H * W;
  int64_t nq = n * (H / 2) * (W / 2);
  if (f16) {
    SDG_LAUNCH(stage_first_conv_kernel<true>, stream_grid(np * 8, 256), 256, 0, s, x, layout, patches, np, H, W);
    SDG_LAUNCH(stage_pooled_input_kernel<true>, stream_grid(nq * 8, 256), 256, 0, s, x, layout, pooled, nq, H, W);
  } else {
    SDG_LAUNCH(stage_first_conv_kernel<false>, stream_grid(np * 8, 256), 256, 0, s, x, layout, patches, np, H, W);
    SDG_LAUNCH(stage_pooled_input_kernel<false>, stream_grid(nq * 8, 256), 256, 0, s, x, layout, pooled, nq, H, W);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ void pooled8(const h16* p, int W2, int C, float* f) {
  float a[8], b[8], c[8], d[8];
  load8<F16>(p, a); load8<F16>(p + C, b); load8<F16>(p + (int64_t)W2 * C, c); load8<F16>(p + (int64_t)W2 * C + C, d);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = (a[j] + b[j] + c[j] + d[j]) * 0.25f;
}

// out = (pool_a ? avgpool2(a) : a) + (pool_b ? avgpool2(b) : b); written rectified to out_relu (input of
// the next conv / the head) and, when out_raw != null, unrectified as well (textbook-shortcut mode)
template <bool F16>
__global__ void __launch_bounds__(256)
combine_h16_kernel(const h16* __restrict__ a, int pool_a, const h16* __restrict__ b, int pool_b,
                   h16* __restrict__ out_relu, h16* __restrict__ out_raw, int64_t total8, int Ho, int Wo, int C) {
  const int C8 = C / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
    int c = (int)(i % C8) * 8;
    int64_t r = i / C8;
    int x = (int)(r % Wo);
    r /= Wo;
    int y = (int)(r % Ho);
    int64_t n = r / Ho;
    int64_t off_out = i * 8;
    int64_t off_in = ((n * 2 * Ho + 2 * y) * (2 * Wo) + 2 * x) * (int64_t)C + c;
    float fa[8], fb[8];
    if (pool_a) pooled8<F16>(a + off_in, 2 * Wo, C, fa); else load8<F16>(a + off_out, fa);
    if (pool_b) pooled8<F16>(b + off_in, 2 * Wo, C, fb); else load8<F16>(b + off_out, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    store8<F16>(out_relu + off_out, fa, true);
    if (out_raw) store8<F16>(out_raw + off_out, fa, false);
  }
}

int combine_h16(const h16* a, int pool_a, const h16* b, int pool_b, h16* out_relu, h16* out_raw, int64_t n, int Ho,
                int Wo, int C, int f16, cudaStream_t s) {
  SDG_REQUIRE(C % 8 == 0, SDG_E_UNSUPPORTED, "combine_h16: C=%d", C);
  int64_t total8 = n * Ho * Wo * (C / 8);
  if (total8 == 0) return 0;
  if (f16) {
    SDG_LAUNCH(combine_h16_kernel<true>, stream_grid(total8, 256), 256, 0, s, a, pool_a, b, pool_b, out_relu, out_raw,
               total8, Ho, Wo, C);
  } else {
    SDG_LAUNCH(combine_h16_kernel<false>, stream_grid(total8, 256), 256, 0, s, a, pool_a, b, pool_b, out_relu, out_raw,
               total8, Ho, Wo, C);
  }
  return 0;
}

// logits[i] = bias + sum_c w[c] * sum_hw hrelu[i,hw,c]; one warp per sample, 8 channels per lane step
template <bool F16>
__global__ void __launch_bounds__(256)
head_h16_kernel(const h16* __restrict__ hrelu, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ logits, int64_t n, int HW, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const h16* hp = hrelu + i * HW * (int64_t)C;
    float part = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
      float s[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.f;
      for (int p = 0; p < HW; ++p) {
        float f[8];
        load8<F16>(hp + (int64_t)p * C + c, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) part = fmaf(s[j], w[c + j], part);
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) logits[i] = part + bias[0];
  }
}

int head_h16(const h16* hrelu, const float* w, const float* bias, float* logits, int64_t n, int HW, int C, int f16,
             cudaStream_t s) {
  if (n == 0) return 0;
  if (f16) {
    SDG_LAUNCH(head_h16_kernel<true>, stream_grid(n, 8), 256, 0, s, hrelu, w, bias, logits, n, HW, C);
  } else {
    SDG_LAUNCH(head_h16_kernel<false>, stream_grid(n, 8), 256, 0, s, hrelu, w, bias, logits, n, HW, C);
  }
  return 0;
}

}  // namespace sdg
