// sg2_fp32.cu -- fp32 pieces of the StyleGAN2 discriminator that are not plain convolutions.
//
// Replaces, from diagan-pkg/diagan/models/stylegan2.py (twin of stylegan2/model.py):
//   Blur.forward -> upfirdn2d(x, k, pad)        :75-90 with op/upfirdn2d.py (native kernel op/upfirdn2d_kernel.cu:107-207)
//   ResBlock's (out + skip) / sqrt(2)           :611-614
//   minibatch-stddev + concat                   :662-670
//   EqualLinear on the NCHW-flattened map       :672-673 (weight re-ordered once so activations stay NHWC)
// fp32 CUDA-core kernels of the exact engine, plus the 16-bit streaming pieces (first conv, blur, split-precision tail
// operands) of the tensor-core engine whose convolutions live in conv_tc.cu.
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
stddev_kernel(const float* __restrict__ in, float* __restrict__ sd, float* __restrict__ sd_sample, int batch, int group, int HW,
              int C) {
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
    const float v = t / (float)per;
    if (sd) sd[blockIdx.y * M + m] = v;
    // stddev.repeat(group, 1, H, W): every sample b of this batch with b % M == m sees this value
    if (sd_sample)
      for (int g = 0; g < group; ++g) sd_sample[b0 + (int64_t)g * M + m] = v;
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
  SDG_LAUNCH(stddev_kernel, dim3(batch / group, (unsigned)(n / batch)), 256, 0, s, in, sd_scratch, (float*)nullptr, batch, group, HW, C);
  const int64_t total = n * HW * (C + 1);
  SDG_LAUNCH(cat_stddev_kernel, stream_grid(total, 256), 256, 0, s, in, sd_scratch, out, total, batch, group, HW, C);
  return 0;
}

int minibatch_stddev_fp32(const float* in, float* sd_sample, int64_t n, int batch, int HW, int C, cudaStream_t s) {
  SDG_REQUIRE(batch >= 1 && n % batch == 0, SDG_E_INVALID, "minibatch_stddev: n=%lld is not a multiple of the batch %d",
              (long long)n, batch);
  const int group = batch < 4 ? batch : 4;
  SDG_REQUIRE(batch % group == 0, SDG_E_INVALID, "minibatch_stddev: batch %d not divisible by the group %d", batch, group);
  if (n == 0) return 0;
  SDG_LAUNCH(stddev_kernel, dim3(batch / group, (unsigned)(n / batch)), 256, 0, s, in, (float*)nullptr, sd_sample, batch, group, HW, C);
  return 0;
}

__global__ void __launch_bounds__(256)
pack_const_channel_kernel(const float* __restrict__ W, float mul, float* __restrict__ wsum, int Cout, int cin_w, int c_extra, int S) {
  const int total = S * S * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int o = i % Cout, p = i / Cout;
    const int y = p / S, x = p % S;
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = y + ky - 1, ix = x + kx - 1;
        if (iy >= 0 && iy < S && ix >= 0 && ix < S) acc += W[((int64_t)o * cin_w + c_extra) * 9 + ky * 3 + kx] * mul;
      }
    wsum[i] = acc;
  }
}

int pack_const_channel_fp32(const float* W, float mul, float* wsum, int Cout, int cin_w, int c_extra, int S, cudaStream_t s) {
  SDG_LAUNCH(pack_const_channel_kernel, stream_grid((int64_t)S * S * Cout, 256), 256, 0, s, W, mul, wsum, Cout, cin_w, c_extra, S);
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
// 16-bit (tensor-core path) helpers: first 1x1 conv from the image, blur
// ---------------------------------------------------------------------------------------------------
// ConvLayer(3, C, 1): out[pix][o] = flrelu(sum_c w[c][o] * x[pix][c] + b[o]); w3 = [3][C] (pack_conv_fp32 layout), already
// carrying 1/sqrt(3).  HBM-write-bound (2*C bytes per pixel).  A thread owns ONE group of 8 output channels for the whole
// kernel (its 24 weights + 8 biases stay in registers) and strides over pixels; consecutive lanes write consecutive 16 B.
// C and S are powers of two, so all index arithmetic is shifts (64-bit division is what made the first version slow).
template <bool F16>
__global__ void __launch_bounds__(256, 4)
sg2_first_conv_kernel(const void* __restrict__ x, int layout, const float* __restrict__ w3, const float* __restrict__ bias,
                      h16* __restrict__ out, int64_t n_pix, int hw_shift, int c8_shift, int C) {
  const int C8 = 1 << c8_shift;
  const int g = threadIdx.x & (C8 - 1);
  // normalisation (x/255 - 0.5)/0.5 with IEEE divisions (transform.py:3-11) through a 256-entry table: the divisions were
  // most of the instruction stream of this otherwise store-bound kernel
  __shared__ float lut[256];
  lut[threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.0f), 0.5f), 0.5f);
  __syncthreads();
  // leaky_relu is positively homogeneous: lrelu(a) * sqrt(2) = lrelu(sqrt(2) * a), so sqrt(2) rides on the weights and the bias
  float w[3][8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    b[j] = bias[g * 8 + j] * 1.4142135623730951f;
#pragma unroll
    for (int c = 0; c < 3; ++c) w[c][j] = w3[c * C + g * 8 + j] * 1.4142135623730951f;
  }
  const int ppb = 256 >> c8_shift;                                   // pixels per block sub-iteration
  const int64_t HWm = ((int64_t)1 << hw_shift) - 1;
  constexpr int U = 4;                                               // pixels in flight per thread (loads batched)
  for (int64_t p0 = ((int64_t)blockIdx.x * U) * ppb + (threadIdx.x >> c8_shift); p0 < n_pix; p0 += (int64_t)gridDim.x * U * ppb) {
    float xv[U][3];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pix = p0 + (int64_t)u * ppb;
      if (pix >= n_pix) { xv[u][0] = xv[u][1] = xv[u][2] = 0.f; continue; }
      if (layout == SDG_LAYOUT_U8_NHWC) {
        const uint8_t* px = reinterpret_cast<const uint8_t*>(x) + pix * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) xv[u][c] = lut[__ldg(px + c)];
      } else {
        const int64_t n = pix >> hw_shift, r = pix & HWm;
#pragma unroll
        for (int c = 0; c < 3; ++c) xv[u][c] = __ldg(reinterpret_cast<const float*>(x) + (((n * 3 + c)) << hw_shift) + r);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pix = p0 + (int64_t)u * ppb;
      if (pix >= n_pix) break;
      uint4 pk;
      uint32_t* h = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int o = 2 * j + e;
          const float a = fmaf(w[2][o], xv[u][2], fmaf(w[1][o], xv[u][1], fmaf(w[0][o], xv[u][0], b[o])));
          v[e] = fmaxf(a, 0.2f * a);
        }
        h[j] = pack_h2<F16>(v[0], v[1]);
      }
      *reinterpret_cast<uint4*>(out + pix * C + g * 8) = pk;
    }
  }
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

int sg2_first_conv_h16(const void* x, int layout, const float* w3, const float* bias, h16* out, int64_t n, int S, int C, int f16,
                       cudaStream_t s) {
  const int64_t n_pix = n * S * S;
  if (n_pix == 0) return 0;
  SDG_REQUIRE((S & (S - 1)) == 0 && (C & (C - 1)) == 0 && C >= 8 && C <= 2048, SDG_E_UNSUPPORTED, "sg2_first_conv: S=%d C=%d", S, C);
  const int c8s = ilog2(C / 8), hws = 2 * ilog2(S);
  const int grid = stream_grid(n_pix * (C / 8), 256);
  if (f16) { SDG_LAUNCH(sg2_first_conv_kernel<true>, grid, 256, 0, s, x, layout, w3, bias, out, n_pix, hws, c8s, C); }
  else { SDG_LAUNCH(sg2_first_conv_kernel<false>, grid, 256, 0, s, x, layout, w3, bias, out, n_pix, hws, c8s, C); }
  return 0;
}

// Blur on 16-bit NHWC (Blur.forward, stylegan2.py:75-90: upfirdn2d with outer([1,3,3,1])/64, zero padding `pad`), separable
// and register-sliding: a thread owns (output column x, 8 channels) and marches down `rows` output rows; per input row it
// filters horizontally (4 loads of 16 B, neighbours hit L1) and keeps the last four filtered rows in registers, so every
// input pixel leaves DRAM once and every output costs 4 (ST = 1) or 8 (ST = 2) L1 loads instead of 16.
// ST = 2 evaluates only the even blur outputs: all that the stride-2 1x1 skip conv reads (stylegan2.py:575-590).
// column taps of one thread: loop-invariant pointers (clamped inside the image) and weights (zero outside)
struct BlurCols {
  const h16* p[4];
  float w[4];
};

template <bool F16>
__device__ __forceinline__ void blur_hrow(const BlurCols& bc, int iy, int H, int64_t row_stride, float (&o)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = 0.f;
  if (iy < 0 || iy >= H) return;
  const int64_t ro = (int64_t)iy * row_stride;
  // the four loads are issued unconditionally (clamped address, zero weight outside the image) so they overlap
  uint4 raw[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) raw[b] = __ldg(reinterpret_cast<const uint4*>(bc.p[b] + ro));
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint32_t* wv = reinterpret_cast<const uint32_t*>(&raw[b]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_h2<F16>(wv[j]);
      o[2 * j] = fmaf(bc.w[b], t.x, o[2 * j]);
      o[2 * j + 1] = fmaf(bc.w[b], t.y, o[2 * j + 1]);
    }
  }
}

template <bool F16, int ST>
__global__ void __launch_bounds__(256, 4)      // <= 64 registers: occupancy, not issue rate, bounds this kernel (ncu: 35 % warps)
blur_h16_kernel(const h16* __restrict__ in, h16* __restrict__ out, int H, int W, int C, int Ho, int Wo, int pad, int c8_shift,
                int x_tiles, int y_segs, int rows) {
  const int C8 = 1 << c8_shift;
  const int cg = threadIdx.x & (C8 - 1);
  int b = blockIdx.x;
  const int xt = b % x_tiles; b /= x_tiles;
  const int ys = b % y_segs;
  const int64_t n = b / y_segs;
  const int x = xt * (256 >> c8_shift) + (threadIdx.x >> c8_shift);
  if (x >= Wo) return;
  const h16* img = in + n * H * W * (int64_t)C + cg * 8;
  const int64_t row_stride = (int64_t)W * C;
  const int y0 = ys * rows, y1 = min(Ho, y0 + rows);
  h16* dst = out + ((n * Ho + y0) * Wo + x) * (int64_t)C + cg * 8;
  const int64_t dst_stride = (int64_t)Wo * C;
  BlurCols bc;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int ix = x * ST - pad + t;
    const bool ok = ix >= 0 && ix < W;
    bc.w[t] = ok ? ((t == 0 || t == 3) ? 0.125f : 0.375f) : 0.f;
    bc.p[t] = img + (int64_t)(ok ? ix : 0) * C;
  }
  float h0[8], h1[8], h2[8], h3[8];
  // filtered input rows y*ST - pad + {0,1,2,3} feed output row y.  The four row buffers rotate roles by NAME (the loop body is
  // written out for one full rotation) so no register is ever copied.
  auto emit = [&](const float (&a)[8], const float (&b2)[8], const float (&c)[8], const float (&d)[8]) {
    uint4 pk;
    uint32_t* hp = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // written out as the compiler contracts it, so that blur_tma.cu can match it bit for bit
      const float e = fmaf(0.125f, a[2 * j] + d[2 * j], 0.375f * (b2[2 * j] + c[2 * j]));
      const float f = fmaf(0.125f, a[2 * j + 1] + d[2 * j + 1], 0.375f * (b2[2 * j + 1] + c[2 * j + 1]));
      hp[j] = pack_h2<F16>(e, f);
    }
    *reinterpret_cast<uint4*>(dst) = pk;
    dst += dst_stride;
  };
  blur_hrow<F16>(bc, y0 * ST - pad, H, row_stride, h0);
  blur_hrow<F16>(bc, y0 * ST - pad + 1, H, row_stride, h1);
  if (ST == 1) {
    blur_hrow<F16>(bc, y0 - pad + 2, H, row_stride, h2);
    int iy = y0 - pad + 3;                       // next input row to filter
    for (int y = y0; y < y1; y += 4, iy += 4) {
      blur_hrow<F16>(bc, iy, H, row_stride, h3); emit(h0, h1, h2, h3);
      if (y + 1 >= y1) break;
      blur_hrow<F16>(bc, iy + 1, H, row_stride, h0); emit(h1, h2, h3, h0);
      if (y + 2 >= y1) break;
      blur_hrow<F16>(bc, iy + 2, H, row_stride, h1); emit(h2, h3, h0, h1);
      if (y + 3 >= y1) break;
      blur_hrow<F16>(bc, iy + 3, H, row_stride, h2); emit(h3, h0, h1, h2);
    }
  } else {
    int iy = y0 * 2 - pad + 2;
    for (int y = y0; y < y1; y += 2, iy += 4) {
      blur_hrow<F16>(bc, iy, H, row_stride, h2); blur_hrow<F16>(bc, iy + 1, H, row_stride, h3); emit(h0, h1, h2, h3);
      if (y + 1 >= y1) break;
      blur_hrow<F16>(bc, iy + 2, H, row_stride, h0); blur_hrow<F16>(bc, iy + 3, H, row_stride, h1); emit(h2, h3, h0, h1);
    }
  }
}

int blur_h16(const h16* in, h16* out, int64_t n, int H, int W, int C, int pad, int stride, int f16, cudaStream_t s) {
  const int Ho = (H + 2 * pad - 4) / stride + 1, Wo = (W + 2 * pad - 4) / stride + 1;
  if (n == 0 || Ho <= 0 || Wo <= 0) return 0;
  SDG_REQUIRE((C & (C - 1)) == 0 && C >= 8 && C <= 2048, SDG_E_UNSUPPORTED, "blur_h16: C=%d must be a power of two in 8..2048", C);
  SDG_REQUIRE(stride == 1 || stride == 2, SDG_E_UNSUPPORTED, "blur_h16: stride=%d", stride);
  {
    const char* e = getenv("SDG_BLUR_TMA");            // read per call: the tests switch it inside one process
    const int variant = e ? atoi(e) : 1;
    if (variant > 0 && blur_tma_applies(H, W, C, stride)) return blur_tma(in, out, n, H, W, C, pad, stride, f16, variant, s);
  }
  const int c8s = ilog2(C / 8);
  const int xpb = 256 >> c8s;
  const int x_tiles = (int)cdiv(Wo, xpb);
  const int rows = Ho >= 64 ? 32 : (Ho >= 16 ? 8 : Ho);
  const int y_segs = (int)cdiv(Ho, rows);
  const int64_t blocks = (int64_t)x_tiles * y_segs * n;
  SDG_REQUIRE(blocks < (1LL << 31), SDG_E_UNSUPPORTED, "blur_h16: grid too large");
#define SDG_BLUR(F, ST) SDG_LAUNCH((blur_h16_kernel<F, ST>), (unsigned)blocks, 256, 0, s, in, out, H, W, C, Ho, Wo, pad, c8s, x_tiles, y_segs, rows)
  if (f16 && stride == 1) { SDG_BLUR(true, 1); }
  else if (f16) { SDG_BLUR(true, 2); }
  else if (stride == 1) { SDG_BLUR(false, 1); }
  else { SDG_BLUR(false, 2); }
#undef SDG_BLUR
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Split-precision operands for the 4x4 tail (final conv, EqualLinear 8192 -> 512).  The last stages feed a 512-term dot
// product that cancels to a small logit, so their 16-bit roundings would dominate the network's error; instead x = hi + lo
// (two 16-bit values, exact to ~2^-22) and W = Wh + Wl, and the GEMM runs over K' = 3K as
//   [xh | xl | xh] . [Wh | Wh | Wl]^T  =  xh.Wh + xl.Wh + xh.Wl      (the dropped xl.Wl term is ~2^-22 relative)
// on the same tcgen05 kernel (fp32 accumulation in TMEM).  3x the MACs of 0.1 % of the network.
// ---------------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ void split_hilo(float v, h16& hi, h16& lo) {
  const uint32_t ph = pack_h2<F16>(v, 0.f);
  const float vh = unpack_h2<F16>(ph).x;
  hi = (h16)(ph & 0xffffu);
  lo = (h16)(pack_h2<F16>(v - vh, 0.f) & 0xffffu);
}

// in fp32 [rows][C] -> out 16-bit [rows][3C] = [hi | lo | hi]
template <bool F16>
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float* __restrict__ in, h16* __restrict__ out, int64_t total, int C, int* ovf) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    h16 hi, lo;
    if (F16 && !(fabsf(in[i]) <= kF16Max)) range_flag_set(ovf, SDG_RANGE_ACT);
    split_hilo<F16>(in[i], hi, lo);
    h16* o = out + r * 3 * C;
    o[c] = hi; o[C + c] = lo; o[2 * C + c] = hi;
  }
}

int split3_rows_h16(const float* in, h16* out, int64_t rows, int C, int f16, cudaStream_t s) {
  const int64_t total = rows * C;
  if (total == 0) return 0;
  if (f16) { SDG_LAUNCH(split3_rows_kernel<true>, stream_grid(total, 256), 256, 0, s, in, out, total, C, t_range_flag); }
  else { SDG_LAUNCH(split3_rows_kernel<false>, stream_grid(total, 256), 256, 0, s, in, out, total, C, t_range_flag); }
  return 0;
}

// wb[o][(g*C + c) * 3 -> g*3C + {0, C, 2C} + c] = {Wh, Wh, Wl} of w(o, g, c) * mul, where (g, c) index the K axis as groups
// of C channels: conv mode (linear = 0): g = tap, w = W[o][c][tap] with cin_w input channels and G = 9 taps;
// linear mode: g = pixel p, w = W[o][c*G + p] (EqualLinear on an NCHW-flattened map, activations NHWC)
template <bool F16>
__global__ void __launch_bounds__(256)
pack_split3_kernel(const float* __restrict__ W, float mul, h16* __restrict__ wb, int O, int C, int G, int cin_w, int linear,
                   int* ovf) {
  const int64_t K = (int64_t)G * C, total = (int64_t)O * K;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int o = (int)(i / K);
    const int k = (int)(i - (int64_t)o * K);
    const int g = k / C, c = k - g * C;
    const float w = (linear ? W[(int64_t)o * K + (int64_t)c * G + g] : W[((int64_t)o * cin_w + c) * G + g]) * mul;
    if (F16 && !(fabsf(w) <= kF16Max)) range_flag_set(ovf, SDG_RANGE_WEIGHT);
    h16 hi, lo;
    split_hilo<F16>(w, hi, lo);
    h16* d = wb + (int64_t)o * 3 * K + (int64_t)g * 3 * C;
    d[c] = hi; d[C + c] = hi; d[2 * C + c] = lo;
  }
}

int pack_split3_h16(const float* W, float mul, h16* wb, int O, int C, int G, int cin_w, int linear, int f16, cudaStream_t s) {
  const int64_t total = (int64_t)O * C * G;
  if (f16) { SDG_LAUNCH(pack_split3_kernel<true>, stream_grid(total, 256), 256, 0, s, W, mul, wb, O, C, G, cin_w, linear, t_range_flag); }
  else { SDG_LAUNCH(pack_split3_kernel<false>, stream_grid(total, 256), 256, 0, s, W, mul, wb, O, C, G, cin_w, linear, t_range_flag); }
  return 0;
}

}  // namespace sdg
