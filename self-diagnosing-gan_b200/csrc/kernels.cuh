// kernels.cuh -- internal launch API between the discriminator engine (dnet.cu) and the kernel files.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace sdg {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_LRELU_SQRT2 = 3 };   // LRELU: slope 0.2 (mnist.py:164)

// 16-bit storage of the tensor-core path: IEEE fp16 (F16 = true) or bfloat16; arithmetic is fp32.
typedef uint16_t h16;

template <bool F16>
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
  if (F16) {
    __half2 h = __floats2half2_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// pack with ReLU in the conversion itself (cvt.rn.relu.{f16x2,bf16x2}.f32: negative inputs become +0)
template <bool F16>
__device__ __forceinline__ uint32_t pack_relu_h2(float x, float y) {
  uint32_t r;
  if (F16) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
  return r;
}

// fp16 range guard: bit 0 = an ACTIVATION left the fp16 range when it was stored as a 16-bit operand, bit 1 = a packed WEIGHT
// did.  The flag is sticky device memory owned by the caller (sdg_ctx_set_range_flag); setting it is the rare path.
__device__ __forceinline__ void range_flag_set(int* flag, int bit) {
  if (flag) atomicOr(flag, bit);
}
// true if either half of a packed fp16x2 word is inf / NaN (exponent all ones): the carry of +1 at the exponent's LSB
__device__ __forceinline__ uint32_t f16x2_nonfinite_bits(uint32_t pk) { return ((pk & 0x7C007C00u) + 0x04000400u) & 0x80008000u; }

template <bool F16>
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  if (F16) return __half22float2(*reinterpret_cast<__half2*>(&v));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

// ---- conv_fp32.cu: IEEE fp32 CUDA-core path (NHWC activations) ---------------------------------
// wp: packed [ks*ks*Cin][Cout] (k = (ky*ks+kx)*Cin + c); pad < 0 means ks/2
int conv_fp32(const float* in, const float* wp, const float* bias, float* out, int64_t n, int H, int W, int Cin,
              int Cout, int ks, int stride, int pre_act, int post_act, cudaStream_t s, int pad = -1);

// ---- sg2_fp32.cu: StyleGAN2 discriminator pieces (fp32) ------------------------------------------------------------
// out[n,H+2*pad-3,...] = upfirdn2d(in, outer([1,3,3,1])/64, pad=(pad,pad))      (Blur, stylegan2.py:75-90)
int blur_fp32(const float* in, float* out, int64_t n, int H, int W, int C, int pad, cudaStream_t s);
int add_div_sqrt2_fp32(const float* a, const float* b, float* out, int64_t total, cudaStream_t s);   // (a + b) / sqrt(2)
// minibatch-stddev (stylegan2.py:662-670) over consecutive reference batches of `batch` samples, appended as channel C:
// in [n,HW,C] -> out [n,HW,C+1]
int minibatch_stddev_cat_fp32(const float* in, float* out, float* sd_scratch, int64_t n, int batch, int HW, int C,
                              cudaStream_t s);
// the same statistic without the concat: sd_sample[b] = the stddev-channel value sample b would see
int minibatch_stddev_fp32(const float* in, float* sd_sample, int64_t n, int batch, int HW, int C, cudaStream_t s);
// contribution of ONE extra input channel (index c_extra of cin_w) of a 3x3 pad-1 conv on an S x S map whose value is constant
// over the map: wsum[p*Cout + o] = mul * sum over the taps that stay inside the map at pixel p of W[o][c_extra][tap]
int pack_const_channel_fp32(const float* W, float mul, float* wsum, int Cout, int cin_w, int c_extra, int S, cudaStream_t s);
// 16-bit (tensor-core path) pieces: first 1x1 conv + FusedLeakyReLU from the image (w3 [3][C] fp32, scaled), Blur
int sg2_first_conv_h16(const void* x, int layout, const float* w3, const float* bias, h16* out, int64_t n, int S, int C, int f16,
                       cudaStream_t s);
// blur_h16: out extent = (H + 2*pad - 4) / stride + 1 (stride 2 = only the blur outputs a stride-2 1x1 conv reads)
int blur_h16(const h16* in, h16* out, int64_t n, int H, int W, int C, int pad, int stride, int f16, cudaStream_t s);
// blur_tma.cu: the same operation, bit for bit, as a TMA-fed streaming kernel (C % 64 == 0, images of >= 32 output rows);
// blur_h16 dispatches to it (SDG_BLUR_TMA=0: never, 1..3: variant)
bool blur_tma_applies(int H, int W, int C, int stride);
int blur_tma(const h16* in, h16* out, int64_t n, int H, int W, int C, int pad, int stride, int f16, int variant, cudaStream_t s);
// split-precision tail operands (see sg2_fp32.cu): activations fp32 [rows][C] -> [rows][hi | lo | hi]; weights -> per K group
// of C channels [Wh | Wh | Wl] (conv: G = 9 taps of W[O][cin_w][3][3]; linear: G = HW pixels of W[O][C*HW], NCHW flatten)
int split3_rows_h16(const float* in, h16* out, int64_t rows, int C, int f16, cudaStream_t s);
int pack_split3_h16(const float* W, float mul, h16* wb, int O, int C, int G, int cin_w, int linear, int f16, cudaStream_t s);
// wp[(p*C + c)*O + o] = W[o][c*HW + p] * mul    (EqualLinear on an NCHW-flattened feature map, activations kept NHWC)
int pack_linear_nchw_fp32(const float* W, float mul, float* wp, int O, int C, int HW, cudaStream_t s);
// out = (pool_a ? avgpool2(a) : a) + (pool_b ? avgpool2(relu_b ? relu(b) : b) : ...); b may be null
int combine_fp32(const float* a, int pool_a, const float* b, int pool_b, int relu_b, float* out, int64_t n,
                 int Ho, int Wo, int C, cudaStream_t s);
int prep_input_fp32(const void* x, int layout, float* out_nhwc, int64_t n, int H, int W, cudaStream_t s);
// logits[i] = bias + sum_c w[c] * sum_hw act(h[i,hw,c])            (SNGAN head: relu, sum-pool, SNLinear)
int head_sumpool_fp32(const float* h, const float* w, const float* bias, float* logits, int64_t n, int HW, int C,
                      int relu, cudaStream_t s);
// logits[i] = bias + sum_j w[j] * h[i,j]                            (DCGAN head: Linear(8192,1))
int head_dot_fp32(const float* h, const float* w, const float* bias, float* logits, int64_t n, int L,
                  cudaStream_t s);

// ---- pack.cu: spectral-norm sigma (eval semantics) and weight packing ---------------------------
struct SnLayer {
  const float* W;     // [cout, K] torch layout flattened (K = cin*ks*ks)
  const float* u;     // [cout]
  int cout, K;
  float* v;           // scratch [K]
  float* t;           // scratch [cout]
  float* vp;          // scratch [sn_slices()][K]: row-slice partials of u.W
};
int sn_slices();
int sn_sigmas(const SnLayer* layers_dev, const SnLayer* layers_host, int n_layers, float* sigma_dev,
              cudaStream_t s);
// wp[(tap*Cin + c)*Cout + o] = W[o][c][tap] * scale[o] / sigma   (scale may be null; sigma may be null)
int pack_conv_fp32(const float* W, const float* sigma, const float* scale, float* wp, int Cout, int Cin, int ks,
                   cudaStream_t s, float mul = 1.0f);
// wb[o*ld + col0 + k] = h16(W[o][c][tap] * scale[o] / sigma) at k = tap*Cin + c, zero padded up to Kpad columns
// cin_w > 0: the weight tensor has cin_w input channels: cin_w >= Cin packs only the first Cin, cin_w < Cin pads the packed
// layout's channels cin_w..Cin-1 with zeros (DCGAN: 16 / 32 real channels inside a 64-channel TMA chunk)
int pack_conv_h16(const float* W, const float* sigma, const float* scale, h16* wb, int Cout, int Cin, int Kpad,
                  int ks, int f16, int ld, int col0, cudaStream_t s, float mul = 1.0f, int cin_w = 0);
// pooled-3x3 weights for the 4x4 stride-2 form: wb[o*ld + (a*4+b)*Cin + c] = 0.25 * sum of W[o][c][ky][kx] / sigma over
// ky in {a-1,a}, kx in {b-1,b} (valid taps); shortcut: wb[o*ld + col0 + t*sc_pad + c] = 0.25 * Wsc[o][c] / sigma_sc, t = 0..3
int pack_pool4_h16(const float* W, const float* sigma, h16* wb, int Cout, int Cin, int f16, int ld, cudaStream_t s);
int pack_pool4_sc_h16(const float* Wsc, const float* sigma, h16* wb, int Cout, int Csc, int sc_pad, int f16, int ld, int col0,
                      cudaStream_t s);
// super-pixel expansion of a packed 16-bit weight matrix (Cout = 64 layers): src [Cout][ty * txs * Cin] ->
// dst [2*Cout][ty * (txs + shift) * Cin]; row par*Cout + c holds src row c shifted right by par*shift taps (zeros elsewhere):
// output pixels 2x and 2x+1 of a conv with x-stride `shift` share one window of txs + shift input columns
int pack_superpix_h16(const h16* src, h16* dst, int Cout, int Cin, int ty, int txs, int shift, cudaStream_t s);
int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s);          // out = a + b
// several small vector operations in one launch: out = (a [+ b]) [/ sigma[0]]; b and sigma may be null.  Jobs must not read
// each other's outputs.
struct VecJob { float* out; const float* a; const float* b; const float* sigma; int n; int pad; };
constexpr int kMaxVecJobs = 40;
int vec_jobs(const VecJob* jobs, int n, cudaStream_t s);
// several 16-bit weight packs (W / sigma) in one launch: PACK_CONV = pack_conv_h16 (scale = null, mul = 1, cin_w = Cin; Kpad, taps,
// col0 as there), PACK_POOL4 = pack_pool4_h16, PACK_POOL4_SC = pack_pool4_sc_h16 (Cin = Csc, Kpad = sc_pad).  total = elements.
enum { PACK_CONV = 0, PACK_POOL4 = 1, PACK_POOL4_SC = 2 };
struct PackJob { const float* W; const float* sigma; h16* wb; int type, Cin, Kpad, taps, ld, col0, total, pad; };
constexpr int kMaxPackJobs = 16;
int pack_jobs(const PackJob* jobs, int n, int f16, cudaStream_t s);
int scale_vec(const float* in, const float* sigma, float* out, int n, cudaStream_t s);   // out = in / sigma
// BatchNorm(eval) folding: scale[o] = gamma/sqrt(var+eps), shift[o] = beta - mean*scale
// mul scales both outputs (DCGAN tensor-core path: 1/sqrt(2) cancels the gain of the FusedLeakyReLU epilogue)
int bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
            float* shift, int C, cudaStream_t s, float mul = 1.0f);
// DCGAN conv1 on the tensor-core path: out[n,S/2,S/2,64] 16-bit = leaky_relu(conv3x3 stride 2 pad 1 (normalise(x)), 0.2) in
// channels 0..15, zeros in 16..63 (one 64-channel TMA chunk for the next layer); wp = pack_conv_fp32 layout [27][16]
int dcgan_first_conv_h16(const void* x, int layout, const float* wp, h16* out, int64_t n, int S, int f16, cudaStream_t s);
// DCGAN fc weight [C*HW] (NCHW flatten) -> NHWC flatten order
int permute_fc(const float* w, float* out, int C, int HW, cudaStream_t s);

// ---- conv_tc.cu / conv_first.cu: tcgen05 implicit-GEMM convolutions (16-bit in, fp32 TMEM accumulate) ----
// One fused residual-block stage: conv (3x3 pad 1 or 1x1, stride 1) [+ the block's 1x1 shortcut conv as extra K
// columns] [+ 3-FMA shortcut from the network input] [2x2 avg-pool] [+ identity residual] -> up to three outputs.
struct TcConv {
  const h16* in = nullptr;        // [n,H,W,Cin] NHWC, Cin % 64 == 0
  const h16* wb = nullptr;        // [Cout][taps*Cin + sc_C] K-major, k = tap*Cin + c, then the shortcut columns
  const float* bias = nullptr;    // [Cout] (conv bias + shortcut bias) or null
  int64_t n = 0;
  int H = 0, W = 0, Cin = 0, Cout = 0, taps = 9;
  const h16* sc_in = nullptr;     // [n,H,W,sc_C]: input of the block's 1x1 shortcut conv (same resolution as `in`)
  int sc_C = 0;
  int sc_sep = 0;                 // 1: sc_in is at the OUTPUT resolution [n,Ho,Wo,sc_C] and its 1x1 conv (the last sc_C columns of
                                  // wb) goes to a second accumulator added AFTER the activation (StyleGAN2 ResBlock skip)
  int pool = 0;                   // avg_pool2d(., 2) of conv (+ shortcut conv) before the adds below
  int pool4 = 0;                  // with pool: run conv3x3 + avg_pool2d as the algebraically equal 4x4 stride-2 conv
                                  // (16/36 of the MACs); wb then holds 16 taps x Cin (+ 4 taps x sc_C), see pack_pool4_h16
  int stride = 1;                 // 2: explicit stride-2 conv; H, W are then the OUTPUT grid and in_H, in_W the input extent
  int in_H = 0, in_W = 0;
  int no_pad = 0;                 // 1: taps start at the output pixel (padding 0) instead of one pixel before ("same")
  int act = 0;                    // 1: FusedLeakyReLU on (acc + bias): leaky_relu(., 0.2) * sqrt(2), before the residual
  float out_scale = 1.0f;         // final multiplier after the residual (StyleGAN2 ResBlock: 1/sqrt(2))
  // general tap grid (super-pixel forms of Cout = 64 layers): taps_y x taps_x taps starting at (offy, offx) relative to
  // (sy * y, sx * x) of the GEMM pixel (y, x); the GEMM grid is H rows x grid_w columns, the input in_H x in_W.
  // wb = [Cout][taps_y * taps_x * Cin] in (ty, tx, c) order.  superpix = 1: a GEMM pixel is two adjacent output pixels
  // (channel 64 + c = pixel 2x+1): only the image shortcut needs to know.
  int general = 0, grid_w = 0, taps_x = 0, taps_y = 0, sx = 1, sy = 1, offx = 0, offy = 0, superpix = 0;
  int img_up = 0;                 // general form: the GEMM rows are at half the resolution of `img` (pooled stage)
  int gemm = 0;                   // 1: plain GEMM rows: in = [W rows][Cin] (H = 1, any W >= 1), taps = 1  (EqualLinear)
  const float* sd = nullptr;      // [n] per-sample scalar of a spatially constant extra input channel (minibatch-stddev) ...
  const float* sd_w = nullptr;    // ... and its summed weights [H*W][Cout]: v += sd[n] * sd_w[pixel][o] before the activation
  // fused SNGAN head (ReLU -> sum over H,W -> SNLinear) for the LAST block: head_out[n] += head_b + sum_c head_w[c] * sum_px
  // relu(v); head_out must be zeroed by the caller; Cout = 128 and H*W in {32, 64} (at most two warps add per image, so the
  // fp32 result is order-independent).  No other output is required in this mode.
  const float* head_w = nullptr;
  const float* head_b = nullptr;
  float* head_out = nullptr;
  const float* res_f32 = nullptr; // identity shortcut [n,Ho,Wo,Cout] fp32
  const h16* res_h16 = nullptr;   // ... or as a 16-bit tensor (role-swapped kernel only; conv_tc rejects it elsewhere)
  int res_relu = 0;               // rectify the identity shortcut (mimicry's in-place ReLU aliasing)
  const void* img = nullptr;      // network input (DBlockOptimized: shortcut = Wsc3 . avg_pool2d(img) at pooled res)
  int img_layout = 0;
  const float* sc_w3 = nullptr;   // [Cout][3] fp32
  h16* out_relu = nullptr;        // relu(v) [n,Ho,Wo,Cout] 16-bit
  h16* out_raw = nullptr;         // v 16-bit
  float* out_f32 = nullptr;       // v fp32
};
int conv_tc_init(int device);
int conv_tc(const TcConv& args, int f16, cudaStream_t s);
bool conv_tc_swap_active();
void conv_tc_set_pair(int on);   // 1 (default): Cout = 128 3x3 layers run on the CTA-pair (cta_group::2) kernel
// out = relu(conv3x3(normalise(x)) + bias): x uint8 NHWC or fp32 NCHW [n,3,S,S]; wb [Cout][64] (k = tap*3+c, 27 real)
int first_conv_init();
// superpix = 1 (Cout = 64): wb is the [128][64] super-pixel weight tile of pack_first_superpix_h16
int first_conv(const void* x, int layout, const h16* wb, const float* bias, h16* out, int64_t n, int S, int Cout,
               int f16, cudaStream_t s, int superpix = 0);
// conv_b1fused.cu: block 1 of SNGAN-32 (ch = 128) / SNGAN-64 (ch = 64) -- c1 -> relu -> c2 in the 4x4 stride-2 form + image
// shortcut -> relu -- in ONE kernel, relu(c1(x)) kept in shared memory.  x uint8 [n,S,S,3], S = 32 | 64; w1 [ch][64]
// (first_conv's operand); w2f [ch][b1_fused_w2_ld(ch)]: pack_pool4_h16's 16 ch columns followed by the shortcut chunk
// b1_fused_pack writes (fp32 W_sc / sigma [ch][3] and bias2 = c2 + shortcut bias as 16-bit hi / lo columns; w2 = null when
// the main columns were packed in place); out_relu [n,S/2,S/2,ch]; dbg_t [n,S,S,ch] or null
int b1_fused_init();
int b1_fused_w2_elems(int ch);
int b1_fused_w2_ld(int ch);
int b1_fused_pack(const h16* w2, const float* sc_w3, const float* bias2, h16* w2f, int ch, int f16, cudaStream_t s);
int b1_fused(const void* x, const h16* w1, const float* b1, const h16* w2f, h16* out_relu, h16* dbg_t, int64_t n, int ch, int f16,
             cudaStream_t s);
// first-conv weights in super-pixel form: wb[(par*Cout + o)*64 + (ky*4 + j)*3 + c] = W[o][c][ky][j - par] / sigma for
// 0 <= j - par <= 2, zero elsewhere (columns 36, 37 receive the bias in the kernel)
int pack_first_superpix_h16(const float* W, const float* sigma, h16* wb, int Cout, int f16, cudaStream_t s);

}  // namespace sdg
