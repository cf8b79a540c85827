// conv_tc.cu -- tcgen05 implicit-GEMM convolutions with fused block epilogues (the throughput path).
//
// Replaces, for one residual-block stage of the discriminators on the path
//   * torch-mimicry DBlock / DBlockOptimized of SNGANDiscriminator32/64 (call site trainer.py:150):
//       F.conv2d(x, W/sigma, b, stride 1, pad ks/2)                       (3x3 or 1x1)
//     [ + F.conv2d(s, Wsc/sigma_sc, bsc)   the block's 1x1 shortcut conv, folded in as extra K columns ]
//     [ + Wsc3 . avg_pool2d(x_img)         DBlockOptimized's 3-channel shortcut, 3 FMAs in the epilogue ]
//     [ avg_pool2d(., 2) ]  [ + identity shortcut ]  [ ReLU for the next conv ]  [ ReLU -> sum-pool -> SNLinear head ]
//   * ConvLayer / ResBlock of the StyleGAN2 discriminator (diagan/models/stylegan2.py:553-616):
//       EqualConv2d (stride 1 pad 1, or stride 2 pad 0 on the blurred tensor) + FusedLeakyReLU
//     [ + skip 1x1 conv in a second accumulator ]  [ * 1/sqrt(2) ]  [ + the minibatch-stddev channel as a rank-1 term ]
//     and EqualLinear as a plain GEMM.
//
// Four kernels share the TMA traversal and the K schedule; conv_tc() picks by shape:
//   conv_swap_kernel          Cout = 128: D^T[128 ch, 256 px] = W . A^T, one CTA, M = 128 x N = 256 per tcgen05.mma
//   conv_pair_stream_kernel   Cout % 256 == 0 (or skip accumulator): CTA pair, cta_group::2, M = 256 px x N = 256 / 128
//   conv_pair_kernel          Cout = 128, weights resident in shared memory, CTA pair (A/B baseline since the swap kernel)
//   conv_tc_kernel<BN>        one CTA, M = 128 px x N = 64 / 128 (Cout = 64 layers, legacy pooled epilogue, baseline)
//
// Pixel-major GEMM view: D[M = pixels, N = Cout] = A[M, K] * B[N, K]^T, 16-bit operands (fp16 or bf16), fp32 accumulate.
//   * A is never materialised: NHWC activations are read by 4-D TMA boxes of 64 channels x 128 pixels, one
//     box set per (tap, 64-channel chunk), shifted by the tap offset; the halo of the 3x3 window is the
//     TMA out-of-bounds zero fill.  A box lands in shared memory as rows of 128 B, 128B-swizzled = the
//     canonical K-major UMMA operand layout.  Stride-2 forms use the tensor map's traversal stride.
//     Tile geometry: "linear" = 128 consecutive pixels (full-width rows; rows of 256+ pixels split in boxes);
//     "box16" = 8 rows x 16 columns (legacy pooled epilogue with W >= 32).
//   * B (weights [Cout][K] K-major, sigma / equalised-lr scale folded in, shortcut columns appended) by 2-D TMA boxes.
//   * one elected thread issues tcgen05.mma into a double-buffered TMEM accumulator; smem stages are recycled with
//     tcgen05.commit -> mbarrier.
//   * epilogue warps: tcgen05.ld, + bias, activation, shortcuts, then up to three stores: ReLU'd 16-bit (operand of the
//     next conv), raw 16-bit, raw fp32 (residual stream / head).
// Persistent CTAs (one per SM), static tile schedule, warp-specialised: warp 0 TMA producer, warp 1 MMA
// issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#include <cstdlib>
#include <cstring>

#include "tc_ptx.cuh"

namespace sdg {

constexpr int TC_BM = 128;            // pixels per tile (UMMA M)
constexpr int TC_BK = 64;             // channels per stage (128 B of 16-bit = one swizzle row)
constexpr int TC_STAGES = 6;
constexpr int TC_THREADS = 256;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;     // 16 KB
constexpr int TC_TMEM_COLS = 256;                 // 2 accumulator stages x up to 128 fp32 columns
constexpr int TC_MAX_COUT = 1024;

// Tap-shared operand schedule of conv_swap_shared_kernel: the pixel operand is loaded once per (64-channel chunk, column
// variant) as a "copy" of rows+halo image rows, and the taps that differ only by a ROW shift read it through a shifted
// descriptor start address (whole 1024-byte swizzle atoms, so no re-layout).
struct ShCopy {
  signed char src;            // 0: the conv operand (map_c), 1: the folded shortcut's tensor (map_s)
  signed char chunk;          // 64-channel chunk of that tensor
  signed char col0, row0;     // box origin in input pixel coordinates
  signed char rows;           // box rows (output rows + halo)
  signed char ntaps;          // taps served by this copy (1..3)
  short wk[3];                // tap j: 64-column chunk of the weight matrix
  signed char roff[3];        // tap j: row shift inside the copy
};
struct ShSched {
  int n_copies;               // copies per tile, in consumption order
  int row_bytes;              // bytes of one copy row (= one row shift of the descriptor start): tile row pixels x 128
  int perm;                   // 0: the tile is one 16 x 16 grid in natural order; 4: four 8 x 8 grids, pixel order (row, image, column)
  int pitch;                  // bytes between the copy buffers (largest copy, a multiple of 1024)
  int w_stages;               // weight ring depth (4, or 3 when the copies need the room)
  ShCopy cp[24];
};

struct TcParams {
  ShSched sh;
  int H, W, Cin, Cout, taps;
  int kchunks;               // Cin / 64
  int sc_chunks;             // extra shortcut K iterations read through map_s (0 = none) = sc_taps * sc_kchunks
  int sc_kchunks;            // shortcut channels / 64
  int tw;                    // taps per kernel row (3: 3x3, 4: pooled-3x3-as-4x4, 1: 1x1)
  int cs;                    // input coordinate scale (rows; columns too unless csx differs): 2 for the stride-2 forms, else 1
  int csx;                   // input coordinate scale of the columns (4 in the super-pixel form of a stride-2 conv)
  int img_up;                // epilogue pixels are at pooled resolution of `img` (stride-2 form)
  int toff;                  // offset of tap 0 relative to the output pixel (times cs): -1 for "same" padding, 0 for none
  int toffx;                 // the same for the columns
  int superpix;              // GEMM pixel = two horizontally adjacent output pixels of a Cout = 64 layer (channels 64..127 = the
                             // odd pixel): only the image-shortcut table of the role-swapped epilogue needs to know
  int act;                   // 1: FusedLeakyReLU on (acc + bias) before the residual
  float out_scale;           // final multiplier (StyleGAN2 ResBlock: 1/sqrt(2))
  int wide;                  // W > 128: a tile is 128 consecutive pixels of one row (tiles_x per row)
  int sc_sep;                // the shortcut K iterations read an OUTPUT-resolution tensor and accumulate into a second TMEM
                             // accumulator that is added AFTER the activation (StyleGAN2 ResBlock skip branch)
  int box16;                 // tile geometry: 0 linear, 1 = 8 rows x 16 columns (W >= 32 with pooling)
  int bh, tiles_y, bn;       // linear: tile = bn images x bh rows x W columns; box16: tiles_y x tiles_x tiles per image
  int tiles_x;
  int n_tiles;               // Cout / BN
  int pool;                  // 2x2 average pooling in the epilogue
  int res_relu;              // ReLU the identity residual before adding (mimicry's in-place aliasing)
  int img_layout;            // layout of `img` for the 3-FMA shortcut
  int debug_skip_a;          // SDG_DEBUG_SKIP_A=1: pair kernel issues no A loads (timing experiment only, wrong results)
  int debug_skip_epi;        // SDG_DEBUG_SKIP_EPI=1: epilogues only drain the barrier protocol (timing experiment only)
  long long n_images;
  long long m_tiles;
  long long total_pixels;
  const float* bias;         // [Cout] (shortcut bias already added)
  const float* sd;           // per-sample scalar of a spatially constant extra input channel (minibatch-stddev) or null
  const float* sd_w;         // its summed weights [H*W][Cout]
  const float* head_w;       // fused SNGAN head (last block): logit[n] += sum_c head_w[c] * sum_px relu(v[px][c]) (+ head_b once)
  const float* head_b;
  float* head_out;           // [n_images] zero-initialised by the caller; at most two warps contribute per image
  const float* res_f32;      // identity shortcut, fp32 [out pixels][Cout] or null
  const h16* res_h16;        // identity shortcut as a 16-bit tensor (role-swapped kernel only) or null
  const void* img;           // network input for the 3-FMA shortcut (DBlockOptimized) or null
  const float* sc_w3;        // [Cout][3] fp32, W_sc / sigma
  h16* out_relu;             // relu(v) 16-bit or null
  h16* out_raw;              // v 16-bit or null
  float* out_f32;            // v fp32 or null
  int* ovf;                  // fp16 range guard flag (sdg_ctx_set_range_flag) or null
};

__device__ __forceinline__ float norm_px(const void* img, int layout, long long n, int y, int x, int c, int H, int W) {
  if (layout == SDG_LAYOUT_U8_NHWC) {
    float v = __fdiv_rn((float)reinterpret_cast<const uint8_t*>(img)[((n * H + y) * W + x) * 3 + c], 255.0f);
    return __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
  }
  return reinterpret_cast<const float*>(img)[((n * 3 + c) * H + y) * (long long)W + x];
}

// ---------------------------------------------------------------------------------------------------
// k iteration -> which tensor map, which 64-channel chunk, which tap offset (in input pixels)
//   main iterations: (tap, chunk) of the conv operand; 3x3: offsets -1..1; pooled-3x3-as-4x4-stride-2: -1..2
//   then the folded 1x1 shortcut conv: offset 0 (same resolution) or the four taps of a 2x2 stride-2 average
__device__ __forceinline__ void tc_k_iter(const TcParams& p, int it, int main_iters, int tap, int kc, bool& is_sc, int& ch,
                                          int& dy, int& dx) {
  is_sc = it >= main_iters;
  dy = 0; dx = 0; ch = kc;
  if (is_sc) {
    const int j = it - main_iters;
    const int st = j / p.sc_kchunks;
    ch = j - st * p.sc_kchunks;
    if (p.cs == 2 && !p.sc_sep) { dy = st >> 1; dx = st & 1; }
  } else if (p.tw == 3) {
    const int ty3 = (tap * 11) >> 5;            // tap / 3 for tap in 0..8
    dy = ty3 + p.toff;
    dx = tap - 3 * ty3 + p.toffx;
  } else if (p.tw == 4) {
    dy = (tap >> 2) + p.toff;
    dx = (tap & 3) + p.toffx;
  } else if (p.tw > 1) {                        // general tap grid (super-pixel forms: 3x4, 4x6)
    const int ty = tap / p.tw;
    dy = ty + p.toff;
    dx = tap - ty * p.tw + p.toffx;
  }
}

// ---------------------------------------------------------------------------------------------------
// tile geometry: origin (image, row, column) of M tile `mt`
__device__ __forceinline__ void tc_tile_origin(const TcParams& p, long long mt, int& n0, int& y0, int& x0) {
  if (p.box16) {
    const int per_img = p.tiles_x * p.tiles_y;
    n0 = (int)(mt / per_img);
    const int r = (int)(mt - (long long)n0 * per_img);
    y0 = (r / p.tiles_x) * 8;
    x0 = (r % p.tiles_x) * 16;
  } else if (p.wide) {
    const long long row = mt / p.tiles_x;
    x0 = (int)(mt - row * p.tiles_x) * TC_BM;
    y0 = (int)(row % p.H);
    n0 = (int)(row / p.H);
  } else {
    n0 = (int)(mt / p.tiles_y) * p.bn;
    y0 = (int)(mt % p.tiles_y) * p.bh;
    x0 = 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// In-register transpose of an 8 x 8 matrix of float4 held by each group of 8 consecutive lanes (lane i of the group
// holds Q[i][0..7] in t[4j..4j+3]); afterwards lane i holds Q[0..7][i].  Three xor-butterfly stages, 48 shuffles.
// The epilogue uses it so that its fp32 global accesses are coalesced: 8 lanes x 16 B = one full 128-byte line of
// ONE pixel row per group, instead of 32 lanes touching 32 different lines (the LSU wavefront count, not the DRAM
// bandwidth, is what made the residual-stream epilogues slow).
__device__ __forceinline__ void warp_transpose8_f4(float (&t)[32], int lane) {
#pragma unroll
  for (int s = 4; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((j & s) == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = t[4 * j + e], b = t[4 * (j + s) + e];
          const float recv = __shfl_xor_sync(0xffffffffu, up ? a : b, s);
          t[4 * j + e] = up ? recv : a;
          t[4 * (j + s) + e] = up ? b : recv;
        }
      }
    }
  }
}

// (warp_transpose4_u4, the 4 x 4 variant for the 16-bit outputs, lives in tc_ptx.cuh: conv_b1fused.cu uses it too)

// store this warp's 32 rows x 32 columns of 16-bit values (pk = the lane's own row, 16 packed words) coalesced
__device__ __forceinline__ void store_rows16_coalesced(h16* base, long long wpix, long long total_pixels, int Cout, int col0,
                                                       uint32_t (&pk)[16], int lane) {
  warp_transpose4_u4(pk, lane);
  const int g4 = lane >> 2, i4 = lane & 3;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const long long rp = wpix + g4 * 4 + m;
    if (rp < total_pixels)
      *reinterpret_cast<uint4*>(base + rp * Cout + col0 + i4 * 8) = make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
  }
}

// ---------------------------------------------------------------------------------------------------
// epilogue of one 128-pixel x BN tile for one warp (TMEM lane quadrant q): waits for the accumulator, then
// pooling / bias / shortcuts / stores.  tmem_acc = TMEM address of column 0 of this accumulator stage.
template <int BN, bool F16>
__device__ __forceinline__ void tc_epilogue_tile(const TcParams& p, const float* s_bias, const float* s_w3,
                                                 uint32_t tmem_acc, long long mt, int nt, int q, int lane,
                                                 uint32_t acc_full_bar, uint32_t acc_phase, float& vmax) {
  if (p.debug_skip_epi) {
    mbar_wait(acc_full_bar, acc_phase);
    tc_fence_after();
    return;
  }
  const int wc = p.W < 16 ? p.W : 16;             // columns per warp-row (pooling partner stride)
  const bool lin = !p.pool && !p.box16;           // this warp's 32 rows are 32 consecutive output pixels
  const long long wpix = mt * TC_BM + q * 32;     // ... starting here
  const int g8 = lane >> 3, i8 = lane & 7;
  const int HW = p.H * p.W;
  const int Ho = p.pool ? p.H >> 1 : p.H, Wo = p.pool ? p.W >> 1 : p.W;
  {
    // this thread's pixel (n, y, x)
    long long n;
    int y, x, r_img = 0;
    bool valid;
    if (p.box16) {
      const int per_img = p.tiles_x * p.tiles_y;
      n = mt / per_img;
      const int r = (int)(mt - n * per_img);
      x = (r % p.tiles_x) * 16 + (lane & 15);
      y = (r / p.tiles_x) * 8 + 2 * q + (lane >> 4);
      valid = n < p.n_images;
    } else {
      const long long pix = mt * TC_BM + q * 32 + lane;    // 128 consecutive NHW pixels
      n = pix / HW;
      const int r = (int)(pix - n * HW);
      r_img = r;
      y = r / p.W;
      x = r - y * p.W;
      valid = pix < p.total_pixels;
    }
    bool active = valid;
    if (p.pool) active = valid && ((x & 1) == 0) && ((y & 1) == 0);
    const long long opix = (n * Ho + (p.pool ? (y >> 1) : y)) * Wo + (p.pool ? (x >> 1) : x);
    float px[3] = {0.f, 0.f, 0.f};
    if (p.img && active) {
      // avg_pool2d of the normalised network input at this pooled pixel (DBlockOptimized shortcut input)
      const int iy = p.img_up ? 2 * y : y, ix = p.img_up ? 2 * x : x;
      const int iH = p.img_up ? 2 * p.H : p.H, iW = p.img_up ? 2 * p.W : p.W;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        px[c] = (norm_px(p.img, p.img_layout, n, iy, ix, c, iH, iW) + norm_px(p.img, p.img_layout, n, iy, ix + 1, c, iH, iW) +
                 norm_px(p.img, p.img_layout, n, iy + 1, ix, c, iH, iW) + norm_px(p.img, p.img_layout, n, iy + 1, ix + 1, c, iH, iW)) * 0.25f;
    }
    const long long obase = opix * p.Cout + nt * BN;
    if (p.res_f32 && active) {
      // pull this thread's residual row (BN fp32 = BN*4 bytes) towards L2 while the MMAs of the tile still run
#pragma unroll
      for (int b = 0; b < BN * 4; b += 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.res_f32 + obase) + b));
    }
    mbar_wait(acc_full_bar, acc_phase);
    tc_fence_after();
    float head_part = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.pool) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += __shfl_down_sync(0xffffffffu, v[j], 1);
          v[j] += __shfl_down_sync(0xffffffffu, v[j], wc);
          v[j] *= 0.25f;
        }
      }
      if (active) {
        const int cb = nt * BN + c0;
        // shared memory bandwidth is what bounds the MMA mainloop (operand fetch + TMA fill), so the epilogue
        // reads its constants with as few wavefronts as possible: 128-bit broadcast loads
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cb + 4 * g);
          v[4 * g] += b4.x; v[4 * g + 1] += b4.y; v[4 * g + 2] += b4.z; v[4 * g + 3] += b4.w;
        }
        if (p.img) {
          float w3[96];
#pragma unroll
          for (int g = 0; g < 24; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(s_w3 + cb * 3 + 4 * g);
            w3[4 * g] = t.x; w3[4 * g + 1] = t.y; w3[4 * g + 2] = t.z; w3[4 * g + 3] = t.w;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = fmaf(w3[3 * j], px[0], fmaf(w3[3 * j + 1], px[1], fmaf(w3[3 * j + 2], px[2], v[j])));
        }
        if (p.sd) {                        // spatially constant extra channel: sd[n] * (sum of its in-bounds tap weights)
          const float sdv = p.sd[n];
          const float4* ws = reinterpret_cast<const float4*>(p.sd_w + (long long)r_img * p.Cout + cb);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 t = ws[g];
            v[4 * g] = fmaf(sdv, t.x, v[4 * g]); v[4 * g + 1] = fmaf(sdv, t.y, v[4 * g + 1]);
            v[4 * g + 2] = fmaf(sdv, t.z, v[4 * g + 2]); v[4 * g + 3] = fmaf(sdv, t.w, v[4 * g + 3]);
          }
        }
        if (p.act) {                       // FusedLeakyReLU: leaky_relu(x + b, 0.2) * sqrt(2)  (op/fused_act.py:104-116)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (v[j] > 0.f ? v[j] : 0.2f * v[j]) * 1.4142135623730951f;
        }
      }
      if (p.sc_sep) {                      // skip branch accumulated beside the main conv (warp-collective TMEM load)
        uint32_t r2[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r2[j]);
      }
      if (lin && p.res_f32) {
        // residual (identity shortcut / StyleGAN2 skip branch), coalesced: lane (g8, i8) loads float4 #i8 of rows g8*8 + m, then an in-register transpose
        // brings every row's 32 values to the lane that owns the row
        float t[32];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const long long rp = wpix + g8 * 8 + m;
          float4 ld = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rp < p.total_pixels) ld = *reinterpret_cast<const float4*>(p.res_f32 + rp * p.Cout + nt * BN + c0 + i8 * 4);
          t[4 * m] = ld.x; t[4 * m + 1] = ld.y; t[4 * m + 2] = ld.z; t[4 * m + 3] = ld.w;
        }
        warp_transpose8_f4(t, lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += p.res_relu ? fmaxf(t[j], 0.f) : t[j];
      }
      if (p.out_scale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
      }
      if (F16 && active && (p.out_raw || p.out_relu)) {
        // fp16 range guard: the largest value about to be rounded to a 16-bit operand (negative ones vanish under ReLU)
#pragma unroll
        for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, p.out_raw ? fabsf(v[j]) : v[j]);
      }
      if (p.head_out && active) {          // s_w3 holds the head weights in this mode
#pragma unroll
        for (int j = 0; j < 32; ++j) head_part = fmaf(s_w3[nt * BN + c0 + j], fmaxf(v[j], 0.f), head_part);
      }
      if (active) {
        if (p.res_f32 && !lin) {
          const float4* rp = reinterpret_cast<const float4*>(p.res_f32 + obase + c0);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 t = rp[g];
            if (p.res_relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
            v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
          }
        }
        if (p.out_f32 && !lin) {
          float4* op = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
          for (int g = 0; g < 8; ++g) op[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        if (p.out_raw && !lin) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            uint32_t* h = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = pack_h2<F16>(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
            *reinterpret_cast<uint4*>(p.out_raw + obase + c0 + g * 8) = pk;
          }
        }
        if (p.out_relu && !lin) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            uint32_t* h = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              h[j] = pack_h2<F16>(fmaxf(v[g * 8 + 2 * j], 0.f), fmaxf(v[g * 8 + 2 * j + 1], 0.f));
            *reinterpret_cast<uint4*>(p.out_relu + obase + c0 + g * 8) = pk;
          }
        }
      }
      if (lin && p.out_raw) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_h2<F16>(v[2 * j], v[2 * j + 1]);
        store_rows16_coalesced(p.out_raw, wpix, p.total_pixels, p.Cout, nt * BN + c0, pk, lane);
      }
      if (lin && p.out_relu) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_relu_h2<F16>(v[2 * j], v[2 * j + 1]);
        store_rows16_coalesced(p.out_relu, wpix, p.total_pixels, p.Cout, nt * BN + c0, pk, lane);
      }
      if (lin && p.out_f32) {
        // fp32 output, coalesced the same way (rows of invalid pixels are skipped).  Bias and shortcuts were added
        // under `active`, which in linear mode is simply `valid`, so every stored row is complete.
        float t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = v[j];
        warp_transpose8_f4(t, lane);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const long long rp = wpix + g8 * 8 + m;
          if (rp < p.total_pixels)
            *reinterpret_cast<float4*>(p.out_f32 + rp * p.Cout + nt * BN + c0 + i8 * 4) =
                make_float4(t[4 * m], t[4 * m + 1], t[4 * m + 2], t[4 * m + 3]);
        }
      }
    }
    if (p.head_out) {
      // fused head: this warp's 32 pixels belong to ONE image (H*W is 32 or 64); fixed shuffle tree, then one atomic per
      // warp.  With at most two addends into a zeroed slot the fp32 sum does not depend on their order: deterministic.
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) head_part += __shfl_xor_sync(0xffffffffu, head_part, o);
      if (lane == 0 && wpix < p.total_pixels) {
        const long long img = wpix / HW;
        if (wpix - img * HW == 0) head_part += p.head_b[0];
        atomicAdd(p.head_out + img, head_part);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <int BN, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_s, const TcParams p) {
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atom (TMA and UMMA must agree on address bits)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  __shared__ __align__(8) uint64_t bar_full[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[TC_MAX_COUT];
  __shared__ __align__(16) float s_w3[TC_MAX_COUT * 3];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  if (p.img)
    for (int i = threadIdx.x; i < p.Cout * 3; i += TC_THREADS) s_w3[i] = p.sc_w3[i];
  if (p.head_out)
    for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) s_w3[i] = p.head_w[i];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.sc_chunks) tma_prefetch_desc(&map_s);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), 4);       // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  const uint32_t tmem_cols = p.sc_sep ? 2 * TC_TMEM_COLS : TC_TMEM_COLS;
  const uint32_t acc_stride = p.sc_sep ? 2 * BN : BN;      // TMEM columns per accumulator stage
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const long long total_tiles = p.m_tiles * p.n_tiles;
  const int main_iters = p.taps * p.kchunks;
  const int k_iters = main_iters + p.sc_chunks;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % p.n_tiles);
        const long long mt = tile / p.n_tiles;
        int n0, y0, x0;
        tc_tile_origin(p, mt, n0, y0, x0);
        int tap = 0, kc = 0;
        for (int it = 0; it < k_iters; ++it) {
          bool is_sc;
          int dy, dx, ch;
          tc_k_iter(p, it, main_iters, tap, kc, is_sc, ch, dy, dx);
          const CUtensorMap* am = is_sc ? &map_s : &map_a;
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_expect_tx(full, STAGE_BYTES);
          const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
          const int cs = (is_sc && p.sc_sep) ? 1 : p.cs;
          const int csx = (is_sc && p.sc_sep) ? 1 : p.csx;
          tma_load_4d(a_dst, am, full, ch * TC_BK, csx * x0 + dx, cs * y0 + dy, n0);
          tma_load_2d(a_dst + TC_A_BYTES, &map_b, full, it * TC_BK, nt * BN);
          if (++kc == p.kchunks) { kc = 0; ++tap; }
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(TC_BM, BN, F16);
      int stage = 0;
      uint32_t phase = 0;
      long long local = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int acc = (int)(local & 1);
        const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)acc * acc_stride;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + TC_A_BYTES);
          // separate-accumulator shortcut: its K iterations start a fresh accumulation BN columns further
          const bool sep = p.sc_sep && it >= main_iters;
          const uint32_t d_tmem = sep ? d_main + BN : d_main;
          const int it0 = sep ? it - main_iters : it;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // advance 16 elements = 32 B inside the swizzle row: +2 in the 16-byte-unit address field
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it0 | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));       // frees the smem stage when these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&bar_acc_full[acc]));        // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp - 4;                       // TMEM lane quadrant this warp may access
    long long local = 0;
    float vmax = 0.f;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int nt = (int)(tile % p.n_tiles);
      const long long mt = tile / p.n_tiles;
      const int acc = (int)(local & 1);
      const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
      tc_epilogue_tile<BN, F16>(p, s_bias, s_w3, tmem_base + (uint32_t)acc * acc_stride, mt, nt, q, lane,
                                smem_u32(&bar_acc_full[acc]), acc_phase, vmax);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
    }
    if (F16 && vmax > kF16Max) range_flag_set(p.ovf, SDG_RANGE_ACT);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant for Cout = 128 layers (all 3x3 convs of SNGAN-32): two CTAs of a cluster run ONE
// tcgen05.mma.cta_group::2 of M = 256 (128 pixels each) x N = 128, each CTA supplying 64 rows of B.
// Why: the 1-CTA kernel above is bounded by shared-memory bandwidth -- per 64-cycle MMA it reads 4 KB of A and
// 4 KB of B from smem while TMA writes the same amount -- which caps it near 50 % of the tensor peak.  In the pair
// every CTA reads 4 KB + 2 KB per MMA, and its half of the weights (<= 20 k-chunks x 8 KB) stays RESIDENT in shared
// memory for the whole kernel, so the only streaming traffic is the A tile: 160 B/clk -> 96 B/clk of smem traffic
// at full tensor rate.
// Synchronisation: "full" and "accumulator-empty" barriers live in the leader CTA (rank 0) and are signalled by both
// CTAs (TMA complete_tx through .cta_group::2 loads, remote mbarrier arrives); "stage-empty" and "accumulator-full"
// are signalled in both CTAs at once by the leader's multicast tcgen05.commit.
constexpr int PAIR_MAX_KI = 20;
constexpr int PAIR_B_TILE = 64 * TC_BK * 2;        // 8 KB: 64 weight rows x 64 k

template <bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_s, const TcParams p, const int n_stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_b;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[128];
  __shared__ __align__(16) float s_w3[128 * 3];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int main_iters = p.taps * p.kchunks;
  const int k_iters = main_iters + p.sc_chunks;
  const uint32_t a_base = smem_base + (uint32_t)k_iters * PAIR_B_TILE;      // A stages follow the resident weights

  for (int i = threadIdx.x; i < 128; i += TC_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  if (p.img)
    for (int i = threadIdx.x; i < 128 * 3; i += TC_THREADS) s_w3[i] = p.sc_w3[i];
  if (p.head_out)
    for (int i = threadIdx.x; i < 128; i += TC_THREADS) s_w3[i] = p.head_w[i];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.sc_chunks) tma_prefetch_desc(&map_s);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), 8);       // 4 epilogue warps x 2 CTAs
    }
    mbar_init(smem_u32(&bar_b), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_slot), TC_TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const long long cluster_id = blockIdx.x >> 1;
  const long long n_clusters = gridDim.x >> 1;
  const long long pair_tiles = (p.m_tiles + 1) >> 1;    // M = 256 tiles; CTA `rank` owns m tile 2*ct + rank

  if (warp == 0) {
    // ================= TMA producer: this CTA's 128-pixel A tile per k iteration =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long ct = cluster_id; ct < pair_tiles; ct += n_clusters) {
        const long long mt = 2 * ct + rank;
        int n0, y0, x0;
        tc_tile_origin(p, mt, n0, y0, x0);
        int tap = 0, kc = 0;
        for (int it = 0; it < k_iters; ++it) {
          bool is_sc;
          int dy, dx, ch;
          tc_k_iter(p, it, main_iters, tap, kc, is_sc, ch, dy, dx);
          const CUtensorMap* am = is_sc ? &map_s : &map_a;
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[stage]), 0);
          if (p.debug_skip_a) {
            if (leader) mbar_arrive(smem_u32(&bar_full[stage]));
          } else {
            if (leader) mbar_expect_tx(smem_u32(&bar_full[stage]), 2 * TC_A_BYTES);     // both CTAs' A tiles
            tma_load_4d_pair(a_base + stage * TC_A_BYTES, am, full_leader, ch * TC_BK, p.csx * x0 + dx, p.cs * y0 + dy, n0);
          }
          if (++kc == p.kchunks) { kc = 0; ++tap; }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc(256, 128, F16);
      mbar_wait(smem_u32(&bar_b), 0);                  // both halves of the weights are resident
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      long long local = 0;
      for (long long ct = cluster_id; ct < pair_tiles; ct += n_clusters, ++local) {
        const int acc = (int)(local & 1);
        const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(a_base + stage * TC_A_BYTES);
          const uint64_t bdesc = make_sw128_desc(smem_base + (uint32_t)it * PAIR_B_TILE);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
          umma_commit_pair(smem_u32(&bar_empty[stage]), 3);      // frees the stage in both CTAs
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(smem_u32(&bar_acc_full[acc]), 3);       // accumulator complete -> both epilogues
      }
    }
  } else if (warp == 2) {
    // ================= resident weights: this CTA's 64 output channels of every k chunk, loaded once =================
    if (elect_one()) {
      const uint32_t b_leader = mapa_u32(smem_u32(&bar_b), 0);
      if (leader) mbar_expect_tx(smem_u32(&bar_b), 2u * (uint32_t)k_iters * PAIR_B_TILE);
      for (int it = 0; it < k_iters; ++it)
        tma_load_2d_pair(smem_base + (uint32_t)it * PAIR_B_TILE, &map_b, b_leader, it * TC_BK, (int)rank * 64);
    }
  } else if (warp >= 4) {
    // ================= epilogue: this CTA's 128 TMEM lanes =================
    const int q = warp - 4;
    long long local = 0;
    float vmax = 0.f;
    for (long long ct = cluster_id; ct < pair_tiles; ct += n_clusters, ++local) {
      const long long mt = 2 * ct + rank;
      const int acc = (int)(local & 1);
      const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
      tc_epilogue_tile<128, F16>(p, s_bias, s_w3, tmem_base + (uint32_t)(acc * 128), mt, 0, q, lane,
                                 smem_u32(&bar_acc_full[acc]), acc_phase, vmax);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_acc_empty[acc]), 0));
    }
    if (F16 && vmax > kF16Max) range_flag_set(p.ovf, SDG_RANGE_ACT);
  }

  tc_fence_before();
  cluster_sync_all();                                   // every MMA retired, every epilogue done, in both CTAs
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TC_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair kernel with STREAMED weights: one tcgen05.mma.cta_group::2 of M = 256 x N = BN per K step, for layers whose
// weights do not fit in shared memory.  BN = 256 (Cout % 256 == 0: StyleGAN2 and SNGAN-64 bodies): every CTA stages its own
// 128-pixel A tile (16 KB) and HALF of the 256-row B tile (16 KB) per stage, i.e. the same shared-memory traffic as the
// single-CTA N = 128 kernel for twice the MMA work -- half the operand bytes per FLOP, which is what the power-capped
// tensor pipe rewards (cuBLAS runs 256-wide tiles for the same reason).  The accumulator is 2 stages x 256 TMEM columns.
// BN = 128 serves the separate-accumulator skip form (main 128 + skip 128 columns per stage).
// Barrier protocol = conv_pair_kernel: full / accumulator-empty live in the leader, stage-empty / accumulator-full are
// multicast by the leader's tcgen05.commit.
template <int BN, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_pair_stream_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_s, const TcParams p, const int n_stages) {
  constexpr int B_HALF = (BN / 2) * TC_BK * 2;
  constexpr int STAGE = TC_A_BYTES + B_HALF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  __shared__ __align__(8) uint64_t bar_full[8];
  __shared__ __align__(8) uint64_t bar_empty[8];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[TC_MAX_COUT];
  __shared__ __align__(16) float s_w3[128 * 3];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int main_iters = p.taps * p.kchunks;
  const int k_iters = main_iters + p.sc_chunks;
  const uint32_t tmem_cols = (BN == 256 || p.sc_sep) ? 512u : 256u;
  const uint32_t acc_stride = p.sc_sep ? 2 * BN : BN;
  // accumulator stages: two, except N = 256 with the skip accumulator (256 + 256 columns fill TMEM: single-buffered, the
  // epilogue of a tile is then exposed, ~10 % of a K = 5120 mainloop, in exchange for N = 256 instructions)
  const int n_acc = (BN == 256 && p.sc_sep) ? 1 : 2;

  for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  if (p.img)
    for (int i = threadIdx.x; i < 128 * 3; i += TC_THREADS) s_w3[i] = p.sc_w3[i];
  if (p.head_out)
    for (int i = threadIdx.x; i < 128; i += TC_THREADS) s_w3[i] = p.head_w[i];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.sc_chunks) tma_prefetch_desc(&map_s);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), 8);       // 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_slot), tmem_cols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const long long cluster_id = blockIdx.x >> 1;
  const long long n_clusters = gridDim.x >> 1;
  const long long pair_tiles = (p.m_tiles + 1) >> 1;
  const long long total = pair_tiles * p.n_tiles;       // n tile fastest: consecutive work items share the A tiles (L2)

  if (warp == 0) {
    // ================= TMA producer: own A tile + own half of the B tile per k iteration =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = cluster_id; t < total; t += n_clusters) {
        const int nt = (int)(t % p.n_tiles);
        const long long mt = 2 * (t / p.n_tiles) + rank;
        int n0, y0, x0;
        tc_tile_origin(p, mt, n0, y0, x0);
        int tap = 0, kc = 0;
        for (int it = 0; it < k_iters; ++it) {
          bool is_sc;
          int dy, dx, ch;
          tc_k_iter(p, it, main_iters, tap, kc, is_sc, ch, dy, dx);
          const CUtensorMap* am = is_sc ? &map_s : &map_a;
          const int cs = (is_sc && p.sc_sep) ? 1 : p.cs;
          const int csx = (is_sc && p.sc_sep) ? 1 : p.csx;
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[stage]), 0);
          if (leader) mbar_expect_tx(smem_u32(&bar_full[stage]), 2 * STAGE);
          const uint32_t dst = smem_base + stage * STAGE;
          tma_load_4d_pair(dst, am, full_leader, ch * TC_BK, csx * x0 + dx, cs * y0 + dy, n0);
          tma_load_2d_pair(dst + TC_A_BYTES, &map_b, full_leader, it * TC_BK, nt * BN + (int)rank * (BN / 2));
          if (++kc == p.kchunks) { kc = 0; ++tap; }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc(256, BN, F16);
      int stage = 0;
      uint32_t phase = 0;
      long long local = 0;
      for (long long t = cluster_id; t < total; t += n_clusters, ++local) {
        const int acc = (int)(local % n_acc);
        const uint32_t acc_phase = (uint32_t)((local / n_acc) & 1);
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)acc * acc_stride;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * STAGE;
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + TC_A_BYTES);
          const bool sep = p.sc_sep && it >= main_iters;
          const uint32_t d_tmem = sep ? d_main + BN : d_main;
          const int it0 = sep ? it - main_iters : it;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it0 | k) != 0 ? 1u : 0u);
          umma_commit_pair(smem_u32(&bar_empty[stage]), 3);
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(smem_u32(&bar_acc_full[acc]), 3);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: this CTA's 128 TMEM lanes x BN columns =================
    const int q = warp - 4;
    long long local = 0;
    float vmax = 0.f;
    for (long long t = cluster_id; t < total; t += n_clusters, ++local) {
      const int nt = (int)(t % p.n_tiles);
      const long long mt = 2 * (t / p.n_tiles) + rank;
      const int acc = (int)(local % n_acc);
      const uint32_t acc_phase = (uint32_t)((local / n_acc) & 1);
      tc_epilogue_tile<BN, F16>(p, s_bias, s_w3, tmem_base + (uint32_t)acc * acc_stride, mt, nt, q, lane,
                                smem_u32(&bar_acc_full[acc]), acc_phase, vmax);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_acc_empty[acc]), 0));
    }
    if (F16 && vmax > kF16Max) range_flag_set(p.ovf, SDG_RANGE_ACT);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

constexpr int SW_STAGE = 3 * TC_A_BYTES;      // 48 KB
constexpr int SW_STAGES = 4;
constexpr int SW_EPI_WARPS = 8;
constexpr int SW_THREADS = 128 + 32 * SW_EPI_WARPS;
constexpr int kSwapSmem = 1024 + SW_STAGES * SW_STAGE;

template <bool F16>
__device__ __forceinline__ uint16_t to_h16(float v) { return (uint16_t)(pack_h2<F16>(v, 0.f) & 0xffffu); }

// ---------------------------------------------------------------------------------------------------
// epilogue of the role-swapped kernels (shared by conv_swap_kernel and conv_swap_shared_kernel): thread = output channel,
// TMEM columns = the tile's 256 pixels
template <bool F16>
__device__ __forceinline__ void swap_epilogue_loop(const TcParams& p, const uint32_t tmem_base, uint64_t* bar_acc_full,
                                                   uint64_t* bar_acc_empty, float* s_px, float* s_head, const long long tiles,
                                                   const int warp, const int lane) {
    // eight epilogue warps: warp w may only read the TMEM lane quadrant w % 4, so two warps share each 32-channel quadrant
    // and split the tile's 256 pixel columns between them (warps 4..7: columns 0..127, warps 8..11: columns 128..255)
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int c = q * 32 + lane;
    const int et = threadIdx.x - 128;                     // 0..255 among the epilogue threads
    const int CO = p.Cout;                                // 128, or 64: weight rows 64..127 are TMA zero fill (half the MMA idles)
    const bool c_ok = c < CO;
    const float bias = (p.bias && c_ok) ? p.bias[c] : 0.f;
    float w3[3] = {0.f, 0.f, 0.f};
    if (p.img && c_ok) { w3[0] = p.sc_w3[c * 3]; w3[1] = p.sc_w3[c * 3 + 1]; w3[2] = p.sc_w3[c * 3 + 2]; }
    const float hw_c = (p.head_out && c_ok) ? p.head_w[c] : 0.f;
    const int spx_n = p.superpix ? 2 : 1, spx_par = p.superpix ? (c >> 6) : 0;    // warp-uniform: a warp owns 32 channels
    const int HW = p.H * p.W;
    long long local = 0;
    float vmax = 0.f;                                     // fp16 range guard: largest value rounded to a 16-bit operand
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++local) {
      const int acc = (int)(local & 1);
      const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
      const long long P0 = t * 256;
      if (p.img) {
        // avg_pool2d of the normalised network input at the tile's (pooled) pixels: one pixel per epilogue thread
        named_bar_sync(1, SW_EPI_WARPS * 32);             // previous tile's readers are done with s_px
        const int npar = p.superpix ? 2 : 1;
        for (int par = 0; par < npar; ++par) {
          const long long pix = P0 + et;
          float px[3] = {0.f, 0.f, 0.f};
          if (pix < p.total_pixels) {
            const long long n = pix / HW;
            const int r = (int)(pix - n * HW);
            const int y = r / p.W, xg = r - y * p.W;
            const int x = p.superpix ? 2 * xg + par : xg;            // output column of this parity
            const int Wout = p.superpix ? 2 * p.W : p.W;
            const int iy = p.img_up ? 2 * y : y, ix = p.img_up ? 2 * x : x;
            const int iH = p.img_up ? 2 * p.H : p.H, iW = p.img_up ? 2 * Wout : Wout;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
              px[ch] = (norm_px(p.img, p.img_layout, n, iy, ix, ch, iH, iW) + norm_px(p.img, p.img_layout, n, iy, ix + 1, ch, iH, iW) +
                        norm_px(p.img, p.img_layout, n, iy + 1, ix, ch, iH, iW) + norm_px(p.img, p.img_layout, n, iy + 1, ix + 1, ch, iH, iW)) * 0.25f;
          }
          float* d = s_px + (et * npar + par) * 3;
          d[0] = px[0]; d[1] = px[1]; d[2] = px[2];
        }
        named_bar_sync(1, SW_EPI_WARPS * 32);
      }
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
      if (p.debug_skip_epi) {                             // timing experiment only (SDG_DEBUG_SKIP_EPI=1): drain the protocol
        mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
        continue;
      }
      float head_sum = 0.f;
      // The residual does not depend on the accumulator: the 32 loads of a chunk are issued one chunk AHEAD (chunk 0 before
      // the accumulator wait), so their latency hides behind the previous chunk's arithmetic and stores.  (The epilogue,
      // not the MMA, paced the residual layers: 34 % tensor-pipe activity in profiles/r1f_ncu_full_swap_summary.txt.)
      float rnext[32];
      auto load_res = [&](int c0n) {
        if (p.res_f32 && c_ok && P0 + c0n < p.total_pixels) {
          const float* rp = p.res_f32 + (P0 + c0n) * CO + c;
#pragma unroll
          for (int j = 0; j < 32; ++j) rnext[j] = __ldg(rp + j * CO);
        } else if (p.res_h16 && c_ok && P0 + c0n < p.total_pixels) {
          // 16-bit residual (mimicry's in-place ReLU makes the identity shortcut relu(h), which IS the stored conv operand):
          // half the bytes of the fp32 stream, and no fp32 copy of the block output has to be written at all
          const uint16_t* rp = p.res_h16 + (P0 + c0n) * CO + c;
#pragma unroll
          for (int j = 0; j < 32; ++j) rnext[j] = __uint_as_float((uint32_t)__ldg(rp + j * CO));      // raw bits, converted at use
        }
      };
      const int cbeg = half * 128, cend = cbeg + 128;
      load_res(cbeg);
#pragma unroll 1
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        const long long pb = P0 + c0;                     // first pixel of this chunk
        const bool chunk_ok = pb < p.total_pixels;        // warp-uniform; chunks never straddle total_pixels (multiple of 32)
        const bool st_ok = chunk_ok && c_ok;
        float rv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) rv[j] = rnext[j];
        if (c0 + 32 < cend) load_res(c0 + 32);
        if (c0 == cbeg) {
          mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
          tc_fence_after();
        }
        uint32_t r[32];
        tmem_ld32(tmem_acc + (uint32_t)c0, r);
        tmem_ld_wait();
        if (chunk_ok) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias;
          if (p.img) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float* px = s_px + ((c0 + j) * spx_n + spx_par) * 3;      // same address in every lane: broadcast
              v[j] = fmaf(w3[0], px[0], fmaf(w3[1], px[1], fmaf(w3[2], px[2], v[j])));
            }
          }
          if (p.act) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]) * 1.4142135623730951f;
          }
          if (p.res_f32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += p.res_relu ? fmaxf(rv[j], 0.f) : rv[j];
          } else if (p.res_h16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float r16 = unpack_h2<F16>(__float_as_uint(rv[j]) & 0xffffu).x;
              v[j] += p.res_relu ? fmaxf(r16, 0.f) : r16;
            }
          }
          if (p.out_scale != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
          }
          if (F16 && st_ok && (p.out_raw || p.out_relu)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, p.out_raw ? fabsf(v[j]) : v[j]);
          }
          if (p.out_f32 && st_ok) {
            float* op = p.out_f32 + pb * CO + c;
#pragma unroll
            for (int j = 0; j < 32; ++j) op[j * CO] = v[j];
          }
          if (p.out_raw && st_ok) {
            uint16_t* op = p.out_raw + pb * CO + c;
#pragma unroll
            for (int j = 0; j < 32; ++j) op[j * CO] = to_h16<F16>(v[j]);
          }
          if (p.out_relu && st_ok) {
            uint16_t* op = p.out_relu + pb * CO + c;
#pragma unroll
            for (int j = 0; j < 32; ++j) op[j * CO] = to_h16<F16>(fmaxf(v[j], 0.f));
          }
          if (p.head_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) head_sum += fmaxf(v[j], 0.f);
          }
        }
        if (p.head_out && ((c0 + 32) % HW) == 0) {
          // an image is complete: sum over this warp's 32 channels (fixed shuffle tree), park the partial for the tile's reduction
          float part = head_sum * hw_c;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          if (lane == 0) s_head[q * 8 + c0 / HW] = part;
          head_sum = 0.f;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
      if (p.head_out) {
        // deterministic cross-warp reduction: one thread per image adds the four warp partials in a fixed order
        named_bar_sync(2, SW_EPI_WARPS * 32);
        const int imgs = 256 / HW;
        if (et < imgs) {
          const long long img = P0 / HW + et;
          if (img < p.n_images)
            p.head_out[img] = p.head_b[0] + ((s_head[et] + s_head[8 + et]) + (s_head[16 + et] + s_head[24 + et]));
        }
        named_bar_sync(2, SW_EPI_WARPS * 32);
      }
    }
    if (F16 && vmax > kF16Max) range_flag_set(p.ovf, SDG_RANGE_ACT);
}

// ---------------------------------------------------------------------------------------------------
// The same epilogue for tiles of FOUR 8 x 8 images whose 256 TMEM columns are in (row, image, column) order -- the order in
// which conv_swap_shared_kernel stages such tiles so that a row shift of the operand is a whole number of swizzle atoms:
// column J <-> image 4t + ((J >> 3) & 3), pixel (J >> 5, J & 7).  Identity residual (fp32 or 16-bit), up to three outputs,
// fused SNGAN head; no image shortcut / super pixels (those layers do not have 8 x 8 outputs).
template <bool F16>
__device__ __forceinline__ void swap_epilogue_loop_perm(const TcParams& p, const uint32_t tmem_base, uint64_t* bar_acc_full,
                                                        uint64_t* bar_acc_empty, float* s_head, const long long tiles,
                                                        const int warp, const int lane) {
  const int q = warp & 3;
  const int half = (warp - 4) >> 2;
  const int c = q * 32 + lane;
  const int et = threadIdx.x - 128;
  constexpr int CO = 128;
  const float bias = p.bias ? p.bias[c] : 0.f;
  const float hw_c = p.head_out ? p.head_w[c] : 0.f;
  long long local = 0;
  float vmax = 0.f;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++local) {
    const int acc = (int)(local & 1);
    const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
    const long long P0 = t * 256;                         // first pixel of image 4t
    const long long img0 = t * 4;
    bool vi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) vi[i] = img0 + i < p.n_images;
    const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 256) + ((uint32_t)(q * 32) << 16);
    if (p.debug_skip_epi) {
      mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
      tc_fence_after();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
      continue;
    }
    float hsum[4] = {0.f, 0.f, 0.f, 0.f};
    float rnext[32];
    // element (pixel, channel c) of column j of the chunk that holds image row y: ((P0 + (j >> 3) * 64 + y * 8 + (j & 7)) * CO + c
    auto load_res = [&](int y) {
      const long long base = (P0 + y * 8) * CO + c;
      if (p.res_f32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (vi[j >> 3]) rnext[j] = __ldg(p.res_f32 + base + ((j >> 3) * 64 + (j & 7)) * CO);
      } else if (p.res_h16) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (vi[j >> 3]) rnext[j] = __uint_as_float((uint32_t)__ldg(p.res_h16 + base + ((j >> 3) * 64 + (j & 7)) * CO));
      }
    };
    const int ybeg = half * 4;
    load_res(ybeg);
#pragma unroll 1
    for (int y = ybeg; y < ybeg + 4; ++y) {
      float rv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) rv[j] = rnext[j];
      if (y + 1 < ybeg + 4) load_res(y + 1);
      if (y == ybeg) {
        mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
        tc_fence_after();
      }
      uint32_t r[32];
      tmem_ld32(tmem_acc + (uint32_t)(y * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias;
      if (p.act) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]) * 1.4142135623730951f;
      }
      if (p.res_f32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += p.res_relu ? fmaxf(rv[j], 0.f) : rv[j];
      } else if (p.res_h16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float r16 = unpack_h2<F16>(__float_as_uint(rv[j]) & 0xffffu).x;
          v[j] += p.res_relu ? fmaxf(r16, 0.f) : r16;
        }
      }
      if (p.out_scale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
      }
      const long long base = (P0 + y * 8) * CO + c;
      if (F16 && (p.out_raw || p.out_relu)) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (vi[j >> 3]) vmax = fmaxf(vmax, p.out_raw ? fabsf(v[j]) : v[j]);
      }
      if (p.out_f32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (vi[j >> 3]) p.out_f32[base + ((j >> 3) * 64 + (j & 7)) * CO] = v[j];
      }
      if (p.out_raw) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (vi[j >> 3]) p.out_raw[base + ((j >> 3) * 64 + (j & 7)) * CO] = to_h16<F16>(v[j]);
      }
      if (p.out_relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (vi[j >> 3]) p.out_relu[base + ((j >> 3) * 64 + (j & 7)) * CO] = to_h16<F16>(fmaxf(v[j], 0.f));
      }
      if (p.head_out) {
#pragma unroll
        for (int j = 0; j < 32; ++j) hsum[j >> 3] += fmaxf(v[j], 0.f);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
    if (p.head_out) {
      // every warp holds, for each of the four images, the sum over its 32 channels and its four image rows: fixed shuffle
      // tree per warp, then one thread per image adds the eight warp partials in a fixed order (deterministic logits)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float part = hsum[i] * hw_c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_head[(half * 4 + q) * 4 + i] = part;
      }
      named_bar_sync(2, SW_EPI_WARPS * 32);
      if (et < 4 && img0 + et < p.n_images) {
        const float* h = s_head + et;
        p.head_out[img0 + et] = p.head_b[0] + (((h[0] + h[4]) + (h[8] + h[12])) + ((h[16] + h[20]) + (h[24] + h[28])));
      }
      named_bar_sync(2, SW_EPI_WARPS * 32);
    }
  }
  if (F16 && vmax > kF16Max) range_flag_set(p.ovf, SDG_RANGE_ACT);
}

// ---------------------------------------------------------------------------------------------------
// Role-swapped kernel for Cout = 128 layers (every 3x3 conv of SNGAN-32, StyleGAN2's 256x256 conv1):
//     D^T[128 channels, 256 pixels] = W[128, K] * A[256 pixels, K]^T
// i.e. the WEIGHTS are the UMMA "A" operand (M = 128) and the pixel tile is the "B" operand with N = 256, so that one
// tcgen05.mma covers 128 x 256 x 16 -- twice the work per instruction of the pixel-major kernels, whose N is capped at
// Cout = 128.  Measured (profiles/r1e_ncu_full_pair_stream_summary.txt): N = 128 instructions keep the tensor pipe 55-58 %
// busy, N = 256 instructions 95 %; the instruction, not shared-memory bandwidth, is the unit that has to be large.
// The accumulator comes out transposed: TMEM lane = output channel, column = pixel.  That suits NHWC: for a fixed pixel the
// 32 lanes of an epilogue warp hold 32 consecutive channels, so every global access of the epilogue (16-bit / fp32 stores,
// fp32 residual loads) is a contiguous 64 / 128-byte segment with no shuffle transposes at all.
// Per stage: weights 16 KB + two 128-pixel boxes 32 KB; 4 stages; accumulator 2 x 256 TMEM columns.

template <bool F16>
__global__ void __launch_bounds__(SW_THREADS, 1)
conv_swap_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_s, const TcParams p, const int n_stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  __shared__ __align__(8) uint64_t bar_full[SW_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[SW_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_px[256 * 2 * 3];          // pooled normalised image pixels of the current tile (image shortcut; two per
                                               // GEMM pixel in the super-pixel form)
  __shared__ float s_head[4 * 8];              // per-warp partial logits of the (up to 8) images of the current tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.sc_chunks) tma_prefetch_desc(&map_s);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SW_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), SW_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const long long tiles = (p.m_tiles + 1) >> 1;          // 256-pixel tiles = pairs of 128-pixel boxes
  const int main_iters = p.taps * p.kchunks;
  const int k_iters = main_iters + p.sc_chunks;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        int n0[2], y0[2], x0[2];
        tc_tile_origin(p, 2 * t, n0[0], y0[0], x0[0]);
        tc_tile_origin(p, 2 * t + 1, n0[1], y0[1], x0[1]);      // beyond the last box: image index >= n, zero filled
        int tap = 0, kc = 0;
        for (int it = 0; it < k_iters; ++it) {
          bool is_sc;
          int dy, dx, ch;
          tc_k_iter(p, it, main_iters, tap, kc, is_sc, ch, dy, dx);
          const CUtensorMap* am = is_sc ? &map_s : &map_a;
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t dst = smem_base + stage * SW_STAGE;
          // timing experiment (SDG_TIMING_EXPERIMENTS builds only, WRONG results): debug_skip_a = m reloads the pixel operand
          // only on every m-th K iteration (stale shared memory otherwise): how much of the time is operand ingest?
          if (p.debug_skip_a > 1 && (it % p.debug_skip_a) != 0) {
            mbar_expect_tx(full, TC_A_BYTES);
            tma_load_2d(dst, &map_b, full, it * TC_BK, 0);
          } else {
          mbar_expect_tx(full, SW_STAGE);
          tma_load_2d(dst, &map_b, full, it * TC_BK, 0);
          tma_load_4d(dst + TC_A_BYTES, am, full, ch * TC_BK, p.csx * x0[0] + dx, p.cs * y0[0] + dy, n0[0]);
          tma_load_4d(dst + 2 * TC_A_BYTES, am, full, ch * TC_BK, p.csx * x0[1] + dx, p.cs * y0[1] + dy, n0[1]);
          }
          if (++kc == p.kchunks) { kc = 0; ++tap; }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: M = 128 channels, N = 256 pixels =================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 256, F16);
      int stage = 0;
      uint32_t phase = 0;
      long long local = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++local) {
        const int acc = (int)(local & 1);
        const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t w_addr = smem_base + stage * SW_STAGE;
          const uint64_t wdesc = make_sw128_desc(w_addr);
          const uint64_t pdesc = make_sw128_desc(w_addr + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(d_tmem, wdesc + (uint64_t)(2 * k), pdesc + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&bar_acc_full[acc]));
      }
    }
  } else if (warp >= 4) {
    swap_epilogue_loop<F16>(p, tmem_base, bar_acc_full, bar_acc_empty, s_px, s_head, tiles, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// Role-swapped kernel with TAP-SHARED operand staging, for layers whose 256-pixel tile is one whole 16 x 16 output grid
// (SNGAN-32 block1.c2 in the 4x4 stride-2 form, block2.c1).
// What bounds conv_swap_kernel is not the tensor pipe but operand INGEST: every K step (64 channels of one tap) pulls
// 16 KB of weights + 32 KB of pixels from L2 into shared memory for 4 x 128 cycles of MMA, 96 B per cycle against the
// ~64 B per cycle an SM takes in (ncu, profiles/r2b_*: l1tex__m_xbar2l1tex_read_bytes at its ceiling while the tensor pipe is
// 67-72 % busy; with the pixel loads skipped -- a timing experiment that produces wrong results -- the pass drops from
// 17.3 to 14.4 ms).  The pixel operand is the redundant part: the taps of a 3x3 (4x4) window re-read the same activations
// 9 (16) times.  Here a tile's activations are staged ONCE per (64-channel chunk, column variant) as a "copy" of 18 (17)
// image rows x 16 pixels, and the taps that differ by a row shift only move the UMMA descriptor's start address by whole
// rows (2 KB = two 1024-byte swizzle atoms, so TMA's 128B swizzle and the descriptor still agree).  A column shift would
// break the 8-row swizzle atom, hence one copy per column variant: 6 copies x 36 KB instead of 18 loads x 32 KB for a 3x3
// layer, 16 x 34 KB instead of 32 x 32 KB for the 4x4 stride-2 form (parity planes through the TMA traversal stride).
// Two producers (pixel copies: warp 0, weights: warp 3) feed independent rings; the epilogue is conv_swap_kernel's.
constexpr int SH_W_STAGES = 4;                // upper bound of the weight ring (sh.w_stages are used)
constexpr int SH_COPIES = 4;
constexpr int kSwapSharedSmem = 1024 + 208 * 1024;      // 4 x 16 KB weights + 4 x 36 KB copies, or 3 x 16 KB + 4 x 40 KB

template <bool F16>
__global__ void __launch_bounds__(SW_THREADS, 1)
conv_swap_shared_kernel(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_s,
                        const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t copy_base = smem_base + (uint32_t)p.sh.w_stages * TC_A_BYTES;
  const uint32_t SH_COPY_PITCH = (uint32_t)p.sh.pitch;
  const int w_stages = p.sh.w_stages;

  __shared__ __align__(8) uint64_t bar_wfull[SH_W_STAGES];
  __shared__ __align__(8) uint64_t bar_wempty[SH_W_STAGES];
  __shared__ __align__(8) uint64_t bar_pfull[SH_COPIES];
  __shared__ __align__(8) uint64_t bar_pempty[SH_COPIES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_px[256 * 2 * 3];
  __shared__ float s_head[4 * 8];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_c);
    tma_prefetch_desc(&map_s);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SH_W_STAGES; ++s) {
      mbar_init(smem_u32(&bar_wfull[s]), 1);
      mbar_init(smem_u32(&bar_wempty[s]), 1);
    }
    for (int s = 0; s < SH_COPIES; ++s) {
      mbar_init(smem_u32(&bar_pfull[s]), 1);
      mbar_init(smem_u32(&bar_pempty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), SW_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const ShSched& sh = p.sh;
  // natural order: one image (16 x 16 outputs) per tile; permuted order: four images (8 x 8 outputs each) per tile
  const long long tiles = sh.perm ? (p.n_images + 3) >> 2 : p.n_images;

  if (warp == 0) {
    // ================= pixel-copy producer =================
    if (elect_one()) {
      int pb = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int k = 0; k < sh.n_copies; ++k) {
          const ShCopy& cp = sh.cp[k];
          mbar_wait(smem_u32(&bar_pempty[pb]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_pfull[pb]);
          if (p.debug_skip_a & 4) { mbar_arrive(full); }        // timing experiment (wrong results): no pixel loads at all
          else {
            mbar_expect_tx(full, (uint32_t)((int)cp.rows * sh.row_bytes));
            const CUtensorMap* m = cp.src ? &map_s : &map_c;
            // permuted maps are {channel, column, image, row}: a copy lands as [row][image][8 pixels], images beyond n zero filled
            if (sh.perm) tma_load_4d(copy_base + pb * SH_COPY_PITCH, m, full, (int)cp.chunk * TC_BK, cp.col0, (int)(t * 4), cp.row0);
            else tma_load_4d(copy_base + pb * SH_COPY_PITCH, m, full, (int)cp.chunk * TC_BK, cp.col0, cp.row0, (int)t);
          }
          if (++pb == SH_COPIES) { pb = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ================= weight producer: one [128 x 64] K chunk per tap, in the order the copies are consumed =================
    if (elect_one()) {
      int ws = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int k = 0; k < sh.n_copies; ++k) {
          const ShCopy& cp = sh.cp[k];
          for (int j = 0; j < cp.ntaps; ++j) {
            mbar_wait(smem_u32(&bar_wempty[ws]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_wfull[ws]);
            if (p.debug_skip_a & 8) { mbar_arrive(full); }      // timing experiment (wrong results): no weight loads at all
            else {
              mbar_expect_tx(full, TC_A_BYTES);
              tma_load_2d(smem_base + ws * TC_A_BYTES, &map_b, full, (int)cp.wk[j] * TC_BK, 0);
            }
            if (++ws == w_stages) { ws = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: M = 128 channels, N = 256 pixels =================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 256, F16);
      int ws = 0, pb = 0;
      uint32_t wphase = 0, pphase = 0;
      long long local = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++local) {
        const int acc = (int)(local & 1);
        const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        uint32_t accumulate = 0;
        for (int k = 0; k < sh.n_copies; ++k) {
          const ShCopy& cp = sh.cp[k];
          mbar_wait(smem_u32(&bar_pfull[pb]), pphase);
          const uint32_t c_addr = copy_base + pb * SH_COPY_PITCH;
          for (int j = 0; j < cp.ntaps; ++j) {
            mbar_wait(smem_u32(&bar_wfull[ws]), wphase);
            tc_fence_after();
            const uint64_t wdesc = make_sw128_desc(smem_base + ws * TC_A_BYTES);
            const uint64_t pdesc = make_sw128_desc(c_addr + (uint32_t)((int)cp.roff[j] * sh.row_bytes));
#pragma unroll
            for (int kk = 0; kk < TC_BK / 16; ++kk) {
              umma_bf16(d_tmem, wdesc + (uint64_t)(2 * kk), pdesc + (uint64_t)(2 * kk), idesc, accumulate);
              accumulate = 1u;
            }
            umma_commit(smem_u32(&bar_wempty[ws]));
            if (++ws == w_stages) { ws = 0; wphase ^= 1u; }
          }
          umma_commit(smem_u32(&bar_pempty[pb]));          // every tap of this copy has been issued: free it when they retire
          if (++pb == SH_COPIES) { pb = 0; pphase ^= 1u; }
        }
        umma_commit(smem_u32(&bar_acc_full[acc]));
      }
    }
  } else if (warp >= 4) {
    if (sh.perm) swap_epilogue_loop_perm<F16>(p, tmem_base, bar_acc_full, bar_acc_empty, s_head, tiles, warp, lane);
    else swap_epilogue_loop<F16>(p, tmem_base, bar_acc_full, bar_acc_empty, s_px, s_head, tiles, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = kNumSMs;

template <int BN>
constexpr int tc_smem_bytes() { return TC_STAGES * (TC_A_BYTES + BN * TC_BK * 2) + 1024; }

int tc_num_sms() { return g_num_sms; }

int tc_max_active_clusters(const void* kernel, const cudaLaunchConfig_t* cfg) {
  struct Entry { const void* k; int dev; size_t smem; int n; };
  static Entry cache[32];
  static std::atomic<int> n_cached{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const int nc = n_cached.load(std::memory_order_acquire);
  for (int i = 0; i < nc; ++i)
    if (cache[i].k == kernel && cache[i].dev == dev && cache[i].smem == cfg->dynamicSmemBytes) return cache[i].n;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  static const int verbose = getenv("SDG_VERBOSE") ? atoi(getenv("SDG_VERBOSE")) : 0;
  if (verbose) fprintf(stderr, "sdg: %d clusters of %u CTAs x %u threads, %zu B dynamic smem resident at once\n", n,
                       cfg->numAttrs ? cfg->attrs[0].val.clusterDim.x : 1, cfg->blockDim.x, cfg->dynamicSmemBytes);
  const int slot = n_cached.load(std::memory_order_relaxed);
  if (slot < 32) { cache[slot] = {kernel, dev, cfg->dynamicSmemBytes, n}; n_cached.store(slot + 1, std::memory_order_release); }
  return n;
}

constexpr int kPairSmemMax = 227 * 1024 - 3072;       // dynamic shared memory budget of the pair kernel (static: ~2.5 KB)
constexpr int kStreamSmem = 1024 + 6 * (TC_A_BYTES + 128 * TC_BK * 2);   // 6 x 32 KB (BN 256) = 8 x 24 KB (BN 128) stages
static int g_pair_mode = 1;                          // 0 = single-CTA pixel-major kernels only, 1 = default selection (role-swapped,
                                                     // streamed / resident CTA pairs where they apply), 2 = CTA pairs but no role swap
void conv_tc_set_pair(int on) { g_pair_mode = on; }
// true when Cout = 128 stages with a linear epilogue run on the role-swapped kernel (the only epilogue that knows the
// super-pixel form of the image shortcut)
bool conv_tc_swap_active() {
  static const int swap_mode = getenv("SDG_SWAP") ? atoi(getenv("SDG_SWAP")) : 1;
  return swap_mode && g_pair_mode == 1;
}

int tc_encode_2d(CUtensorMap* map, const void* ptr, int f16, uint64_t inner, uint64_t outer, uint32_t box_inner,
                 uint32_t box_outer) {
  SDG_REQUIRE(g_encode, SDG_E_STATE, "conv_tc: conv_tc_init not called");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDG_REQUIRE(r == CUDA_SUCCESS, SDG_E_INVALID, "conv_tc: cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
  return 0;
}

// 4-D map over NHWC activations; box = 64 channels x bw x bh x bn pixels; es = traversal stride over pixels (2 for the
// stride-2 form: the box then spans es*bw x es*bh input pixels and every es-th one is loaded)
static int encode_act(CUtensorMap* map, const void* ptr, int f16, int64_t n, int H, int W, int C, int bw, int bh, int bn,
                      int es, int esx = 0) {
  if (esx <= 0) esx = es;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(bw * esx), (cuuint32_t)(bh * es), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)esx, (cuuint32_t)es, 1};
  CUresult r = g_encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDG_REQUIRE(r == CUDA_SUCCESS, SDG_E_INVALID, "conv_tc: cuTensorMapEncodeTiled(act) failed: %d", (int)r);
  return 0;
}

// plain (un-swizzled) box of bc channels x bw columns x bh rows of one image over NHWC activations: the streaming kernels
// (blur_tma.cu) read such a box with ordinary shared-memory loads; out-of-bounds elements arrive as zeros
int tc_encode_act_box(CUtensorMap* map, const void* ptr, int f16, int64_t n, int H, int W, int C, int bc, int bw, int bh) {
  SDG_REQUIRE(g_encode, SDG_E_STATE, "conv_tc: conv_tc_init not called");
  SDG_REQUIRE(bc <= 256 && bw <= 256 && bh <= 256 && (bc * 2) % 16 == 0, SDG_E_INVALID, "tc_encode_act_box: box %d x %d x %d", bc, bw, bh);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDG_REQUIRE(r == CUDA_SUCCESS, SDG_E_INVALID, "conv_tc: cuTensorMapEncodeTiled(act box) failed: %d", (int)r);
  return 0;
}

// Function attributes (opt-in shared memory) are per DEVICE, the driver entry point is per process: one bit per device id
// records which devices have been initialised, so that a single process may drive several GPUs.
static std::atomic<unsigned long long> g_dev_init{0};

// the same tensor viewed as {channel, column, IMAGE, row} (strides are free in a tiled map): a box {64, bw, bn, bh} then lands in
// shared memory as [row][image][column], the layout conv_swap_shared_kernel needs for tiles of several small images
static int encode_act_perm(CUtensorMap* map, const void* ptr, int f16, int64_t n, int H, int W, int C, int bw, int bh, int bn,
                           int es) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)n, (cuuint64_t)H};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(bw * es), (cuuint32_t)bn, (cuuint32_t)(bh * es)};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, 1, (cuuint32_t)es};
  CUresult r = g_encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDG_REQUIRE(r == CUDA_SUCCESS, SDG_E_INVALID, "conv_tc: cuTensorMapEncodeTiled(act, permuted) failed: %d", (int)r);
  return 0;
}

int conv_tc_init(int device) {
  if (device >= 0 && device < 64 && ((g_dev_init.load(std::memory_order_acquire) >> device) & 1ULL)) return 0;
  cudaDeviceProp prop;
  SDG_CUDA(cudaGetDeviceProperties(&prop, device));
  SDG_REQUIRE(prop.major == 10, SDG_E_DEVICE, "conv_tc: device %d is sm_%d%d, need sm_100 (B200)", device, prop.major,
              prop.minor);
  g_num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SDG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  SDG_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, SDG_E_DEVICE, "conv_tc: cuTensorMapEncodeTiled unavailable");
  SDG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<128>()));
  SDG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<64>()));
  SDG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<128>()));
  SDG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<64>()));
  SDG_CUDA(cudaFuncSetAttribute(conv_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemMax));
  SDG_CUDA(cudaFuncSetAttribute(conv_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemMax));
  SDG_CUDA(cudaFuncSetAttribute(conv_swap_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwapSmem));
  SDG_CUDA(cudaFuncSetAttribute(conv_swap_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwapSmem));
  SDG_CUDA(cudaFuncSetAttribute(conv_swap_shared_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwapSharedSmem));
  SDG_CUDA(cudaFuncSetAttribute(conv_swap_shared_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwapSharedSmem));
  SDG_CUDA(cudaFuncSetAttribute((conv_pair_stream_kernel<256, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
  SDG_CUDA(cudaFuncSetAttribute((conv_pair_stream_kernel<256, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
  SDG_CUDA(cudaFuncSetAttribute((conv_pair_stream_kernel<128, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
  SDG_CUDA(cudaFuncSetAttribute((conv_pair_stream_kernel<128, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
  g_encode = (EncodeTiledFn)fn;
  int rc = first_conv_init();
  if (rc) return rc;
  rc = b1_fused_init();
  if (rc) return rc;
  if (device >= 0 && device < 64) g_dev_init.fetch_or(1ULL << device, std::memory_order_release);
  return 0;
}

int conv_tc(const TcConv& a, int f16, cudaStream_t s) {
  SDG_REQUIRE(g_encode, SDG_E_STATE, "conv_tc: conv_tc_init not called");
  const int H = a.H, W = a.W, Cin = a.Cin, Cout = a.Cout, taps = a.taps;
  const bool s2 = a.pool4 != 0;          // conv3x3 + avg_pool2d(2) evaluated as the equivalent 4x4 stride-2 conv
  const bool strided = s2 || a.stride == 2;
  SDG_REQUIRE(a.general || taps == 9 || taps == 1, SDG_E_UNSUPPORTED, "conv_tc: taps=%d", taps);
  SDG_REQUIRE(!a.general || (a.taps_x >= 1 && a.taps_y >= 1 && a.grid_w >= 4 && (a.grid_w & (a.grid_w - 1)) == 0 && a.in_H > 0 &&
                             a.in_W > 0 && !a.pool && !a.pool4 && a.stride == 1 && !a.sc_in && !a.gemm && a.sx * a.grid_w <= 256),
              SDG_E_INVALID, "conv_tc: malformed general tap grid");
  SDG_REQUIRE(!s2 || (taps == 9 && a.pool), SDG_E_INVALID, "conv_tc: pool4 needs a pooled 3x3 stage");
  SDG_REQUIRE(a.stride == 1 || a.stride == 2, SDG_E_UNSUPPORTED, "conv_tc: stride=%d", a.stride);
  SDG_REQUIRE(!(a.stride == 2 && a.pool), SDG_E_UNSUPPORTED, "conv_tc: strided conv with pooling");
  SDG_REQUIRE(Cin % TC_BK == 0 && Cout % 64 == 0 && Cout <= TC_MAX_COUT, SDG_E_UNSUPPORTED, "conv_tc: Cin=%d Cout=%d", Cin, Cout);
  if (a.gemm) {
    SDG_REQUIRE(taps == 1 && H == 1 && W >= 1 && a.n == 1 && !a.pool && a.stride == 1 && !a.sc_in && !a.img && !a.sd, SDG_E_INVALID,
                "conv_tc: gemm mode takes a [W rows][Cin] matrix (n = 1, H = 1, taps = 1) and no conv extras");
  } else if (a.general) {
    SDG_REQUIRE(H >= 4 && (H & (H - 1)) == 0, SDG_E_UNSUPPORTED, "conv_tc: general grid H=%d", H);
  } else {
    SDG_REQUIRE(W >= 4 && W <= 512 && (W & (W - 1)) == 0 && H == W, SDG_E_UNSUPPORTED, "conv_tc: H=%d W=%d", H, W);
  }
  SDG_REQUIRE(!a.sc_sep || (a.sc_in && !a.pool && !a.pool4 && Cout % 128 == 0), SDG_E_INVALID,
              "conv_tc: separate-accumulator shortcut needs a shortcut tensor, no pooling and Cout %% 128 == 0");
  SDG_REQUIRE((a.sd == nullptr) == (a.sd_w == nullptr) && (!a.sd || !a.pool), SDG_E_INVALID, "conv_tc: sd / sd_w mismatch");
  SDG_REQUIRE(a.sc_C % TC_BK == 0, SDG_E_UNSUPPORTED, "conv_tc: shortcut channels %d", a.sc_C);
  SDG_REQUIRE((a.sc_C == 0) == (a.sc_in == nullptr), SDG_E_INVALID, "conv_tc: shortcut tensor / channels mismatch");
  SDG_REQUIRE(!a.img || ((a.pool || a.general) && a.sc_w3), SDG_E_INVALID, "conv_tc: image shortcut needs pooling and weights");
  SDG_REQUIRE(a.out_relu || a.out_raw || a.out_f32 || a.head_out, SDG_E_INVALID, "conv_tc: no output");
  SDG_REQUIRE(!a.head_out || (a.head_w && a.head_b && Cout == 128 && !a.pool && !a.img && (H * W == 32 || H * W == 64) && !a.gemm),
              SDG_E_UNSUPPORTED, "conv_tc: fused head needs Cout = 128, an un-pooled stage and 32 or 64 pixels per image");
  auto al16 = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
  SDG_REQUIRE(al16(a.in) && al16(a.wb) && al16(a.sc_in) && al16(a.res_f32) && al16(a.res_h16) && al16(a.out_relu) &&
                  al16(a.out_raw) && al16(a.out_f32), SDG_E_INVALID, "conv_tc: pointers must be 16-byte aligned");
  SDG_REQUIRE(!(a.res_f32 && a.res_h16), SDG_E_INVALID, "conv_tc: one residual tensor at most");
  if (a.n == 0) return 0;
  // H, W describe the GEMM's M grid for a stride-1 conv and for pool4 (where the grid is H/2 x W/2); for an explicit
  // stride-2 conv they are the OUTPUT grid and in_H / in_W give the input tensor's extent
  const int Hc = s2 ? H / 2 : H, Wc = a.general ? a.grid_w : (s2 ? W / 2 : W);
  const int Hin = a.in_H ? a.in_H : H, Win = a.in_W ? a.in_W : W;
  TcParams p;
  p.H = Hc; p.W = Wc; p.Cin = Cin; p.Cout = Cout;
  p.taps = s2 ? 16 : taps;
  p.tw = s2 ? 4 : (taps == 9 ? 3 : 1);
  p.cs = strided ? 2 : 1;
  p.toff = a.no_pad ? 0 : -1;
  p.csx = p.cs; p.toffx = p.toff; p.superpix = 0;
  if (a.general) {
    p.taps = a.taps_x * a.taps_y; p.tw = a.taps_x;
    p.cs = a.sy; p.csx = a.sx; p.toff = a.offy; p.toffx = a.offx; p.superpix = a.superpix;
  }
  p.act = a.act;
  p.out_scale = a.out_scale;
  p.img_up = (s2 || (a.general && a.img_up)) ? 1 : 0;
  p.kchunks = Cin / TC_BK;
  p.sc_kchunks = a.sc_C / TC_BK;
  p.sc_chunks = (s2 ? 4 : 1) * p.sc_kchunks;
  p.pool = (a.pool && !s2) ? 1 : 0;
  p.box16 = (p.pool && W >= 32) ? 1 : 0;
  p.wide = 0;
  p.sc_sep = a.sc_sep;
  p.tiles_x = 1;
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  p.n_tiles = Cout / BN;
  int bw = Wc, bh, bn;
  if (p.box16) {
    SDG_REQUIRE(Wc <= 128, SDG_E_UNSUPPORTED, "conv_tc: pooled epilogue with W=%d", Wc);
    bw = 16; bh = 8; bn = 1;
    p.bh = 8; p.bn = 1;
    p.tiles_x = Wc / 16;
    p.tiles_y = Hc / 8;
    p.m_tiles = a.n * p.tiles_x * p.tiles_y;
  } else if (Wc > TC_BM || a.gemm) {
    p.wide = 1;
    bw = TC_BM; bh = 1; bn = 1;
    p.bh = 1; p.bn = 1;
    p.tiles_x = (int)cdiv(Wc, TC_BM);
    p.tiles_y = Hc;
    p.m_tiles = a.n * (long long)Hc * p.tiles_x;
  } else {
    int rows = TC_BM / Wc;                      // grid rows per tile if one image is big enough
    if (rows >= Hc) { p.bh = Hc; p.bn = TC_BM / (Hc * Wc); } else { p.bh = rows; p.bn = 1; }
    bh = p.bh; bn = p.bn;
    p.tiles_y = Hc / p.bh;
    p.m_tiles = cdiv(a.n, p.bn) * p.tiles_y;
  }
  p.res_relu = a.res_relu; p.img_layout = a.img_layout;
  // Timing experiments that produce WRONG results (operand loads or epilogues skipped) exist only in a library built with
  // `make EXTRA=-DSDG_TIMING_EXPERIMENTS`; the shipped build ignores these environment variables.
#ifdef SDG_TIMING_EXPERIMENTS
  static const int dbg_skip_a = getenv("SDG_DEBUG_SKIP_A") ? atoi(getenv("SDG_DEBUG_SKIP_A")) : 0;
  static const int dbg_skip_epi = getenv("SDG_DEBUG_SKIP_EPI") ? atoi(getenv("SDG_DEBUG_SKIP_EPI")) : 0;
#else
  static const int dbg_skip_a = 0, dbg_skip_epi = 0;
#endif
  static const int dbg_stages = getenv("SDG_PAIR_STAGES") ? atoi(getenv("SDG_PAIR_STAGES")) : 0;
  p.debug_skip_a = dbg_skip_a;
  p.debug_skip_epi = dbg_skip_epi;
  p.n_images = a.n;
  p.total_pixels = a.n * Hc * Wc;
  p.bias = a.bias; p.sd = a.sd; p.sd_w = a.sd_w; p.res_f32 = a.res_f32; p.res_h16 = a.res_h16;
  p.head_w = a.head_w; p.head_b = a.head_b; p.head_out = a.head_out; p.img = a.img; p.sc_w3 = a.sc_w3;
  p.out_relu = a.out_relu; p.out_raw = a.out_raw; p.out_f32 = a.out_f32;
  p.ovf = t_range_flag;
  const int es = a.general ? a.sy : (strided ? 2 : 1);      // TMA traversal stride over input pixels (rows; columns: esx)
  const int esx = a.general ? a.sx : es;

  CUtensorMap map_a, map_b, map_s;
  const uint64_t k_cols = (uint64_t)p.taps * Cin + (uint64_t)p.sc_chunks * TC_BK;
  { int rc = encode_act(&map_a, a.in, f16, a.n, Hin, Win, Cin, bw, bh, bn, es, esx); if (rc) return rc; }
  if (a.sc_in && a.sc_sep) { int rc = encode_act(&map_s, a.sc_in, f16, a.n, Hc, Wc, a.sc_C, bw, bh, bn, 1); if (rc) return rc; }
  else if (a.sc_in) { int rc = encode_act(&map_s, a.sc_in, f16, a.n, Hin, Win, a.sc_C, bw, bh, bn, es); if (rc) return rc; }
  else map_s = map_a;
  { int rc = tc_encode_2d(&map_b, a.wb, f16, k_cols, Cout, TC_BK, BN); if (rc) return rc; }

  const int k_iters = p.taps * p.kchunks + p.sc_chunks;
  // role-swapped kernel (N = 256 pixels per MMA) for Cout = 128 layers with a "linear" epilogue
  static const int swap_mode = getenv("SDG_SWAP") ? atoi(getenv("SDG_SWAP")) : 1;     // SDG_SWAP=0: A/B against the pixel-major kernels
  static const int swap64 = getenv("SDG_SWAP64") ? atoi(getenv("SDG_SWAP64")) : 0;
  // (a single 128-pixel box is allowed when the 16-bit residual is used, which only this kernel implements: the second box
  // of the tile then lies beyond the tensor and is zero filled)
  // tap-shared operand staging (conv_swap_shared_kernel) applies when a tile is one whole 16 x 16 output grid, or four 8 x 8
  // grids in (row, image, column) order; 3x3 stride 1 or the 4x4 stride-2 form, optionally with the folded 1x1 shortcut.
  // It is chosen by SHAPE only (any n >= 1), so that a sample's logit never depends on the batch it is evaluated in.
  static const int shared_taps = getenv("SDG_SHARED_TAPS") ? atoi(getenv("SDG_SHARED_TAPS")) : 1;   // 2: 16 x 16 grids only
  const bool grid16 = Hc == 16 && Wc == 16, grid8 = Hc == 8 && Wc == 8 && shared_taps == 1;
  const bool sh_ok = shared_taps && Cout == 128 && (grid16 || grid8) && !a.general && (p.sc_chunks == 0 || s2) &&
                     (s2 || (taps == 9 && !strided)) && p.toff == -1 && (grid16 || !a.img);
  if (swap_mode && g_pair_mode == 1 && (Cout == 128 || (swap64 && Cout == 64)) && !p.pool && !p.box16 && !p.sc_sep && !a.sd &&
      !a.gemm && (p.m_tiles >= 2 || a.res_h16 || sh_ok) && p.total_pixels % 32 == 0 &&
      (!a.head_out || (256 % (Hc * Wc) == 0 && Hc * Wc >= 32))) {
    // Cout = 64 (SDG_SWAP64=1, experiment): the weight box still has 128 rows; rows 64..127 lie outside the tensor and are
    // zero filled by TMA, so the M = 128 instruction runs half empty and N stays 256
    CUtensorMap map_w;
    { int rc = tc_encode_2d(&map_w, a.wb, f16, k_cols, Cout, TC_BK, 128); if (rc) return rc; }
    if (sh_ok) {
      ShSched& sh = p.sh;
      memset(&sh, 0, sizeof(sh));
      sh.perm = grid8 ? 4 : 0;
      sh.row_bytes = (grid8 ? 4 * 8 : 16) * 128;
      int nc = 0;
      for (int c = 0; c < p.kchunks; ++c) {
        if (s2) {                                 // 4x4 stride-2 form: copy = (row class, column tap b); taps a = class, class + 2
          for (int rcl = 0; rcl < 2; ++rcl)
            for (int b = 0; b < 4; ++b) {
              ShCopy& cp = sh.cp[nc++];
              cp.src = 0; cp.chunk = (signed char)c; cp.col0 = (signed char)(b - 1);
              cp.row0 = (signed char)(rcl == 0 ? -1 : 0);           // input rows -1, 1, 3, ... (a = 0, 2) or 0, 2, 4, ... (a = 1, 3)
              cp.rows = (signed char)(Hc + 1); cp.ntaps = 2;
              const int a0 = rcl == 0 ? 0 : 1;
              cp.wk[0] = (short)((a0 * 4 + b) * p.kchunks + c); cp.roff[0] = 0;
              cp.wk[1] = (short)(((a0 + 2) * 4 + b) * p.kchunks + c); cp.roff[1] = 1;
            }
        } else {                                  // 3x3 stride 1: copy = column tap, taps dy = -1, 0, 1
          for (int kx = 0; kx < 3; ++kx) {
            ShCopy& cp = sh.cp[nc++];
            cp.src = 0; cp.chunk = (signed char)c; cp.col0 = (signed char)(kx - 1); cp.row0 = -1;
            cp.rows = (signed char)(Hc + 2); cp.ntaps = 3;
            for (int j = 0; j < 3; ++j) { cp.wk[j] = (short)((j * 3 + kx) * p.kchunks + c); cp.roff[j] = (signed char)j; }
          }
        }
      }
      // folded 1x1 shortcut of a pooled block: the four taps of the 2x2 stride-2 average, each its own parity plane
      const int main_chunks = p.taps * p.kchunks;
      for (int st = 0; st < (p.sc_chunks ? 4 : 0); ++st)
        for (int ch = 0; ch < p.sc_kchunks; ++ch) {
          ShCopy& cp = sh.cp[nc++];
          cp.src = 1; cp.chunk = (signed char)ch; cp.col0 = (signed char)(st & 1); cp.row0 = (signed char)(st >> 1);
          cp.rows = (signed char)Hc; cp.ntaps = 1;
          cp.wk[0] = (short)(main_chunks + st * p.sc_kchunks + ch); cp.roff[0] = 0;
        }
      sh.n_copies = nc;
      SDG_REQUIRE(nc <= 24, SDG_E_UNSUPPORTED, "conv_tc: tap-shared schedule with %d copies", nc);
      const int rows_main = s2 ? Hc + 1 : Hc + 2;
      sh.pitch = (rows_main * sh.row_bytes + 1023) / 1024 * 1024;
      sh.w_stages = (SH_W_STAGES * TC_A_BYTES + SH_COPIES * sh.pitch <= kSwapSharedSmem - 1024) ? SH_W_STAGES : SH_W_STAGES - 1;
      SDG_REQUIRE(sh.w_stages * TC_A_BYTES + SH_COPIES * sh.pitch <= kSwapSharedSmem - 1024, SDG_E_UNSUPPORTED,
                  "conv_tc: tap-shared copies of %d bytes do not fit", sh.pitch);
      CUtensorMap map_c, map_s2;
      if (grid8) {
        { int rc = encode_act_perm(&map_c, a.in, f16, a.n, Hin, Win, Cin, 8, rows_main, 4, es); if (rc) return rc; }
        if (a.sc_in) { int rc = encode_act_perm(&map_s2, a.sc_in, f16, a.n, Hin, Win, a.sc_C, 8, Hc, 4, es); if (rc) return rc; }
        else map_s2 = map_c;
      } else {
        { int rc = encode_act(&map_c, a.in, f16, a.n, Hin, Win, Cin, Wc, rows_main, 1, es, esx); if (rc) return rc; }
        // folded shortcut of a pooled block whose output is one 16 x 16 grid (SNGAN-64 block2.c2): Hc rows per parity plane
        if (a.sc_in && p.sc_chunks) { int rc = encode_act(&map_s2, a.sc_in, f16, a.n, Hin, Win, a.sc_C, Wc, Hc, 1, es, esx); if (rc) return rc; }
        else map_s2 = map_c;
      }
      const long long tiles_sh = grid8 ? (a.n + 3) / 4 : a.n;
      const int grid = (int)(tiles_sh < g_num_sms ? tiles_sh : g_num_sms);
      if (f16) { SDG_LAUNCH(conv_swap_shared_kernel<true>, grid, SW_THREADS, kSwapSharedSmem, s, map_c, map_s2, map_w, p); }
      else { SDG_LAUNCH(conv_swap_shared_kernel<false>, grid, SW_THREADS, kSwapSharedSmem, s, map_c, map_s2, map_w, p); }
      return 0;
    }
    const long long tiles = (p.m_tiles + 1) / 2;
    const int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);
    static const int sw_stages_env = getenv("SDG_SWAP_STAGES") ? atoi(getenv("SDG_SWAP_STAGES")) : SW_STAGES;   // timing experiment
    const int sw_stages = sw_stages_env >= 2 && sw_stages_env <= SW_STAGES ? sw_stages_env : SW_STAGES;
    if (f16) { SDG_LAUNCH(conv_swap_kernel<true>, grid, SW_THREADS, kSwapSmem, s, map_a, map_w, map_s, p, sw_stages); }
    else { SDG_LAUNCH(conv_swap_kernel<false>, grid, SW_THREADS, kSwapSmem, s, map_a, map_w, map_s, p, sw_stages); }
    return 0;
  }
  SDG_REQUIRE(!a.res_h16, SDG_E_UNSUPPORTED, "conv_tc: the 16-bit residual needs the role-swapped kernel (Cout = 128, linear "
              "epilogue, pixel count a multiple of 32)");
  static const int stream_all = getenv("SDG_PAIR_STREAM128") ? atoi(getenv("SDG_PAIR_STREAM128")) == 2 : 0;
  if (g_pair_mode && Cout == 128 && taps == 9 && k_iters <= PAIR_MAX_KI && p.m_tiles >= 2 && !p.sc_sep && !stream_all) {
    // CTA-pair kernel: weights resident (64 rows per CTA), A streamed through as many 16 KB stages as fit
    CUtensorMap map_bh;
    { int rc = tc_encode_2d(&map_bh, a.wb, f16, k_cols, Cout, TC_BK, 64); if (rc) return rc; }
    int n_stages = (kPairSmemMax - 1024 - k_iters * PAIR_B_TILE) / TC_A_BYTES;
    if (n_stages > 8) n_stages = 8;
    if (dbg_stages > 0 && dbg_stages < n_stages) n_stages = dbg_stages;
    SDG_REQUIRE(n_stages >= 3, SDG_E_UNSUPPORTED, "conv_tc: pair kernel needs >= 3 stages, K chunks = %d", k_iters);
    const size_t smem = 1024 + (size_t)k_iters * PAIR_B_TILE + (size_t)n_stages * TC_A_BYTES;
    const long long pair_tiles = (p.m_tiles + 1) / 2;
    long long clusters = g_num_sms / 2;
    if (pair_tiles < clusters) clusters = pair_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    {
      cfg.gridDim = dim3((unsigned)(g_num_sms / 2 * 2));
      const long long fit = tc_max_active_clusters(f16 ? (const void*)conv_pair_kernel<true> : (const void*)conv_pair_kernel<false>, &cfg);
      if (fit >= 1 && fit < clusters) clusters = fit;
      cfg.gridDim = dim3((unsigned)(2 * clusters));
    }
    if (f16) { SDG_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<true>, map_a, map_bh, map_s, p, n_stages)); }
    else { SDG_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<false>, map_a, map_bh, map_s, p, n_stages)); }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  // Cout = 128 layers whose weights are too large to stay resident (SNGAN-32 block1.c2 / block2.c2 in the 4x4 stride-2 form)
  // also run as CTA pairs, N = 128, weights streamed (measured +1.9 % on the whole SNGAN-32 pass); 0 disables, 2 = experiment:
  // streamed weights for EVERY Cout = 128 3x3 layer instead of the resident-weight kernel
  static const int stream128 = getenv("SDG_PAIR_STREAM128") ? atoi(getenv("SDG_PAIR_STREAM128")) : 1;
  // the skip-accumulator form is single-buffered at N = 256 (TMEM is full): worth it only when the mainloop is long
  const bool sep_short = p.sc_sep && k_iters < 40;
  const int sbn = (Cout % 256 == 0 && !sep_short) ? 256 : ((p.sc_sep || (stream128 && Cout == 128)) && Cout % 128 == 0 ? 128 : 0);
  if (g_pair_mode && sbn && p.m_tiles >= 2 && !p.pool && (!p.img || Cout == 128)) {
    // CTA-pair kernel with streamed weights: M = 256 x N = sbn per MMA
    p.n_tiles = Cout / sbn;
    CUtensorMap map_bh;
    { int rc = tc_encode_2d(&map_bh, a.wb, f16, k_cols, Cout, TC_BK, sbn / 2); if (rc) return rc; }
    const int n_stages = sbn == 256 ? 6 : 8;
    const long long total = ((p.m_tiles + 1) / 2) * p.n_tiles;
    long long clusters = g_num_sms / 2;
    if (total < clusters) clusters = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = kStreamSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    {
      // static tile schedule: all clusters must be resident at once (a GPC with an odd number of free SMs strands one)
      const void* kfn = sbn == 256 ? (f16 ? (const void*)conv_pair_stream_kernel<256, true> : (const void*)conv_pair_stream_kernel<256, false>)
                                   : (f16 ? (const void*)conv_pair_stream_kernel<128, true> : (const void*)conv_pair_stream_kernel<128, false>);
      cfg.gridDim = dim3((unsigned)(g_num_sms / 2 * 2));
      const long long fit = tc_max_active_clusters(kfn, &cfg);
      if (fit >= 1 && fit < clusters) clusters = fit;
      cfg.gridDim = dim3((unsigned)(2 * clusters));
    }
    if (sbn == 256 && f16) { SDG_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_stream_kernel<256, true>, map_a, map_bh, map_s, p, n_stages)); }
    else if (sbn == 256) { SDG_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_stream_kernel<256, false>, map_a, map_bh, map_s, p, n_stages)); }
    else if (f16) { SDG_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_stream_kernel<128, true>, map_a, map_bh, map_s, p, n_stages)); }
    else { SDG_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_stream_kernel<128, false>, map_a, map_bh, map_s, p, n_stages)); }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  const long long total_tiles = p.m_tiles * p.n_tiles;
  const int grid = (int)(total_tiles < g_num_sms ? total_tiles : g_num_sms);
  if (BN == 128 && f16) {
    SDG_LAUNCH((conv_tc_kernel<128, true>), grid, TC_THREADS, tc_smem_bytes<128>(), s, map_a, map_b, map_s, p);
  } else if (BN == 128) {
    SDG_LAUNCH((conv_tc_kernel<128, false>), grid, TC_THREADS, tc_smem_bytes<128>(), s, map_a, map_b, map_s, p);
  } else if (f16) {
    SDG_LAUNCH((conv_tc_kernel<64, true>), grid, TC_THREADS, tc_smem_bytes<64>(), s, map_a, map_b, map_s, p);
  } else {
    SDG_LAUNCH((conv_tc_kernel<64, false>), grid, TC_THREADS, tc_smem_bytes<64>(), s, map_a, map_b, map_s, p);
  }
  return 0;
}

}  // namespace sdg
