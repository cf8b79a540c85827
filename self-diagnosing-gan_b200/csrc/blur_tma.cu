// blur_tma.cu -- StyleGAN2's Blur (upfirdn2d with outer([1,3,3,1])/64 and zero padding; diagan/models/stylegan2.py:75-90,
// op/upfirdn2d.py) on 16-bit NHWC activations as a TMA-fed streaming kernel.
//
// blur_h16_kernel (sg2_fp32.cu) reads its four column taps through L1 with four loads in flight per thread and runs at
// 0.55-0.73 of the HBM rate (ncu: issue slots ~50 % busy, 35 % of the warps resident: latency AND issue bound).  Here
//   * a CTA owns (image, 64-channel chunk, strip of output columns, segment of output rows) and marches down the rows;
//   * one thread feeds a ring of shared-memory stages with cp.async.bulk.tensor boxes of 4 input rows x (strip + halo) columns
//     x 64 channels.  The zero padding of the blur IS the out-of-bounds zero fill of the tensor map, so no tap is ever
//     predicated and no address is ever clamped; 3-4 stages per CTA keep 50-100 KB per SM in flight without holding registers;
//   * a thread owns PPT adjacent output columns x 8 channels: it filters each input row horizontally from shared memory
//     (neighbouring outputs share the converted pixels) and keeps the last four filtered rows in registers for the vertical pass;
//   * all fp32 arithmetic is issued as packed pairs (fma.rn.f32x2 / add.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FADD2 / FMUL2 on
//     sm_100): half the issue slots of the scalar form, bit-identical results.
// The operation order per output equals blur_h16_kernel's, so the two kernels agree bit for bit (tested).
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace sdg {

namespace {

constexpr int BT_ROWS = 4;            // input rows per stage
constexpr int BT_CH = 64;             // channels per CTA (one 128-byte row of the box)
constexpr int BT_THREADS = 256;       // 8 channel groups x 32 column slots

template <int ST, int PPT>
struct BtGeom {
  static constexpr int XW = 32 * PPT;                        // output columns per CTA
  static constexpr int WIN = ST * (XW - 1) + 4;              // input columns per stage row
  static constexpr int STAGE = BT_ROWS * WIN * BT_CH * 2;    // bytes per stage
};

typedef unsigned long long f32x2;     // two packed floats (lo = even channel)

__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  return ((f32x2)__float_as_uint(hi) << 32) | (f32x2)__float_as_uint(lo);
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
template <bool F16>
__device__ __forceinline__ f32x2 widen2(uint32_t v) {
  const float2 t = unpack_h2<F16>(v);
  return pk2(t.x, t.y);
}
template <bool F16>
__device__ __forceinline__ uint32_t narrow2(f32x2 v) {
  return pack_h2<F16>(__uint_as_float((uint32_t)v), __uint_as_float((uint32_t)(v >> 32)));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// horizontally filtered input row for this thread's PPT outputs x 8 channels: h[o][j] = sum_b w[b] * in[ST*o + b], taps in
// ascending order (the order of blur_hrow in sg2_fp32.cu)
template <bool F16, int ST, int PPT>
__device__ __forceinline__ void hrow(uint32_t row_addr, f32x2 (&h)[PPT][4]) {
  constexpr int NIN = ST * (PPT - 1) + 4;
  const f32x2 w_edge = pk2(0.125f, 0.125f), w_mid = pk2(0.375f, 0.375f);
#pragma unroll
  for (int o = 0; o < PPT; ++o)
#pragma unroll
    for (int j = 0; j < 4; ++j) h[o][j] = 0ULL;
#pragma unroll
  for (int p = 0; p < NIN; ++p) {
    const uint4 raw = lds128(row_addr + p * (BT_CH * 2));
    const f32x2 t[4] = {widen2<F16>(raw.x), widen2<F16>(raw.y), widen2<F16>(raw.z), widen2<F16>(raw.w)};
#pragma unroll
    for (int o = 0; o < PPT; ++o) {
      const int b = p - ST * o;
      if (b < 0 || b > 3) continue;
      const f32x2 w = (b == 0 || b == 3) ? w_edge : w_mid;
#pragma unroll
      for (int j = 0; j < 4; ++j) h[o][j] = fma2(w, t[j], h[o][j]);
    }
  }
}

// vertical pass + store of one output row: 0.125 * (a + d) + 0.375 * (b + c), evaluated as fma(0.125, a + d, 0.375 * (b + c))
template <bool F16, int PPT>
__device__ __forceinline__ void emit(h16* dst, int C, int n_valid, const f32x2 (&a)[PPT][4], const f32x2 (&b)[PPT][4],
                                     const f32x2 (&c)[PPT][4], const f32x2 (&d)[PPT][4]) {
  const f32x2 w_edge = pk2(0.125f, 0.125f), w_mid = pk2(0.375f, 0.375f);
#pragma unroll
  for (int o = 0; o < PPT; ++o) {
    uint4 pk;
    uint32_t* hp = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) hp[j] = narrow2<F16>(fma2(w_edge, add2(a[o][j], d[o][j]), mul2(w_mid, add2(b[o][j], c[o][j]))));
    if (o < n_valid) *reinterpret_cast<uint4*>(dst + (int64_t)o * C) = pk;
  }
}

template <bool F16, int ST, int PPT, int NS>
__global__ void __launch_bounds__(BT_THREADS, PPT == 1 ? 3 : 2)
blur_tma_kernel(const __grid_constant__ CUtensorMap map_in, h16* __restrict__ out, int Ho, int Wo, int C, int pad,
                int c_chunks, int x_strips, int y_segs, int seg_rows) {
  constexpr int XW = BtGeom<ST, PPT>::XW, WIN = BtGeom<ST, PPT>::WIN, STAGE = BtGeom<ST, PPT>::STAGE;
  extern __shared__ unsigned char bt_smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[NS];
  const uint32_t smem0 = (smem_u32(bt_smem_raw) + 127u) & ~127u;
  const int tid = threadIdx.x;
  int b = blockIdx.x;
  const int cc = b % c_chunks; b /= c_chunks;
  const int xs = b % x_strips; b /= x_strips;
  const int ys = b % y_segs;
  const int img = b / y_segs;
  const int y0 = ys * seg_rows, y1 = min(Ho, y0 + seg_rows), rows_out = y1 - y0;
  const int x0 = xs * XW;
  const int col_in0 = x0 * ST - pad, row_in0 = y0 * ST - pad;
  // input rows this segment touches: ST * (rows_out - 1) + 4, in stages of 4
  const int n_stages = (ST * (rows_out - 1) + 4 + BT_ROWS - 1) / BT_ROWS;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    tma_prefetch_desc(&map_in);
    for (int k = 0; k < NS && k < n_stages; ++k) {
      mbar_expect_tx(smem_u32(&full_bar[k]), STAGE);
      tma_load_4d(smem0 + k * STAGE, &map_in, smem_u32(&full_bar[k]), cc * BT_CH, col_in0, row_in0 + k * BT_ROWS, img);
    }
  }
  const int cg = tid & 7, xl = tid >> 3;
  const int x = x0 + xl * PPT;
  const int n_valid = min(PPT, Wo - x);                  // <= 0: a column slot beyond the image (computes, never stores)
  h16* dst = out + (((int64_t)img * Ho + y0) * Wo + x) * C + cc * BT_CH + cg * 8;
  const int64_t dst_stride = (int64_t)Wo * C;
  const uint32_t t_off = (uint32_t)(xl * PPT * ST) * (BT_CH * 2) + cg * 16;
  f32x2 h0[PPT][4], h1[PPT][4], h2[PPT][4], h3[PPT][4];
  int y = y0;                                            // next output row to emit
  for (int k = 0; k < n_stages; ++k) {
    const int s = k % NS;
    mbar_wait(smem_u32(&full_bar[s]), (uint32_t)(k / NS) & 1u);
    const uint32_t base = smem0 + s * STAGE + t_off;
    constexpr uint32_t ROW = WIN * BT_CH * 2;
    if (ST == 1) {
      // stage row r = input row y0 - pad + 4k + r feeds the rotating buffer h_r; output y0 + j - 3 leaves after row j
      hrow<F16, ST, PPT>(base, h0);
      if (k > 0 && y < y1) { emit<F16, PPT>(dst, C, n_valid, h1, h2, h3, h0); dst += dst_stride; ++y; }
      hrow<F16, ST, PPT>(base + ROW, h1);
      if (k > 0 && y < y1) { emit<F16, PPT>(dst, C, n_valid, h2, h3, h0, h1); dst += dst_stride; ++y; }
      hrow<F16, ST, PPT>(base + 2 * ROW, h2);
      if (k > 0 && y < y1) { emit<F16, PPT>(dst, C, n_valid, h3, h0, h1, h2); dst += dst_stride; ++y; }
      hrow<F16, ST, PPT>(base + 3 * ROW, h3);
      if (y < y1) { emit<F16, PPT>(dst, C, n_valid, h0, h1, h2, h3); dst += dst_stride; ++y; }
    } else {
      // output 2k uses this stage's four rows; output 2k - 1 the previous stage's rows 2, 3 (kept in h2, h3) and rows 0, 1
      hrow<F16, ST, PPT>(base, h0);
      hrow<F16, ST, PPT>(base + ROW, h1);
      if (k > 0 && y < y1) { emit<F16, PPT>(dst, C, n_valid, h2, h3, h0, h1); dst += dst_stride; ++y; }
      hrow<F16, ST, PPT>(base + 2 * ROW, h2);
      hrow<F16, ST, PPT>(base + 3 * ROW, h3);
      if (y < y1) { emit<F16, PPT>(dst, C, n_valid, h0, h1, h2, h3); dst += dst_stride; ++y; }
    }
    __syncthreads();                                     // every thread is done reading stage s
    if (tid == 0 && k + NS < n_stages) {
      mbar_expect_tx(smem_u32(&full_bar[s]), STAGE);
      tma_load_4d(smem0 + s * STAGE, &map_in, smem_u32(&full_bar[s]), cc * BT_CH, col_in0, row_in0 + (k + NS) * BT_ROWS, img);
    }
  }
}

template <bool F16, int ST, int PPT, int NS>
int launch(const CUtensorMap& map, h16* out, int64_t n, int Ho, int Wo, int C, int pad, cudaStream_t s) {
  constexpr int SMEM = NS * BtGeom<ST, PPT>::STAGE + 128;
  static std::atomic<unsigned long long> attr_set{0};
  int dev = 0;
  SDG_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((attr_set.load() >> dev) & 1ULL)) {
    SDG_CUDA(cudaFuncSetAttribute((blur_tma_kernel<F16, ST, PPT, NS>), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (dev < 64) attr_set.fetch_or(1ULL << dev);
  }
  const int c_chunks = C / BT_CH;
  const int x_strips = (int)cdiv(Wo, BtGeom<ST, PPT>::XW);
  const int seg_target = ST == 1 ? 64 : 32;
  const int y_segs = (int)cdiv(Ho, seg_target);
  const int seg_rows = (int)cdiv(Ho, y_segs);
  const int64_t blocks = (int64_t)c_chunks * x_strips * y_segs * n;
  SDG_REQUIRE(blocks < (1LL << 31), SDG_E_UNSUPPORTED, "blur_tma: grid too large");
  SDG_LAUNCH((blur_tma_kernel<F16, ST, PPT, NS>), (unsigned)blocks, BT_THREADS, SMEM, s, map, out, Ho, Wo, C, pad, c_chunks,
             x_strips, y_segs, seg_rows);
  return 0;
}

}  // namespace

bool blur_tma_applies(int H, int W, int C, int stride) {
  const int ho = stride == 1 ? H : H / 2;
  return C % BT_CH == 0 && ho >= 32 && W >= 32;
}

// variant: 1 = one output column per thread, deep ring (4 / 3 stages); 2 = two adjacent columns per thread (shared pixel
// conversions), 2 stages; 3 = one column per thread, 2 stages (most resident CTAs); 0 = the default choice
int blur_tma(const h16* in, h16* out, int64_t n, int H, int W, int C, int pad, int stride, int f16, int variant, cudaStream_t s) {
  const int Ho = (H + 2 * pad - 4) / stride + 1, Wo = (W + 2 * pad - 4) / stride + 1;
  if (n == 0 || Ho <= 0 || Wo <= 0) return 0;
  SDG_REQUIRE(C % BT_CH == 0, SDG_E_UNSUPPORTED, "blur_tma: C=%d is not a multiple of %d", C, BT_CH);
  SDG_REQUIRE(stride == 1 || stride == 2, SDG_E_UNSUPPORTED, "blur_tma: stride=%d", stride);
  SDG_REQUIRE(variant >= 0 && variant <= 3, SDG_E_INVALID, "blur_tma: variant=%d", variant);
  if (variant == 0) variant = 1;
  const int ppt = variant == 2 ? 2 : 1;
  CUtensorMap map;
  const int win = stride * (32 * ppt - 1) + 4;
  { int rc = tc_encode_act_box(&map, in, f16, n, H, W, C, BT_CH, win, BT_ROWS); if (rc) return rc; }
#define SDG_BT(F, ST, PPT, NS) return launch<F, ST, PPT, NS>(map, out, n, Ho, Wo, C, pad, s)
#define SDG_BT_V(F, ST, NS1)                                \
  do {                                                      \
    if (variant == 1) SDG_BT(F, ST, 1, NS1);                \
    else if (variant == 2) SDG_BT(F, ST, 2, 2);             \
    else SDG_BT(F, ST, 1, 2);                               \
  } while (0)
  if (f16) { if (stride == 1) SDG_BT_V(true, 1, 4); else SDG_BT_V(true, 2, 3); }
  else { if (stride == 1) SDG_BT_V(false, 1, 4); else SDG_BT_V(false, 2, 3); }
#undef SDG_BT_V
#undef SDG_BT
}

}  // namespace sdg
