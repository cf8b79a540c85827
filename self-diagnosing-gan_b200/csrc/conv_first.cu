// conv_first.cu -- first convolution of the SNGAN discriminators straight from the dataset bytes.
//
// Replaces transform.py:3-11 (ToTensor + Normalize) + DBlockOptimized.c1 (SNConv2d 3->C, 3x3, pad 1) + ReLU
// (torch-mimicry resblocks.py; SURVEY 8(a) "b1.c1"): out = relu(conv3x3(norm(x)) + b), written as the 16-bit
// NHWC operand of the next conv.  K = 27 is too thin to stream through TMA, so builder warps assemble the
// im2col tile (128 pixels x 32, 128B-swizzled K-major) in shared memory:
//   the raw rows of the tile (+1 halo row each side) are prefetched two tiles ahead with cp.async; each builder thread
//   gathers its pixel's 27 values, uint8 through a 256-entry lookup table holding (v/255 - 0.5)/0.5 rounded to the
//   operand type (27 independent load chains).
// One thread issues two tcgen05.mma (K = 2 x 16) per tile.  The bias rides in the GEMM: operand columns 27 and 28 of
// every pixel are 1.0 and the matching weight columns hold bias_hi and bias_lo (bias = hi + lo to ~22 bits), so the
// epilogue is only TMEM load -> convert-with-ReLU (cvt.rn.relu) -> swizzled shared memory -> TMA store, several tiles
// in flight.  Profiled, the kernel is latency-bound (warps issue 17 % of the time), so TWO CTAs share each SM
// (<= 112 registers, ~94 KB of shared memory, 2 x 256 TMEM columns).  3 KB in, 256 KB out per CIFAR-shaped sample.
// Warp roles: 0-3 epilogue (TMEM lane quadrant = warp), 4-7 builders, 8 weights TMA / MMA / TMEM.
#include "tc_ptx.cuh"

namespace sdg {

constexpr int FC_THREADS = 9 * 32;
constexpr int FC_A_BYTES = 128 * 128;
constexpr int FC_OUT_BUFS = 1;       // output staging tiles per CTA (two CTAs share an SM and overlap each other)

struct FcParams {
  const void* x;
  int layout, S, Cout;
  long long n_images, tiles;
  const float* bias;
  h16* out;
  int* ovf;                  // fp16 range guard flag or null
};

// SP = super-pixel form for Cout = 64 (BN = 128): a GEMM row is two horizontally adjacent pixels, its 128 columns are
// 64 channels of the even + 64 of the odd pixel; the window is 3 x 4 taps at x-stride 2 (K = 36 + 2 bias columns, 3 MMAs),
// the weights hold each 3x3 filter twice, shifted by one column.  Same bytes out per tile as the Cout = 128 case, half the
// tiles, 36 instead of 54 gathered values per pixel pair (the builders, not HBM, bound the plain Cout = 64 kernel).
template <int BN, bool F16, bool SP>
__global__ void __launch_bounds__(FC_THREADS, 2)
first_conv_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out, const FcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_addr = smem_base + 2 * FC_A_BYTES;
  // output staging: FC_OUT_BUFS buffers x (BN/64) boxes of 128 rows x 128 B, 128B-swizzled, drained by TMA stores
  constexpr int OUT_BOX = 128 * 128;
  constexpr int OUT_BYTES = (BN / 64) * OUT_BOX;
  const uint32_t out_off = 2 * FC_A_BYTES + BN * 128;

  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2], b_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_rawf[3][1152];     // raw rows: (R+2) x S x 3 bytes (u8) or floats (fp32 NCHW)
  __shared__ uint16_t s_lut[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.S;
  const int SX = SP ? S / 2 : S;          // GEMM pixels per image row
  const int R = 128 / SX;                 // image rows per tile
  const int tiles_y = S / R;

  if (threadIdx.x < 256) {
    float v = __fdiv_rn((float)threadIdx.x, 255.0f);
    v = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
    s_lut[threadIdx.x] = (uint16_t)(pack_h2<F16>(v, 0.f) & 0xffffu);
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&a_full[s]), 128);
        mbar_init(smem_u32(&a_empty[s]), 1);
        mbar_init(smem_u32(&acc_full[s]), 1);
        mbar_init(smem_u32(&acc_empty[s]), 4);
      }
      mbar_init(smem_u32(&b_full), 1);
      fence_barrier_init();
      tma_prefetch_desc(&map_b);
      tma_prefetch_desc(&map_out);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), 2 * BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 8) {
    // ================= weights (once) + MMA issuer =================
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&b_full), BN * 128);
      tma_load_2d(b_addr, &map_b, smem_u32(&b_full), 0, 0);
    }
    __syncwarp();
    mbar_wait(smem_u32(&b_full), 0);
    // bias as two extra K columns (k = 27: hi, k = 28: lo) of the resident weight tile; A carries 1.0 there.
    // row o of the swizzled tile: 16-byte chunk 3 (k = 24..31) sits at chunk position 3 ^ (o & 7)
    for (int o = lane; o < BN; o += 32) {
      const float b = p.bias[SP ? (o & 63) : o];
      const uint32_t hi = pack_h2<F16>(b, 0.f) & 0xffffu;
      const float bhi = unpack_h2<F16>(hi).x;
      const uint32_t lo = pack_h2<F16>(b - bhi, 0.f) & 0xffffu;
      // plain: k = 27, 28 (chunk 3, elements 3, 4); super-pixel: k = 36, 37 (chunk 4, elements 4, 5)
      uint16_t* chunk = reinterpret_cast<uint16_t*>(smem_gen + 2 * FC_A_BYTES + o * 128 + (((SP ? 4 : 3) ^ (o & 7)) << 4));
      chunk[SP ? 4 : 3] = (uint16_t)hi;
      chunk[SP ? 5 : 4] = (uint16_t)lo;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, BN, F16);
      const uint64_t bdesc = make_sw128_desc(b_addr);
      long long local = 0;
      for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
        const int buf = (int)(local & 1);
        const uint32_t ph = (uint32_t)((local >> 1) & 1);
        mbar_wait(smem_u32(&acc_empty[buf]), ph ^ 1u);
        mbar_wait(smem_u32(&a_full[buf]), ph);
        tc_fence_after();
        const uint64_t adesc = make_sw128_desc(smem_base + buf * FC_A_BYTES);
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        umma_bf16(d_tmem, adesc, bdesc, idesc, 0u);                 // k = 0..15
        umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);         // k = 16..31 (27, 28: the bias columns; 29..31 zero)
        if (SP) umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u); // k = 32..47 (36, 37: the bias columns)
        umma_commit(smem_u32(&a_empty[buf]));
        umma_commit(smem_u32(&acc_full[buf]));
      }
    }
  } else if (warp >= 4) {
    // ================= builders: raw rows -> normalised patch -> swizzled im2col tile =================
    const int bt = threadIdx.x - 128;           // 0..127 = pixel of the tile
    const int ly = bt / SX, lx = bt - ly * SX;      // row of the tile, GEMM pixel within the row
    const bool u8 = p.layout == SDG_LAYOUT_U8_NHWC;
    auto prefetch_raw = [&](long long tile, int rb) {
      if (tile < p.tiles) {
        const long long n = tile / tiles_y;
        const int y0 = (int)(tile % tiles_y) * R;
        float* rawf = s_rawf[rb];
        if (u8) {
          const int words_per_row = S * 3 / 4;
          const int words = (R + 2) * words_per_row;
          for (int w = bt; w < words; w += 128) {
            const int pr = w / words_per_row, wi = w - pr * words_per_row;
            const int iy = y0 - 1 + pr;
            if (iy >= 0 && iy < S)
              cp_async_4(smem_u32(reinterpret_cast<uint32_t*>(rawf) + w),
                         reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(p.x) + ((n * S + iy) * (long long)S) * 3) + wi);
          }
        } else {
          const int per_c = (R + 2) * S;
          for (int w = bt; w < 3 * per_c; w += 128) {
            const int c = w / per_c, r = w - c * per_c;
            const int pr = r / S, ix = r - pr * S;
            const int iy = y0 - 1 + pr;
            if (iy >= 0 && iy < S)
              cp_async_4(smem_u32(rawf + w), reinterpret_cast<const float*>(p.x) + ((n * 3 + c) * S + iy) * (long long)S + ix);
          }
        }
      }
      cp_async_commit();          // one group per call, empty or not: keeps the wait_group arithmetic uniform
    };
    prefetch_raw(blockIdx.x, 0);
    prefetch_raw((long long)blockIdx.x + gridDim.x, 1);
    long long local = 0;
    for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
      const int buf = (int)(local & 1);
      const uint32_t ph = (uint32_t)((local >> 1) & 1);
      const int y0 = (int)(tile % tiles_y) * R;
      const float* rawf = s_rawf[local % 3];
      const uint8_t* rawb = reinterpret_cast<const uint8_t*>(rawf);
      cp_async_wait_1();          // everything but the newest group has landed -> this tile's rows are in smem
      named_bar_sync(1, 128);     // ... for every builder thread
      // the raw buffer of tile local-1 is free now: prefetch two tiles ahead
      prefetch_raw(tile + 2 * (long long)gridDim.x, (int)((local + 2) % 3));
      // ---- gather this pixel's 3x3x3 neighbourhood: k = (ky*3+kx)*3 + c (27 independent byte -> LUT chains) ----
      constexpr uint32_t kOne = F16 ? 0x3C00u : 0x3F80u;      // 1.0 in the operand type: the bias columns
      constexpr int TX = SP ? 4 : 3;                           // window columns
      constexpr int NV = 3 * TX * 3;                           // 27 or 36 gathered values
      constexpr int NW = SP ? 24 : 16;                         // packed 32-bit words written per row (K = 48 or 32)
      uint16_t vals[NW * 2];
#pragma unroll
      for (int tap = 0; tap < 3 * TX; ++tap) {
        const int ky = tap / TX, kx = tap % TX;
        const int iy = y0 + ly + ky - 1, ix = (SP ? 2 * lx : lx) + kx - 1;
        const bool ok = iy >= 0 && iy < S && ix >= 0 && ix < S;
        const int pr = ly + ky;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint16_t h = 0;
          if (ok) {
            if (u8) h = s_lut[rawb[(pr * S + ix) * 3 + c]];
            else h = (uint16_t)(pack_h2<F16>(rawf[(c * (R + 2) + pr) * S + ix], 0.f) & 0xffffu);
          }
          vals[tap * 3 + c] = h;
        }
      }
      vals[NV] = (uint16_t)kOne; vals[NV + 1] = (uint16_t)kOne;   // bias hi / lo columns
#pragma unroll
      for (int j = NV + 2; j < NW * 2; ++j) vals[j] = 0;
      uint32_t packed[NW];
#pragma unroll
      for (int j = 0; j < NW; ++j) packed[j] = (uint32_t)vals[2 * j] | ((uint32_t)vals[2 * j + 1] << 16);
      mbar_wait(smem_u32(&a_empty[buf]), ph ^ 1u);          // the MMAs that read this A buffer have retired
      uint8_t* row = smem_gen + buf * FC_A_BYTES + bt * 128;
#pragma unroll
      for (int ch = 0; ch < NW / 4; ++ch)
        *reinterpret_cast<uint4*>(row + ((ch ^ (bt & 7)) << 4)) =
            make_uint4(packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
      fence_proxy_async_smem();                             // generic-proxy writes -> visible to the tensor core
      mbar_arrive(smem_u32(&a_full[buf]));
    }
  } else {
    // ================= epilogue: TMEM -> ReLU + 16-bit convert -> swizzled smem -> TMA store =================
    const int q = warp;
    const int row = q * 32 + lane;
    long long local = 0;
    uint32_t nonfinite = 0;           // fp16 range guard: OR of the inf / NaN marks of every packed output word
    for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
      const int buf = (int)(local & 1);
      const uint32_t ph = (uint32_t)((local >> 1) & 1);
      mbar_wait(smem_u32(&acc_full[buf]), ph);
      tc_fence_after();
      // the store that read this staging buffer FC_OUT_BUFS tiles ago must have finished reading it
      const int ob = (int)(local % FC_OUT_BUFS);
      if (threadIdx.x == 0) bulk_wait_read<FC_OUT_BUFS - 1>();
      named_bar_sync(2, 128);
      uint8_t* stage = smem_gen + out_off + ob * OUT_BYTES;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), r);
        tmem_ld_wait();
        uint8_t* srow = stage + (c0 >> 6) * OUT_BOX + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_relu_h2<F16>(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
          pk.y = pack_relu_h2<F16>(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
          pk.z = pack_relu_h2<F16>(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
          pk.w = pack_relu_h2<F16>(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
          if (F16) nonfinite |= f16x2_nonfinite_bits(pk.x) | f16x2_nonfinite_bits(pk.y) | f16x2_nonfinite_bits(pk.z) |
                                f16x2_nonfinite_bits(pk.w);
          const int chunk = ((c0 & 63) >> 3) + g;                 // 16-byte chunk within the 128-byte row
          *reinterpret_cast<uint4*>(srow + ((chunk ^ (row & 7)) << 4)) = pk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));       // TMEM stage drained
      fence_proxy_async_smem();                                    // staging writes -> visible to the TMA engine
      named_bar_sync(2, 128);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int bx = 0; bx < BN / 64; ++bx)
          tma_store_2d(&map_out, smem_base + out_off + ob * OUT_BYTES + bx * OUT_BOX, bx * 64, (int)(tile * 128));
        bulk_commit();
      }
    }
    if (threadIdx.x == 0) bulk_wait_all();
    if (F16 && nonfinite) range_flag_set(p.ovf, SDG_RANGE_ACT);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN>
constexpr int fc_smem_bytes() { return 2 * FC_A_BYTES + BN * 128 + FC_OUT_BUFS * (BN / 64) * 128 * 128 + 1024; }

int first_conv_init() {
  SDG_CUDA(cudaFuncSetAttribute((first_conv_kernel<128, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, fc_smem_bytes<128>()));
  SDG_CUDA(cudaFuncSetAttribute((first_conv_kernel<128, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, fc_smem_bytes<128>()));
  SDG_CUDA(cudaFuncSetAttribute((first_conv_kernel<128, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, fc_smem_bytes<128>()));
  SDG_CUDA(cudaFuncSetAttribute((first_conv_kernel<128, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, fc_smem_bytes<128>()));
  SDG_CUDA(cudaFuncSetAttribute((first_conv_kernel<64, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, fc_smem_bytes<64>()));
  SDG_CUDA(cudaFuncSetAttribute((first_conv_kernel<64, false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, fc_smem_bytes<64>()));
  return 0;
}

int first_conv(const void* x, int layout, const h16* wb, const float* bias, h16* out, int64_t n, int S, int Cout,
               int f16, cudaStream_t s, int superpix) {
  SDG_REQUIRE(S == 32 || S == 64, SDG_E_UNSUPPORTED, "first_conv: image size %d", S);
  SDG_REQUIRE(Cout == 64 || Cout == 128, SDG_E_UNSUPPORTED, "first_conv: Cout=%d", Cout);
  SDG_REQUIRE(!superpix || Cout == 64, SDG_E_INVALID, "first_conv: the super-pixel form is for Cout = 64");
  SDG_REQUIRE(bias, SDG_E_INVALID, "first_conv: bias required");
  SDG_REQUIRE(((uintptr_t)x % 4) == 0 && ((uintptr_t)out % 16) == 0, SDG_E_INVALID, "first_conv: misaligned pointer");
  if (n == 0) return 0;
  // super-pixel form: wb is [128][64] (pack_first_superpix_h16), the output [n*S*S/2 rows][128] is the same memory as
  // [n*S*S rows][64]
  const int bn = superpix ? 128 : Cout;
  const uint64_t rows = (uint64_t)n * S * S / (superpix ? 2 : 1);
  CUtensorMap map_b, map_out;
  { int rc = tc_encode_2d(&map_b, wb, f16, 64, bn, 64, bn); if (rc) return rc; }
  { int rc = tc_encode_2d(&map_out, out, f16, bn, rows, 64, 128); if (rc) return rc; }
  FcParams p;
  p.x = x; p.layout = layout; p.S = S; p.Cout = Cout; p.n_images = n;
  p.tiles = (long long)(rows / 128);
  p.bias = bias; p.out = out; p.ovf = t_range_flag;
  const int sms = tc_num_sms();
  const int grid = (int)(p.tiles < 2 * sms ? p.tiles : 2 * sms);     // two persistent CTAs per SM
  if (superpix) {
    if (f16) { SDG_LAUNCH((first_conv_kernel<128, true, true>), grid, FC_THREADS, fc_smem_bytes<128>(), s, map_b, map_out, p); }
    else { SDG_LAUNCH((first_conv_kernel<128, false, true>), grid, FC_THREADS, fc_smem_bytes<128>(), s, map_b, map_out, p); }
  } else if (Cout == 128 && f16) {
    SDG_LAUNCH((first_conv_kernel<128, true, false>), grid, FC_THREADS, fc_smem_bytes<128>(), s, map_b, map_out, p);
  } else if (Cout == 128) {
    SDG_LAUNCH((first_conv_kernel<128, false, false>), grid, FC_THREADS, fc_smem_bytes<128>(), s, map_b, map_out, p);
  } else if (f16) {
    SDG_LAUNCH((first_conv_kernel<64, true, false>), grid, FC_THREADS, fc_smem_bytes<64>(), s, map_b, map_out, p);
  } else {
    SDG_LAUNCH((first_conv_kernel<64, false, false>), grid, FC_THREADS, fc_smem_bytes<64>(), s, map_b, map_out, p);
  }
  return 0;
}

}  // namespace sdg
